"""ctypes mirror of the C ABI (include/b2_batch.h, include/mujoco/mujoco.h).  No physics here: every number comes from
libb2sim.so.  Importing fails loudly when the library has not been built (python mujoco_sim_b200/build.py)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    return os.path.join(_HERE, "lib", "libb2sim.so")


def asset(name):
    """Path of a model file shipped under mujoco_sim_b200/assets."""
    return os.path.join(_HERE, "assets", name)


class B2Error(RuntimeError):
    pass


if not os.path.exists(lib_path()):
    raise ImportError("%s is missing: build it with `python mujoco_sim_b200/build.py` "
                      "(there is no Python or CPU fallback for the CUDA engine)" % lib_path())
lib = C.CDLL(lib_path(), mode=C.RTLD_GLOBAL)

TICK_CONTROLLER, TICK_INVERSE, TICK_INTEGRATE, TICK_ODOM, TICK_NOSOLVE, TICK_READ_POST = 1, 2, 4, 8, 1 << 9, 1 << 10
F32, F64, EXPORT_STAGES = 4, 8, 0x100
ENV_MAJOR, NATIVE = 0, 1
OBJ_BODY, OBJ_JOINT, OBJ_GEOM, OBJ_MESH = 1, 3, 5, 9

_vp, _cp, _i = C.c_void_p, C.c_char_p, C.c_int
_sig = {
    "mj_loadXML": (_vp, [_cp, _vp, _cp, _i]),
    "mj_loadXMLString": (_vp, [_cp, _cp, _cp, _i]),
    "mj_saveLastXML": (_i, [_cp, _vp, _cp, _i]),
    "mj_saveModel": (None, [_vp, _cp, _vp, _i]),
    "mj_loadModel": (_vp, [_cp, _vp]),
    "mj_makeData": (_vp, [_vp]),
    "mj_deleteData": (None, [_vp]),
    "mj_deleteModel": (None, [_vp]),
    "mj_resetData": (None, [_vp, _vp]),
    "mj_step": (None, [_vp, _vp]),
    "mj_step1": (None, [_vp, _vp]),
    "mj_step2": (None, [_vp, _vp]),
    "mj_forward": (None, [_vp, _vp]),
    "mj_inverse": (None, [_vp, _vp]),
    "mj_mulM": (None, [_vp, _vp, _vp, _vp]),
    "mj_name2id": (_i, [_vp, _i, _cp]),
    "mj_id2name": (_cp, [_vp, _i, _i]),
    "mj_printModel": (None, [_vp, _cp]),
    "mj_printData": (None, [_vp, _vp, _cp]),
    "b2_model_int": (_i, [_vp, _cp, C.POINTER(_i)]),
    "b2_model_array": (_i, [_vp, _cp, C.POINTER(_vp), C.POINTER(_i)]),
    "b2_data_array": (_i, [_vp, _vp, _cp, C.POINTER(_vp), C.POINTER(_i)]),
    "b2_model_set_opt": (_i, [_vp, _cp, C.c_double]),
    "b2_data_contacts": (_i, [_vp, _i, _vp, _vp, _vp]),
    "b2_last_error": (_cp, []),
    "b2_device_count": (_i, []),
    "b2_create": (_vp, [_vp, _i, _i, _i]),
    "b2_destroy": (None, [_vp]),
    "b2_path_name": (_cp, [_vp]),
    "b2_nenv": (_i, [_vp]),
    "b2_nenv_padded": (_i, [_vp]),
    "b2_precision": (_i, [_vp]),
    "b2_set_controlled": (_i, [_vp, _vp]),
    "b2_set_odom": (_i, [_vp, _i, _vp, _vp]),
    "b2_set_timestep": (_i, [_vp, C.c_double]),
    "b2_set_option": (_i, [_vp, _cp, C.c_double]),
    "b2_field_size": (_i, [_vp, _cp]),
    "b2_set_field_f32": (_i, [_vp, _cp, _vp, _i, _i, _i]),
    "b2_set_field_f64": (_i, [_vp, _cp, _vp, _i, _i, _i]),
    "b2_get_field_f32": (_i, [_vp, _cp, _vp, _i, _i, _i]),
    "b2_get_field_f64": (_i, [_vp, _cp, _vp, _i, _i, _i]),
    "b2_get_field_i32": (_i, [_vp, _cp, _vp, _i, _i, _i]),
    "b2_device_ptr": (_vp, [_vp, _cp]),
    "b2_reset": (_i, [_vp, _i, _i]),
    "b2_tick": (_i, [_vp, _i]),
    "b2_step": (_i, [_vp, _i]),
    "b2_set_tick_flags": (_i, [_vp, _i]),
    "b2_forward": (_i, [_vp]),
    "b2_sync": (_i, [_vp]),
    "b2_stream": (_vp, [_vp]),
    "b2_launch_count": (C.c_longlong, [_vp]),
    "b2_set_hw_joints": (_i, [_vp, _i, _vp]),
    "b2_write_commands": (_i, [_vp, _vp, _vp]),
    "b2_read_joints": (_i, [_vp, _vp, _vp, _vp]),
    "b2_tick_host": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "b2_tick_resident": (_i, [_vp]),
    "b2_register_host": (_i, [_vp, _vp, C.c_longlong]),
    "b2_unregister_host": (_i, [_vp, _vp]),
    "b2_pack_obs": (_i, [_vp, _vp]),
    "b2_transfer_state": (_i, [_vp, _vp]),
    "b2_obs_create": (_vp, [_vp, _i, _i]),
    "b2_obs_handle": (_i, [_vp, _vp]),
    "b2_obs_attach": (_i, [_vp, _vp, _vp]),
    "b2_obs_enable": (_i, [_vp, _i]),
    "b2_obs_read": (_i, [_vp, _vp]),
    "b2_create_multi": (_vp, [_vp, _i, _vp, _i, _i, _i]),
    "b2_multi_destroy": (None, [_vp]),
    "b2_multi_count": (_i, [_vp]),
    "b2_multi_shard": (_vp, [_vp, _i]),
    "b2_multi_tick": (_i, [_vp, _i]),
    "b2_multi_sync": (_i, [_vp]),
    "b2_set_pd": (_i, [_vp, _vp, _vp]),
    "b2_set_slots": (_i, [_vp, _i, _vp]),
    "b2_spawn": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "b2_destroy_slots": (_i, [_vp, _i, _vp, _vp]),
    "b2_slot_active": (_i, [_vp, _vp, _i, _i]),
    "b2_l2_flush": (_i, [_vp, C.c_longlong]),
    "b2_profile_begin": (_i, [_vp, _i]),
    "b2_profile_end": (_i, [_vp, _vp, _i]),
    "b2_mirror_env": (_i, [_vp, _i, _vp]),
    "b2_load_env": (_i, [_vp, _i, _vp]),
    "b2_shim_batch": (_vp, [_vp, _vp]),
}
for _n, (_r, _a) in _sig.items():
    _f = getattr(lib, _n)
    _f.restype = _r
    _f.argtypes = _a

INT_FIELDS = {"ncon", "nefc", "efc_type", "efc_id", "contact_int", "solver_iter", "status", "efc_nwords", "env_order", "nisl", "isl_off", "isl_end", "nblk", "blk_row0", "blk_off", "efc_tree"}


def _err():
    return (lib.b2_last_error() or b"").decode()


_KIND_DTYPE = {0: np.float64, 1: np.int32, 2: np.uint8, 3: np.float32}


class Model:
    """mjModel handle (mj_loadXML: reference include/mujoco_sim/mj_util.h:190)."""

    def __init__(self, path=None, xml=None, basedir="."):
        err = C.create_string_buffer(1000)
        if path is not None and path.endswith(".mjb"):     # binary image of a compiled model (mj_loadModel)
            self.ptr = lib.mj_loadModel(path.encode(), None)
            if not self.ptr:
                raise B2Error("mj_loadModel: cannot load " + path)
        elif path is not None:
            self.ptr = lib.mj_loadXML(path.encode(), None, err, 1000)
        else:
            self.ptr = lib.mj_loadXMLString(xml.encode(), basedir.encode(), err, 1000)
        if not self.ptr:
            raise B2Error("mj_loadXML: " + err.value.decode())

    def __del__(self):
        if getattr(self, "ptr", None):
            lib.mj_deleteModel(self.ptr)
            self.ptr = None

    def save(self, path):
        """mj_saveModel: binary image of the compiled model."""
        lib.mj_saveModel(self.ptr, path.encode(), None, 0)
        if not os.path.exists(path):
            raise B2Error("mj_saveModel: cannot write " + path)

    def int(self, name):
        v = _i(0)
        if lib.b2_model_int(self.ptr, name.encode(), C.byref(v)) < 0:
            raise KeyError(name)
        return v.value

    def array(self, name):
        """numpy VIEW of a model array (writes go through to the model)."""
        p, k = _vp(), _i()
        n = lib.b2_model_array(self.ptr, name.encode(), C.byref(p), C.byref(k))
        if n < 0:
            raise KeyError(name)
        if n == 0 or not p.value:
            return np.zeros(0, _KIND_DTYPE[k.value])
        dt = _KIND_DTYPE[k.value]
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,))

    def __getattr__(self, name):
        if name.startswith("n") and not name.startswith("name"):
            try:
                return self.int(name)
            except KeyError:
                pass
        try:
            return self.array(name)
        except KeyError:
            raise AttributeError(name)

    def set_opt(self, name, value):
        if lib.b2_model_set_opt(self.ptr, name.encode(), float(value)) < 0:
            raise KeyError(name)

    @property
    def timestep(self):
        return float(self.array("opt.timestep")[0])

    def name2id(self, objtype, name):
        return lib.mj_name2id(self.ptr, objtype, name.encode())

    def id2name(self, objtype, i):
        s = lib.mj_id2name(self.ptr, objtype, i)
        return s.decode() if s else None


class Data:
    """mjData handle (mj_makeData: reference src/mujoco_sim/mj_sim.cpp:816)."""

    def __init__(self, model):
        self.model = model
        self.ptr = lib.mj_makeData(model.ptr)
        if not self.ptr:
            raise B2Error("mj_makeData failed")

    def __del__(self):
        if getattr(self, "ptr", None):
            lib.mj_deleteData(self.ptr)
            self.ptr = None

    def array(self, name):
        p, k = _vp(), _i()
        n = lib.b2_data_array(self.model.ptr, self.ptr, name.encode(), C.byref(p), C.byref(k))
        if n < 0:
            raise KeyError(name)
        dt = _KIND_DTYPE[k.value]
        if n == 0 or not p.value:
            return np.zeros(0, dt)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,))

    def __getattr__(self, name):
        if name in ("model", "ptr"):
            raise AttributeError(name)
        try:
            a = self.array(name)
        except KeyError:
            raise AttributeError(name)
        if name in ("time", "ncon", "nefc", "solver_iter"):
            return a[0]
        return a


def data_contacts(data, nmax=None):
    """(geom1, geom2, dist) arrays of the contact list held by an mjData (the legacy single-environment view)."""
    n = int(data.ncon) if nmax is None else int(nmax)
    g1, g2, dist = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float64)
    if n:
        lib.b2_data_contacts(data.ptr, n, g1.ctypes.data, g2.ctypes.data, dist.ctypes.data)
    return g1, g2, dist


class Batch:
    """b2_batch handle: nenv environments of one model on one GPU."""

    def __init__(self, model, nenv, device=0, precision=F32, export_stages=False):
        self.model = model
        self.ptr = lib.b2_create(model.ptr, int(nenv), int(device), int(precision) | (EXPORT_STAGES if export_stages else 0))
        if not self.ptr:
            raise B2Error("b2_create: " + _err())
        self.nenv = int(nenv)
        self.precision = precision

    def close(self):
        if getattr(self, "ptr", None):
            lib.b2_destroy(self.ptr)
            self.ptr = None

    __del__ = close

    def _ck(self, rc, what):
        if rc < 0:
            raise B2Error(what + ": " + _err())
        return rc

    def field_size(self, name):
        return self._ck(lib.b2_field_size(self.ptr, name.encode()), "b2_field_size")

    def set(self, name, value, env_lo=0, env_hi=None, layout=ENV_MAJOR):
        """value: array [env][n] (ENV_MAJOR) or [n][env] (NATIVE) for environments [env_lo, env_hi)."""
        env_hi = self.nenv if env_hi is None else env_hi
        n = self.field_size(name)
        nenv = env_hi - env_lo
        shape = (nenv, n) if layout == ENV_MAJOR else (n, nenv)
        if np.asarray(value).dtype == np.float32:
            a = np.ascontiguousarray(np.broadcast_to(np.asarray(value, np.float32), shape))
            rc = lib.b2_set_field_f32(self.ptr, name.encode(), a.ctypes.data, env_lo, env_hi, layout)
        else:
            a = np.ascontiguousarray(np.broadcast_to(np.asarray(value, np.float64), shape))
            rc = lib.b2_set_field_f64(self.ptr, name.encode(), a.ctypes.data, env_lo, env_hi, layout)
        self._ck(rc, "b2_set_field(%s)" % name)

    def get(self, name, env_lo=0, env_hi=None, layout=ENV_MAJOR, dtype=np.float64):
        env_hi = self.nenv if env_hi is None else env_hi
        n = self.field_size(name)
        nenv = env_hi - env_lo
        shape = (nenv, n) if layout == ENV_MAJOR else (n, nenv)
        if name in INT_FIELDS:
            out = np.empty(shape, np.int32)
            rc = lib.b2_get_field_i32(self.ptr, name.encode(), out.ctypes.data, env_lo, env_hi, layout)
        elif dtype == np.float32:
            out = np.empty(shape, np.float32)
            rc = lib.b2_get_field_f32(self.ptr, name.encode(), out.ctypes.data, env_lo, env_hi, layout)
        else:
            out = np.empty(shape, np.float64)
            rc = lib.b2_get_field_f64(self.ptr, name.encode(), out.ctypes.data, env_lo, env_hi, layout)
        self._ck(rc, "b2_get_field(%s)" % name)
        return out

    def device_ptr(self, name):
        p = lib.b2_device_ptr(self.ptr, name.encode())
        if not p:
            raise B2Error("b2_device_ptr: " + _err())
        return p

    def set_controlled(self, mask):
        a = np.ascontiguousarray(mask, np.uint8)
        self._ck(lib.b2_set_controlled(self.ptr, a.ctypes.data), "b2_set_controlled")

    def set_odom(self, dof, qposadr):
        d = np.ascontiguousarray(dof, np.int32).reshape(-1, 6)
        q = np.ascontiguousarray(qposadr, np.int32).reshape(-1, 3)
        self._ck(lib.b2_set_odom(self.ptr, d.shape[0], d.ctypes.data, q.ctypes.data), "b2_set_odom")

    def set_timestep(self, h):
        self._ck(lib.b2_set_timestep(self.ptr, float(h)), "b2_set_timestep")

    def set_option(self, name, value):
        self._ck(lib.b2_set_option(self.ptr, name.encode(), float(value)), "b2_set_option")

    def reset(self, env_lo=0, env_hi=None):
        self._ck(lib.b2_reset(self.ptr, env_lo, self.nenv if env_hi is None else env_hi), "b2_reset")

    def tick(self, flags=TICK_INTEGRATE):
        self._ck(lib.b2_tick(self.ptr, flags), "b2_tick")

    def step(self, nsteps=1):
        self._ck(lib.b2_step(self.ptr, nsteps), "b2_step")

    def set_tick_flags(self, flags):
        self._ck(lib.b2_set_tick_flags(self.ptr, flags), "b2_set_tick_flags")

    def forward(self):
        self._ck(lib.b2_forward(self.ptr), "b2_forward")

    def sync(self):
        self._ck(lib.b2_sync(self.ptr), "b2_sync")

    @property
    def path_name(self):
        return (lib.b2_path_name(self.ptr) or b"").decode()

    @property
    def stream(self):
        return lib.b2_stream(self.ptr)

    @property
    def launch_count(self):
        return lib.b2_launch_count(self.ptr)

    def set_hw_joints(self, jnt_ids):
        a = np.ascontiguousarray(jnt_ids, np.int32)
        self._ck(lib.b2_set_hw_joints(self.ptr, a.size, a.ctypes.data), "b2_set_hw_joints")
        self.nhw = a.size

    def write_commands(self, vel_cmd, effort_cmd):
        v = np.ascontiguousarray(vel_cmd, np.float32)
        e = np.ascontiguousarray(effort_cmd, np.float32)
        self._ck(lib.b2_write_commands(self.ptr, v.ctypes.data, e.ctypes.data), "b2_write_commands")

    def read_joints(self):
        out = [np.empty((self.nhw, self.nenv), np.float32) for _ in range(3)]
        self._ck(lib.b2_read_joints(self.ptr, *[o.ctypes.data for o in out]), "b2_read_joints")
        return out

    def tick_host_raw(self, vel_ptr, eff_ptr, pos_ptr, velo_ptr, effo_ptr):
        self._ck(lib.b2_tick_host(self.ptr, vel_ptr, eff_ptr, pos_ptr, velo_ptr, effo_ptr), "b2_tick_host")

    def register_host(self, arr):
        """Pin a caller-owned numpy buffer for the zero-copy exchange of tick_host (b2_register_host)."""
        self._ck(lib.b2_register_host(self.ptr, arr.ctypes.data, arr.nbytes), "b2_register_host")

    def unregister_host(self, arr):
        self._ck(lib.b2_unregister_host(self.ptr, arr.ctypes.data), "b2_unregister_host")

    def set_pd(self, kp=None, kd=None):
        """Device-side PD stage: the effort-command buffer then carries position targets (b2_set_pd)."""
        if kp is None:
            self._ck(lib.b2_set_pd(self.ptr, None, None), "b2_set_pd")
            return
        kp = np.ascontiguousarray(np.broadcast_to(np.asarray(kp, np.float32), (self.nhw,)))
        kd = np.ascontiguousarray(np.broadcast_to(np.asarray(kd, np.float32), (self.nhw,)))
        self._ck(lib.b2_set_pd(self.ptr, kp.ctypes.data, kd.ctypes.data), "b2_set_pd")

    # ---- runtime spawn / destroy as slot activation (b2_set_slots) ----
    def set_slots(self, body_ids):
        a = np.ascontiguousarray(body_ids, np.int32)
        self._ck(lib.b2_set_slots(self.ptr, a.size, a.ctypes.data), "b2_set_slots")
        self.nslot = a.size

    def spawn(self, env, slot, pose7, twist6=None):
        e = np.ascontiguousarray(env, np.int32); s = np.ascontiguousarray(slot, np.int32)
        p = np.ascontiguousarray(np.broadcast_to(np.asarray(pose7, np.float32), (e.size, 7)))
        t = None if twist6 is None else np.ascontiguousarray(np.broadcast_to(np.asarray(twist6, np.float32), (e.size, 6)))
        self._ck(lib.b2_spawn(self.ptr, e.size, e.ctypes.data, s.ctypes.data, p.ctypes.data, None if t is None else t.ctypes.data), "b2_spawn")

    def destroy_slots(self, env, slot):
        e = np.ascontiguousarray(env, np.int32); s = np.ascontiguousarray(slot, np.int32)
        self._ck(lib.b2_destroy_slots(self.ptr, e.size, e.ctypes.data, s.ctypes.data), "b2_destroy_slots")

    def slot_active(self, env_lo=0, env_hi=None):
        env_hi = self.nenv if env_hi is None else env_hi
        out = np.zeros((env_hi - env_lo, self.nslot), np.uint8)
        self._ck(lib.b2_slot_active(self.ptr, out.ctypes.data, env_lo, env_hi), "b2_slot_active")
        return out

    def pack_obs(self, dev_ptr):
        """[qpos | qvel] as fp32 [nq + nv][nenv] into a device buffer (the payload of the per-tick all-gather)."""
        self._ck(lib.b2_pack_obs(self.ptr, dev_ptr), "b2_pack_obs")

    def transfer_state_to(self, other):
        """The reference's add_old_state for whole batches: carry the state of every body that exists (by name) in both
        models over to `other` (a batch of the re-compiled world); returns the number of bodies carried."""
        n = lib.b2_transfer_state(self.ptr, other.ptr)
        if n < 0:
            raise B2Error("b2_transfer_state: " + _err())
        return n

    # ---- fused observation exchange (include/b2_batch.h) ----
    def obs_create(self, world, rank):
        """Allocate this GPU's observation buffer [world][nq + nv][nenv] fp32; returns (device pointer, 64-byte IPC handle)."""
        p = lib.b2_obs_create(self.ptr, world, rank)
        if not p:
            raise B2Error("b2_obs_create: " + _err())
        h = C.create_string_buffer(64)
        self._ck(lib.b2_obs_handle(self.ptr, h), "b2_obs_handle")
        return p, h.raw

    def obs_attach(self, handles=None, ptrs=None):
        """Open the peers' buffers (handles: world x 64 bytes from the other processes, or ptrs: device pointers of a
        single process driving several devices) and switch the exchange on."""
        if handles is not None:
            buf = C.create_string_buffer(bytes(handles), len(handles))
            self._ck(lib.b2_obs_attach(self.ptr, buf, None), "b2_obs_attach")
        else:
            arr = (C.c_void_p * len(ptrs))(*ptrs)
            self._ck(lib.b2_obs_attach(self.ptr, None, arr), "b2_obs_attach")

    def obs_read(self, world):
        """This GPU's observation buffer as a host array [world][nq + nv][nenv] (synchronises the batch's stream only)."""
        out = np.zeros((world, self.model.nq + self.model.nv, self.nenv), np.float32)
        self._ck(lib.b2_obs_read(self.ptr, out.ctypes.data), "b2_obs_read")
        return out

    def obs_enable(self, on=True):
        self._ck(lib.b2_obs_enable(self.ptr, int(bool(on))), "b2_obs_enable")

    def tick_resident(self):
        self._ck(lib.b2_tick_resident(self.ptr), "b2_tick_resident")

    def l2_flush(self, nbytes=256 << 20):
        self._ck(lib.b2_l2_flush(self.ptr, int(nbytes)), "b2_l2_flush")

    PROFILE_SLOTS = ["hw_write", "smooth", "collide", "make_constraint", "project", "pgs", "integrate", "hw_read"]

    def profile_begin(self, max_ticks):
        self._ck(lib.b2_profile_begin(self.ptr, int(max_ticks)), "b2_profile_begin")

    def profile_end(self):
        """-> (ticks profiled, {slot: summed ms})"""
        ms = np.zeros(len(self.PROFILE_SLOTS))
        n = self._ck(lib.b2_profile_end(self.ptr, ms.ctypes.data, ms.size), "b2_profile_end")
        return n, dict(zip(self.PROFILE_SLOTS, ms.tolist()))

    def mirror_env(self, env, data):
        self._ck(lib.b2_mirror_env(self.ptr, env, data.ptr), "b2_mirror_env")

    def load_env(self, env, data):
        self._ck(lib.b2_load_env(self.ptr, env, data.ptr), "b2_load_env")
