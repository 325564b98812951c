"""ctypes binding of the fp64 CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by mujoco_sim_b200."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
if not os.path.exists(LIB_PATH):
    raise ImportError(LIB_PATH + " is missing: build it with `python mujoco_sim_b200/build.py`")
olib = C.CDLL(LIB_PATH)
_vp, _i = C.c_void_p, C.c_int
for _n in ["omj_kinematics", "omj_comPos", "omj_crb", "omj_factorM", "omj_collision", "omj_makeConstraint",
           "omj_projectConstraint", "omj_fwdPosition", "omj_comVel", "omj_passive", "omj_referenceConstraint",
           "omj_fwdVelocity", "omj_fwdAcceleration", "omj_fwdConstraint", "omj_Euler", "omj_energy", "omj_step1",
           "omj_step2", "omj_step", "omj_forward", "omj_inverse", "omj_invConstraint", "omj_rnePostConstraint"]:
    getattr(olib, _n).restype = None
    getattr(olib, _n).argtypes = [_vp, _vp]
olib.omj_rne.restype = None
olib.omj_rne.argtypes = [_vp, _vp, _i, _vp]
olib.omj_mulM.restype = None
olib.omj_mulM.argtypes = [_vp, _vp, _vp, _vp]
olib.omj_solveM.restype = None
olib.omj_solveM.argtypes = [_vp, _vp, _vp, _i]
olib.omj_fullM.restype = None
olib.omj_fullM.argtypes = [_vp, _vp, _vp]
olib.omj_jac.restype = None
olib.omj_jac.argtypes = [_vp, _vp, _vp, _vp, _vp, _i]
olib.omj_controller.restype = None
olib.omj_controller.argtypes = [_vp, _vp, _vp, _vp, _vp]
olib.omj_set_odom_vels.restype = None
olib.omj_set_odom_vels.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp]
olib.omj_tick.restype = None
olib.omj_tick.argtypes = [_vp, _vp, _vp, _vp, _vp, _i]
olib.omj_tick_batch.restype = _i
olib.omj_tick_batch.argtypes = [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]
olib.omj_tick_batch_pd.restype = _i
olib.omj_tick_batch_pd.argtypes = [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]
olib.omj_pair_supported.restype = _i
olib.omj_pair_supported.argtypes = [_i, _i]


def _p(a):
    return None if a is None else a.ctypes.data


def call(name, model, data):
    getattr(olib, "omj_" + name)(model.ptr, data.ptr)


def tick(model, data, ddq=None, dq=None, controlled=None, do_inverse=False):
    olib.omj_tick(model.ptr, data.ptr, _p(ddq), _p(dq), _p(controlled), int(do_inverse))


def controller(model, data, ddq, dq, controlled):
    olib.omj_controller(model.ptr, data.ptr, _p(ddq), _p(dq), _p(controlled))


def set_odom_vels(model, data, lin_dof, ang_dof, ang_qpos, vels):
    a = [np.ascontiguousarray(x, np.int32) for x in (lin_dof, ang_dof, ang_qpos)]
    v = np.ascontiguousarray(vels, np.float64)
    olib.omj_set_odom_vels(model.ptr, data.ptr, a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, v.ctypes.data)


def full_M(model, data):
    nv = model.nv
    out = np.zeros((nv, nv))
    olib.omj_fullM(model.ptr, data.ptr, out.ctypes.data)
    return out


def rne(model, data, flg_acc):
    out = np.zeros(model.nv)
    olib.omj_rne(model.ptr, data.ptr, int(flg_acc), out.ctypes.data)
    return out


def tick_batch(model, pool, nsteps, qpos, qvel, qacc_warmstart=None, qfrc_applied=None, ddq=None, dq=None, controlled=None,
               do_inverse=False, qfrc_inverse_out=None, pd_kp=None, pd_kd=None):
    """Advance every row of qpos/qvel ([nenv][n] float64, updated in place) nsteps ticks; one thread per pool entry.
    pd_kp / pd_kd ([nv] float64): dofs with a non-zero gain read ddq as a position target (PD stage, b2_set_pd)."""
    arr = (C.c_void_p * len(pool))(*[d.ptr for d in pool])
    nenv = qpos.shape[0]
    for a in (qpos, qvel, qacc_warmstart, qfrc_applied, ddq, dq, qfrc_inverse_out):
        assert a is None or (a.dtype == np.float64 and a.flags.c_contiguous)
    for a in (pd_kp, pd_kd):
        assert a is None or (a.dtype == np.float64 and a.flags.c_contiguous and a.size == model.nv)
    return olib.omj_tick_batch_pd(model.ptr, arr, len(pool), nenv, nsteps, _p(qpos), _p(qvel), _p(qacc_warmstart), _p(qfrc_applied),
                                  _p(ddq), _p(dq), _p(controlled), int(do_inverse), _p(qfrc_inverse_out), _p(pd_kp), _p(pd_kd))
