#!/usr/bin/env python
"""bench.py — env-steps/s of the batched rigid-body tick (BASELINE.json metric) on N B200 GPUs of one node.

  python bench.py --gpus N --steps K --warmup W [--config c3|c2|c4|c5] [--nenv E] [--impl reference]

One "step" = one control tick (write -> mj_step1 + controller -> read/mj_inverse -> mj_step2, reference
src/mj_main.cpp:82-112) of every environment of the batch.  The headline configuration is C3 (BASELINE.json
configs[2]: arm + tabletop objects with contacts, PGS, 16384 environments per GPU — the configuration of the
north-star target and of the 1 -> 8 GPU sweep); the other configurations are measured time-boxed in the same run and
reported compactly under "configs".  N > 1 is launched by torchrun, one rank per GPU; the environments shard across
ranks with no data-path collective, torch.distributed (NCCL) is used only for the barriers
and the max-over-ranks reduction of the device time.  Rank 0 prints ONE JSON line.

Timing: a block of exactly K steps is timed with CUDA events on the launching stream (L2 flushed before every step,
the flush outside the event pairs), bracketed by barrier + synchronize, max over ranks; the block is repeated until
at least 0.5 s of device time has been measured and the MEDIAN block is reported ("repeats" says how many).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC, UNIT = "env_steps_per_sec", "env-steps/s"
SETTLE_TICKS = {"c1": 0, "c2": 0, "c3": 150, "c4": 150, "c5": 120}
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # 148 SMs x 128 FMA lanes x 2 flop x 1.965 GHz (nominal, B200_PROFILING.md)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def b_alg(model, ncon_mean):
    """Algorithmic HBM bytes per env-step, fp32 SoA:
    read {qpos, qvel, qfrc_applied, qacc_warmstart, ddq, dq} + write {qpos, qvel, qacc, qacc_warmstart, qfrc_bias,
    qfrc_inverse} = 2 nq + 8 nv floats, + write {xpos, xquat} = 7 nbody floats, + 72 B per contact."""
    return 4 * (2 * model.nq + 8 * model.nv + 7 * model.nbody) + 72.0 * ncon_mean


def flops_per_env_step(model, ncon, nefc, iters):
    """Analytic fp32 operation count of one tick of one environment (adds + multiplies, FMA = 2), from the model's tree
    tables and the measured mean contact / row / iteration counts.  Stage formulas follow the loops of the oracle
    (oracle/oracle_smooth.cpp, oracle_constraint.cpp): this is the count an instrumented scalar run of those loops gives
    to within the few-percent spread of the per-contact chain depths, which are taken as the model's mean."""
    nv, nb, njnt, nM = model.nv, model.nbody, model.njnt, model.nM
    cnt = np.array([0] * nv)
    par = np.array(model.dof_parentid)
    for i in range(nv):
        j, c = i, 0
        while j >= 0:
            c += 1
            j = par[j]
        cnt[i] = c
    depth = float(cnt.mean()) if nv else 0.0
    fk = 190.0 * (nb - 1) + 75.0 * njnt                     # parent transform, quaternion products, quat2mat, inertial frame
    com = 16.0 * nb + 95.0 * (nb - 1) + 20.0 * nv           # subtree CoM, cinert (R I R^T + parallel axis), cdof
    crb = 10.0 * (nb - 1) + nv * 36.0 + 11.0 * float(cnt.sum())
    ldl = float(sum(sum(2 * (c - a) + 1 for a in range(1, c)) for c in cnt))
    vel = 60.0 * nv + 130.0 * (nb - 1) + 24.0 * nv + 12.0 * (nb - 1)   # comVel, RNE forward / backward, projection
    passive = 3.0 * nv + 20.0 * depth * float((np.array(model.body_gravcomp) != 0).sum())
    mulm = 2 * 4.0 * nM                                      # controller M ddq and inverse M qacc
    solve = 4.0 * nM + nv
    euler = (ldl + solve + 2 * nv) if (np.array(model.dof_damping) > 0).any() else 2.0 * nv
    geoms = 75.0 * model.ngeom
    collide = 14.0 * model.npair + 160.0 * ncon
    # constraint rows: per contact two chains of ~depth dofs, 45 flop per dof; B = M^-1 J^T per base direction ~ 4 nM of the
    # trees touched (approximated by 4 * wrow * depth); local matrix, parameters
    nb_dir = 3.2                                             # mean base directions per contact (condim 3 / 4 mix)
    wrow = min(nv, 2.0 * max(1.0, depth) * 2)                # compact row width ~ dofs of the touched trees
    rows = ncon * (2 * depth * 45.0 + nb_dir * (4.0 * wrow * max(1.0, depth) + 2.0 * wrow) + nb_dir * nb_dir * wrow + 80.0)
    rows += (nefc - ncon * 2 * (nb_dir - 1)) * (4.0 * wrow * max(1.0, depth) + 60.0)   # scalar rows (limits, equalities)
    nblk = ncon + max(0.0, nefc - ncon * 2 * (nb_dir - 1))
    pgs = iters * nblk * (2.0 * nb_dir * wrow * 2 + 2 * (nb_dir - 1) * 16.0)
    total = fk + com + crb + ldl + vel + passive + mulm + solve + euler + geoms + collide + rows + pgs
    return total, {"smooth": fk + com + crb + ldl + vel + passive + mulm + solve + geoms, "collide": collide, "rows": rows, "pgs": pgs, "euler": euler}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", os.environ.get("B2_BENCH_SMI_MS", "100")],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def real_mujoco():
    """probe for the real thing (python `mujoco` or libmujoco.so.2.3.7).  Returns (kind, handle) or
    (None, None).  Never found in this image; when found, tools/real_mujoco_diff.py diffs it against the oracle."""
    try:
        import mujoco  # noqa: F401
        return "python-mujoco " + getattr(mujoco, "__version__", "?"), mujoco
    except Exception:
        pass
    import ctypes
    for name in ("libmujoco.so.2.3.7", "libmujoco.so"):
        try:
            return name, ctypes.CDLL(name)
        except OSError:
            continue
    return None, None


def real_mujoco_diff(cfg):
    """When a real MuJoCo is importable, step the same MJCF / states with it and with the oracle and report the
    difference (tools/real_mujoco_diff.py); otherwise say that it is absent."""
    kind, _ = real_mujoco()
    if kind is None:
        return {"present": False}
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import real_mujoco_diff as rmd
        out = rmd.diff_config(cfg)
        out["present"] = True; out["kind"] = kind
        return out
    except Exception as e:   # the probe must never take the bench down
        return {"present": True, "kind": kind, "error": repr(e)[:200]}


def config_dict(cfg, asset, desc, nenv, world, model, kp_n):
    """Static description of the workload: identical in both arms."""
    slots = cfg == "c5"
    return {"workload": "%s: %s" % (cfg, desc), "model": asset, "envs_per_gpu": nenv, "envs_total": nenv * world, "timestep": 0.005,
            "tick": ("step1+step2 with 8 of 20 object slots live per environment; one destroy + one spawn per environment every 60 ticks in the end-to-end loop" if slots else
                     "write+step1+controller+inverse+step2+read" + (" with PD (kp 200, kd 50) on %d arm joints" % kp_n if kp_n else "")),
            "solver": "PGS, %d iterations max" % int(model.int("opt.iterations")),
            "l2": "flushed before every timed step (256 MiB memset, outside the event pairs)",
            "states": "seeded per environment (workloads.config_state), draws penetrating deeper than 1 cm redrawn; %d settle ticks before timing; "
                      "the end-to-end loop restarts from the same settled state as the device-timed blocks (the tick time drifts with the contact state)" % SETTLE_TICKS[cfg]}


def oracle_redraw(cfg, m, envs, pool, max_depth=0.01, max_rounds=8):
    """The reference arm's states: the same seeded draws as the GPU arm (workloads.load_config), with the same redraw of
    deeply penetrating starts, checked here with the oracle's collision pass."""
    import mujoco_sim_b200 as b2
    from mujoco_sim_b200 import workloads as w
    from oracle import pyoracle as orc
    base = 0xB200 + int(cfg[1])
    qpos, qvel, frc = w.config_state(cfg, m, envs, seed=base)
    if m.npair == 0 or cfg == "c5":
        return qpos, qvel, frc
    d = pool[0]
    for rnd in range(1, max_rounds + 1):
        bad = []
        for i in range(envs.size):
            d.qpos[:] = qpos[i]
            orc.call("kinematics", m, d); orc.call("collision", m, d)
            if int(d.ncon) and b2.data_contacts(d)[2].min() < -max_depth:
                bad.append(i)
        if not bad:
            break
        bad = np.array(bad)
        q2, v2, f2 = w.config_state(cfg, m, envs[bad], seed=base + 7919 * rnd)
        qpos[bad], qvel[bad], frc[bad] = q2, v2, f2
    return qpos, qvel, frc


def cpu_reference(cfg, nenv_sample, steps, warmup, threads=None, min_seconds=0.0, redraw=True):
    """The reference's CPU path for this tick, restated by the fp64 oracle (libmujoco is not available: SURVEY 8c),
    on `threads` host threads (default: all).  The K-step block is repeated until min_seconds have been measured; returns
    (median env-steps/s, threads used, median block seconds, repeats)."""
    import mujoco_sim_b200 as b2
    from mujoco_sim_b200 import workloads as w
    from oracle import pyoracle as orc
    asset = w.CONFIGS[cfg][0]
    m = b2.Model(b2.asset(asset))
    threads = threads or os.cpu_count() or 1
    threads = max(1, min(threads, nenv_sample))
    pool = [b2.Data(m) for _ in range(threads)]
    envs = np.arange(nenv_sample)
    if redraw:
        qpos, qvel, frc = oracle_redraw(cfg, m, envs, pool)
    else:
        qpos, qvel, frc = w.config_state(cfg, m, envs)
    qpos = np.ascontiguousarray(qpos); qvel = np.ascontiguousarray(qvel)
    ws = np.zeros((nenv_sample, m.nv))
    if cfg == "c5":
        return cpu_reference_c5(m, pool, qpos, qvel, steps, warmup, min_seconds)
    hw, ctl, kp, kd = w.control_spec(cfg, m)
    dadr = np.array(m.jnt_dofadr)[hw]
    ddq = np.zeros((nenv_sample, m.nv))
    ddq[:, dadr] = w.commands(cfg, m, envs)
    dq = np.zeros((nenv_sample, m.nv))
    pd = {}
    if kp is not None:
        kpv, kdv = np.zeros(m.nv), np.zeros(m.nv)
        kpv[dadr], kdv[dadr] = kp, kd
        pd = {"pd_kp": kpv, "pd_kd": kdv}
    settle = SETTLE_TICKS[cfg]
    if settle:
        orc.tick_batch(m, pool, settle, qpos, qvel, ws, None, ddq, dq, ctl, True, **pd)
    if warmup:
        orc.tick_batch(m, pool, warmup, qpos, qvel, ws, None, ddq, dq, ctl, True, **pd)
    blocks, used, total = [], 1, 0.0
    while True:
        t0 = time.perf_counter()
        used = orc.tick_batch(m, pool, steps, qpos, qvel, ws, None, ddq, dq, ctl, True, **pd)
        dt = time.perf_counter() - t0
        blocks.append(dt); total += dt
        if total >= min_seconds or len(blocks) >= 1000:
            break
    med = float(np.median(blocks))
    return nenv_sample * steps / med, used, med, len(blocks)


def cpu_reference_c5(m, pool, qpos, qvel, steps, warmup, min_seconds=0.0):
    """C5 on the CPU: the same request stream (8 of 20 slots live, one destroy + one spawn per environment every 60
    ticks) applied to the oracle's state arrays between chunks of ticks; inactive slots are put back to their parking
    place at every chunk boundary."""
    from mujoco_sim_b200 import workloads as w
    from oracle import pyoracle as orc
    nenv = qpos.shape[0]
    envs = np.arange(nenv)
    slots = w.c5_slot_bodies(m)
    qadr = np.array([m.jnt_qposadr[m.body_jntadr[b]] for b in slots]); dadr = np.array([m.jnt_dofadr[m.body_jntadr[b]] for b in slots])
    live = np.zeros((nenv, w.NSLOT_C5), bool)
    ws = np.zeros((nenv, m.nv))

    def park():
        for s in range(w.NSLOT_C5):
            off = ~live[:, s]
            qpos[off, qadr[s]:qadr[s] + 7] = w.park_pose(s); qvel[off, dadr[s]:dadr[s] + 6] = 0; ws[off, dadr[s]:dadr[s] + 6] = 0

    def spawn(k):
        s = (envs + 3 * k) % w.NSLOT_C5
        pose = w.c5_spawn_pose(envs, k)
        for e in range(nenv):
            qpos[e, qadr[s[e]]:qadr[s[e]] + 7] = pose[e]; qvel[e, dadr[s[e]]:dadr[s[e]] + 6] = 0; live[e, s[e]] = True
    park()
    for k in range(w.C5_INITIAL):
        spawn(k)
    rnd = 0

    def run(n):
        nonlocal rnd
        used = 1
        done = 0
        while done < n:
            c = min(60, n - done)
            used = orc.tick_batch(m, pool, c, qpos, qvel, ws)
            done += c
            if c == 60:
                live[envs, (envs + 3 * rnd) % w.NSLOT_C5] = False
                park(); spawn(rnd + w.C5_INITIAL); rnd += 1
        return used
    run(SETTLE_TICKS["c5"] + warmup)
    blocks, total, used = [], 0.0, 1
    while True:
        t0 = time.perf_counter()
        used = run(steps)
        dt = time.perf_counter() - t0
        blocks.append(dt); total += dt
        if total >= min_seconds or len(blocks) >= 1000:
            break
    med = float(np.median(blocks))
    return nenv * steps / med, used, med, len(blocks)


def drift_report(cfg, nenv=64, horizons=(1, 10, 100, 1000), precision=None):
    """BASELINE metric, second part: relative L2 drift of qpos (fp32 CUDA tick vs the fp64 CPU oracle) from identical
    states.  Both sides advance tick by tick; after every tick the contact lists (geom1, geom2 in order) of the two are
    compared, and an environment leaves the drift statistics at the first tick where they differ (drift
    "up to the first tick where the contact index sets differ, which must itself be reported").  Part of the cpu_baseline
    leg: the oracle is the checker here, not the thing measured."""
    import mujoco_sim_b200 as b2
    from mujoco_sim_b200 import workloads as w
    from oracle import pyoracle as orc
    m = b2.Model(b2.asset(w.CONFIGS[cfg][0]))
    bt = b2.Batch(m, nenv, precision=precision or b2.engine.F32)
    qpos, qvel, frc, _ = w.load_config(cfg, bt)
    rq, rv, rf = (np.ascontiguousarray(x, np.float64).copy() for x in (qpos, qvel, frc))
    contacts = m.npair > 0
    out = {"envs": nenv, "ticks": list(horizons), "median": [], "max": [], "frac_below_1e-4": [], "envs_compared": []}
    last = max(horizons)
    if not contacts:
        # contact-free: no index sets to compare, the oracle advances in bulk between the horizons
        ws = np.zeros((nenv, m.nv))
        pool = [b2.Data(m) for _ in range(min(16, os.cpu_count() or 1, nenv))]
        done = 0
        for k in horizons:
            bt.step(k - done); bt.sync()
            orc.tick_batch(m, pool, k - done, rq, rv, ws, rf)
            done = k
            rel = np.linalg.norm(bt.get("qpos") - rq, axis=1) / np.maximum(np.linalg.norm(rq, axis=1), 1e-12)
            out["median"].append(float(np.median(rel))); out["max"].append(float(rel.max()))
            out["frac_below_1e-4"].append(float((rel < 1e-4).mean())); out["envs_compared"].append(nenv)
        bt.close()
        return out
    ds = [b2.Data(m) for _ in range(nenv)]
    for e in range(nenv):
        ds[e].qpos[:] = rq[e]; ds[e].qvel[:] = rv[e]; ds[e].qfrc_applied[:] = rf[e]; ds[e].qacc[:] = 0; ds[e].qacc_warmstart[:] = 0
    ncm = m.nconmax
    first_diff = np.full(nenv, -1, int)
    for k in range(1, last + 1):
        bt.step(1)
        for e in range(nenv):
            orc.call("step", m, ds[e])
        gi = bt.get("contact_int")          # [env][5 * nconmax]: geom1 | geom2 | dim | pair | efc (contacts of the tick just run)
        gn = bt.get("ncon")[:, 0]
        for e in range(nenv):
            if first_diff[e] >= 0:
                continue
            nc = int(ds[e].ncon)
            same = nc == int(gn[e])
            if same and nc:
                g1, g2, _ = b2.data_contacts(ds[e], nc)
                same = np.array_equal(g1, gi[e, :nc]) and np.array_equal(g2, gi[e, ncm:ncm + nc])
            if not same:
                first_diff[e] = k
        if k in horizons:
            gq = bt.get("qpos")
            rqk = np.array([np.array(ds[e].qpos) for e in range(nenv)])
            rel = np.linalg.norm(gq - rqk, axis=1) / np.maximum(np.linalg.norm(rqk, axis=1), 1e-12)
            ok = first_diff < 0
            out["envs_compared"].append(int(ok.sum()))
            if ok.any():
                out["median"].append(float(np.median(rel[ok]))); out["max"].append(float(rel[ok].max()))
                out["frac_below_1e-4"].append(float((rel[ok] < 1e-4).mean()))
            else:
                out["median"].append(None); out["max"].append(None); out["frac_below_1e-4"].append(None)
    diverged = first_diff[first_diff >= 0]
    out["first_contact_set_mismatch_tick"] = {"envs_with_mismatch": int(diverged.size), "earliest": int(diverged.min()) if diverged.size else None,
                                              "median": float(np.median(diverged)) if diverged.size else None}
    bt.close()
    return out


def measure_exchange(bt, m, nenv, K, ctx, tick_resident, do_flush, starts, ends, stream, allmax, barrier, blocks=5):
    """Per-tick observation exchange: device time of K-step blocks with (a) the fused peer-store epilogue and (b) for
    N > 1 the pack kernel + NCCL all-gather; the fused buffer is verified against the ranks' own states."""
    import torch
    rank, world, dist = ctx["rank"], ctx["world"], ctx["dist"]
    nobs = m.nq + m.nv
    _, handle = bt.obs_create(world, rank)
    if dist is not None:
        hs = [None] * world
        dist.all_gather_object(hs, handle)
        bt.obs_attach(handles=b"".join(hs))
    else:
        bt.obs_attach(handles=handle)

    def block(after=None):
        barrier(); torch.cuda.synchronize(); bt.sync()
        for k in range(K):
            do_flush()
            starts[k].record(stream)
            tick_resident()
            if after is not None:
                after()
            ends[k].record(stream)
        bt.sync(); torch.cuda.synchronize(); barrier()
        return allmax(sum(s.elapsed_time(e) for s, e in zip(starts, ends))) / K

    # The tick time depends on the (evolving) contact state, so the variants are INTERLEAVED block by block — off, fused,
    # (off, pack + NCCL) — and compared by their medians; toggling the exchange re-captures the tick's CUDA graph, which
    # the three untimed ticks after every toggle absorb.
    obs_local = obs_all = None
    if dist is not None:
        obs_local = torch.empty((nobs, nenv), dtype=torch.float32, device="cuda")
        obs_all = torch.empty((world, nobs, nenv), dtype=torch.float32, device="cuda")

    def nccl_exchange():
        bt.pack_obs(obs_local.data_ptr())
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(obs_all, obs_local)
    off_ms, fused_ms, nccl_ms = [], [], []
    for _ in range(blocks):
        bt.obs_enable(False)
        for _ in range(3):
            tick_resident()
        off_ms.append(block())
        bt.obs_enable(True)
        for _ in range(3):
            tick_resident()
        fused_ms.append(block())
        if dist is not None:
            bt.obs_enable(False)
            for _ in range(3):
                tick_resident()
            nccl_ms.append(block(nccl_exchange))
    # verification: after a barrier every rank's buffer holds every rank's state
    bt.obs_enable(True)
    for _ in range(2):
        tick_resident()
    bt.sync(); barrier()
    got = bt.obs_read(world)
    mine = np.concatenate([bt.get("qpos", layout=1, dtype=np.float32), bt.get("qvel", layout=1, dtype=np.float32)])
    ok = bool(np.array_equal(got[rank], mine))
    if dist is not None:
        allst = torch.empty((world, nobs, nenv), dtype=torch.float32, device="cuda")
        dist.all_gather_into_tensor(allst, torch.from_numpy(mine).cuda())
        ok = ok and bool(np.array_equal(got, allst.cpu().numpy()))
        flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item() > 0.5)
    bt.obs_enable(False)
    res = {"no_exchange_ms_per_step": float(np.median(off_ms)), "fused_peer_store_ms_per_step": float(np.median(fused_ms)),
           "verified_equal_to_all_ranks_state": ok, "bytes_per_rank_per_tick": int(4 * nobs * nenv), "blocks_interleaved": blocks,
           "how": "stores from k_integrate's epilogue into every GPU's buffer through NVLink peer mappings (CUDA IPC); no pack kernel, no collective"}
    if dist is not None:
        res["nccl_allgather_ms_per_step"] = float(np.median(nccl_ms))
    return res


def measure(cfg, nenv, K, W, ctx, min_dev_s=0.5, detail=True, max_repeats=400, precision=None):
    """Settle + warm up, then timed K-step blocks (device events, max over ranks, median block) and the end-to-end loop
    through host buffers.  Returns a dict with everything the JSON line needs for this configuration."""
    import torch
    import mujoco_sim_b200 as b2
    from mujoco_sim_b200 import workloads as w
    rank, world, local_rank, dist, flush, gather = ctx["rank"], ctx["world"], ctx["local_rank"], ctx["dist"], ctx["flush"], ctx["gather"]
    asset, _, desc = w.CONFIGS[cfg]

    def barrier():
        if dist is not None:
            dist.barrier()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    m = b2.Model(b2.asset(asset))
    bt = b2.Batch(m, nenv, device=local_rank, precision=precision or b2.engine.F32)
    env_offset = rank * nenv  # contiguous shards of one global batch
    w.load_config(cfg, bt, env_offset=env_offset)
    hw, ctl, kp, kd = w.control_spec(cfg, m)
    nhw = hw.size
    slots = cfg == "c5"   # no ros_control joints: the host exchange of this config is spawn / destroy requests in, body poses out
    keep = []
    if not slots:
        bt.set_controlled(ctl)
        bt.set_hw_joints(hw)
        if kp is not None:
            bt.set_pd(kp, kd)
        cmd = w.commands(cfg, m, np.arange(env_offset, env_offset + nenv))
        eff_cmd = torch.from_numpy(np.ascontiguousarray(cmd.T.astype(np.float32))).pin_memory()
        vel_cmd = torch.zeros((nhw, nenv), dtype=torch.float32).pin_memory()
        pos_o = torch.empty((nhw, nenv), dtype=torch.float32).pin_memory()
        vel_o = torch.empty_like(pos_o).pin_memory()
        eff_o = torch.empty_like(pos_o).pin_memory()
        keep = [eff_cmd, vel_cmd, pos_o, vel_o, eff_o]
        host_args = (vel_cmd.data_ptr(), eff_cmd.data_ptr(), pos_o.data_ptr(), vel_o.data_ptr(), eff_o.data_ptr())
        bt.write_commands(vel_cmd.numpy(), eff_cmd.numpy())  # uploads the commands once: they stay resident in HBM for the device-timed loop
        tick_resident = bt.tick_resident
        h2d, d2h = 2 * nhw * nenv * 4, 3 * nhw * nenv * 4

        def tick_e2e(k):
            bt.tick_host_raw(*host_args)
            return pos_o
    else:
        w.c5_init(bt, env_offset)
        churn = {"round": 0}
        tick_resident = lambda: bt.step(1)   # noqa: E731
        # per tick: body poses out (what the ROS layer publishes); every 60 ticks one destroy + one spawn per environment in
        h2d, d2h = (2 * 2 * 4 + 7 * 4) * nenv // 60, 3 * m.nbody * nenv * 4

        def tick_e2e(k):
            if k % 60 == 0:
                w.c5_churn(bt, churn["round"], env_offset); churn["round"] += 1
            bt.step(1)
            return bt.get("xpos", dtype=np.float32)   # synchronises

    stream = torch.cuda.ExternalStream(bt.stream, device=torch.device("cuda", local_rank))
    nobs = m.nq + m.nv
    if gather:
        obs_local = torch.empty((nobs, nenv), dtype=torch.float32, device="cuda")
        obs_all = torch.empty((world, nobs, nenv), dtype=torch.float32, device="cuda")

    def exchange():
        if gather:
            bt.pack_obs(obs_local.data_ptr())
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(obs_all, obs_local)

    def do_flush():
        if flush:
            bt.l2_flush(256 << 20)  # 256 MiB memset on the batch's stream > 126 MB L2: evicts the state between timed steps

    # settle (contacts form) + warm-up, untimed
    for _ in range(SETTLE_TICKS[cfg] + W):
        tick_resident()
    bt.sync()
    ncon_mean = float(bt.get("ncon").mean()) if m.npair > 0 else 0.0
    nefc_mean = float(bt.get("nefc").mean())
    iter_mean = float(bt.get("solver_iter").mean()) if m.npair > 0 else 0.0
    # The tick time drifts with the contact state of the simulation (the arms keep pushing props: more environments need
    # all 100 solver iterations as time goes on), so the end-to-end loop below must see the SAME stretch of the simulation
    # as the device-timed blocks: the settled state is kept and restored in front of it.
    state_fields = ("qpos", "qvel", "qacc", "qacc_warmstart", "qfrc_applied", "time")
    dt_np = np.float64 if precision == b2.engine.F64 else np.float32
    snap = None if slots else {f: bt.get(f, layout=b2.engine.NATIVE, dtype=dt_np) for f in state_fields}

    # ---- device-timed blocks: K steps each, inputs resident in HBM, CUDA events on the launching stream ----
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    blocks, total_ms, launches = [], 0.0, 0
    sampler = ClockSampler(local_rank) if (detail and not os.environ.get("B2_BENCH_NO_SAMPLER")) else None
    if sampler:
        sampler.start()
    t_wall = time.perf_counter()
    while True:
        barrier(); torch.cuda.synchronize(); bt.sync()
        l0 = bt.launch_count
        for k in range(K):
            do_flush()
            starts[k].record(stream)
            tick_resident()             # the tick kernels (one launch for a limit-only chain, else a CUDA-graph replay)
            exchange()
            ends[k].record(stream)
        bt.sync(); torch.cuda.synchronize()
        launches = bt.launch_count - l0
        barrier()
        blk = allmax(sum(s.elapsed_time(e) for s, e in zip(starts, ends)))
        blocks.append(blk); total_ms += blk
        # every rank sees the same max-reduced numbers, so every rank leaves the loop at the same repeat
        if total_ms >= 1e3 * min_dev_s or len(blocks) >= max_repeats:
            break
    max_ms = float(np.median(blocks))
    # per-kernel device time: K ticks again, launched eagerly with CUDA events between the kernels
    bt.profile_begin(K)
    for k in range(K):
        do_flush()
        tick_resident()
    bt.sync(); torch.cuda.synchronize()
    nprof, slot_ms = bt.profile_end()

    # ---- end to end: the same tick through the C ABI with HOST buffers (H2D commands, D2H joint states), over the same
    #      stretch of the simulation as the device-timed blocks ----
    if snap is not None:
        for k in range(3):          # first use of the host-buffer tick: eager, capture, replay
            tick_e2e(k)
        for f in state_fields:
            bt.set(f, snap[f], layout=b2.engine.NATIVE)
        bt.sync()
    e2e_blocks, e2e_total = [], 0.0
    while True:
        barrier()
        e2e_s = 0.0
        for k in range(K):
            do_flush()
            bt.sync()
            t0 = time.perf_counter()
            last = tick_e2e(k)
            if gather:
                exchange()
                stream.synchronize()
            e2e_s += time.perf_counter() - t0
        e2e_s = allmax(e2e_s)
        e2e_blocks.append(e2e_s); e2e_total += e2e_s
        if e2e_total >= min_dev_s or len(e2e_blocks) >= max_repeats:
            break
    e2e_max = float(np.median(e2e_blocks))
    _ = float(last.sum())   # the result is read on the host
    clocks = sampler.stop() if sampler else None
    wall = time.perf_counter() - t_wall

    # ---- observation exchange, measured next to the exchange-free number: the fused peer-store epilogue
    #      (every rank writes its [qpos | qvel] into every GPU's buffer from inside the integrate kernel; no pack kernel, no
    #      collective call) and, for N > 1, the pack + NCCL all-gather baseline ----
    exch = None
    if detail and not slots:
        try:
            exch = measure_exchange(bt, m, nenv, K, ctx, tick_resident, do_flush, starts, ends, stream, allmax, barrier)
        except Exception as e:   # never take the bench line down
            exch = {"error": repr(e)[:300]}

    total_envs = nenv * world
    value = total_envs * K / (max_ms * 1e-3)
    peak, peak_src = peaks()
    kern = {k: v / max(1, nprof) for k, v in slot_ms.items()}
    inner = {k: v for k, v in kern.items() if k not in ("hw_write", "hw_read")}
    dom = max(inner, key=inner.get)
    tick_ms = max_ms / K
    balg = b_alg(m, ncon_mean)
    achieved = balg * nenv / (tick_ms * 1e-3) / 1e9          # whole tick: B_alg is the tick's bytes, so the tick's time divides it
    flops, fsplit = flops_per_env_step(m, ncon_mean, nefc_mean, iter_mean)
    fp32_tf = flops * nenv / (tick_ms * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_%s.json" % cfg)
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                tj = json.load(f)
            traffic = tj.get("tick_total", tj.get(dom))
        except Exception:
            traffic = None
    hbm_frac, fp32_frac = achieved / peak, fp32_tf / FP32_PEAK_TFLOPS
    res = {
        "value": value, "ms_per_step": tick_ms, "repeats": len(blocks), "block_ms_min_med_max": [float(min(blocks)), max_ms, float(max(blocks))],
        "config": config_dict(cfg, asset, desc, nenv, world, m, nhw if kp is not None else 0),
        "stats": {"mean_ncon": ncon_mean, "mean_nefc": nefc_mean, "mean_solver_iter": iter_mean, "kernels": bt.path_name,
                  "obs_allgather": ("NCCL all_gather of [qpos|qvel] fp32, %d B per rank per tick, inside the timed region" % (4 * nobs * nenv)) if gather else "none: shards are independent, no data-path collective"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": hbm_frac, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_env_step": balg,
                     "scope": "whole tick (all kernels of one graph replay): achieved = B_alg x envs / tick time",
                     "fp32_flop_per_env_step": flops, "fp32_flop_split": fsplit, "fp32_achieved_tflops": fp32_tf, "fp32_peak_tflops": FP32_PEAK_TFLOPS,
                     "fp32_frac": fp32_frac, "fp32_count": "analytic, from the model's tree tables and the measured contact / row / iteration means (bench.py flops_per_env_step)",
                     "binding": ("fp32 issue / latency" if fp32_frac > hbm_frac else "hbm") + ": hbm_frac %.4f, fp32_frac %.4f" % (hbm_frac, fp32_frac),
                     "dominant_kernel": {"pgs": "k_pgs", "smooth": bt.path_name.split("+")[0]}.get(dom, "k_" + dom),
                     "kernel_ms": inner[dom], "kernel_share_of_step": inner[dom] / max(1e-12, sum(kern.values())), "kernel_ms_all": kern,
                     "kernel_timing": "CUDA events between the kernels, K ticks re-run eagerly right after the graph-replayed timed blocks"},
        "e2e": {"value": total_envs * K / e2e_max, "unit": UNIT, "h2d_bytes_per_step": int(h2d * world),
                "d2h_bytes_per_step": int(d2h * world), "ms_per_step": 1e3 * e2e_max / K, "repeats": len(e2e_blocks)},
        "gpu_launches": int(launches),
        "clocks": clocks, "wall_s": wall, "obs_exchange": exch,
    }
    del keep
    bt.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default=os.environ.get("B2_BENCH_CONFIG", "c3"), choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--nenv", type=int, default=0, help="environments per GPU (default: the config's)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the time-boxed c2 / c4 / c5 measurements of the 'configs' object")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--gather-obs", action="store_true",
                    help="N > 1: add the optional per-tick NCCL all-gather of observations (SURVEY.md 8e) to the timed region; "
                         "off by default because the path itself has no exchange step (environments are independent)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from mujoco_sim_b200 import workloads as w
    asset, nenv_default, desc = w.CONFIGS[args.config]
    nenv = args.nenv or nenv_default

    if args.impl == "reference":
        # the reference arm: the CPU tick on the box's host cores, rank 0 only
        if rank != 0:
            return
        import mujoco_sim_b200 as b2
        m = b2.Model(b2.asset(asset))
        _, _, kp, _ = w.control_spec(args.config, m)
        nhw = w.control_spec(args.config, m)[0].size
        # a bounded sample of the workload per step, sized so that the K-step block takes 0.2 - 2 s on the host cores
        sample = min(nenv, {"c2": 4096, "c3": 2048, "c4": 512, "c5": 512}[args.config])
        val, used, med, reps = cpu_reference(args.config, sample, args.steps, args.warmup, min_seconds=2.0)
        out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": 1e3 * med / args.steps, "repeats": reps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic",
               "config": config_dict(args.config, asset, desc, nenv, world, m, nhw if kp is not None else 0),
               "cpu_baseline": {"value": val, "unit": UNIT, "cores": used, "kind": "port",
                                "sample": "%d of the %d envs per step x %d ticks per block, median of %d blocks (>= 2 s); fp64 oracle restatement of MuJoCo 2.3.7 semantics (real MuJoCo present: %s)"
                                          % (sample, nenv, args.steps, reps, real_mujoco()[0] or "no")},
               "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(out))
        return

    import torch
    import mujoco_sim_b200 as b2
    if not torch.cuda.is_available() or b2.lib.b2_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = {"rank": rank, "world": world, "local_rank": local_rank, "dist": dist, "flush": not args.no_flush,
           "gather": dist is not None and args.gather_obs}

    res = measure(args.config, nenv, args.steps, args.warmup, ctx)
    others = {}
    if not args.no_other_configs:
        # the other BASELINE configurations, time-boxed: fewer steps per block, 0.25 s of device time each
        for c in ("c2", "c4", "c5"):
            if c == args.config:
                continue
            r = measure(c, w.CONFIGS[c][1], min(args.steps, 20), 3, ctx, min_dev_s=0.25, detail=False, max_repeats=100)
            others[c] = {"workload": r["config"]["workload"], "envs_per_gpu": r["config"]["envs_per_gpu"], "value": r["value"], "e2e": r["e2e"]["value"],
                         "ms_per_step": r["ms_per_step"], "roofline_frac": r["roofline"]["frac"], "fp32_frac": r["roofline"]["fp32_frac"],
                         "kernel_ms_all": r["roofline"]["kernel_ms_all"], "mean_ncon": r["stats"]["mean_ncon"], "mean_nefc": r["stats"]["mean_nefc"],
                         "mean_solver_iter": r["stats"]["mean_solver_iter"], "kernels": r["stats"]["kernels"], "gpu_launches": r["gpu_launches"], "repeats": r["repeats"]}

        # the contact-free chain once more with fp64 arithmetic: what the north star's drift bound (< 1e-4 over 1000 ticks)
        # costs on a chaotic arm, where fp32 rounding is amplified exponentially whatever the integrator does
        if "c2" in others:
            r = measure("c2", w.CONFIGS["c2"][1], min(args.steps, 20), 3, ctx, min_dev_s=0.25, detail=False, max_repeats=100, precision=b2.engine.F64)
            others["c2_f64"] = {"workload": r["config"]["workload"] + " — fp64 state and arithmetic (b2_create precision = B2_F64)", "envs_per_gpu": r["config"]["envs_per_gpu"],
                                "value": r["value"], "e2e": r["e2e"]["value"], "ms_per_step": r["ms_per_step"], "kernels": r["stats"]["kernels"], "gpu_launches": r["gpu_launches"], "repeats": r["repeats"]}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    out = {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"], "repeats": res["repeats"], "block_ms_min_med_max": res["block_ms_min_med_max"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": res["config"], "stats": res["stats"], "roofline": res["roofline"], "e2e": res["e2e"],
        "gpu_launches": res["gpu_launches"], "clocks": res["clocks"], "obs_exchange": res["obs_exchange"],
    }
    if not args.no_cpu_baseline and world == 1:
        sample = {"c2": 4096, "c3": 2048, "c4": 512, "c5": 512}[args.config]
        csteps = {"c2": 500, "c3": 100, "c4": 50, "c5": 60}[args.config]
        val, used, med, reps = cpu_reference(args.config, sample, csteps, 3, min_seconds=10.0)   # about 10 s of CPU work
        out["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": used, "kind": "port",
                               "sample": "%d envs x %d ticks per block, median of %d blocks (>= 10 s); fp64 oracle restatement of MuJoCo 2.3.7 semantics "
                                         "(real MuJoCo present: %s)" % (sample, csteps, reps, real_mujoco()[0] or "no")}
        out["drift"] = drift_report(args.config, nenv=32 if args.config == "c4" else 64)
        out["real_mujoco_diff"] = real_mujoco_diff(args.config)
        if others:
            for c in others:
                if c == "c5":
                    continue
                try:
                    d = drift_report(c[:2], nenv=16 if c == "c4" else 64, horizons=(1, 10, 100, 1000) if c.startswith("c2") else (1, 10, 100, 300),
                                     precision=b2.engine.F64 if c.endswith("_f64") else None)
                    others[c]["drift_frac_below_1e-4"] = dict(zip([str(t) for t in d["ticks"]], d["frac_below_1e-4"]))
                    others[c]["drift_max"] = dict(zip([str(t) for t in d["ticks"]], d["max"]))
                    if "first_contact_set_mismatch_tick" in d:
                        others[c]["first_contact_set_mismatch_tick"] = d["first_contact_set_mismatch_tick"]
                except Exception as e:
                    others[c]["drift_error"] = repr(e)[:200]
    if others:
        out["configs"] = others
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
