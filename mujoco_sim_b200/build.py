"""In-tree build of the native libraries (no JIT cache: the built .so files travel with the repo snapshot).

  mujoco_sim_b200/lib/libb2sim.so   CUDA kernels (sm_100a) + C ABI (include/b2_batch.h, include/mujoco/mujoco.h)
  oracle/liboracle.so               fp64 CPU oracle (test infrastructure only)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mujoco_sim_b200", "csrc")
LIB = os.path.join(ROOT, "mujoco_sim_b200", "lib", "libb2sim.so")
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("CXX", "g++")

LIB_SOURCES = ["batch.cu", "chain_f32.cu", "chain_f64.cu", "chain1_f32.cu", "chain1_f64.cu", "shim_step.cpp", "shim_host.cpp", "mjcf_compile.cpp", "urdf_import.cpp", "set0.cpp", "model_store.cpp", "mirror_post.cpp"]
ORACLE_SOURCES = ["oracle_smooth.cpp", "oracle_collision.cpp", "oracle_constraint.cpp", "oracle_top.cpp"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _headers(d):
    return [os.path.join(d, f) for f in os.listdir(d) if f.endswith((".h", ".cuh"))]


def _run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_lib(force=False, verbose_ptxas=False):
    """Compile every source to an object under build/ (incremental), then link the shared library."""
    hdrs = _headers(CSRC) + [os.path.join(ROOT, "include", "b2_batch.h"), os.path.join(ROOT, "include", "mujoco", "mujoco.h")]
    objdir = os.path.join(ROOT, "build", "obj")
    os.makedirs(objdir, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    objs, relink = [], force or not os.path.exists(LIB)
    jobs = []
    for s in LIB_SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            cmd = [NVCC, "-c", "-Xcompiler", "-fPIC", "-O3", "-std=c++17", "-lineinfo",
                   "-gencode", "arch=compute_100a,code=sm_100a",
                   "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-o", obj, src]
            if os.environ.get("B2_EXTRA_NVCC") and s.endswith(".cu"):
                cmd[1:1] = os.environ["B2_EXTRA_NVCC"].split()
            if verbose_ptxas and s.endswith(".cu"):
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)
    if jobs:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(_run, jobs))
        relink = True
    if relink or _newer(LIB, objs):
        _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs)
    return LIB


def build_oracle(force=False):
    srcs = [os.path.join(ORACLE_DIR, s) for s in ORACLE_SOURCES]
    deps = srcs + _headers(ORACLE_DIR) + [os.path.join(ROOT, "include", "mujoco", "mujoco.h")]
    if not force and not _newer(ORACLE_LIB, deps):
        return ORACLE_LIB
    cmd = [CXX, "-shared", "-fPIC", "-O3", "-march=x86-64-v3", "-pthread", "-std=c++17",
           "-I" + os.path.join(ROOT, "include"), "-I" + ORACLE_DIR, "-o", ORACLE_LIB] + srcs
    _run(cmd)
    return ORACLE_LIB


if __name__ == "__main__":
    force = "--force" in sys.argv
    build_lib(force, verbose_ptxas="-v" in sys.argv)
    build_oracle(force)
