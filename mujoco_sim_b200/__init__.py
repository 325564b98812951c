"""mujoco_sim_b200 — B200-native batched rigid-body step behind the MuJoCo-named C API used by HoangGiang93/mujoco_sim.

The product is the native library (mujoco_sim_b200/lib/libb2sim.so: CUDA kernels for sm_100a + the extern "C" ABI of
include/b2_batch.h and include/mujoco/mujoco.h).  This Python package is a thin ctypes mirror of that ABI used by the
tests and the benchmark; it contains no physics.
"""
from .engine import Batch, Model, Data, lib, lib_path, asset, B2Error, data_contacts  # noqa: F401
from . import engine  # noqa: F401
