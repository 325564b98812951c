import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # build the native libraries once per session (incremental; a no-op when they are up to date)
    from mujoco_sim_b200 import build
    if os.environ.get("B2_SKIP_BUILD") != "1" and os.path.exists("/usr/local/cuda/bin/nvcc"):
        try:
            build.build_lib()
            build.build_oracle()
        except Exception as e:  # the prebuilt .so files may still be usable (GPU box)
            print("build failed:", e)


@pytest.fixture(scope="session")
def b2():
    import mujoco_sim_b200
    return mujoco_sim_b200


@pytest.fixture(scope="session")
def orc():
    from oracle import pyoracle
    return pyoracle


def has_gpu():
    import mujoco_sim_b200
    return mujoco_sim_b200.lib.b2_device_count() > 0
