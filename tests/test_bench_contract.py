"""bench.py contract pieces that run without a GPU: the reference arm's JSON line, and the product arm failing loudly when
there is no CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, env=e)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = run("--impl", "reference", "--steps", "3", "--warmup", "3", "--nenv", "256")
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["steps"] == 3 and d["warmup"] == 3 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = run("--impl", "reference", "--steps", "3", "--gpus", "2", env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_fails_loudly_without_a_gpu():
    import mujoco_sim_b200 as b2
    if b2.lib.b2_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = run("--steps", "2")
    assert r.returncode != 0 and "no CUDA device" in (r.stdout + r.stderr)
