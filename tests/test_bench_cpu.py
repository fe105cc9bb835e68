"""CPU checks of bench.py's host logic: the reference arm (`--impl reference`, the CPU oracle port) must print one valid
JSON line with the contract's keys for every workload, and every workload's config() must be available before setup()."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("workload", ["xvector_train", "logmel"])
def test_reference_arm_prints_contract_line(workload):
    env = dict(os.environ, LBX_REF_BUDGET_S="0.5")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload,
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["gpu_launches"] == 0
    for key in ("metric", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in line["config"]


def test_workload_configs_do_not_need_setup():
    sys.path.insert(0, ROOT)
    import bench

    class Args:
        batch, seconds = 0, 0
    for cls in bench.WORKLOADS.values():
        json.dumps(cls(Args, 0, 1).config())
