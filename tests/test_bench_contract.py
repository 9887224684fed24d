"""bench.py's reference arm (the CPU port on the host cores) prints one JSON line with the contract's keys; the GPU arm has
no CPU fallback."""

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    # (a 256-scenario cut of the batch keeps the CPU suite short; the driver runs the whole 4096 x 64 every step)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--scenarios", "256"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "participant_steps_per_sec" and d["unit"] == "participant-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("C2 256x64") and d["config"]["scenarios_per_gpu"] == 256 and d["config"]["participants"] == 64


def test_gpu_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from tactics2d_b200 import BatchedWorld, TypeTable

    with pytest.raises(RuntimeError, match="CUDA"):
        BatchedWorld(2, 2, TypeTable.vehicles())
