"""CPU: the parts of bench.py that run without a GPU — the reference arm (`--impl reference`: the unmodified reference
through oracle/_ref on the host cores) prints one JSON line with the contract's keys, and our arm refuses to run without
the CUDA library's device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFLIB = os.path.join(ROOT, "oracle", "_ref", "libmorpho_refharness.so")


@pytest.mark.skipif(not os.path.exists(REFLIB), reason="oracle/_ref (reference build) is not present")
def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--freq", "24", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "area_volume_gradient_throughput" and d["unit"] == "Melem/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Melem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "icosphere f=24" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--freq", "24"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == "", (p.stdout, p.stderr[-500:])


def test_our_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--freq", "8", "--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-configs"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode != 0
    assert not any(l.startswith("{") for l in p.stdout.splitlines())        # no number from a CPU path
