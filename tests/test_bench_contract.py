"""CPU: bench.py's contract pieces that do not need a GPU -- the reference arm prints one JSON
line with the agreed keys (rank 0 only), and the B200 arm refuses to run without a CUDA device
(no CPU fallback)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _run(args, env_extra=None, timeout=600):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, env=env)


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--workload", "100Kx384", "--steps", "1", "--warmup", "1", "--batch", "64"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["metric"] == "queries/sec exact top-10 cosine" and d["value"] > 0
    assert d["config"]["workload"].startswith("100Kx384") and d["config"]["batch"] == 64
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--workload", "100Kx384", "--steps", "1", "--warmup", "1", "--gpus", "2"],
             env_extra={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_b200_arm_fails_loudly_without_a_gpu():
    r = _run(["--steps", "1", "--warmup", "1"])
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
