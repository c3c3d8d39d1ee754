"""CPU-side checks of bench.py's contract (no GPU): the reference arm prints ONE
JSON line with the agreed keys, ranks other than 0 stay silent, and our own arm
refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=e,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _bench("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "particles/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1
    assert d["config"]["workload"].startswith("periodic box, 1e8 uniform particles, 1024^3 mesh, TSC")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert "scaled twin" in cb["sample"] or "full workload" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "particles/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_nothing():
    r = _bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
               env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_own_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _bench("--steps", "1", "--warmup", "1")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
    assert not any(ln.startswith("{") for ln in r.stdout.splitlines())
