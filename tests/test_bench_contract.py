"""bench.py without a GPU: the reference arm (`--impl reference`, the reference's own CPU matcher on a bounded sample)
prints one JSON line with the contract's keys; the CUDA arm refuses to run (there is no CPU path to fall back to)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT, have_gpu


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                          cwd=ROOT, timeout=600)


def test_reference_arm_prints_the_contract_line(built):
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "64"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gallery matches/sec per latent" and d["unit"] == "matches/s"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "64" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "matches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_on_rank_zero_only(built):
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-sample", "64"],
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cuda_arm_refuses_to_run_without_a_device(built):
    if have_gpu():
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)
