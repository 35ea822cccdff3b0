"""bench.py's contract, as far as a machine without a GPU can check it: the reference arm (the unmodified reference's CPU
`hist`, oracle/_ref) prints ONE JSON line with the keys the driver reads; non-zero ranks of a torchrun launch leave
without work; our arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _env(**kw):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "B200_BENCH_REEXEC")}
    env.update(kw)
    return env


def _reference_available():
    sys.path.insert(0, ROOT)
    from oracle import ref_capi as rc
    return rc.available("cpu")


@pytest.mark.skipif(not _reference_available(), reason="oracle/_ref/libxgboost.so not built (needs /root/reference)")
def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1",
                        "--ref-rows", "60000", "--features", "32", "--depth", "4"],
                       capture_output=True, text=True, timeout=600, env=_env(), cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "iters/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["timed_rows"] == 60000 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_on_a_non_zero_rank_leaves_without_work():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, env=_env(RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"),
                       cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_cuda_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                       timeout=300, env=_env(), cwd=ROOT)
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
