"""bench.py's driver contract, as far as it can be checked without a GPU: the reference arm's JSON line (the reference's
own NumPy step() from the staged notebooks in oracle/_ref), its behaviour under a multi-rank launch, and that the
product arm refuses to run without the CUDA path instead of falling back to anything on the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def run_bench(*args, env=None, timeout=300):
    e = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        e.pop(k, None)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH, *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=e)


def test_reference_arm_line_has_the_contract_keys():
    r = run_bench("--impl", "reference", "--num-envs", "4096", "--steps", "3", "--warmup", "1", "--cpu-seconds", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "quadrotor env-steps/sec" and d["unit"] == "env-steps/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("e2e_zigzag_ga1_N4096") and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # env-steps/s and ms per step of N envs describe the same measurement
    assert d["value"] == pytest.approx(4096 / (d["ms_per_step"] * 1e-3), rel=0.35)
    if os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "reference")):  # staged by build(): then it is the real thing
        assert cb["kind"] == "reference" and "NumPy" in cb["what"]


def test_reference_arm_other_ranks_exit_quietly():
    r = run_bench("--impl", "reference", "--gpus", "2", "--num-envs", "4096", "--steps", "2", "--warmup", "1",
                  env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29533"},
                  timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    r = run_bench("--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--no-configs", timeout=180)
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")], "no bench line may come from a box without the CUDA path"
    assert "needs a gpu" in (r.stderr + r.stdout).lower()
