"""bench.py's contract as far as it can be checked without a GPU: the reference arm prints exactly one JSON
line on stdout with the keys the driver reads, the GPU arm refuses to run without CUDA (no CPU fallback),
and the roofline inputs (algorithmic bytes, measured DRAM traffic) are where bench.py expects them."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True,
                          text=True, cwd=ROOT, env=e, timeout=600)


def test_reference_arm_prints_one_json_line(oracle):
    r = _run("--impl", "reference", "--workload", "c2", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sediment cell-updates/sec"
    assert d["unit"] == "cell-updates/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "slab" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("C2")


def test_reference_arm_other_ranks_do_nothing(oracle):
    r = _run("--impl", "reference", "--workload", "c2", "--steps", "1", "--warmup", "0", "--gpus", "2",
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)


def test_roofline_inputs():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.b_alg(40) == pytest.approx(141.4) and bench.b_alg(30) == pytest.approx(143.2)
    single, pair = bench.traffic_per_cell("column_kernel"), bench.traffic_per_cell("pair_kernel")
    assert 100.0 < single <= 141.4          # no re-reads: measured DRAM bytes <= algorithmic
    assert pair == pytest.approx(single / 2, rel=0.02)   # two steps per HBM round trip
    peak, src = bench.measured_peak()
    assert peak > 1000.0 and isinstance(src, str)
    for name, wl in bench.WORKLOADS.items():
        assert len(wl) == 7 and wl[2] <= 64
    assert bench.SEGMENT % int(round(bench.COUPLING_SECONDS / bench.DT)) == 0
