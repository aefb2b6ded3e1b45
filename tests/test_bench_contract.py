"""The reference arm of bench.py runs on CPU: check the JSON-line contract on a small system."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*extra, env=None):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cells", "16", *extra], capture_output=True, text=True, timeout=300,
                       cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    return [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_line():
    lines = run()
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["metric"] == "PCG iterations/sec" and d["unit"] == "iter/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "BASELINE configs[1]" in d["config"]["workload"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    assert run("--gpus", "2", env=env) == []


def test_both_arms_build_the_same_config_dict():
    """The driver compares the two arms' `config`: it comes from one function fed with the rank-0
    block of the same decomposition, and names the workload BASELINE.json is quoted on."""
    sys.path.insert(0, ROOT)
    import bench
    for n_gpus in (1, 2, 8):
        s = bench.build_rank_system(6, n_gpus, 0)
        c = bench.common_config(6, n_gpus, s)
        assert c["rows_per_gpu"] == 216 and c["cells_global"] == 216 * n_gpus
        assert c["decomposition"] == list(bench.procs_for(n_gpus))
        assert c["halo_per_gpu"] == {1: 0, 2: 36, 8: 108}[n_gpus]
        assert set(c) == {"workload", "rows_per_gpu", "nnz_per_gpu", "halo_per_gpu", "cells_global",
                          "decomposition", "value_definition", "l2"}
    assert "BASELINE configs[4]" in bench.workload_text(200, 8) and "64 M cells" in bench.workload_text(200, 8)
    # algorithmic bytes at 200^3
    n, nnz = 8_000_000, 55_760_000
    assert bench.alg_bytes_spmv(n, nnz) == 12 * nnz + 4 * (n + 1) + 16 * n == 829_120_004
    assert bench.alg_bytes_pcg(n, nnz) == 1_469_120_004
    assert bench.alg_bytes_bicgstab(n, nnz) == 2 * (12 * nnz + 4 * (n + 1)) + 200 * n
    # the oracle's pinned iteration counts exist for every default workload of the scaling run
    for k in ("pressure_200_x1", "pressure_200_x2", "pressure_200_x4", "pressure_200_x8", "pressure_100_x1"):
        assert bench.expected(k)["iterations"] > 100


import pytest  # noqa: E402


@pytest.mark.gpu
def test_gpu_arm_line_at_a_small_size():
    """The GPU arm end to end on a small system (72^3 per GPU: the kernels of the default run): one JSON
    line with the contract's keys, a passing in-line check, e2e bytes counted, launches counted."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--cells", "72", "--steps", "2",
                        "--warmup", "3", "--no-extra"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = lines[0]
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks",
                "check"):
        assert key in d, key
    assert d["metric"] == "PCG iterations/sec" and d["dtype"] == "f64" and d["n_gpus"] == 1
    assert d["check"]["ok"] is True and d["check"]["true_residual"] < 1e-6
    assert d["gpu_launches"] > 0 and d["value"] > 0 and 0 < d["e2e"]["value"] < d["value"]
    n, f = 72 ** 3, 3 * 72 * 72 * 71
    assert d["e2e"]["h2d_bytes_per_step"] == 8 * f + 24 * n and d["e2e"]["d2h_bytes_per_step"] == 8 * n
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert "k_spmv_ell" in r["kernel"] and d["cpu_baseline"]["kind"] == "port"
