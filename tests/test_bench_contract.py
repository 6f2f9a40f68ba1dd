"""The driver-facing contract of `bench.py --impl reference` (the CPU arm; runs without a GPU): one JSON line
with the keys the round-end driver reads, rank 0 only under a multi-rank launch."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, MRA_BENCH_SAMPLE_GRID="60", **env_extra)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout.strip()


def test_reference_arm_prints_one_contract_line():
    out = _run({"OMP_NUM_THREADS": "1"})            # what torchrun exports; the arm must lift it
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("getLikelihood() evals/sec") and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["config"]["workload"] == "cfg5" and d["config"]["n_locs"] == 4000000
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == os.cpu_count() and "sample" in cb
    modes = cb["modes"]
    assert modes["serial"]["evals_per_s"] > 0
    assert modes["serial"]["blas_threads"] >= 1
    assert "error" not in modes["fork_critDepth0"] and modes["fork_critDepth0"]["processes"] == 4     # real forks
    assert "error" not in modes["side_by_side"] and modes["side_by_side"]["processes"] >= 1
    assert d["value"] == max(m["evals_per_s"] for m in modes.values())
    # the line says what it is: an extrapolated estimate from a smaller sample, with consistent per-step time
    assert d["same_config"] is False and d["extrapolated"] is True and d["sample_grid"] == [60, 60]
    assert abs(d["ms_per_step"] - 1e3 / d["value"]) < 1e-6 * d["ms_per_step"]


def test_bench_and_parity_inputs_are_the_same_field():
    import numpy as np
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench
    from _util import FULLSIZE_CASES, fullsize_inputs
    a = bench.make_inputs(40, 0.4)
    b = fullsize_inputs(40, 0.4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1], equal_nan=True)
    for k in ("cfg3", "cfg4", "cfg5"):
        assert bench.WORKLOADS[k] == FULLSIZE_CASES[k]


def test_reference_arm_other_ranks_exit_quietly():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out == ""
