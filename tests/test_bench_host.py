"""Host-side logic of bench.py that runs without a GPU: the CPU arm's extrapolation rule, the reference-arm line on a
tiny sample, the option parser.  (The GPU arms are exercised by the gpurun scripts under tools/gpurun_scripts/.)"""
import json
import math
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def sample(n, ms, setup):
    return {"n": n, "ms_per_selection": ms, "setup_inverse_s": setup}


def test_one_sample_uses_the_algorithmic_laws():
    ex = bench.cpu_extrapolate([sample(8192, 8.0, 5.0)], 50000, 20)
    r = 50000 / 8192
    assert ex["seconds_per_selection"] == pytest.approx(8e-3 * r ** 2)
    assert ex["setup_seconds"] == pytest.approx(5.0 * r ** 3)
    assert ex["fitted_exponents"] is None
    assert ex["whole_call_selections_per_s"] == pytest.approx(20 / (ex["setup_seconds"] + 20 * ex["seconds_per_selection"]))


def test_two_samples_keep_the_prediction_more_favourable_to_the_cpu():
    # the round-end measurements: step exponent 1.68, setup exponent 1.94 (LAPACK still gaining efficiency)
    s = [sample(8192, 8.108704349996287, 5.529435728999999), sample(16384, 25.967866749994073, 21.172488131000023)]
    ex = bench.cpu_extrapolate(s, 50000, 20, dgemm_gflops=1498.5)
    f = ex["fitted_exponents"]
    assert f["step_exponent"] == pytest.approx(1.679, abs=1e-3) and f["step_fit_used"]
    assert f["setup_exponent"] == pytest.approx(1.937, abs=1e-3) and f["setup_fit_used"]
    c = ex["candidates"]
    assert ex["seconds_per_selection"] == min(c["seconds_per_selection"].values())
    assert ex["setup_seconds"] == pytest.approx(c["setup_seconds"]["fitted"])
    assert ex["setup_seconds"] > c["setup_seconds"]["floor_n3_flop_at_host_dgemm_rate"]


def test_a_wild_fit_is_ignored_and_the_dgemm_floor_holds():
    s = [sample(1024, 1.0, 0.5), sample(2048, 1.05, 0.52)]          # thread-pool start-up noise: exponents near 0
    ex = bench.cpu_extrapolate(s, 50000, 20, dgemm_gflops=1000.0)
    f = ex["fitted_exponents"]
    assert not f["step_fit_used"] and not f["setup_fit_used"]
    r = 50000 / 2048
    assert ex["seconds_per_selection"] == pytest.approx(1.05e-3 * r ** 2)
    s = [sample(8192, 8.0, 1e-3), sample(16384, 32.0, 8e-3)]         # an impossibly fast inverse: floored
    ex = bench.cpu_extrapolate(s, 50000, 20, dgemm_gflops=1000.0)
    assert ex["setup_seconds"] == pytest.approx(50000.0 ** 3 / 1e12)


def test_option_flag_is_recorded_in_the_config():
    ns = type("A", (), {"n": 2000, "k": 5, "gpus": 1, "exchange": "peer", "option": ["gemm_emulate_slices=0"]})()
    cfg = bench.config_dict(ns)
    assert cfg["library_options"] == ["gemm_emulate_slices=0"]
    assert cfg["workload"].startswith("greedy_mi_placement_n2000_k5")
    ns.option = []
    assert "library_options" not in bench.config_dict(ns)


@pytest.mark.timeout(300)
def test_reference_arm_prints_one_json_line_on_a_small_sample():
    env = dict(os.environ, VGP_BENCH_CPU_N="512,1024")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "4000", "--steps", "4",
                          "--warmup", "1"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=280)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["extrapolated"] is True and d["gpu_launches"] == 0
    assert d["metric"] == "greedy_mi_selections_per_s" and d["unit"] == "selections/s" and d["higher_is_better"] is True
    assert d["config"]["n"] == 4000 and d["config"]["k"] == 4
    assert [m["n"] for m in d["measured_at"]] == [512, 1024]
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert math.isfinite(d["value"]) and d["value"] > 0 and d["e2e"]["value"] > 0
