"""bench.py's contract where it can be checked without a GPU: the reference arm (the CPU port of the
reference's loops, timed on the host) prints one JSON line with the keys the driver reads, and the
product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    res = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--entities", "4000",
                    "--cpu-sample-entities", "1000")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    out = json.loads(lines[0])
    assert out["impl"] == "reference" and out["n_gpus"] == 1 and out["steps"] == 2 and out["warmup"] == 1
    assert out["metric"] == "vi_iterations_per_sec_at_10M_ground_factors" and out["unit"] == "it/s"
    assert out["higher_is_better"] is True and out["vs_baseline"] is None and out["data"] == "synthetic"
    assert out["value"] > 0 and abs(out["value"] * out["ms_per_step"] / 1000.0 - 1.0) < 1e-6
    assert "workload" in out["config"] and "model" not in out["config"]
    cb = out["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == out["value"] and cb["sample"]
    assert out["e2e"] == {"value": out["value"], "unit": out["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is visible")
    res = run_bench("--steps", "1", "--warmup", "1", "--entities", "2000", "--no-cpu-baseline")
    assert res.returncode != 0
    assert "no CPU fallback" in (res.stdout + res.stderr)
    assert not [ln for ln in res.stdout.splitlines() if ln.startswith("{")]


def test_c2f_figure_of_the_bench_line(monkeypatch):
    """`bench.c2f_probe` (the auxiliary coarse-to-fine figure) with the numpy oracle standing in for
    the device: keys, consistency of the phase times."""
    import importlib.util
    import types
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import lhvi_b200
    from oracle_engine import OracleEngine
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    monkeypatch.setattr(lhvi_b200.lifting.C2FArrayVI, "_make_engine",
                        lambda self, model: OracleEngine(model, var_threshold=0.1))
    a = types.SimpleNamespace(entities=3000, groups=3, K=2, T=3, dtype="float64")
    out = bench.c2f_probe(a, iterations=20)
    assert out["rounds"] == 2 and out["iterations"] == 20 and out["free_energy_finite"]
    assert out["ground_factors"] == 1000 * 3 + 1000 + 3 * 2 or out["ground_factors"] > 3000
    assert out["classes_per_round"] == sorted(out["classes_per_round"])
    parts = out["lifting_passes_s"] + out["upload_s"] + out["device_iterations_s"] + out["readback_s"] + out["setup_s"]
    assert 0 < parts <= out["run_s"] * 1.001
    assert out["lifting_passes"].startswith("host")          # no GPU here: the passes run in the host library
