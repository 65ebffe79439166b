"""GPU tests of the reference-facing boundary: the get_cuda_model plugin called through the
reference's own ModelFunctions / MonteCarloParams / MonteCarloResult structs, and the reference's
UNMODIFIED main_runner.c driving it as a drop-in next to the C-serial model."""
import csv
import ctypes as C
import os
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle_lib
from mcr_b200 import api

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "gpu-montecarlo-risk_b200"
GOLD = ROOT / "tests" / "golden"


@pytest.fixture
def plugin():
    so = PKG / "libmcr_b200_model.so"
    assert so.exists(), "libmcr_b200_model.so must be prebuilt and shipped to the GPU box"
    return C.CDLL(str(so))


def test_plugin_matches_engine_and_oracle(plugin, engine, oracle, monkeypatch):
    g = np.load(GOLD / "pyref_n15.npz")
    mu, Sigma = g["mu"], g["Sigma"]
    M, seed = 250000, 67
    for mode_name, mode in (("lower", api.TRANSFORM_LOWER), ("gsl_full", api.TRANSFORM_GSL_FULL)):
        monkeypatch.setenv("MCR_TRANSFORM", mode_name)
        monkeypatch.setenv("MCR_S_VALUES", "1")
        r = oracle_lib.run_model_struct(plugin.get_cuda_model, mu, Sigma, 5, 0.02, M, seed)
        assert r["rc"] == 0 and r["name"] == "cuda"
        A = engine.upload_sigma(mu, Sigma, mode)
        Z, _ = engine.dump_normals(M, seed)
        ro = oracle.run_normals(A, mu, 5, 0.02, Z)
        np.testing.assert_array_equal(r["S_values"], ro["S_values"])
        assert r["count"] == ro["count"]
        assert (r["P_hat"], r["std_error"], r["ci_lower"], r["ci_upper"]) == oracle.stats(ro["count"], M)
        monkeypatch.setenv("MCR_S_VALUES", "0")
        r2 = oracle_lib.run_model_struct(plugin.get_cuda_model, mu, Sigma, 5, 0.02, M, seed)
        assert r2["rc"] == 0 and r2["S_values"] is None and r2["count"] == r["count"]


def test_plugin_error_convention(plugin, capfd):
    g = np.load(GOLD / "pyref_n15.npz")
    r = oracle_lib.run_model_struct(plugin.get_cuda_model, g["mu"], -g["Sigma"], 5, 0.02, 1000, 67)
    assert r["rc"] == -1 and r["S_values"] is None
    assert "Cholesky decomposition failed" in capfd.readouterr().err


def test_reference_main_runner_drop_in(tmp_path):
    """The reference's unmodified main_runner.c (built -DSERIAL_BUILD -DOPENMP_BUILD -DCUDA_BUILD)
    selects the B200 model for `cuda` in config.yaml and writes its 22-column results row."""
    exe = PKG / "bin" / "monte_carlo"
    assert exe.exists(), "bin/monte_carlo must be prebuilt and shipped to the GPU box"
    (tmp_path / "data").mkdir()
    (tmp_path / "results").mkdir()
    for f in ("mu.bin", "sigma.bin", "params.bin"):
        shutil.copy(GOLD / "data15" / f, tmp_path / "data" / f)
    shutil.copy(GOLD / "config_dropin.yaml", tmp_path / "config.yaml")
    env = dict(os.environ, MCR_TRANSFORM="gsl_full", OMP_NUM_THREADS="2", MCR_TIMING_CSV=str(tmp_path / "results" / "b200_timing.csv"))
    p = subprocess.run([str(exe)], cwd=tmp_path, stdin=subprocess.DEVNULL, capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
    rows = list(csv.DictReader(open(tmp_path / "results" / "simulation_results.csv")))
    assert [r["model_name"] for r in rows] == ["cuda", "serial"]
    M = 300000
    pc, ps = float(rows[0]["P_hat"]), float(rows[1]["P_hat"])
    assert int(rows[0]["M"]) == M and int(rows[0]["count"]) == round(pc * M)
    se = np.sqrt(pc * (1 - pc) / M + ps * (1 - ps) / M)
    assert abs(pc - ps) <= 3 * se + 1e-6, (pc, ps, se)   # independent RNG streams, same transform matrix
    assert "CUDA Simulation complete" in p.stdout
    timing = list(csv.DictReader(open(tmp_path / "results" / "b200_timing.csv")))
    assert len(timing) == 1 and int(timing[0]["M"]) == M and float(timing[0]["device_ms"]) > 0
    assert int(timing[0]["count"]) == int(rows[0]["count"])


def test_cli_runs_from_preprocessing_binaries(engine):
    """mcr_cli: hardened loader (model_loader.c) -> mcr_build_transform -> mcr_simulate_host, no GSL involved."""
    exe = PKG / "bin" / "mcr_cli"
    assert exe.exists(), "bin/mcr_cli must be prebuilt and shipped to the GPU box"
    p = subprocess.run([str(exe), str(GOLD / "data15"), "500000", "0.02", "5", "67", "1", "gsl_full"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    g = np.load(GOLD / "pyref_n15.npz")
    engine.upload_sigma(g["mu"], g["Sigma"], api.TRANSFORM_GSL_FULL)
    r = engine.run(500000, 67, 0.02, 5)
    assert f"count={r.count} sum_S={r.sum_S} " in p.stdout
    bad = subprocess.run([str(exe), str(GOLD)], capture_output=True, text=True, timeout=60)
    assert bad.returncode == 1 and "Cannot open" in bad.stderr
