"""The oracle against every pinned vector we have (no GPU):
 - MT19937 raw outputs from numpy's legacy seeding (tests/golden/mt19937_kat.json),
 - the reference's own Python model run on recorded normals (tests/golden/pyref_*.npz),
 - Random123's published Philox4x32-10 known-answer vectors,
 - the reference's unmodified C models compiled against the GSL shim (oracle/_ref), when built.
"""
import json
from pathlib import Path

import numpy as np
import pytest

import oracle_lib

GOLD = Path(__file__).parent / "golden"


def test_mt19937_known_answers(oracle):
    kat = json.load(open(GOLD / "mt19937_kat.json"))
    for seed, v in kat.items():
        raw = oracle.mt19937_raw(int(seed), 1000)
        assert [int(t) for t in raw[:8]] == v["first8"]
        assert int(raw[999]) == v["n1000"]
    # the classic MT19937 check value (also GSL's own rng test): seed 4357, 1000th output
    assert int(oracle.mt19937_raw(4357, 1000)[999]) == 1186927261
    # gsl_rng_set(r, 0) uses the default seed 4357
    assert (oracle.mt19937_raw(0, 16) == oracle.mt19937_raw(4357, 16)).all()


def test_polar_gaussian_is_standard_normal(oracle):
    z = oracle.mt19937_gaussians(67, 400000)
    assert abs(z.mean()) < 4 / np.sqrt(z.size)
    assert abs(z.var() - 1.0) < 0.01
    assert abs((z ** 4).mean() - 3.0) < 0.06
    # polar acceptance is pi/4: the generator consumes 2/(pi/4) uniforms per normal on average
    np.testing.assert_allclose(z[:4], [1.6118409272803023, 0.9836920427306165, 0.942148531578204, -1.0465246247230229], rtol=0, atol=0)


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32 10 rounds
    assert [hex(v) for v in oracle.philox([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    f = 0xFFFFFFFF
    assert [hex(v) for v in oracle.philox([f, f, f, f], [f, f])] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(v) for v in oracle.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_philox_counter_layout(oracle):
    U = oracle.philox_uniforms(seed=(5 << 32) | 9, path_offset=(1 << 32) + 3, M=2, N=6)
    assert U.shape == (2, 8)
    np.testing.assert_array_equal(U[1, 4:8], oracle.philox([4, 1, 1, 0], [9, 5]))
    Z = oracle.philox_normals(67, 0, 200000, 4)
    assert abs(Z.mean()) < 5 / np.sqrt(Z.size) and abs(Z.var() - 1) < 0.01


@pytest.mark.parametrize("name", ["pyref_n15.npz", "pyref_n25.npz"])
def test_oracle_matches_python_reference_on_identical_normals(oracle, name):
    g = np.load(GOLD / name)
    A = oracle.transform(g["Sigma"], oracle_lib.LOWER)  # python_serial.py:40 uses np.linalg.cholesky (lower)
    np.testing.assert_allclose(np.tril(A), np.linalg.cholesky(g["Sigma"]), rtol=1e-12, atol=1e-18)
    r = oracle.run_normals(A, g["mu"], int(g["k"]), float(g["x"]), g["Z"])
    np.testing.assert_array_equal(r["S_values"], g["S_values"])
    assert r["count"] == int(g["count"])
    P, se, lo, hi = oracle.stats(r["count"], int(g["M"]))
    assert P == float(g["P_hat"]) and se == float(g["std_error"])
    assert lo == float(g["ci_lower"]) and hi == float(g["ci_upper"])


def test_stats_formulas(oracle):
    P, se, lo, hi = oracle.stats(4001326, 100000000)  # the reference's published serial row (results/a.csv:3)
    assert f"{P:.6f}" == "0.040013" and f"{se:.6f}" == "0.000020"
    assert f"{lo:.6f}" == "0.039975" and f"{hi:.6f}" == "0.040052"
    assert oracle.stats(0, 10) == (0.0, 0.0, 0.0, 0.0)
    assert oracle.stats(10, 10)[2:] == (1.0, 1.0)


def test_transform_conventions(oracle):
    g = np.load(GOLD / "pyref_n15.npz")
    S = g["Sigma"]
    L = oracle.transform(S, oracle_lib.LOWER)
    F = oracle.transform(S, oracle_lib.GSL_FULL)
    np.testing.assert_array_equal(np.tril(L), np.tril(F))
    assert (np.triu(L, 1) == 0).all()
    np.testing.assert_array_equal(np.triu(F, 1), np.triu(S, 1))
    np.testing.assert_allclose(np.tril(L) @ np.tril(L).T, S, rtol=1e-12)
    with pytest.raises(ValueError):
        oracle.transform(-S, oracle_lib.LOWER)


def test_real_data_goldens(oracle):
    """Shipped 15-index data, config.yaml parameters, M = 1e6 (BASELINE config 1 scale-down is M itself)."""
    g = np.load(GOLD / "pyref_n15.npz")
    F = oracle.transform(g["Sigma"], oracle_lib.GSL_FULL)
    L = oracle.transform(g["Sigma"], oracle_lib.LOWER)
    rf = oracle.run_mt19937(F, g["mu"], 5, 0.02, 1000000, 67)
    rl = oracle.run_mt19937(L, g["mu"], 5, 0.02, 1000000, 67, want_S=False)
    assert (rf["count"], rf["sum_S"]) == (47605, 674239)
    assert (rl["count"], rl["sum_S"]) == (47283, 673423)
    assert list(rf["S_values"][:8]) == [0, 0, 3, 0, 3, 0, 0, 0]
    assert int((rf["S_values"] >= 5).sum()) == rf["count"]


def test_goldens_do_not_depend_on_the_cholesky_diagonal_rounding(oracle):
    """GSL is not installed here, and its Cholesky is restated from recall: the unblocked kernel scales the whole column
    A(j:n, j), diagonal included, by 1/sqrt(a_jj) (what the shim, the oracle and mcr_build_transform do); a build that stored
    L_jj = sqrt(a_jj) instead would differ by at most one ulp on the diagonal.  Per-trial S over the whole C1 run is the same
    either way, so the frozen counts are not an artefact of that choice."""
    g = np.load(GOLD / "pyref_n15.npz")
    differ = 0
    for mode in (oracle_lib.GSL_FULL, oracle_lib.LOWER):
        A0 = oracle.transform(g["Sigma"], mode)
        A1 = oracle.transform(g["Sigma"], mode | oracle_lib.DIAG_SQRT)
        off = ~np.eye(15, dtype=bool)
        assert np.array_equal(A0[off], A1[off])
        d0, d1 = np.diag(A0), np.diag(A1)
        assert np.all(np.abs(d0 - d1) <= np.spacing(d1))
        differ += int((d0 != d1).sum())
        r0 = oracle.run_mt19937(A0, g["mu"], 5, 0.02, 1000000, 67)
        r1 = oracle.run_mt19937(A1, g["mu"], 5, 0.02, 1000000, 67)
        np.testing.assert_array_equal(r0["S_values"], r1["S_values"])
    print(f"diagonal entries that differ by one ulp between the two roundings: {differ} of 30")


@pytest.mark.skipif(not oracle_lib.ref_models_path().exists(), reason="oracle/_ref not built (reference tree absent)")
def test_oracle_matches_compiled_reference_models(oracle):
    """The reference's monte_carlo_serial.c / monte_carlo_omp.c, compiled unmodified against the GSL shim."""
    ref = oracle_lib.RefModels()
    g = np.load(GOLD / "pyref_n15.npz")
    M = 60000
    rs = ref.serial(g["mu"], g["Sigma"], 5, 0.02, M, 67)
    assert rs["rc"] == 0 and rs["name"] == "serial"
    F = oracle.transform(g["Sigma"], oracle_lib.GSL_FULL)
    ro = oracle.run_mt19937(F, g["mu"], 5, 0.02, M, 67)
    np.testing.assert_array_equal(rs["S_values"], ro["S_values"])
    assert rs["count"] == ro["count"]
    assert (rs["P_hat"], rs["std_error"], rs["ci_lower"], rs["ci_upper"]) == oracle.stats(ro["count"], M)
    # OpenMP: same multiset of trials (one shared RNG under omp critical, omp.c:77-82) => same count
    rm = ref.openmp(g["mu"], g["Sigma"], 5, 0.02, M, 67)
    assert rm["rc"] == 0 and rm["count"] == rs["count"]
    assert sorted(rm["S_values"]) == sorted(rs["S_values"])
