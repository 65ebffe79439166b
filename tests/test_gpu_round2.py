"""GPU parity tests added in round 2 (run on a B200 with -m gpu; everything goes through the C ABI).

  * fast kernels against the fp64 kernel at FULL BASELINE sizes (no oracle can follow there: the fp64 GPU kernel evaluates
    the reference's summation order on the same Philox normals, and is itself pinned to the oracle at small sizes);
  * models in other units (Sigma x 1e4 = basis points, Sigma x 1e-6): the per-row power-of-two scale of the fp16 operand
    images keeps every tensor kernel exact;
  * k <= 0, runs on foreign streams, the plugin's single-process multi-GPU route (MCR_NUM_GPUS);
  * the reference's own CUDA model (monte_carlos_cuda.cu compiled unmodified for sm_100a) as an independent-stream check.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np
import pytest

import oracle_lib
from mcr_b200 import api
from mcr_b200.synthetic import synthetic_model

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "gpu-montecarlo-risk_b200"
GOLD = ROOT / "tests" / "golden"


@pytest.mark.parametrize("N,M", [(50, 100_000_000), (100, 30_000_000), (600, 20_000_000)])
def test_fast_kernels_equal_the_fp64_kernel_at_full_size(engine, N, M):
    """AUTO (tensor-core kernel + guard band + fp64 fix-up) == MCR_KERNEL_EXACT on (count, sum_S) over the full BASELINE
    trial counts: 1e8 trials at N=50 (config 2), 2e7 at N=600 (config 4 shape), i.e. 5e9 / 1.2e10 per-asset decisions."""
    m = synthetic_model(N)
    engine.upload_sigma(m.mu, m.Sigma)
    fast = engine.run(M, 67, m.x, m.k)
    assert fast.kernel == "tensor"
    exact = engine.run(M, 67, m.x, m.k, kernel=api.KERNEL_EXACT)
    assert exact.kernel == "exact"
    assert (fast.count, fast.sum_S, fast.n_paths) == (exact.count, exact.sum_S, M)
    assert 0 < fast.n_rechecked < M // 20


@pytest.mark.parametrize("N,M", [(50, 200_000), (100, 100_000), (128, 40_000), (160, 20_000), (320, 12_000), (600, 6_000)])
@pytest.mark.parametrize("unit", [1e4, 1e-6, 1.0])
def test_models_in_other_units_stay_exact(engine, oracle, N, M, unit):
    """Returns in basis points (Sigma x 1e8, mu and x x 1e4) or tiny units (x 1e-6): the same events, so the per-trial S must
    not change, whatever kernel AUTO picks.  A fixed operand scale overflows fp16 / drops into its subnormals here."""
    m = synthetic_model(N)
    A0 = engine.upload_sigma(m.mu, m.Sigma)
    Z, _ = engine.dump_normals(M, 67)
    base = oracle.run_normals(A0, m.mu, m.k, m.x, Z)
    A = engine.upload_sigma(m.mu * unit, m.Sigma * unit * unit)
    ro = oracle.run_normals(A, m.mu * unit, m.k, m.x * unit, Z)
    h = api.simulate_host(m.mu * unit, A, M, 67, m.x * unit, m.k, a_is_lower=True, want_S_values=True)
    assert h.kernel == "tensor"
    np.testing.assert_array_equal(h.S_values, ro["S_values"])
    assert (h.count, h.sum_S) == (ro["count"], ro["sum_S"])
    # and the events are the same as in the original units up to fp64 rounding of the rescaled model
    assert abs(ro["sum_S"] - base["sum_S"]) <= 2 + base["sum_S"] // 100000


def test_rows_of_very_different_magnitude(engine, oracle):
    """One asset quoted in basis points next to assets quoted as fractions: rows of the transform 1e4 apart."""
    m = synthetic_model(64)
    s = np.ones(64)
    s[5], s[40], s[63] = 1e4, 1e-3, 3e2
    mu, Sigma = m.mu * s, m.Sigma * np.outer(s, s)
    # thresholds are per-model (one x): use x in the units of the unscaled assets; the scaled ones then (almost) never / always hit
    A = engine.upload_sigma(mu, Sigma)
    M = 200_000
    Z, _ = engine.dump_normals(M, 67)
    ro = oracle.run_normals(A, mu, m.k, m.x, Z)
    for kernel in (api.KERNEL_TENSOR, api.KERNEL_FFMA, api.KERNEL_AUTO):
        r = engine.run(M, 67, m.x, m.k, kernel=kernel)
        assert (r.count, r.sum_S) == (ro["count"], ro["sum_S"]), r.kernel


@pytest.mark.parametrize("N,mode,M", [(129, api.TRANSFORM_LOWER, 20_000), (144, api.TRANSFORM_GSL_FULL, 12_000), (256, api.TRANSFORM_LOWER, 12_000),
                                      (383, api.TRANSFORM_LOWER, 8_000), (385, api.TRANSFORM_LOWER, 8_000), (400, api.TRANSFORM_GSL_FULL, 5_000),
                                      (600, api.TRANSFORM_LOWER, 148 * 128 * 2 + 133), (768, api.TRANSFORM_LOWER, 4_000),
                                      (1000, api.TRANSFORM_GSL_FULL, 2_500), (1024, api.TRANSFORM_LOWER, 3_000)])
def test_streamed_kernel_every_pass_shape(engine, oracle, N, mode, M):
    """mcr_stream.cu: one pass (N <= 384), two (N = 385 is 16 + 384 rows) and three (N > 768) passes, ragged last K-block and
    last row group, lower-triangular (K-blocks skip the rows above them) and dense transforms, several tiles per CTA plus a
    ragged one at N = 600.  Per-trial S identical to the fp64 oracle on the dumped normals."""
    m = synthetic_model(N)
    A = engine.upload_sigma(m.mu, m.Sigma, mode)
    Z, _ = engine.dump_normals(M, 19)
    ro = oracle.run_normals(A, m.mu, m.k, m.x, Z)
    h = api.simulate_host(m.mu, A, M, 19, m.x, m.k, a_is_lower=(mode == api.TRANSFORM_LOWER), want_S_values=True)
    assert h.kernel == "tensor" and h.padded_N == (N + 15) // 16 * 16
    np.testing.assert_array_equal(h.S_values, ro["S_values"])
    assert (h.count, h.sum_S) == (ro["count"], ro["sum_S"])
    assert h.n_rechecked < M // 5


@pytest.mark.parametrize("N,mode,M", [(160, api.TRANSFORM_LOWER, 6_000), (600, api.TRANSFORM_LOWER, 3_000), (600, api.TRANSFORM_GSL_FULL, 1_500),
                                      (1024, api.TRANSFORM_LOWER, 1_000)])
def test_supplied_normals_wide_models(engine, oracle, N, mode, M):
    """The C-serial model's own normal stream (MT19937 + polar method, seed 67) fed through mcr_run_with_normals to the
    tensor path at Stoxx600-class sizes (round 1: fp64 kernel only)."""
    m = synthetic_model(N)
    A = engine.upload_sigma(m.mu, m.Sigma, mode)
    ro = oracle.run_mt19937(A, m.mu, m.k, m.x, M, 67, want_Z=True)
    for kernel in (api.KERNEL_TENSOR, api.KERNEL_AUTO, api.KERNEL_EXACT):
        r = engine.run_with_normals(ro["Z"], m.x, m.k, kernel=kernel)
        assert r.kernel == ("exact" if kernel == api.KERNEL_EXACT else "tensor")
        np.testing.assert_array_equal(r.S_values, ro["S_values"])
        assert (r.count, r.sum_S, r.n_paths) == (ro["count"], ro["sum_S"], M)


@pytest.mark.parametrize("N,mode,M,weights", [(128, api.TRANSFORM_LOWER, 300_000, None), (200, api.TRANSFORM_GSL_FULL, 200_000, "ramp"),
                                              (600, api.TRANSFORM_LOWER, 120_000, "ramp")])
def test_portfolio_outputs_for_wide_models(engine, oracle, N, mode, M, weights):
    """North-star extension A7 (parity unpinned: the reference has no portfolio / VaR / ES code; checked against the oracle's own
    definition on the dumped normals) beyond the tile kernel: the streamed kernel carries w.R as one more image row.
    Stated tolerances as for N <= 126: moments 1e-4 relative (fp16x3 products, fp32 per-thread partial sums); hits may differ by
    the trials whose P sits within 1e-6 of the threshold; VaR within one histogram bin (sigma_p / 128); ES within 0.2 %."""
    m = synthetic_model(N)
    A = engine.upload_sigma(m.mu, m.Sigma, mode)
    w = np.full(N, 1.0 / N) if weights is None else np.arange(1, N + 1) / np.arange(1, N + 1).sum()
    xp, seed = 0.012, 31
    r = engine.run(M, seed, m.x, m.k, portfolio_threshold=xp, weights=w)
    assert r.kernel == "tensor"
    Z, _ = engine.dump_normals(M, seed)
    ro = oracle.run_normals(A, m.mu, m.k, m.x, Z)
    assert (r.count, r.sum_S) == (ro["count"], ro["sum_S"])       # the base outputs are untouched and exact
    P = oracle.portfolio(A, m.mu, w, Z)
    pf = r.portfolio
    near = int((np.abs(P + xp) < 1e-6).sum())
    assert abs(pf["hits"] - int((P < -xp).sum())) <= near + 1
    assert pf["mean"] == pytest.approx(P.mean(), rel=1e-4, abs=1e-8)
    assert pf["var"] == pytest.approx(P.var(), rel=1e-4)
    assert pf["tail_loss_sum"] == pytest.approx(-P[P < -xp].sum(), rel=1e-4, abs=1e-6)
    assert int(pf["hist"].sum()) == M
    assert float(pf["hist_sum"].sum()) == pytest.approx(P.sum(), rel=1e-5, abs=1e-6 * M * P.std())   # per-bin sums: exact to the 2^-24 quantisation
    width = (pf["hist_hi"] - pf["hist_lo"]) / len(pf["hist"])
    for a in (0.95, 0.99):
        q = np.quantile(-P, a)
        es = (-P)[-P >= q].mean()
        assert abs(pf["VaR"][a] - q) <= width
        assert pf["ES"][a] == pytest.approx(es, rel=2e-3)
    # determinism: integer histogram words -> identical bins and bin sums on a second run and over two shards
    r2 = engine.run(M, seed, m.x, m.k, portfolio_threshold=xp, weights=w)
    np.testing.assert_array_equal(r2.portfolio["hist"], pf["hist"])
    engine.run(1000, seed, m.x, m.k, weights=np.full(N, 1.0 / N))      # restore default weights


@pytest.mark.parametrize("N", [15, 50, 160, 600])
def test_non_positive_k_counts_every_trial(engine, N):
    """S >= k is a signed comparison in the reference (serial.c:137-141): k <= 0 counts every trial, on every kernel."""
    m = synthetic_model(N)
    engine.upload_sigma(m.mu, m.Sigma)
    M = 50_000
    want = engine.run(M, 67, m.x, 1, kernel=api.KERNEL_EXACT)
    for kernel in (api.KERNEL_AUTO, api.KERNEL_EXACT) + ((api.KERNEL_FFMA,) if N <= 128 else ()):
        for k in (0, -3):
            r = engine.run(M, 67, m.x, k, kernel=kernel)
            assert (r.count, r.sum_S) == (M, want.sum_S), (r.kernel, k)


def test_runs_on_different_streams_are_ordered(engine):
    """Every run of an engine shares its scratch buffers; a run enqueued on another stream waits for the previous one."""
    import torch
    m = synthetic_model(50)
    engine.upload_sigma(m.mu, m.Sigma)
    want = [engine.run(3_000_000, 67, m.x, m.k, path_offset=o) for o in (0, 3_000_000, 6_000_000)]
    streams = [torch.cuda.Stream() for _ in range(3)]
    bufs = [torch.zeros(api.MCR_PARTIAL_WORDS, dtype=torch.int64, device="cuda") for _ in range(3)]
    for rep in range(3):
        for i, (s, b) in enumerate(zip(streams, bufs)):
            engine.run_async(3_000_000, 67, m.x, m.k, b.data_ptr(), s.cuda_stream, path_offset=3_000_000 * i)
        torch.cuda.synchronize()
        for w, b in zip(want, bufs):
            r = api.result_from_partials(b.cpu().numpy().astype(np.uint64))
            assert (r.count, r.sum_S) == (w.count, w.sum_S)


@pytest.mark.parametrize("N,M", [(50, (2 << 23) + 12_345), (300, (1 << 23) + 777)])
def test_streamed_S_values_over_several_parts(engine, oracle, N, M):
    """S_values on demand (reference: MonteCarloResult.S_values, src/model_interface.h:12): the run is enqueued in parts of 2^23
    trials, each narrowed on the device to one byte per trial (two for N > 255), copied to pinned memory while the next part runs
    and widened into the caller's int array.  Totals equal the plain run; windows at the part boundaries equal the oracle."""
    m = synthetic_model(N)
    A = engine.upload_sigma(m.mu, m.Sigma)
    plain = engine.run(M, 67, m.x, m.k)
    h = api.simulate_host(m.mu, A, M, 67, m.x, m.k, a_is_lower=True, want_S_values=True, want_hist=True)
    assert (h.count, h.sum_S, h.n_paths) == (plain.count, plain.sum_S, M)
    S = h.S_values
    assert S.dtype == np.int32 and S.shape == (M,)
    assert int(S.sum(dtype=np.int64)) == plain.sum_S and int((S >= m.k).sum()) == plain.count
    np.testing.assert_array_equal(np.bincount(S, minlength=N + 1), h.hist_S)
    for start in (0, (1 << 23) - 700, (2 << 23) - 700 if M > (2 << 23) else M - 2000, M - 1400):
        cnt = 1400
        Z, _ = engine.dump_normals(cnt, 67, path_offset=start)
        ro = oracle.run_normals(A, m.mu, m.k, m.x, Z)
        np.testing.assert_array_equal(S[start:start + cnt], ro["S_values"])


def test_mma_flops_accounting(engine):
    m = synthetic_model(50)
    engine.upload_sigma(m.mu, m.Sigma)
    engine.run(10_000, 67, m.x, m.k)
    assert engine.mma_flops_per_trial() == 3 * 2 * 64 * 64
    engine.run(10_000, 67, m.x, m.k, kernel=api.KERNEL_FFMA)
    assert engine.mma_flops_per_trial() == 0
    m = synthetic_model(600)
    engine.upload_sigma(m.mu, m.Sigma)
    engine.run(10_000, 67, m.x, m.k)
    f = engine.mma_flops_per_trial()
    assert 3 * 600 * 601 <= f <= 3 * 2 * 640 * 640, f   # at least the triangle, at most the padded dense matrix, x3 passes


# ---------------------------------------------------------------- the plugin's own multi-GPU route
@pytest.mark.skipif(not (PKG / "libmcr_b200_model.so").exists(), reason="plugin not built")
def test_plugin_multi_gpu_route_is_bit_identical(engine, monkeypatch):
    """MCR_NUM_GPUS=2 through get_cuda_model (single process, ncclCommInitAll, one grouped all-reduce): same S_values / count
    as one GPU, bit for bit.  Skipped on a single-GPU box; the driver's 2/4/8-GPU runs exercise the same route through
    bench.py's e2e leg."""
    if api.device_count() < 2:
        pytest.skip("needs two GPUs")
    plugin = C.CDLL(str(PKG / "libmcr_b200_model.so"))
    m = synthetic_model(50)
    M = 3_000_001
    monkeypatch.setenv("MCR_S_VALUES", "1")
    monkeypatch.setenv("MCR_NUM_GPUS", "1")
    one = oracle_lib.run_model_struct(plugin.get_cuda_model, m.mu, m.Sigma, m.k, m.x, M, 67)
    monkeypatch.setenv("MCR_NUM_GPUS", "2")
    two = oracle_lib.run_model_struct(plugin.get_cuda_model, m.mu, m.Sigma, m.k, m.x, M, 67)
    assert one["rc"] == 0 and two["rc"] == 0
    assert one["count"] == two["count"] and one["P_hat"] == two["P_hat"]
    np.testing.assert_array_equal(one["S_values"], two["S_values"])
    m6 = synthetic_model(600)
    A = api.build_transform(m6.Sigma)
    h1 = api.simulate_host(m6.mu, A, 400_000, 67, m6.x, m6.k, a_is_lower=True, n_gpus=1, want_hist=True)
    h2 = api.simulate_host(m6.mu, A, 400_000, 67, m6.x, m6.k, a_is_lower=True, n_gpus=2, want_hist=True)
    assert (h1.count, h1.sum_S) == (h2.count, h2.sum_S)
    np.testing.assert_array_equal(h1.hist_S, h2.hist_S)


# ---------------------------------------------------------------- the reference's own GPU model
@pytest.mark.skipif(not oracle_lib.ref_cuda_path().exists(), reason="oracle/_ref/libref_cuda.so not built (reference tree absent at build time)")
def test_reference_cuda_model_agrees_within_3_sigma(engine):
    """monte_carlos_cuda.cu, compiled UNMODIFIED for sm_100a, against this library on the same model: independent RNG streams
    (cuRAND XORWOW vs Philox), same lower-triangular transform -> P_hat within 3 sigma of the Monte Carlo standard error
    (north-star check 2)."""
    ref = oracle_lib.RefCuda()
    for N, M in ((15, 2_000_000), (50, 4_000_000)):
        if N == 15:
            g = np.load(GOLD / "pyref_n15.npz")
            mu, Sigma, x, k = g["mu"], g["Sigma"], 0.02, 5
        else:
            m = synthetic_model(N)
            mu, Sigma, x, k = m.mu, m.Sigma, m.x, m.k
        r = ref.cuda(mu, Sigma, k, x, M, 67)
        assert r["rc"] == 0 and r["name"] == "cuda"
        assert int((r["S_values"] >= k).sum()) == r["count"]
        engine.upload_sigma(mu, Sigma, api.TRANSFORM_LOWER)
        ours = engine.run(20_000_000, 67, x, k)
        pr, po = r["P_hat"], ours.P_hat
        se = np.sqrt(pr * (1 - pr) / M + po * (1 - po) / ours.n_paths)
        assert abs(pr - po) <= 3 * se, (N, pr, po, se)
        assert abs(r["S_values"].mean() - ours.sum_S / ours.n_paths) < 0.02 * max(r["S_values"].mean(), 1e-3) + 3e-3


def test_supplied_normals_reach_both_tensor_kernels(engine, oracle):
    """MCR_KERNEL_TENSOR with caller-supplied normals: odd N <= 128 and N < 36 run the supplied-normals instantiation of the tile
    kernel (tc_path_kernel<128, 3, 3, true>: same generator hand-over, MMA issue and epilogue code as the headline kernel, the
    normals read from HBM element by element), even N >= 36 and every N > 128 the streamed kernel (even N: through TMA boxes).
    Ragged trial counts, adversarial normals of very different magnitude, zeros."""
    rng = np.random.default_rng(5)
    for N, M in ((15, 1), (34, 300), (36, 129), (50, 20_001), (51, 5_000), (127, 2_000), (128, 4_097), (129, 1_000), (600, 700)):
        m = synthetic_model(N)
        A = engine.upload_sigma(m.mu, m.Sigma)
        Z = rng.standard_normal((M, N))
        Z[::7] *= 1e-3
        Z[3::11] *= 4.0
        Z[5::13] = 0.0
        ro = oracle.run_normals(A, m.mu, m.k, m.x, Z)
        r = engine.run_with_normals(Z, m.x, m.k, kernel=api.KERNEL_TENSOR)
        tile = N <= 128 and (N % 2 == 1 or N < 36)
        assert r.kernel == "tensor" and r.padded_N == (128 if tile else (N + 15) // 16 * 16)
        np.testing.assert_array_equal(r.S_values, ro["S_values"])
        assert (r.count, r.sum_S) == (ro["count"], ro["sum_S"])


# ---------------------------------------------------------------- F4: statistics and factorisation on the device
@pytest.mark.parametrize("N", [15, 50, 200, 600, 1024])
def test_device_cholesky_is_bitwise_the_host_one(engine, oracle, N):
    """mcr_device_transform: left-looking Cholesky, one thread per row, inner products accumulated in the host routine's order
    with separate roundings -> the SAME bits as mcr_build_transform (and the oracle for the sizes it covers)."""
    if N == 15:
        g = np.load(GOLD / "pyref_n15.npz")
        Sigma = g["Sigma"]
    else:
        Sigma = synthetic_model(N).Sigma
    for mode in (api.TRANSFORM_LOWER, api.TRANSFORM_GSL_FULL):
        A_dev = engine.device_transform(Sigma, mode)
        np.testing.assert_array_equal(A_dev, api.build_transform(Sigma, mode))
        if N <= 50:
            np.testing.assert_array_equal(A_dev, oracle.transform(Sigma, mode))
    with pytest.raises(api.McrError, match="not positive definite"):
        engine.device_transform(-np.eye(5))


@pytest.mark.parametrize("T,N", [(2800, 15), (1500, 600), (33, 70)])
def test_device_moments_match_numpy(engine, T, N):
    """mcr_device_moments = train.mean(), train.cov() of the reference's preprocessing (preprocessing_stats.py:18,21; ddof = 1).
    Different summation order than numpy's: 1e-12 relative to the scale of the entries."""
    rng = np.random.default_rng(T + N)
    m = synthetic_model(N)
    L = np.linalg.cholesky(m.Sigma)
    X = m.mu + rng.standard_normal((T, N)) @ L.T
    mu, Sigma = engine.device_moments(X)
    np.testing.assert_allclose(mu, X.mean(axis=0), rtol=0, atol=1e-15 + 1e-13 * np.abs(X).max())
    ref = np.cov(X, rowvar=False, ddof=1)
    np.testing.assert_allclose(Sigma, ref, rtol=0, atol=1e-12 * np.abs(ref).max())
    np.testing.assert_array_equal(Sigma, Sigma.T)
    if T > N:   # full rank: the estimated model runs end to end
        A = engine.device_transform(Sigma)
        engine.upload_model(mu, A, a_is_lower=True)
        r = engine.run(100_000, 67, m.x, m.k)
        x = engine.run(100_000, 67, m.x, m.k, kernel=api.KERNEL_EXACT)
        assert (r.count, r.sum_S) == (x.count, x.sum_S)


@pytest.mark.skipif(not (PKG / "libmcr_b200_model.so").exists(), reason="plugin not built")
def test_plugin_device_cholesky_option(monkeypatch):
    """MCR_DEVICE_CHOLESKY=1 through get_cuda_model: identical results (the factor is bitwise the host's)."""
    plugin = C.CDLL(str(PKG / "libmcr_b200_model.so"))
    m = synthetic_model(100)
    monkeypatch.setenv("MCR_S_VALUES", "1")
    a = oracle_lib.run_model_struct(plugin.get_cuda_model, m.mu, m.Sigma, m.k, m.x, 300_000, 67)
    monkeypatch.setenv("MCR_DEVICE_CHOLESKY", "1")
    b = oracle_lib.run_model_struct(plugin.get_cuda_model, m.mu, m.Sigma, m.k, m.x, 300_000, 67)
    assert a["rc"] == 0 and b["rc"] == 0 and a["count"] == b["count"]
    np.testing.assert_array_equal(a["S_values"], b["S_values"])


_PAIR_CHILD = r"""
import json, sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import oracle_lib
from mcr_b200 import api
from mcr_b200.synthetic import synthetic_model
oracle = oracle_lib.Oracle()
e = api.Engine(0)
out = []
for N, mode, M in [(144, api.TRANSFORM_GSL_FULL, 5_000), (385, api.TRANSFORM_LOWER, 6_000), (600, api.TRANSFORM_LOWER, 148 * 128 + 77),
                   (600, api.TRANSFORM_LOWER, 100), (1000, api.TRANSFORM_LOWER, 2_500)]:
    m = synthetic_model(N)
    A = e.upload_sigma(m.mu, m.Sigma, mode)
    Z, _ = e.dump_normals(M, 19)
    ro = oracle.run_normals(A, m.mu, m.k, m.x, Z)
    h = api.simulate_host(m.mu, A, M, 19, m.x, m.k, a_is_lower=(mode == api.TRANSFORM_LOWER), want_S_values=True)
    out.append({"N": N, "kernel": h.kernel, "same_S": bool(np.array_equal(h.S_values, ro["S_values"])),
                "same_counts": (h.count, h.sum_S) == (ro["count"], ro["sum_S"]), "rechecked": h.n_rechecked, "M": M})
m = synthetic_model(600)
e.upload_sigma(m.mu, m.Sigma)
fast = e.run(3_000_000, 67, m.x, m.k)
exact = e.run(3_000_000, 67, m.x, m.k, kernel=api.KERNEL_EXACT)
out.append({"N": 600, "kernel": fast.kernel, "same_S": True, "same_counts": (fast.count, fast.sum_S) == (exact.count, exact.sum_S),
            "rechecked": fast.n_rechecked, "M": 3_000_000})
print(json.dumps(out))
"""


def test_cta_pair_form_of_the_streamed_kernel_is_bit_identical():
    """MCR_STREAM_PAIR=1 (tcgen05 cta_group::2: one M = 256 MMA per CTA pair, half of every transform piece per CTA; read once per
    process, hence the child): one, two and three passes, dense and lower-triangular, a run of a single tile (the second CTA of
    the pair works on a ghost tile), per-trial S against the fp64 oracle and AUTO == EXACT over 3e6 trials at N = 600."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, MCR_STREAM_PAIR="1")
    p = subprocess.run([sys.executable, "-c", _PAIR_CHILD, str(PKG), str(ROOT / "tests")], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    res = json.loads(p.stdout.strip().splitlines()[-1])
    assert len(res) == 6
    for r in res:
        assert r["kernel"] == "tensor" and r["same_S"] and r["same_counts"], r
        assert r["rechecked"] < max(r["M"] // 5, 50), r


@pytest.mark.parametrize("N,mode", [(600, api.TRANSFORM_LOWER), (600, api.TRANSFORM_GSL_FULL), (300, api.TRANSFORM_LOWER), (1000, api.TRANSFORM_LOWER),
                                    (257, api.TRANSFORM_LOWER)])
def test_supplied_normals_streamed_repeatedly(engine, N, mode):
    """The streamed kernel's supplied-normals path reads the caller's matrix through a TMA tensor map into a ring of boxes that
    generator warps read with ordinary loads (even N; odd N = 257 keeps the element loads).  The hand-back of a box to the loader
    needs a proxy fence (generic reads before the async-proxy refill): without it a few trials in a thousand came out wrong, in
    some runs only - so this one repeats a run that keeps the tensor core busy, per-trial S against the fp64 kernel on the same
    normals, several tiles per CTA."""
    m = synthetic_model(N)
    engine.upload_sigma(m.mu, m.Sigma, mode)
    M = 148 * 128 * 2 + 77
    Z = np.random.default_rng(N).standard_normal((M, N))
    exact = engine.run_with_normals(Z, m.x, m.k, kernel=api.KERNEL_EXACT)
    for _ in range(6):
        r = engine.run_with_normals(Z, m.x, m.k, kernel=api.KERNEL_TENSOR)
        assert r.kernel == "tensor"
        np.testing.assert_array_equal(r.S_values, exact.S_values)
        assert (r.count, r.sum_S) == (exact.count, exact.sum_S)


def test_supplied_normals_tma_edge_cases(engine, oracle):
    """Supplied normals through the TMA boxes of the streamed kernel: a single trial, one short of a tile, a ragged last K-block at
    the largest even N, and a device matrix that is NOT 16-byte aligned (the tensor map cannot describe it: element loads)."""
    import torch
    rng = np.random.default_rng(11)
    for N, M in ((600, 1), (600, 127), (64, 1), (130, 129), (1022, 300)):
        m = synthetic_model(N)
        A = engine.upload_sigma(m.mu, m.Sigma)
        Z = rng.standard_normal((M, N))
        ro = oracle.run_normals(A, m.mu, m.k, m.x, Z)
        r = engine.run_with_normals(Z, m.x, m.k, kernel=api.KERNEL_TENSOR)
        assert r.kernel == "tensor" and r.padded_N == (N + 15) // 16 * 16
        np.testing.assert_array_equal(r.S_values, ro["S_values"])
        assert (r.count, r.sum_S) == (ro["count"], ro["sum_S"])
    N, M = 256, 5_000
    m = synthetic_model(N)
    A = engine.upload_sigma(m.mu, m.Sigma)
    Z = rng.standard_normal((M, N))
    ro = oracle.run_normals(A, m.mu, m.k, m.x, Z)
    buf = torch.zeros(M * N + 1, dtype=torch.float64, device="cuda")
    buf[1:] = torch.from_numpy(Z.reshape(-1)).cuda()
    for off in (0, 1):   # 16-byte aligned copy (TMA) and the copy one double further (element loads)
        if off == 0:
            src = buf[1:].clone()
            assert src.data_ptr() % 16 == 0
        else:
            src = buf[1:]
            assert src.data_ptr() % 16 == 8
        r = engine.run_with_normals_device(src.data_ptr(), M, m.x, m.k, kernel=api.KERNEL_TENSOR)
        assert (r.kernel, r.count, r.sum_S) == ("tensor", ro["count"], ro["sum_S"])
