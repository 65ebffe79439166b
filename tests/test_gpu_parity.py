"""GPU parity tests (run on a B200 with -m gpu).  Everything goes through the C ABI.

Bars:
  1. identical normals  -> S_values / count / sum_S BIT-EXACT vs the fp64 oracle (both the fp32
     FFMA2 path with its guard band + fp64 fix-up, and the fp64 exact kernel); P_hat, std_error,
     CI are then the same fp64 expressions -> compared with == (tolerance 0).
  2. in-register generator: Philox words bit-exact vs the CPU Philox; normals within 2e-4 absolute
     of the libm Box-Muller (MUFU approximations); the run on generated normals is bit-exact vs the
     oracle fed the DUMPED normals.
  3. independent streams (Philox vs MT19937) agree within 3 sigma of the binomial error.
  4. shard invariance: totals do not depend on how the path range is partitioned.
"""
from pathlib import Path

import numpy as np
import pytest

import oracle_lib
from mcr_b200 import api
from mcr_b200.synthetic import synthetic_model

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def real15():
    g = np.load(GOLD / "pyref_n15.npz")
    return g["mu"], g["Sigma"]


def model(N):
    if N == 15:
        mu, Sigma = real15()
        return mu, Sigma, 0.02, 5
    m = synthetic_model(N)
    return m.mu, m.Sigma, m.x, m.k


# ---------------------------------------------------------------- generator
@pytest.mark.parametrize("offset", [0, 12345, (1 << 32) - 7, (3 << 40) + 1])
def test_philox_words_bit_exact(engine, oracle, offset):
    mu, Sigma, x, k = model(15)
    engine.upload_sigma(mu, Sigma)
    seed = (0xDEADBEEF << 32) | 67
    _, U = engine.dump_normals(3000, seed, path_offset=offset, normals=False, uniforms=True)
    np.testing.assert_array_equal(U, oracle.philox_uniforms(seed, offset, 3000, 15))


def test_normals_close_to_libm_box_muller_and_standard(engine, oracle):
    mu, Sigma, x, k = model(50)
    engine.upload_sigma(mu, Sigma)
    Z, _ = engine.dump_normals(40000, 67)
    Zo = oracle.philox_normals(67, 0, 40000, 50)
    assert np.abs(Z - Zo).max() < 2e-4          # MUFU sin/cos/lg2 vs libm, |z| up to 6.8
    assert np.abs(Z - Zo).mean() < 2e-6
    z = Z.astype(np.float64).ravel()
    assert abs(z.mean()) < 5 / np.sqrt(z.size)
    assert abs(z.var() - 1.0) < 5 * np.sqrt(2.0 / z.size)
    assert abs((z ** 3).mean()) < 0.02 and abs((z ** 4).mean() - 3.0) < 0.05
    assert np.abs(Z).max() <= 6.8
    # columns are uncorrelated
    c = np.corrcoef(Z[:, :8].T)
    assert np.abs(c - np.eye(8)).max() < 0.03


# ---------------------------------------------------------------- identical normals
@pytest.mark.parametrize("name", ["pyref_n15.npz", "pyref_n25.npz"])
@pytest.mark.parametrize("kernel", [api.KERNEL_FFMA, api.KERNEL_EXACT, api.KERNEL_TENSOR])
def test_golden_python_reference_bit_exact(engine, name, kernel):
    g = np.load(GOLD / name)
    engine.upload_sigma(g["mu"], g["Sigma"], api.TRANSFORM_LOWER)
    r = engine.run_with_normals(g["Z"], float(g["x"]), int(g["k"]), kernel=kernel)
    np.testing.assert_array_equal(r.S_values, g["S_values"])
    assert r.count == int(g["count"]) and r.sum_S == int(g["S_values"].sum())
    assert (r.P_hat, r.std_error, r.ci_lower, r.ci_upper) == \
        (float(g["P_hat"]), float(g["std_error"]), float(g["ci_lower"]), float(g["ci_upper"]))
    assert r.kernel == {api.KERNEL_FFMA: "ffma", api.KERNEL_EXACT: "exact", api.KERNEL_TENSOR: "tensor"}[kernel]


@pytest.mark.parametrize("N,mode,M", [(15, api.TRANSFORM_GSL_FULL, 200000), (15, api.TRANSFORM_LOWER, 100000),
                                      (25, api.TRANSFORM_GSL_FULL, 60000), (50, api.TRANSFORM_LOWER, 60000),
                                      (50, api.TRANSFORM_GSL_FULL, 40000), (100, api.TRANSFORM_LOWER, 30000),
                                      (100, api.TRANSFORM_GSL_FULL, 20000), (128, api.TRANSFORM_LOWER, 20000)])
def test_c_serial_stream_bit_exact(engine, oracle, N, mode, M):
    """Normals drawn exactly like the C-serial model (MT19937 + polar, seed 67), fed to both sides."""
    mu, Sigma, x, k = model(N)
    A = engine.upload_sigma(mu, Sigma, mode)
    np.testing.assert_array_equal(A, oracle.transform(Sigma, mode))
    ro = oracle.run_mt19937(A, mu, k, x, M, 67, want_Z=True)
    # the tensor path too (north-star check 1: identical pre-generated normals fed through the interface): the streamed
    # tcgen05 kernel reads the caller's fp64 normals from HBM in place of its Philox generator
    for kernel in (api.KERNEL_FFMA, api.KERNEL_TENSOR, api.KERNEL_EXACT):
        r = engine.run_with_normals(ro["Z"], x, k, kernel=kernel)
        np.testing.assert_array_equal(r.S_values, ro["S_values"])
        assert (r.count, r.sum_S, r.n_paths) == (ro["count"], ro["sum_S"], M)
        assert (r.P_hat, r.std_error, r.ci_lower, r.ci_upper) == oracle.stats(ro["count"], M)
        assert r.kernel == {api.KERNEL_FFMA: "ffma", api.KERNEL_EXACT: "exact", api.KERNEL_TENSOR: "tensor"}[kernel]
    assert r.n_rechecked == 0  # exact kernel rechecks nothing


def test_real_data_golden_counts(engine, oracle):
    """Shipped 15-index data, config.yaml x/k/seed, M = 1e6: the counts frozen in SURVEY.md §8c."""
    mu, Sigma = real15()
    for mode, want in ((api.TRANSFORM_GSL_FULL, (47605, 674239)), (api.TRANSFORM_LOWER, (47283, 673423))):
        A = engine.upload_sigma(mu, Sigma, mode)
        Z = oracle.run_mt19937(A, mu, 5, 0.02, 1000000, 67, want_S=False, want_Z=True)["Z"]
        r = engine.run_with_normals(Z, 0.02, 5)
        assert (r.count, r.sum_S) == want
        assert list(r.S_values[:8]) == [0, 0, 3, 0, 3, 0, 0, 0]


def test_guard_band_catches_near_threshold_trials(engine, oracle):
    """Adversarial normals: trials constructed so that some R_i sits within 1e-9 of -x.  fp32 alone
    would mis-decide some; the band + fp64 fix-up must still be bit-exact."""
    mu, Sigma, x, k = model(15)
    A = engine.upload_sigma(mu, Sigma, api.TRANSFORM_LOWER)
    rng = np.random.default_rng(5)
    M = 20000
    Z = rng.standard_normal((M, 15))
    # solve for z_0 such that R_0 = mu_0 + A_00 z_0 = -x + eps
    eps = rng.uniform(-2e-9, 2e-9, size=M)
    Z[:, 0] = (-x + eps - mu[0]) / A[0, 0]
    ro = oracle.run_normals(A, mu, k, x, Z)
    r = engine.run_with_normals(Z, x, k, kernel=api.KERNEL_FFMA)
    np.testing.assert_array_equal(r.S_values, ro["S_values"])
    assert r.count == ro["count"] and r.n_rechecked >= M // 2
    raw = engine.run_with_normals(Z, x, k, kernel=api.KERNEL_FFMA, no_recheck=True)
    assert raw.n_rechecked == 0
    assert (raw.S_values != ro["S_values"]).sum() > 0  # fp32 alone does flip decisions here


# ---------------------------------------------------------------- in-register generator
@pytest.mark.parametrize("N,mode,M", [(15, api.TRANSFORM_LOWER, 300000), (15, api.TRANSFORM_GSL_FULL, 100000),
                                      (25, api.TRANSFORM_LOWER, 100000), (50, api.TRANSFORM_LOWER, 200000),
                                      (50, api.TRANSFORM_GSL_FULL, 50000), (64, api.TRANSFORM_LOWER, 50000),
                                      (100, api.TRANSFORM_LOWER, 100000), (128, api.TRANSFORM_LOWER, 30000)])
def test_philox_run_bit_exact_vs_oracle_on_dumped_normals(engine, oracle, N, mode, M):
    mu, Sigma, x, k = model(N)
    A = engine.upload_sigma(mu, Sigma, mode)
    seed, off = 67, 1000003
    Z, _ = engine.dump_normals(M, seed, path_offset=off)
    ro = oracle.run_normals(A, mu, k, x, Z)
    for kernel, name in ((api.KERNEL_FFMA, "ffma"), (api.KERNEL_TENSOR, "tensor"), (api.KERNEL_EXACT, "exact"), (api.KERNEL_AUTO, None)):
        r = engine.run(M, seed, x, k, path_offset=off, kernel=kernel)
        assert (name is None or r.kernel == name) and r.padded_N >= N
        assert (r.count, r.sum_S, r.n_paths) == (ro["count"], ro["sum_S"], M), (r.kernel, r)
    # simulate_host starts at path 0: compare with its own dump; S_values from both fast kernels
    Z0, _ = engine.dump_normals(M, seed)
    r0 = oracle.run_normals(A, mu, k, x, Z0)
    for kernel in (api.KERNEL_FFMA, api.KERNEL_TENSOR):
        h = api.simulate_host(mu, A, M, seed, x, k, a_is_lower=(mode == api.TRANSFORM_LOWER), want_S_values=True, kernel=kernel)
        np.testing.assert_array_equal(h.S_values, r0["S_values"])
        assert (h.count, h.sum_S) == (r0["count"], r0["sum_S"])


def test_independent_streams_agree_within_3_sigma(engine, oracle):
    for N, mode in ((15, api.TRANSFORM_GSL_FULL), (50, api.TRANSFORM_LOWER)):
        mu, Sigma, x, k = model(N)
        A = engine.upload_sigma(mu, Sigma, mode)
        Mo, Mg = 400000, 20000000
        ro = oracle.run_mt19937(A, mu, k, x, Mo, 67, want_S=False)
        r = engine.run(Mg, 67, x, k)
        po, pg = ro["count"] / Mo, r.P_hat
        se = np.sqrt(po * (1 - po) / Mo + pg * (1 - pg) / Mg)
        assert abs(po - pg) <= 3 * se, (N, po, pg, se)
        mean_S_o, mean_S_g = ro["sum_S"] / Mo, r.sum_S / Mg
        assert abs(mean_S_o - mean_S_g) < 0.05 * max(mean_S_o, 1e-3) + 3e-3


# ---------------------------------------------------------------- sharding / full size properties
def test_shard_invariance(engine):
    mu, Sigma, x, k = model(50)
    engine.upload_sigma(mu, Sigma)
    M = 3000001
    whole = engine.run(M, 99, x, k)
    cuts = [0, 1, 777777, 2000000, M]
    parts = [engine.run(b - a, 99, x, k, path_offset=a) for a, b in zip(cuts[:-1], cuts[1:])]
    assert sum(p.count for p in parts) == whole.count
    assert sum(p.sum_S for p in parts) == whole.sum_S
    assert sum(p.n_paths for p in parts) == M
    assert engine.run(M, 100, x, k).sum_S != whole.sum_S  # the seed matters


def test_full_size_stoxx50_shape_properties(engine):
    """BASELINE config 2 at full size (N=50, M=1e8): additivity over shards, determinism,
    statistics consistent with the counts, chunking (M > one chunk of 2^27 paths)."""
    mu, Sigma, x, k = model(50)
    engine.upload_sigma(mu, Sigma)
    M = 100000000
    a = engine.run(M, 67, x, k)
    b = engine.run(M, 67, x, k)
    assert (a.count, a.sum_S, a.n_rechecked) == (b.count, b.sum_S, b.n_rechecked)
    half = [engine.run(M // 2, 67, x, k, path_offset=o) for o in (0, M // 2)]
    assert half[0].count + half[1].count == a.count and half[0].sum_S + half[1].sum_S == a.sum_S
    assert a.P_hat == a.count / M
    big = engine.run((1 << 27) + 12345, 67, x, k)  # crosses a chunk boundary
    parts = [engine.run(1 << 27, 67, x, k), engine.run(12345, 67, x, k, path_offset=1 << 27)]
    assert big.count == parts[0].count + parts[1].count and big.sum_S == parts[0].sum_S + parts[1].sum_S
    assert 0 < a.n_rechecked < M // 100


# ---------------------------------------------------------------- edge cases and errors
def test_edge_cases(engine, oracle):
    mu, Sigma, x, k = model(15)
    A = engine.upload_sigma(mu, Sigma)
    assert engine.run(0, 1, x, k).n_paths == 0
    for M in (1, 31, 127, 128, 129, 1000):
        Z, _ = engine.dump_normals(M, 5)
        ro = oracle.run_normals(A, mu, k, x, Z)
        r = engine.run(M, 5, x, k)
        assert (r.count, r.sum_S, r.n_paths) == (ro["count"], ro["sum_S"], M)
    r = engine.run(5000, 5, x, 0)
    assert r.count == 5000 and r.P_hat == 1.0 and r.ci_upper == 1.0       # S >= 0 always
    r = engine.run(5000, 5, x, 16)
    assert r.count == 0 and r.std_error == 0.0 and r.ci_lower == 0.0      # S <= N < k
    # N = 1 and a huge threshold that every trial crosses
    engine.upload_model(np.array([0.0]), np.array([[0.01]]), a_is_lower=True)
    r = engine.run(10000, 3, -1.0, 1)   # R < 1 always
    assert r.count == 10000 and r.sum_S == 10000
    r = engine.run(10000, 3, 1.0, 1)    # R < -1 never
    assert r.count == 0 and r.sum_S == 0


def test_tensor_kernel_raw_decisions_are_close_and_recheck_makes_them_exact(engine, oracle):
    """fp16x3 tensor-core correlate: without the fp64 recheck a handful of near-threshold decisions may
    differ from fp64 (stated tolerance: at most 1e-5 of the per-asset decisions); with it, none."""
    for N in (50, 100, 128):
        mu, Sigma, x, k = model(N)
        A = engine.upload_sigma(mu, Sigma)
        M = 400000
        Z, _ = engine.dump_normals(M, 31337)
        ro = oracle.run_normals(A, mu, k, x, Z)
        raw = api.simulate_host(mu, A, M, 31337, x, k, a_is_lower=True, want_S_values=True, kernel=api.KERNEL_TENSOR, no_recheck=True)
        assert raw.n_rechecked == 0
        assert np.abs(raw.S_values.astype(np.int64) - ro["S_values"]).sum() <= 1e-5 * M * N
        fix = api.simulate_host(mu, A, M, 31337, x, k, a_is_lower=True, want_S_values=True, kernel=api.KERNEL_TENSOR)
        np.testing.assert_array_equal(fix.S_values, ro["S_values"])
        assert 0 < fix.n_rechecked < M // 50


def test_portfolio_extension_against_oracle(engine, oracle):
    """North-star extension (A7; the reference has no portfolio / VaR / ES code, so this is checked
    against the oracle's own definition on the dumped normals): P = w.R per trial, hits = #{P < -x_p},
    mean/variance, ES numerator, VaR/ES from the merged 2048-bin histogram.
    Stated tolerances: moments 1e-5 relative (fp16x3 products, fp32 per-thread partial sums);
    hits may differ by the trials whose P sits within 1e-6 of the threshold; VaR within one
    histogram bin (sigma_p / 128); ES within 0.2 % (bin interpolation at the VaR boundary)."""
    for N, weights in ((15, None), (50, None), (100, "ramp")):
        mu, Sigma, x, k = model(N)
        A = engine.upload_sigma(mu, Sigma)
        w = np.full(N, 1.0 / N) if weights is None else np.arange(1, N + 1) / np.arange(1, N + 1).sum()
        M, seed, xp = 1000000, 2025, 0.015
        r = engine.run(M, seed, x, k, portfolio_threshold=xp, weights=w)
        assert r.kernel == "tensor"
        Z, _ = engine.dump_normals(M, seed)
        ro = oracle.run_normals(A, mu, k, x, Z)
        assert (r.count, r.sum_S) == (ro["count"], ro["sum_S"])       # the base outputs are untouched and exact
        P = oracle.portfolio(A, mu, w, Z)
        pf = r.portfolio
        near = int((np.abs(P + xp) < 1e-6).sum())
        assert abs(pf["hits"] - int((P < -xp).sum())) <= near + 1
        assert pf["mean"] == pytest.approx(P.mean(), rel=1e-4, abs=1e-8)
        assert pf["var"] == pytest.approx(P.var(), rel=1e-4)
        assert pf["tail_loss_sum"] == pytest.approx(-P[P < -xp].sum(), rel=1e-4, abs=1e-6)
        assert int(pf["hist"].sum()) == M
        width = (pf["hist_hi"] - pf["hist_lo"]) / len(pf["hist"])
        for a in (0.95, 0.99):
            q = np.quantile(-P, a)                                    # VaR_a of the loss
            es = (-P)[-P >= q].mean()
            assert abs(pf["VaR"][a] - q) <= width
            assert pf["ES"][a] == pytest.approx(es, rel=2e-3)
        engine.run(1000, seed, x, k, weights=np.full(N, 1.0 / N))      # restore default weights


def test_histogram_of_crash_counts(engine, oracle):
    """want_hist: hist_S[s] = #{trials : S == s}, for every kernel flavour; exact like S itself."""
    for N, kernels in ((15, (api.KERNEL_FFMA, api.KERNEL_EXACT)), (50, (api.KERNEL_TENSOR, api.KERNEL_FFMA)), (160, (api.KERNEL_AUTO,))):
        mu, Sigma, x, k = model(N)
        A = engine.upload_sigma(mu, Sigma)
        M = 300000 if N < 100 else 30000
        Z, _ = engine.dump_normals(M, 5)
        ro = oracle.run_normals(A, mu, k, x, Z)
        want = np.bincount(ro["S_values"], minlength=N + 1)
        for kern in kernels:
            r = engine.run(M, 5, x, k, kernel=kern, want_hist=True)
            np.testing.assert_array_equal(r.hist_S, want)
            assert int(r.hist_S[k:].sum()) == r.count and int((np.arange(N + 1) * r.hist_S).sum()) == r.sum_S


def test_extension_partials_shard_like_the_counts(engine, oracle):
    """mcr_run_async_ext leaves the extension outputs of a shard as plain sums in two device buffers; adding the buffers of
    the shards of a path range (what a sum-allreduce across GPUs does) must reproduce mcr_run over the whole range:
    integer parts (histogram of S, portfolio hits and histogram) exactly, the sums to 1e-6 relative, per-bin sums and ES to
    1e-4 (they are staged per CTA with fp32 shared-memory atomics, whose grouping changes with the shard boundaries).  mcr_simulate_host carries the same outputs."""
    import torch
    N = 50
    mu, Sigma, x, k = model(N)
    A = engine.upload_sigma(mu, Sigma)
    M, seed, xp = 700001, 77, 0.015
    whole = engine.run(M, seed, x, k, portfolio_threshold=xp, want_hist=True)
    stream = torch.cuda.Stream()
    words = torch.zeros(api.MCR_PARTIAL_WORDS, dtype=torch.int64, device="cuda")
    counts = torch.zeros(api.MCR_EXT_COUNT_WORDS, dtype=torch.int64, device="cuda")
    sums = torch.zeros(api.MCR_EXT_SUM_WORDS, dtype=torch.float64, device="cuda")
    tw, tc, ts = torch.zeros_like(words), torch.zeros_like(counts), torch.zeros_like(sums)
    cuts = [0, 1, 250000, 250128, M]
    with torch.cuda.stream(stream):
        for a, b in zip(cuts[:-1], cuts[1:]):
            engine.run_async_ext(b - a, seed, x, k, words.data_ptr(), counts.data_ptr(), sums.data_ptr(), stream.cuda_stream,
                                 path_offset=a, portfolio_threshold=xp, want_hist=True)
            tw += words; tc += counts; ts += sums
    stream.synchronize()
    r = api.result_from_partials_ext(tw.cpu().numpy().astype(np.uint64), tc.cpu().numpy().astype(np.uint64), ts.cpu().numpy(),
                                     engine.portfolio_hist_range(), N=N, portfolio=True)
    assert (r.n_paths, r.count, r.sum_S) == (M, whole.count, whole.sum_S)
    np.testing.assert_array_equal(r.hist_S, whole.hist_S)
    pw, pr = whole.portfolio, r.portfolio
    assert pr["hits"] == pw["hits"] and (pr["hist_lo"], pr["hist_hi"]) == (pw["hist_lo"], pw["hist_hi"])
    np.testing.assert_array_equal(pr["hist"], pw["hist"])
    np.testing.assert_allclose(pr["hist_sum"], pw["hist_sum"], rtol=1e-4, atol=1e-7)
    for key in ("mean", "var", "tail_loss_sum"):
        assert pr[key] == pytest.approx(pw[key], rel=1e-6)
    for a in (0.95, 0.99):
        assert pr["VaR"][a] == pytest.approx(pw["VaR"][a], rel=1e-6) and pr["ES"][a] == pytest.approx(pw["ES"][a], rel=1e-4)
    # without the extension flags the buffers come back zero
    engine.run_async_ext(1000, seed, x, k, words.data_ptr(), counts.data_ptr(), sums.data_ptr(), stream.cuda_stream)
    stream.synchronize()
    assert int(counts.abs().sum()) == 0 and float(sums.abs().sum()) == 0.0 and int(words[0]) == 1000
    # one-shot host entry point
    h = api.simulate_host(mu, A, M, seed, x, k, a_is_lower=True, portfolio_threshold=xp, want_hist=True)
    assert (h.count, h.sum_S) == (whole.count, whole.sum_S)
    np.testing.assert_array_equal(h.hist_S, whole.hist_S)
    assert h.portfolio["hits"] == pw["hits"]
    np.testing.assert_array_equal(h.portfolio["hist"], pw["hist"])
    assert h.portfolio["VaR"][0.99] == pytest.approx(pw["VaR"][0.99], rel=1e-6)


def test_kernel_selection(engine):
    mu, Sigma, x, k = model(15)
    engine.upload_sigma(mu, Sigma)
    assert engine.run(1000, 1, x, k).kernel == "ffma"          # small N: FFMA kernel
    m = synthetic_model(64)
    engine.upload_sigma(m.mu, m.Sigma)
    assert engine.run(1000, 1, m.x, m.k).kernel == "tensor"    # N >= 36: tensor-core kernel
    # supplied normals: AUTO keeps the FFMA kernel below N = 64 (and for odd N up to 128), from there the streamed tensor kernel
    # reads the matrix faster (TMA boxes); an explicit request reaches a tensor kernel at every N
    Z = np.zeros((4, 64))
    assert engine.run_with_normals(Z, m.x, m.k).kernel == "tensor"
    m50 = synthetic_model(50)
    engine.upload_sigma(m50.mu, m50.Sigma)
    assert engine.run_with_normals(np.zeros((4, 50)), m50.x, m50.k).kernel == "ffma"
    engine.upload_sigma(m.mu, m.Sigma)
    rt = engine.run_with_normals(Z, m.x, m.k, kernel=api.KERNEL_TENSOR)
    assert rt.kernel == "tensor" and rt.count == 0 and list(rt.S_values) == [0, 0, 0, 0]
    with pytest.raises(api.McrError, match="portfolio outputs are not available with supplied normals"):
        engine._lib.mcr_run_with_normals   # (the error path itself is exercised through the C ABI below)
        import ctypes as C
        from mcr_b200 import _lib
        o = api.make_options(api.KERNEL_TENSOR, portfolio_threshold=0.01)
        r = _lib.Result()
        _lib.check(engine._lib.mcr_run_with_normals(engine._h, 4, Z.ctypes.data, float(m.x), int(m.k), C.byref(o), None, C.byref(r)))


@pytest.mark.parametrize("N", [50, 100, 160])
def test_tensor_kernels_ragged_trial_counts_and_offsets(engine, oracle, N):
    """Tiles are 128 trials: partial tiles, fewer tiles than CTAs, tile counts not divisible by the number of
    generator groups, and a path range crossing 2^32 (Philox counter carry)."""
    m = synthetic_model(N)
    A = engine.upload_sigma(m.mu, m.Sigma)
    for M, off in ((1, 0), (127, 5), (128, 0), (129, 7), (385, (1 << 32) - 100), (148 * 128 * 3 + 17, 12345)):
        Z, _ = engine.dump_normals(M, 9, path_offset=off)
        ro = oracle.run_normals(A, m.mu, m.k, m.x, Z)
        r = engine.run(M, 9, m.x, m.k, path_offset=off, kernel=api.KERNEL_TENSOR)
        assert r.kernel == "tensor" and (r.count, r.sum_S, r.n_paths) == (ro["count"], ro["sum_S"], M), (N, M, off)
    assert engine.run(0, 9, m.x, m.k, kernel=api.KERNEL_TENSOR).n_paths == 0


@pytest.mark.parametrize("N", [33, 43, 45, 46, 47, 72, 119])
def test_tensor_kernel_every_tail_shape(engine, oracle, N):
    """The last 8-column chunk of a trial row holds N % 8 normals, the constant 1 and zeros, and only the Philox blocks /
    Box-Muller pairs it needs are drawn (N % 8 = 1..7 take different code paths; 0 is covered by N = 64, 128 above).
    Enough trials that every warpgroup runs several tiles: all-zero chunks of single-K-block tiles are written once."""
    m = synthetic_model(N)
    A = engine.upload_sigma(m.mu, m.Sigma)
    M = 148 * 6 * 128 * 3 + 77
    Z, _ = engine.dump_normals(M, 31)
    ro = oracle.run_normals(A, m.mu, m.k, m.x, Z)
    r = engine.run(M, 31, m.x, m.k, kernel=api.KERNEL_TENSOR)
    assert r.kernel == "tensor" and (r.count, r.sum_S, r.n_paths) == (ro["count"], ro["sum_S"], M)
    h = api.simulate_host(m.mu, A, M, 31, m.x, m.k, a_is_lower=True, want_S_values=True, kernel=api.KERNEL_TENSOR)
    np.testing.assert_array_equal(h.S_values, ro["S_values"])


@pytest.mark.parametrize("N,mode,M", [(160, api.TRANSFORM_LOWER, 20000), (200, api.TRANSFORM_GSL_FULL, 8000),
                                      (300, api.TRANSFORM_GSL_FULL, 6000), (500, api.TRANSFORM_LOWER, 5000),
                                      (600, api.TRANSFORM_LOWER, 5000), (1024, api.TRANSFORM_LOWER, 1500)])
def test_large_n_row_slab_tensor_kernel_bit_exact(engine, oracle, N, mode, M):
    """N > 128 (Stoxx600 shape): the streamed tensor kernel (normals in TMEM, one to three passes over row ranges),
    lower-triangular and dense transforms; still bit-exact."""
    m = synthetic_model(N)
    A = engine.upload_sigma(m.mu, m.Sigma, mode)
    Z, _ = engine.dump_normals(M, 11, path_offset=77)
    ro = oracle.run_normals(A, m.mu, m.k, m.x, Z)
    r = engine.run(M, 11, m.x, m.k, path_offset=77)
    assert r.kernel == "tensor" and r.padded_N >= N
    assert (r.count, r.sum_S, r.n_paths) == (ro["count"], ro["sum_S"], M)
    rx = engine.run(M, 11, m.x, m.k, path_offset=77, kernel=api.KERNEL_EXACT)
    assert rx.kernel == "exact" and (rx.count, rx.sum_S) == (ro["count"], ro["sum_S"])
    Z0, _ = engine.dump_normals(M, 11)
    r0 = oracle.run_normals(A, m.mu, m.k, m.x, Z0)
    h = api.simulate_host(m.mu, A, M, 11, m.x, m.k, a_is_lower=(mode == api.TRANSFORM_LOWER), want_S_values=True)
    np.testing.assert_array_equal(h.S_values, r0["S_values"])
    with pytest.raises(api.McrError, match="no FFMA instantiation"):
        engine.run(1000, 11, m.x, m.k, kernel=api.KERNEL_FFMA)


@pytest.mark.parametrize("N,mode,tiles", [(320, api.TRANSFORM_LOWER, 3), (600, api.TRANSFORM_LOWER, 3), (300, api.TRANSFORM_GSL_FULL, 2)])
def test_single_pass_kernel_several_tiles_per_cta(engine, oracle, N, mode, tiles):
    """The single-pass kernel (257 <= N <= 640) pipelines across the tiles of a CTA: the accumulators of a tile are thresholded
    and re-initialised (same pass) while the next tile's first K-blocks are already queued for the tensor core - at the tile
    boundary by default, slab by slab inside the K loop in builds with MCR_WIDE_ELAG > 0.  Enough trials for several tiles per
    CTA plus a ragged one; per-trial S identical to the oracle."""
    m = synthetic_model(N)
    A = engine.upload_sigma(m.mu, m.Sigma, mode)
    nt = 128 if N <= 512 else 96
    M = 148 * nt * tiles + nt + 5
    Z, _ = engine.dump_normals(M, 13)
    ro = oracle.run_normals(A, m.mu, m.k, m.x, Z)
    h = api.simulate_host(m.mu, A, M, 13, m.x, m.k, a_is_lower=(mode == api.TRANSFORM_LOWER), want_S_values=True)
    assert h.kernel == "tensor"
    np.testing.assert_array_equal(h.S_values, ro["S_values"])
    assert (h.count, h.sum_S) == (ro["count"], ro["sum_S"])


def test_slab_fixup_paths_under_a_wide_band():
    """The tensor kernels have two fix-ups: a single near-threshold decision is re-evaluated alone (exact_pairs_kernel), a
    trial with several goes to the full fp64 kernel.  A band widened 32x (env knob, read once per process) makes both
    common -- and trials hitting both in different slabs -- and the totals must still equal the fp64 kernel's."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from mcr_b200 import api\n"
        "from mcr_b200.synthetic import synthetic_model\n"
        "e = api.Engine(0)\n"
        "for N, M in ((50, 400000), (128, 200000), (160, 40000), (600, 20000), (1000, 6000)):\n"
        "    m = synthetic_model(N); e.upload_sigma(m.mu, m.Sigma)\n"
        "    r = e.run(M, 21, m.x, m.k, path_offset=3); x = e.run(M, 21, m.x, m.k, path_offset=3, kernel=api.KERNEL_EXACT)\n"
        "    assert r.kernel == 'tensor' and x.kernel == 'exact'\n"
        "    assert (r.count, r.sum_S) == (x.count, x.sum_S), (N, r.count, x.count, r.sum_S, x.sum_S)\n"
        "    print(N, r.n_rechecked)\n"
    ) % str(Path(__file__).resolve().parent.parent / "gpu-montecarlo-risk_b200")
    import os, subprocess, sys
    for large_n in ("stream", "slab", "wide"):   # the streamed kernel (default) and the two round-1 kernels for N > 128, at every size
        p = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, MCR_TC_BAND_SCALE="32", MCR_LARGE_N=large_n),
                           capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, (large_n, p.stderr[-2000:])
        rechecked = {int(l.split()[0]): int(l.split()[1]) for l in p.stdout.strip().splitlines()}
        assert rechecked[50] > 2000 and rechecked[128] > 5000, rechecked   # the wide band was in force
        assert rechecked[160] > 400 and rechecked[600] > 2000 and rechecked[1000] > 1000, (large_n, rechecked)


def test_error_behaviour():
    e = api.Engine(0)
    with pytest.raises(api.McrError, match="no model uploaded"):
        e.run(10, 1, 0.02, 1)
    with pytest.raises(api.McrError, match="not finite"):
        e.upload_model(np.array([np.nan, 0.0]), np.eye(2))
    with pytest.raises(api.McrError, match="a_is_lower"):
        e.upload_model(np.zeros(2), np.array([[1.0, 0.5], [0.0, 1.0]]), a_is_lower=True)
    with pytest.raises(api.McrError, match="out of range"):
        api.Engine(99)
    e.close()
