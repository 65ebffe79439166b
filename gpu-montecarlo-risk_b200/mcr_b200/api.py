"""Thin object wrapper over the C ABI — used by tests, bench.py and smoke(); mirrors the calls
the C plugin makes (src/06-B200/monte_carlo_b200.c)."""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import (KERNEL_AUTO, KERNEL_EXACT, KERNEL_FFMA, KERNEL_NAMES, KERNEL_TENSOR, MCR_EXT_COUNT_WORDS,
                   MCR_EXT_SUM_WORDS, MCR_MAX_HIST, MCR_PARTIAL_WORDS, TRANSFORM_GSL_FULL, TRANSFORM_LOWER, HostRequest,
                   McrError, Result, RunOptions, check)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def device_count() -> int:
    n = C.c_int(0)
    check(_lib.load().mcr_device_count(C.byref(n)))
    return n.value


def build_transform(Sigma, mode=TRANSFORM_LOWER) -> np.ndarray:
    """Host fp64 Cholesky + the chosen upper-triangle convention (mcr_build_transform)."""
    Sigma = _f64(Sigma)
    N = Sigma.shape[0]
    A = np.zeros((N, N), dtype=np.float64)
    check(_lib.load().mcr_build_transform(Sigma.ctypes.data, N, N, int(mode), A.ctypes.data))
    return A


@dataclass
class RunResult:
    n_paths: int
    count: int
    sum_S: int
    n_rechecked: int
    P_hat: float
    std_error: float
    ci_lower: float
    ci_upper: float
    device_ms: float
    kernel: str
    padded_N: int
    launches: int
    S_values: np.ndarray = None
    portfolio: dict = None   # extension outputs of a want_portfolio run (see Engine.run)
    hist_S: np.ndarray = None  # hist_S[s] = #{trials with S == s} (want_hist)

    @staticmethod
    def from_c(r: Result, S_values=None):
        return RunResult(int(r.n_paths), int(r.count), int(r.sum_S), int(r.n_rechecked), r.P_hat, r.std_error,
                         r.ci_lower, r.ci_upper, float(r.device_ms), KERNEL_NAMES.get(r.kernel_used, "?"),
                         int(r.padded_N), int(r.launches), S_values)


def make_options(kernel=KERNEL_AUTO, no_recheck=False, d_S_values=0, portfolio_threshold=None, want_hist=False, sm_reserve=0) -> RunOptions:
    o = RunOptions()
    o.sm_reserve = int(sm_reserve)
    o.struct_size = C.sizeof(RunOptions)
    o.kernel = int(kernel)
    o.no_recheck = int(bool(no_recheck))
    o.d_S_values = d_S_values or None
    o.want_hist = int(bool(want_hist))
    if portfolio_threshold is not None:
        o.want_portfolio = 1
        o.portfolio_loss_threshold = float(portfolio_threshold)
    return o


def portfolio_outputs(r: Result, alphas=(0.95, 0.99)) -> dict:
    """Extension outputs (north-star A7): moments, tail-event count, and VaR / ES of the loss -w.R."""
    n = int(r.n_paths)
    out = {"hits": int(r.portfolio_hits), "mean": r.portfolio_sum / n, "tail_loss_sum": r.portfolio_tail_loss_sum,
           "var": r.portfolio_sumsq / n - (r.portfolio_sum / n) ** 2,
           "hist": np.ctypeslib.as_array(r.portfolio_hist).copy(), "hist_sum": np.ctypeslib.as_array(r.portfolio_hist_sum).copy(),
           "hist_lo": r.portfolio_hist_lo, "hist_hi": r.portfolio_hist_hi, "VaR": {}, "ES": {}}
    for a in alphas:
        v, e = C.c_double(), C.c_double()
        check(_lib.load().mcr_portfolio_var_es(C.byref(r), float(a), C.byref(v), C.byref(e)))
        out["VaR"][a], out["ES"][a] = v.value, e.value
    return out


class Engine:
    """One engine = one CUDA device (one process per GPU in the multi-GPU launch)."""

    def __init__(self, device=0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        check(self._lib.mcr_create(int(device), C.byref(self._h)))
        self.device = device
        self.N = 0

    def close(self):
        if self._h:
            self._lib.mcr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_model(self, mu, A, a_is_lower=False):
        mu, A = _f64(mu), _f64(A)
        N = mu.shape[0]
        assert A.shape == (N, N)
        check(self._lib.mcr_upload_model(self._h, N, mu.ctypes.data, A.ctypes.data, int(bool(a_is_lower))))
        self.N = N

    def upload_sigma(self, mu, Sigma, mode=TRANSFORM_LOWER):
        A = build_transform(Sigma, mode)
        self.upload_model(mu, A, a_is_lower=(mode == TRANSFORM_LOWER))
        return A

    def run(self, M, seed, x, k, path_offset=0, kernel=KERNEL_AUTO, no_recheck=False, portfolio_threshold=None,
            weights=None, want_hist=False) -> RunResult:
        """portfolio_threshold=x_p additionally reduces the portfolio return w.R (weights default 1/N):
        result.portfolio = {hits = #{w.R < -x_p}, mean, var, tail_loss_sum, hist..., VaR{alpha}, ES{alpha}}."""
        r = Result()
        o = make_options(kernel, no_recheck, portfolio_threshold=portfolio_threshold, want_hist=want_hist)
        if weights is not None:
            w = _f64(weights)
            check(self._lib.mcr_set_portfolio_weights(self._h, w.ctypes.data, w.shape[0]))
        check(self._lib.mcr_run(self._h, int(M), int(path_offset), int(seed), float(x), int(k), C.byref(o), C.byref(r)))
        res = RunResult.from_c(r)
        if portfolio_threshold is not None:
            res.portfolio = portfolio_outputs(r)
        if want_hist:
            res.hist_S = np.ctypeslib.as_array(r.hist_S)[: self.N + 1].astype(np.int64)
        return res

    def run_async(self, M, seed, x, k, d_partials_ptr, stream_ptr, path_offset=0, kernel=KERNEL_AUTO,
                  no_recheck=False, d_S_values=0, sm_reserve=0):
        """Enqueue on a caller-owned stream; partial words land in the caller's device buffer.  sm_reserve: SMs left free for
        a collective that runs next to the kernels."""
        o = make_options(kernel, no_recheck, d_S_values, sm_reserve=sm_reserve)
        check(self._lib.mcr_run_async(self._h, int(M), int(path_offset), int(seed), float(x), int(k), C.byref(o),
                                      C.c_void_p(int(d_partials_ptr)), C.c_void_p(int(stream_ptr) if stream_ptr else None)))

    def run_async_ext(self, M, seed, x, k, d_partials_ptr, d_ext_counts_ptr, d_ext_sums_ptr, stream_ptr, path_offset=0,
                      kernel=KERNEL_AUTO, portfolio_threshold=None, want_hist=False):
        """run_async plus the extension partials of this shard (histogram of S, portfolio sums and histogram) in two
        caller-owned device buffers of MCR_EXT_COUNT_WORDS uint64 / MCR_EXT_SUM_WORDS float64 plain sums."""
        o = make_options(kernel, False, portfolio_threshold=portfolio_threshold, want_hist=want_hist)
        check(self._lib.mcr_run_async_ext(self._h, int(M), int(path_offset), int(seed), float(x), int(k), C.byref(o),
                                          C.c_void_p(int(d_partials_ptr)), C.c_void_p(int(d_ext_counts_ptr)),
                                          C.c_void_p(int(d_ext_sums_ptr)), C.c_void_p(int(stream_ptr) if stream_ptr else None)))

    def portfolio_hist_range(self):
        lo, hi = C.c_double(), C.c_double()
        check(self._lib.mcr_portfolio_hist_range(self._h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def run_with_normals(self, Z, x, k, kernel=KERNEL_AUTO, no_recheck=False, want_S=True) -> RunResult:
        Z = _f64(Z)
        M = Z.shape[0]
        assert Z.ndim == 2 and Z.shape[1] == self.N
        S = np.zeros(M, dtype=np.int32) if want_S else None
        r = Result()
        o = make_options(kernel, no_recheck)
        check(self._lib.mcr_run_with_normals(self._h, M, Z.ctypes.data, float(x), int(k), C.byref(o),
                                             S.ctypes.data if want_S else None, C.byref(r)))
        return RunResult.from_c(r, S)

    def run_with_normals_device(self, d_Z_ptr, M, x, k, kernel=KERNEL_AUTO, no_recheck=False) -> RunResult:
        r = Result()
        o = make_options(kernel, no_recheck)
        check(self._lib.mcr_run_with_normals_device(self._h, int(M), C.c_void_p(int(d_Z_ptr)), float(x), int(k),
                                                    C.byref(o), C.byref(r)))
        return RunResult.from_c(r)

    def dump_normals(self, M, seed, path_offset=0, normals=True, uniforms=False):
        nb = (self.N + 3) // 4
        Z = np.zeros((M, self.N), dtype=np.float32) if normals else None
        U = np.zeros((M, nb * 4), dtype=np.uint32) if uniforms else None
        check(self._lib.mcr_dump_normals(self._h, int(M), int(path_offset), int(seed),
                                         Z.ctypes.data if normals else None, U.ctypes.data if uniforms else None))
        return Z, U

    def last_main_kernel_ms(self):
        """(average ms per launch of the path kernel in the last run, launches timed)."""
        ms, n = C.c_float(0), C.c_int(0)
        check(self._lib.mcr_last_main_kernel_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def device_moments(self, returns):
        """(mu, Sigma) = mean vector and sample covariance (ddof=1) of a T x N matrix of returns, computed on the device
        (the reference's preprocessing_stats.compute_stats)."""
        X = _f64(returns)
        T, N = X.shape
        mu, Sigma = np.zeros(N), np.zeros((N, N))
        check(self._lib.mcr_device_moments(self._h, X.ctypes.data, T, N, mu.ctypes.data, Sigma.ctypes.data))
        return mu, Sigma

    def device_transform(self, Sigma, mode=TRANSFORM_LOWER) -> np.ndarray:
        """build_transform on the device (same summation order: bitwise the host result)."""
        Sigma = _f64(Sigma)
        N = Sigma.shape[0]
        A = np.zeros((N, N))
        check(self._lib.mcr_device_transform(self._h, Sigma.ctypes.data, N, N, int(mode), A.ctypes.data))
        return A

    def mma_flops_per_trial(self) -> float:
        """fp16 MMA flops per trial executed by the tensor pipe in the last run (0 for the FFMA / fp64 kernels)."""
        f = C.c_double(0)
        check(self._lib.mcr_last_mma_flops_per_trial(self._h, C.byref(f)))
        return f.value

    def measure_fp32_peak(self, ms_target=20.0) -> float:
        t = C.c_double(0)
        check(self._lib.mcr_measure_fp32_peak(self._h, float(ms_target), C.byref(t)))
        return t.value


def result_from_partials(words) -> RunResult:
    w = np.ascontiguousarray(words, dtype=np.uint64)
    assert w.size == MCR_PARTIAL_WORDS
    r = Result()
    check(_lib.load().mcr_result_from_partials(w.ctypes.data, C.byref(r)))
    return RunResult.from_c(r)


def result_from_partials_ext(words, ext_counts, ext_sums, hist_range=(0.0, 0.0), N=None, portfolio=False) -> RunResult:
    """Result of (all-reduced) partial words plus extension partials: .hist_S and, with portfolio=True, .portfolio."""
    w = np.ascontiguousarray(words, dtype=np.uint64)
    c = np.ascontiguousarray(ext_counts, dtype=np.uint64)
    s = np.ascontiguousarray(ext_sums, dtype=np.float64)
    assert w.size == MCR_PARTIAL_WORDS and c.size == MCR_EXT_COUNT_WORDS and s.size == MCR_EXT_SUM_WORDS
    r = Result()
    check(_lib.load().mcr_result_from_partials_ext(w.ctypes.data, c.ctypes.data, s.ctypes.data, float(hist_range[0]),
                                                   float(hist_range[1]), C.byref(r)))
    res = RunResult.from_c(r)
    res.hist_S = np.ctypeslib.as_array(r.hist_S)[: (N + 1) if N is not None else MCR_MAX_HIST].astype(np.int64)
    if portfolio:
        res.portfolio = portfolio_outputs(r)
    return res


def simulate_host(mu, A, M, seed, x, k, a_is_lower=False, n_gpus=1, first_device=0, want_S_values=False,
                  kernel=KERNEL_AUTO, no_recheck=False, reupload=False, portfolio_threshold=None, want_hist=False) -> RunResult:
    """The one-shot call the C plugin makes: host buffers in, result (and optional S_values) out.  The extension outputs
    (want_hist, portfolio_threshold; equal weights) are combined over n_gpus devices like the counts."""
    mu, A = _f64(mu), _f64(A)
    req = HostRequest()
    req.struct_size = C.sizeof(HostRequest)
    req.N, req.k, req.M, req.seed, req.x = mu.shape[0], int(k), int(M), int(seed), float(x)
    req.mu, req.A = mu.ctypes.data, A.ctypes.data
    req.a_is_lower, req.n_gpus, req.first_device = int(bool(a_is_lower)), int(n_gpus), int(first_device)
    S = np.zeros(int(M), dtype=np.int32) if want_S_values else None
    req.want_S_values = int(bool(want_S_values))
    req.reupload = int(bool(reupload))
    req.S_values_out = S.ctypes.data if want_S_values else None
    req.opt = make_options(kernel, no_recheck, portfolio_threshold=portfolio_threshold, want_hist=want_hist)
    r = Result()
    check(_lib.load().mcr_simulate_host(C.byref(req), C.byref(r)))
    res = RunResult.from_c(r, S)
    if portfolio_threshold is not None:
        res.portfolio = portfolio_outputs(r)
    if want_hist:
        res.hist_S = np.ctypeslib.as_array(r.hist_S)[: mu.shape[0] + 1].astype(np.int64)
    return res
