"""ctypes wrapper of oracle/liboracle.so (the CPU restatement of the reference algorithm) and of
oracle/_ref/libref_models.so (the reference's own unmodified C models built against the GSL shim).
TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
LOWER, GSL_FULL = 0, 1
DIAG_SQRT = 2   # OR-ed into the mode: L_jj = sqrt(a_jj) (see mcr_oracle.c)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    def __init__(self):
        so = ORACLE_DIR / "liboracle.so"
        if not so.exists():
            subprocess.run(["make", "-C", str(ORACLE_DIR), "liboracle.so"], check=True, capture_output=True)
        self.lib = C.CDLL(str(so))
        self.lib.mcr_oracle_transform.restype = C.c_int

    def transform(self, Sigma, mode=LOWER):
        Sigma = np.ascontiguousarray(Sigma, np.float64)
        N = Sigma.shape[0]
        A = np.zeros((N, N))
        rc = self.lib.mcr_oracle_transform(_p(Sigma), N, int(mode), _p(A))
        if rc != 0:
            raise ValueError("Sigma is not positive definite")
        return A

    def run_mt19937(self, A, mu, k, x, M, seed, want_S=True, want_Z=False):
        A, mu = np.ascontiguousarray(A, np.float64), np.ascontiguousarray(mu, np.float64)
        N = mu.shape[0]
        S = np.zeros(M, np.int32) if want_S else None
        Z = np.zeros((M, N)) if want_Z else None
        cnt, ss = C.c_int64(), C.c_int64()
        rc = self.lib.mcr_oracle_run_mt19937(_p(A), _p(mu), N, int(k), C.c_double(x), C.c_int64(M), C.c_ulong(seed),
                                             _p(S), C.byref(cnt), C.byref(ss), _p(Z))
        assert rc == 0
        return dict(count=cnt.value, sum_S=ss.value, S_values=S, Z=Z)

    def run_normals(self, A, mu, k, x, Z):
        A, mu = np.ascontiguousarray(A, np.float64), np.ascontiguousarray(mu, np.float64)
        N = mu.shape[0]
        M = Z.shape[0]
        assert Z.shape == (M, N)
        S = np.zeros(M, np.int32)
        cnt, ss = C.c_int64(), C.c_int64()
        if Z.dtype == np.float32:
            Z = np.ascontiguousarray(Z)
            fn = self.lib.mcr_oracle_run_normals_f32
        else:
            Z = np.ascontiguousarray(Z, np.float64)
            fn = self.lib.mcr_oracle_run_normals
        rc = fn(_p(A), _p(mu), N, int(k), C.c_double(x), C.c_int64(M), _p(Z), _p(S), C.byref(cnt), C.byref(ss))
        assert rc == 0
        return dict(count=cnt.value, sum_S=ss.value, S_values=S)

    def portfolio(self, A, mu, w, Zf32):
        A, mu, w = (np.ascontiguousarray(v, np.float64) for v in (A, mu, w))
        Z = np.ascontiguousarray(Zf32, np.float32)
        M, N = Z.shape
        P = np.zeros(M)
        assert self.lib.mcr_oracle_portfolio_normals_f32(_p(A), _p(mu), _p(w), N, C.c_int64(M), _p(Z), _p(P)) == 0
        return P

    def stats(self, count, M):
        out = [C.c_double() for _ in range(4)]
        self.lib.mcr_oracle_stats(C.c_int64(count), C.c_int64(M), *[C.byref(o) for o in out])
        return tuple(o.value for o in out)

    def philox(self, ctr, key):
        ctr = np.asarray(ctr, np.uint32)
        key = np.asarray(key, np.uint32)
        out = np.zeros(4, np.uint32)
        self.lib.mcr_oracle_philox4x32_10(_p(ctr), _p(key), _p(out))
        return out

    def philox_uniforms(self, seed, path_offset, M, N):
        nb = (N + 3) // 4
        U = np.zeros((M, nb * 4), np.uint32)
        self.lib.mcr_oracle_philox_uniforms(C.c_uint64(seed), C.c_uint64(path_offset), C.c_int64(M), N, _p(U))
        return U

    def philox_normals(self, seed, path_offset, M, N):
        Z = np.zeros((M, N), np.float32)
        self.lib.mcr_oracle_philox_normals(C.c_uint64(seed), C.c_uint64(path_offset), C.c_int64(M), N, _p(Z))
        return Z

    def mt19937_raw(self, seed, n):
        out = np.zeros(n, np.uint32)
        self.lib.mcr_oracle_mt19937_raw(C.c_ulong(seed), n, _p(out))
        return out

    def mt19937_gaussians(self, seed, n):
        out = np.zeros(n)
        self.lib.mcr_oracle_mt19937_gaussians(C.c_ulong(seed), n, _p(out))
        return out


# ---- the reference's own compiled models (oracle/_ref) ------------------------------------------
class _GslBlock(C.Structure):
    _fields_ = [("size", C.c_size_t), ("data", C.POINTER(C.c_double))]


class _GslVector(C.Structure):
    _fields_ = [("size", C.c_size_t), ("stride", C.c_size_t), ("data", C.POINTER(C.c_double)),
                ("block", C.c_void_p), ("owner", C.c_int)]


class _GslMatrix(C.Structure):
    _fields_ = [("size1", C.c_size_t), ("size2", C.c_size_t), ("tda", C.c_size_t),
                ("data", C.POINTER(C.c_double)), ("block", C.c_void_p), ("owner", C.c_int)]


class RefParams(C.Structure):  # MonteCarloParams, reference src/model_interface.h:19-27
    _fields_ = [("x", C.c_double), ("N", C.c_int), ("k", C.c_int), ("mu", C.POINTER(_GslVector)),
                ("Sigma", C.POINTER(_GslMatrix)), ("M", C.c_int), ("random_seed", C.c_ulong)]


class RefResult(C.Structure):  # MonteCarloResult, reference src/model_interface.h:9-16
    _fields_ = [("P_hat", C.c_double), ("count", C.c_int), ("S_values", C.POINTER(C.c_int)),
                ("std_error", C.c_double), ("ci_lower", C.c_double), ("ci_upper", C.c_double)]


class _ModelFunctions(C.Structure):
    _fields_ = [("name", C.c_char_p), ("run_model", C.CFUNCTYPE(C.c_int, C.POINTER(RefParams), C.POINTER(RefResult)))]


def ref_models_path():
    return ORACLE_DIR / "_ref" / "libref_models.so"


def run_model_struct(getter, mu, Sigma, k, x, M, seed, keep_S=True):
    """Call a `ModelFunctions get_*_model(void)` plugin through the reference's own interface."""
    getter.restype = _ModelFunctions
    mf = getter()
    mu = np.ascontiguousarray(mu, np.float64)
    Sigma = np.ascontiguousarray(Sigma, np.float64)
    N = mu.shape[0]
    v = _GslVector(N, 1, mu.ctypes.data_as(C.POINTER(C.c_double)), None, 0)
    m = _GslMatrix(N, N, N, Sigma.ctypes.data_as(C.POINTER(C.c_double)), None, 0)
    p = RefParams(x, N, k, C.pointer(v), C.pointer(m), M, seed)
    r = RefResult()
    rc = mf.run_model(C.byref(p), C.byref(r))
    out = dict(rc=rc, name=mf.name.decode(), P_hat=r.P_hat, count=r.count, std_error=r.std_error,
               ci_lower=r.ci_lower, ci_upper=r.ci_upper, S_values=None)
    if r.S_values:
        if keep_S and rc == 0:
            out["S_values"] = np.ctypeslib.as_array(r.S_values, shape=(M,)).copy()
        C.CDLL(None).free(r.S_values)  # the driver's make_clean (main_runner.c:47-52)
    return out


class RefModels:
    """get_serial_model / get_openmp_model of the reference, compiled unmodified (oracle/_ref)."""

    def __init__(self):
        self.lib = C.CDLL(str(ref_models_path()))

    def serial(self, *a, **kw):
        return run_model_struct(self.lib.get_serial_model, *a, **kw)

    def openmp(self, *a, **kw):
        return run_model_struct(self.lib.get_openmp_model, *a, **kw)


def ref_cuda_path():
    return ORACLE_DIR / "_ref" / "libref_cuda.so"


class RefCuda:
    """get_cuda_model of the reference (src/05-GPU/monte_carlos_cuda.cu), compiled UNMODIFIED for sm_100a against the GSL
    shim (oracle/Makefile): the GPU baseline.  N <= 64 only (monte_carlos_cuda.cu:78-81); always fills S_values."""

    def __init__(self):
        self.lib = C.CDLL(str(ref_cuda_path()))

    def cuda(self, *a, **kw):
        return run_model_struct(self.lib.get_cuda_model, *a, **kw)
