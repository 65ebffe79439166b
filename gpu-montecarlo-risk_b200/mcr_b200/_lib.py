"""ctypes binding of libmcr_b200.so (include/mcr_b200.h).  No fallback: if the shared
library has not been built, importing the binding raises."""
import ctypes as C
import os
from pathlib import Path

PKG_ROOT = Path(__file__).resolve().parent.parent
LIB_PATH = Path(os.environ.get("MCR_B200_LIB", PKG_ROOT / "libmcr_b200.so"))  # override: kernel-tuning builds

MCR_MAX_HIST = 1025
MCR_PORTFOLIO_BINS = 2048
MCR_PARTIAL_WORDS = 8
MCR_EXT_COUNT_WORDS = MCR_MAX_HIST + 1 + MCR_PORTFOLIO_BINS
MCR_EXT_SUM_WORDS = 3 + MCR_PORTFOLIO_BINS

TRANSFORM_LOWER, TRANSFORM_GSL_FULL = 0, 1
KERNEL_AUTO, KERNEL_FFMA, KERNEL_TENSOR, KERNEL_EXACT = 0, 1, 2, 3
KERNEL_NAMES = {0: "auto", 1: "ffma", 2: "tensor", 3: "exact"}


class RunOptions(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("kernel", C.c_int32),
        ("no_recheck", C.c_int32),
        ("want_hist", C.c_int32),
        ("want_portfolio", C.c_int32),
        ("sm_reserve", C.c_int32),
        ("portfolio_loss_threshold", C.c_double),
        ("d_S_values", C.c_void_p),
    ]


class Result(C.Structure):
    _fields_ = [
        ("n_paths", C.c_uint64),
        ("count", C.c_uint64),
        ("sum_S", C.c_uint64),
        ("n_rechecked", C.c_uint64),
        ("P_hat", C.c_double),
        ("std_error", C.c_double),
        ("ci_lower", C.c_double),
        ("ci_upper", C.c_double),
        ("device_ms", C.c_float),
        ("kernel_used", C.c_int32),
        ("padded_N", C.c_int32),
        ("launches", C.c_int32),
        ("hist_S", C.c_uint64 * MCR_MAX_HIST),
        ("portfolio_hits", C.c_uint64),
        ("portfolio_sum", C.c_double),
        ("portfolio_sumsq", C.c_double),
        ("portfolio_tail_loss_sum", C.c_double),
        ("portfolio_hist_lo", C.c_double),
        ("portfolio_hist_hi", C.c_double),
        ("portfolio_hist", C.c_uint64 * MCR_PORTFOLIO_BINS),
        ("portfolio_hist_sum", C.c_double * MCR_PORTFOLIO_BINS),
    ]


class HostRequest(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("N", C.c_int32),
        ("k", C.c_int32),
        ("M", C.c_uint64),
        ("seed", C.c_uint64),
        ("x", C.c_double),
        ("mu", C.c_void_p),
        ("A", C.c_void_p),
        ("a_is_lower", C.c_int32),
        ("n_gpus", C.c_int32),
        ("first_device", C.c_int32),
        ("want_S_values", C.c_int32),
        ("reupload", C.c_int32),
        ("reserved1", C.c_int32),
        ("S_values_out", C.c_void_p),
        ("opt", RunOptions),
    ]


# every symbol include/mcr_b200.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "mcr_abi_version": (C.c_int, []),
    "mcr_abi_sizeof": (C.c_int, [C.c_int]),
    "mcr_last_error": (C.c_char_p, []),
    "mcr_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "mcr_build_transform": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mcr_host_alloc_pinned": (C.c_void_p, [C.c_size_t]),
    "mcr_host_free_pinned": (None, [C.c_void_p]),
    "mcr_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "mcr_destroy": (C.c_int, [C.c_void_p]),
    "mcr_upload_model": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "mcr_set_portfolio_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "mcr_device_moments": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "mcr_device_transform": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mcr_run": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_int,
                          C.POINTER(RunOptions), C.POINTER(Result)]),
    "mcr_run_async": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_int,
                                C.POINTER(RunOptions), C.c_void_p, C.c_void_p]),
    "mcr_run_async_ext": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_int,
                                    C.POINTER(RunOptions), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mcr_portfolio_hist_range": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mcr_result_from_partials_ext": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.POINTER(Result)]),
    "mcr_last_main_kernel_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "mcr_last_mma_flops_per_trial": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "mcr_result_from_partials": (C.c_int, [C.c_void_p, C.POINTER(Result)]),
    "mcr_run_with_normals": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_double, C.c_int,
                                       C.POINTER(RunOptions), C.c_void_p, C.POINTER(Result)]),
    "mcr_run_with_normals_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_double, C.c_int,
                                              C.POINTER(RunOptions), C.POINTER(Result)]),
    "mcr_dump_normals": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "mcr_simulate_host": (C.c_int, [C.POINTER(HostRequest), C.POINTER(Result)]),
    "mcr_portfolio_var_es": (C.c_int, [C.POINTER(Result), C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mcr_measure_fp32_peak": (C.c_int, [C.c_void_p, C.c_float, C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """Load libmcr_b200.so once and attach prototypes.  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C gpu-montecarlo-risk_b200 lib`). There is no CPU fallback.")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class McrError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise McrError(load().mcr_last_error().decode("utf-8", "replace"))
