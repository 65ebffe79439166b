"""The C-ABI library loads and exports exactly what include/mcr_b200.h declares (no GPU needed);
compute entry points fail loudly — never fall back — when no CUDA device is usable."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from mcr_b200 import _lib, api

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "mcr_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mcr_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == names, "python binding and header disagree"


def test_abi_version_and_struct_layout():
    lib = _lib.load()
    assert lib.mcr_abi_version() == 1
    assert lib.mcr_abi_sizeof(0) == C.sizeof(_lib.RunOptions)
    assert lib.mcr_abi_sizeof(1) == C.sizeof(_lib.Result)
    assert lib.mcr_abi_sizeof(2) == C.sizeof(_lib.HostRequest)
    assert lib.mcr_abi_sizeof(99) == -1


def test_build_transform_host_only(oracle):
    """mcr_build_transform is pure host code: identical (bitwise) to the oracle's Cholesky."""
    import oracle_lib
    g = np.load(ROOT / "tests" / "golden" / "pyref_n15.npz")
    for mode in (api.TRANSFORM_LOWER, api.TRANSFORM_GSL_FULL):
        A = api.build_transform(g["Sigma"], mode)
        np.testing.assert_array_equal(A, oracle.transform(g["Sigma"], mode))
    from mcr_b200.synthetic import synthetic_model
    for N in (25, 50, 100, 600):
        m = synthetic_model(N)
        A = api.build_transform(m.Sigma, api.TRANSFORM_LOWER)
        np.testing.assert_allclose(A @ A.T, m.Sigma, rtol=1e-10, atol=1e-18)
    with pytest.raises(api.McrError, match="not positive definite"):
        api.build_transform(-np.eye(3))
    with pytest.raises(api.McrError, match="unknown mode"):
        api.build_transform(np.eye(3), 7)


def test_result_from_partials_matches_reference_statistics(oracle):
    r = api.result_from_partials([100000000, 4001326, 0, 0, 0, 0, 0, 0])
    assert (r.P_hat, r.std_error, r.ci_lower, r.ci_upper) == oracle.stats(4001326, 100000000)
    with pytest.raises(api.McrError, match="overflow"):
        api.result_from_partials([10, 1, 1, 0, 1, 0, 0, 0])


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is for GPU-less hosts")
    with pytest.raises(api.McrError, match="no usable CUDA device"):
        api.device_count()
    with pytest.raises(api.McrError, match="no usable CUDA device"):
        api.Engine(0)
    with pytest.raises(api.McrError, match="no usable CUDA device"):
        api.simulate_host(np.zeros(2), np.eye(2), 10, 1, 0.02, 1)
