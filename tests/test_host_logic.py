"""Host-side logic without a GPU: binary model format, plugin export and error path, sharding
arithmetic, and the N>1 reduction plumbing over gloo (world_size 2)."""
import ctypes as C
import os
import struct
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle_lib
from mcr_b200 import api, model_io
from mcr_b200.synthetic import synthetic_model

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "gpu-montecarlo-risk_b200"
GOLD = ROOT / "tests" / "golden"


def test_binary_format_round_trip(tmp_path):
    m = synthetic_model(25)
    model_io.save_model(tmp_path, m.mu, m.Sigma, 0.25)
    # exactly the layout preprocessing/save_binary.py:23-45 writes
    raw = (tmp_path / "mu.bin").read_bytes()
    assert struct.unpack("<i", raw[:4])[0] == 25 and len(raw) == 4 + 25 * 8
    raw = (tmp_path / "sigma.bin").read_bytes()
    assert struct.unpack("<ii", raw[:8]) == (25, 25) and len(raw) == 8 + 625 * 8
    mu, Sigma, freq = model_io.load_model(tmp_path)
    np.testing.assert_array_equal(mu, m.mu)
    np.testing.assert_array_equal(Sigma, m.Sigma)
    assert freq == 0.25


def test_reference_preprocessing_output_loads():
    mu, Sigma, freq = model_io.load_model(GOLD / "data15")
    g = np.load(GOLD / "pyref_n15.npz")
    np.testing.assert_array_equal(mu, g["mu"])
    np.testing.assert_array_equal(Sigma, g["Sigma"])
    assert freq == pytest.approx(0.014)


def test_loader_rejects_inconsistent_files(tmp_path):
    m = synthetic_model(8)
    model_io.save_model(tmp_path, m.mu, m.Sigma, 0.0)
    (tmp_path / "mu.bin").write_bytes((tmp_path / "mu.bin").read_bytes()[:-8])  # truncated
    with pytest.raises(ValueError):
        model_io.load_model(tmp_path)
    model_io.save_model(tmp_path, m.mu[:7], m.Sigma, 0.0)                       # N mismatch
    with pytest.raises(ValueError):
        model_io.load_model(tmp_path)


def test_c_loader_rejects_bad_files_before_touching_the_gpu(tmp_path):
    """model_loader.c (the hardened twin of the reference's src/utilities/load_binary.c:13-36,56-87) through mcr_cli: every header /
    length / dimension check runs before the first allocation, so the rejections need no GPU and name the offending file."""
    import subprocess
    exe = PKG / "bin" / "mcr_cli"
    if not exe.exists():
        pytest.skip("bin/mcr_cli not built")
    m = synthetic_model(8)

    def run(d):
        return subprocess.run([str(exe), str(d), "1000"], capture_output=True, text=True, timeout=60)

    cases = {}
    d = tmp_path / "trunc_mu"; d.mkdir()
    model_io.save_model(d, m.mu, m.Sigma, 0.0)
    (d / "mu.bin").write_bytes((d / "mu.bin").read_bytes()[:-8])
    cases["trunc_mu"] = (d, "mu.bin has the wrong length for N=8")
    d = tmp_path / "trunc_sigma"; d.mkdir()
    model_io.save_model(d, m.mu, m.Sigma, 0.0)
    (d / "sigma.bin").write_bytes((d / "sigma.bin").read_bytes()[:-16])
    cases["trunc_sigma"] = (d, "sigma.bin has the wrong length for a 8x8 matrix")
    d = tmp_path / "mismatch"; d.mkdir()
    model_io.save_model(d, m.mu[:7], m.Sigma, 0.0)
    cases["mismatch"] = (d, "is 8x8 but")
    d = tmp_path / "bad_header"; d.mkdir()
    model_io.save_model(d, m.mu, m.Sigma, 0.0)
    raw = bytearray((d / "mu.bin").read_bytes()); raw[0:4] = (-3).to_bytes(4, "little", signed=True)
    (d / "mu.bin").write_bytes(bytes(raw))
    cases["bad_header"] = (d, "Invalid dimension header")
    d = tmp_path / "no_sigma"; d.mkdir()
    model_io.save_model(d, m.mu, m.Sigma, 0.0)
    (d / "sigma.bin").unlink()
    cases["no_sigma"] = (d, "Cannot open")
    d = tmp_path / "padded"; d.mkdir()
    model_io.save_model(d, m.mu, m.Sigma, 0.0)
    (d / "sigma.bin").write_bytes((d / "sigma.bin").read_bytes() + b"\0" * 8)   # trailing garbage is a length error too
    cases["padded"] = (d, "wrong length")
    for name, (d, msg) in cases.items():
        p = run(d)
        assert p.returncode == 1 and msg in p.stderr, (name, p.returncode, p.stderr)


def test_shard_ranges_partition_the_path_range():
    for M, n in ((100, 1), (100, 3), (7, 8), (10**9, 8), (0, 4)):
        r = model_io.shard_ranges(M, n)
        assert len(r) == n and r[0][0] == 0 and sum(c for _, c in r) == M
        for (o0, c0), (o1, _) in zip(r[:-1], r[1:]):
            assert o0 + c0 == o1
        assert max(c for _, c in r) - min(c for _, c in r) <= 1


@pytest.mark.skipif(not (PKG / "libmcr_b200_model.so").exists(), reason="plugin not built (reference tree absent)")
def test_plugin_exports_get_cuda_model_and_fails_loudly_without_gpu():
    import torch
    lib = C.CDLL(str(PKG / "libmcr_b200_model.so"))
    assert hasattr(lib, "get_cuda_model")
    if torch.cuda.is_available():
        pytest.skip("GPU present: the success path is covered by the gpu tests")
    g = np.load(GOLD / "pyref_n15.npz")
    r = oracle_lib.run_model_struct(lib.get_cuda_model, g["mu"], g["Sigma"], 5, 0.02, 1000, 67)
    assert r["name"] == "cuda"
    assert r["rc"] == -1 and r["S_values"] is None  # error convention of the reference models: -1, nothing leaked


GLOO_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["MCR_PKG"]); sys.path.insert(0, os.environ["MCR_TESTS"])
import oracle_lib
from mcr_b200 import api, model_io
from mcr_b200.synthetic import synthetic_model
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
m = synthetic_model(16); o = oracle_lib.Oracle(); A = o.transform(m.Sigma, 0)
M, seed = 30001, 99
off, cnt = model_io.shard_ranges(M, world)[rank]
# the shard's partial words, computed by the ORACLE on the Philox stream of exactly this path range
Z = o.philox_normals(seed, off, cnt, 16)
r = o.run_normals(A, m.mu, m.k, m.x, Z)
words = torch.tensor([cnt, r["count"], r["sum_S"], 0, 0, 0, 0, 0], dtype=torch.int64)
dist.all_reduce(words)                      # the one cross-rank step (NCCL on GPUs, gloo here)
res = api.result_from_partials(words.numpy().astype(np.uint64))
Zall = o.philox_normals(seed, 0, M, 16); ref = o.run_normals(A, m.mu, m.k, m.x, Zall)
assert (res.n_paths, res.count, res.sum_S) == (M, ref["count"], ref["sum_S"]), (res, ref["count"])
assert (res.P_hat, res.std_error, res.ci_lower, res.ci_upper) == o.stats(ref["count"], M)
print(f"rank {rank} ok count={res.count}")
dist.destroy_process_group()
"""


def test_two_rank_reduction_over_gloo(tmp_path):
    """world_size 2 on CPU: disjoint Philox path ranges + sum all-reduce of the partial words give the
    single-shard result bit-for-bit (shard invariance of the counter layout, reduction plumbing)."""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, MCR_PKG=str(PKG), MCR_TESTS=str(ROOT / "tests"), MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577", str(script)],
                       env=env, capture_output=True, text=True, timeout=240)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "rank 0 ok" in p.stdout and "rank 1 ok" in p.stdout
