"""On-disk model format and path sharding — host-side mirror of the reference's loader.

File layout (little-endian), as written by the reference's preprocessing/save_binary.py:23-45 and
read by src/utilities/load_binary.c:5-124:
    mu.bin     [int32 N][N x float64]
    sigma.bin  [int32 rows][int32 cols][rows*cols x float64, row-major]
    params.bin [float64 actual_freq]
Unlike the reference loader, lengths are checked against the headers and mu/sigma must agree on N
(the C twin of this check is src/utilities/model_loader.c).
"""
from pathlib import Path

import numpy as np


def save_model(directory, mu, Sigma, actual_freq=0.0):
    d = Path(directory)
    d.mkdir(parents=True, exist_ok=True)
    mu = np.ascontiguousarray(mu, dtype="<f8")
    Sigma = np.ascontiguousarray(Sigma, dtype="<f8")
    with open(d / "mu.bin", "wb") as f:
        np.array([mu.shape[0]], dtype="<i4").tofile(f)
        mu.tofile(f)
    with open(d / "sigma.bin", "wb") as f:
        np.array(Sigma.shape, dtype="<i4").tofile(f)
        Sigma.tofile(f)
    with open(d / "params.bin", "wb") as f:
        np.array([actual_freq], dtype="<f8").tofile(f)


def load_model(directory):
    d = Path(directory)
    raw = (d / "mu.bin").read_bytes()
    if len(raw) < 4:
        raise ValueError("mu.bin: missing header")
    N = int(np.frombuffer(raw[:4], dtype="<i4")[0])
    if N <= 0 or len(raw) != 4 + 8 * N:
        raise ValueError(f"mu.bin: header says N={N} but the file has {len(raw)} bytes")
    mu = np.frombuffer(raw[4:], dtype="<f8").copy()
    raw = (d / "sigma.bin").read_bytes()
    if len(raw) < 8:
        raise ValueError("sigma.bin: missing header")
    r, c = (int(v) for v in np.frombuffer(raw[:8], dtype="<i4"))
    if r != N or c != N:
        raise ValueError(f"sigma.bin is {r}x{c} but mu.bin has N={N}")
    if len(raw) != 8 + 8 * r * c:
        raise ValueError(f"sigma.bin: wrong length for a {r}x{c} matrix")
    Sigma = np.frombuffer(raw[8:], dtype="<f8").reshape(r, c).copy()
    p = d / "params.bin"
    freq = float(np.frombuffer(p.read_bytes()[:8], dtype="<f8")[0]) if p.exists() else float("nan")
    return mu, Sigma, freq


def shard_ranges(M, n_shards):
    """Contiguous (offset, count) path ranges, sizes differing by at most one — the same split
    mcr_simulate_host applies over devices.  Any split gives identical totals (Philox is keyed by the
    absolute path index)."""
    out, off = [], 0
    for g in range(n_shards):
        c = M // n_shards + (1 if g < M % n_shards else 0)
        out.append((off, c))
        off += c
    return out
