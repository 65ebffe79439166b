"""Seeded synthetic risk models of the shapes BASELINE.json names.

The reference ships data for 15 indices only (config.yaml:15-30); "Stoxx50 (50 assets)",
"EuroNext100 (100)", "Stoxx600 (600)" in BASELINE.json are shapes, so their covariance is
synthetic SPD with magnitudes matched to the real data:
daily vol 0.009-0.014, mu -5e-5..5e-4, a low-rank factor structure plus a common market
factor.  Reproducible: rng = default_rng(1000 + N).
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class RiskModel:
    N: int
    mu: np.ndarray      # (N,) float64
    Sigma: np.ndarray   # (N, N) float64, SPD
    x: float            # crash threshold (reference config.yaml: x)
    k: int              # min simultaneous crashes (reference config.yaml: k)
    seed: int = 67      # reference config.yaml: random_seed


def synthetic_model(N: int) -> RiskModel:
    rng = np.random.default_rng(1000 + N)
    F = max(3, N // 8)
    B = rng.standard_normal((N, F))
    C0 = B @ B.T / F
    d = np.sqrt(np.diag(C0))
    C0 = C0 / np.outer(d, d)
    C = 0.5 * C0 + 0.3 * np.ones((N, N)) + 0.2 * np.eye(N)
    sig = rng.uniform(0.009, 0.014, size=N)
    Sigma = np.outer(sig, sig) * C
    Sigma = 0.5 * (Sigma + Sigma.T)
    mu = rng.uniform(-5e-5, 5e-4, size=N)
    return RiskModel(N=N, mu=mu, Sigma=Sigma, x=0.02, k=max(1, round(N / 3)))
