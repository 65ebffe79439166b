#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/.

Run ONCE in the build container, where the reference tree is mounted at /root/reference
(it does not exist on the GPU box, so nothing at test time reads it):

    python tests/golden/make_golden.py

What it produces, and from which reference code (paths relative to /root/reference):

  data15/{mu,sigma,params}.bin  the reference's own preprocessing pipeline
        (preprocessing/fetch_split.py, preprocessing_stats.py, evaluation.py,
        save_binary.py — imported unmodified) on the shipped dataset/ + config.yaml
        (15 indices, 2015-2025, train_ratio 0.8, x 0.02, k 5).
  pyref_n15.npz   a run of the reference's Python model
        (src/01-python-serial/python_serial.py:monte_carlo_crash_simulation, imported
        unmodified) on that data with numpy's legacy global RNG seeded; the normals it
        consumed are replayed from the same seed and stored beside its outputs, so the
        oracle / GPU path can be fed IDENTICAL normals.
  pyref_n25.npz   same on the synthetic N=25 model (SURVEY.md §8 row D recipe).
  mt19937_kat.json  raw MT19937 outputs from numpy's legacy init_genrand seeding
        (the same seeding GSL's gsl_rng_mt19937 uses) for the oracle's generator.
"""
import contextlib
import io
import json
import sys
from pathlib import Path

import numpy as np
import yaml

REF = Path("/root/reference")
HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(REF / "preprocessing"))
sys.path.insert(0, str(REF / "src" / "01-python-serial"))
sys.path.insert(0, str(HERE.parent.parent / "gpu-montecarlo-risk_b200"))


def quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


def main():
    from fetch_split import fetch_and_split
    from preprocessing_stats import compute_stats
    from evaluation import compute_metrics
    from save_binary import save_results_binary
    from python_serial import monte_carlo_crash_simulation, compute_accuracy
    from mcr_b200.synthetic import synthetic_model

    cfg = yaml.safe_load(open(REF / "config.yaml"))
    train, test = quiet(fetch_and_split, indices=cfg["indices"], start=cfg["start"], end=cfg["end"],
                        train_ratio=cfg["train_ratio"])
    mu, Sigma = quiet(compute_stats, train)
    actual_freq = quiet(compute_metrics, test, x=cfg["x"], k=cfg["k"])
    quiet(save_results_binary, mu, Sigma, actual_freq, HERE / "data15")
    print(f"data15: N={len(mu)} train={train.shape} test={test.shape} actual_freq={actual_freq}")

    def pyref(name, mu, Sigma, x, k, M, seed):
        N = len(mu)
        np.random.seed(seed)
        P_hat, count, S_values = quiet(monte_carlo_crash_simulation, x, N, k, mu, Sigma, M)
        se, lo, hi = compute_accuracy(P_hat, M)
        np.random.seed(seed)  # replay the normals the model consumed (standard_normal(N) per trial)
        Z = np.stack([np.random.standard_normal(N) for _ in range(M)])
        np.savez_compressed(HERE / name, mu=mu, Sigma=Sigma, x=x, k=k, M=M, seed=seed, Z=Z,
                            S_values=S_values.astype(np.int32), count=count, P_hat=P_hat,
                            std_error=se, ci_lower=lo, ci_upper=hi)
        print(f"{name}: N={N} M={M} count={count} P_hat={P_hat} sumS={int(S_values.sum())}")

    pyref("pyref_n15.npz", mu, Sigma, float(cfg["x"]), int(cfg["k"]), 4000, 20251017)
    m25 = synthetic_model(25)
    pyref("pyref_n25.npz", m25.mu, m25.Sigma, m25.x, m25.k, 2000, 1234)

    kat = {}
    for seed in (67, 4357, 1, 20251017):
        mt = np.random.MT19937()
        mt._legacy_seeding(seed)
        raw = mt.random_raw(1000)
        kat[str(seed)] = {"first8": [int(v) for v in raw[:8]], "n1000": int(raw[999])}
    json.dump(kat, open(HERE / "mt19937_kat.json", "w"), indent=1)
    print("mt19937_kat:", {k: v["n1000"] for k, v in kat.items()})


if __name__ == "__main__":
    main()
