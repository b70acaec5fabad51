"""Seeded synthetic Food.com-shaped rating lists.

The reference ships no data (``data/raw/.gitkeep`` only) and no network is
available, so every parity test and benchmark runs on ratings drawn here.  The
shape follows the reference's exploratory plots: power-law user activity and
item popularity, a rating histogram heavily skewed to 5 stars, duplicates
allowed, rows in generation (unsorted) order -- the order matters because the
reference's ``_build_index_lists`` (poisson_mf_cavi.py:73-84) keeps each row's
observations in original DataFrame order.

Pure host NumPy; shared by the engine's benchmarks, the tests and the oracle.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

RATING_PMF = (0.05, 0.01, 0.02, 0.05, 0.17, 0.70)  # ratings 0..5, Food.com-like skew


@dataclass(frozen=True)
class Workload:
    """One BASELINE.json config: sizes + model + K."""

    name: str
    model: str
    n_users: int
    n_items: int
    nnz: int
    n_factors: int
    seed: int


# BASELINE.json "configs", in order (C1..C5 in SURVEY.md §8).
WORKLOADS = {
    "c1": Workload("c1", "gaussian_mf", 20_000, 10_000, 200_000, 10, 20261),
    "c2": Workload("c2", "poisson_mf", 200_000, 230_000, 1_100_000, 50, 20262),
    "c3": Workload("c3", "hpf_cavi", 200_000, 230_000, 1_100_000, 50, 20263),
    "c4": Workload("c4", "hpf_pytorch", 200_000, 230_000, 1_100_000, 100, 20264),
    "c5": Workload("c5", "hpf_cavi", 2_000_000, 500_000, 100_000_000, 64, 20265),
}


def _power_law_ids(rng, n, count, gamma, chunk=1 << 24):
    out = np.empty(count, dtype=np.int32)
    for s in range(0, count, chunk):
        e = min(s + chunk, count)
        r = rng.random(e - s)  # float64 uniform
        np.power(r, gamma, out=r)
        r *= n
        ids = r.astype(np.int64)
        np.minimum(ids, n - 1, out=ids)
        out[s:e] = ids
    return out


def make_ratings(n_users, n_items, nnz, seed, gamma_u=2.0, gamma_i=2.5):
    """Return (u int32[nnz], i int32[nnz], rating float32[nnz]) with ratings in 0..5.

    ``u = min(floor(N * U1**gamma_u), N-1)`` relabelled by a random permutation so
    heavy rows are not contiguous; likewise for items.  The last observation is
    (N-1, M-1) so the reference's ``max(id)+1`` size inference
    (poisson_mf_cavi.py:44-46) yields exactly (N, M).
    """
    if nnz < 1:
        raise ValueError("nnz must be >= 1")
    rng = np.random.default_rng(seed)
    u = _power_law_ids(rng, n_users, nnz, gamma_u)
    i = _power_law_ids(rng, n_items, nnz, gamma_i)
    relabel_u = rng.permutation(n_users).astype(np.int32)
    relabel_i = rng.permutation(n_items).astype(np.int32)
    u = relabel_u[u]
    i = relabel_i[i]
    u[-1] = n_users - 1
    i[-1] = n_items - 1
    cdf = np.cumsum(np.asarray(RATING_PMF, dtype=np.float64))
    cdf[-1] = 1.0
    x = np.empty(nnz, dtype=np.float32)
    chunk = 1 << 24
    for s in range(0, nnz, chunk):
        e = min(s + chunk, nnz)
        x[s:e] = np.searchsorted(cdf, rng.random(e - s), side="right").astype(np.float32)
    return u, i, x


def make_splits(n_users, n_items, nnz, seed, val_frac=0.10, test_frac=0.05):
    """Train / validation / test triples drawn from the same generator stream.

    Validation and test ids may fall outside the train-inferred (N, M) only if
    they exceed max(train id); with the pinned last train row they never do, so
    tests that need out-of-range ids add them explicitly.
    """
    n_val = max(1, int(nnz * val_frac))
    n_test = max(1, int(nnz * test_frac))
    u, i, x = make_ratings(n_users, n_items, nnz + n_val + n_test, seed)
    tr = (u[:nnz].copy(), i[:nnz].copy(), x[:nnz].copy())
    tr[0][-1] = n_users - 1
    tr[1][-1] = n_items - 1
    va = (u[nnz:nnz + n_val], i[nnz:nnz + n_val], x[nnz:nnz + n_val])
    te = (u[nnz + n_val:], i[nnz + n_val:], x[nnz + n_val:])
    return tr, va, te


def to_frame(u, i, x):
    """DataFrame with the reference's schema (load_data.py:93-105): columns u, i, rating."""
    import pandas as pd

    return pd.DataFrame({"u": u, "i": i, "rating": x})


def workload_ratings(name):
    w = WORKLOADS[name]
    return w, make_ratings(w.n_users, w.n_items, w.nnz, w.seed)
