"""Deterministic synthetic inputs for parity tests and the CPU baseline.
TEST INFRASTRUCTURE ONLY (see oracle/vs_oracle.py header).

The reference's own tests use unseeded `np.random.rand` / `np.random.normal`
(tests/test_integration.py:83, tests/demo.py:167, benchmarks/large_scale_benchmark.py:59-61);
these generators keep the two distributions and fix the seeds.
"""
from __future__ import annotations

import numpy as np

DB_SEED = 1234
QUERY_SEED = 4321


def make_db(n: int, d: int, dist: str = "normal", seed: int = DB_SEED) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if dist == "normal":      # i.i.d. N(0,1): cosine ~ N(0, 1/D), no structure
        return rng.standard_normal((n, d), dtype=np.float32)
    if dist == "uniform":     # the reference's `np.random.rand`: all-positive, scores cluster (near-ties)
        return rng.random((n, d), dtype=np.float32)
    raise ValueError(dist)


def make_queries(b: int, d: int, dist: str = "normal", seed: int = QUERY_SEED) -> np.ndarray:
    return make_db(b, d, dist, seed)


def make_adversarial(n: int, d: int, seed: int = 99):
    """DB with exact duplicates (ties), a zero row (1e-8 clamp), a tiny-norm row and a
    scaled copy (cosine tie, different L2); queries that equal stored rows (self-match pin,
    reference tests/test_integration.py:133-136)."""
    rng = np.random.default_rng(seed)
    db = rng.standard_normal((n, d), dtype=np.float32)
    if n >= 16:
        db[5] = db[2]            # exact duplicate -> tie, lower id first
        db[n - 1] = db[2]        # duplicate at the very end (tile tail)
        db[7] = 0.0              # zero row -> score 0 via the clamp, not NaN
        db[9] = db[3] * np.float32(2.0)   # same direction, different length
        db[11] = db[4] * np.float32(1e-12)  # norm below the clamp
    q = np.stack([db[2], db[3], db[0], rng.standard_normal(d, dtype=np.float32)]).astype(np.float32)
    return db, q
