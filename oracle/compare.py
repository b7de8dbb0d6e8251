"""Parity comparator implementing BASELINE.json's contract.  TEST INFRASTRUCTURE ONLY.

Contract (BASELINE.json north_star): "the fp32 path reproduces the reference top-k
ids exactly, except for ties within 1e-6 relative distance, with scores within 1e-5".

Interpretation used here (stated in DESIGN.md): two rows are *tied* when their oracle
scores differ by at most `tie_rtol * max(1, |s|)` -- cosine scores live on the unit
scale, so 1e-6 is relative to the normalised vectors; for euclidean / dot it is
relative to the score itself once |s| > 1.  Inside a tie group any order / member is
accepted; outside it ids must match position by position.
"""
from __future__ import annotations

from dataclasses import dataclass
import numpy as np


@dataclass
class ParityReport:
    queries: int
    positions: int
    id_exact: int          # positions where ids are identical
    id_tie_ok: int         # positions that differ but lie in an oracle tie group
    id_wrong: int          # real mismatches
    max_score_err: float
    first_failure: str = ""

    @property
    def ok(self) -> bool:
        return self.id_wrong == 0


def compare_topk(ref_ids, ref_scores, got_ids, got_scores, score_matrix,
                 tie_rtol: float = 1e-6, score_atol: float = 1e-5,
                 score_rtol: float = 1e-5) -> ParityReport:
    """`score_matrix` is the oracle's full (B, N) score array, used to look up the oracle
    score of an id the engine returned at a position where the ids differ."""
    ref_ids = np.asarray(ref_ids)
    got_ids = np.asarray(got_ids)
    ref_scores = np.asarray(ref_scores, dtype=np.float64)
    got_scores = np.asarray(got_scores, dtype=np.float64)
    assert ref_ids.shape == got_ids.shape, (ref_ids.shape, got_ids.shape)
    B, K = ref_ids.shape
    exact = tie_ok = wrong = 0
    max_err = 0.0
    first = ""
    for b in range(B):
        if K and len(set(got_ids[b].tolist())) != K:
            wrong += 1
            first = first or f"query {b}: duplicate ids {got_ids[b].tolist()}"
        for j in range(K):
            gi = int(got_ids[b, j])
            ri = int(ref_ids[b, j])
            if gi < 0 or gi >= score_matrix.shape[1]:
                wrong += 1
                first = first or f"query {b} rank {j}: id {gi} out of range"
                continue
            s_ref_rank = float(ref_scores[b, j])
            s_ref_of_got = float(score_matrix[b, gi])
            scale = max(1.0, abs(s_ref_rank))
            # score check is against the oracle score of the id actually returned
            err = abs(float(got_scores[b, j]) - s_ref_of_got)
            max_err = max(max_err, err)
            if err > score_atol + score_rtol * abs(s_ref_of_got):
                wrong += 1
                first = first or (f"query {b} rank {j}: score {got_scores[b, j]!r} vs oracle "
                                  f"{s_ref_of_got!r} for id {gi}")
                continue
            if gi == ri:
                exact += 1
            elif abs(s_ref_of_got - s_ref_rank) <= tie_rtol * scale:
                tie_ok += 1
            else:
                wrong += 1
                first = first or (f"query {b} rank {j}: id {gi} (oracle score {s_ref_of_got!r}) "
                                  f"!= oracle id {ri} (score {s_ref_rank!r})")
    return ParityReport(B, B * K, exact, tie_ok, wrong, max_err, first)


def recall_at_k(ref_ids, got_ids) -> float:
    ref_ids = np.asarray(ref_ids)
    got_ids = np.asarray(got_ids)
    hit = 0
    for r, g in zip(ref_ids, got_ids):
        hit += len(set(r.tolist()) & set(g.tolist()))
    return hit / max(1, ref_ids.size)
