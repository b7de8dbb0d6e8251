"""CPU model of K3's adaptive filter threshold (csrc/gemm_topk.cu `tau_select_kernel`, the hit
queues and the ladder warp of csrc/gemm_kernel.cuh).

The kernel's exactness rests on the certification, which only needs ONE property of the filter:
every row whose 16-bit key is at or above the FINAL threshold is among the candidates.  That (and
the "enough candidates" property the ladder is built for) must hold however the threads interleave:
thresholds are re-read late (stale), hit-queue entries may be dropped when a ring is full, only one
entry per (lane, 32-column chunk) is counted, CTAs walk the tiles in any order.  This model replays
the algorithm in NumPy with all of that randomised and asserts the invariants."""
import zlib

import numpy as np
import pytest

K_LEVELS = 12


def ladder_levels(sample_max: np.ndarray, ks: int) -> np.ndarray:
    """tau_select_kernel: level 0 = ks-th largest sample maximum, then ranks ceil(ks/2^j) down to 1,
    then extrapolated levels (spacing = half of lvl2 - lvl0, shrinking 5 % per level)."""
    L = np.full(K_LEVELS, np.inf, np.float32)
    if sample_max.size < ks:
        L[0] = -np.inf
        return L
    srt = np.sort(sample_max)[::-1]
    L[0] = srt[ks - 1]
    nl, rk = 1, ks
    for j in range(1, K_LEVELS):
        if rk <= 1:
            break
        rk = (rk + 1) >> 1
        L[j] = srt[rk - 1]
        nl = j + 1
    d = np.float32(0.5) * (L[2] - L[0]) if nl >= 3 else (L[1] - L[0] if nl == 2 else np.float32(0))
    for j in range(nl, K_LEVELS):
        d = np.float32(d * np.float32(0.95))
        L[j] = L[j - 1] + d if d > 0 else np.inf
    for j in range(1, K_LEVELS):
        if not L[j] > L[j - 1]:
            L[j] = np.inf
    return L


def run_filter(keys: np.ndarray, ks: int, rank: int, rng, tile=128, n_ctas=8, sample_tiles=24,
               drop_prob=0.3, stale_prob=0.5):
    """One query's pass 1 + pass 2.  Returns (final threshold, candidate row ids)."""
    n = keys.size
    n_tiles = (n + tile - 1) // tile
    full_tiles = n // tile
    s_tiles = min(sample_tiles, full_tiles)
    stride = max(1, full_tiles // max(1, s_tiles))
    sample_max = np.array([keys[t * stride * tile:(t * stride + 1) * tile].max() for t in range(s_tiles)], np.float32)
    L = ladder_levels(sample_max, ks)
    cnt = np.zeros(K_LEVELS, np.int64)
    tau = L[0]                                   # tau_cur (global)
    cands = []
    # every CTA has its own (possibly stale) view of tau; tiles are dealt round-robin and the CTAs
    # advance in a random interleaving
    view = np.full(n_ctas, L[0], np.float32)
    pos = np.zeros(n_ctas, np.int64)
    my = [list(range(c, n_tiles, n_ctas)) for c in range(n_ctas)]
    live = [c for c in range(n_ctas) if my[c]]
    while live:
        c = live[rng.integers(len(live))]
        t = my[c][pos[c]]
        pos[c] += 1
        if pos[c] == len(my[c]):
            live.remove(c)
        if rng.random() > stale_prob:            # the threshold is re-read one accumulator ahead: sometimes stale
            view[c] = max(view[c], tau)
        rows = np.arange(t * tile, min(n, (t + 1) * tile))
        for c0 in range(0, rows.size, 32):       # 32-column chunks
            chunk = rows[c0:c0 + 32]
            kv = keys[chunk]
            hit = kv >= view[c]
            if not hit.any():
                continue
            cands.extend(chunk[hit].tolist())    # every hit becomes a candidate
            # the ladder sees at most the chunk maximum, only for whole chunks, and only if the ring had room
            if chunk.size == 32 and rng.random() > drop_prob:
                m = kv.max()
                for j in range(1, K_LEVELS):
                    if m >= L[j]:
                        cnt[j] += 1
                        if cnt[j] == rank:
                            tau = max(tau, L[j])
    return np.float32(tau), np.array(sorted(set(cands)), np.int64)


@pytest.mark.parametrize("dist", ["normal", "uniform", "clustered", "ties"])
@pytest.mark.parametrize("ks,rank", [(16, 16), (16, 32), (150, 250)])
def test_filter_invariants_hold_under_any_interleaving(dist, ks, rank):
    rng = np.random.default_rng(zlib.crc32(f"{dist}-{ks}-{rank}".encode()))
    n = 60_000
    for trial in range(6):
        if dist == "normal":
            keys = rng.standard_normal(n).astype(np.float32) * np.float32(0.09)
        elif dist == "uniform":
            keys = (0.75 + 0.02 * rng.standard_normal(n)).astype(np.float32)
        elif dist == "clustered":                # a time-ordered ingest: the best rows sit at the end
            keys = np.sort(rng.standard_normal(n).astype(np.float32))
        else:                                    # many exact ties at the top
            keys = np.round(rng.standard_normal(n), 1).astype(np.float32)
        tau, cands = run_filter(keys, ks, rank, rng, drop_prob=rng.choice([0.0, 0.3, 0.9]),
                                stale_prob=rng.choice([0.0, 0.5, 0.95]), sample_tiles=max(24, 2 * ks))
        # 1. the certification's premise: no row outside the candidate set has a key at or above the final threshold
        outside = np.ones(n, bool)
        outside[cands] = False
        assert not (keys[outside] >= tau).any()
        # 2. the threshold never overshoots: at least ks rows reach it (ks from pass 1, `rank` >= ks from the ladder)
        assert (keys >= tau).sum() >= min(ks, rank)
        # 3. hence the ks best rows are all candidates
        best = np.argsort(-keys, kind="stable")[:ks]
        kth = keys[best[-1]]
        assert set(np.flatnonzero(keys > kth).tolist()) <= set(cands.tolist())


def test_ladder_is_ascending_and_starts_at_the_sample_bound():
    rng = np.random.default_rng(5)
    for ks in (7, 16, 150, 256):
        m = rng.standard_normal(4 * ks + 13).astype(np.float32)
        L = ladder_levels(m, ks)
        assert L[0] == np.sort(m)[::-1][ks - 1]
        fin = L[np.isfinite(L)]
        assert (np.diff(fin) > 0).all()
    # degenerate: all maxima equal -> no level above the first
    L = ladder_levels(np.full(100, 0.5, np.float32), 16)
    assert L[0] == np.float32(0.5) and np.isinf(L[1:]).all()
    # too few maxima -> -inf (the host never launches the filter like that: `sampled` is false)
    assert ladder_levels(np.ones(5, np.float32), 16)[0] == -np.inf


def test_ladder_tightens_the_threshold():
    """With a small sample the ladder must end well above pass 1's bound (that is its purpose)."""
    rng = np.random.default_rng(11)
    keys = rng.standard_normal(400_000).astype(np.float32)
    tau, cands = run_filter(keys, 16, 16, rng, n_ctas=16, sample_tiles=64, drop_prob=0.0, stale_prob=0.2)
    n_pass1_only = (keys >= np.sort(np.array([keys[t * 48 * 128:(t * 48 + 1) * 128].max() for t in range(64)]))[::-1][15]).sum()
    assert cands.size < n_pass1_only / 3, (cands.size, n_pass1_only)
    assert (keys >= tau).sum() >= 16
