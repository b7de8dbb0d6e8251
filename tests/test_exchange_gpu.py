"""GPU: the peer-memory candidate exchange (csrc/exchange.cu) on ONE device -- every emulated rank
pushes its (2, B, k) block into its slot of a single buffer (as it would into a peer's), then the
fused wait + merge kernel must give what K4 (vs_merge) and a NumPy merge give."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _numpy_merge(scores, ids, k, l2):
    G, B, _ = scores.shape
    out_s = np.zeros((B, k), np.float32)
    out_i = np.full((B, k), -1, np.int32)
    for b in range(B):
        s = scores[:, b, :].reshape(-1)
        i = ids[:, b, :].reshape(-1)
        live = i >= 0
        s, i = s[live], i[live]
        key = -s if l2 else s
        order = np.lexsort((i, -key))          # key desc, id asc
        order = order[:k]
        out_s[b, :len(order)] = s[order]
        out_i[b, :len(order)] = i[order]
    return out_s, out_i


@pytest.mark.parametrize("G,B,k", [(8, 1024, 10), (2, 1, 10), (4, 37, 1), (8, 130, 32), (3, 5, 7)])
@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_push_wait_merge_equals_k4_and_numpy(native_lib, G, B, k, metric):
    from b200vs import _cabi
    lib = native_lib
    l2 = metric == "euclidean"
    rng = np.random.default_rng(G * 1000 + B + k)
    scores = rng.random((G, B, k), dtype=np.float32)
    scores = np.round(scores, 2)                 # plenty of ties across ranks
    scores = np.sort(scores, axis=2)
    if not l2:
        scores = scores[:, :, ::-1].copy()
    ids = np.empty((G, B, k), np.int32)
    for g in range(G):
        ids[g] = g * 1_000_000 + rng.integers(0, 1_000_000, size=(B, k), dtype=np.int32)
    # some ranks hold fewer than k rows: trailing slots are (0, -1)
    ids[G - 1, :, k // 2:] = -1
    scores[G - 1, :, k // 2:] = 0
    dev = torch.device("cuda", 0)
    words = 2 * B * k
    block_words = (words + 3) // 4 * 4
    buf = torch.zeros((G * block_words + 64,), dtype=torch.int32, device=dev)
    flags = torch.zeros((G,), dtype=torch.int32, device=dev)
    counter = torch.zeros((4,), dtype=torch.int32, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    step = 7
    P1 = C.c_void_p * 1
    srcs = []
    for g in range(G):
        raw = torch.zeros((block_words,), dtype=torch.int32, device=dev)
        raw[:B * k] = torch.from_numpy(scores[g].reshape(-1).view(np.int32)).to(dev)
        raw[B * k:2 * B * k] = torch.from_numpy(ids[g].reshape(-1)).to(dev)
        srcs.append(raw)
        _cabi.check(lib.vs_exchange_push(0, C.c_void_p(raw.data_ptr()), block_words * 4,
                                         P1(buf.data_ptr() + g * block_words * 4), P1(flags.data_ptr() + g * 4),
                                         1, step, C.c_void_p(counter.data_ptr()), stream))
    out_s = torch.empty((B, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((B, k), dtype=torch.int32, device=dev)
    _cabi.check(lib.vs_exchange_wait_merge(0, _cabi.METRICS[metric], C.c_void_p(flags.data_ptr()), G, step,
                                           C.c_void_p(buf.data_ptr()), block_words, B, k,
                                           C.c_void_p(out_s.data_ptr()), C.c_void_p(out_i.data_ptr()), stream))
    k4_s = torch.empty((B, k), dtype=torch.float32, device=dev)
    k4_i = torch.empty((B, k), dtype=torch.int32, device=dev)
    _cabi.check(lib.vs_exchange_wait(0, C.c_void_p(flags.data_ptr()), G, step, stream))
    _cabi.check(lib.vs_merge(0, _cabi.METRICS[metric], C.c_void_p(buf.data_ptr()), C.c_void_p(buf.data_ptr() + 4 * B * k),
                             G, B, k, block_words, C.c_void_p(k4_s.data_ptr()), C.c_void_p(k4_i.data_ptr()), stream))
    torch.cuda.synchronize()
    assert (flags.cpu().numpy() == step).all() and int(counter[0]) == 0
    ref_s, ref_i = _numpy_merge(scores, ids, k, l2)
    np.testing.assert_array_equal(out_i.cpu().numpy(), ref_i)
    np.testing.assert_array_equal(out_s.cpu().numpy(), ref_s)
    np.testing.assert_array_equal(k4_i.cpu().numpy(), ref_i)
    np.testing.assert_array_equal(k4_s.cpu().numpy(), ref_s)
