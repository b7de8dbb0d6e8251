import sys, ctypes as C, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/mlx-vector-db_b200")
from b200vs import _cabi
from b200vs.sharded import ShardedVectorStore
lib = _cabi.lib(); dev = torch.device("cuda", 0)
for n, d in ((4_000_000, 128), (2_000_000, 384), (500_000, 1536)):
    st = ShardedVectorStore(d, "cosine", device=dev, max_vectors_per_shard=n + 16)
    rows = torch.randn((n, d), device=dev)
    st.add_vectors(rows); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        _cabi.check(lib.vs_reset(st.shard.handle)); st.total = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); st.add_vectors(rows); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    ld16 = (d + 63) // 64 * 64
    nbytes = n * (d * 8 + ld16 * 2 + 12)
    print(f"K1 {n}x{d}: {best:.3f} ms  {nbytes / best / 1e6:.0f} GB/s  ({nbytes / best / 1e6 / 6466.5:.2f} of HBM copy peak)")
    st.close(); del rows; torch.cuda.empty_cache()
