"""K3 timing probe: one store, several environment-knob configurations, CUDA-event time per
batched search.  Diagnostic only (some knobs produce unusable results on purpose).

    python tools/k3_probe.py [--rows 10000000] [--dim 128] [--batch 1024] NAME:KNOB=V,KNOB=V ...

e.g.  base:  cg2:B200VS_GEMM_CG=2  nop:B200VS_GEMM_DBGMODE=3   (DBGMODE needs a B200VS_DEBUG_BUILD=1 library)
"""
import argparse
import ctypes as C
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "mlx-vector-db_b200"))
from b200vs import _cabi  # noqa: E402
from b200vs.sharded import ShardedVectorStore  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--mode", default="gemm_nocert")
ap.add_argument("--metric", default="cosine")
ap.add_argument("--iters", type=int, default=6)
ap.add_argument("--rounds", type=int, default=6)
ap.add_argument("configs", nargs="*", default=["base:"])
args = ap.parse_args()

lib = _cabi.lib()
dev = torch.device("cuda", 0)
st = ShardedVectorStore(args.dim, args.metric, device=dev, max_vectors_per_shard=args.rows + 16, search_mode=args.mode)
g = torch.Generator(device=dev)
g.manual_seed(1234)
step = 1_000_000
for r0 in range(0, args.rows, step):
    st.add_vectors(torch.randn((min(step, args.rows - r0), args.dim), device=dev, generator=g))
q = torch.randn((args.batch, args.dim), device=dev, generator=g)
torch.cuda.synchronize()
lib.vs_profile(1)

# Configurations are measured round-robin (`--rounds` times each): the box drifts by 10-15 % as it
# warms up and hits its power cap, so back-to-back blocks of one configuration are not comparable.
import statistics
import subprocess


def sm_clock():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                              capture_output=True, text=True, timeout=5).stdout.strip()
    except Exception:
        return "?"


results = {}
for rnd in range(args.rounds):
    order = args.configs[rnd % len(args.configs):] + args.configs[:rnd % len(args.configs)]   # rotate the start
    for cfg in order:
        name, _, kv = cfg.partition(":")
        knobs = dict(x.split("=", 1) for x in kv.split(",") if x)
        for k_, v in knobs.items():
            os.environ[k_] = v
        try:
            for _ in range(2):
                st.search(q, args.k)
            torch.cuda.synchronize()
            fr0 = int(lib.vs_fallback_count(st.shard.handle)), int(lib.vs_retry_count(st.shard.handle))
            ms0, n0 = C.c_double(), C.c_int64()
            lib.vs_profile_read(1, C.byref(ms0), C.byref(n0))          # slot 1 = K3 launches (kProfGemm)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                ids, sc = st.search(q, args.k)
            e1.record()
            torch.cuda.synchronize()
            ms1, n1 = C.c_double(), C.c_int64()
            lib.vs_profile_read(1, C.byref(ms1), C.byref(n1))
            per = e0.elapsed_time(e1) / args.iters
            kms = ms1.value / args.iters      # vs_profile_read clears the record on every read
            chk = int(ids.to(torch.int64).sum().item())
            fr1 = int(lib.vs_fallback_count(st.shard.handle)), int(lib.vs_retry_count(st.shard.handle))
            results.setdefault(cfg, []).append((per, kms, chk, n1.value / args.iters,
                                                (fr1[0] - fr0[0]) / args.iters, (fr1[1] - fr0[1]) / args.iters))
        except Exception as e:  # noqa: BLE001
            print(f"{name:14s} ERROR {e}", flush=True)
        for k_ in knobs:
            os.environ.pop(k_, None)
    if rnd == args.rounds - 1:
        print(f"# sm clock / power right after the last round: {sm_clock()}", flush=True)
fb = int(lib.vs_fallback_count(st.shard.handle)), int(lib.vs_retry_count(st.shard.handle))
for cfg, rows in results.items():
    name, _, kv = cfg.partition(":")
    per = [r[0] for r in rows]
    kms = [r[1] for r in rows]
    print(f"{name:14s} search min {min(per):7.3f} med {statistics.median(per):7.3f} ms   K3 kernels min {min(kms):7.3f} "
          f"med {statistics.median(kms):7.3f} ms ({rows[0][3]:.1f} launches)  qps(med) {args.batch / statistics.median(per) * 1e3:9.0f}"
          f"  idsum {rows[-1][2]}  fb/retry per search {rows[-1][4]:.1f}/{rows[-1][5]:.1f}  {kv}", flush=True)
print(f"# fallbacks {fb[0]} retries {fb[1]}")
