#!/usr/bin/env python
"""bench.py -- queries/sec of exact top-10 cosine search (BASELINE.json's metric).

    python bench.py [--gpus N --steps K --warmup W]            # this engine on N B200s
    python bench.py --impl reference [...]                      # reference arithmetic on host cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of B synthetic queries.  Headline
workload: 10M x 128 fp32 cosine top-10, batch 1024 (named by BASELINE.json's metric; the
multi-GPU config), row-sharded over the N ranks (strong scaling: the database is fixed).
Rank 0 prints ONE JSON line.  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for _p in (str(ROOT), str(ROOT / "mlx-vector-db_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

WORKLOADS = {          # name -> (rows, dim)
    "10Mx128": (10_000_000, 128),
    "1Mx1536": (1_000_000, 1536),
    "1Mx768": (1_000_000, 768),
    "5Mx384": (5_000_000, 384),
    "100Kx384": (100_000, 384),
    "1.25Mx128": (1_250_000, 128),   # one shard of the 8-GPU headline run (diagnostic)
}
N_BLOCKS = 8           # the database is generated in 8 seeded blocks so every N in {1,2,4,8} sees the same rows
DB_SEED, QUERY_SEED = 1234, 4321
METRIC_NAME = "queries/sec exact top-10 cosine"
_T0 = time.perf_counter()


def log(msg: str) -> None:
    """Progress to stderr (stdout carries the ONE JSON line): phase + seconds since start."""
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"[bench {time.perf_counter() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="10Mx128", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--mode", default="auto")
    ap.add_argument("--dist", default="normal", choices=["normal", "uniform"],
                    help="row / query distribution: N(0,1) or the reference tests' np.random.rand U[0,1)")
    ap.add_argument("--extras", type=int, default=-1,
                    help="also measure the other BASELINE shapes/batches (default: on at N=1)")
    ap.add_argument("--cpu-baseline", type=int, default=1)
    ap.add_argument("--clocks", default="smi", choices=["smi", "nvml", "off"],
                    help="how SM clocks / throttle reasons are sampled during the timed region")
    ap.add_argument("--verify", type=int, default=1)
    return ap.parse_args()


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], 0.0, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = max(smax, float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax or None,
                "power_w_max": max(power) if power else None, "samples": len(sm),
                "reasons": sorted(reasons)}


class NvmlSampler:
    """Same record through NVML from a background thread of this process (no nvidia-smi process
    polling the driver): SM clock, power and throttle reasons every 50 ms."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples, self.power, self.bits = [], [], 0
        self.smax = None
        self.stop_flag = False
        self.thread = None

    def start(self):
        import threading
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.thread = None
            return

        def loop():
            while not self.stop_flag:
                try:
                    self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                    try:
                        self.bits |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                    except Exception:
                        self.bits |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                except Exception:
                    pass
                time.sleep(0.05)

        self.thread = threading.Thread(target=loop, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        time.sleep(0.06)
        self.stop_flag = True
        self.thread.join(timeout=2)
        sm = sorted(self.samples)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.smax,
                "power_w_max": max(self.power) if self.power else None, "samples": len(sm),
                "reasons": sorted(v for b, v in self.REASONS.items() if self.bits & b), "source": "nvml"}


class NoSampler:
    def __init__(self, gpu_index): pass
    def start(self): pass
    def stop(self): return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling disabled"]}


def make_sampler(kind: str, gpu_index: int):
    return {"smi": ClockSampler, "nvml": NvmlSampler, "off": NoSampler}[kind](gpu_index)


# --------------------------------------------------------------------------- reference arm
def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def set_host_threads(threads: int):
    """torchrun exports OMP_NUM_THREADS=1 for nproc > 1, which would leave the BLAS of the CPU arm
    single-threaded at N >= 2: pin every native thread pool to the cores this process may use."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=threads)
    except Exception:
        return None


def cpu_sample_queries(threads: int, B: int) -> int:
    """Queries per CPU sample: one per host thread (the per-row argsort runs one row per thread, so
    fewer would leave cores idle and bias the extrapolation), the same at every N."""
    return max(1, min(B, threads))


def reference_step(db_np, q_np, k, sample_q, threads, return_scores=False):
    """One bounded sample of the reference's batch search on the host: the first `sample_q`
    queries of the batch against the FULL database, in the reference's op order (per-call
    database re-normalisation, fp32 GEMM, full stable argsort).  Returns (phase times, ids, scores)."""
    from oracle import vs_oracle
    ids, sc, t = vs_oracle.batch_similarity_search_timed(q_np[:sample_q], db_np, k, threads=threads,
                                                        return_scores=return_scores)
    return t, ids, sc


def extrapolate_qps(t: dict, sample_q: int, B: int) -> float:
    """Whole-batch throughput implied by a sample: the database normalisation is paid once per
    call, GEMM and argsort scale with the number of queries."""
    per_q = (t["matmul_s"] + t["argsort_s"]) / sample_q
    return B / (t["normalize_s"] + B * per_q)


def optimized_cpu_line(db_np, q_np, k, sample_q, threads, B):
    """BASELINE.md section 2's "optimised CPU" line: database normalised once (untimed), GEMM,
    argpartition to k + sort of k.  Context for how much of the reference's cost is algorithmic."""
    from oracle import vs_oracle
    vn = vs_oracle.normalize_vectors(db_np)
    vs_oracle.batch_similarity_search_optimized_timed(q_np[:sample_q], vn, k, threads)      # warm-up
    _, _, t = vs_oracle.batch_similarity_search_optimized_timed(q_np[:sample_q], vn, k, threads)
    per_q = (t["matmul_s"] + t["select_s"]) / sample_q
    return {"value": 1.0 / per_q, "unit": "queries/s", "cores": threads,
            "what": "pre-normalised database (once, untimed) + fp32 GEMM + argpartition(k) + sort of k",
            "sample": f"first {sample_q} of {B} queries vs the full database; GEMM {t['matmul_s']:.2f}s, "
                      f"select {t['select_s']:.2f}s"}


def verify_against_oracle(ref_ids, ref_scores, score_matrix, got_ids, got_scores):
    """--verify: the engine's results for the sampled queries against the oracle's, BASELINE.json's
    contract (ids exact outside 1e-6 ties, scores within 1e-5) -- oracle/compare.py."""
    from oracle import compare
    q = ref_ids.shape[0]
    rep = compare.compare_topk(ref_ids, ref_scores, got_ids[:q], got_scores[:q], score_matrix)
    return {"queries": int(q), "ok": bool(rep.ok), "ids_exact": int(rep.id_exact), "ids_tie_ok": int(rep.id_tie_ok),
            "ids_wrong": int(rep.id_wrong), "max_score_err": float(rep.max_score_err),
            "against": "NumPy oracle (reference op order) over the FULL database, tie rule of oracle/compare.py",
            "first_failure": rep.first_failure}


def make_host_data(n, d, B):
    import numpy as np
    blocks = []
    for b in range(N_BLOCKS):
        rng = np.random.default_rng(DB_SEED + b)
        blocks.append(rng.standard_normal((n // N_BLOCKS, d), dtype=np.float32))
    db = np.concatenate(blocks, axis=0)
    q = np.random.default_rng(QUERY_SEED).standard_normal((B, d), dtype=np.float32)
    return db, q


def run_reference(args):
    """--impl reference: the reference's own CPU arithmetic for this path (oracle port -- `mlx`
    is not installable here) on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, d = WORKLOADS[args.workload]
    B, k = args.batch, args.k
    threads = host_cores()
    limiter = set_host_threads(threads)
    db, q = make_host_data(n, d, B)
    sample_q = cpu_sample_queries(threads, B)
    # bound the run: warmup + steps samples within ~3 minutes
    t, _, _ = reference_step(db, q, k, sample_q, threads)
    one = t["normalize_s"] + t["matmul_s"] + t["argsort_s"]
    steps = max(1, min(args.steps, int(170.0 / max(one, 1e-3)) - 1))
    warm = max(0, min(args.warmup - 1, int(170.0 / max(one, 1e-3)) - 1 - steps))
    for _ in range(warm):
        reference_step(db, q, k, sample_q, threads)
    vals, wall = [], 0.0
    for _ in range(steps):
        t0 = time.perf_counter()
        t, _, _ = reference_step(db, q, k, sample_q, threads)
        wall += time.perf_counter() - t0
        vals.append(extrapolate_qps(t, sample_q, B))
    value = len(vals) / sum(1.0 / v for v in vals)
    sample = (f"per step: first {sample_q} of {B} queries (one per host thread) vs the full {n}x{d} database "
              f"(normalise DB + fp32 GEMM + full stable argsort); QPS extrapolated to the batch as "
              f"B/(t_normalise + B*(t_gemm+t_argsort)/{sample_q}); {steps} timed samples")
    line = {
        "impl": "reference", "metric": METRIC_NAME, "value": value, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / max(1, steps), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload} fp32 cosine top-{k}, batch {B}", "rows": n, "dim": d,
                   "batch": B, "k": k, "sample_queries_per_step": sample_q, "timed_samples": steps},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    del limiter


# --------------------------------------------------------------------------- extra configs (N = 1)
def config_c_l2_bf16(dev, lib, _cabi, ShardedVectorStore, args):
    """BASELINE config C: 1M x 1536 L2 top-100 over the bf16 shadow of the database with fp32
    rescoring of the candidates.  recall@100 is against the ORACLE's ids (direct-difference
    euclidean distance + stable argsort over a host copy of the same rows, first 8 queries);
    throughput for the tensor-core path (K3, certified: batches) and the scans (batch 1)."""
    import numpy as np
    import torch
    from oracle import vs_oracle
    n, d, k = 1_000_000, 1536, 100
    out = {"workload": "1Mx1536 L2 top-100, bf16 database + fp32 rescoring (config C)"}
    st = ShardedVectorStore(d, "euclidean", device=dev, shadow_bf16=True, max_vectors_per_shard=n + 16,
                            search_mode="auto")
    host_blocks = []
    for b in range(N_BLOCKS):
        g = torch.Generator(device=dev).manual_seed(DB_SEED + b)
        rows = torch.randn((n // N_BLOCKS, d), generator=g, device=dev, dtype=torch.float32)
        st.shard.append(rows, b * (n // N_BLOCKS))
        host_blocks.append(rows.cpu().numpy())
        del rows
    st.total = n
    q = torch.randn((1024, d), generator=torch.Generator().manual_seed(QUERY_SEED), dtype=torch.float32).to(dev)
    nq = 8
    q_np = q[:nq].cpu().numpy()
    dist_rows = np.concatenate([np.stack([vs_oracle.euclidean_distance(qi, blk) for qi in q_np]) for blk in host_blocks],
                               axis=1)                                   # (nq, n) oracle distances
    del host_blocks
    ref = np.argsort(dist_rows, axis=1, kind="stable")[:, :k]

    def recall(ids):
        got = ids[:nq].cpu().numpy()
        return sum(len(set(r.tolist()) & set(g_.tolist())) for r, g_ in zip(ref, got)) / ref.size

    def timed(B, mode, steps=10):
        st.shard.flags = _cabi.SEARCH_MODES[mode]
        qq = q[:B].contiguous()
        fb0 = int(lib.vs_fallback_count(st.shard.handle))
        for _ in range(3):
            ids, _ = st.search(qq, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pend = st.submit(qq, k)
        for _ in range(steps - 1):
            nxt = st.submit(qq, k)
            st.result(pend)
            pend = nxt
        ids, _ = st.result(pend)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        r = {"mode": mode, "batch": B, "qps": B / (ms * 1e-3), "ms_per_step": ms,
             "exact_fallback_queries_per_step": (int(lib.vs_fallback_count(st.shard.handle)) - fb0) / (steps + 3)}
        if B >= nq:
            r["recall_at_100_vs_oracle"] = recall(ids)
        if B == 1:
            bytes_per_row = d * (4 if mode == "scan_fp32" else 2)
            r["hbm_gbs_of_scanned_rows"] = n * bytes_per_row / (ms * 1e-3) / 1e9
        else:
            r["tflops"] = 2.0 * B * n * d / (ms * 1e-3) / 1e12
        return r

    out["runs"] = [timed(1024, "auto"), timed(16, "auto"), timed(8, "scan_bf16", 5),
                   timed(1, "scan_bf16", 20), timed(1, "gemm", 20), timed(1, "scan_fp32", 20)]
    out["note"] = ("auto = K3 (tcgen05 over the bf16 shadow, key 2 q.x - ||x||^2, certified, exact fp32 results) for "
                   "batches, the fp32 scan for single queries; scan_bf16 = K2 over the bf16 shadow + K5 (recall-reported)")
    st.close()
    torch.cuda.empty_cache()
    return out


def fp8_variant(dev, lib, _cabi, ShardedVectorStore, args):
    """fp8 (e4m3) database variant: K3 with kind::f8f6f4 MMAs over the e4m3 shadow + fp32 rescoring
    of 4x over-fetched candidates; reported as recall@10 against the engine's exact results."""
    import torch
    out = {"workload": "fp8 (e4m3) database variant, cosine top-10: recall@10 vs the exact path and QPS"}
    for name, batches in (("1Mx1536", (1024, 32)), ("10Mx128", (1024, 32))):
        n, d = WORKLOADS[name]
        st = ShardedVectorStore(d, "cosine", device=dev, shadow_bf16=3, max_vectors_per_shard=n + 16,
                                search_mode="auto")
        for b in range(N_BLOCKS):
            g = torch.Generator(device=dev).manual_seed(DB_SEED + b)
            rows = torch.randn((n // N_BLOCKS, d), generator=g, device=dev, dtype=torch.float32)
            st.shard.append(rows, b * (n // N_BLOCKS))
            del rows
        st.total = n
        for B in batches:
            q = torch.randn((B, d), generator=torch.Generator().manual_seed(QUERY_SEED), dtype=torch.float32).to(dev)
            st.shard.flags = _cabi.SEARCH_MODES["auto"]
            ref, _ = st.search(q, args.k)
            st.shard.flags = _cabi.SEARCH_MODES["gemm_fp8"]
            got, _ = st.search(q, args.k)
            r, g_ = ref.cpu().numpy(), got.cpu().numpy()
            hits = sum(len(set(a.tolist()) & set(b_.tolist())) for a, b_ in zip(r, g_))
            for _ in range(3):
                st.search(q, args.k)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            steps = 20
            e0.record()
            for _ in range(steps):
                st.search(q, args.k)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[f"{name} batch {B}"] = {"recall_at_10": hits / r.size, "qps": B / (ms * 1e-3), "ms_per_step": ms}
        st.close()
        torch.cuda.empty_cache()
    return out


def k1_append_throughput(dev, lib, _cabi, ShardedVectorStore, args):
    """K1 append_norm alone: device-resident rows appended into a fresh store; bytes = rows read
    + fp32 master written + 16-bit shadow written + norms (HBM-bound copy-like kernel)."""
    import torch
    out = {"workload": "K1 append_norm (device-resident rows -> arena, norms, 16-bit shadow)"}
    for n, d in ((4_000_000, 128), (500_000, 1536)):
        st = ShardedVectorStore(d, "cosine", device=dev, shadow_bf16=True, max_vectors_per_shard=n + 16,
                                search_mode=args.mode)
        rows = torch.randn((n, d), device=dev, dtype=torch.float32)
        st.add_vectors(rows)                       # maps the arena chunks (host-side VMM calls)
        torch.cuda.synchronize()
        _cabi.check(lib.vs_reset(st.shard.handle))  # forget the rows, keep the mapped arena
        st.total = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st.add_vectors(rows)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        ld16 = (d + 63) // 64 * 64
        nbytes = n * (d * 4 + d * 4 + ld16 * 2 + 12)
        out[f"{n}x{d}"] = {"ms": ms, "rows_per_s": n / (ms * 1e-3), "GBps": nbytes / (ms * 1e-3) / 1e9,
                           "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / 6466.5}
        st.close()
        del rows
        torch.cuda.empty_cache()
    return out


def config_e_streaming(dev, lib, _cabi, ShardedVectorStore, args):
    """BASELINE config E: 5M x 384 mixed streaming workload -- the store starts at 1M rows and
    grows to 5M by 10 000-row appends (K1, in-place arena growth), each followed by a batch-256
    top-10 query against everything appended so far."""
    import torch
    d, k, B, step_rows, n0, n1 = 384, 10, 256, 10_000, 1_000_000, 5_000_000
    st = ShardedVectorStore(d, "cosine", device=dev, shadow_bf16=True, max_vectors_per_shard=n1 + 16,
                            search_mode=args.mode)
    g = torch.Generator(device=dev).manual_seed(DB_SEED)
    first = torch.randn((n0, d), generator=g, device=dev, dtype=torch.float32)
    st.add_vectors(first)
    del first
    q = torch.randn((B, d), generator=torch.Generator().manual_seed(QUERY_SEED), dtype=torch.float32).to(dev)
    chunk = torch.empty((step_rows, d), device=dev, dtype=torch.float32)
    st.search(q, k)
    torch.cuda.synchronize()
    cycles = (n1 - n0) // step_rows
    fb0 = int(lib.vs_fallback_count(st.shard.handle))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for c in range(cycles):
        chunk.normal_(generator=g)             # fresh rows every cycle, generated on the device
        st.add_vectors(chunk)
        ids, scores = st.search(q, k)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    # the last appended chunk must be searchable: its first row finds itself
    probe_ids, probe_scores = st.search(chunk[:1].contiguous(), 1)
    ok = int(probe_ids[0, 0].item()) == st.total - step_rows and float(probe_scores[0, 0].item()) > 0.9999
    out = {"workload": "5Mx384 streaming: 1M -> 5M rows by 10k-row appends, each followed by a batch-256 "
                       "top-10 query (config E)",
           "cycles": cycles, "total_ms": ms, "wall_ms": wall * 1e3, "ms_per_cycle": ms / cycles,
           "query_qps_during_ingest": B * cycles / (ms * 1e-3),
           "append_rows_per_s_during_queries": step_rows * cycles / (ms * 1e-3),
           "final_rows": st.total, "last_chunk_searchable": ok,
           "exact_fallbacks": int(lib.vs_fallback_count(st.shard.handle)) - fb0}
    st.close()
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: this engine has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner to stdout; the contract is ONE JSON line there
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    from b200vs import _cabi, build
    from b200vs.sharded import ShardedVectorStore
    build.build_library()
    lib = _cabi.lib()
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured" if "hbm_gbs" in peaks else "fallback"
    tc_burst = float(peaks.get("bf16_tflops", 1590.0))
    tc_sust = float(peaks.get("bf16_tflops_sustained", 1400.0))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def build_store(n, d, mode):
        st = ShardedVectorStore(d, "cosine", device=dev, shadow_bf16=True,
                                max_vectors_per_shard=n // world + 16, search_mode=mode)
        per_block = n // N_BLOCKS
        for b in range(N_BLOCKS):
            owner = b * world // N_BLOCKS
            if owner != rank:
                continue
            g = torch.Generator(device=dev).manual_seed(DB_SEED + b)
            # same stream on any N: block b is always generated whole with its own seed
            if args.dist == "uniform":
                rows = torch.rand((per_block, d), generator=g, device=dev, dtype=torch.float32)
            else:
                rows = torch.randn((per_block, d), generator=g, device=dev, dtype=torch.float32)
            st.shard.append(rows, b * per_block)
            del rows
        st.total = per_block * N_BLOCKS
        torch.cuda.synchronize()
        return st

    def read_profile(kind):
        ms, cnt = C.c_double(), C.c_int64()
        _cabi.check(lib.vs_profile_read(kind, C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    # searches in flight in the timed loops: 2 on one GPU (the certification count of batch i travels to the
    # host while batch i + 1 runs); 3 across GPUs, where a step also depends on every other rank having
    # finished the step two before it, so that one late rank (host jitter) does not stall all of them
    DEPTH = 2 if world == 1 else 3

    def measure(st, n, d, B, k, steps, warmup, with_e2e=True, wl_name=""):
        gq = torch.Generator().manual_seed(QUERY_SEED)
        if args.dist == "uniform":
            q_host = torch.rand((B, d), generator=gq, dtype=torch.float32).pin_memory()
        else:
            q_host = torch.randn((B, d), generator=gq, dtype=torch.float32).pin_memory()
        q_dev = q_host.to(dev)
        fb0 = int(lib.vs_fallback_count(st.shard.handle))
        rt0 = int(lib.vs_retry_count(st.shard.handle))
        def pipelined(nsteps, begin=lambda: st.submit(q_dev, k), finish=st.result):
            """nsteps searches with DEPTH in flight: submit i + DEPTH - 1, then collect i."""
            from collections import deque
            inflight = deque()
            out = None
            for _ in range(nsteps):
                inflight.append(begin())
                if len(inflight) >= DEPTH:
                    out = finish(inflight.popleft())
            while inflight:
                out = finish(inflight.popleft())
            return out

        # warm-up with the timed loop's own pipeline depth: everything a step needs (workspace blocks of
        # the searches in flight, exchange buffers, streams) exists before the timed region
        barrier()
        tw = time.perf_counter()
        wsteps = max(warmup, DEPTH + 1)
        ids, scores = pipelined(wsteps)
        barrier()
        # expected length of the timed region (the same on every rank: the decision below shapes collective loops)
        est_ms = max_over_ranks(1e3 * (time.perf_counter() - tw) / wsteps * steps)
        # ---- device-resident timing (value) ----
        # per-kernel event brackets over the timed region -- unless the store overlaps searches on two
        # streams (N > 1): a bracket would then include queueing behind the other search, so the kernel is
        # timed in a serial pass right after the timed region instead (below) and the region runs unbracketed
        lib.vs_profile(0 if getattr(st, "overlap_streams", False) else 1)
        read_profile(0), read_profile(1)
        barrier()
        launches0 = lib.vs_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = make_sampler(args.clocks, local_rank)
        # A timed region shorter than ~0.7 s is too short for nvidia-smi's 100 ms sampling anyway (the clocks
        # are then sampled under the same load right after it, below) -- and spawning the sampler right before
        # it disturbs the host for milliseconds, which a 10 ms region (8 GPUs, 20 steps) cannot absorb
        sample_inside = est_ms >= 700.0
        if rank == 0 and sample_inside:
            sampler.start()
        # every step submits one batch and collects one: two searches are in flight, so a batch's
        # certification count reaches the host (and its candidates cross NVLink at N > 1) while
        # the next batch already runs -- ShardedVectorStore.submit / result
        e0.record()
        ids, scores = pipelined(steps)
        e1.record()
        barrier()
        clocks = sampler.stop() if (rank == 0 and sample_inside) else {}
        ms = max_over_ranks(e0.elapsed_time(e1))
        launches = lib.vs_launch_count() - launches0
        lib.vs_profile(0)
        resample = ms < 700.0 or not sample_inside
        if resample:
            # the timed region was too short for nvidia-smi's 100 ms sampling: keep the same step
            # running (untimed, every rank -- the all-gather is collective) and sample under load
            extra = int(min(20000, 700.0 / max(ms / steps, 1e-3))) + 1
            sampler = make_sampler(args.clocks, local_rank)
            if rank == 0:
                sampler.start()
                time.sleep(0.05)
            pipelined(extra)
            barrier()
            if rank == 0:
                clocks = sampler.stop()
                clocks["sampled"] = f"under the same load right after the timed region ({extra} extra steps)"
        scan_ms, scan_n = read_profile(0)
        gemm_ms, gemm_n = read_profile(1)
        prof_steps, prof_note = steps, "CUDA events around every launch of the kernel, on its launch stream, over the timed region"
        if getattr(st, "overlap_streams", False):
            # two searches were in flight on two streams: an event bracket around a kernel then also holds
            # the time it queued behind the other search's kernels.  The kernel's own duration is taken
            # from a serial pass (one stream, one search at a time) of the same step right after.
            st.overlap_streams = False
            prof_steps = max(3, min(steps, 10))
            st.search(q_dev, k)
            barrier()
            lib.vs_profile(1)
            read_profile(0), read_profile(1)
            for _ in range(prof_steps):
                st.search(q_dev, k)
            barrier()
            lib.vs_profile(0)
            scan_ms, scan_n = read_profile(0)
            gemm_ms, gemm_n = read_profile(1)
            st.overlap_streams = True
            prof_note = (f"CUDA events around every launch of the kernel in a SERIAL pass of {prof_steps} steps right after the "
                         "timed region (the timed region keeps two searches in flight on two streams, where a per-kernel "
                         "bracket would include queueing behind the other search)")
        res = {"ms_per_step": ms / steps, "qps": B * steps / (ms / 1e3), "launches": int(launches),
               "exact_fallback_queries_per_step": (int(lib.vs_fallback_count(st.shard.handle)) - fb0) /
                                                  max(1, wsteps + steps + (extra if resample else 0)),
               "wide_retry_queries_per_step": (int(lib.vs_retry_count(st.shard.handle)) - rt0) /
                                              max(1, wsteps + steps + (extra if resample else 0)),
               "clocks": clocks, "ids": ids, "scores": scores}
        n_local = n // world
        if gemm_n and gemm_ms >= scan_ms:
            # K3 runs twice per step (sample pass + full pass): the algorithmic work of a step is
            # 2*B*N_local*D flops and one pass over the N_local*D 16-bit rows, set against the summed
            # duration of the step's GEMM launches; the slower of the two rooflines bounds it
            flops = 2.0 * B * n_local * d
            bytes16 = float(n_local) * d * 2
            per_step = gemm_ms / prof_steps
            peak = tc_sust if ms > 1000 else tc_burst
            kname = ("gemm_topk (K3: tcgen05 kind::f16, fp16 operands for cosine / bf16 for dot and "
                     "euclidean, fp32 accumulators in TMEM; pass 1 sample + pass 2 filter)")
            if flops / (peak * 1e12) >= bytes16 / (hbm_peak * 1e9):
                ach = flops / (per_step * 1e-3) / 1e12
                res["roofline"] = {"kernel": kname, "bound": "tensor", "achieved": ach,
                                   "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                                   "peak_source": f"{peak_src} MEASURED_PEAKS.json bf16 cuBLAS "
                                                  f"({'sustained' if ms > 1000 else 'burst'})",
                                   "launches": int(gemm_n), "launches_per_step": gemm_n / prof_steps,
                                   "kernel_ms_per_step": per_step, "kernel_share_of_step": per_step / (ms / steps),
                                   "kernel_timing": prof_note,
                                   "scan_fallback_ms_per_step": scan_ms / prof_steps,
                                   # the runs are power-capped (clocks.reasons): the sustained cuBLAS figure
                                   "frac_of_sustained_peak": ach / tc_sust, "sustained_peak": tc_sust,
                                   "hbm_gbs_of_16bit_rows": bytes16 / (per_step * 1e-3) / 1e9}
            else:
                # small batches: K3 is HBM-bound on the 16-bit rows it streams (half the bytes of the
                # fp32 rows the exact scan reads, with the same exact results by certification)
                ach = bytes16 / (per_step * 1e-3) / 1e9
                res["roofline"] = {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": hbm_peak,
                                   "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                                   "peak_source": f"{peak_src} MEASURED_PEAKS.json hbm_gbs",
                                   "algorithmic_bytes": "N_local * D * 2 (one pass over the 16-bit shadow rows)",
                                   "launches": int(gemm_n), "launches_per_step": gemm_n / prof_steps,
                                   "kernel_ms_per_step": per_step, "kernel_share_of_step": per_step / (ms / steps),
                                   "kernel_timing": prof_note,
                                   "scan_fallback_ms_per_step": scan_ms / prof_steps,
                                   # north_star's roofline for the fp32 path: fp32 database bytes / HBM peak
                                   "fp32_scan_roofline_qps": B / (float(n_local) * d * 4 / (hbm_peak * 1e9)),
                                   "qps_vs_fp32_scan_roofline": (B * steps / (ms / 1e3)) /
                                                                (B / (float(n_local) * d * 4 / (hbm_peak * 1e9)))}
        elif scan_n:
            per = scan_ms / scan_n
            bytes_per_launch = float(n_local) * d * 4   # one pass over the fp32 rows of this shard
            ach = bytes_per_launch / (per * 1e-3) / 1e9
            res["roofline"] = {"kernel": "scan_topk (K2, fp32)", "bound": "hbm", "achieved": ach,
                               "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                               "peak_source": f"{peak_src} MEASURED_PEAKS.json hbm_gbs",
                               "launches": int(scan_n), "avg_launch_ms": per, "kernel_timing": prof_note,
                               "kernel_share_of_step": (scan_ms / prof_steps) / (ms / steps)}
        if world == 1 and "roofline" in res:
            tp = ROOT / "profiles" / "traffic_r02.json"
            if tp.exists():
                kind = "gemm" if res["roofline"]["bound"] == "tensor" else "scan"
                ent = json.loads(tp.read_text()).get(f"{wl_name}|{B}|{kind}")
                if ent:
                    res["roofline"]["traffic"] = ent["bytes"]
                    res["roofline"]["traffic_note"] = ("DRAM bytes per step from the committed ncu capture "
                                                       "(profiles/traffic_r02.json): " + ent["note"])
        # ---- end to end: host buffers in, host buffers out, every step ----
        if with_e2e:
            out_s = torch.empty((B, k), dtype=torch.float32).pin_memory()
            out_i = torch.empty((B, k), dtype=torch.int32).pin_memory()
            flags = _cabi.SEARCH_MODES[args.mode]

            def e2e_run(nsteps):
                if world == 1:
                    # the C-ABI call MLXVectorStore.query()/batch_query() make (host pointers), one per step
                    for _ in range(nsteps):
                        _cabi.check(lib.vs_search_host(st.shard.handle, C.c_void_p(q_host.data_ptr()), B, k,
                                                       flags, None, -1, C.c_void_p(out_s.data_ptr()),
                                                       C.c_void_p(out_i.data_ptr())))
                    return
                # sharded store: every step copies its queries from pinned host memory, searches, and copies
                # its results back to pinned host memory; like the device-resident loop DEPTH steps are in flight
                # (submit i+1, then collect i), the host waits for the last copy at the end
                def begin():
                    return st.submit(q_host.to(dev, non_blocking=True), k)

                def finish(p_):
                    i_, s_ = st.result(p_)
                    out_i.copy_(i_, non_blocking=True)
                    out_s.copy_(s_, non_blocking=True)

                pipelined(nsteps, begin, finish)
                torch.cuda.current_stream().synchronize()

            e2e_run(max(1, warmup))
            barrier()
            t0 = time.perf_counter()
            e2e_run(steps)
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            res["e2e"] = {"value": B * steps / dt, "unit": "queries/s", "ms_per_step": 1e3 * dt / steps,
                          "h2d_bytes_per_step": B * d * 4 * world, "d2h_bytes_per_step": B * k * 8 * world}
        return res

    n, d = WORKLOADS[args.workload]
    B, k = args.batch, args.k
    log(f"building the {args.workload} store ({n // world} rows per rank)")
    st = build_store(n, d, args.mode)
    log(f"headline: batch {B}, {args.steps} steps")
    main = measure(st, n, d, B, k, args.steps, args.warmup, wl_name=args.workload)
    log(f"headline done: {main['qps']:.0f} QPS, {main['ms_per_step']:.3f} ms/step")

    # cheap end-of-run sanity (not parity -- tests/ do that): rank-1 score bound, sortedness
    ids_h = main["ids"].cpu().numpy()
    sc_h = main["scores"].cpu().numpy()
    assert ids_h.shape == (B, k) and (ids_h >= 0).all() and (ids_h < n).all()
    assert (np.diff(sc_h, axis=1) <= 0).all(), "scores not sorted"
    checksum = int(np.bitwise_xor.reduce(ids_h.astype(np.int64).ravel() * 2654435761 % (1 << 31)))

    def entry(name, r):
        return {"workload": name, "qps": r["qps"], "ms_per_step": r["ms_per_step"], "roofline": r.get("roofline"),
                "e2e_qps": r["e2e"]["value"] if "e2e" in r else None,
                "exact_fallback_queries_per_step": r["exact_fallback_queries_per_step"],
                "wide_retry_queries_per_step": r["wide_retry_queries_per_step"]}

    extras = []
    do_extras = args.extras if args.extras >= 0 else (1 if world == 1 else 0)
    # The extras are context next to the headline: they run inside a time budget (seconds since
    # start; the CPU leg behind them needs ~90 s) and whatever does not fit is listed as skipped,
    # so the default run always ends within a few minutes.
    budget = float(os.environ.get("B200VS_BENCH_BUDGET_S", "330"))

    def within_budget(name):
        if time.perf_counter() - _T0 < budget:
            log(f"extra: {name}")
            return True
        log(f"extra skipped (time budget): {name}")
        extras.append({"workload": name, "skipped": f"time budget of {budget:.0f} s reached"})
        return False

    # batch 1 on the same sharded store at every N (north_star's 8-GPU target covers batch 1 too)
    log("batch 1 (AUTO)")
    r = measure(st, n, d, 1, k, max(50, args.steps), args.warmup, wl_name=args.workload)
    extras.append(entry(f"{args.workload} batch 1 (AUTO: 16-bit tensor-core prefilter + certified fp32 rescoring)", r))
    if do_extras:
        short = max(3, min(args.steps, 10))
        if within_budget(f"{args.workload} batch 1 via the fp32 scan (K2, mode scan_fp32)"):
            st.shard.flags = _cabi.SEARCH_MODES["scan_fp32"]
            r = measure(st, n, d, 1, k, max(20, args.steps), args.warmup, wl_name=args.workload)
            extras.append(entry(f"{args.workload} batch 1 via the fp32 scan (K2, mode scan_fp32)", r))
            st.shard.flags = _cabi.SEARCH_MODES[args.mode]
        if within_budget(f"{args.workload} batch 32"):
            r = measure(st, n, d, 32, k, max(20, args.steps), args.warmup)
            extras.append(entry(f"{args.workload} batch 32", r))
        st.close()
        del st
        torch.cuda.empty_cache()
        for name, batches in (("1Mx1536", (1, 1024)), ("1Mx768", (1, 1024))):
            if not within_budget(f"{name} batches {batches}"):
                continue
            n2, d2 = WORKLOADS[name]
            st2 = build_store(n2, d2, args.mode)
            for b2 in batches:
                log(f"  {name} batch {b2}")
                r = measure(st2, n2, d2, b2, k, max(20, args.steps) if b2 == 1 else short, args.warmup, wl_name=name)
                extras.append(entry(f"{name} batch {b2}", r))
            if name == "1Mx1536":
                log(f"  {name} batch 1 scan_fp32")
                st2.shard.flags = _cabi.SEARCH_MODES["scan_fp32"]
                r = measure(st2, n2, d2, 1, k, max(20, args.steps), args.warmup, wl_name=name)
                extras.append(entry(f"{name} batch 1 via the fp32 scan (K2, mode scan_fp32)", r))
            st2.close()
            del st2
            torch.cuda.empty_cache()
        for name, fn in (("config C (1Mx1536 L2 top-100)", config_c_l2_bf16), ("config E (5Mx384 streaming)", config_e_streaming),
                         ("K1 append throughput", k1_append_throughput), ("fp8 variant", fp8_variant)):
            if within_budget(name):
                extras.append(fn(dev, lib, _cabi, ShardedVectorStore, args))
    else:
        st.close()

    # ---- CPU leg (rank 0): the oracle over the FULL database for one query per host thread.  Its ids
    # verify the engine's (--verify, any N); its timing is the cpu_baseline (N = 1 only).
    cpu_baseline = None
    verified = None
    if rank == 0 and (args.verify or (world == 1 and args.cpu_baseline)):
        threads = host_cores()
        limiter = set_host_threads(threads)
        log(f"CPU leg: generating the host copy of the database ({threads} threads)")
        # the host copy holds the SAME rows and queries the engine searched: every block is
        # regenerated on this rank's GPU from its seed (torch's CUDA generator, as build_store does)
        # and copied to the host; the queries come from torch's CPU generator as in measure()
        db_np = np.empty((n, d), np.float32)
        per_block = n // N_BLOCKS
        for b in range(N_BLOCKS):
            g = torch.Generator(device=dev).manual_seed(DB_SEED + b)
            gen = torch.rand if args.dist == "uniform" else torch.randn
            blk = gen((per_block, d), generator=g, device=dev, dtype=torch.float32)
            db_np[b * per_block:(b + 1) * per_block] = blk.cpu().numpy()
            del blk
        gq = torch.Generator().manual_seed(QUERY_SEED)
        q_np = (torch.rand if args.dist == "uniform" else torch.randn)((B, d), generator=gq, dtype=torch.float32).numpy()
        torch.cuda.empty_cache()
        sample_q = cpu_sample_queries(threads, B)
        log(f"CPU leg: oracle over the full database for {sample_q} queries")
        t, ref_ids, ref_sc = reference_step(db_np, q_np, k, sample_q, threads, return_scores=bool(args.verify))
        if args.verify:
            verified = verify_against_oracle(ref_ids, ref_sc, t.pop("score_matrix"), ids_h, sc_h)
        if world == 1 and args.cpu_baseline:
            log("CPU leg: timed sample + optimised-CPU line")
            t, _, _ = reference_step(db_np, q_np, k, sample_q, threads)        # second sample: warm caches
            cpu_baseline = {
                "value": extrapolate_qps(t, sample_q, B), "unit": "queries/s", "cores": threads, "kind": "port",
                "sample": (f"first {sample_q} of {B} queries (one per host thread) vs the full {n}x{d} database on "
                           f"the host, reference op order (per-call DB re-normalisation {t['normalize_s']:.2f}s, fp32 "
                           f"GEMM {t['matmul_s']:.2f}s, full stable argsort {t['argsort_s']:.2f}s); QPS extrapolated "
                           f"to the batch as B/(t_norm + B*(t_gemm+t_sort)/{sample_q}); NumPy port of the reference "
                           f"(mlx not installable)"),
                "optimized_cpu": optimized_cpu_line(db_np, q_np, k, sample_q, threads, B),
            }
        del db_np, limiter

    if rank == 0:
        line = {
            "metric": METRIC_NAME, "value": main["qps"], "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.workload} fp32 cosine top-{k}, batch {B}", "rows": n, "dim": d,
                       "batch": B, "k": k, "sharding": f"rows/{world}", "search_mode": args.mode,
                       "arithmetic": "fp16 tensor-core prefilter (tcgen05, fp32 accumulate) over a 16-bit shadow of the "
                                     "normalised rows + per-query certified fp32 rescoring: results are the exact fp32 "
                                     "ones (uncertified queries are re-run through the fp32 scan; counted below)",
                       "pipeline": f"{2 if world == 1 else 3} batches in flight (submit ahead, then collect the oldest): "
                                   "ShardedVectorStore.submit/result",
                       "l2_policy": "database (5.1 GB fp32 + 2.6 GB 16-bit) is far larger than the 126 MB L2; "
                                    "no flush needed between steps",
                       "data_detail": f"{'U[0,1)' if args.dist == 'uniform' else 'N(0,1)'} rows, 8 seeded blocks "
                                      f"(seed 1234+b), queries seed 4321",
                       "result_checksum": checksum},
            "e2e": main.get("e2e"),
            "gpu_launches": main["launches"],
            "exact_fallback_queries_per_step": main["exact_fallback_queries_per_step"],
            "wide_retry_queries_per_step": main["wide_retry_queries_per_step"],
            "clocks": main["clocks"],
            "roofline": main.get("roofline"),
            "cpu_baseline": cpu_baseline,
            "verified": verified,
            "workloads": extras,
        }
        print(json.dumps(line), flush=True)
        if verified is not None and not verified["ok"]:
            raise SystemExit(f"--verify failed: {verified}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    # watchdog: a bench that stops making progress dumps every thread's Python stack to stderr
    # (B200VS_BENCH_WATCHDOG seconds between dumps; 0 = off)
    wd = float(os.environ.get("B200VS_BENCH_WATCHDOG", "240") or 0)
    if wd > 0:
        import faulthandler
        faulthandler.dump_traceback_later(wd, repeat=True, file=sys.stderr)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
