#!/usr/bin/env python
"""bench.py - hybrid retrieval queries/sec @ top-10 on the hot path (BASELINE.json).

Workload = BASELINE config 3: BM25 over 1M synthetic documents (50k vocabulary, Zipf terms,
~200 tokens) fused via RRF (rrf_k 60) with the dense top-100 of a 1M x 768 two-stage quantised
search (binary Hamming scan -> float32 rescore of 400 candidates), batch of 1024 queries x 8
tokens, fused top-10.  One step = ONE product call, ``HybridSearch.search_batch``
(radiant-rag_b200/hybrid.py): dense -> BM25 -> RRF on device.  With --gpus N the documents are
row-sharded over N ranks (torchrun, NCCL): both halves exchange their per-shard candidates and
every rank ends with the same fused lists (strong scaling: the corpus is fixed).

One JSON line on rank 0:
  value         hybrid queries/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e           the same with HOST (pinned) queries + term ids in and host results out
  roofline      the dominant kernel of the step, timed live with CUDA events on its stream
  cpu_baseline  the NumPy oracle on a bounded sample of the batch, 1 thread, and the GPU's fused
                ids checked against it in the same run (at every N)
  extras        BASELINE config 2 (dense only, batch 256) on the same GPUs; config 4 at N = 2 / 4
                and config 5 at N = 8 when the run has those GPUs

`--impl reference` times the CPU path only (all host cores, bounded sample per step).
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CFG = dict(n=1_000_000, dim=768, q=1024, v=50_000, mean_len=200, q_len=8, dense_k=100, bm25_k=100,
           top_k=10, mult=4.0, rrf_k=60, seed=2, k1=1.5, b=0.75)
CFG2 = dict(n=1_000_000, dim=768, q=256, cand_k=200, top_k=10, seed=1)
METRIC = "hybrid retrieval queries/sec @ top-10"
WORKLOAD = ("config3: BM25 over 1M synthetic docs (50k vocab, Zipf, avg 200 tokens) fused via RRF with the dense "
            "top-100 of a 1M x 768 binary Hamming scan + fp32 rescore, batch 1024 queries x 8 tokens, fused top-10")
WORKLOAD2 = "config2: 1M x 768-dim binary Hamming scan + fp32 rescore of top-200 candidates, batch 256 queries, top-10"


def _peaks() -> dict:
    """Measured peaks: MEASURED_PEAKS.json (driver-written HBM copy / cuBLAS bf16) and
    profiles/r2_peaks.json (tools/peak_probe.py: int8 tensor pipe, POPC, shared memory)."""
    out = {"hbm_gbs": 6650.0, "hbm_src": "fallback (B200_PROFILING.md)", "bf16_tflops": 1590.0}
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        out.update(hbm_gbs=float(d["hbm_gbs"]), hbm_src="measured copy bandwidth (MEASURED_PEAKS.json)",
                   bf16_tflops=float(d.get("bf16_tflops", 1590.0)))
    p2 = ROOT / "profiles" / "r2_peaks.json"
    if p2.exists():
        d = json.loads(p2.read_text())
        out.update(i8_ts_tops=d.get("i8_mma_m128n128_ts_tops"), i8_ss_tops=d.get("i8_mma_m128n256_ss_tops"),
                   popc_tera=d.get("popc32_tera_per_s"), smem_tbs=d.get("smem_load_tb_per_s"),
                   gather_tbs=d.get("gather_3kb_rows_1024x400_tb_per_s"),
                   probe_src="measured by tools/peak_probe.py (profiles/r2_peaks.json)")
    return out


def _traffic(kernel: str, shape: str):
    """dram bytes per launch of `kernel` from a tracked ncu --set full capture keyed by shape
    (profiles/r2_traffic.json), or None when no capture of this shape is tracked."""
    p = ROOT / "profiles" / "r2_traffic.json"
    if not p.exists():
        return None
    return json.loads(p.read_text()).get(kernel, {}).get(shape)


# ---------------------------------------------------------------- CPU side (oracle port)
def _cpu_dense_one(codes, dim, seed, n, qi, cand_k, top_k):
    """One dense query through the CPU restatement: ubinary-quantise, exact Hamming top-k' over all
    rows, float32 rescore of the candidates (rows regenerated from the counter-based generator)."""
    import oracle
    from radiant_rag_b200 import synthetic

    q = synthetic.hash_query_rows_f32(qi, 1, dim, seed, n)[0]
    qc = oracle.quantize_ubinary(q[None, :])
    _d, cand = oracle.hamming_topk(codes, qc, cand_k)
    ids = cand[0][cand[0] >= 0]
    rows = np.concatenate([synthetic.hash_rows_f32(int(r), 1, dim, seed) for r in ids])
    got, _s = oracle.rescore_f32(q, rows, ids, top_k=top_k, min_similarity=0.0)
    return got


def _cpu_hybrid_one(w, qi):
    """dense top-100 + BM25 top-100 -> RRF top-10 for query qi (reference call chain
    radiant/app.py:1178-1249, one query at a time)."""
    import oracle

    d_ids = _cpu_dense_one(w["codes"], CFG["dim"], CFG["seed"], CFG["n"], qi, int(CFG["dense_k"] * CFG["mult"]),
                           CFG["dense_k"])
    b_rows, _ = w["orc"].search(w["qt"][qi].tolist(), CFG["bm25_k"])
    ids, _sc = oracle.rrf_fuse([d_ids.tolist(), b_rows.tolist()], CFG["top_k"], CFG["rrf_k"])
    return ids.tolist()


_W = None


def _cpu_worker_query(qi):
    return _cpu_hybrid_one(_W, qi)


def _gen_codes_chunk(args):
    import oracle
    from radiant_rag_b200 import synthetic

    lo, m, dim, seed = args
    return oracle.quantize_ubinary(synthetic.hash_rows_f32(lo, m, dim, seed))


def _gen_tokens_chunk(args):
    from radiant_rag_b200 import synthetic

    pos, m, seed, v = args
    return synthetic.zipf_tokens(pos, m, seed, synthetic.zipf_cdf_u32(v))


def run_reference_arm(args) -> None:
    """The reference's CPU implementation of the path (NumPy port under oracle/ - the reference
    itself is Python and is not present on the GPU box), all host cores, one query per task."""
    global _W
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    from oracle.bm25 import BM25Oracle
    from radiant_rag_b200 import synthetic

    cores = len(os.sched_getaffinity(0))
    n, dim, seed, v = CFG["n"], CFG["dim"], CFG["seed"], CFG["v"]
    ctx = mp.get_context("fork")
    t0 = time.time()
    per_step = max(cores, 8)
    total_q = min(CFG["q"], per_step * (args.steps + args.warmup))
    qt = synthetic.zipf_queries(CFG["q"], CFG["q_len"], v, seed)
    need = set(int(t) for qi in range(total_q) for t in qt[qi])
    lens = synthetic.doc_lengths(0, n, seed, CFG["mean_len"])
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    total_tok = int(ptr[-1])
    with ctx.Pool(cores) as pool:
        parts = pool.map(_gen_codes_chunk, [(lo, min(25_000, n - lo), dim, seed) for lo in range(0, n, 25_000)])
        tchunk = 4_000_000
        tparts = pool.map(_gen_tokens_chunk, [(p, min(tchunk, total_tok - p), seed, v) for p in range(0, total_tok, tchunk)])
    codes = np.concatenate(parts)
    toks = np.concatenate(tparts)
    del parts, tparts
    orc = BM25Oracle(ptr, toks, v, CFG["k1"], CFG["b"], only_terms=need)
    del toks
    _W = dict(codes=codes, orc=orc, qt=qt)  # inherited by the forked workers (copy on write)
    setup_s = time.time() - t0
    with ctx.Pool(cores) as pool:
        for w in range(args.warmup):
            pool.map(_cpu_worker_query, [(w * per_step + i) % total_q for i in range(per_step)])
        times = []
        for s in range(args.steps):
            t = time.perf_counter()
            pool.map(_cpu_worker_query, [((s + args.warmup) * per_step + i) % total_q for i in range(per_step)])
            times.append(time.perf_counter() - t)
    total = sum(times)
    value = per_step * args.steps / total
    sample = (f"{per_step} queries per step (one per worker process) against the full 1M-doc BM25 index and 1M x 768 "
              f"codes ({args.steps} steps)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8 codes + f32 rescore + f64 BM25/RRF",
        "data": "synthetic", "config": {"workload": WORKLOAD, "queries_per_step": per_step, "setup_s": round(setup_s, 1)},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------- GPU arm
class ClockSampler:
    """nvidia-smi clock / throttle-reason samples every 100 ms, time-stamped; `summary` keeps the
    samples that fall inside the given wall-clock windows (the timed regions)."""

    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def summary(self, windows) -> dict:
        import datetime

        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 8:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    c, m = float(f[1]), float(f[2])
                except ValueError:
                    continue
                if not any(a - 0.05 <= ts <= b + 0.05 for a, b in windows):
                    continue
                sm.append(c)
                mx.append(m)
                for nm, v in zip(names, f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


class Bench:
    """Shared plumbing of the GPU arm: process group, timing loops, L2 flush."""

    def __init__(self, args) -> None:
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.sampler = ClockSampler(self.local_rank) if self.rank == 0 else None
        self.windows = []
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.comm = None
        self.comm_sparse = None
        if self.world > 1:
            # NCCL prints its version banner to stdout on the first communicator; rank 0 must print ONE
            # JSON line, so stdout is pointed at stderr (fd level) until the first collective is done
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                warm = torch.zeros(1, device=self.dev)
                dist.all_reduce(warm)
                torch.cuda.synchronize()
                # candidate exchanges are NCCL calls on the step's own stream, so that a sharded step
                # (kernels + collectives) replays as ONE CUDA graph
                from radiant_rag_b200.nccl import NcclComm
                self.comm = NcclComm(self.dev)
                # the BM25 half of the hybrid step runs on its own stream next to the dense half: its
                # exchanges need their own communicator (one communicator = one order of collectives)
                self.comm_sparse = NcclComm(self.dev)
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)  # > 126 MB L2

    def exchange_us(self, shapes, reps: int = 20) -> float:
        """Event-timed latency of one step's collectives alone (same payload shapes, this rank)."""
        if self.comm is None:
            return 0.0
        torch = self.torch
        bufs = []
        for kind, shape, dtype in shapes:
            t = torch.zeros(shape, dtype=dtype, device=self.dev)
            bufs.append((kind, t, torch.empty((self.world,) + tuple(shape), dtype=dtype, device=self.dev)))

        def run():
            for kind, t, out in bufs:
                if kind == "all_gather":
                    self.comm.all_gather(t, out)
                else:
                    self.comm.all_reduce(t, "max")

        for _ in range(3):
            run()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        self.barrier()
        return e0.elapsed_time(e1) / reps * 1e3

    def barrier(self) -> None:
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def timed(self, fn, steps: int, warmup: int) -> float:
        """Total ms of `steps` calls, each bracketed by CUDA events on the current stream, L2 flushed
        (256 MB fill) before every call outside the event pair; max over ranks."""
        torch = self.torch
        for _ in range(warmup):
            self.flush.fill_(1)
            fn()
        self.barrier()
        evs = []
        for _ in range(steps):
            self.flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        self.barrier()
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([total_ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def time_stage(self, fn, reps: int) -> float:
        """Average ms of one call on THIS rank (no collective inside fn), L2 flushed before each."""
        torch = self.torch
        for _ in range(3):
            self.flush.fill_(1)
            fn()
        torch.cuda.synchronize()
        evs = []
        for _ in range(reps):
            self.flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / reps

    def sustain(self, fn, ms_per_step: float) -> None:
        """Keep the step running ~1.5 s so the 100 ms clock sampler sees the GPU under this load."""
        w1 = time.time()
        n_sustain = max(20, min(20000, int(1500.0 / max(ms_per_step, 1e-3))))  # same count on every rank
        for i in range(n_sustain):
            fn()
            if i % 50 == 49:
                self.torch.cuda.synchronize()
        self.torch.cuda.synchronize()
        self.windows.append((w1, time.time()))

    def sum_over_ranks(self, value: int) -> int:
        if self.world == 1:
            return int(value)
        t = self.torch.tensor([int(value)], dtype=self.torch.int64, device=self.dev)
        self.dist.all_reduce(t)
        return int(t.item())


def _tc_parts(lib, _lib, index, qc, cand_k, flush, reps):
    """Per-kernel ms of the tensor-core stage 1 (rr_tc_timing brackets the four kernels with CUDA
    events on the launching stream), averaged over flushed repetitions."""
    lib.rr_tc_timing(1)
    parts = [0.0, 0.0, 0.0, 0.0]
    buf = (C.c_float * 4)()
    for _ in range(reps):
        flush.fill_(1)
        index.hamming_topk(qc, cand_k, check_overflow=False)
        if lib.rr_tc_last_timing_ms(buf) != 0:
            raise SystemExit("rr_tc_last_timing_ms: " + _lib.last_error())
        parts = [p + float(b) for p, b in zip(parts, buf)]
    lib.rr_tc_timing(0)
    return [p / reps for p in parts]


def run_config2(bx: Bench, steps: int, warmup: int) -> dict:
    """BASELINE config 2 (round 1's headline, kept as an extra): dense two-stage search only,
    batch 256, strong-scaled over the ranks."""
    torch = bx.torch
    from radiant_rag_b200 import _lib
    from radiant_rag_b200.graphed import GraphedSearch, GraphedShardedSearch
    from radiant_rag_b200.index import DenseIndex, synth_query_rows_device, synth_rows_device
    from radiant_rag_b200.sharded import GpuShardOps, ShardedDenseSearch, shard_range

    n, dim, nq, cand_k, top_k, seed = (CFG2[k] for k in ("n", "dim", "q", "cand_k", "top_k", "seed"))
    mult = cand_k / top_k
    lo, hi = shard_range(n, bx.rank, bx.world)
    index = DenseIndex(dim, device=bx.local_rank, store_int8=False, store_f32=True, row_base=lo, capacity=hi - lo)
    for a in range(lo, hi, 125_000):
        index.add(synth_rows_device(a, min(125_000, hi - a), dim, seed, bx.dev))
    queries_dev = synth_query_rows_device(0, nq, dim, seed, n, bx.dev)
    queries_host = queries_dev.cpu().pin_memory()
    search = ShardedDenseSearch(GpuShardOps(index), comm=bx.comm)
    out_host = {"idx": torch.empty((nq, top_k), dtype=torch.int64).pin_memory(),
                "score": torch.empty((nq, top_k), dtype=torch.float32).pin_memory(),
                "count": torch.empty((nq,), dtype=torch.int32).pin_memory()}
    if bx.world > 1:
        graphed = GraphedShardedSearch(search, nq, dim, top_k, rescore_multiplier=mult, prefer_int8=False)
    else:
        graphed = GraphedSearch(lambda q: search.search_quantized(q, top_k, rescore_multiplier=mult, prefer_int8=False,
                                                                  check_overflow=False), nq, dim, bx.dev)
    graphed.load(queries_dev)
    torch.cuda.synchronize()

    def step_e2e():
        idx, score, count = graphed(queries_host)
        out_host["idx"].copy_(idx, non_blocking=True)
        out_host["score"].copy_(score, non_blocking=True)
        out_host["count"].copy_(count, non_blocking=True)

    launches0 = _lib.launch_count
    total_ms = bx.timed(graphed.replay, steps, warmup)
    launches = (_lib.launch_count - launches0) // (steps + warmup)
    e2e_ms = bx.timed(step_e2e, steps, warmup)
    overflow = search.overflow_total()  # collective: summed over the ranks
    out = {"workload": WORKLOAD2, "value": nq / (total_ms / steps * 1e-3), "unit": "queries/s",
           "ms_per_step": total_ms / steps,
           "e2e": {"value": nq / (e2e_ms / steps * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms / steps,
                   "h2d_bytes_per_step": nq * dim * 4, "d2h_bytes_per_step": nq * top_k * 12 + nq * 4},
           "gpu_launches_per_step": launches, "tc_list_overflows": overflow, "rows_per_gpu": hi - lo}
    if bx.rank == 0:
        lib = _lib.load()
        qf, qc = index.quantize_queries(queries_dev)
        reps = max(5, min(steps, 20))
        sample_ms, tau_ms, filter_ms, select_ms = _tc_parts(lib, _lib, index, qc, cand_k, bx.flush, reps)
        _d, cand = index.hamming_topk(qc, cand_k)
        rescore_ms = bx.time_stage(lambda: index.rescore(qf, cand, top_k, 0.0, prefer_int8=False), reps)
        pk = _peaks()
        int8_ops = 2.0 * nq * (hi - lo) * index.words * 32
        tensor_peak = pk.get("i8_ts_tops") or 2.0 * pk["bf16_tflops"]
        out["stage1_kernels_ms"] = {"sample_pass": sample_ms, "tau": tau_ms, "filter_pass": filter_ms,
                                    "list_select": select_ms}
        out["rescore_ms"] = rescore_ms
        out["filter_pass_roofline"] = {
            "bound": "tensor", "achieved": int8_ops / (filter_ms * 1e-3) / 1e12, "peak": tensor_peak, "unit": "TOP/s",
            "frac": int8_ops / (filter_ms * 1e-3) / 1e12 / tensor_peak,
            "peak_source": pk.get("probe_src", "2 x measured bf16 (no probe file)")}
        c_bytes = nq * cand_k * dim * 4 + nq * dim * 4
        out["rescore_roofline"] = {"bound": "hbm", "achieved": c_bytes / (rescore_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"],
                                   "unit": "GB/s", "frac": c_bytes / (rescore_ms * 1e-3) / 1e9 / pk["hbm_gbs"]}
    del graphed, index
    torch.cuda.empty_cache()
    return out


def _gather_obj(bx: Bench, obj):
    """Python object from every rank -> list on every rank (host-side; parity checks only)."""
    if bx.world == 1:
        return [obj]
    out = [None] * bx.world
    bx.dist.all_gather_object(out, obj)
    return out


def run_config4(bx: Bench, steps: int, n_total: int = 10_000_000) -> dict:
    """BASELINE config 4: 10M x 1024 int8 exact search on the tensor cores, batch 4096, top-10,
    row-sharded over the ranks (local exact top-k -> all_gather -> merge).  Parity: every rank
    scores a sample of the queries against ITS rows with the oracle (exact int32 through float32
    BLAS), rank 0 merges the per-shard lists on the CPU and compares with the GPUs' merged result."""
    torch = bx.torch
    import oracle
    from radiant_rag_b200 import synthetic
    from radiant_rag_b200.index import DenseIndex, synth_query_rows_device, synth_rows_device
    from radiant_rag_b200.sharded import GpuShardOps, ShardedInt8Search, shard_range

    nq, dim, seed, top_k = 4096, 1024, 3, 10
    lo, hi = shard_range(n_total, bx.rank, bx.world)
    bound = 131070.0 * 2.0 ** -synthetic.value_shift(dim)
    ranges = np.stack([np.full(dim, -bound, np.float32), np.full(dim, bound, np.float32)])
    index = DenseIndex(dim, device=bx.local_rank, store_int8=True, store_f32=False, int8_ranges=ranges, row_base=lo,
                       capacity=hi - lo)
    for a in range(lo, hi, 250_000):
        index.add(synth_rows_device(a, min(250_000, hi - a), dim, seed, bx.dev))
    q8 = index.quantize_int8_queries(synth_query_rows_device(0, nq, dim, seed, n_total, bx.dev))
    search = ShardedInt8Search(GpuShardOps(index), comm=bx.comm)
    out = {}

    def step():
        out["r"] = search.search(q8, top_k)

    total_ms = bx.timed(step, steps, 2)
    idx, score = out["r"]
    ms = total_ms / steps
    sample = np.arange(0, nq, 512)  # 8 queries
    sel = torch.from_numpy(sample).to(bx.dev)
    t0 = time.perf_counter()
    w_r, w_s = oracle.int8_exact_topk_blas(q8[sel].cpu().numpy(), index.int8[: hi - lo].cpu().numpy(), top_k)
    cpu_s = time.perf_counter() - t0
    parts = _gather_obj(bx, (w_r + lo, w_s))
    res = {"workload": f"config4: {n_total} x {dim} int8 exact search (tcgen05 kind::i8), batch {nq}, top-{top_k}, "
                       f"row-sharded x{bx.world} (NCCL all_gather of per-shard top-k + merge)",
           "ms_per_batch": ms, "value": nq / (ms * 1e-3), "unit": "queries/s", "rows_per_gpu": hi - lo,
           "int8_TOPS_per_gpu": 2.0 * (hi - lo) * dim * nq / (ms * 1e-3) / 1e12}
    if bx.rank == 0:
        agree = 0
        gi, gs = idx[sel].cpu().numpy(), score[sel].cpu().numpy()
        for j in range(sample.size):
            r = np.concatenate([p[0][j] for p in parts])
            sc = np.concatenate([p[1][j] for p in parts]).astype(np.int64)
            order = np.lexsort((r, -sc))[:top_k]
            agree += int(np.array_equal(gi[j], r[order]) and np.array_equal(gs[j].astype(np.int64), sc[order]))
        pk = _peaks()
        peak = pk.get("i8_ss_tops") or 2.0 * pk["bf16_tflops"]
        res["frac_of_int8_peak_per_gpu"] = res["int8_TOPS_per_gpu"] / peak
        res["cpu_baseline"] = {"value": sample.size / cpu_s, "unit": "queries/s over one shard", "cores": "BLAS threads",
                               "kind": "port", "sample": f"{sample.size} queries x {hi - lo} rows per rank (float32 BLAS, exact)",
                               "gpu_matches_cpu_on_sample": f"{agree}/{sample.size}"}
    del index, search, q8
    torch.cuda.empty_cache()
    return res


def run_config5(bx: Bench, n_total: int = 100_000_000) -> dict:
    """BASELINE config 5: 100M x 1024 binary codes + int8 rescore (k' = 40) and BM25 over the same
    100M documents (50k vocabulary, ~200 tokens), fused by RRF, top-10, batch 8192, row-sharded
    over 8 GPUs: the hybrid product call with NCCL exchanges of the per-shard candidates.
    Parity on a sample of the queries: the dense half against a distributed oracle at full size
    (per-shard Hamming lists merged on the CPU, owners rescore), BM25 by exact float64 scores of
    every returned document recomputed from its regenerated tokens plus a selection check over
    the deterministic first 1M documents, RRF against oracle.rrf_fuse."""
    torch, dist = bx.torch, bx.dist
    import oracle
    from oracle.bm25 import BM25Oracle
    from radiant_rag_b200 import synthetic
    from radiant_rag_b200.bm25_index import Bm25DeviceIndex, synth_zipf_corpus_device
    from radiant_rag_b200.hybrid import HybridSearch
    from radiant_rag_b200.index import DenseIndex, synth_query_rows_device, synth_rows_device
    from radiant_rag_b200.sharded import shard_range

    dim, nq, v, mean_len, q_len, seed = 1024, 8192, 50_000, 200, 8, 4
    top_k, mult = 10, 4.0
    dev = bx.dev
    t0 = time.time()
    lo, hi = shard_range(n_total, bx.rank, bx.world)
    n_local = hi - lo
    bound = 131070.0 * 2.0 ** -synthetic.value_shift(dim)
    ranges = np.stack([np.full(dim, -bound, np.float32), np.full(dim, bound, np.float32)])
    index = DenseIndex(dim, device=bx.local_rank, store_int8=True, store_f32=False, int8_ranges=ranges, row_base=lo,
                       capacity=n_local)
    for a in range(lo, hi, 500_000):
        index.add(synth_rows_device(a, min(500_000, hi - a), dim, seed, dev))
    ptr, toks = synth_zipf_corpus_device(n_local, v, seed, mean_len, device=bx.local_rank, row_start=lo)
    bm = Bm25DeviceIndex.build(ptr, toks, v, None, None, 1.5, 0.75, device=bx.local_rank, row_base=lo, sharded=True)
    del ptr, toks
    torch.cuda.empty_cache()
    build_s = time.time() - t0
    queries = synth_query_rows_device(0, nq, dim, seed, n_total, dev)
    qt_np = synthetic.zipf_queries(nq, q_len, v, seed)
    qt = torch.from_numpy(qt_np).to(dev)
    hybrid = HybridSearch(index, bm, rescore_multiplier=mult, prefer_int8=True, comm=bx.comm)
    kw = dict(top_k=top_k, dense_top_k=top_k, bm25_top_k=top_k, rrf_k=60)
    out = {}

    redone = [0, 0]

    def step():
        # CHECKED step: flagged BM25 queries / overflowed candidate lists are redone on the exact
        # paths inside the timed region (one device sync per half), so every result is proven exact
        out["r"] = hybrid.search_batch(queries, qt, check=True, **kw)
        redone[0] += index.last_tc_redone
        redone[1] += bm.last_flagged

    index.last_tc_redone = 0
    total_ms = bx.timed(step, 3, 2)
    ms = total_ms / 3
    events = bx.sum_over_ranks(redone[0] + redone[1])
    res_g = out["r"]
    # per-half times on this rank (each half contains its own collectives)
    d_ms = bx.timed(lambda: hybrid.dense.search_quantized(queries, top_k, rescore_multiplier=mult, check_overflow=False), 2, 1) / 2
    b_ms = bx.timed(lambda: hybrid.sparse.search_batch(qt, top_k, check=False), 2, 1) / 2
    res = {"workload": f"config5: {n_total} x {dim} binary codes + int8 rescore (k'=40) + BM25 over {n_total} docs "
                       f"(50k vocab, ~200 tokens) fused by RRF, top-{top_k}, batch {nq}, row-sharded x{bx.world}",
           "ms_per_batch": ms, "value": nq / (ms * 1e-3), "unit": "hybrid queries/s", "rows_per_gpu": n_local,
           "postings_per_gpu": bm.n_postings, "index_build_s": round(build_s, 1),
           "ms": {"dense_top10": d_ms, "bm25_top10": b_ms}, "queries_redone_on_exact_paths_in_5_steps_all_ranks": events,
           "dense_queries_per_s": nq / (d_ms * 1e-3), "bm25_queries_per_s": nq / (b_ms * 1e-3)}

    # ---- parity on a sample
    sample = [int(x) for x in np.linspace(1, nq - 1, 8).astype(int)]
    sel = torch.tensor(sample, device=dev)
    qh = queries[sel].cpu().numpy()
    # dense: every rank's local Hamming top-40 from the oracle over ITS codes (full shard size)
    codes = index.codes[:n_local].cpu().numpy()[:, : dim // 8]
    loc_lists = []
    for j in range(len(sample)):
        dd, cc = oracle.hamming_topk(codes, oracle.quantize_ubinary(qh[j:j + 1]), 40)
        loc_lists.append((dd[0], cc[0] + lo))
    del codes
    all_lists = _gather_obj(bx, loc_lists)
    own_scores = []
    for j in range(len(sample)):
        d_all = np.concatenate([p[j][0] for p in all_lists]).astype(np.int64)
        r_all = np.concatenate([p[j][1] for p in all_lists])
        order = np.lexsort((r_all, d_all))[:40]
        cand = r_all[order]
        mine = (cand >= lo) & (cand < hi)
        rows = index.int8[torch.from_numpy(cand[mine] - lo).to(dev)].cpu().numpy()
        sc = (rows.astype(np.float64) @ qh[j].astype(np.float64)).astype(np.float32)
        own_scores.append((cand, mine, sc))
    all_scores = _gather_obj(bx, own_scores)
    dense_ok = 0
    if bx.rank == 0:
        for j, qi in enumerate(sample):
            cand = all_scores[0][j][0]
            s = np.full(cand.size, -np.inf, dtype=np.float32)
            for p in all_scores:
                s[p[j][1]] = p[j][2]
            order = np.argsort(-s.astype(np.float64), kind="stable")[:top_k]
            order = order[s[order] >= 0.0]
            m = int(res_g.dense_count[qi])
            dense_ok += int(res_g.dense_idx[qi, :m].cpu().tolist() == cand[order].tolist())
    # BM25: exact float64 scores of EVERY returned document recomputed from its regenerated tokens
    # (global idf / avgdl are the index's inputs, R7), and no document of the deterministic first
    # 1M documents beats the k-th returned one
    from radiant_rag_b200 import _lib
    bm25_ok = rrf_ok = 0
    need = sorted(set(int(t) for qi in sample for t in qt_np[qi]))
    need_t = torch.tensor(need, device=dev)
    df_need = (bm.tile_term_ptr[:, 1:] - bm.tile_term_ptr[:, :-1])[:, need_t].sum(dim=0)
    lens_local = torch.empty(n_local, dtype=torch.int32, device=dev)
    _lib.call("rr_synth_doc_lengths", lens_local.data_ptr(), lo, n_local, seed, mean_len,
              torch.cuda.current_stream().cuda_stream)
    ptr_local = torch.zeros(n_local + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens_local.to(torch.int64), 0, out=ptr_local[1:])
    totals = _gather_obj(bx, int(ptr_local[-1].item()))
    pos0 = sum(totals[: bx.rank])  # global token position of this shard's first token
    if bx.world > 1:
        dist.all_reduce(df_need)
    rows_all = res_g.bm25_idx[sel].flatten()
    mine = rows_all[(rows_all >= lo) & (rows_all < hi)] - lo
    where = {int(r) + lo: (pos0 + int(p), int(ln)) for r, p, ln in
             zip(mine.cpu().tolist(), ptr_local[mine].cpu().tolist(), lens_local[mine].cpu().tolist())}
    where_all = _gather_obj(bx, where)
    if bx.rank == 0:
        loc = {}
        for w in where_all:
            loc.update(w)
        n_glob, tok_glob = n_total, sum(totals)
        avgdl = tok_glob / n_glob
        idf = np.zeros(v, dtype=np.float64)
        for t, d in zip(need, df_need.cpu().tolist()):
            if d:
                idf[t] = np.log((n_glob - d + 0.5) / (d + 0.5) + 1.0)
        cdf = synthetic.zipf_cdf_u32(v)
        sub = min(1_000_000, n_total)
        sub_ptr, sub_toks = synth_zipf_corpus_device(sub, v, seed, mean_len, device=bx.local_rank)
        orc_sub = BM25Oracle(sub_ptr.cpu().numpy(), sub_toks.cpu().numpy(), v, 1.5, 0.75, idf=idf, avgdl=avgdl,
                             only_terms=need)
        orc_sub.known = idf > 0
        del sub_ptr, sub_toks
        for j, qi in enumerate(sample):
            m = int(res_g.bm25_count[qi])
            rows = res_g.bm25_idx[qi, :m].cpu().numpy()
            got = res_g.bm25_score[qi, :m].cpu().numpy()
            ok = m == top_k
            # (a) a mini-corpus of just the returned documents, tokens regenerated on the CPU
            mini_toks = [synthetic.zipf_tokens(loc[int(r)][0], loc[int(r)][1], seed, cdf) for r in rows]
            mini_ptr = np.concatenate([[0], np.cumsum([t.size for t in mini_toks])])
            orc_mini = BM25Oracle(mini_ptr, np.concatenate(mini_toks), v, 1.5, 0.75, idf=idf, avgdl=avgdl)
            orc_mini.known = idf > 0
            ok = ok and np.array_equal(orc_mini.scores(qt_np[qi].tolist()), got)
            ok = ok and all((got[i] > got[i + 1]) or (got[i] == got[i + 1] and rows[i] < rows[i + 1]) for i in range(m - 1))
            # (b) selection: nothing among the first 1M documents outranks the returned k-th
            s_sub = orc_sub.scores(qt_np[qi].tolist())
            kth_s, kth_r = got[-1], rows[-1]
            better = np.nonzero((s_sub > kth_s) | ((s_sub == kth_s) & (np.arange(sub) < kth_r)))[0]
            ok = ok and set(better.tolist()) <= set(rows.tolist())
            bm25_ok += int(ok)
            dm = int(res_g.dense_count[qi])
            ids, sc = oracle.rrf_fuse([res_g.dense_idx[qi, :dm].cpu().tolist(), rows.tolist()], top_k, 60)
            fm = int(res_g.count[qi])
            rrf_ok += int(res_g.idx[qi, :fm].cpu().tolist() == ids.tolist() and res_g.score[qi, :fm].cpu().tolist() == sc.tolist())
        res["parity_on_sample"] = {
            "queries": len(sample),
            "dense_matches_distributed_oracle": f"{dense_ok}/{len(sample)}",
            "bm25_returned_scores_exact_and_selection_holds_on_first_1M_docs": f"{bm25_ok}/{len(sample)}",
            "rrf_matches_oracle": f"{rrf_ok}/{len(sample)}"}
    del hybrid, bm, index
    torch.cuda.empty_cache()
    return res


def run_gpu_arm(args) -> None:
    bx = Bench(args)
    torch, dist = bx.torch, bx.dist
    from radiant_rag_b200 import _lib, synthetic
    from radiant_rag_b200.bm25_index import Bm25DeviceIndex, synth_zipf_corpus_device
    from radiant_rag_b200.hybrid import GraphedHybridSearch, HybridSearch
    from radiant_rag_b200.index import DenseIndex, synth_query_rows_device, synth_rows_device
    from radiant_rag_b200.sharded import shard_range

    world, rank, dev = bx.world, bx.rank, bx.dev
    n, dim, nq, v, seed = CFG["n"], CFG["dim"], CFG["q"], CFG["v"], CFG["seed"]
    top_k, dense_k, bm25_k, mult, rrf_k, q_len = (CFG[k] for k in ("top_k", "dense_k", "bm25_k", "mult", "rrf_k", "q_len"))
    cand_k = int(dense_k * mult)

    # ---- resident indexes of this rank's shard, generated on device
    t_build = time.time()
    lo, hi = shard_range(n, rank, world)
    index = DenseIndex(dim, device=bx.local_rank, store_int8=False, store_f32=True, row_base=lo, capacity=hi - lo)
    for a in range(lo, hi, 125_000):
        index.add(synth_rows_device(a, min(125_000, hi - a), dim, seed, dev))
    ptr, toks = synth_zipf_corpus_device(hi - lo, v, seed, CFG["mean_len"], device=bx.local_rank, row_start=lo)
    bm = Bm25DeviceIndex.build(ptr, toks, v, None, None, CFG["k1"], CFG["b"], device=bx.local_rank, row_base=lo,
                               sharded=True)  # global df / avgdl over the ranks, built by the product from tokens
    bm.fast_min_docs = 0  # shards of a large corpus keep the batched path
    del ptr, toks
    queries_dev = synth_query_rows_device(0, nq, dim, seed, n, dev)
    qt_np = synthetic.zipf_queries(nq, q_len, v, seed)
    qt_dev = torch.from_numpy(qt_np).to(dev)
    queries_host = queries_dev.cpu().pin_memory()
    qt_host = torch.from_numpy(qt_np).pin_memory()
    torch.cuda.synchronize()
    build_s = time.time() - t_build
    hybrid = HybridSearch(index, bm, rescore_multiplier=mult, prefer_int8=False, comm=bx.comm,
                          overlap=not args.no_overlap, comm_sparse=bx.comm_sparse)
    kw = dict(top_k=top_k, dense_top_k=dense_k, bm25_top_k=bm25_k, rrf_k=rrf_k)

    out_host = {"idx": torch.empty((nq, top_k), dtype=torch.int64).pin_memory(),
                "score": torch.empty((nq, top_k), dtype=torch.float64).pin_memory(),
                "count": torch.empty((nq,), dtype=torch.int32).pin_memory()}

    graphed = None
    if not args.no_graph:  # at N > 1 the NCCL exchanges are part of the graph (nccl.NcclComm)
        graphed = GraphedHybridSearch(hybrid, nq, dim, q_len, **kw)
        graphed.load(queries_dev, qt_dev)
        torch.cuda.synchronize()

    def step_device():
        if graphed is not None:
            return graphed.replay()  # static inputs already hold the resident batch
        return hybrid.search_batch(queries_dev, qt_dev, check=False, **kw)

    def step_e2e():
        # pipelined serving loop: H2D of this step's queries + term ids (pinned), the step, D2H of
        # the fused lists.  Exactness counters are accumulated on device and asserted below.
        if graphed is not None:
            res = graphed(queries_host, qt_host)
        else:
            res = hybrid.search_batch(queries_host.to(dev, non_blocking=True), qt_host.to(dev, non_blocking=True),
                                      check=False, **kw)
        out_host["idx"].copy_(res.idx, non_blocking=True)
        out_host["score"].copy_(res.score, non_blocking=True)
        out_host["count"].copy_(res.count, non_blocking=True)

    hybrid.reset_unchecked_events()
    launches0 = _lib.launch_count
    w0 = time.time()
    total_ms = bx.timed(step_device, args.steps, args.warmup)
    launches_per_step = (_lib.launch_count - launches0) // (args.steps + args.warmup)
    e2e_ms = bx.timed(step_e2e, args.steps, args.warmup)
    bx.windows.append((w0, time.time()))
    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": total_ms / args.steps,
                              "e2e_ms_per_step": e2e_ms / args.steps, "gpu_launches_per_step": launches_per_step}), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    bx.sustain(step_device, total_ms / args.steps)
    # every rank learns about inexact events on ANY rank before anyone decides to stop
    events = bx.sum_over_ranks(hybrid.unchecked_events())
    if events != 0:
        # flagged BM25 queries / overflowed candidate lists in an unchecked step: the timed results
        # were not all proven exact - report it instead of a number
        if rank == 0:
            print(json.dumps({"metric": METRIC, "error": f"{events} unchecked exactness events in the timed region"}),
                  flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        raise SystemExit(2)
    exchange_us = bx.exchange_us([("all_gather", (nq, cand_k), torch.int64), ("all_reduce", (nq, cand_k), torch.float32),
                                  ("all_gather", (nq, bm25_k), torch.float64), ("all_gather", (nq, bm25_k), torch.int64)])
    res_dev = step_device()
    fused_idx = res_dev.idx.clone()
    dense_idx = res_dev.dense_idx.clone()
    torch.cuda.synchronize()

    # ---- rank 0 needs the whole corpus for the CPU check at N > 1: regenerate it on its own GPU
    cpu = None
    lib = _lib.load()
    if rank == 0:
        sample_q = 48
        sample = [int(x) for x in np.linspace(0, nq - 1, sample_q).astype(int)]
        if world == 1:
            codes_host = index.codes[: hi - lo].cpu().numpy()[:, : dim // 8]
        else:
            parts = []
            for a in range(0, n, 125_000):
                rows = synth_rows_device(a, min(125_000, n - a), dim, seed, dev)
                c = torch.empty((rows.shape[0], index.words * 4), dtype=torch.uint8, device=dev)
                _lib.call("rr_quantize_ubinary", rows.data_ptr(), rows.shape[0], dim, c.data_ptr(), index.words * 4,
                          torch.cuda.current_stream().cuda_stream)
                parts.append(c.cpu().numpy()[:, : dim // 8])
            codes_host = np.concatenate(parts)
        ptr_all, toks_all = synth_zipf_corpus_device(n, v, seed, CFG["mean_len"], device=bx.local_rank)
        ptr_h, toks_h = ptr_all.cpu().numpy(), toks_all.cpu().numpy()
        del ptr_all, toks_all
        import oracle
        from oracle.bm25 import BM25Oracle

        need = set(int(t) for qi in sample for t in qt_np[qi])
        orc = BM25Oracle(ptr_h, toks_h, v, CFG["k1"], CFG["b"], only_terms=need)
        del toks_h
        w = dict(codes=codes_host, orc=orc, qt=qt_np)
        t0 = time.perf_counter()
        cpu_ids = [_cpu_hybrid_one(w, qi) for qi in sample]
        cpu_s = time.perf_counter() - t0
        gi = fused_idx.cpu().numpy()
        agree = sum(1 for j, qi in enumerate(sample) if gi[qi][: len(cpu_ids[j])].tolist() == cpu_ids[j]
                    and (len(cpu_ids[j]) == top_k or gi[qi][len(cpu_ids[j])] == -1))
        cpu = {"value": sample_q / cpu_s, "unit": "queries/s", "cores": 1, "kind": "port",
               "sample": f"{sample_q} of the {nq} queries against the full 1M-doc corpus (dense two-stage + BM25 + RRF), "
                         "NumPy oracle, 1 thread; index build not timed",
               "gpu_matches_cpu_on_sample": f"{agree}/{sample_q}"}

    # ---- stage breakdown + roofline (rank 0's shard)
    line = None
    if rank == 0:
        reps = max(5, min(args.steps, 20))
        n_local = hi - lo
        qf, qc = index.quantize_queries(queries_dev)
        dense_ms = bx.time_stage(lambda: index.search_quantized(queries_dev, dense_k, rescore_multiplier=mult,
                                                                prefer_int8=False, check_overflow=False), reps)
        sample_ms, tau_ms, filter_ms, select_ms = _tc_parts(lib, _lib, index, qc, cand_k, bx.flush, reps)
        _d, cand = index.hamming_topk(qc, cand_k)
        rescore_ms = bx.time_stage(lambda: index.rescore(qf, cand, dense_k, 0.0, prefer_int8=False), reps)
        bm25_ms = bx.time_stage(lambda: bm.search_batch(qt_dev, bm25_k, check=False), reps)
        lib.rr_bm25_timing(1)
        bparts = [0.0] * 4
        buf = (C.c_float * 4)()
        for _ in range(reps):
            bx.flush.fill_(1)
            bm.search_batch(qt_dev, bm25_k, check=False)
            if lib.rr_bm25_last_timing_ms(buf) != 0:
                raise SystemExit("rr_bm25_last_timing_ms: " + _lib.last_error())
            bparts = [p + float(b) for p, b in zip(bparts, buf)]
        lib.rr_bm25_timing(0)
        b_sample, b_tau, b_filter, b_refine = (p / reps for p in bparts)
        from radiant_rag_b200.agents import rrf_fuse_runs_device
        b_idx = res_dev.bm25_idx
        rrf_ms = bx.time_stage(lambda: rrf_fuse_runs_device([dense_idx, b_idx], top_k, rrf_k), reps)
        clocks = bx.sampler.summary(bx.windows) if bx.sampler else {}
        pk = _peaks()
        ms_per_step = total_ms / args.steps
        # dominant kernel = the largest single launch of the step
        int8_ops = 2.0 * nq * n_local * index.words * 32
        tensor_peak = pk.get("i8_ts_tops") or 2.0 * pk["bf16_tflops"]
        qt_flat = qt_np.ravel()
        # algorithmic bytes of the BM25 filter pass: ONE pass over the shard's index per batch
        # (packed postings 8 B + dense float16 head columns 2 B/doc/term); without sharing every query
        # would pull sum_t df(t) * 12 B on its own (SURVEY.md 8d's per-query figure)
        index_bytes = bm.n_postings * 8 + (bm.head_imp.numel() * 2 if bm.n_head else 0)
        qt_valid = torch.from_numpy(qt_flat[qt_flat >= 0].astype(np.int64)).to(dev)
        head_tok = float((bm.head_slot[qt_valid] >= 0).sum().item()) / nq if bm.n_head else 0.0  # head tokens per query
        ttp = bm.tile_term_ptr
        df_local = (ttp[:, 1:] - ttp[:, :-1]).sum(dim=0)
        unshared_bytes = float(df_local[torch.from_numpy(qt_flat[qt_flat >= 0].astype(np.int64)).to(dev)].sum().item()) * 12.0
        tc_roof = {
            "kernel": "tc_i8_search_kernel<EPI_FILTER> (tcgen05 kind::i8 filter pass of the dense stage 1)",
            "bound": "tensor", "achieved": int8_ops / (filter_ms * 1e-3) / 1e12, "peak": tensor_peak, "unit": "TOP/s",
            "frac": int8_ops / (filter_ms * 1e-3) / 1e12 / tensor_peak, "launch_ms": filter_ms,
            "share_of_step": filter_ms / ms_per_step,
            "traffic": _traffic("tc_i8_search_kernel_filter", f"{n_local}x{dim}x{nq}"),
            "peak_source": pk.get("probe_src", "2 x measured dense bf16 cuBLAS burst (no probe file)") +
                           ": back-to-back tcgen05.mma kind::i8 M128 N128, A in tensor memory",
            "algorithmic_ops_per_launch": int8_ops, "algorithmic_bytes_per_launch": n_local * index.words * 4}
        bm_roof = {
            "kernel": "bm25_fast_kernel<FILTER> (batched float32 filter pass: dense head columns in shared memory + "
                      "tail postings scattered per query)",
            "bound": "hbm", "achieved": index_bytes / (b_filter * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
            "limiter": "the latency of each warp's dependent chain at 8 warps per scheduler (ncu: IPC 2.7, issue active 68 %, LSU wavefronts 52 %: profiles/r2_ncu_full_summary.md), not HBM: the batch shares one pass over the index",
            "frac": index_bytes / (b_filter * 1e-3) / 1e9 / pk["hbm_gbs"], "launch_ms": b_filter,
            "share_of_step": b_filter / ms_per_step,
            "traffic": _traffic("bm25_fast_kernel_filter", f"{n_local}x{v}x{nq}"),
            "peak_source": pk["hbm_src"],
            "algorithmic_bytes_per_launch": index_bytes,
            "algorithmic_bytes_without_batch_sharing": unshared_bytes,
            "achieved_without_batch_sharing_GBs": unshared_bytes / (b_filter * 1e-3) / 1e9,
            "note": "one pass over the index serves the whole batch (SURVEY.md 8d asks for both accountings); the "
                    "kernel's own limiter is the latency of each warp's dependent chain at 8 warps per scheduler, then shared memory: per (query, document) "
                    "4 B of tail accumulator zeroed + 4 B read + 2 B per head token",
            "smem_view": None if not pk.get("smem_tbs") else {
                "bytes_per_launch": float(nq) * n_local * (8.0 + 2.0 * head_tok),
                "achieved_TBs": float(nq) * n_local * (8.0 + 2.0 * head_tok) / (b_filter * 1e-3) / 1e12,
                "peak_TBs": pk["smem_tbs"], "frac": float(nq) * n_local * (8.0 + 2.0 * head_tok) / (b_filter * 1e-3) / 1e12 / pk["smem_tbs"]}}
        c_bytes = float(nq) * cand_k * dim * 4 + nq * dim * 4
        rs_roof = {"kernel": "rescore_ring_kernel (candidate rows bulk-copied into a shared-memory ring)", "bound": "hbm",
                   "traffic": _traffic("rescore_ring_kernel_f32", f"{nq}x{cand_k}x{dim}"),
                   "achieved": c_bytes / (rescore_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                   "frac": c_bytes / (rescore_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "launch_ms": rescore_ms,
                   "share_of_step": rescore_ms / ms_per_step, "algorithmic_bytes_per_launch": c_bytes}
        if pk.get("gather_tbs"):  # what the memory system delivers for this access pattern with the arithmetic left out
            rs_roof["gather_view"] = {"peak_TBs": pk["gather_tbs"], "what": "rr_probe_gather: the same bulk-copy ring over "
                                      "random 3 KB rows, rows released unread (profiles/r2_peaks.json)",
                                      "frac": c_bytes / (rescore_ms * 1e-3) / 1e12 / pk["gather_tbs"]}
        dominant = max((tc_roof, bm_roof, rs_roof), key=lambda r: r["launch_ms"])
        roofline = dict(dominant)
        roofline["timing"] = ("CUDA events on the launching stream around the kernel (rr_tc_timing / rr_bm25_timing), "
                              "L2 flushed before every call")
        roofline["other_kernels"] = [r for r in (tc_roof, bm_roof, rs_roof) if r is not dominant]
        line = {
            "metric": METRIC, "value": nq / (ms_per_step * 1e-3), "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "u8 sign codes (0/255 u8 x +-1 s8 on tcgen05, int32) + f32 rescore + f32-filtered / f64-exact BM25 + f64 RRF",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "docs": n, "dim": dim, "batch_queries": nq, "query_tokens": q_len,
                       "dense_top_k": dense_k, "dense_candidates": cand_k, "bm25_top_k": bm25_k, "top_k": top_k,
                       "rrf_k": rrf_k, "docs_per_gpu": n_local, "parallelism": f"row-sharded x{world}",
                       "postings_per_gpu": bm.n_postings, "index_build_s": round(build_s, 1),
                       "l2": "flushed between timed iterations (256 MB fill)",
                       "issue": ("CUDA graph replay of the step" + (" (NCCL exchanges inside the graph)" if world > 1 else ""))
                                if graphed is not None else "eager launches + NCCL exchanges"},
            "e2e": {"value": nq / (e2e_ms / args.steps * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": nq * dim * 4 + nq * q_len * 4, "d2h_bytes_per_step": nq * top_k * 16 + nq * 4},
            "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
            "exchange_us": exchange_us,
            "stages_ms": {"dense_top100": dense_ms, "bm25_top100": bm25_ms, "rrf_top10": rrf_ms,
                          "dense_stage1": {"sample_pass": sample_ms, "tau": tau_ms, "filter_pass": filter_ms,
                                           "list_select": select_ms},
                          "dense_rescore": rescore_ms,
                          "bm25": {"sample_pass": b_sample, "tau": b_tau, "filter_pass": b_filter, "refine": b_refine}},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": dict({k: clocks.get(k) for k in ("sm_mhz", "sm_max_mhz", "reasons", "samples")},
                           window="timed regions + 1.5 s of the same step back to back"),
        }
    del graphed, hybrid, bm, index
    torch.cuda.empty_cache()

    # ---- extras: the other BASELINE configs on the same GPUs (collective: every rank takes part)
    extras = {}
    if not args.no_extras:
        ex_steps = max(3, min(args.steps, 10))
        extras["config2"] = run_config2(bx, ex_steps, 3)
        if world in (2, 4):
            extras["config4"] = run_config4(bx, ex_steps)
        if world == 8:
            extras["config5"] = run_config5(bx)
    if rank == 0:
        line["extras"] = extras
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=0, help="debug only: override the corpus size")
    ap.add_argument("--no-graph", action="store_true", help="issue every step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra BASELINE configs")
    ap.add_argument("--no-overlap", action="store_true",
                    help="run the BM25 half after the dense half on one stream instead of next to it on a second one")
    ap.add_argument("--profile", action="store_true",
                    help="profiling runs (ncu): build, warm up and run the timed steps only, print a short line")
    ap.add_argument("--debug-extra", default="", help="debug only: run just this extra (config2 | config4 | config5)")
    ap.add_argument("--debug-rows", type=int, default=0, help="debug only: total rows of the --debug-extra config")
    args = ap.parse_args()
    if args.rows:
        CFG["n"] = args.rows
        CFG2["n"] = args.rows
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.debug_extra:
        bx = Bench(args)
        fn = {"config2": lambda: run_config2(bx, 5, 3), "config4": lambda: run_config4(bx, 3, args.debug_rows or 10_000_000),
              "config5": lambda: run_config5(bx, args.debug_rows or 100_000_000)}[args.debug_extra]
        res = fn()
        if bx.rank == 0:
            print(json.dumps(res), flush=True)
        if bx.world > 1:
            bx.dist.barrier()
            bx.dist.destroy_process_group()
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
