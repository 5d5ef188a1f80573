#!/usr/bin/env python
"""bench.py — NIQKI hot path on B200: sketch -> index -> query, whole job per step.

Contract (one JSON line on rank 0):  python bench.py --gpus N --steps K --warmup W
  Every GPU runs BASELINE.json configs[1] at every N (weak scaling): 10k synthetic 5 Mbp genomes --index,
  then --query 1k mutated copies, K=31 S=15 W=12 H=4, minjac 0.1.  Rank r indexes genomes [r*G,(r+1)*G)
  (index sharded by genome id), sketches its slice of the queries, the query sketches are all-gathered
  through the library's own exchange step (nq_allgather_sketches: NCCL behind the C ABI, u16 on the wire),
  every shard counts all of them, the sorted per-shard hit lists go to the host and are merged on rank 0
  (nq_hits_merge) — all inside the timed step.  `value` = bases sketched by all ranks / max-over-ranks
  step time with the sequences resident in HBM; `e2e` = the same job through the host-buffer C ABI (pinned
  host characters in — packed to 2 bits per base by the library on the way —, exchange, merged hits out).
  The same line carries `query_100k`: query sketches/s of 10k queries against a 100k-genome index sharded
  over the N GPUs (strong scaling; BASELINE's second metric), `parity_ok` (merged hit lists against a
  brute-force count) and the rooflines.
--workload q100k / c4 / c5: secondary lines (query-only shapes; configs[3] reads; configs[4] --matrix tiles).
--impl reference times the reference's own CPU path (oracle/_ref, else the oracle port) on a bounded
sample with all host threads.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 42
GENOME_LEN = 5_000_000
K, S, W, H, J = 31, 15, 12, 4, 0.1
RATES = (0.001, 0.01, 0.05)
OPS_PER_BASE = 38  # algorithmic int32 ops per base of the sketch scan
METRIC = "Gbases/s sketched (configs[1] per GPU: index 10k x 5 Mbp genomes + query 1k mutated copies); query sketches/s vs 100k-genome index in query_100k"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genomes", type=int, default=0, help="genomes per GPU (default: 10000 = configs[1], at every N)")
    ap.add_argument("--queries", type=int, default=0, help="queries per GPU (default: genomes/10)")
    ap.add_argument("--genome-len", type=int, default=GENOME_LEN)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-q100k", action="store_true", help="skip the query_100k section of the contract line")
    ap.add_argument("--no-q100k-auto", action="store_true", help="N > 1: skip the replicated-layout timing of query_100k")
    ap.add_argument("--q100k-genomes", type=int, default=100_000)
    ap.add_argument("--q100k-queries", type=int, default=10_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-genomes", type=int, default=0, help="genomes per GPU in the e2e leg (default: all that fit host RAM)")
    ap.add_argument("--reads", type=int, default=0, help="c4: reads per GPU (default 10M)")
    ap.add_argument("--query-groups", type=int, default=1,
                    help="q100k: R query groups x (GPUs / R) genome shards; every shard is held by R GPUs, each counting 1/R of "
                         "the queries (default 1 = the north-star layout, one gid shard per GPU)")
    ap.add_argument("--workload", default="c2", choices=["c2", "q100k", "c4", "c5"],
                    help="c2: the contract line (default).  q100k: secondary line, query sketches/s against a 100k-genome "
                         "index split over the GPUs (BASELINE metric, second half); genomes are sketched in batches.  "
                         "c4: configs[3], --indexlines/--querylines on 10M synthetic 150 bp reads (S=8).  "
                         "c5: configs[4], all-vs-all --matrix of 20k genomes at S=18, rows tiled over the GPUs")
    return ap.parse_args()


def thresholds(rates):
    return np.array([int(np.ldexp(np.longdouble(r), 64)) if r > 0 else 0 for r in rates], dtype=np.uint64)


def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except OSError:
        pass
    return 0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nme, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O

    cores = os.cpu_count() or 1
    use_ref = O.ref_available()
    kind = "reference" if use_ref else "port"
    o = O.Oracle(K=K, S=S, W=W, H=H, J=J)
    L = args.genome_len
    per_core = 4
    n_idx = max(8, min(cores * per_core, 512))
    n_q = max(1, n_idx // 10)
    # sample: n_idx genomes + n_q mutated copies, generated by the (multi-threaded) C generator
    seqs = [o.synth_genome(g, L) for g in range(n_idx)]
    seqs += [o.synth_mutant(q % n_idx, q, RATES[q % 3], L) for q in range(n_q)]
    bases, offs = O.concat_entries(seqs)
    del seqs
    nbases = int(offs[-1])
    if use_ref:
        r = O.Ref(K=K, S=S, W=W, H=H, J=J)  # 2^27 std::vector headers, ~5 s, untimed
        sk, _ = r.sketch_batch(bases, offs, nthreads=cores)
        for g in range(n_idx):
            r.insert_sketch(sk[g], g, f"g{g}")  # index built once, untimed (Index cannot be reset)
    else:
        sk = o.sketch_batch(bases, offs, cores)
        o.insert_sketches(sk[:n_idx])

    def step():
        t0 = time.perf_counter()
        if use_ref:
            s2, _ = r.sketch_batch(bases, offs, nthreads=cores, want_out=True)
            tq, _ = r.query_batch_timed(s2[n_idx:], nthreads=cores)
        else:
            s2 = o.sketch_batch(bases, offs, cores)
            o.query_batch(s2[n_idx:], cores)
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        step()
    times = [step() for _ in range(args.steps)]
    total = sum(times)
    value = nbases * args.steps / total / 1e9
    sample = (f"{n_idx} genomes + {n_q} mutated queries of {L} bp per step: compute_sketch on every entry "
              f"(OpenMP, {cores} threads) + query_sketch of the {n_q} queries against a {n_idx}-genome index built once")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Gbases/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "configs[1] sample: " + sample, "K": K, "S": S, "W": W, "H": H, "minjac": J},
            "cpu_baseline": {"value": value, "unit": "Gbases/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline(args, cores):
    """Bounded sample of the same workload on the host cores (rank 0, N=1 only): ~10-30 s of CPU."""
    from oracle import oracle as O

    use_ref = O.ref_available()
    o = O.Oracle(K=K, S=S, W=W, H=H, J=J)
    L = args.genome_len
    n_idx = max(8, min(cores * 4, 256))
    n_q = max(1, n_idx // 10)
    seqs = [o.synth_genome(g, L) for g in range(n_idx)] + [o.synth_mutant(q % n_idx, q, RATES[q % 3], L) for q in range(n_q)]
    bases, offs = O.concat_entries(seqs)
    del seqs
    t0 = time.perf_counter()
    if use_ref:
        r = O.Ref(K=K, S=S, W=W, H=H, J=J)
        sk, secs = r.sketch_batch(bases, offs, nthreads=cores)
        for g in range(n_idx):
            r.insert_sketch(sk[g], g, f"g{g}")
        tq, _ = r.query_batch_timed(sk[n_idx:], nthreads=cores)
        r.close()
    else:
        t1 = time.perf_counter()
        sk = o.sketch_batch(bases, offs, cores)
        secs = time.perf_counter() - t1
        o.insert_sketches(sk[:n_idx])
        t1 = time.perf_counter()
        o.query_batch(sk[n_idx:], cores)
        tq = time.perf_counter() - t1
    return {"value": int(offs[-1]) / secs / 1e9, "unit": "Gbases/s", "cores": cores,
            "kind": "reference" if use_ref else "port",
            "sample": f"compute_sketch on {n_idx}+{n_q} entries of {L} bp ({int(offs[-1])/1e9:.2f} Gbases) with {cores} threads",
            "query_sketches_per_s": n_q / tq if tq > 0 else None,
            "query_sample": f"query_sketch of {n_q} queries vs a {n_idx}-genome index, {cores} threads",
            "wall_s": time.perf_counter() - t0}


# ------------------------------------------------------------------------------------------------
def run_q100k(args):
    """Secondary line: query sketches/s against a 100k-genome index (configs[2]: 10k queries, minjac 0.1),
    the index split by gid over the GPUs (strong scaling: total work fixed).  Sequences never persist:
    genomes are generated and sketched in batches, only the sketches and the index stay in HBM."""
    import torch
    import torch.distributed as dist

    import niqki_b200
    from niqki_b200.capi import check, lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    Gt = args.genomes or 100_000
    Qt = args.queries or 10_000
    R = max(1, args.query_groups)
    if world % R:
        raise SystemExit("--query-groups must divide the number of GPUs")
    shards = world // R                # genome shards; rank r holds shard r % shards and serves query group r // shards
    shard, group = rank % shards, rank // shards
    G, Q, L = Gt // shards, Qt // world, args.genome_len
    Lc = lib()
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = niqki_b200.Context(local, stream)
    ix = niqki_b200.Index(S=S, K=K, W=W, H=H, min_fract=J, ctx=ctx)
    F = ix.F
    g0, q0 = shard * G, rank * Q
    B = 4000
    buf = torch.empty(B * L + 64, dtype=torch.uint8, device=dev)
    sk_idx = torch.empty((G, F), dtype=torch.int32, device=dev)
    sk_qry = torch.empty((Q, F), dtype=torch.int32, device=dev)
    t_build = time.perf_counter()
    for b0 in range(0, G, B):
        nb = min(B, G - b0)
        check(Lc.nq_synth_genomes_device(ctx.h, SEED, g0 + b0, nb, L, C.c_void_p(buf.data_ptr())))
        ix.compute_sketches(buf, np.arange(nb + 1, dtype=np.uint64) * L, out=sk_idx[b0:b0 + nb])
    ix.insert_sketches(sk_idx, gid_base=g0)
    del sk_idx
    for b0 in range(0, Q, B):
        nb = min(B, Q - b0)
        qid = np.arange(q0 + b0, q0 + b0 + nb, dtype=np.uint64)
        parents = (qid % np.uint64(Gt)).astype(np.uint64)
        thr = thresholds([RATES[int(q) % 3] for q in qid])
        check(Lc.nq_synth_mutants_device(ctx.h, SEED, parents.ctypes.data, qid.ctypes.data, thr.ctypes.data, nb, L,
                                         C.c_void_p(buf.data_ptr())))
        ix.compute_sketches(buf, np.arange(nb + 1, dtype=np.uint64) * L, out=sk_qry[b0:b0 + nb])
    del buf
    sk_all = torch.empty((Q * world, F), dtype=torch.int32, device=dev) if world > 1 else sk_qry
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    per_group = (Q * world) // R       # queries this rank counts: the group's slice of the all-gathered sketches

    def step():
        if world > 1:
            dist.all_gather_into_tensor(sk_all, sk_qry)
        ix.query_sketches(sk_all[group * per_group:(group + 1) * per_group] if R > 1 else sk_all, fetch=False)

    for _ in range(args.warmup):
        step()
    ctx.set_timing(True)
    ctx.timing_reset()
    launches0 = ctx.launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    q_ms, q_n = ctx.timing()["query"]
    gathered = ctx.last_query_gathered
    launches = ctx.launches - launches0
    ctx.set_timing(False)
    # e2e: host query sketches in, sorted hit lists out, through the C ABI
    h_sk = torch.empty((Q * world, F), dtype=torch.int32, pin_memory=True)
    h_sk.copy_(sk_all)
    torch.cuda.synchronize()
    n_sk = h_sk.numpy()[group * per_group:(group + 1) * per_group] if R > 1 else h_sk.numpy()
    ptr, cnt, gid = ix.query_sketches(n_sk)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ptr, cnt, gid = ix.query_sketches(n_sk)
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    first_hits = [int(gid[int(ptr[i])]) if ptr[i + 1] > ptr[i] else -1 for i in range(8)]
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        nq_total = Q * world
        q_bytes = 4 * gathered + (per_group if R > 1 else nq_total) * F * (8 + 2)  # this rank's launch
        q_gbs = q_bytes / (q_ms / max(q_n, 1) / 1e3) / 1e9 if q_ms else None
        info = ix.info()
        line = {"metric": "query sketches/s vs 100k-genome index (10k mutated-copy queries, minjac 0.1)",
                "value": nq_total * args.steps / (ms_total / 1e3), "unit": "query sketches/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "config": {"workload": f"secondary (not the contract line): {Gt} synthetic {L} bp genomes indexed, gid-sharded over "
                                       f"{world} GPU(s), {nq_total} mutated-copy queries all-gathered and counted on every shard"
                                       + (f"; layout {shards} gid shards x {R} query groups (each shard held by {R} GPUs, each "
                                          f"counting 1/{R} of the queries)" if R > 1 else ""),
                           "K": K, "S": S, "W": W, "H": H, "minjac": J, "genomes_per_gpu": G, "queries_total": nq_total, "query_groups": R,
                           "l2": "per-step index traffic far larger than L2 (>= 1 GB of postings gathered per GPU)"},
                "index_postings_per_gpu": info["n_postings"], "index_build_wall_s": t_build,
                "roofline": {"kernel": "query_count_kernel", "bound": "hbm", "achieved": q_gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": (q_gbs / hbm_peak) if q_gbs else None, "traffic": None,
                             "algorithmic_bytes": q_bytes, "gathered_postings": gathered, "launches": int(q_n),
                             "ms_per_launch": q_ms / max(q_n, 1),
                             "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback"},
                "e2e": {"value": nq_total * args.steps / float(dt.item()), "unit": "query sketches/s",
                        "h2d_bytes_per_step": int(n_sk.nbytes), "d2h_bytes_per_step": int(ptr.nbytes + cnt.nbytes + gid.nbytes),
                        "note": "nq_query_batch with pinned host sketches; sorted hit lists copied out"},
                "gpu_launches": int(launches), "clocks": clocks, "first_hits": first_hits}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _setup(args):
    import torch
    import torch.distributed as dist

    import niqki_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = niqki_b200.Context(local, stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    return torch, dist, niqki_b200, world, rank, local, dev, stream, ctx, barrier, timed


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return {}


def run_c4(args):
    """Secondary line, configs[3]: --indexlines on N synthetic 150 bp reads (S=8, W=12, K=31: SURVEY 8d),
    then --querylines on a subset.  Step = sketch + densify every read, build the index, count the
    query subset.  Reads are sharded by id over the GPUs like genomes."""
    from niqki_b200.capi import check, lib

    torch, dist, niqki_b200, world, rank, local, dev, stream, ctx, barrier, timed = _setup(args)
    S4, RL = 8, 150
    N = args.reads or 10_000_000
    Q = args.queries or 2_000
    Lc = lib()
    ix = niqki_b200.Index(S=S4, K=K, W=W, H=H, min_fract=J, ctx=ctx)
    F = ix.F
    r0 = rank * N
    d_reads = torch.empty(N * RL + 64, dtype=torch.uint8, device=dev)
    check(Lc.nq_synth_reads_device(ctx.h, SEED, r0, N, GENOME_LEN, RL, C.c_void_p(d_reads.data_ptr())))
    offs = np.arange(N + 1, dtype=np.uint64) * RL
    sk = torch.empty((N, F), dtype=torch.int32, device=dev)
    fl = torch.empty(N, dtype=torch.int32, device=dev)
    sk_all = torch.empty((Q * world, F), dtype=torch.int32, device=dev) if world > 1 else None
    torch.cuda.synchronize()

    def step_index():
        ix.compute_sketches(d_reads, offs, out=sk, flags=fl)
        ix.insert_sketches(sk, gid_base=r0)

    def step_query():
        q = sk[:Q]
        if world > 1:
            dist.all_gather_into_tensor(sk_all, q.contiguous())
            q = sk_all
        ix.query_sketches(q, fetch=False)

    for _ in range(args.warmup):
        step_index()
    step_query()
    ctx.set_timing(True)
    ctx.timing_reset()
    launches0 = ctx.launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_index = timed(step_index, args.steps)
    ms_query = timed(step_query, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    kt = ctx.timing()
    launches = ctx.launches - launches0
    gathered = ctx.last_query_gathered
    ctx.set_timing(False)
    ptr, cnt, gid = ix.query_sketches(sk[:8])
    first_hits = [int(gid[int(ptr[i])]) if ptr[i + 1] > ptr[i] else -1 for i in range(8)]
    # e2e: host reads in (pinned), sketches out, index built from them, sorted hits of the query subset out
    Ne = min(N, 2_000_000)
    h_reads = torch.empty(Ne * RL, dtype=torch.uint8, pin_memory=True)
    h_reads.copy_(d_reads[: Ne * RL])
    h_sk = torch.empty((Ne, F), dtype=torch.int32, pin_memory=True)
    torch.cuda.synchronize()
    n_reads, n_sk = h_reads.numpy(), h_sk.numpy()
    eo = np.arange(Ne + 1, dtype=np.uint64) * RL
    d2h = [0]

    def step_e2e():
        ix.compute_sketches(n_reads, eo, out=n_sk)
        ix.insert_sketches(n_sk, gid_base=r0)
        p_, c_, g_ = ix.query_sketches(n_sk[:Q])
        d2h[0] = n_sk.nbytes + p_.nbytes + c_.nbytes + g_.nbytes

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    step_e2e()
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        peaks = _peaks()
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        q_ms, q_n = kt["query"]
        nq_total = Q * world
        q_bytes = 4 * gathered + nq_total * F * (8 + 2)
        q_gbs = q_bytes / (q_ms / max(q_n, 1) / 1e3) / 1e9 if q_ms else None
        line = {"metric": "Gbases/s sketched + indexed (--indexlines on 150 bp reads)", "value": N * RL * world * args.steps / (ms_index / 1e3) / 1e9,
                "unit": "Gbases/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_index / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": f"secondary (not the contract line) configs[3]: {N} synthetic {RL} bp reads per GPU --indexlines, "
                                       f"then --querylines on {nq_total} of them, x{world} GPUs", "K": K, "S": S4, "W": W, "H": H, "minjac": J,
                           "l2": "inputs larger than L2 (1.5 GB of reads, 10 GB of sketches per step)"},
                "reads_per_s": N * world * args.steps / (ms_index / 1e3),
                "query_sketches_per_s": nq_total * args.steps / (ms_query / 1e3), "query_ms_per_step": ms_query / args.steps,
                "kernel_ms_per_step": {k: v[0] / args.steps for k, v in kt.items()},
                "roofline": {"kernel": "query_count_kernel", "bound": "hbm", "achieved": q_gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": (q_gbs / hbm_peak) if q_gbs else None, "traffic": None, "algorithmic_bytes": q_bytes,
                             "gathered_postings": gathered, "launches": int(q_n), "ms_per_launch": q_ms / max(q_n, 1)},
                "e2e": {"value": Ne * RL * world / float(dt.item()) / 1e9, "unit": "Gbases/s", "reads": Ne,
                        "h2d_bytes_per_step": int(n_reads.nbytes + n_sk.nbytes + Q * F * 4), "d2h_bytes_per_step": int(d2h[0])},
                "gpu_launches": int(launches), "clocks": clocks, "first_hits": first_hits}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_c5(args):
    """Secondary line, configs[4]: all-vs-all --matrix of N genomes at S=18, the genome x genome grid tiled
    over the GPUs: the index is sharded by genome id like everywhere else; the rows of one
    shard at a time are broadcast (nq_bcast_sketches, NCCL, u16 on the wire) and every GPU counts them
    against its own columns (nq_matrix_tile), dense counts to the host.  Timed: the whole matrix."""
    from niqki_b200.capi import check, lib
    from niqki_b200.shard import Comm, torch_bcast_bytes

    torch, dist, niqki_b200, world, rank, local, dev, stream, ctx, barrier, timed = _setup(args)
    S5 = 18
    N = args.genomes or 20_000
    L = args.genome_len
    Lc = lib()
    ix = niqki_b200.Index(S=S5, K=K, W=W, H=H, min_fract=J, ctx=ctx)
    comm = Comm(ctx, rank, world, torch_bcast_bytes(dist, dev)) if world > 1 else None
    F = ix.F
    per = (N + world - 1) // world
    g0, g1 = min(N, rank * per), min(N, (rank + 1) * per)
    B = 2000
    buf = torch.empty(B * L + 64, dtype=torch.uint8, device=dev)
    mine = torch.empty((max(g1 - g0, 1), F), dtype=torch.int32, device=dev)
    ctx.set_timing(True)
    ctx.timing_reset()
    torch.cuda.synchronize()
    t_sk = time.perf_counter()
    for b0 in range(g0, g1, B):
        nb = min(B, g1 - b0)
        check(Lc.nq_synth_genomes_device(ctx.h, SEED, b0, nb, L, C.c_void_p(buf.data_ptr())))
        ix.compute_sketches(buf, np.arange(nb + 1, dtype=np.uint64) * L, out=mine[b0 - g0:b0 - g0 + nb])
    torch.cuda.synchronize()
    t_sk = time.perf_counter() - t_sk
    scan_ms, scan_n = ctx.timing()["scan"]
    del buf
    t_ix = time.perf_counter()
    ix.insert_sketches(mine[: g1 - g0], gid_base=g0)
    torch.cuda.synchronize()
    t_ix = time.perf_counter() - t_ix
    RB = 1000   # rows per broadcast block
    d_rows = torch.empty((RB, F), dtype=torch.int32, device=dev)
    out = {}

    def step(check_tiles=False):
        diag = True
        incr = 0.0
        for owner in range(world):
            o0, o1 = min(N, owner * per), min(N, (owner + 1) * per)
            for b0 in range(o0, o1, RB):
                nb = min(RB, o1 - b0)
                if rank == owner:
                    d_rows[:nb].copy_(mine[b0 - g0:b0 - g0 + nb])
                if world > 1:
                    comm.bcast_sketches(ix.p, d_rows[:nb], owner)
                tile = ix.matrix_tile(d_rows[:nb], wrap16=True)    # [nb][g1 - g0] on the host
                if check_tiles:  # untimed pass only: checksum of the tiles and the diagonal (a genome against itself: F mod 2^16)
                    incr += float(tile.sum(dtype=np.float64))
                    if rank == owner:
                        diag = diag and bool(np.all(tile[np.arange(nb), np.arange(b0 - g0, b0 - g0 + nb)] == (F & 0xFFFF)))
        if check_tiles:
            out["diag"], out["incr"] = diag, incr

    step(check_tiles=True)
    ctx.timing_reset()
    launches0 = ctx.launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    agg = torch.tensor([out["incr"], 1.0 if out["diag"] else 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    clocks = sampler.stop() if rank == 0 else None
    m_ms, m_n = ctx.timing()["matrix"]
    launches = ctx.launches - launches0
    if rank == 0:
        sec = float(dt.item()) / args.steps
        peaks = _peaks()
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        info = ix.info()
        # K4b algorithmic bytes of this rank's tiles: the postings its probes gather (4 B each,
        # ~ the pair increments of wrapped counters are a lower bound; the true figure is the gathered ids),
        # row pairs + u16 sketch per probed cell, dense counts out
        m_bytes = 4.0 * float(agg[0].item()) / world + N * F * (8 + 2) + 4.0 * N * (g1 - g0)
        m_gbs = m_bytes / (m_ms / args.steps / 1e3) / 1e9 if m_ms else None
        line = {"metric": "genome pairs/s (all-vs-all --matrix, S=18)", "value": N * N / sec, "unit": "pairs/s", "n_gpus": world,
                "steps": args.steps, "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "config": {"workload": f"secondary (not the contract line) configs[4]: --matrix of {N} synthetic {L} bp genomes, S={S5}, "
                                       f"genome x genome grid tiled over {world} GPU(s): index sharded by genome id, row blocks of {RB} "
                                       f"broadcast (nq_bcast_sketches), every GPU counts them against its own columns (nq_matrix_tile)",
                           "K": K, "S": S5, "W": W, "H": H, "l2": "index (directory + postings) far larger than L2"},
                "matrix_device_ms_per_step": m_ms / args.steps, "matrix_kernel_launches": int(m_n),
                "pair_increments_per_s": float(agg[0].item()) / sec,
                "sketch_gbases_per_s": (g1 - g0) * L / (scan_ms / 1e3) / 1e9 if scan_ms else None, "sketch_wall_s": t_sk,
                "index_build_wall_s": t_ix, "index_bytes_per_gpu": info["device_bytes"],
                "diag_is_F_mod_65536": bool(agg[1].item() == world),
                "roofline": {"kernel": "matrix tiles = the query kernels with a dense epilogue", "bound": "hbm", "achieved": m_gbs, "peak": hbm_peak,
                             "unit": "GB/s", "frac": (m_gbs / hbm_peak) if m_gbs else None, "traffic": None, "algorithmic_bytes": m_bytes,
                             "note": "rank 0's tiles; gathered postings counted as the wrapped pair increments (a lower bound)"},
                "e2e": {"value": N * N / sec, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(4 * N * (g1 - g0)),
                        "note": "nq_matrix_tile: dense u32 counts of every tile copied to the host"},
                "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _brute_force_hits(torch, sk_shard, gid_base, sk_q, F, min_score, range_):
    """Independent check of the query path (torch eager, no library code): per-genome counts of equal
    valid fingerprints of `sk_q` rows against this rank's shard sketches, thresholded like :661-665."""
    out = []
    n = sk_shard.shape[0]
    for qi in range(sk_q.shape[0]):
        q = sk_q[qi]
        valid = (q >= 0) & (q < range_)
        cnt = torch.zeros(n, dtype=torch.int64, device=sk_shard.device)
        for g0 in range(0, n, 2048):
            blk = sk_shard[g0:g0 + 2048]
            cnt[g0:g0 + 2048] = ((blk == q[None, :]) & valid[None, :]).sum(1)
        hit = torch.nonzero(cnt >= max(min_score, 0)).flatten()
        out.append(sorted(((int(cnt[g]), gid_base + int(g)) for g in hit.tolist()), reverse=True))
    return out


def run_contract(args):
    """The contract line: configs[1] per GPU (weak scaling) + the BASELINE metric's second half in the
    same JSON line: query sketches/s against a 100k-genome index sharded over the N GPUs."""
    import torch
    import torch.distributed as dist

    import niqki_b200
    from niqki_b200.capi import check, lib
    from niqki_b200.shard import Comm, merge_hits, torch_bcast_bytes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    gloo = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)   # barriers + max-over-ranks of the timings
        gloo = dist.new_group(backend="gloo")             # per-rank hit lists travel to rank 0 on the host

    G = args.genomes or 10000          # configs[1] per GPU at every N
    Q = args.queries or max(1, G // 10)
    L = args.genome_len
    Lc = lib()
    stream = torch.cuda.Stream(device=dev)  # a real (non-NULL) stream shared by torch and the library
    torch.cuda.set_stream(stream)
    ctx = niqki_b200.Context(local, stream)
    # the ranks of one box share its cores: each rank's packer gets its share
    ctx.set_host_packing(-1, max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world)))))
    ix = niqki_b200.Index(S=S, K=K, W=W, H=H, min_fract=J, ctx=ctx)
    F = ix.F
    comm = Comm(ctx, rank, world, torch_bcast_bytes(dist, dev)) if world > 1 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def gather_merge(part):
        """Per-shard (ptr, counts, gids) -> merged lists on rank 0 (nq_hits_merge); None elsewhere."""
        if world == 1:
            return part
        parts = [None] * world if rank == 0 else None
        dist.gather_object(part, parts, dst=0, group=gloo)
        return merge_hits(parts) if rank == 0 else None

    # ---- synthetic inputs, generated in HBM (untimed)
    g0, q0 = rank * G, rank * Q
    d_idx = torch.empty(G * L + 64, dtype=torch.uint8, device=dev)
    check(Lc.nq_synth_genomes_device(ctx.h, SEED, g0, G, L, C.c_void_p(d_idx.data_ptr())))
    qid = np.arange(q0, q0 + Q, dtype=np.uint64)
    parents = (qid % np.uint64(G * world)).astype(np.uint64)  # queries are copies of genomes 0..Q*world-1
    thr = thresholds([RATES[int(q) % 3] for q in qid])
    d_qry = torch.empty(Q * L + 64, dtype=torch.uint8, device=dev)
    check(Lc.nq_synth_mutants_device(ctx.h, SEED, parents.ctypes.data, qid.ctypes.data, thr.ctypes.data, Q, L,
                                     C.c_void_p(d_qry.data_ptr())))
    idx_offs = np.arange(G + 1, dtype=np.uint64) * L
    qry_offs = np.arange(Q + 1, dtype=np.uint64) * L
    sk_idx = torch.empty((G, F), dtype=torch.int32, device=dev)
    sk_qry = torch.empty((Q, F), dtype=torch.int32, device=dev)
    sk_all = torch.empty((Q * world, F), dtype=torch.int32, device=dev) if world > 1 else sk_qry
    fl_idx = torch.empty(G, dtype=torch.int32, device=dev)
    fl_qry = torch.empty(Q, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    last = {}

    def step_device():
        ix.compute_sketches(d_idx, idx_offs, out=sk_idx, flags=fl_idx)
        ix.insert_sketches(sk_idx, gid_base=g0)
        ix.compute_sketches(d_qry, qry_offs, out=sk_qry, flags=fl_qry)
        if world > 1:
            comm.allgather_sketches(ix.p, sk_qry, sk_all)       # nq_allgather_sketches: NCCL, u16 on the wire
        part = ix.query_sketches(sk_all)                          # hit lists on the host, sorted (:685)
        last["hits"] = gather_merge(part)                         # nq_hits_merge of the shards' lists on rank 0

    for _ in range(args.warmup):
        step_device()
    ctx.set_timing(True)
    ctx.timing_reset()
    launches0 = ctx.launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    kt = ctx.timing()
    launches = ctx.launches - launches0
    gathered = ctx.last_query_gathered
    ctx.set_timing(False)
    bases_per_step = (G + Q) * L * world
    value = bases_per_step * args.steps / (ms_total / 1e3) / 1e9
    info = ix.info()

    # ---- parity of the merged result: the first queries against a brute-force count over the shards'
    # sketches (torch eager, every rank its own shard, merged on rank 0 by a plain sort)
    ncheck = min(64 if world > 1 else 16, Q * world)
    bf = _brute_force_hits(torch, sk_idx, g0, sk_all[:ncheck], F, ix.min_score, int(ix.p.range))
    if world > 1:
        allbf = [None] * world if rank == 0 else None
        dist.gather_object(bf, allbf, dst=0, group=gloo)
    else:
        allbf = [bf]
    parity_ok = None
    first_hits = None
    if rank == 0:
        ptr, cnt, gid = last["hits"]
        parity_ok = True
        for qi in range(ncheck):
            exp = sorted((h for r in allbf for h in r[qi]), reverse=True)
            got = list(zip(cnt[int(ptr[qi]):int(ptr[qi + 1])].tolist(), gid[int(ptr[qi]):int(ptr[qi + 1])].tolist()))
            parity_ok = parity_ok and got == exp
        first_hits = [int(gid[int(ptr[i])]) if ptr[i + 1] > ptr[i] else -1 for i in range(min(8, Q))]

    # ---- e2e: the same job through the host-buffer C ABI: pinned host sequences in (packed to 2 bits
    # per base by the library on the way, K1), sketches stay in HBM, query sketches exchanged over NCCL,
    # sorted hit lists out and merged on rank 0 — every step.
    e2e = None
    if not args.no_e2e:
        avail = mem_available_bytes() // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
        Ge = args.e2e_genomes or G
        while Ge > 64 and avail and (Ge + Q) * L * 1.6 > avail:
            Ge //= 2
        Qe = max(1, min(Q, Ge // 10))
        h_idx = torch.empty(Ge * L, dtype=torch.uint8, pin_memory=True)
        h_qry = torch.empty(Qe * L, dtype=torch.uint8, pin_memory=True)
        h_idx.copy_(d_idx[: Ge * L]); h_qry.copy_(d_qry[: Qe * L])
        torch.cuda.synchronize()
        n_idx_b, n_qry_b = h_idx.numpy(), h_qry.numpy()
        eo_idx = np.arange(Ge + 1, dtype=np.uint64) * L
        eo_qry = np.arange(Qe + 1, dtype=np.uint64) * L
        e_sk_idx, e_sk_qry = sk_idx[:Ge], sk_qry[:Qe]
        e_sk_all = torch.empty((Qe * world, F), dtype=torch.int32, device=dev) if world > 1 else e_sk_qry
        h_flags = np.zeros(max(Ge, Qe), np.uint32)
        d2h = [0]

        def step_e2e():
            ix.sketch_records_to_device(n_idx_b, eo_idx, e_sk_idx, h_flags)      # H2D (packed), sketches stay on the device
            ix.insert_sketches(e_sk_idx, gid_base=g0)
            ix.sketch_records_to_device(n_qry_b, eo_qry, e_sk_qry, h_flags)
            if world > 1:
                comm.allgather_sketches(ix.p, e_sk_qry, e_sk_all)
            part = ix.query_sketches(e_sk_all)                                    # D2H sorted hits
            d2h[0] = sum(int(x.nbytes) for x in part) + 4 * (Ge + Qe)
            last["e2e_hits"] = gather_merge(part)

        step_e2e()  # one warm-up pass (staging buffers, packer threads)
        h2d0 = ctx.h2d_bytes
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 2))
        for _ in range(e2e_steps):
            step_e2e()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_bases = (Ge + Qe) * L * world
        e2e = {"value": e2e_bases * e2e_steps / float(dt.item()) / 1e9, "unit": "Gbases/s",
               "h2d_bytes_per_step": int((ctx.h2d_bytes - h2d0) // e2e_steps), "d2h_bytes_per_step": int(d2h[0]),
               "host_bytes_read_per_step": int((Ge + Qe) * L), "genomes": Ge, "queries": Qe, "steps": e2e_steps,
               "note": "nq_sketch_records (host characters packed to 2 bits/base by the library, K1) / nq_index_build_device / "
                       "nq_allgather_sketches / nq_query_batch_device + nq_hits_merge; h2d = bytes that crossed PCIe"}
        del h_idx, h_qry
    del d_idx, d_qry, sk_idx, sk_qry, sk_all
    ix.close_index()
    torch.cuda.empty_cache()

    q100k = None
    if not args.no_q100k:
        q100k = q100k_section(args, torch, dist, niqki_b200, ctx, comm, gloo, world, rank, local, dev, stream, barrier, gather_merge)

    if rank == 0:
        peaks = _peaks()
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        scan_ms, scan_n = kt["scan"]
        q_ms, q_n = kt["query"]
        bases_per_scan = (G + Q) * L * args.steps / max(scan_n, 1)
        scan_s = scan_ms / max(scan_n, 1) / 1e3
        scan_gbases = bases_per_scan / scan_s / 1e9 if scan_ms else None
        nq_total = Q * world
        q_bytes = 4 * gathered + nq_total * F * (8 + 2)  # gids + row pairs + u16 sketch (+ 8 B/hit, negligible)
        q_gbs = q_bytes / (q_ms / max(q_n, 1) / 1e3) / 1e9 if q_ms else None
        sm_clk = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
        int_peak, int_src = _int_peak(peaks, sm_clk)
        b_ms, b_n = kt["cell_sort"]
        s_ms, s_n = kt["slab"]
        elem = 2 if G <= 65400 else 4
        b_bytes = info["n_postings"] * 2 * elem + F * (1 << W) * 2 * elem
        b_gbs = b_bytes / (b_ms / max(b_n, 1) / 1e3) / 1e9 if b_ms else None
        line = {
            "metric": METRIC, "value": value, "unit": "Gbases/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"configs[1] on every GPU: {G} synthetic {L} bp genomes --index + {Q} mutated copies --query per GPU, "
                                   f"x{world} GPUs (index sharded by genome id; query sketches all-gathered, hits merged on rank 0)",
                       "K": K, "S": S, "W": W, "H": H, "minjac": J, "genomes_per_gpu": G, "queries_per_gpu": Q,
                       "l2": "inputs larger than L2 (>= 50 GB of sequence per step)", "parallelism": f"shard-by-gid x{world}",
                       "collective": "nq_allgather_sketches (NCCL, u16 on the wire)" if world > 1 else "none"},
            "query_sketches_per_s": nq_total / (q_ms / max(q_n, 1) / 1e3) if q_ms else None,
            "index_postings": info["n_postings"],
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in kt.items()},
            "roofline": {"kernel": "sketch_scan_kernel", "bound": "int32-alu",
                         "achieved": scan_gbases * OPS_PER_BASE / 1e3 if scan_gbases else None, "peak": int_peak,
                         "unit": "Tint32-op/s", "frac": (scan_gbases * OPS_PER_BASE / 1e3 / int_peak) if scan_gbases else None,
                         "traffic": _ncu_traffic("sketch_scan_kernel", G, nq_total, bases_per_scan / (G * L)),
                         "algorithmic_ops_per_base": OPS_PER_BASE, "gbases_per_s": scan_gbases,
                         "launches": int(scan_n), "ms_per_launch": scan_ms / max(scan_n, 1), "peak_source": int_src,
                         "hbm": {"achieved": scan_gbases, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": scan_gbases / hbm_peak if scan_gbases else None,
                                 "note": "1 B/base algorithmic: this kernel is not HBM-bound"}},
            "roofline_query": {"kernel": "query_count_seg_kernel (CSR segment-table form: probed lists of <= 12 ids)" if G * 7.1e-4 <= 12
                               else "slab_resolve_kernel + query_slab_kernel (granule form)", "bound": "hbm", "achieved": q_gbs, "peak": hbm_peak,
                               "unit": "GB/s", "frac": (q_gbs / hbm_peak) if q_gbs else None,
                               "traffic": _ncu_traffic("query_count_seg_kernel" if G * 7.1e-4 <= 12 else "query_slab_kernel", G, nq_total),
                               "peak_source": peak_src,
                               "algorithmic_bytes": q_bytes, "gathered_postings": gathered,
                               "launches": int(q_n), "ms_per_launch": q_ms / max(q_n, 1)},
            "roofline_build": {"kernel": "cell_build_kernel", "bound": "hbm", "achieved": b_gbs, "peak": hbm_peak, "unit": "GB/s",
                               "frac": (b_gbs / hbm_peak) if b_gbs else None, "traffic": _ncu_traffic("cell_build_kernel", G, nq_total),
                               "algorithmic_bytes": b_bytes, "launches": int(b_n), "ms_per_launch": b_ms / max(b_n, 1),
                               "slab_ms_per_build": s_ms / max(args.steps, 1)},
            "query_100k": q100k,
            "parity_ok": parity_ok, "parity_checked_queries": ncheck,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "first_hits": first_hits,
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline(args, os.cpu_count() or 1)
            except Exception as exc:  # the baseline must never take the bench line down
                line["cpu_baseline"] = {"error": repr(exc)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def q100k_section(args, torch, dist, niqki_b200, ctx, comm, gloo, world, rank, local, dev, stream, barrier, gather_merge):
    """BASELINE metric, second half: query sketches/s of 10k mutated-copy queries (minjac 0.1) against a
    100k-genome index sharded by genome id over the N GPUs (strong scaling: total work fixed).  A step =
    all-gather of the query sketches (each rank holds 1/N of them) + counting on every shard + sorted hit
    lists to the host + merge on rank 0 — all inside the timed region.  When N > 1 the same queries are
    also timed in the layout the library would pick when the whole index fits one GPU's HBM (every GPU
    holds all shards, queries are split, no exchange)."""
    from niqki_b200.capi import check, lib
    from niqki_b200.shard import merge_hits

    Lc = lib()
    Gt, Qt, L = args.q100k_genomes, args.q100k_queries, args.genome_len
    G, Q = Gt // world, Qt // world
    ix = niqki_b200.Index(S=S, K=K, W=W, H=H, min_fract=J, ctx=ctx)
    F = ix.F
    g0, q0 = rank * G, rank * Q
    B = 2500
    buf = torch.empty(B * L + 64, dtype=torch.uint8, device=dev)
    sk_idx = torch.empty((G, F), dtype=torch.int32, device=dev)
    sk_qry = torch.empty((Q, F), dtype=torch.int32, device=dev)
    t_build = time.perf_counter()
    for b0 in range(0, G, B):
        nb = min(B, G - b0)
        check(Lc.nq_synth_genomes_device(ctx.h, SEED, g0 + b0, nb, L, C.c_void_p(buf.data_ptr())))
        ix.compute_sketches(buf, np.arange(nb + 1, dtype=np.uint64) * L, out=sk_idx[b0:b0 + nb])
    ix.insert_sketches(sk_idx, gid_base=g0)
    for b0 in range(0, Q, B):
        nb = min(B, Q - b0)
        qid = np.arange(q0 + b0, q0 + b0 + nb, dtype=np.uint64)
        parents = (qid % np.uint64(Gt)).astype(np.uint64)
        thr = thresholds([RATES[int(q) % 3] for q in qid])
        check(Lc.nq_synth_mutants_device(ctx.h, SEED, parents.ctypes.data, qid.ctypes.data, thr.ctypes.data, nb, L,
                                         C.c_void_p(buf.data_ptr())))
        ix.compute_sketches(buf, np.arange(nb + 1, dtype=np.uint64) * L, out=sk_qry[b0:b0 + nb])
    del buf
    sk_all = torch.empty((Q * world, F), dtype=torch.int32, device=dev) if world > 1 else sk_qry
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build
    last = {}

    def step():
        if world > 1:
            comm.allgather_sketches(ix.p, sk_qry, sk_all)
        last["hits"] = gather_merge(ix.query_sketches(sk_all))

    def timed_wall(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    steps = max(1, min(args.steps, 5))
    for _ in range(2):
        step()
    ctx.set_timing(True)
    ctx.timing_reset()
    ms_total = timed_wall(step, steps)
    q_ms, q_n = ctx.timing()["query"]
    gathered = ctx.last_query_gathered
    ctx.set_timing(False)
    info = ix.info()

    # parity of the merged lists: first queries against a brute-force count over the shards' sketches
    ncheck = min(64, Q * world)
    bf = _brute_force_hits(torch, sk_idx, g0, sk_all[:ncheck], F, ix.min_score, int(ix.p.range))
    if world > 1:
        allbf = [None] * world if rank == 0 else None
        dist.gather_object(bf, allbf, dst=0, group=gloo)
    else:
        allbf = [bf]
    parity_ok = None
    if rank == 0:
        ptr, cnt, gid = last["hits"]
        parity_ok = True
        for qi in range(ncheck):
            exp = sorted((h for r in allbf for h in r[qi]), reverse=True)
            got = list(zip(cnt[int(ptr[qi]):int(ptr[qi + 1])].tolist(), gid[int(ptr[qi]):int(ptr[qi + 1])].tolist()))
            parity_ok = parity_ok and got == exp

    # the layout the library picks when every shard fits one GPU: all shards on every GPU, queries split
    auto = None
    if world > 1 and not args.no_q100k_auto:
        shards = [ix]
        all_idx = torch.empty((Gt, F), dtype=torch.int32, device=dev)
        comm.allgather_sketches(ix.p, sk_idx, all_idx)
        full = niqki_b200.Index(S=S, K=K, W=W, H=H, min_fract=J, ctx=ctx)
        ix.close_index()
        full.insert_sketches(all_idx, gid_base=0)
        del all_idx

        def step_auto():
            last["auto"] = full.query_sketches(sk_qry)   # this rank's queries against the whole index: final lists

        for _ in range(2):
            step_auto()
        ms_auto = timed_wall(step_auto, steps)
        auto = {"layout": f"whole index on each of the {world} GPUs (fits HBM: {full.info()['device_bytes'] / 1e9:.1f} GB), queries split {world} ways, no exchange",
                "value": Qt * steps / (ms_auto / 1e3), "unit": "query sketches/s", "ms_per_step": ms_auto / steps}
        full.close()
        del shards
    del sk_idx
    out = None
    if rank == 0:
        peaks = _peaks()
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        q_bytes = 4 * gathered + Qt * F * (8 + 2)  # this rank's launch: all queries against its shard
        q_gbs = q_bytes / (q_ms / max(q_n, 1) / 1e3) / 1e9 if q_ms else None
        out = {"metric": "query sketches/s vs 100k-genome index (10k mutated-copy queries, minjac 0.1)",
               "value": Qt * steps / (ms_total / 1e3), "unit": "query sketches/s", "ms_per_step": ms_total / steps, "steps": steps,
               "scaling": "strong", "layout": f"north star: index sharded by genome id, {G} genomes per GPU x {world}; query sketches "
                                              f"all-gathered (nq_allgather_sketches), hits merged on rank 0 (nq_hits_merge), inside the timed region",
               "genomes": Gt, "queries": Qt, "index_postings_per_gpu": info["n_postings"], "index_bytes_per_gpu": info["device_bytes"],
               "build_wall_s": t_build, "parity_ok": parity_ok, "parity_checked_queries": ncheck,
               "count_kernel_ms_per_step": q_ms / steps if q_ms else None,
               "roofline": {"kernel": "query kernels of this shard size (slab form up to 65.4k genomes, split16 segment form above)", "bound": "hbm",
                            "achieved": q_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": (q_gbs / hbm_peak) if q_gbs else None,
                            "traffic": _ncu_traffic("query_100k", G, Qt), "algorithmic_bytes": q_bytes, "gathered_postings": gathered,
                            "launches": int(q_n), "ms_per_launch": q_ms / max(q_n, 1)},
               "auto_layout": auto}
    ix.close()
    return out


def _int_peak(peaks, sm_clk):
    """INT32 roofline denominator: the issue-rate microbenchmark of tools/microbench.cu when its record is
    committed (profiles/int32_peak.json: measured int32 ops per clock per SM), else lanes x clock."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "int32_peak.json")))
        per_clk = float(rec["int32_ops_per_clk_per_sm"])
        return 148 * per_clk * sm_clk * 1e6 / 1e12, f"measured: {per_clk:.1f} int32 ops/clk/SM (tools/microbench.cu, {rec.get('source', 'profiles/int32_peak.json')}) x 148 SMs x SM clock sampled during the run"
    except (OSError, ValueError, KeyError):
        return 148 * 128 * sm_clk * 1e6 / 1e12, "148 SMs x 128 int32 lanes x SM clock sampled during the run (no measured INT32 figure available)"


def _ncu_traffic(kernel, genomes_per_gpu, queries, scale=1.0):
    """DRAM bytes per launch from a committed ncu capture of the SAME configuration, else None."""
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except (OSError, ValueError):
        return None
    t = traffic.get(kernel)
    if isinstance(t, list):
        t = next((x for x in t if x.get("genomes_per_gpu") == genomes_per_gpu and x.get("queries") == queries), None)
    ok = t and t.get("genomes_per_gpu") == genomes_per_gpu and t.get("queries") == queries
    return int(t["dram_bytes_per_launch"] * scale) if ok else None


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "q100k":
        return run_q100k(args)
    if args.workload == "c4":
        return run_c4(args)
    if args.workload == "c5":
        return run_c5(args)
    return run_contract(args)


if __name__ == "__main__":
    main()
