#!/usr/bin/env python
"""bench.py -- headline benchmark of sourmash_b200 (contract: see the task statement / DESIGN.md §Measurement).

Primary workload (BASELINE.json configs[2]): `compare` of 10 000 synthetic FracMinHash sketches
(k=31, scaled=1000, ~5000 hashes each, 100 families) -> all-vs-all Jaccard matrix; metric
sketch-pairs/s.  Secondary workload (configs[1], reported under "sketch"): `sketch dna`
k=21,31,51 scaled=1000 over 100 synthetic 5 Mbp genomes; metric k-mers hashed/s.

    python bench.py --gpus 1 --steps 5 --warmup 3                  # this framework on cuda:0
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                           # CPU restatement on host cores

One JSON line on stdout (rank 0).  Only the cpu_baseline / --impl reference legs touch oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SKETCHES = 10_000
KSIZES = (21, 31, 51)
N_GENOMES = 100
COMPARE_WORKLOAD = ("configs[2]: compare 10000 synthetic sketches (k=31, scaled=1000, ~5000 hashes, "
                    "100 families) all-vs-all jaccard float64 matrix")
SKETCH_WORKLOAD = "configs[1]: sketch dna k=21,31,51 scaled=1000 on 100 synthetic 5 Mbp genomes"
N_DB_SEARCH, N_DB_GATHER, N_QUERY_SEARCH = 300_000, 50_000, 10_000_000   # configs[3] / configs[4] shapes
SM_COUNT = 148                                        # B200
GENOME_LEN = 5_000_000
SCALED = 1000


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly one JSON line: keep a private handle to the real stdout and point fd 1
# at stderr, so that library banners (e.g. "NCCL version ..." printed from C) cannot pollute it.
_JSON_FD = None


def protect_stdout():
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit_json(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """`nvidia-smi -lms 20` running for the whole benchmark (it needs ~0.3 s to deliver its first
    line, longer than some timed regions); every timed region reports the samples whose arrival
    time falls inside it, widened to the nearest samples when the region is shorter than the
    sampling period."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def window(self, t_begin, t_end):
        "Clock statistics of the samples taken in [t_begin, t_end] (monotonic seconds)."
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)                                  # let the last in-region line arrive
        rows = list(self.rows)
        inside = [r for r in rows if t_begin <= r[0] <= t_end + 0.03]
        note = "timed region"
        if len(inside) < 2:                               # region shorter than two sampling periods
            mid = 0.5 * (t_begin + t_end)
            inside = sorted(rows, key=lambda r: abs(r[0] - mid))[:3]
            note = "3 samples nearest to the timed region (region shorter than two 20 ms sampling periods)"
        sm, mx, reasons = [], None, set()
        for _, r in inside:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(self.NAMES, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "samples": len(sm), "reasons": sorted(reasons), "window": note}


# ----------------------------------------------------------------------------- workloads
def compare_workload():
    from sourmash_b200.synth import synth_sketches
    t = time.time()
    h, off = synth_sketches(N_SKETCHES)
    log(f"[bench] compare workload: {N_SKETCHES} sketches, {len(h)} hashes ({time.time() - t:.1f}s)")
    return h, off


def sketch_workload(n_genomes=N_GENOMES):
    from sourmash_b200.synth import synth_genome
    t = time.time()
    total = n_genomes * GENOME_LEN
    seqs = np.empty(total, dtype=np.uint8)
    for g in range(n_genomes):
        seqs[g * GENOME_LEN:(g + 1) * GENOME_LEN] = synth_genome(GENOME_LEN, 1000 + g)
    offs = (np.arange(n_genomes + 1, dtype=np.uint64) * np.uint64(GENOME_LEN))
    log(f"[bench] sketch workload: {n_genomes} genomes x {GENOME_LEN} bp ({time.time() - t:.1f}s)")
    return seqs, offs


def host_cores():
    """Usable host threads: the smallest of cpu_count, the affinity mask and the cgroup CPU quota."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        with open("/sys/fs/cgroup/cpu.max") as fh:
            quota, period = fh.read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period))))
    except Exception:
        pass
    return n


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def git_blob_hash(path):
    "the id `git hash-object` gives the file: sha1 over 'blob <size>\\0' + content"
    import hashlib
    with open(path, "rb") as fh:
        data = fh.read()
    return hashlib.sha1(b"blob %d\0" % len(data) + data).hexdigest()


def ncu_record(key):
    """An ncu-measured constant (DRAM bytes per launch, warp instructions per k-mer) from profiles/ncu_traffic.json,
    or None when it is absent or STALE: every entry lists the kernel source files it was measured on with their
    git blob ids, and counts only while those files are byte-identical (scripts/record_traffic.py writes entries)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as fh:
        entry = json.load(fh).get("entries", {}).get(key)
    if not entry or not entry.get("files"):
        return None
    for rel, blob in entry["files"].items():
        path = os.path.join(ROOT, rel)
        if not os.path.exists(path) or git_blob_hash(path) != blob:
            return None
    return entry


def ncu_traffic(key):
    entry = ncu_record(key)
    return entry["value"] if entry else None


def ncu_source(key):
    entry = ncu_record(key)
    return entry.get("source") if entry else None


# ----------------------------------------------------------------------------- reference arm
def cpu_compare_sample(h, off, ncores, target_pairs):
    """Time the oracle's compare_serial restatement on rows [0, R) (all columns j > i)."""
    import oracle as orc
    n = len(off) - 1
    rows, pairs = 0, 0
    while rows < n and pairs < target_pairs:
        pairs += n - 1 - rows
        rows += 1
    t = time.perf_counter()
    orc.compare_all_pairs(h, off, first_row=0, n_rows=rows, nthreads=ncores)
    dt = time.perf_counter() - t
    return pairs, dt, f"rows 0..{rows - 1} x all later columns of the {n}-sketch matrix = {pairs} pairs"


def cpu_sketch_sample(seqs, offs, ncores, n_genomes):
    import oracle as orc
    sub_off = offs[: n_genomes + 1]
    sub = seqs[: int(sub_off[-1])]
    mx = orc.max_hash_for_scaled(SCALED)
    kmers, t = 0, time.perf_counter()
    for k in KSIZES:
        orc.sketch_batch(sub, sub_off, k, mx, nthreads=ncores, cap_per_seq=20000)
        kmers += int(sum(int(sub_off[i + 1] - sub_off[i]) - k + 1 for i in range(n_genomes)))
    dt = time.perf_counter() - t
    return kmers, dt, f"{n_genomes} of the {N_GENOMES} genomes x k={list(KSIZES)} = {kmers} k-mers"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = host_cores()
    if args.workload == "compare":
        h, off = compare_workload()
        target = int(1.5e5 * ncores)            # ~8 s per step at ~1.5e4 pairs/s/core
        times = []
        for i in range(args.warmup + args.steps):
            units, dt, sample = cpu_compare_sample(h, off, ncores, target)
            if i >= args.warmup:
                times.append(dt)
        metric, unit = "sketch-pairs/sec (compare)", "pairs/s"
        config = {"workload": COMPARE_WORKLOAD, "n_sketches": N_SKETCHES,
                  "pairs_per_step": N_SKETCHES * (N_SKETCHES - 1) // 2,
                  "parallelism": "%d host threads (OpenMP over rows)" % ncores}
    else:
        ng = N_GENOMES                           # the whole configs[1] workload (a few seconds on 16 cores)
        seqs, offs = sketch_workload(ng)
        times = []
        for i in range(args.warmup + args.steps):
            units, dt, sample = cpu_sketch_sample(seqs, offs, ncores, ng)
            if i >= args.warmup:
                times.append(dt)
        metric, unit = "k-mers hashed/sec (sketch)", "k-mers/s"
        config = {"workload": SKETCH_WORKLOAD, "parallelism": "%d host threads (one genome per thread)" % ncores}
    ms = 1000.0 * float(np.mean(times))
    value = units / (ms / 1000.0)
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": unit, "cores": ncores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "CPU restatement of src/core (oracle/oracle.c, OpenMP), not the Rust binary: no Rust toolchain in the image"}
    emit_json(line)


# ----------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        # keep stdout clean for the single JSON line: NCCL's banner / debug output goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from sourmash_b200 import batch as B
    B.set_device(local_rank)
    stream = torch.cuda.current_stream()
    B.set_stream(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                                   # runs until the process exits (daemon reader)
        import atexit
        atexit.register(sampler.stop)

    def timed(step_fn, steps, warmup):
        for _ in range(warmup):
            step_fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = B.kernel_launches()
        t_begin = time.monotonic()
        e0.record(stream)
        extra = [step_fn() for _ in range(steps)]
        e1.record(stream)
        barrier()
        t_end = time.monotonic()
        ms = e0.elapsed_time(e1)
        clocks = sampler.window(t_begin, t_end) if rank == 0 else None
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, B.kernel_launches() - launches0, clocks, extra

    hbm_peak, peak_src = peaks()
    if args.workload in ("search", "gather"):
        r = bench_search_gather(args, torch, dist, B, rank, world, timed, args.workload)
        if rank == 0:
            r.update({"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "data": "synthetic", "dtype": "u64"})
            emit_json(r)
        if dist is not None:
            dist.destroy_process_group()
        return
    out = {}
    if args.workload in ("compare", "both"):
        out["compare"] = bench_compare(args, torch, dist, B, rank, world, timed, hbm_peak, peak_src)
    if args.workload in ("sketch", "both"):
        out["sketch"] = bench_sketch(args, torch, dist, B, rank, world, timed, hbm_peak, peak_src)
    if rank == 0:
        primary = out.get("compare") or out.get("sketch")
        line = dict(primary)
        if "compare" in out and "sketch" in out:
            line["sketch"] = out["sketch"]
        emit_json(line)
    if dist is not None:
        dist.destroy_process_group()


def bench_compare(args, torch, dist, B, rank, world, timed, hbm_peak, peak_src):
    h, off = compare_workload()
    n = len(off) - 1
    n_pairs = n * (n - 1) // 2
    sizes = np.diff(off.astype(np.int64))
    # algorithmic bytes of the intersection kernel: every pair reads both rows once, writes a u32
    alg_bytes = 8.0 * (n - 1) * float(sizes.sum()) + 4.0 * n_pairs
    dev = torch.device("cuda")
    B.set_profiling(True)

    if world == 1:
        sset = B.SketchSet.from_host(h, off)
        d_out = torch.empty((n, n), dtype=torch.float64, device=dev)
        kernel_ms = []

        def step():
            B.compare_jaccard_device(sset, d_out.data_ptr())

        def step_prof():
            step()
            kernel_ms.append(B.last_kernel_ms(0))

        ms, launches, clocks, _ = timed(step_prof, args.steps, args.warmup)
        kernel_ms = kernel_ms[-args.steps:]
        # end to end: host CSR (pinned) -> H2D -> kernels -> D2H of the float64 matrix (pinned)
        ph, po = B.pinned_empty(len(h), np.uint64), B.pinned_empty(len(off), np.uint64)
        ph.array[:] = h
        po.array[:] = off
        pout = B.pinned_empty((n, n), np.float64)

        def step_e2e():
            s2 = B.SketchSet.from_host(ph.array, po.array)
            B.compare_jaccard(s2, out=pout.array)
            return float(pout.array[0, 1])

        ms_e2e, _, _, _ = timed(step_e2e, max(2, args.steps // 2), 1)
        h2d, d2h = int(h.nbytes + off.nbytes), int(n * n * 8)
        parallelism = "1 gpu"
    else:
        # shard rows across ranks as if each rank had sketched its own genomes; one all-gather
        # of the shards, cyclic row tiles per rank, all-reduce of the partial count matrices,
        # each rank finalises a block of rows.
        from sourmash_b200.distributed import CompareShard
        cs = CompareShard(torch, dist, B, h, off, rank, world)
        kernel_ms = []

        def step_prof():
            cs.step(e2e=False)
            kernel_ms.append(B.last_kernel_ms(0))

        ms, launches, clocks, _ = timed(step_prof, args.steps, args.warmup)
        kernel_ms = kernel_ms[-args.steps:]
        ms_e2e, _, _, _ = timed(lambda: cs.step(e2e=True), max(2, args.steps // 2), 1)
        h2d, d2h = cs.h2d_bytes, cs.d2h_bytes
        alg_bytes = alg_bytes / world
        parallelism = f"{world} gpus: sketches all-gathered, row tiles cyclic, counts all-reduced"

    value = n_pairs / (ms / 1e3)
    kms = float(np.mean(kernel_ms))
    achieved = alg_bytes / (kms / 1e3) / 1e9
    plan = B.last_compare_plan()
    if plan["algo"] == "join":
        knote = ("counts by sorting the hashes of the set once and incrementing one counter per pair of rows sharing a hash "
                 "(planner estimate: %.3g increments over %.3g elements); algorithmic bytes keep SURVEY 8d's "
                 "definition, 8*(|A|+|B|) per pair + 4 B out" % (plan["est_increments"], plan["est_elements"]))
        if os.environ.get("SMB_JOIN_LAYOUT", "") == "plain":      # A/B runs only: the global-reduction join
            kname = "inverted join, global reductions (join_gather + cub::DeviceRadixSort + join_count_kernel)"
            tkey = "inverted_join"
            plan = dict(plan, layout="plain")
        else:
            kname = ("inverted join, stripe layout (stripe_keys + cub::DeviceRadixSort on 32-bit keys + stripe_tag + "
                     "join_stripe_kernel + stripe_mirror; counters in shared memory, no global atomics)")
            tkey = "inverted_join_stripe"
            plan = dict(plan, layout=os.environ.get("SMB_JOIN_LAYOUT") or "stripe")
    else:
        kname, tkey = "pairwise_tile_split_kernel", "pairwise_tile_split_kernel"
        knote = ("algorithmic bytes = 8*(|A|+|B|) per pair + 4 B out; rows are reused from shared "
                 "memory / L2, so DRAM traffic is far below this")
    res = {
        "metric": "sketch-pairs/sec (compare)", "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": COMPARE_WORKLOAD,
                   "n_sketches": n, "pairs_per_step": n_pairs, "parallelism": parallelism,
                   "l2_policy": "inputs 400 MB + 800 MB output per step exceed the 126 MB L2; no flush"},
        "e2e": {"value": n_pairs / (ms_e2e / 1e3), "unit": "pairs/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"kernel": kname, "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                     "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": ncu_traffic(tkey),
                     "kernel_ms": kms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                     "algorithm": plan, "note": knote},
    }
    traffic = res["roofline"]["traffic"]
    if traffic:                                           # what actually crossed the HBM interface (ncu), same launches
        res["roofline"]["dram"] = {"bytes_per_launch": traffic, "achieved": traffic / world / (kms / 1e3) / 1e9,
                                   "unit": "GB/s", "frac": traffic / world / (kms / 1e3) / 1e9 / hbm_peak}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:      # CPU baseline: rank 0 at N=1 only
        ncores = host_cores()
        units, dt, sample = cpu_compare_sample(h, off, ncores, int(1.5e5 * ncores))
        res["cpu_baseline"] = {"value": units / dt, "unit": "pairs/s", "cores": ncores, "kind": "port",
                               "sample": sample, "seconds": dt}
    return res


def bench_sketch(args, torch, dist, B, rank, world, timed, hbm_peak, peak_src):
    seqs, offs = sketch_workload()
    ng = len(offs) - 1
    mine = list(range(rank, ng, world))                   # genomes dealt round-robin to ranks
    my_off = np.zeros(len(mine) + 1, dtype=np.uint64)
    my_off[1:] = np.cumsum([int(offs[g + 1] - offs[g]) for g in mine])
    my_seqs = np.concatenate([seqs[int(offs[g]):int(offs[g + 1])] for g in mine]) if world > 1 else seqs
    total_kmers = sum(int(offs[g + 1] - offs[g]) - k + 1 for g in range(ng) for k in KSIZES)
    my_kmers = sum(int(offs[g + 1] - offs[g]) - k + 1 for g in mine for k in KSIZES)
    dev = torch.device("cuda")
    d_bases = torch.empty(len(my_seqs) + 64, dtype=torch.uint8, device=dev)
    d_bases[: len(my_seqs)].copy_(torch.from_numpy(my_seqs))
    lens = np.diff(my_off.astype(np.int64)).astype(np.uint64)
    B.set_profiling(True)
    kernel_ms = []

    def gather_shards(sset):
        if dist is None:
            return
        from sourmash_b200.distributed import allgather_sketchset
        allgather_sketchset(torch, dist, B, sset)

    def step():
        sset, nk = B.sketch_streams_device(d_bases.data_ptr(), my_off[:-1], lens, KSIZES, scaled=SCALED)
        assert nk == my_kmers
        kernel_ms.append(B.last_kernel_ms(1))
        gather_shards(sset)
        return sset

    ms, launches, clocks, _ = timed(step, args.steps, args.warmup)
    kernel_ms = kernel_ms[-args.steps:]
    pin = B.pinned_empty(len(my_seqs), np.uint8)
    pin.array[:] = my_seqs

    def step_e2e():
        sset, nk = B.sketch_sequences(pin.array, my_off, KSIZES, scaled=SCALED)
        gather_shards(sset)
        hh, oo = sset.to_host()
        return int(oo[-1])

    ms_e2e, _, _, ex = timed(step_e2e, max(2, args.steps // 2), 1)
    kms = float(np.mean(kernel_ms))
    alg_bytes = float(my_kmers) * (1.0 + 8.0 / SCALED)
    achieved = alg_bytes / (kms / 1e3) / 1e9
    res = {
        "metric": "k-mers hashed/sec (sketch)", "value": total_kmers / (ms / 1e3), "unit": "k-mers/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": SKETCH_WORKLOAD,
                   "kmers_per_step": total_kmers,
                   "parallelism": "1 gpu" if world == 1 else f"{world} gpus: genomes round-robin, sketches all-gathered",
                   "l2_policy": "500 MB of bases per pass exceed the 126 MB L2; no flush"},
        "e2e": {"value": total_kmers / (ms_e2e / 1e3), "unit": "k-mers/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(len(my_seqs)), "d2h_bytes_per_step": int(ex[-1]) * 8},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"kernel": "hash_kmers_kernel (3 launches, k=21,31,51)", "bound": "hbm", "achieved": achieved,
                     "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": ncu_traffic("hash_kmers_kernel"), "kernel_ms": kms,
                     "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                     "note": "integer-issue bound by construction (~150 int ops per k-mer vs 1 B): "
                             "HBM fraction is expected to be small; roofline.issue is the bound that applies"},
    }
    fused = os.environ.get("SMB_SKETCH_FUSED") != "0"      # default: k=21,31,51 in one pass; =0: three launches (A/B runs)
    if fused:
        res["roofline"]["kernel"] = "hash_kmers_fused_kernel (1 launch, k=21,31,51)"
        res["roofline"]["traffic"] = ncu_traffic("hash_kmers_fused_kernel")
    wipk = ncu_traffic("hash_kmers_fused_kernel_warp_instr_per_kmer" if fused else "hash_kmers_kernel_warp_instr_per_kmer")
    if wipk:
        # the bound that applies: warp instructions issued (ncu count of the same three launches, per k-mer)
        # against 4 issue slots per SM per clock at the SM clock sampled during the timed region
        clk = float((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
        issued = wipk * my_kmers / (kms / 1e3)
        res["roofline"]["issue"] = {"warp_instructions_per_kmer": wipk, "achieved": issued / 1e9,
                                    "peak": SM_COUNT * 4 * clk / 1e9, "unit": "G warp-instr/s",
                                    "frac": issued / (SM_COUNT * 4 * clk),
                                    "source": ncu_source("hash_kmers_fused_kernel_warp_instr_per_kmer" if fused else "hash_kmers_kernel_warp_instr_per_kmer")}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ncores = host_cores()
        units, dt, sample = cpu_sketch_sample(seqs, offs, ncores, ng)
        res["cpu_baseline"] = {"value": units / dt, "unit": "k-mers/s", "cores": ncores, "kind": "port",
                               "sample": sample, "seconds": dt}
    return res


def _maybe_index(args, B, db):
    """--index: build the inverted index of the resident database once, outside the timed region (it
    belongs to loading the database, like the upload); the build time is reported beside the result."""
    if not getattr(args, "index", False):
        return {}
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n_keys = db.build_index()
    torch.cuda.synchronize()
    return {"index": {"distinct_hashes": int(n_keys), "build_ms": (time.perf_counter() - t0) * 1e3,
                      "note": "inverted index (hash -> rows) of the resident database, built once at load"}}


def bench_search_gather(args, torch, dist, B, rank, world, timed, which):
    """configs[3] / configs[4] shapes (parity-test cases, not the headline): one query vs a
    large resident database; reported for completeness, single GPU only."""
    from sourmash_b200.synth import MAX_HASH_1000, rows_of, synth_sketches
    h, off = synth_sketches(N_SKETCHES)
    rows = rows_of(h, off)
    rng = np.random.Generator(np.random.PCG64(4000))
    if world > 1:
        return bench_search_gather_sharded(args, torch, dist, B, rank, world, timed, which, h, off, rows, rng)
    if which == "search":
        n_db = N_DB_SEARCH
        reps = n_db // N_SKETCHES
        db_h = np.tile(h, reps)
        db_off = np.concatenate([[0], np.cumsum(np.tile(np.diff(off.astype(np.int64)), reps))]).astype(np.uint64)
        planted = rng.choice(N_SKETCHES, size=100, replace=False)
        query = np.unique(np.concatenate([rng.integers(1, MAX_HASH_1000, size=N_QUERY_SEARCH, dtype=np.uint64)] +
                                         [rows[j][: len(rows[j]) // 2] for j in planted]))
        db = B.SketchSet.from_host(db_h, db_off)
        index_info = _maybe_index(args, B, db)
        pq = B.pinned_empty(len(query), np.uint64)
        pq.array[:] = query
        ms, launches, clocks, ex = timed(lambda: int(B.one_vs_many(pq.array, db).sum()), args.steps, args.warmup)
        alg = 8.0 * (len(db_h) + len(query))
        return {"metric": "query-vs-DB passes/sec (search)", "value": 1e3 / ms, "unit": "queries/s", "ms_per_step": ms,
                "config": {"workload": "configs[3]: 1e7-hash query vs 300000-sketch DB (12 GB resident), containment counts",
                           "db_hashes": int(len(db_h)), "query_hashes": int(len(query))},
                "subjects_per_s": n_db / (ms / 1e3), "algorithmic_GBps": alg / (ms / 1e3) / 1e9, "gpu_launches": launches,
                "note": "query uploaded from pinned host memory every step; counts downloaded", **index_info}
    n_db = N_DB_GATHER
    reps = n_db // N_SKETCHES
    db_h = np.tile(h, reps)
    db_off = np.concatenate([[0], np.cumsum(np.tile(np.diff(off.astype(np.int64)), reps))]).astype(np.uint64)
    planted = rng.choice(N_SKETCHES, size=200, replace=False)
    query = np.unique(np.concatenate([rows[j][rng.random(len(rows[j])) < 0.6] for j in planted] +
                                     [rng.integers(1, MAX_HASH_1000, size=20_000, dtype=np.uint64)]))
    db = B.SketchSet.from_host(db_h, db_off)
    index_info = _maybe_index(args, B, db)
    res = {}

    def step():
        ids, sizes = B.gather(query, db, threshold=50)
        res["rounds"] = len(ids)
        return len(ids)

    ms, launches, clocks, ex = timed(step, args.steps, args.warmup)
    return {"metric": "gather wall time", "value": ms, "unit": "ms", "higher_is_better": False, "ms_per_step": ms,
            "config": {"workload": "configs[4]: ~1e5-hash query vs 50000-sketch DB, 200 planted overlapping matches, "
                                   "threshold 50 hashes", "query_hashes": int(len(query)), "db_hashes": int(len(db_h))},
            "rounds": res["rounds"], "rounds_per_s": res["rounds"] / (ms / 1e3), "gpu_launches": launches, **index_info}


def bench_search_gather_sharded(args, torch, dist, B, rank, world, timed, which, h, off, rows, rng):
    """configs[3] / configs[4] on N GPUs: the database sharded by subject (rank r keeps rows
    [b[r], b[r+1]) resident, built locally), the query replicated; search = local counts + one all-gather; gather =
    the session rounds with (best count, row) all-gathered and the winner's intersection broadcast
    (sourmash_b200.distributed.ShardedDatabase).  Same databases and queries as the single-GPU workloads."""
    from sourmash_b200.distributed import ShardedDatabase, shard_bounds
    from sourmash_b200.synth import MAX_HASH_1000
    n_db = N_DB_SEARCH if which == "search" else N_DB_GATHER
    b = shard_bounds(n_db, world)
    lo, hi = b[rank], b[rank + 1]
    sizes = np.diff(off.astype(np.int64))
    idx = np.arange(lo, hi) % N_SKETCHES                                     # the tiled database, this rank's rows only
    local_off = np.concatenate([[0], np.cumsum(sizes[idx])]).astype(np.uint64)
    local_h = np.concatenate([rows[j] for j in idx]) if len(idx) else np.zeros(0, np.uint64)
    local = B.SketchSet.from_host(local_h, local_off)
    index_info = _maybe_index(args, B, local)
    db = ShardedDatabase(torch, dist, B, local, n_db, lo)
    if which == "search":
        planted = rng.choice(N_SKETCHES, size=100, replace=False)
        query = np.unique(np.concatenate([rng.integers(1, MAX_HASH_1000, size=N_QUERY_SEARCH, dtype=np.uint64)] +
                                         [rows[j][: len(rows[j]) // 2] for j in planted]))
        ms, launches, clocks, ex = timed(lambda: int(db.search_counts(query).sum()), args.steps, args.warmup)
        return {"metric": "query-vs-DB passes/sec (search)", "value": 1e3 / ms, "unit": "queries/s", "ms_per_step": ms,
                "config": {"workload": "configs[3]: 1e7-hash query vs 300000-sketch DB sharded by subject over %d GPUs" % world,
                           "db_hashes_per_rank": int(len(local_h)), "query_hashes": int(len(query)),
                           "parallelism": "%d gpus: database sharded by subject, query replicated, counts all-gathered" % world},
                "subjects_per_s": n_db / (ms / 1e3), "gpu_launches": launches, "scaling": "strong", **index_info}
    planted = rng.choice(N_SKETCHES, size=200, replace=False)
    query = np.unique(np.concatenate([rows[j][rng.random(len(rows[j])) < 0.6] for j in planted] +
                                     [rng.integers(1, MAX_HASH_1000, size=20_000, dtype=np.uint64)]))
    res = {}

    def step():
        ids, _sizes = db.gather(query, threshold=50)
        res["rounds"] = len(ids)
        return len(ids)

    ms, launches, clocks, ex = timed(step, args.steps, args.warmup)
    return {"metric": "gather wall time", "value": ms, "unit": "ms", "higher_is_better": False, "ms_per_step": ms,
            "config": {"workload": "configs[4]: ~1e5-hash query vs 50000-sketch DB sharded by subject over %d GPUs, "
                                   "200 planted overlapping matches, threshold 50 hashes" % world,
                       "query_hashes": int(len(query)), "db_hashes_per_rank": int(len(local_h)),
                       "parallelism": "%d gpus: database sharded by subject; per round an all-gather of (count, row) and a "
                                      "broadcast of the winner's intersection" % world},
            "rounds": res["rounds"], "rounds_per_s": res["rounds"] / (ms / 1e3), "gpu_launches": launches,
            "scaling": "strong", **index_info}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=["compare", "sketch", "both", "search", "gather"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--index", action="store_true",
                    help="search / gather workloads: query through the inverted index of the resident database")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "compare" if args.impl == "reference" else "both"
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    protect_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
