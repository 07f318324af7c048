#!/usr/bin/env python
"""Measurements for the widened rows on one GPU; prints one JSON object.

  protein    `sketch protein` / `sketch translate` throughput (hash_aa_kernel), with the oracle port
             timed on a sample on the host cores, and the kernel's HBM roofline fraction
  files      end-to-end `sketch dna` from FASTA files (native reader -> pinned buffer -> GPU -> .sig JSON)
  sigs       bulk .sig load (native parser) -> SketchSet -> N x N compare

    python tests/tools/bench_extra.py [--what protein,files,sigs]
"""
import argparse
import gzip
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--what", default="protein,files,sigs")
ap.add_argument("--genomes", type=int, default=100)
ap.add_argument("--sigs", type=int, default=2000)
a = ap.parse_args()
what = set(a.what.split(","))

import torch  # noqa: E402
from sourmash_b200 import batch as B  # noqa: E402
from sourmash_b200.synth import synth_genome, synth_sketches  # noqa: E402

assert torch.cuda.is_available()
stream = torch.cuda.current_stream()
B.set_stream(stream.cuda_stream)
out = {"gpu": torch.cuda.get_device_name(0), "host_cores": len(os.sched_getaffinity(0))}
with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
    HBM = float(json.load(fh)["hbm_gbs"])


def gpu_time(fn, steps=5, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


if "protein" in what:
    import oracle as orc
    ng = 20
    dna = np.concatenate([synth_genome(5_000_000, 1000 + g) for g in range(ng)])
    offs = (np.arange(ng + 1, dtype=np.uint64) * np.uint64(5_000_000))
    rng = np.random.default_rng(1)
    alphabet = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    prot = alphabet[rng.integers(0, 20, size=100_000_000)]
    poffs = (np.arange(101, dtype=np.uint64) * np.uint64(1_000_000))
    res = {}
    for tag, seqs, so, mol, inp, ks, scaled in (
            ("translate_protein_k7_k10", dna, offs, "protein", False, [7, 10], 200),
            ("translate_dayhoff_k16", dna, offs, "dayhoff", False, [16], 200),
            ("translate_hp_k42", dna, offs, "hp", False, [42], 200),
            ("protein_k7_k10", prot, poffs, "protein", True, [7, 10], 200)):
        B.set_profiling(True)
        nk = [0]
        kms = []

        def step():
            s, n = B.sketch_sequences(seqs, so, ks, scaled=scaled, moltype=mol, input_is_protein=inp)
            nk[0] = n
            kms.append(B.last_kernel_ms(1))
        ms = gpu_time(step)
        kernel_ms = float(np.mean(kms[-5:]))
        # oracle port on a sample: first record(s), all k, one thread per record
        t = time.perf_counter()
        sample = bytes(seqs[: int(so[1])])
        cnt = 0
        for k in ks:
            hh = (orc.seq_to_hashes_protein if inp else orc.seq_to_hashes_translate)(sample, k, mol)
            cnt += len(hh)
        cpu_dt = time.perf_counter() - t
        bytes_in = float(len(seqs)) * len(ks)
        res[tag] = {"kmers": nk[0], "ms_e2e_host_input": round(ms, 3), "kmers_per_s_e2e": nk[0] / ms * 1e3,
                    "hash_kernel_ms": round(kernel_ms, 3), "kmers_per_s_kernel": nk[0] / kernel_ms * 1e3,
                    "roofline": {"bound": "hbm", "achieved": bytes_in / kernel_ms / 1e6, "peak": HBM, "unit": "GB/s",
                                 "frac": bytes_in / kernel_ms / 1e6 / HBM,
                                 "note": "1 B read per window start per ksize; integer-issue bound like the DNA kernel"},
                    "cpu_port_1thread_kmers_per_s": cnt / cpu_dt}
    out["protein"] = res

if "files" in what:
    from sourmash_b200.sketch import sketch_fasta_files
    import sourmash_b200 as smb
    td = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    plain, gz = [], []
    for g in range(a.genomes):
        seq = bytes(synth_genome(5_000_000, 1000 + g))
        body = b">genome %d\n" % g + b"\n".join(seq[i:i + 80] for i in range(0, len(seq), 80)) + b"\n"
        p = os.path.join(td, f"g{g}.fa")
        with open(p, "wb") as fh:
            fh.write(body)
        plain.append(p)
        if g < 32:
            with gzip.open(p + ".gz", "wb", compresslevel=1) as fh:
                fh.write(body)
            gz.append(p + ".gz")
    res = {}
    for tag, paths in (("fasta_plain", plain), ("fasta_gz", gz)):
        ts = []
        for it in range(4):
            t = time.perf_counter()
            sigs = sketch_fasta_files(paths, ksizes=[21, 31, 51], scaled=1000)
            text = smb.save_signatures_to_json(sigs)
            ts.append(time.perf_counter() - t)
        kmers = sum(5_000_000 - k + 1 for k in (21, 31, 51)) * len(paths)
        # phases of the last iteration, timed separately
        from sourmash_b200.sketch import RecordBatch
        t0 = time.perf_counter()
        rb = RecordBatch(paths)
        t1 = time.perf_counter()
        sset, _ = rb.sketch(rb.files.copy(), len(paths), [21, 31, 51], scaled=1000)
        B.synchronize()
        t2 = time.perf_counter()
        del rb, sset
        t3 = time.perf_counter()
        sigs = sketch_fasta_files(paths, ksizes=[21, 31, 51], scaled=1000)
        t4 = time.perf_counter()
        text = smb.save_signatures_to_json(sigs)
        t5 = time.perf_counter()
        res[tag] = {"files": len(paths), "wall_s_best": round(min(ts[1:]), 4), "kmers_per_s": kmers / min(ts[1:]),
                    "sig_json_bytes": len(text),
                    "phases_s": {"read_parse": round(t1 - t0, 4), "upload_hash_sort": round(t2 - t1, 4),
                                 "all_but_json": round(t4 - t3, 4), "sig_json": round(t5 - t4, 4)}}
    out["files"] = res

if "sigs" in what:
    from sourmash_b200.sigset import compare_signature_files
    td = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    h, off = synth_sketches(a.sigs, mean=5000, sd=500, lo=3000, hi=7000, n_families=20, pool=6000, seed=3)
    paths = []
    for i in range(a.sigs):
        doc = [{"class": "sourmash_signature", "email": "", "hash_function": "0.murmur64", "filename": f"g{i}.fa",
                "name": f"genome {i}", "license": "CC0", "version": 0.4,
                "signatures": [{"num": 0, "ksize": 31, "seed": 42, "max_hash": 18446744073709552,
                                "mins": h[int(off[i]):int(off[i + 1])].tolist(), "md5sum": "0" * 32, "molecule": "DNA"}]}]
        p = os.path.join(td, f"g{i}.sig")
        with open(p, "w") as fh:
            json.dump(doc, fh, separators=(",", ":"))
        paths.append(p)
    nbytes = sum(os.path.getsize(p) for p in paths)
    ts = []
    for it in range(3):
        t = time.perf_counter()
        m, labels = compare_signature_files(paths, ksize=31)
        ts.append(time.perf_counter() - t)
    t = time.perf_counter()
    for p in paths[: a.sigs // 10]:
        with open(p) as fh:
            json.load(fh)
    py = (time.perf_counter() - t) * 10
    out["sigs"] = {"files": a.sigs, "MB": round(nbytes / 1e6, 1), "load_plus_compare_s": round(min(ts), 4),
                   "python_json_load_only_s": round(py, 3), "matrix": list(m.shape)}
print(json.dumps(out))
