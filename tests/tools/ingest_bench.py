#!/usr/bin/env python
"""Host-side ingest throughput: native readers of csrc/ingest.cu against the
per-object / per-record Python path they replace.  CPU only; prints one JSON line.

    python tests/tools/ingest_bench.py [--sigs 2000] [--genomes 16]
"""
import argparse
import gzip
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sourmash_b200.sigset import SignatureSet            # noqa: E402
from sourmash_b200.sketch import RecordBatch             # noqa: E402
from sourmash_b200.synth import synth_genome, synth_sketches  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sigs", type=int, default=2000)
ap.add_argument("--genomes", type=int, default=16)
ap.add_argument("--threads", type=int, default=0)
a = ap.parse_args()

out = {"cores": len(os.sched_getaffinity(0))}
with tempfile.TemporaryDirectory() as td:
    # ---- .sig files: one sketch of ~5000 hashes per file (config 3 shape) ----------------
    h, off = synth_sketches(a.sigs, mean=5000, sd=500, lo=3000, hi=7000, n_families=20, pool=6000, seed=3)
    paths = []
    for i in range(a.sigs):
        row = h[int(off[i]):int(off[i + 1])]
        doc = [{"class": "sourmash_signature", "email": "", "hash_function": "0.murmur64", "filename": f"g{i}.fa",
                "name": f"genome {i}", "license": "CC0", "version": 0.4,
                "signatures": [{"num": 0, "ksize": 31, "seed": 42, "max_hash": 18446744073709552,
                                "mins": row.tolist(), "md5sum": "0" * 32, "molecule": "DNA"}]}]
        p = os.path.join(td, f"g{i}.sig")
        with open(p, "w") as fh:
            json.dump(doc, fh, separators=(",", ":"))
        paths.append(p)
    nbytes = sum(os.path.getsize(p) for p in paths)
    t0 = time.perf_counter()
    ss = SignatureSet.from_files(paths, a.threads)
    t_native = time.perf_counter() - t0
    assert len(ss) == a.sigs and int(ss.offsets[-1]) == int(off[-1])
    assert np.array_equal(ss.mins, h)
    t0 = time.perf_counter()
    rows = []
    for p in paths[: max(a.sigs // 10, 1)]:                # python json, a tenth of the files
        with open(p) as fh:
            for rec in json.load(fh):
                for sk in rec["signatures"]:
                    rows.append(np.array(sk["mins"], dtype=np.uint64))
    t_py = (time.perf_counter() - t0) * (a.sigs / max(a.sigs // 10, 1))
    out["sig"] = {"files": a.sigs, "MB": round(nbytes / 1e6, 1), "native_s": round(t_native, 3),
                  "native_MBps": round(nbytes / 1e6 / t_native, 1), "python_json_s": round(t_py, 3),
                  "speedup": round(t_py / t_native, 1)}
    # ---- the same signatures as one .zip collection (members signatures/<i>.sig.gz, like `sig cat -o x.zip`) ----
    import zipfile
    zpath = os.path.join(td, "db.zip")
    with zipfile.ZipFile(zpath, "w", compression=zipfile.ZIP_STORED) as z:
        for i, p in enumerate(paths):
            with open(p, "rb") as fh:
                z.writestr(f"signatures/{i}.sig.gz", gzip.compress(fh.read(), compresslevel=6))
    t0 = time.perf_counter()
    zs = SignatureSet.from_files([zpath], a.threads)
    t_native = time.perf_counter() - t0
    assert len(zs) == a.sigs and np.array_equal(zs.mins, h)
    t0 = time.perf_counter()
    with zipfile.ZipFile(zpath) as z:                      # python zipfile + gzip + json, a tenth of the members
        for name in z.namelist()[: max(a.sigs // 10, 1)]:
            for rec in json.loads(gzip.decompress(z.read(name))):
                for sk in rec["signatures"]:
                    rows.append(np.array(sk["mins"], dtype=np.uint64))
    t_py = (time.perf_counter() - t0) * (a.sigs / max(a.sigs // 10, 1))
    out["zip"] = {"members": a.sigs, "MB": round(os.path.getsize(zpath) / 1e6, 1), "json_MB": round(nbytes / 1e6, 1),
                  "native_s": round(t_native, 3), "python_zipfile_json_s": round(t_py, 3),
                  "speedup": round(t_py / t_native, 1)}
    # ---- FASTA.gz genomes ------------------------------------------------------------------
    fpaths = []
    for g in range(a.genomes):
        seq = bytes(synth_genome(5_000_000, seed=1000 + g))
        p = os.path.join(td, f"g{g}.fa.gz")
        with gzip.open(p, "wb", compresslevel=1) as fh:
            fh.write(b">genome %d\n" % g)
            for i in range(0, len(seq), 80):
                fh.write(seq[i:i + 80] + b"\n")
        fpaths.append(p)
    t0 = time.perf_counter()
    rb = RecordBatch(fpaths, a.threads)
    t_native = time.perf_counter() - t0
    assert len(rb) == a.genomes and rb.total_bytes == a.genomes * 5_000_000
    t0 = time.perf_counter()
    with gzip.open(fpaths[0], "rb") as fh:                 # python: one file
        data = fh.read()
    recs = [c.partition(b"\n")[2].replace(b"\n", b"") for c in data.split(b">")[1:]]
    t_py = (time.perf_counter() - t0) * a.genomes
    out["fasta_gz"] = {"files": a.genomes, "Mbp": a.genomes * 5, "native_s": round(t_native, 3),
                       "native_Mbp_per_s": round(a.genomes * 5 / t_native, 1), "python_gzip_s": round(t_py, 3),
                       "speedup": round(t_py / t_native, 1)}
print(json.dumps(out))
