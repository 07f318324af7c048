"""Randomised differential campaign of the emulated library against the oracle (adversarial keys: 0, UINT64_MAX,
dense small integers, clashing low words, empty / huge rows; queries below and above the shared-memory limit).
Test infrastructure; run by hand:  [SMB_FUZZ_SWITCHED=1] python tests/host_emul/fuzz_emulated.py [seconds] [seed] [sets|sketch|protein]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import emulated_boot  # noqa: E402

emulated_boot.install()

import numpy as np  # noqa: E402

import oracle as orc  # noqa: E402
from sourmash_b200 import batch as B  # noqa: E402

BIG = np.uint64(2**64 - 1)
SWITCHED = os.environ.get("SMB_FUZZ_SWITCHED") == "1"      # also drive the stripe / cluster / ranges / index paths


def random_row(rng, scale):
    kind = rng.integers(0, 8)
    n = int(rng.integers(0, scale))
    if kind == 0:
        return np.zeros(0, np.uint64)
    if kind == 1:
        r = rng.integers(0, 2**64 - 1, size=n, dtype=np.uint64)
    elif kind == 2:
        r = rng.integers(0, max(2 * n, 4), size=n, dtype=np.uint64)                  # dense small integers
    elif kind == 3:
        r = (rng.integers(0, 50, size=n, dtype=np.uint64) << np.uint64(32)) | np.uint64(0xabcdef)   # one low word
    else:
        r = rng.integers(1, 2**54, size=n, dtype=np.uint64)
    extra = [x for x, p in ((np.uint64(0), 0.2), (BIG, 0.2), (BIG - np.uint64(1), 0.1)) if rng.random() < p]
    return np.unique(np.concatenate([r, np.array(extra, dtype=np.uint64)]))


def trial(rng):
    n = int(rng.integers(1, 40)) if rng.random() < 0.7 else int(rng.integers(40, 130))   # > 32 rows: whole blocks of swizzled counter columns
    scale = int(rng.choice([5, 60, 600])) if n < 40 else int(rng.choice([5, 60]))
    rows = [random_row(rng, scale) for _ in range(n)]
    for _ in range(int(rng.integers(0, 4))):                                        # related rows
        i, j = rng.integers(0, n, size=2)
        if len(rows[i]):
            rows[j] = np.unique(np.concatenate([rows[j], rows[i][rng.random(len(rows[i])) < 0.6]]))
    h, off = orc.to_csr(rows)
    db = B.SketchSet.from_host(h, off)
    want = orc.compare_all_pairs(h, off, nthreads=2)
    for algo in ("tile", "join"):
        os.environ["SMB_COMPARE_ALGO"] = algo
        got = B.compare_jaccard(db)
        assert np.array_equal(got, want), ("compare", algo, n, scale)
    if SWITCHED:                                                                     # the paths behind switches
        for layout, tags in (("stripe_full", "u16"), ("stripe", "u32"), ("plain", "u16")):
            os.environ["SMB_JOIN_LAYOUT"], os.environ["SMB_STRIPE_TAGS"] = layout, tags
            assert np.array_equal(B.compare_jaccard(db), want), ("compare", layout, tags, n, scale)
        os.environ.pop("SMB_JOIN_LAYOUT"); os.environ.pop("SMB_STRIPE_TAGS")
        for key in ("SMB_STRIPE_CTAS", "SMB_STRIPE_SWIZZLE"):                        # one CTA per SM / counters in column order
            os.environ[key] = "1" if key.endswith("CTAS") else "0"
            assert np.array_equal(B.compare_jaccard(db), want), ("compare", key, n, scale)
            os.environ.pop(key)
        os.environ["SMB_SEARCH_LAYOUT"] = "global"
        if rng.random() < 0.5 and len(h) and len(h) < 2**31:
            db.build_index()
    os.environ.pop("SMB_COMPARE_ALGO")
    qkind = rng.integers(0, 3)
    if qkind == 0:
        q = rows[int(rng.integers(0, n))]
    elif qkind == 1:
        q = np.unique(np.concatenate([random_row(rng, 3000)] + [rows[int(rng.integers(0, n))]]))
    else:                                                                            # above the shared-memory limit
        q = np.unique(np.concatenate([rng.integers(0, 2**64 - 1, size=40_000, dtype=np.uint64),
                                      rows[int(rng.integers(0, n))], np.array([0, BIG] if rng.random() < 0.5 else [], dtype=np.uint64)]))
    if len(h):
        wc = orc.one_vs_many(q, h, off).astype(np.uint32) if len(q) else np.zeros(n, np.uint32)
        assert np.array_equal(B.one_vs_many(q, db), wc), ("one_vs_many", qkind, n, scale, len(q))
        if len(q):
            ids, sizes = B.gather(q, db, threshold=1)
            cur, cnt, ref = q.copy(), np.array([orc.count_common(q, r) for r in rows], dtype=np.int64), []
            while True:
                j = int(np.argmax(cnt))
                if cnt[j] < 1:
                    break
                isect = np.intersect1d(cur, rows[j])
                ref.append((j, len(isect)))
                cnt = cnt - np.array([orc.count_common(isect, r) for r in rows], dtype=np.int64)
                cur = np.setdiff1d(cur, isect)
                if not len(cur):
                    break
            assert list(zip(ids.tolist(), sizes.tolist())) == ref, ("gather", qkind, n, scale)


def sketch_trial(rng):
    "random records (invalid bases, lower case, short) x k in {21, 31, 51} or any k, scaled / num / abundance"
    recs = []
    for _ in range(int(rng.integers(1, 6))):
        n = int(rng.choice([0, 5, 20, 21, 52, 300, 3000]))
        g = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
        for _ in range(int(rng.integers(0, 4))):
            if n:
                g[int(rng.integers(0, n))] = rng.choice(np.frombuffer(b"NnRacgt\x00", dtype=np.uint8))
        recs.append(g.astype(np.uint8))
    seqs = np.concatenate(recs) if recs else np.zeros(0, np.uint8)
    offs = np.cumsum([0] + [len(g) for g in recs]).astype(np.uint64)
    ks = [21, 31, 51] if rng.random() < 0.6 else sorted(set(int(x) for x in rng.choice([4, 15, 21, 24, 31, 33, 51, 63], size=3)))
    mode = int(rng.integers(0, 3))
    kw = [dict(scaled=int(rng.choice([1, 7, 50]))), dict(num=int(rng.choice([5, 60]))), dict(scaled=9, track_abundance=True)][mode]
    for fused in ("0", "1") if SWITCHED else ("0",):
        os.environ["SMB_SKETCH_FUSED"] = fused
        sset, _ = B.sketch_sequences(seqs, offs, ks, **kw)
        hh, oo, ab = sset.to_host(with_abunds=True) if mode == 2 else (*sset.to_host(), None)
        for gi, g in enumerate(recs):
            for ki, k in enumerate(ks):
                om = orc.OracleMinHash(scaled=kw.get("scaled", 0), num=kw.get("num", 0), ksize=k, track_abundance=mode == 2)
                om.add_sequence(bytes(g), force=True)
                lo, hi = int(oo[gi * len(ks) + ki]), int(oo[gi * len(ks) + ki + 1])
                assert hh[lo:hi].tolist() == om.mins().tolist(), ("sketch", fused, ks, kw, gi, k)
                if mode == 2:
                    assert ab[lo:hi].tolist() == om.abunds().tolist(), ("abund", fused, ks, gi, k)
    os.environ.pop("SMB_SKETCH_FUSED")


def protein_trial(rng):
    "protein / dayhoff / hp sketches from residues or from DNA translated in six frames, any k, scaled / num"
    moltype = str(rng.choice(["protein", "dayhoff", "hp"]))
    translate = bool(rng.random() < 0.5)
    recs = []
    for _ in range(int(rng.integers(1, 5))):
        n = int(rng.choice([0, 2, 20, 21, 64, 400, 2000]))
        if translate:
            g = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
            junk = b"NnRacgt"
        else:
            g = rng.choice(np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8), size=n)
            junk = b"XxBZ*acd"
        for _ in range(int(rng.integers(0, 4))):
            if n:
                g[int(rng.integers(0, n))] = rng.choice(np.frombuffer(junk, dtype=np.uint8))
        recs.append(g.astype(np.uint8))
    seqs = np.concatenate(recs) if recs else np.zeros(0, np.uint8)
    offs = np.cumsum([0] + [len(g) for g in recs]).astype(np.uint64)
    ks = sorted(set(int(x) for x in rng.choice([1, 2, 7, 10, 16, 21, 33, 42], size=2)))
    scaled = int(rng.choice([1, 5, 40]))
    sset, _ = B.sketch_sequences(seqs, offs, ks, scaled=scaled, moltype=moltype, input_is_protein=not translate)
    rows = sset.rows()
    mx = orc.max_hash_for_scaled(scaled)
    for gi, g in enumerate(recs):
        for ki, k in enumerate(ks):
            fn = orc.seq_to_hashes_translate if translate else orc.seq_to_hashes_protein
            hs = np.asarray(fn(bytes(g), k, moltype, keep_zeros=False), dtype=np.uint64)
            want = np.unique(hs[hs <= np.uint64(mx)])
            assert np.array_equal(rows[gi * len(ks) + ki], want), ("protein", moltype, translate, k, scaled, len(g))


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    what = sys.argv[3] if len(sys.argv) > 3 else "sets"
    rng = np.random.Generator(np.random.PCG64(seed))
    t0, k = time.time(), 0
    while time.time() - t0 < seconds:
        {"sketch": sketch_trial, "protein": protein_trial}.get(what, trial)(rng)
        k += 1
    print("%d trials in %.0f s, no difference" % (k, time.time() - t0))


if __name__ == "__main__":
    main()
