"""End-to-end checks of the library through its C ABI and Python layer with the EMULATED build (emul_lib.py: the
product's host glue and kernels, the "device" being the SIMT emulator on the CPU).  Run as a script by
tests/test_emulated_library.py; every check compares with the oracle.  The point is the host glue that no other
CPU test reaches -- buffer sizes, launch geometry, cub call sequences, row chunking, caches on the resident set --
above all for the paths that sit behind switches and have not run on a GPU yet.

    python tests/host_emul/emulated_checks.py [name ...]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import emulated_boot  # noqa: E402

LIB = emulated_boot.install()

import numpy as np  # noqa: E402

import oracle as orc  # noqa: E402
from sourmash_b200 import batch as B  # noqa: E402
from sourmash_b200.synth import rows_of, synth_genome, synth_sketches  # noqa: E402

CHECKS = {}


def check(fn):
    CHECKS[fn.__name__] = fn
    return fn


class env:
    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        for k, v in self.kw.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _device_matrix(sset, n):
    out = np.full((n, n), -1.0)                      # "device" memory of the emulated build is host memory
    B.compare_jaccard_device(sset, out.ctypes.data)
    return out


@check
def compare_default_and_join():
    h, off = synth_sketches(150, mean=200, sd=40, lo=0, hi=400, n_families=4, pool=260, seed=2)
    want = orc.compare_all_pairs(h, off, nthreads=4)
    sset = B.SketchSet.from_host(h, off)
    for algo in ("tile", "join"):
        with env(SMB_COMPARE_ALGO=algo):
            assert np.array_equal(B.compare_jaccard(sset), want), algo
            assert np.array_equal(_device_matrix(sset, 150), want), algo
            assert B.last_compare_plan()["algo"] == algo
    with env(SMB_COMPARE_ALGO="join", SMB_JOIN_LAYOUT="plain"):         # the global-reduction join behind the stripe layout
        assert np.array_equal(B.compare_jaccard(sset), want)
        assert np.array_equal(_device_matrix(sset, 150), want)


@check
def compare_stripe_layouts_resident():
    h, off = synth_sketches(150, mean=200, sd=40, lo=0, hi=400, n_families=4, pool=260, seed=3)
    rows = rows_of(h, off)
    top = np.uint64(0x001234ab) << np.uint64(32)      # below max_hash(1000) ~ 2^54: the sort key is key >> 22
    for i in range(0, 150, 3):                        # hashes equal in their top 32 significant bits: the repair of the 32-bit sort
        extra = [top | np.uint64(v) for v in (9, 3, 7, 1) if (i + v) % 3]
        rows[i] = np.unique(np.concatenate([rows[i], np.array(extra, dtype=np.uint64)]))
    rows[7] = np.zeros(0, np.uint64)
    h, off = orc.to_csr(rows)
    want = orc.compare_all_pairs(h, off, nthreads=4)
    sset = B.SketchSet.from_host(h, off)
    for layout in (None, "stripe_full"):              # default: upper triangle + mirror; stripe_full: both directions
        for tags in (None, "u32"):
            with env(SMB_COMPARE_ALGO="join", SMB_JOIN_LAYOUT=layout, SMB_STRIPE_TAGS=tags):
                assert np.array_equal(B.compare_jaccard(sset), want), (layout, tags)
                assert np.array_equal(_device_matrix(sset, 150), want), (layout, tags)
                block = np.full((37, 150), -1.0)
                B.compare_jaccard_rows_device(sset, 50, 87, block.ctypes.data)
                assert np.array_equal(block, want[50:87]), (layout, tags)
    block = np.full((37, 150), -1.0)                  # the default path of the rows entry point
    B.compare_jaccard_rows_device(sset, 50, 87, block.ctypes.data)
    assert np.array_equal(block, want[50:87])


@check
def compare_host_path_row_chunks():
    "n >= 1024: smb_compare_jaccard takes the join, finalises / mirrors / downloads chunk of rows by chunk of rows"
    h, off = synth_sketches(1060, mean=40, sd=10, lo=0, hi=90, n_families=12, pool=60, seed=4)
    want = orc.compare_all_pairs(h, off, nthreads=8)
    sset = B.SketchSet.from_host(h, off)
    with env(SMB_COMPARE_ALGO="join"):
        assert np.array_equal(B.compare_jaccard(sset), want)
        for layout in ("stripe_full", "plain"):
            with env(SMB_JOIN_LAYOUT=layout):
                assert np.array_equal(B.compare_jaccard(sset), want), layout


@check
def compare_shards_sum_to_full():
    "the multi-GPU building blocks: partial counts of key-range / tile shards add up; rows finalised from the sum"
    n = 200
    h, off = synth_sketches(n, mean=150, sd=40, lo=30, hi=300, n_families=5, pool=200, seed=8)
    sset = B.SketchSet.from_host(h, off)
    want = orc.pairwise_common(h, off, nthreads=4)
    jac = orc.compare_all_pairs(h, off, nthreads=4)
    iu = np.triu_indices(n, 1)
    for algo in ("join", "tile"):
        with env(SMB_COMPARE_ALGO=algo):
            for shards in (1, 3):
                total = np.zeros((n, n), dtype=np.uint32)
                for r in range(shards):
                    part = np.zeros((n, n), dtype=np.uint32)
                    B.pairwise_counts_shard_device(sset, r, shards, part.ctypes.data)
                    total += part
                assert np.array_equal(total[iu], want[iu]), (algo, shards)
            rows = np.full((60, n), -1.0)
            B.finalize_jaccard_rows_device(sset, total.ctypes.data, 70, 130, rows.ctypes.data)
            assert np.array_equal(rows, jac[70:130]), algo
    # whole-row counters by shard (the unit of a reduce-scatter by row blocks): stripe layout = key-range shards of the
    # sorted stream; plain join / tile kernel = upper-triangle shards, mirrored
    off_diag = ~np.eye(n, dtype=bool)
    for envs in (dict(SMB_COMPARE_ALGO="join"), dict(SMB_COMPARE_ALGO="join", SMB_JOIN_LAYOUT="plain"),
                 dict(SMB_COMPARE_ALGO="tile"), dict(SMB_COMPARE_ALGO="join", SMB_STRIPE_TAGS="u32")):
        with env(**envs):
            for bits, dt in ((32, np.uint32), (16, np.uint16)):
                for shards in (1, 3, 8):
                    total = np.zeros((n, n), dtype=dt)
                    for r in range(shards):
                        part = np.full((n, n), 0xbeef, dtype=dt)                       # every cell must be written
                        B.compare_counts_shard_device(sset, r, shards, part.ctypes.data, bits=bits)
                        np.fill_diagonal(part, 0)
                        total += part
                    assert np.array_equal(total[off_diag], want[off_diag].astype(dt)), (envs, bits, shards)
                    if bits == 16:                        # what the reduce-scatter does: pairs of counters added as one u32
                        assert np.array_equal((total.view(np.uint32) + 0).view(np.uint16), total)
                rows = np.full((60, n), -1.0)
                block = np.ascontiguousarray(total[70:130])
                B.finalize_counts_rows_device(sset, block.ctypes.data, 70, 130, rows.ctypes.data, bits=bits)
                assert np.array_equal(rows, jac[70:130]), (envs, bits)
    # rows too large for the shared-memory tables (warp-per-pair kernel): the shards must still split
    # the pairs, not each count all of them (round-1 advisor finding: counts came out x world_size)
    rng = np.random.Generator(np.random.PCG64(77))
    pool = np.unique(rng.integers(1, 1 << 62, size=60_000, dtype=np.uint64))
    big = [np.sort(rng.choice(pool, size=40_000, replace=False)) for _ in range(3)]
    bh = np.concatenate(big)
    boff = np.array([0, 40_000, 80_000, 120_000], dtype=np.uint64)
    bset = B.SketchSet.from_host(bh, boff)
    bwant = orc.pairwise_common(bh, boff, nthreads=4)
    biu = np.triu_indices(3, 1)
    for shards in (1, 2):
        total = np.zeros((3, 3), dtype=np.uint32)
        for r in range(shards):
            part = np.zeros((3, 3), dtype=np.uint32)
            B.pairwise_counts_shard_device(bset, r, shards, part.ctypes.data)
            total += part
        assert np.array_equal(total[biu], bwant[biu]), ("large rows", shards, total, bwant)


@check
def search_layouts_and_index():
    h, off = synth_sketches(400, mean=300, sd=60, lo=0, hi=600, n_families=4, pool=400, seed=5)
    rows = rows_of(h, off)
    rng = np.random.Generator(np.random.PCG64(6))
    q_small = rows[7]
    q_large = np.unique(np.concatenate([rng.integers(1, 2**54, size=60_000, dtype=np.uint64)] + rows[:30] +
                                       [np.array([2**63, 2**64 - 1], dtype=np.uint64)]))
    db = B.SketchSet.from_host(h, off)
    want_small = orc.one_vs_many(q_small, h, off).astype(np.uint32)
    want_large = orc.one_vs_many(q_large, h, off).astype(np.uint32)
    assert np.array_equal(B.one_vs_many(q_small, db), want_small)
    with env(SMB_RM_RANGES="7"):                                               # a large query streams the range-major copy of the set
        assert np.array_equal(B.one_vs_many(q_large, db), want_large)
        assert np.array_equal(B.one_vs_many(q_large[::2], db), orc.one_vs_many(q_large[::2], h, off).astype(np.uint32))   # cached layout
    with env(SMB_SEARCH_LAYOUT="global"):                                      # global directory + bitmap kernel
        assert np.array_equal(B.one_vs_many(q_large, db), want_large)
    n_keys = db.build_index()
    assert db.has_index and n_keys == len(np.unique(h)) and db.build_index() == n_keys
    assert np.array_equal(B.one_vs_many(q_small, db), want_small)
    assert np.array_equal(B.one_vs_many(q_large, db), want_large)
    assert np.array_equal(B.one_vs_many(np.zeros(0, np.uint64), db), np.zeros(400, np.uint32))
    db.drop_index()
    assert not db.has_index and np.array_equal(B.one_vs_many(q_large, db), want_large)


def _gather_oracle(query, rows, threshold):
    q = np.array(query, dtype=np.uint64)
    counts = np.array([orc.count_common(q, r) for r in rows], dtype=np.int64)
    out = []
    while True:
        j = int(np.argmax(counts))
        if counts[j] < threshold or counts[j] == 0:
            break
        isect = np.intersect1d(q, rows[j])
        out.append((j, len(isect)))
        counts = counts - np.array([orc.count_common(isect, r) for r in rows], dtype=np.int64)
        q = np.setdiff1d(q, isect)
        if not len(q):
            break
    return out


@check
def gather_default_and_index():
    h, off = synth_sketches(300, mean=250, sd=50, lo=0, hi=500, n_families=3, pool=320, seed=7)
    rows = rows_of(h, off)
    rows[17] = rows[4].copy()
    h, off = orc.to_csr(rows)
    query = np.unique(np.concatenate([rows[4], rows[9][:150], rows[25][50:250], rows[41][::2]]))
    want = _gather_oracle(query, rows, 5)
    assert len(want) >= 4
    db = B.SketchSet.from_host(h, off)
    ids, sizes = B.gather(query, db, threshold=5)
    assert list(zip(ids.tolist(), sizes.tolist())) == want
    db.build_index()
    ids, sizes = B.gather(query, db, threshold=5)
    assert list(zip(ids.tolist(), sizes.tolist())) == want
    sess = B.GatherSession(query, db, min_count=5)
    picked = []
    while True:
        cnt, row = sess.peek()
        if cnt < 5:
            break
        isect = sess.intersect(row)
        picked.append((row, len(isect)))
        if sess.apply(isect) == 0:
            break
    assert picked == want


@check
def sketch_default_and_fused():
    genomes = [synth_genome(6000 + 500 * i, seed=10 + i, n_every=97 if i == 1 else 0) for i in range(3)]
    genomes.append(synth_genome(40, seed=3))
    genomes.append(synth_genome(20, seed=4))
    genomes[2][500:580] = np.frombuffer(bytes(genomes[2][500:580]).lower(), dtype=np.uint8)
    genomes.append(np.concatenate([genomes[0][:3000]] * 2))                  # repeats: abundances above one
    seqs = np.concatenate(genomes)
    offs = np.cumsum([0] + [len(g) for g in genomes]).astype(np.uint64)
    mx = orc.max_hash_for_scaled(20)
    for fused in (None, "0"):                         # default: one pass over the bases; "0": one launch per ksize
        with env(SMB_SKETCH_FUSED=fused):
            for ks in ([21, 31, 51], [51, 21, 31]):
                sset, nk = B.sketch_sequences(seqs, offs, ks, scaled=20)
                rows = sset.rows()
                assert nk == sum(max(len(g) - k + 1, 0) for g in genomes for k in ks)
                for gi, g in enumerate(genomes):
                    for ki, k in enumerate(ks):
                        assert np.array_equal(rows[gi * 3 + ki], orc.sketch_scaled(g, k, mx)), (fused, gi, k)
            sset, _ = B.sketch_sequences(seqs, offs, [21, 31, 51], scaled=20, track_abundance=True)
            hh, oo, ab = sset.to_host(with_abunds=True)
            for gi, g in enumerate(genomes):
                om = orc.OracleMinHash(scaled=20, ksize=31, track_abundance=True)
                om.add_sequence(bytes(g), force=True)
                lo, hi = int(oo[gi * 3 + 1]), int(oo[gi * 3 + 2])
                assert hh[lo:hi].tolist() == om.mins().tolist() and ab[lo:hi].tolist() == om.abunds().tolist(), (fused, gi)
            nset, _ = B.sketch_sequences(seqs, offs, [21, 31, 51], num=50)
            for gi, g in enumerate(genomes):
                om = orc.OracleMinHash(num=50, ksize=21)
                om.add_sequence(bytes(g), force=True)
                assert nset.rows()[gi * 3].tolist() == om.mins().tolist(), (fused, gi)


def main(names):
    names = names or list(CHECKS)
    for name in names:
        t0 = time.time()
        CHECKS[name]()
        print("ok  %-36s %.1f s" % (name, time.time() - t0), flush=True)
    print("emulated checks passed (%s)" % os.path.basename(LIB))


if __name__ == "__main__":
    main(sys.argv[1:])
