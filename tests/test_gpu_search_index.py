"""GPU parity tests of the one-vs-many paths beyond the shared-memory table: the streaming pass over the range-major copy
of a resident set (queries too large for shared memory), the global-directory kernel behind SMB_SEARCH_LAYOUT=global, the
inverted index of a resident set, and the one-pass k = 21, 31, 51 hash kernel against the three launches.  All against
the oracle; the same kernels run on the CPU in tests/test_host_emulation.py."""
import os

import numpy as np
import pytest

import oracle as orc
from sourmash_b200.synth import MAX_HASH_1000, rows_of, synth_sketches

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    from sourmash_b200 import batch
    return batch


class _DeviceMatrix:
    """float64 device buffer for the *_device entry points: a torch CUDA tensor on a GPU box; plain host memory
    when the suite runs against the emulated library (tests/host_emul/run_gpu_tests_emulated.py), whose
    "device" pointers are host pointers."""

    def __init__(self, shape):
        import torch
        if torch.cuda.is_available():
            self._t = torch.empty(shape, dtype=torch.float64, device="cuda")
            self.ptr = self._t.data_ptr()
        else:
            self._t = None
            self._a = np.full(shape, -1.0)
            self.ptr = self._a.ctypes.data

    def numpy(self):
        if self._t is None:
            return self._a
        import torch
        torch.cuda.synchronize()
        return self._t.cpu().numpy()


def _edge_rows():
    rng = np.random.Generator(np.random.PCG64(77))
    big = np.uint64(2**64 - 1)
    rows = [np.zeros(0, np.uint64), np.array([0], np.uint64), np.array([0, 1, 2, big - 1, big], np.uint64),
            np.array([big], np.uint64), np.arange(1, 300, dtype=np.uint64), np.arange(1, 300, dtype=np.uint64),
            np.unique(rng.integers(0, 2**64 - 1, size=500, dtype=np.uint64))]
    rows += [np.array([7, 1000 + i], dtype=np.uint64) for i in range(100)]      # one hash shared by 100 rows
    return rows


def _big_query(rows, seed=4000, extra=400_000):
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.unique(np.concatenate([rng.integers(1, MAX_HASH_1000, size=extra, dtype=np.uint64)] + rows[:40]))


@pytest.mark.parametrize("layout", [None, "global"])
def test_large_query_search_matches_oracle(B, monkeypatch, layout):
    "query too large for shared memory: the streaming pass over the range-major copy (default) / the global-directory kernel"
    if layout:
        monkeypatch.setenv("SMB_SEARCH_LAYOUT", layout)
    h, off = synth_sketches(3000, mean=2000, sd=200, lo=0, hi=3000, n_families=5, pool=2500, seed=3)
    rows = rows_of(h, off)
    db = B.SketchSet.from_host(h, off)
    q = _big_query(rows)                                                      # too large for shared memory: global path
    want = orc.one_vs_many(q, h, off).astype(np.uint32)
    assert np.array_equal(B.one_vs_many(q, db), want)
    assert np.array_equal(B.one_vs_many(q, db), want)                        # second call: cached bounds
    q2 = np.unique(np.concatenate([q[::3], np.array([MAX_HASH_1000 + 5, 2**63, 2**64 - 1], dtype=np.uint64)]))
    assert np.array_equal(B.one_vs_many(q2, db), orc.one_vs_many(q2, h, off).astype(np.uint32))
    eh, eoff = orc.to_csr(_edge_rows())
    edb = B.SketchSet.from_host(eh, eoff)
    rng = np.random.Generator(np.random.PCG64(5))
    qe = np.unique(np.concatenate([rng.integers(0, 2**64 - 1, size=300_000, dtype=np.uint64), eh]))
    assert np.array_equal(B.one_vs_many(qe, edb), orc.one_vs_many(qe, eh, eoff).astype(np.uint32))


def test_db_index_search_and_gather(B):
    from tests.test_gpu_kernels import _gather_oracle
    h, off = synth_sketches(3000, mean=2000, sd=200, lo=0, hi=3000, n_families=5, pool=2500, seed=3)
    rows = rows_of(h, off)
    db = B.SketchSet.from_host(h, off)
    queries = [rows[7], _big_query(rows), np.zeros(0, np.uint64), np.array([1, 2**64 - 1], dtype=np.uint64)]
    plain = [B.one_vs_many(q, db) for q in queries]
    g_query = np.unique(np.concatenate([rows[4], rows[9][:1500], rows[25][500:2500], rows[30][::2]]))
    ids0, sizes0 = B.gather(g_query, db, threshold=5)
    n_keys = db.build_index()
    assert db.has_index and n_keys == len(np.unique(h)) and db.build_index() == n_keys
    for q, p in zip(queries, plain):
        got = B.one_vs_many(q, db)
        assert np.array_equal(got, p)
        if len(q):
            assert np.array_equal(got, orc.one_vs_many(q, h, off).astype(np.uint32))
    ids1, sizes1 = B.gather(g_query, db, threshold=5)
    assert ids1.tolist() == ids0.tolist() and sizes1.tolist() == sizes0.tolist()
    assert list(zip(ids1.tolist(), sizes1.tolist())) == _gather_oracle(g_query, rows, threshold=5)
    sess = B.GatherSession(g_query, db, min_count=5)                          # step API on the indexed set
    picked = []
    while True:
        cnt, row = sess.peek()
        if cnt < 5:
            break
        isect = sess.intersect(row)
        picked.append((row, len(isect)))
        if sess.apply(isect) == 0:
            break
    assert picked == list(zip(ids0.tolist(), sizes0.tolist()))
    db.drop_index()
    assert not db.has_index and np.array_equal(B.one_vs_many(queries[1], db), plain[1])
    eh, eoff = orc.to_csr(_edge_rows())
    edb = B.SketchSet.from_host(eh, eoff)
    edb.build_index()
    for q in (np.array([7], dtype=np.uint64), np.unique(eh), np.array([0, 2**64 - 1], dtype=np.uint64)):
        assert np.array_equal(B.one_vs_many(q, edb), orc.one_vs_many(q, eh, eoff).astype(np.uint32))


def test_db_index_behind_linear_index(B):
    "LinearIndex.build_device_index: search / prefetch / gather results unchanged."
    import glob
    import sourmash_b200 as smb
    from sourmash_b200.index import LinearIndex, gather
    from tests.conftest import GOLDEN
    d = os.path.join(GOLDEN, "gather")
    query = smb.signature.load_one_signature_from_json(os.path.join(d, "combined.sig"), ksize=21)
    sigs = [smb.signature.load_one_signature_from_json(p, ksize=21) for p in sorted(glob.glob(os.path.join(d, "GCF*.sig")))]
    plain, indexed = LinearIndex(sigs), LinearIndex(sigs)
    assert indexed.build_device_index() == len(set(h for s in sigs for h in s.minhash.hashes))

    def view(results):
        return [(r.score, r.signature.md5sum()) for r in results]
    assert view(indexed.search(query, threshold=0.0, do_containment=True)) == view(plain.search(query, threshold=0.0, do_containment=True))
    assert view(indexed.prefetch(query, 50000)) == view(plain.prefetch(query, 50000))
    assert [(g.match.md5sum(), g.intersect_size) for g in gather(query, indexed)] == \
        [(g.match.md5sum(), g.intersect_size) for g in gather(query, plain)]


def test_fused_sketch_kernel_matches_default_and_oracle(B, monkeypatch):
    "k = 21, 31, 51 in one pass (the default) == three launches (SMB_SKETCH_FUSED=0) == the oracle (scaled, num, abundance)."
    from sourmash_b200.synth import synth_genome
    genomes = [synth_genome(60_000 + 1000 * i, seed=10 + i, n_every=97 if i == 1 else 0) for i in range(4)]
    genomes.append(synth_genome(40, seed=3))                               # shorter than 51: only k = 21 and 31
    genomes.append(synth_genome(20, seed=4))                               # shorter than every k
    genomes[2][5000:5400] = np.frombuffer(bytes(genomes[2][5000:5400]).lower(), dtype=np.uint8)
    genomes[3][777] = ord("R")
    seqs = np.concatenate(genomes)
    offs = np.cumsum([0] + [len(g) for g in genomes]).astype(np.uint64)
    for ks in ([21, 31, 51], [51, 21, 31]):
        for kw in (dict(scaled=100), dict(scaled=1), dict(num=500), dict(scaled=50, track_abundance=True)):
            monkeypatch.setenv("SMB_SKETCH_FUSED", "0")
            plain, nk0 = B.sketch_sequences(seqs, offs, ks, **kw)
            monkeypatch.delenv("SMB_SKETCH_FUSED", raising=False)
            fused, nk1 = B.sketch_sequences(seqs, offs, ks, **kw)
            assert nk0 == nk1
            a, b = plain.rows(), fused.rows()
            assert len(a) == len(b) == len(genomes) * 3
            for x, y in zip(a, b):
                assert np.array_equal(x, y), (ks, kw)
            if kw.get("track_abundance"):
                assert np.array_equal(plain.to_host(with_abunds=True)[2], fused.to_host(with_abunds=True)[2])
    monkeypatch.delenv("SMB_SKETCH_FUSED", raising=False)
    sset, _ = B.sketch_sequences(seqs, offs, [21, 31, 51], scaled=100)
    rows = sset.rows()
    mx = orc.max_hash_for_scaled(100)
    for gi, g in enumerate(genomes):
        for ki, k in enumerate((21, 31, 51)):
            assert np.array_equal(rows[gi * 3 + ki], orc.sketch_scaled(g, k, mx)), (gi, k)
