"""BASELINE.json-sized workloads on the GPU, checked through size-independent properties
(the oracle is only run on samples that finish in seconds)."""
import numpy as np
import pytest

import oracle as orc
from sourmash_b200.synth import MAX_HASH_1000, rows_of, synth_genome, synth_sketches

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    from sourmash_b200 import batch
    assert batch.device_count() > 0
    return batch


@pytest.fixture(scope="module")
def config3():
    return synth_sketches(10_000)


def test_compare_10k_properties(B, config3):
    h, off = config3
    n = len(off) - 1
    sset = B.SketchSet.from_host(h, off)
    m = B.compare_jaccard(sset)
    assert m.shape == (n, n) and m.dtype == np.float64
    assert np.array_equal(m, m.T)                                   # symmetric (mirrored, not recomputed)
    assert np.all(np.diag(m) == 1.0)                                # compare.py:38 np.ones
    assert m.min() >= 0.0 and m.max() <= 1.0
    sizes = np.diff(off.astype(np.int64))
    # jaccard -> common is invertible: c = J (|A|+|B|) / (1+J); must be integers
    i = np.arange(0, n, 97)
    sub = m[np.ix_(i, i)]
    c = sub * (sizes[i][:, None] + sizes[i][None, :]) / (1.0 + sub)
    off_diag = ~np.eye(len(i), dtype=bool)
    assert np.allclose(c[off_diag], np.round(c[off_diag]), atol=1e-6)
    # families of 100: members i, i+100 share a pool -> high similarity; i, i+1 do not
    assert m[0, 100] > 0.2 and m[0, 1] < 0.01
    # oracle on a sample of rows (bit-exact)
    rows = rows_of(h, off)
    for a in (0, 1234, 9999):
        for b in (7, 100, 1334, 5000, 9998):
            if a != b:
                assert m[a, b] == orc.jaccard(rows[a], rows[b])
    # one-vs-many of a member against the whole set reproduces its matrix row (checksum of counts)
    for q in (3, 4321):
        cm = B.one_vs_many(rows[q], sset).astype(np.int64)
        un = sizes[q] + sizes - cm
        want = cm / np.maximum(un, 1)
        want[q] = 1.0
        assert np.array_equal(want, m[q])


def test_sketch_100_genomes_properties(B):
    n_g, L = 100, 5_000_000
    seqs = np.empty(n_g * L, dtype=np.uint8)
    for g in range(n_g):
        seqs[g * L:(g + 1) * L] = synth_genome(L, 1000 + g)
    offs = np.arange(n_g + 1, dtype=np.uint64) * np.uint64(L)
    ks = [21, 31, 51]
    sset, nk = B.sketch_sequences(seqs, offs, ks, scaled=1000)
    assert nk == sum(n_g * (L - k + 1) for k in ks)
    h, off = sset.to_host()
    assert len(off) == n_g * 3 + 1
    sizes = np.diff(off.astype(np.int64))
    assert abs(sizes.mean() - L / 1000) < 30 and sizes.min() > 4600 and sizes.max() < 5400
    for r in range(0, n_g * 3, 17):                                  # sorted-unique, below max_hash
        row = h[int(off[r]):int(off[r + 1])]
        assert np.all(row[1:] > row[:-1]) and row[-1] <= MAX_HASH_1000 and row[0] > 0
    for g, ki in ((0, 1), (57, 0), (99, 2)):                        # bit-exact vs oracle on a sample
        want = orc.sketch_scaled(seqs[g * L:(g + 1) * L], ks[ki], MAX_HASH_1000)
        r = g * 3 + ki
        assert np.array_equal(h[int(off[r]):int(off[r + 1])], want)
    # idempotence / independence of batching: genome 57 alone == its row in the batch
    alone, _ = B.sketch_sequences(seqs[57 * L:58 * L], [0, L], ks, scaled=1000)
    ah, aoff = alone.to_host()
    for ki in range(3):
        r = 57 * 3 + ki
        assert np.array_equal(ah[int(aoff[ki]):int(aoff[ki + 1])], h[int(off[r]):int(off[r + 1])])
    # splitting a genome into two records feeding one sketch loses exactly the k-1 junction windows
    two, _ = B.sketch_sequences(seqs[:L], [0, L // 2, L], [31], scaled=1000,
                                seq_to_sketch=np.zeros(2, np.uint32), n_sketches=1)
    (row2,) = two.rows()
    full = h[int(off[1]):int(off[2])]
    assert set(row2.tolist()) <= set(full.tolist()) and len(full) - len(row2) <= 1


def test_search_large_query_properties(B, config3):
    """configs[3] shape at 1/10 database scale: 1e7-hash query vs 30 000 sketches (containment)."""
    h, off = config3
    rng = np.random.Generator(np.random.PCG64(4000))
    rows = rows_of(h, off)
    db_rows = [rows[i % 10_000] for i in range(30_000)]
    planted = list(range(0, 30_000, 301))
    query = np.unique(np.concatenate([rng.integers(1, MAX_HASH_1000, size=10_000_000, dtype=np.uint64)] +
                                     [db_rows[j][: len(db_rows[j]) // 2] for j in planted]))
    db = B.SketchSet.from_rows(db_rows)
    cm = B.one_vs_many(query, db)
    sizes = db.sizes()
    assert cm.shape == (30_000,) and np.all(cm <= sizes)
    for j in planted[:20]:
        assert cm[j] >= len(db_rows[j]) // 2
    # identical subjects get identical counts; sample vs oracle
    assert np.array_equal(cm[:10_000], cm[10_000:20_000])
    for j in (0, 301, 12_345, 29_999):
        assert cm[j] == orc.count_common(query, db_rows[j])
    # containment scores / threshold filter as Index.find would compute them (search.py:143-160)
    cont = cm / len(query)
    assert cont.max() <= 1.0 and np.count_nonzero(cm * 1000 >= 2_000_000) >= len(planted)


def test_gather_large_properties(B, config3):
    """configs[4] shape: ~1e5-hash metagenome query vs 50 000 sketches with planted, overlapping matches."""
    h, off = config3
    rows = rows_of(h, off)
    db_rows = [rows[i % 10_000] for i in range(50_000)]
    db = B.SketchSet.from_rows(db_rows)
    rng = np.random.Generator(np.random.PCG64(5000))
    planted = [int(x) for x in rng.choice(10_000, size=40, replace=False)]
    parts = [db_rows[j][rng.random(len(db_rows[j])) < 0.8] for j in planted]
    query = np.unique(np.concatenate(parts + [rng.integers(1, MAX_HASH_1000, size=20_000, dtype=np.uint64)]))
    ids, sizes = B.gather(query, db, threshold=50)
    assert len(ids) >= 40 and len(set(ids.tolist())) == len(ids)
    assert np.all(sizes[1:] <= sizes[:-1]) and sizes[-1] >= 50          # greedy: non-increasing gains
    # the gains partition the covered part of the query
    covered = np.zeros(0, dtype=np.uint64)
    remaining = query
    for j, s in zip(ids.tolist(), sizes.tolist()):
        isect = np.intersect1d(remaining, db_rows[j])
        assert len(isect) == s
        remaining = np.setdiff1d(remaining, db_rows[j])
    assert int(sizes.sum()) == len(query) - len(remaining)
    # first pick is the best single overlap (lowest index among ties)
    cm = B.one_vs_many(query, db)
    assert ids[0] == int(np.argmax(cm)) and sizes[0] == cm.max()


def test_compare_10k_first_256_rows_equal_the_oracle(B, config3):
    """configs[2]: 256 complete rows of the 10 000 x 10 000 matrix (2.5 million pairs) against the CPU oracle's
    compare_serial restatement, float64 bit for bit -- resident entry point and host entry point."""
    import os
    import torch
    h, off = config3
    n = len(off) - 1
    want = orc.compare_all_pairs(h, off, first_row=0, n_rows=256, nthreads=os.cpu_count() or 8)[:256]
    sset = B.SketchSet.from_host(h, off)
    d_out = torch.empty((n, n), dtype=torch.float64, device="cuda")
    B.compare_jaccard_device(sset, d_out.data_ptr())
    torch.cuda.synchronize()
    got = d_out[:256].cpu().numpy()
    iu = np.triu_indices(256, 1, n)
    assert np.array_equal(got[iu], want[iu])
    full = d_out.cpu().numpy()
    assert np.array_equal(full, full.T) and np.array_equal(full[:, :256].T[iu], want[iu])     # mirrored cells too
    assert np.array_equal(B.compare_jaccard(sset), full)                                          # host path: identical matrix


@pytest.mark.timeout(1500)
def test_search_300k_subjects_all_counts_equal_the_oracle(B):
    """configs[3] at FULL size: a 1e7-hash query against 300 000 resident sketches (12 GB), every one of the
    300 000 counters compared with the CPU oracle (binary-search form of count_common, held equal to the faithful walk
    in tests/test_oracle_golden.py; the faithful walk itself checks the first 2 000 subjects)."""
    import os
    import torch
    import bench
    from sourmash_b200.synth import database_plan, search_query
    query = search_query(bench.N_QUERY_SEARCH)
    sizes, frac = database_plan(bench.N_DB_SEARCH, 4001, planted_frac=0.01)
    db, d_h, h_off = bench.build_database(torch, B, sizes, frac, query, 4002, 0, bench.N_DB_SEARCH)
    got = B.one_vs_many(query, db)
    hh = d_h.cpu().numpy().view(np.uint64)
    ncores = os.cpu_count() or 8
    want = orc.one_vs_many_bsearch(query, hh, h_off, nthreads=ncores)
    assert np.array_equal(got.astype(np.uint64), want)
    planted = np.nonzero(frac > 0)[0]
    assert 2000 < len(planted) < 4000 and np.all(got[planted] >= (frac[planted] * sizes[planted]).astype(np.int64))
    assert got[np.setdiff1d(np.arange(len(got)), planted)].max() <= 3          # random subjects share next to nothing
    assert np.array_equal(orc.one_vs_many(query, hh[: int(h_off[2000])], h_off[:2001], nthreads=ncores), want[:2000])
    # the device entry point and the global-directory kernel give the same counters
    d_q = torch.from_numpy(query.view(np.int64)).cuda()
    d_c = torch.zeros(len(got), dtype=torch.int32, device="cuda")
    B.one_vs_many_device(d_q.data_ptr(), len(query), db, d_c.data_ptr())
    assert np.array_equal(d_c.cpu().numpy().astype(np.uint32), got)
    os.environ["SMB_SEARCH_LAYOUT"] = "global"
    try:
        assert np.array_equal(B.one_vs_many(query, db), got)
    finally:
        del os.environ["SMB_SEARCH_LAYOUT"]


def test_gather_50k_picks_equal_the_cpu_rounds(B):
    "configs[4] (SURVEY 8d): ~1e5-hash metagenome vs 50 000 sketches, 200 planted in overlapping clusters; the pick list = the CPU rounds"
    import torch
    import bench
    from sourmash_b200.synth import gather_workload
    query, sizes, frac, overrides = gather_workload(bench.N_DB_GATHER)
    db, d_h, h_off = bench.build_database(torch, B, sizes, frac, query, 5002, 0, bench.N_DB_GATHER, overrides)
    ids, isizes = B.gather(query, db, threshold=50)
    want = bench._gather_on_host(query, overrides, 50)
    assert list(zip(ids.tolist(), isizes.tolist())) == want and len(want) > 80
    cm = B.one_vs_many(query, db)
    others = np.setdiff1d(np.arange(len(cm)), np.array(sorted(overrides)))
    assert cm[others].max() < 50                                               # only planted rows can reach the threshold
