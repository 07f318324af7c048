"""compare.py host glue without a GPU: the sketches of N objects are pulled out of the library in
one call (SignatureSet.from_objects), validated with the reference's per-pair error semantics,
downsampled by prefix cut and handed to the batched kernels.  The kernels are replaced by the
oracle here (like tests/test_distributed_gloo.py does for the collectives), so what is checked is
everything around them: collection, order of the compatibility errors, mixed scaled values,
abundance arrays, containment / ANI post-processing."""
import numpy as np
import pytest

import oracle as orc
import sourmash_b200 as smb
from sourmash_b200 import batch as B
from sourmash_b200 import compare as C
from sourmash_b200 import distance_utils as DU


class _FakeSet:
    def __init__(self, h, off, ab=None):
        self.h, self.off, self.ab = np.array(h, dtype=np.uint64), np.array(off, dtype=np.uint64), ab

    def __len__(self):
        return len(self.off) - 1


@pytest.fixture
def cpu_kernels(monkeypatch):
    monkeypatch.setattr(B.SketchSet, "from_host", classmethod(lambda cls, h, off, ab=None: _FakeSet(h, off, ab)))
    monkeypatch.setattr(B, "compare_jaccard", lambda s, num=0, out=None: orc.compare_all_pairs(s.h, s.off, num=num))
    monkeypatch.setattr(B, "pairwise_common", lambda a, b=None, num=0, want_usize=False: orc.pairwise_common(a.h, a.off))

    def angular(s):
        n = len(s)
        m = np.ones((n, n))
        for i in range(n):
            for j in range(i + 1, n):
                a, b = slice(int(s.off[i]), int(s.off[i + 1])), slice(int(s.off[j]), int(s.off[j + 1]))
                m[i, j] = m[j, i] = orc.angular_similarity(s.h[a], s.ab[a], s.h[b], s.ab[b])
        return m
    monkeypatch.setattr(B, "compare_angular", angular)


def _sigs(n=9, scaleds=(1000,), track=(False,), seed=0, ksize=31):
    rng = np.random.default_rng(seed)
    pool = np.unique(rng.integers(1, 2**64 // 400, size=1500, dtype=np.uint64))
    out = []
    for i in range(n):
        mh = smb.MinHash(0, ksize, scaled=scaleds[i % len(scaleds)], track_abundance=track[i % len(track)])
        vals = rng.choice(pool, size=int(rng.integers(200, 900)), replace=False)
        if mh.track_abundance:
            mh.set_abundances({int(v): int(rng.integers(1, 9)) for v in vals})
        else:
            mh.add_many(vals)
        out.append(smb.SourmashSignature(mh, name=f"s{i}"))
    return out


def _rows(sigs, scaled):
    mx = B.max_hash_for_scaled(scaled)
    return [np.array(sorted(h for h in s.minhash.hashes if h <= mx), dtype=np.uint64) for s in sigs]


def test_collect_matches_per_object_path():
    sigs = _sigs(scaleds=(200, 1000, 500), track=(True, False))
    for objs in (sigs, [s.minhash for s in sigs]):
        a = C._collect(objs, downsample=True, with_abunds=True)
        b = C._collect_per_object(list(objs), downsample=True, need_scaled=False, with_abunds=True)
        for k in ("hashes", "offsets", "abunds", "sizes", "has_abund", "orig_sizes", "orig_scaled"):
            assert np.array_equal(a[k], b[k]), k
        assert (a["num"], a["scaled"], a["ksize"]) == (b["num"], b["scaled"], b["ksize"]) == (0, 1000, 31)
        rows = _rows(sigs, 1000)
        assert [a["hashes"][int(a["offsets"][i]):int(a["offsets"][i + 1])].tolist() for i in range(len(sigs))] == \
            [r.tolist() for r in rows]
    mixed = [sigs[0], sigs[1].minhash]                                   # falls back, same contract
    assert C._collect(mixed, downsample=True)["scaled"] == 1000


def test_error_semantics_and_order():
    sigs = _sigs(4)
    other_k = smb.SourmashSignature(smb.MinHash(0, 21, scaled=1000), name="k21")
    prot = smb.SourmashSignature(smb.MinHash(0, 31, scaled=1000, is_protein=True), name="prot")
    seed = smb.SourmashSignature(smb.MinHash(0, 31, scaled=1000, seed=43), name="seed")
    num = smb.SourmashSignature(smb.MinHash(500, 31), name="num")
    for bad, exc, msg in ((other_k, ValueError, "different ksizes"), (seed, ValueError, "mismatch in seed"),
                          (num, ValueError, "mismatch in scaled")):         # scaled vs num: max_hash differs
        with pytest.raises(exc, match=msg):
            C._collect(sigs + [bad], downsample=False)
    prot93 = smb.SourmashSignature(smb.MinHash(0, 31, scaled=1000, is_protein=True), name="p")
    with pytest.raises(ValueError, match="different ksizes"):            # protein k=31 is stored as 93
        C._collect(sigs + [prot93], downsample=False)
    dna93 = smb.SourmashSignature(smb.MinHash(0, 93, scaled=1000), name="d93")
    with pytest.raises(ValueError, match="DNA/prot"):
        C._collect([dna93, prot], downsample=False)
    with pytest.raises(ValueError, match="different ksizes"):            # the first offender decides
        C._collect(sigs + [other_k, seed], downsample=False)
    with pytest.raises(ValueError, match="mismatch in scaled"):
        C._collect(_sigs(4, scaleds=(100, 1000)), downsample=False)
    with pytest.raises(TypeError, match="can only calculate containment"):
        C._collect([num, num], downsample=False, need_scaled=True)
    assert C._collect([num, num], downsample=False)["num"] == 500


def test_compare_all_pairs_glue(cpu_kernels):
    sigs = _sigs(scaleds=(200, 1000), track=(False,))
    with pytest.raises(ValueError, match="mismatch in scaled"):
        C.compare_all_pairs(sigs, True)
    # different scaled values: the reference downsamples PER PAIR to max(scaled_i, scaled_j)
    # (similarity(other, downsample=True), minhash.rs:682-702), not everything to the coarsest of the list
    got = C.compare_all_pairs(sigs, True, downsample=True)
    rows_at = {sc: _rows(sigs, sc) for sc in (200, 1000)}
    for i in range(len(sigs)):
        for j in range(len(sigs)):
            sc = max(sigs[i].minhash.scaled, sigs[j].minhash.scaled)
            want = 1.0 if i == j else orc.jaccard(rows_at[sc][i], rows_at[sc][j])
            assert got[i, j] == want, (i, j, sc)
    assert any(got[i, j] != orc.jaccard(rows_at[1000][i], rows_at[1000][j])            # ... and that is observable
               for i in range(len(sigs)) for j in range(i) if max(sigs[i].minhash.scaled, sigs[j].minhash.scaled) == 200)
    # ANI of such a list: every pair as jaccard_ani(other, downsample=True) sees it -- both sketches cut at the pair's max
    # scaled, size_is_accurate() of the sketches as given (minhash.py:749-785) -- i.e. the one-scaled path on that pair
    ani = C.compare_all_pairs(sigs, True, downsample=True, return_ani=True)
    acc = C._sizes_accurate_arrays(np.array([len(s.minhash) for s in sigs]), np.array([s.minhash.scaled for s in sigs]))
    seen = set()
    for i in range(len(sigs)):
        for j in range(i + 1, len(sigs)):
            sc = max(sigs[i].minhash.scaled, sigs[j].minhash.scaled)
            pair = [smb.SourmashSignature(s.minhash.downsample(scaled=sc), name=s.name) for s in (sigs[i], sigs[j])]
            c2 = C._collect(pair, downsample=False)
            want, _u, _f = C.DU.jaccard_to_ani_matrix(orc.compare_all_pairs(c2["hashes"], c2["offsets"]), c2["sizes"], 31, sc,
                                                      size_accurate=acc[[i, j]])
            assert ani[i, j] == ani[j, i] == want[0, 1], (i, j, sc)
            seen.add((sc, bool(want[0, 1] > 0)))
    assert {sc for sc, _ in seen} == {200, 1000} and any(pos for _, pos in seen)
    # one scaled value (what `sourmash compare` hands over after its own downsampling): one batched call
    same = [smb.SourmashSignature(s.minhash.downsample(scaled=1000), name=s.name) for s in sigs]
    h, off = orc.to_csr(_rows(sigs, 1000))
    assert np.array_equal(C.compare_all_pairs(same, True), orc.compare_all_pairs(h, off))
    # abundance sketches with different scaled values: angular where both track abundance, per pair too
    msigs = _sigs(6, scaleds=(200, 1000, 500), track=(True, True, False), seed=11)
    gotm = C.compare_all_pairs(msigs, False, downsample=True)
    for i in range(6):
        for j in range(6):
            if i == j:
                continue
            sc = max(msigs[i].minhash.scaled, msigs[j].minhash.scaled)
            a, b = msigs[i].minhash.downsample(scaled=sc), msigs[j].minhash.downsample(scaled=sc)
            ra, rb = np.array(sorted(a.hashes), dtype=np.uint64), np.array(sorted(b.hashes), dtype=np.uint64)
            if a.track_abundance and b.track_abundance:
                ha, hb = a.hashes, b.hashes
                want = orc.angular_similarity(ra, np.array([ha[int(x)] for x in ra], dtype=np.uint64),
                                              rb, np.array([hb[int(x)] for x in rb], dtype=np.uint64))
            else:
                want = orc.jaccard(ra, rb)
            assert gotm[i, j] == want, (i, j)
    assert C.compare_all_pairs([], True).shape == (0, 0)
    # abundance: angular where both track abundance, Jaccard elsewhere
    sigs = _sigs(8, track=(True, True, False), seed=5)
    got = C.compare_all_pairs(sigs, False)
    rows = _rows(sigs, 1000)
    for i in range(8):
        for j in range(8):
            if i == j:
                assert got[i, j] == 1.0
                continue
            a, b = sigs[i].minhash, sigs[j].minhash
            if a.track_abundance and b.track_abundance:
                ha, hb = a.hashes, b.hashes
                want = orc.angular_similarity(rows[i], np.array([ha[int(x)] for x in rows[i]], dtype=np.uint64),
                                              rows[j], np.array([hb[int(x)] for x in rows[j]], dtype=np.uint64))
            else:
                want = orc.jaccard(rows[i], rows[j])
            assert got[i, j] == want, (i, j)
    assert np.array_equal(C.compare_all_pairs(sigs, True), orc.compare_all_pairs(*orc.to_csr(rows)))


def test_containment_and_ani_glue(cpu_kernels):
    sigs = _sigs(7, scaleds=(500, 1000), seed=2)
    tiny = smb.MinHash(0, 31, scaled=1000)
    tiny.add_many(list(sigs[0].minhash.downsample(scaled=1000).hashes)[:3])
    sigs.append(smb.SourmashSignature(tiny, name="tiny"))
    n = len(sigs)

    def cont(c, size, scaled=1000):
        if size == 0:
            return 0.0
        v = c / (size * (1.0 - (1.0 - 1.0 / scaled) ** float(size * scaled)))
        return 1.0 if v >= 1 else 0.0 if v <= 0 else v
    # different scaled values, downsample=True: count_common per pair at max(scaled_i, scaled_j); the
    # denominators keep len(self) and self.scaled of the sketches as given (minhash.py:819-905)
    mixed = {name: fn(sigs, downsample=True) for name, fn in (("c", C.compare_serial_containment),
                                                             ("mc", C.compare_serial_max_containment),
                                                             ("ac", C.compare_serial_avg_containment))}
    rows_at = {sc: _rows(sigs, sc) for sc in (500, 1000)}
    size = [len(s.minhash) for s in sigs]
    sc_of = [s.minhash.scaled for s in sigs]
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            sc = max(sc_of[i], sc_of[j])
            c = orc.count_common(rows_at[sc][i], rows_at[sc][j])
            assert mixed["c"][i, j] == cont(c, size[j], sc_of[j]), (i, j)    # siglist[j].contained_by(siglist[i], True)
            hi = max(i, j)                                                   # siglist[hi].max_containment(siglist[lo], True)
            assert mixed["mc"][i, j] == cont(c, min(size[i], size[j]), sc_of[hi]), (i, j)
            assert mixed["ac"][i, j] == (cont(c, size[j], sc_of[j]) + cont(c, size[i], sc_of[i])) / 2
    # the ANI forms take everything from the pair downsampled to its max scaled (minhash.py:843-945): the one-scaled
    # functions on that pair, except that containment_ani / max_containment_ani ask size_is_accurate() of the sketches as
    # given and the avg form (FracMinHashComparison) of the downsampled ones
    acc = C._sizes_accurate_arrays(np.array(size), np.array(sc_of))
    got_ani = {name: fn(sigs, downsample=True, return_ani=True) for name, fn in (("c", C.compare_serial_containment),
                                                                                  ("mc", C.compare_serial_max_containment),
                                                                                  ("ac", C.compare_serial_avg_containment))}
    for i in range(n):
        for j in range(i + 1, n):
            sc = max(sc_of[i], sc_of[j])
            pair = [smb.SourmashSignature(s.minhash.downsample(scaled=sc), name=s.name) for s in (sigs[i], sigs[j])]
            sizes2 = np.array([len(p.minhash) for p in pair], dtype=np.int64)
            common2 = np.array([[sizes2[0], 0], [0, sizes2[1]]], dtype=np.float64)
            common2[0, 1] = common2[1, 0] = orc.count_common(rows_at[sc][i], rows_at[sc][j])
            want_c, _ = C._containment_block(common2, sizes2, sc, 31, acc[[i, j]], True)
            want_mc, _ = C._max_containment_block(common2, sizes2, sc, 31, acc[[i, j]], True)
            assert (got_ani["c"][i, j], got_ani["c"][j, i]) == (want_c[0, 1], want_c[1, 0]), (i, j)
            assert got_ani["mc"][i, j] == got_ani["mc"][j, i] == want_mc[0, 1], (i, j)
            assert got_ani["ac"][i, j] == got_ani["ac"][j, i] == C.compare_serial_avg_containment(pair, return_ani=True)[0, 1]
    assert (got_ani["c"] > 0).sum() > n and np.array_equal(np.diagonal(got_ani["c"]), np.ones(n))
    # from here on: one scaled value, as `sourmash compare` hands the list over after its own downsampling
    given = sigs
    sigs = [smb.SourmashSignature(s.minhash.downsample(scaled=1000), name=s.name) for s in given]
    rows = _rows(sigs, 1000)
    m = C.compare_serial_containment(sigs)
    mm = C.compare_serial_max_containment(sigs)
    ma = C.compare_serial_avg_containment(sigs)
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            c = orc.count_common(rows[i], rows[j])
            assert m[i, j] == cont(c, len(rows[j])), (i, j)                 # siglist[j].contained_by(siglist[i])
            assert mm[i, j] == cont(c, min(len(rows[i]), len(rows[j])))
            assert ma[i, j] == (cont(c, len(rows[j])) + cont(c, len(rows[i]))) / 2
    # ANI: size accuracy is judged on the sketches as given (before downsampling), like MinHash.*_ani
    acc = [bool(DU.set_size_exact_prob(len(s.minhash) * s.minhash.scaled, s.minhash.scaled, relative_error=0.2) >= 0.95)
           for s in sigs]
    assert acc[-1] is False and sum(acc) >= n - 2           # both outcomes occur
    ani = C.compare_all_pairs(sigs, True, return_ani=True)
    jac = orc.compare_all_pairs(*orc.to_csr(rows))
    for i in range(n):
        for j in range(i + 1, n):
            r = DU.jaccard_to_distance(jac[i, j], 31, 1000, n_unique_kmers=round((len(rows[i]) + len(rows[j])) / 2 * 1000))
            want = 0.0 if (r.je_exceeds_threshold or not (acc[i] and acc[j])) else 1 - r.dist
            assert abs(ani[i, j] - want) < 1e-12 and ani[i, j] == ani[j, i]
    cani = C.compare_serial_containment(sigs, return_ani=True)
    for i in range(n):
        for j in range(n):
            if i != j:
                want = DU.containment_to_distance(m[i, j], 31, 1000, n_unique_kmers=len(rows[j]) * 1000).ani if acc[i] and acc[j] else 0.0
                assert abs(cani[i, j] - want) < 1e-12
    with pytest.raises(TypeError, match="can only calculate ANI"):
        num = [smb.SourmashSignature(smb.MinHash(50, 31), name=f"n{i}") for i in range(3)]
        C.compare_all_pairs(num, True, return_ani=True)


def test_reference_ani_matrices(cpu_kernels, golden):
    """The four ANI matrices the reference asserts for 2.fa / 2+63.fa / 47.fa / 63.fa at k=31
    (tests/test_compare.py:94-190, decimal=3), from the reference-written .sig fixtures through the
    native loader and the matrix post-processing (kernels replaced by the oracle on the CPU)."""
    import os
    from tests.conftest import GOLDEN
    kat = golden["meta"]["compare_ani_k31"]
    sigs = []
    for f in kat["order"]:
        these = [s for s in smb.load_signatures(os.path.join(GOLDEN, f), ksize=31) if s.minhash.scaled]
        sigs.extend(these)
    assert len(sigs) == 4
    np.testing.assert_array_almost_equal(C.compare_all_pairs(sigs, True, return_ani=True), np.array(kat["jaccard"]), decimal=3)
    np.testing.assert_array_almost_equal(C.compare_serial_containment(sigs, return_ani=True), np.array(kat["containment"], dtype=float), decimal=3)
    np.testing.assert_array_almost_equal(C.compare_serial_max_containment(sigs, return_ani=True), np.array(kat["max_containment"]), decimal=3)
    np.testing.assert_array_almost_equal(C.compare_serial_avg_containment(sigs, return_ani=True), np.array(kat["avg_containment"]), decimal=3)

