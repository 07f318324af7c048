"""compare --ani family on the GPU counts: matrix forms against the per-pair MinHash ANI methods
(reference: compare.py:14-187 loops over jaccard_ani / containment_ani / max_containment_ani /
avg_containment_ani; minhash.py:749-976).  Tolerance 1e-12 (numpy's vector pow vs libm pow)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sigs():
    import sourmash_b200 as smb
    from sourmash_b200.synth import synth_sketches
    assert smb.batch.device_count() > 0
    h, off = synth_sketches(14, mean=400, sd=80, lo=150, hi=700, n_families=3, pool=600, seed=11)
    out = []
    for i in range(14):
        mh = smb.MinHash(0, 31, scaled=1000)
        mh.add_many(h[int(off[i]):int(off[i + 1])])
        out.append(smb.SourmashSignature(mh, name=f"s{i}"))
    tiny = smb.MinHash(0, 31, scaled=1000)          # too few hashes: size estimate not trustworthy
    tiny.add_many(h[:3])
    out.append(smb.SourmashSignature(tiny, name="tiny"))
    return out


def _loop(sigs, fn, symmetric):
    n = len(sigs)
    m = np.ones((n, n))
    for i in range(n):
        for j in range(n):
            if i == j or (symmetric and j < i):
                continue
            v = fn(i, j)
            v = 0.0 if v is None else v
            m[i][j] = v
            if symmetric:
                m[j][i] = v
    return m


def test_compare_ani_matrices_match_pairwise(sigs):
    from sourmash_b200 import compare as C
    mh = [s.minhash for s in sigs]
    assert not mh[-1].size_is_accurate() and mh[0].size_is_accurate()
    got = C.compare_all_pairs(sigs, True, return_ani=True)
    want = _loop(sigs, lambda i, j: mh[i].jaccard_ani(mh[j]).ani, True)
    assert np.abs(got - want).max() < 1e-12 and (got[-1, :-1] == 0).all()
    got = C.compare_serial_containment(sigs, return_ani=True)
    want = _loop(sigs, lambda i, j: mh[j].containment_ani(mh[i]).ani, False)
    assert np.abs(got - want).max() < 1e-12
    got = C.compare_serial_max_containment(sigs, return_ani=True)
    want = _loop(sigs, lambda i, j: mh[j].max_containment_ani(mh[i]).ani, True)
    assert np.abs(got - want).max() < 1e-12
    got = C.compare_serial_avg_containment(sigs, return_ani=True)
    want = _loop(sigs, lambda i, j: mh[j].avg_containment_ani(mh[i]), True)
    assert np.abs(got - want).max() < 1e-12


def test_minhash_ani_methods(sigs):
    import sourmash_b200 as smb
    from sourmash_b200 import distance_utils as du
    a, b = sigs[0].minhash, sigs[1].minhash
    r = a.jaccard_ani(b)
    j = a.jaccard(b)
    assert r.dist == du.jaccard_to_distance(j, 31, 1000, n_unique_kmers=round((len(a) + len(b)) / 2 * 1000)).dist
    c = a.containment_ani(b, estimate_ci=True)
    want = du.containment_to_distance(a.contained_by(b), 31, 1000, n_unique_kmers=len(a) * 1000, estimate_ci=True)
    assert (c.dist, c.dist_low, c.dist_high) == (want.dist, want.dist_low, want.dist_high)
    assert a.max_containment_ani(b).dist == du.containment_to_distance(
        a.max_containment(b), 31, 1000, n_unique_kmers=min(len(a), len(b)) * 1000).dist
    assert a.avg_containment_ani(b) == (a.containment_ani(b).ani + b.containment_ani(a).ani) / 2
    # downsampling to the coarser scaled first (minhash.py:763-766)
    fine = smb.MinHash(0, 31, scaled=500)
    fine.add_many(list(a.hashes)[:100] + [7, 9])
    assert fine.jaccard_ani(a, downsample=True).dist == fine.downsample(scaled=1000).jaccard_ani(a).dist
    num = smb.MinHash(10, 31)
    with pytest.raises(TypeError, match="can only calculate ANI for scaled MinHashes"):
        num.jaccard_ani(a)
    # inflate: abundances of a flat sketch taken from an abundance sketch (minhash.py:1071-1092)
    ab = smb.MinHash(0, 31, scaled=1000, track_abundance=True)
    hs = sorted(a.hashes)[:50]
    ab.set_abundances({h: i + 1 for i, h in enumerate(hs)})
    flat = smb.MinHash(0, 31, scaled=1000)
    flat.add_many(hs[10:20] + [sorted(b.hashes)[-1]])
    inf = flat.inflate(ab)
    assert dict(inf.hashes) == {h: i + 1 for i, h in enumerate(hs) if 10 <= i < 20}
    with pytest.raises(ValueError):
        ab.inflate(flat)
