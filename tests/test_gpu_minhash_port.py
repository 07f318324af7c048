"""Scenarios of the reference's tests/test_minhash.py re-expressed against sourmash_b200.MinHash
(each test cites the reference test it mirrors).  Comparisons run on the GPU through the
reference-compatible ABI."""
import math

import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import sourmash_b200
    assert sourmash_b200.batch.device_count() > 0
    return sourmash_b200.MinHash


@pytest.fixture(params=[True, False])
def track_abundance(request):
    return request.param


def _scaled(max_hash):
    from sourmash_b200.minhash import _get_scaled_for_max_hash
    return _get_scaled_for_max_hash(max_hash)


A_VALUES = {1: 5, 3: 3, 5: 2, 8: 2}
B_VALUES = {1: 3, 3: 2, 5: 1, 6: 1, 8: 1, 10: 1}


def _fill(mh, values, track):
    if track:
        mh.set_abundances(values)
    else:
        mh.add_many(values.keys())
    return mh


def test_div_zero(M, track_abundance):                       # test_minhash.py:115-132
    mh, mh2 = M(1, 4, track_abundance=track_abundance), M(1, 4, track_abundance=track_abundance)
    mh2.add_sequence("ATGC")
    assert mh.similarity(mh2) == 0 and mh2.similarity(mh) == 0
    s, s2 = M(0, 4, scaled=1, track_abundance=track_abundance), M(0, 4, scaled=1, track_abundance=track_abundance)
    s2.add_sequence("ATGC")
    assert s.contained_by(s2) == 0 and s2.contained_by(s) == 0


def test_contained_requires_scaled(M, track_abundance):      # test_minhash.py:135-178
    a, b = M(1, 4, track_abundance=track_abundance), M(0, 4, scaled=1, track_abundance=track_abundance)
    for x, y in ((a, b), (b, a)):
        with pytest.raises(TypeError, match="can only calculate containment for scaled MinHashes"):
            x.contained_by(y)
        with pytest.raises(TypeError):
            x.max_containment(y)
        with pytest.raises(TypeError):
            x.avg_containment(y)


def test_bytes_and_str_dna(M, track_abundance):              # test_minhash.py:180-195
    mh = M(1, 4, track_abundance=track_abundance)
    mh.add_sequence("ATGC")
    mh.add_sequence(b"ATGC")
    assert len(mh.hashes) == 1 and list(mh.hashes) == [12415348535738636339]
    if track_abundance:
        assert mh.hashes[12415348535738636339] == 2


def test_size_limit_and_scaled_filter(M, track_abundance):   # test_minhash.py:464-491
    mh = M(3, 4, track_abundance=track_abundance)
    for h in (10, 20, 30):
        mh.add_hash(h)
    assert list(sorted(mh.hashes)) == [10, 20, 30]
    mh.add_hash(5)
    assert list(sorted(mh.hashes)) == [5, 10, 20]
    s = M(0, 4, track_abundance=track_abundance, scaled=_scaled(35))
    assert s._max_hash == 35
    for h in (10, 20, 30):
        s.add_hash(h)
    s.add_hash(40)
    assert list(sorted(s.hashes)) == [10, 20, 30]
    s.add_hash(36)
    assert list(sorted(s.hashes)) == [10, 20, 30]


def test_jaccard_and_downsample_values(M):                   # test_minhash.py:542-564
    a, b = M(0, 20, scaled=_scaled(50)), M(0, 20, scaled=_scaled(50))
    a.add_many([1, 3, 5, 8]); b.add_many([1, 3, 5, 6, 8, 10])
    assert a.similarity(b) == 4.0 / 6.0
    a, b = M(0, 20, scaled=_scaled(50)), M(0, 20, scaled=_scaled(100))
    a.add_many([1, 3, 5, 8, 70]); b.add_many([1, 3, 5, 6, 8, 10, 70])
    assert a.similarity(b, downsample=True) == 4.0 / 6.0      # 70 is dropped by the downsample


def test_angular_similarity_values(M):                       # test_minhash.py:567-614
    a, b = M(0, 20, scaled=_scaled(50), track_abundance=True), M(0, 20, scaled=_scaled(50), track_abundance=True)
    a.set_abundances(A_VALUES); b.set_abundances(B_VALUES)
    angular = 1 - 2 * math.acos(0.9356) / math.pi
    assert round(angular, 4) == 0.7703 == round(a.similarity(b), 4)
    a, b = M(0, 20, scaled=_scaled(100), track_abundance=True), M(0, 20, scaled=_scaled(100), track_abundance=True)
    a.set_abundances({**A_VALUES, 70: 70}); b.set_abundances({**B_VALUES, 70: 70})
    assert round(a.similarity(b), 4) == 0.9728
    assert a.similarity(b, ignore_abundance=True) == 5.0 / 7.0
    a2 = M(0, 20, scaled=_scaled(50), track_abundance=True)
    a2.set_abundances({**A_VALUES, 70: 70})
    assert round(a2.similarity(b, downsample=True), 4) == 0.7703
    assert a2.similarity(b, downsample=True, ignore_abundance=True) == 4.0 / 6.0


def test_similarity_downsample_symmetry_and_errors(M, track_abundance):   # test_minhash.py:648-708
    a = _fill(M(0, 20, scaled=_scaled(50), track_abundance=track_abundance), A_VALUES, track_abundance)
    b = _fill(M(0, 20, scaled=_scaled(100), track_abundance=track_abundance), B_VALUES, track_abundance)
    for ia in (True, False):
        assert a.similarity(b, ignore_abundance=ia, downsample=True) == b.similarity(a, ignore_abundance=ia, downsample=True)
        for x, y in ((a, b), (b, a)):
            with pytest.raises(ValueError, match="mismatch in scaled; comparison fail"):
                x.similarity(y, ignore_abundance=ia)


def test_similarity_1(M, track_abundance):                   # test_minhash.py:768-794
    import oracle as orc
    s1 = "TGCCGCCCAGCACCGGGTGACTAGGTTGAGCCATGATTAACCTGCAATGA"
    s2 = "GATTGGTGCACACTTAACTGGGTGCCGCGCTGGTGCTGATCCATGAAGTT"
    a, b = M(20, 10, track_abundance=track_abundance), M(20, 10, track_abundance=track_abundance)
    a.add_sequence(s1); b.add_sequence(s1)
    for x, y in ((a, b), (b, b), (b, a), (a, a)):
        assert round(x.similarity(y), 3) == 1.0
    b.add_sequence(s1)                                        # same sequence again
    for x, y in ((a, b), (b, b), (b, a), (a, a)):
        assert round(x.similarity(y), 3) == 1.0
    b.add_sequence(s2)
    assert a.similarity(b) >= 0.3 and b.similarity(a) >= 0.3
    assert round(a.similarity(a), 3) == 1.0 and round(b.similarity(b), 3) == 1.0
    # and the exact value, from the oracle
    oa = orc.OracleMinHash(ksize=10, num=20, track_abundance=track_abundance)
    ob = orc.OracleMinHash(ksize=10, num=20, track_abundance=track_abundance)
    oa.add_sequence(s1); ob.add_sequence(s1); ob.add_sequence(s1); ob.add_sequence(s2)
    if track_abundance:
        want = orc.angular_similarity(oa.mins(), oa.abunds(), ob.mins(), ob.abunds())
        assert abs(a.similarity(b) - want) < 1e-12
    else:
        assert a.similarity(b) == orc.jaccard(oa.mins(), ob.mins(), num=20)


def test_count_common_and_incompatibilities(M, track_abundance):   # test_minhash.py:845-908
    a, b = M(20, 10, track_abundance=track_abundance), M(20, 10, track_abundance=track_abundance)
    for i in range(0, 40, 2):
        a.add_hash(i)
    for i in range(0, 80, 4):
        b.add_hash(i)
    assert a.count_common(b) == 10 == b.count_common(a)
    with pytest.raises(ValueError):
        a.count_common(M(20, 5, is_protein=True, track_abundance=track_abundance))
    with pytest.raises(ValueError):
        M(0, 5, scaled=_scaled(1), track_abundance=track_abundance).count_common(
            M(0, 5, scaled=_scaled(2), track_abundance=track_abundance))
    with pytest.raises(ValueError):
        M(20, 5, seed=1, track_abundance=track_abundance).count_common(M(20, 5, seed=2, track_abundance=track_abundance))
    with pytest.raises(ValueError):
        M(20, 5, track_abundance=track_abundance).count_common(M(20, 6, track_abundance=track_abundance))
    with pytest.raises(TypeError, match="Must be a MinHash!"):
        a.count_common(set())


def test_jaccard_asymmetric_num(M, track_abundance):         # test_minhash.py:916-937
    a, b = M(20, 10, track_abundance=track_abundance), M(10, 10, track_abundance=track_abundance)
    for i in range(0, 40, 2):
        a.add_hash(i)
    for i in range(0, 80, 4):
        b.add_hash(i)
    with pytest.raises(TypeError):
        a.jaccard(b)
    a = a.downsample(num=10)
    assert a.count_common(b) == 5 and b.count_common(a) == 5
    assert a.jaccard(b) == 0.5 and b.jaccard(a) == 0.5        # bottom-10 of the union: 0,2,..,18 -> 5 shared


def test_merge_semantics(M, track_abundance):                # test_minhash.py:945-1041
    a, b = M(100, 10, track_abundance=track_abundance), M(100, 10, track_abundance=track_abundance)
    for i in range(0, 40, 2):
        a.add_hash(i)
    for i in range(0, 80, 4):
        b.add_hash(i)
    c, d = a.__copy__(), b.__copy__()
    c.merge(b); d.merge(a)
    assert len(c) == len(d) == 30 and sorted(c.hashes.items()) == sorted(d.hashes.items())
    assert round(c.similarity(d), 3) == 1.0 and round(d.similarity(c), 3) == 1.0
    e = M(100, 10, track_abundance=track_abundance)
    e.merge(a)
    assert list(e.hashes) == list(a.hashes)
    with pytest.raises(TypeError):
        a.merge(set())
    with pytest.raises(TypeError):
        a += set()


def test_asymmetric_merge_and_concat(M, track_abundance):    # test_minhash.py:1043-1141
    a, b = M(20, 10, track_abundance=track_abundance), M(10, 10, track_abundance=track_abundance)
    for i in range(0, 40, 2):
        a.add_hash(i)
    for i in range(0, 80, 4):
        b.add_hash(i)
    c, d = a.__copy__(), b.__copy__()
    c.merge(b); d.merge(a)
    assert len(a) == 20 and len(b) == 10 and len(c) == len(a) and len(d) == len(b)
    with pytest.raises(TypeError):
        d.jaccard(a)
    a1 = a.downsample(num=d.num)
    if track_abundance:
        assert round(d.similarity(a1), 3) == 0.795
    else:
        assert round(d.similarity(a1), 3) == 1.0
    c1 = c.downsample(num=b.num)
    if track_abundance:
        assert round(c1.similarity(b), 3) == 0.436
    else:
        assert c1.similarity(b) == 0.5


def test_abundance_bookkeeping(M):                           # test_minhash.py:1267-1495
    a = M(20, 5, track_abundance=True)
    a.add_sequence("AAAAA")
    assert list(a.hashes) == [2110480117637990133] and a.hashes[2110480117637990133] == 1
    a.add_sequence("AAAAA")
    assert a.hashes[2110480117637990133] == 2
    a.add_hash_with_abundance(10, 3)
    assert a.hashes[10] == 3
    with pytest.raises(RuntimeError):
        M(20, 5).add_hash_with_abundance(10, 1)
    a.clear()
    assert len(a) == 0 and a.track_abundance
    a.set_abundances({1: 3, 2: 4})
    a.set_abundances({1: 0}, clear=False)                   # abundance 0 removes (test_clear_abundance_on_zero)
    assert dict(a.hashes) == {2: 4}
    a.set_abundances({2: 1}, clear=False)
    assert a.hashes[2] == 5
    with pytest.raises(RuntimeError):
        M(20, 5).set_abundances({1: 1})
    n = M(2, 10, track_abundance=True)
    n.set_abundances({1: 3, 2: 4, 3: 5})                     # test_set_abundance_num
    assert dict(n.hashes) == {1: 3, 2: 4}


def test_abundance_count_common_and_similarity(M):           # test_minhash.py:1337-1381
    a, b = M(20, 5, track_abundance=True), M(20, 5)
    a.add_sequence("AAAAA"); a.add_sequence("AAAAA"); b.add_sequence("AAAAA")
    assert a.count_common(b) == 1 == b.count_common(a)
    b.add_sequence("GGGGG")
    assert sorted(b.hashes) == [2110480117637990133, 10798773792509008305]
    assert a.count_common(b) == 1 == b.count_common(a)
    s1 = "TGCCGCCCAGCACCGGGTGACTAGGTTGAGCCATGATTAACCTGCAATGA"
    a2, b2 = M(20, 10, track_abundance=True), M(20, 10, track_abundance=False)
    a2.add_sequence(s1); b2.add_sequence(s1)
    assert round(a2.similarity(b2), 3) == 1.0 and round(b2.similarity(a2), 3) == 1.0   # falls back to jaccard
    b2.add_sequence("GATTGGTGCACACTTAACTGGGTGCCGCGCTGGTGCTGATCCATGAAGTT")
    assert a2.similarity(b2) >= 0.3 and b2.similarity(a2) >= 0.3


def test_add_remove_many_and_views(M, track_abundance):      # test_minhash.py:1748-1828, 1988-2017
    a = M(0, 10, track_abundance=track_abundance, scaled=_scaled(5000))
    b = M(0, 10, track_abundance=track_abundance, scaled=_scaled(5000))
    a.add_many(list(range(0, 100, 2))); a.add_many(list(range(0, 100, 2)))
    for h in range(0, 100, 2):
        b.add_hash(h); b.add_hash(h)
    assert len(a) == 50 == len(b) and a == b
    a.remove_many(range(0, 100, 4))
    assert len(a) == 25 and all(c % 4 != 0 for c in a.hashes)
    c = M(0, 10, track_abundance=track_abundance, scaled=_scaled(5000))
    c.add_many(range(2, 100, 4))
    a.remove_many(c)                                          # remove a whole MinHash
    assert len(a) == 0
    b.add_many(c)
    assert len(b) == 50
    assert set(b.get_mins()) == set(b.hashes) == set(b.get_hashes())
    with pytest.raises(RuntimeError):
        b.hashes[100] = 1


def test_add_kmer(M, track_abundance):                       # test_minhash.py:1962-1986
    a, b = M(0, 7, scaled=1, track_abundance=track_abundance), M(0, 7, scaled=1, track_abundance=track_abundance)
    seq = "AAAAAAATGCCGTCGTT"
    a.add_sequence(seq)
    for i in range(len(seq) - 7 + 1):
        b.add_kmer(seq[i:i + 7])
    assert a == b
    with pytest.raises(ValueError, match="kmer to add is not 7 in length"):
        b.add_kmer(seq[:8])


def test_addition_and_iaddition(M):                          # test_minhash.py:2097-2165
    with pytest.raises(TypeError, match="incompatible num values"):
        M(10, 21) + M(20, 21)
    for track in (True, False):
        a, b = M(10, 21, track_abundance=track), M(10, 21, track_abundance=track)
        a.add_hash(10); b.add_hash(10); b.add_hash(20)
        c = a + b
        assert len(c) == 2 and (c.hashes[10] == (2 if track else 1))
        a += b
        assert dict(a.hashes) == dict(c.hashes)
    with pytest.raises(TypeError):
        M(10, 21) + 5


def test_intersections(M):                                   # test_minhash.py:2167-2314
    a, b = M(10, 21), M(10, 21)
    a.add_hash(10); b.add_hash(10); b.add_hash(20)
    for mh in (a.intersection(b), a & b, b & a):
        assert list(mh.hashes) == [10]
    s1, s2 = M(0, 21, scaled=1), M(0, 21, scaled=1)
    s1.add_hash(10); s2.add_hash(10); s2.add_hash(20)
    assert list((s1 & s2).hashes) == [10]
    with pytest.raises(TypeError, match="can only intersect flat MinHash objects"):
        M(0, 21, scaled=1, track_abundance=True).intersection(s1)
    with pytest.raises(ValueError, match="different ksizes cannot be compared"):
        M(0, 31, scaled=1).intersection(s1)
    with pytest.raises(TypeError):
        s1.intersection(set())
    # full num sketches: intersection is restricted to the bottom-num of the union (test_intersection_6_full_num)
    n1, n2 = M(20, 21), M(20, 21)
    for i in range(20):
        n1.add_hash(i)
    for i in range(10, 30):
        n2.add_hash(i)
    common = n1 & n2
    assert sorted(common.hashes) == list(range(10, 20))
    assert n1.intersection_and_union_size(n2) == (10, 20) and n1.jaccard(n2) == 0.5
    # full scaled sketches (test_intersection_7_full_scaled)
    f1, f2 = M(0, 21, scaled=100), M(0, 21, scaled=100)
    for i in range(20):
        f1.add_hash(i)
    for i in range(10, 30):
        f2.add_hash(i)
    assert sorted((f1 & f2).hashes) == list(range(10, 20))
    assert f1.intersection_and_union_size(f2) == (10, 30)
    with pytest.raises(TypeError, match="incompatible MinHash objects"):
        M(0, 31, scaled=100).intersection_and_union_size(f1)


def test_flatten_inflate(M):                                 # test_minhash.py:1830-1960
    mh = M(0, 21, scaled=1, track_abundance=True)
    mh.set_abundances({10: 2, 20: 3, 30: 1})
    flat = mh.flatten()
    assert not flat.track_abundance and list(flat.hashes) == [10, 20, 30] and flat.hashes[20] == 1
    assert flat.flatten() is flat
    scaled_mh = M(0, 21, scaled=1)
    scaled_mh.add_many([10, 20, 40])
    inflated = scaled_mh.inflate(mh)
    assert dict(inflated.hashes) == {10: 2, 20: 3}           # 40 has no abundance -> dropped
    with pytest.raises(ValueError, match="inflate operates on a flat MinHash"):
        mh.inflate(mh)
    with pytest.raises(ValueError):
        scaled_mh.inflate(scaled_mh)


def test_distance_matrix_and_copy(M, track_abundance):       # test_minhash.py:797-821, 1722-1745
    import numpy as np
    seqs = ["TGCCGCCCAGCACCGGGTGACTAGGTTGAGCCATGATTAACCTGCAATGA", "TGCCGCCCAGCACCGGGTGACTAGGTTGAGCCATGATTAACCTGCAATGG",
            "TGCCGCCCAGCACCGGGTGACTAGGTTGAGCCATGATTAACCTGCAATGC"]
    mhs = []
    for s in seqs:
        mh = M(500, 12, track_abundance=track_abundance)
        mh.add_sequence(s)
        mhs.append(mh)
    D = np.array([[a.similarity(b) for b in mhs] for a in mhs])
    assert np.allclose(D, D.T) and np.all(np.diag(D) == 1.0)
    c = mhs[0].__copy__()
    assert c == mhs[0] and c.similarity(mhs[0]) == 1.0
    f = mhs[0].to_frozen()
    assert f.__copy__() is f and f.similarity(mhs[1]) == mhs[0].similarity(mhs[1])
