"""Pin the CPU oracle (oracle/oracle.c) to the reference's own golden vectors."""
import numpy as np
import pytest

import oracle as orc


def test_hash_murmur_kat(golden):
    kat = golden["meta"]["kat"]
    assert orc.hash_murmur("ACG", 42) == kat["hash_murmur_ACG_42"]
    mh = orc.OracleMinHash(scaled=0, ksize=4, num=1)
    mh.add_sequence("ATGC")
    assert mh.mins().tolist() == kat["n1_k4_ATGC"]


def test_max_hash_for_scaled(golden):
    kat = golden["meta"]["kat"]
    assert orc.max_hash_for_scaled(100) == kat["max_hash_scaled_100"]
    assert orc.max_hash_for_scaled(1000) == kat["max_hash_scaled_1000"]
    assert orc.max_hash_for_scaled(0) == 0
    assert orc.max_hash_for_scaled(1) == 2**64 - 1


def test_merge_kat(golden):
    kat = golden["meta"]["kat"]["merge_k10_num20"]
    a = orc.OracleMinHash(scaled=0, ksize=10, num=20)
    b = orc.OracleMinHash(scaled=0, ksize=10, num=20)
    for s in kat["a"]:
        assert a.add_sequence(s, force=False) is None
    for s in kat["b"]:
        assert b.add_sequence(s, force=False) is None
    a.merge(b)
    assert a.mins().tolist() == kat["merged"]


def test_invalid_dna_kat(golden):
    for seq, n in golden["meta"]["kat"]["invalid_dna_k3"].items():
        mh = orc.OracleMinHash(scaled=0, ksize=3, num=20)
        assert mh.add_sequence(seq, force=True) is None
        assert len(mh) == n
    mh = orc.OracleMinHash(scaled=0, ksize=3, num=20)
    assert mh.add_sequence("AAANNCCCTN", force=False) == 1   # window 1 = "AAN" is the first bad one
    assert len(mh) == 1                                        # "AAA" was added before the error
    hs, err = orc.seq_to_hashes("ATGR", 4, force=False)
    assert err == 0 and len(hs) == 0
    hs, err = orc.seq_to_hashes("AAANNCCCTN", 3, force=True, keep_zeros=True)
    assert err is None and len(hs) == 8 and int((hs == 0).sum()) == 5
    hs, _ = orc.seq_to_hashes("acgtacgt", 4)
    hs2, _ = orc.seq_to_hashes("ACGTACGT", 4)
    assert hs.tolist() == hs2.tolist()
    assert len(orc.seq_to_hashes("ACG", 4)[0]) == 0


@pytest.mark.parametrize("k", [21, 31, 51])
def test_ecoli_golden_sketch(golden, ecoli_seq, k):
    info = golden["meta"]["ecoli"][str(k)]
    want = golden["arrays"][f"ecoli_k{k}"]
    got = orc.sketch_scaled(ecoli_seq, k, info["max_hash"], seed=info["seed"])
    assert len(got) == info["n"]
    assert np.array_equal(got, want)
    assert orc.md5sum(k, got) == info["md5sum"]


def test_ecoli_golden_via_minhash_object(golden, ecoli_seq):
    # the literal add_sequence -> add_hash path on a prefix (full genome would be slow-ish)
    seq = ecoli_seq[:300000]
    mh = orc.OracleMinHash(scaled=1000, ksize=31)
    assert mh.add_sequence(seq, force=True) is None
    assert np.array_equal(mh.mins(), orc.sketch_scaled(seq, 31, mh.max_hash))


@pytest.mark.parametrize("k", [21, 30])
def test_genome_s10_num_sketch(golden, s10_records, k):
    info = golden["meta"]["genome_s10"][str(k)]
    mh = orc.OracleMinHash(scaled=0, ksize=k, num=info["num"], seed=info["seed"])
    for _, seq in s10_records:
        assert mh.add_sequence(seq, force=True) is None
    assert np.array_equal(mh.mins(), golden["arrays"][f"s10_k{k}"])
    assert mh.md5sum() == info["md5sum"]


def test_47_63_counts(golden):
    a, b = golden["arrays"]["s47"], golden["arrays"]["s63"]
    want = golden["meta"]["s47_s63"]
    assert (len(a), len(b)) == (want["n47"], want["n63"])
    assert orc.intersection_size(a, b) == (want["common"], want["union"])
    assert orc.count_common(a, b) == want["common"]
    assert orc.jaccard(a, b) == want["common"] / want["union"]
    assert orc.md5sum(31, a) == golden["meta"]["s47_md5"]
    assert orc.md5sum(31, b) == golden["meta"]["s63_md5"]


def test_demo_matrix_num500(golden):
    rows = [golden["arrays"][f"demo{i}"] for i in range(7)]
    hashes, offsets = orc.to_csr(rows)
    got = orc.compare_all_pairs(hashes, offsets, num=500)
    assert np.array_equal(got, np.array(golden["meta"]["demo_matrix"]))


def test_scaled100_real_data(golden):
    a, b = golden["arrays"]["scaled100_ecoli"], golden["arrays"]["scaled100_salmonella"]
    want = golden["meta"]["scaled100_jaccard"]
    assert orc.intersection_size(a, b) == (1522, 92559)
    for scaled, digits in ((100, 5), (1000, 5), (10000, 3), (100000, 2)):
        mx = orc.max_hash_for_scaled(scaled)
        j = orc.jaccard(orc.downsample(a, mx), orc.downsample(b, mx))
        assert round(j, digits) == want[str(scaled)]


def test_n10000_real_data(golden):
    a, b = golden["arrays"]["n10000_ecoli"], golden["arrays"]["n10000_salmonella"]
    want = golden["meta"]["n10000_jaccard"]
    for num in (10000, 1000, 100, 10):
        assert orc.jaccard(a[:num], b[:num], num=num) == want[str(num)]


def test_angular_similarity_small():
    # self-similarity is 1; disjoint is 0 (src/core/src/sketch/minhash.rs:635-680)
    a = np.array([1, 5, 9], dtype=np.uint64)
    ab = np.array([2, 1, 7], dtype=np.uint64)
    assert orc.angular_similarity(a, ab, a, ab) == pytest.approx(1.0, abs=1e-7)
    b = np.array([2, 6], dtype=np.uint64)
    bb = np.array([1, 1], dtype=np.uint64)
    assert orc.angular_similarity(a, ab, b, bb) == 0.0
