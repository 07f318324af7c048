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


# ------------------------------------------------------------------ protein family
def _fasta_records(path):
    name, seq, out = None, [], []
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            if name is not None:
                out.append((name, "".join(seq)))
            name, seq = line[1:], []
        elif line:
            seq.append(line)
    out.append((name, "".join(seq)))
    return out


def test_protein_encoding_kats(golden):
    kat = golden["meta"]["kat"]
    for codon, aa in kat["translate_codon"].items():
        assert orc.translate_codon(codon) == aa
    with pytest.raises(ValueError):
        orc.translate_codon("")
    with pytest.raises(ValueError):
        orc.translate_codon("TCTA")
    # tests/test_minhash.py:390-454
    assert orc.seq_to_hashes_protein("CADHIFC", 7, "dayhoff").tolist() == [orc.hash_murmur(kat["dayhoff_CADHIFC"])]
    assert orc.seq_to_hashes_protein("CADHIF*", 7, "dayhoff").tolist() == [orc.hash_murmur(kat["dayhoff_CADHIF*"])]
    assert orc.seq_to_hashes_protein("ANA", 3, "hp").tolist() == [orc.hash_murmur(kat["hp_ANA"])]
    assert orc.seq_to_hashes_protein("AN*", 3, "hp").tolist() == [orc.hash_murmur(kat["hp_AN*"])]
    for mol, n in kat["AGYYG_k2"].items():
        assert len(set(orc.seq_to_hashes_protein("AGYYG", 2, mol).tolist())) == n
    for mol in ("protein", "dayhoff", "hp"):
        assert len(set(orc.seq_to_hashes_translate("ACTGAC", 2, mol).tolist())) == kat["ACTGAC_translate_k2"]
    assert len(orc.seq_to_hashes_translate("ACTGA", 2, "dayhoff")) == 0      # test_minhash.py:283-287
    assert len(orc.seq_to_hashes_protein("AG", 9, "protein")) == 0          # test_minhash.py:456-461
    assert len(orc.seq_to_hashes_protein("AGY", 2, "protein")) == kat["AGY_k2_protein"]
    assert len(orc.seq_to_hashes_protein("AGY", 1, "protein")) == kat["AGY_k1_protein"]
    with pytest.raises(ValueError):
        orc.seq_to_hashes_protein("ATGAGAGACGATAGACAGATGACC", 7, "dna")     # test_minhash.py:242-249
    # lower case is folded before hashing (signature.rs:214)
    assert orc.seq_to_hashes_protein("agyyg", 2).tolist() == orc.seq_to_hashes_protein("AGYYG", 2).tolist()
    # force && bad_kmers_as_zeroes: the iterator's two bookkeeping zeros surround the hashes
    hz = orc.seq_to_hashes_translate("ACTGACTGA", 2, keep_zeros=True)
    h = orc.seq_to_hashes_translate("ACTGACTGA", 2)
    assert hz[0] == 0 and hz[-1] == 0 and hz[1:-1].tolist() == h.tolist() and len(h) == 2 * (9 - 6 + 1)


def test_protein_benchmark_sigs(golden, golden_dir):
    """tests/test_sourmash_sketch.py:1340-1376: the reference's own known-good protein sketches."""
    info = golden["meta"]["protein_benchmarks"]
    prots = dict((n.split()[0], s) for n, s in _fasta_records(golden_dir / "ecoli.faa"))
    genes = dict((n.split()[0], s) for n, s in _fasta_records(golden_dir / "ecoli.genes.fna"))
    i = info["input_prot"]
    mh = orc.OracleMinHash(scaled=0, ksize=i["ksize"], num=i["num"], seed=i["seed"])
    mh.add_protein_family(prots[i["name"].split()[0]], "protein", True)
    assert np.array_equal(mh.mins(), golden["arrays"]["bench_input_prot"])
    assert mh.md5sum() == i["md5sum"]
    t = info["translate_prot"]
    mh = orc.OracleMinHash(scaled=0, ksize=t["ksize"], num=t["num"], seed=t["seed"])
    mh.add_protein_family(genes[t["name"].split()[0]], "protein", False)
    assert np.array_equal(mh.mins(), golden["arrays"]["bench_translate_prot"])
    assert mh.md5sum() == t["md5sum"]


def test_protein_2x2_similarities(golden, golden_dir):
    """tests/test_sourmash_compute.py:810-860: protein input vs translated genes, k=21, num=500."""
    want = golden["meta"]["protein_2x2"]
    prots = sorted(_fasta_records(golden_dir / "ecoli.faa"))
    genes = sorted(_fasta_records(golden_dir / "ecoli.genes.fna"))
    aa, tr = [], []
    for _, s in prots:
        mh = orc.OracleMinHash(scaled=0, ksize=21, num=500)
        mh.add_protein_family(s, "protein", True)
        aa.append(mh.mins())
    for _, s in genes:
        mh = orc.OracleMinHash(scaled=0, ksize=21, num=500)
        mh.add_protein_family(s, "protein", False)
        tr.append(mh.mins())
    assert round(orc.jaccard(aa[0], tr[0], num=500), 3) == want["aa1_trans1"]
    assert round(orc.jaccard(aa[1], tr[0], num=500), 3) == want["aa2_trans1"]
    assert round(orc.jaccard(aa[0], tr[1], num=500), 3) == want["aa1_trans2"]
    assert round(orc.jaccard(aa[1], tr[1], num=500), 3) == want["aa2_trans2"]


def test_translate_golden_genome_s10(golden, s10_records):
    """genome-s10.fa.gz.sig also carries protein sketches (k = 7 and 10 residues, num=500) that the
    reference computed from the DNA by six-frame translation: a 500 kbp golden for that path."""
    info = golden["meta"]["genome_s10_protein"]
    assert sorted(info) == ["21", "30"]
    for k3, meta in info.items():
        assert meta["molecule"] == "protein" and meta["num"] == 500
        mh = orc.OracleMinHash(scaled=0, ksize=int(k3), num=500, seed=meta["seed"])
        for _, seq in s10_records:
            mh.add_protein_family(seq, "protein", False)
        assert np.array_equal(mh.mins(), golden["arrays"][f"s10_prot_k{k3}"])
        assert mh.md5sum() == meta["md5sum"]


def test_one_vs_many_bsearch_equals_the_two_pointer_walk():
    """oracle.one_vs_many_bsearch (used to check full-size configs[3] results) gives the counts of the faithful
    count_common walk: random, planted, empty, extreme-key rows and queries."""
    rng = np.random.default_rng(11)
    big = np.uint64(2**64 - 1)
    q = np.unique(np.concatenate([rng.integers(0, 2**64 - 1, size=20_000, dtype=np.uint64), np.array([0, big], dtype=np.uint64)]))
    rows = [np.unique(rng.integers(0, 2**64 - 1, size=int(rng.integers(0, 400)), dtype=np.uint64)) for _ in range(60)]
    rows += [np.unique(np.concatenate([r, q[rng.integers(0, len(q), size=50)]])) for r in rows[:20]]
    rows += [np.zeros(0, np.uint64), np.array([0], np.uint64), np.array([big], np.uint64), q.copy(), q[::7].copy()]
    h, off = orc.to_csr(rows)
    for query in (q, q[:1], np.zeros(0, np.uint64), rows[3]):
        a = orc.one_vs_many(query, h, off, nthreads=2)
        b = orc.one_vs_many_bsearch(query, h, off, nthreads=2)
        assert np.array_equal(a, b)
    assert int(orc.one_vs_many_bsearch(q, h, off)[-2]) == len(q)
