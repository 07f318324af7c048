"""Protein / dayhoff / hp sketches on the GPU: residues through add_protein,
DNA through six-frame translation.  Scenarios of the reference's tests/test_minhash.py
(:221-461, :2630-2870) and tests/test_sourmash_sketch.py:1340-1376, each checked bit-for-bit
against the oracle and the reference's golden sketches."""
import numpy as np
import pytest

import oracle as orc

pytestmark = pytest.mark.gpu

PROT = ("MVKVYAPASSANMSVGFDVLGAAVTPVDGALLGDVVTVEAAETFSLNNLGRFADKLPSEPRENIVYQCWERFCQELGKQIPVAMTLEKNMPIGSGLGSSACSVVAALMAMNEH"
        "CGKPLNDTRLLALMGELEGRISGSIHYDNVAPCFLGGMQLMIEENDIISQQVPGFDEWLWVLAYPGIKVSTAEARAILPAQYRRQDCIAHGRHLAGFIHACYSRQPELAAKLM"
        "KDVIAEPYRERLLPGFRQARQAVAEIGAVASGISGSGPTLFALCDKPETAQRVADWLGKNYLQNQEGFVHICRLDTAGARVLEN*")
DNA = ("atggttaaagtttatgccccggcttccagtgccaatatgagcgtcgggtttgatgtgctcggggcggcggtgacacctgttgatggtgcattgctcggagatgtagtcacggt"
       "tgaggcggcagagacattcagtctcaacaacctcggacgctttgccgataagctgccgtcagaaccacgggaaaatatcgtttatcagtgctgggagcgtttttgccaggaactg"
       "ggtaagcaaattccagtggcgatgaccctggaaaagaatatgccgatcggttcgggcttaggctccagtgcctgttcggtggtcgcggcgctgatggcgatgaatgaacactgcg"
       "gcaagccgcttaatgacactcgtttgctggctttgatgggcgagctggaaggccgtatctccggcagcattcattacgacaacgtggcaccgtgttttctcggtggtatgcagtt"
       "gatgatcgaagaaaacgacatcatcagccagcaagtgccagggtttgatgagtggctgtgggtgctggcgtatccggggattaaagtctcgacggcagaagccagggctatttta"
       "ccggcgcagtatcgccgccaggattgcattgcgcacgggcgacatctggcaggcttcattcacgcctgctattcccgtcagcctgagcttgccgcgaagctgatgaaagatgtta"
       "tcgctgaaccctaccgtgaacggttactgccaggcttccggcaggcgcggcaggcggtcgcggaaatcggcgcggtagcgagcggtatctccggctccggcccgaccttgttcgc"
       "tctgtgtgacaagccggaaaccgcccagcgcgttgccgactggttgggtaagaactacctgcaaaatcaggaaggttttgttcatatttgccggctggatacggcgggcgcacga"
       "gtactggaaaactaa")


@pytest.fixture(scope="module")
def smb():
    import sourmash_b200
    assert sourmash_b200.batch.device_count() > 0
    return sourmash_b200


def _mh(smb, moltype, k, **kw):
    return smb.MinHash(kw.pop("n", 0), k, is_protein=moltype == "protein", dayhoff=moltype == "dayhoff",
                       hp=moltype == "hp", **kw)


def _records(path):
    from tests.conftest import read_fasta
    return [(n, s) for n, s in read_fasta(str(path))]


@pytest.mark.parametrize("moltype", ["protein", "dayhoff", "hp"])
def test_seq_to_hashes_protein_in_order(smb, moltype):                 # test_minhash.py:2630-2776
    mh = _mh(smb, moltype, 7, scaled=1)
    got = mh.seq_to_hashes(PROT, is_protein=True)
    want = orc.seq_to_hashes_protein(PROT, 7, moltype)
    assert got == want.tolist() and len(got) == len(PROT) - 7 + 1
    mh.add_protein(PROT)
    assert set(got) == set(mh.hashes)
    for kmer, h in mh.kmers_and_hashes(PROT[:40], is_protein=True):
        single = mh.copy_and_clear()
        single.add_protein(kmer)
        assert list(single.hashes) == [h]


@pytest.mark.parametrize("moltype", ["protein", "dayhoff", "hp"])
def test_translate_six_frames(smb, moltype):                            # test_minhash.py:2780-2870
    mh = _mh(smb, moltype, 7, scaled=1)
    ht = mh.seq_to_hashes(DNA)
    assert ht == orc.seq_to_hashes_translate(DNA, 7, moltype).tolist()
    hp_ = mh.seq_to_hashes(PROT, is_protein=True)
    assert set(hp_).issubset(set(ht)) and not set(ht).issubset(set(hp_))
    # kmers_and_hashes pairs every DNA k-mer of every coding frame with its forward-only hash
    pairs = list(mh.kmers_and_hashes(DNA.upper()))
    assert len(pairs) == 2 * (len(DNA) - 21 + 1)
    for kmer, h in pairs[:50] + pairs[-50:]:
        aa = "".join(smb.translate_codon(kmer[i:i + 3]) for i in range(0, 21, 3))
        assert orc.seq_to_hashes_protein(aa, 7, moltype).tolist() == [h]
    mh.add_sequence(DNA)
    assert set(mh.hashes) == set(ht)


def test_small_kats(smb, golden):                                       # test_minhash.py:221-461
    kat = golden["meta"]["kat"]
    for moltype, n in kat["AGYYG_k2"].items():
        for track in (False, True):
            mh = _mh(smb, moltype, 2, n=10, track_abundance=track)
            mh.add_protein("AGYYG"); mh.add_protein("AGYYG"); mh.add_protein(b"AGYYG")
            assert len(mh.hashes) == n and mh.moltype == moltype
            assert set(mh.hashes) == set(mh.seq_to_hashes("AGYYG", is_protein=True))
            if track:
                assert sum(mh.hashes.values()) == 12
    for moltype in ("protein", "dayhoff", "hp"):
        mh = _mh(smb, moltype, 2, n=10)
        mh.add_sequence("ACTGAC")
        assert len(mh.hashes) == kat["ACTGAC_translate_k2"]
        assert set(mh.hashes) == set(mh.seq_to_hashes("ACTGAC"))
    assert len(_mh(smb, "dayhoff", 2, scaled=1).seq_to_hashes("ACTGA")) == 0
    mh = _mh(smb, "protein", 9, n=10)
    mh.add_protein("AG")
    assert len(mh.hashes) == 0
    d = _mh(smb, "dayhoff", 7, scaled=1, track_abundance=True)
    d.add_protein("CADHIFC")
    assert list(d.hashes) == [smb.hash_murmur(kat["dayhoff_CADHIFC"])]
    d = d.copy_and_clear(); d.add_protein("CADHIF*")
    assert list(d.hashes) == [smb.hash_murmur(kat["dayhoff_CADHIF*"])]
    h = _mh(smb, "hp", 3, scaled=1)
    h.add_protein("ANA")
    assert list(h.hashes) == [smb.hash_murmur(kat["hp_ANA"])]
    h = h.copy_and_clear(); h.add_protein("AN*")
    assert list(h.hashes) == [smb.hash_murmur(kat["hp_AN*"])]
    dna = smb.MinHash(0, 21, scaled=1)
    with pytest.raises(ValueError):
        dna.seq_to_hashes("ATGAGAGACGATAGACAGATGACC", is_protein=True)
    with pytest.raises(ValueError):                                     # InvalidHashFunction
        dna.add_protein("ATGAGAGACGATAGACAGATGACC")


def test_benchmark_sigs_and_2x2(smb, golden, golden_dir):     # test_sourmash_sketch.py:1340-1376
    info = golden["meta"]["protein_benchmarks"]
    prots = dict((n.split()[0], s) for n, s in _records(golden_dir / "ecoli.faa"))
    genes = dict((n.split()[0], s) for n, s in _records(golden_dir / "ecoli.genes.fna"))
    i = info["input_prot"]
    mh = smb.MinHash(i["num"], i["ksize"] // 3, is_protein=True, seed=i["seed"])
    mh.add_protein(prots[i["name"].split()[0]])
    assert np.array_equal(np.array(sorted(mh.hashes), dtype=np.uint64), golden["arrays"]["bench_input_prot"])
    assert mh.md5sum() == i["md5sum"]
    t = info["translate_prot"]
    mt = smb.MinHash(t["num"], t["ksize"] // 3, is_protein=True, seed=t["seed"])
    mt.add_sequence(genes[t["name"].split()[0]])
    assert np.array_equal(np.array(sorted(mt.hashes), dtype=np.uint64), golden["arrays"]["bench_translate_prot"])
    assert mt.md5sum() == t["md5sum"]
    # tests/test_sourmash_compute.py:810-860
    want = golden["meta"]["protein_2x2"]
    aa, tr = [], []
    for _, s in sorted(prots.items()):
        m = smb.MinHash(500, 7, is_protein=True); m.add_protein(s); aa.append(m)
    for _, s in sorted(genes.items()):          # by name: "gi|...:2801-3733" < "gi|...:337-2799"
        m = smb.MinHash(500, 7, is_protein=True); m.add_sequence(s); tr.append(m)
    assert round(aa[0].similarity(tr[0]), 3) == want["aa1_trans1"]
    assert round(aa[1].similarity(tr[0]), 3) == want["aa2_trans1"]
    assert round(aa[0].similarity(tr[1]), 3) == want["aa1_trans2"]
    assert round(aa[1].similarity(tr[1]), 3) == want["aa2_trans2"]


@pytest.mark.parametrize("moltype", ["protein", "dayhoff", "hp"])
@pytest.mark.parametrize("kaa", [1, 5, 10, 16, 19, 42])
def test_random_sequences_vs_oracle(smb, moltype, kaa):
    from sourmash_b200.synth import synth_genome
    rng = np.random.default_rng(kaa)
    g = synth_genome(20_000, seed=kaa, n_every=211)
    g[500:900] = np.frombuffer(bytes(g[500:900]).lower(), dtype=np.uint8)
    g[1000] = ord("R"); g[1001] = ord("n")
    dna = bytes(g)
    alphabet = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYXBZ*acdefghiklmnpqrstvwy", dtype=np.uint8)
    prot = bytes(alphabet[rng.integers(0, len(alphabet), size=7000)])
    for scaled, num, track in ((1, 0, False), (20, 0, True), (0, 300, True)):
        ms = _mh(smb, moltype, kaa, n=num, scaled=scaled, track_abundance=track)
        ms.add_sequence(dna)
        o = orc.OracleMinHash(scaled=scaled, ksize=3 * kaa, num=num, track_abundance=track)
        o.add_protein_family(dna, moltype, False)
        assert np.array_equal(np.array(sorted(ms.hashes), dtype=np.uint64), o.mins())
        if track:
            assert [ms.hashes[int(h)] for h in o.mins()] == o.abunds().tolist()
        mp = _mh(smb, moltype, kaa, n=num, scaled=scaled, track_abundance=track)
        mp.add_protein(prot)
        o = orc.OracleMinHash(scaled=scaled, ksize=3 * kaa, num=num, track_abundance=track)
        o.add_protein_family(prot, moltype, True)
        assert np.array_equal(np.array(sorted(mp.hashes), dtype=np.uint64), o.mins())
        if track:
            assert [mp.hashes[int(h)] for h in o.mins()] == o.abunds().tolist()
    assert _mh(smb, moltype, kaa, scaled=1).seq_to_hashes(dna[:3000], force=True, bad_kmers_as_zeroes=True) == \
        orc.seq_to_hashes_translate(dna[:3000], kaa, moltype, keep_zeros=True).tolist()


def test_signature_template_all_moltypes(smb, golden_dir):             # signature.rs:1041-1064
    p = smb.ComputeParameters(ksizes=[21, 31, 51], num_hashes=500, scaled=0)
    p.protein = True; p.dayhoff = True; p.hp = True; p.dna = True
    sig = smb.SourmashSignature.from_params(p)
    for _, s in _records(golden_dir / "ecoli.genes.fna"):
        sig.add_sequence(s.decode(), False)
    sk = list(sig.sketches())
    assert len(sk) == 12 and all(len(m) == 500 for m in sk)
    assert [m.moltype for m in sk[:4]] == ["protein", "dayhoff", "hp", "DNA"]
    p2 = smb.ComputeParameters(ksizes=[3, 6], num_hashes=10, scaled=0, dna=False)
    p2.protein = True
    sig2 = smb.SourmashSignature.from_params(p2)                       # signature.rs:1020-1039
    sig2.add_protein("AGY")
    assert [len(m) for m in sig2.sketches()] == [3, 2]


@pytest.mark.parametrize("moltype,input_is_protein", [("protein", True), ("dayhoff", True), ("hp", False), ("protein", False)])
def test_batched_protein_sketching(smb, moltype, input_is_protein, golden_dir):
    """smb_sketch_sequences_aa: many records, two ksizes, records grouped into sketches."""
    from sourmash_b200 import batch as B
    from sourmash_b200.synth import synth_genome
    rng = np.random.default_rng(7)
    if input_is_protein:
        alphabet = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY*X", dtype=np.uint8)
        recs = [bytes(alphabet[rng.integers(0, len(alphabet), size=n)]) for n in (5000, 3, 0, 1200, 257, 256, 9000)]
    else:
        recs = [bytes(synth_genome(n, seed=n, n_every=97 if n % 2 else 0)) for n in (30000, 20, 0, 7001, 768, 40000, 29)]
    owner = [0, 0, 1, 1, 2, 3, 3]
    offs = np.cumsum([0] + [len(r) for r in recs]).astype(np.uint64)
    seqs = np.frombuffer(b"".join(recs), dtype=np.uint8)
    ks = [7, 10]
    sset, nk = B.sketch_sequences(seqs, offs, ks, scaled=10, moltype=moltype, input_is_protein=input_is_protein,
                                  seq_to_sketch=np.array(owner, dtype=np.uint32), n_sketches=4, track_abundance=True)
    h, off, ab = sset.to_host(with_abunds=True)
    total = 0
    for s in range(4):
        for j, k in enumerate(ks):
            o = orc.OracleMinHash(scaled=10, ksize=3 * k, track_abundance=True)
            for r, ow in zip(recs, owner):
                if ow == s:
                    o.add_protein_family(r, moltype, input_is_protein)
                    span = k if input_is_protein else 3 * k
                    total += max(len(r) - span + 1, 0) * (1 if input_is_protein else 2)
            row = slice(int(off[s * 2 + j]), int(off[s * 2 + j + 1]))
            assert np.array_equal(h[row], o.mins()), (s, k)
            assert np.array_equal(ab[row], o.abunds()), (s, k)
    assert nk == total
