"""CPU-only tests: the C-ABI library loads, exports every declared symbol, and the host-side
object model (sorted containers, num/scaled bookkeeping, md5, errors, pickling, JSON) behaves
like the reference -- checked against the oracle's KmerMinHash restatement.  No compute calls
(those need the GPU and live in test_gpu_*.py)."""
import os
import pickle

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

import oracle as orc
import sourmash_b200 as smb
from sourmash_b200 import FrozenMinHash, MinHash, SourmashSignature
from sourmash_b200._lowlevel import declared_symbols, lib


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) > 100
    for n in names:
        assert getattr(lib, n) is not None
    # the reference-ABI subset used by the Python object model is present
    for n in ("kmerminhash_new kmerminhash_add_sequence kmerminhash_seq_to_hashes kmerminhash_count_common "
              "kmerminhash_intersection_union_size kmerminhash_similarity kmerminhash_md5sum "
              "signature_first_mh signature_add_sequence sourmash_err_get_last_code hash_murmur "
              "computeparams_new smb_sketch_sequences smb_compare_jaccard smb_one_vs_many smb_gather").split():
        assert n in names


def test_no_gpu_fails_loudly_not_silently():
    if smb.batch.device_count() > 0:
        pytest.skip("GPU present")
    mh = MinHash(0, 21, scaled=10)
    with pytest.raises(smb.exceptions.CudaUnavailable, match="no CPU fallback"):
        mh.add_sequence("ACGT" * 20)
    a, b = MinHash(0, 21, scaled=10), MinHash(0, 21, scaled=10)
    a.add_many([1, 2, 3]); b.add_many([2, 3, 4])
    with pytest.raises(smb.exceptions.CudaUnavailable):
        a.count_common(b)
    with pytest.raises(smb.exceptions.CudaUnavailable):
        smb.hash_murmur("ACG")


def test_max_hash_for_scaled_matches_reference_formula(golden):
    kat = golden["meta"]["kat"]
    assert smb.batch.max_hash_for_scaled(100) == kat["max_hash_scaled_100"]
    assert smb.batch.max_hash_for_scaled(1000) == kat["max_hash_scaled_1000"]
    for s in (0, 1, 2, 93, 99, 1923, 4102, 10**6):
        assert smb.batch.max_hash_for_scaled(s) == orc.max_hash_for_scaled(s)
    mh = MinHash(0, 31, scaled=1000)
    assert mh._max_hash == kat["max_hash_scaled_1000"] and mh.scaled == 1000


def test_constructor_validation():
    with pytest.raises(ValueError):
        MinHash(0, 21)
    with pytest.raises(ValueError):
        MinHash(10, 21, scaled=100)
    with pytest.raises(ValueError):
        MinHash(0, 21, scaled=100, max_hash=5)
    p = MinHash(10, 7, is_protein=True)
    assert p.ksize == 7 and p.moltype == "protein" and not p.is_dna
    assert MinHash(0, 21, max_hash=18446744073709552).scaled == 1000


@given(st.lists(st.tuples(st.integers(0, 2**64 - 1), st.integers(0, 5)), max_size=80),
       st.sampled_from([(0, 1), (0, 1000), (5, 0), (20, 0), (0, 3)]), st.booleans())
@settings(max_examples=150, deadline=None)
def test_add_hash_semantics_match_oracle(items, params, track):
    num, scaled = params
    # scale hashes so that some pass a scaled filter
    mh = MinHash(num, 21, scaled=scaled, track_abundance=track)
    om = orc.OracleMinHash(scaled=scaled, ksize=21, num=num, track_abundance=track)
    for h, a in items:
        h = h if scaled <= 1 else h >> 1 if a % 2 else h // (scaled // 2 + 1)
        if track:
            mh.add_hash_with_abundance(h, a)
            om.add_hash_with_abundance(h, a)
        else:
            mh.add_hash(h)
            om.add_hash(h)
    assert mh._mins_array().tolist() == om.mins().tolist()
    if track:
        assert mh._abunds_array().tolist() == om.abunds().tolist()
    assert mh.md5sum() == om.md5sum()
    assert len(mh) == len(om)


@given(st.lists(st.integers(0, 500), max_size=60), st.lists(st.integers(0, 500), max_size=60),
       st.sampled_from([(0, 1), (10, 0)]), st.booleans(), st.booleans())
@settings(max_examples=100, deadline=None)
def test_merge_and_remove_match_oracle(xs, ys, params, ta, tb):
    num, scaled = params
    a, b = MinHash(num, 21, scaled=scaled, track_abundance=ta), MinHash(num, 21, scaled=scaled, track_abundance=tb)
    oa = orc.OracleMinHash(scaled=scaled, ksize=21, num=num, track_abundance=ta)
    ob = orc.OracleMinHash(scaled=scaled, ksize=21, num=num, track_abundance=tb)
    for x in xs:
        a.add_hash(x); oa.add_hash(x)
    for y in ys:
        b.add_hash(y); ob.add_hash(y)
    a.merge(b); oa.merge(ob)
    assert a._mins_array().tolist() == oa.mins().tolist()
    assert a.track_abundance == (ta and tb)
    if ta and tb:
        assert a._abunds_array().tolist() == oa.abunds().tolist()
    a.remove_many(ys[:5])
    for y in ys[:5]:
        oa.remove_hash(y)
    assert a._mins_array().tolist() == oa.mins().tolist()


def test_merge_kat_host_only(golden):
    # src/core/tests/minhash.rs:29-54 -- here with the k-mer hashes taken from the oracle
    kat = golden["meta"]["kat"]["merge_k10_num20"]
    a, b = MinHash(20, 10), MinHash(20, 10)
    for mh, seqs in ((a, kat["a"]), (b, kat["b"])):
        for s in seqs:
            mh.add_many(orc.seq_to_hashes(s, 10)[0])
    a.merge(b)
    assert list(a.hashes) == kat["merged"]


def test_incompatible_merge_raises_valueerror_with_reference_messages():
    a = MinHash(0, 21, scaled=10)
    with pytest.raises(ValueError, match="different ksizes cannot be compared"):
        a.merge(MinHash(0, 31, scaled=10))
    with pytest.raises(ValueError, match="mismatch in scaled; comparison fail"):
        a.merge(MinHash(0, 21, scaled=20))
    with pytest.raises(ValueError, match="mismatch in seed; comparison fail"):
        a.merge(MinHash(0, 21, scaled=10, seed=43))
    with pytest.raises(ValueError, match="DNA/prot minhashes cannot be compared"):
        a.merge(MinHash(0, 7, scaled=10, is_protein=True))
    assert not a.is_compatible(MinHash(0, 31, scaled=10))
    with pytest.raises(TypeError):
        a.intersection_and_union_size(MinHash(0, 31, scaled=10))
    with pytest.raises(TypeError):
        a.jaccard(MinHash(5, 21))


def test_hashes_view_set_abundances_and_flatten():
    mh = MinHash(0, 21, scaled=1, track_abundance=True)
    mh.set_abundances({10: 3, 5: 1, 7: 0, 20: 2})
    assert dict(mh.hashes) == {5: 1, 10: 3, 20: 2}
    mh.set_abundances({5: 4}, clear=False)
    assert mh.hashes[5] == 5 and mh.sum_abundances == 10
    with pytest.raises(ValueError):
        mh.set_abundances({1: -1})
    with pytest.raises(RuntimeError):
        mh.hashes[1] = 2
    flat = mh.flatten()
    assert not flat.track_abundance and list(flat.hashes) == [5, 10, 20]
    with pytest.raises(RuntimeError):
        flat.add_hash_with_abundance(1, 2)
    inflated = flat.inflate(mh)
    assert dict(inflated.hashes) == dict(mh.hashes)
    with pytest.raises(RuntimeError, match="MinHash is empty"):
        flat.track_abundance = True
    mh.track_abundance = False
    assert not mh.track_abundance


def test_downsample_rules():
    mh = MinHash(0, 21, scaled=2)
    mh.add_many([1, 2**62, 2**63 - 5, 2**63 + 5])           # last one is above max_hash for scaled=2
    assert len(mh) == 3
    d = mh.downsample(scaled=4)
    assert list(d.hashes) == [1, 2**62] and d.scaled == 4
    with pytest.raises(ValueError, match="lower than current sample scaled"):
        d.downsample(scaled=2)
    with pytest.raises(ValueError):
        mh.downsample(num=2)
    with pytest.raises(ValueError):
        mh.downsample()
    n = MinHash(4, 21)
    n.add_many([9, 3, 7, 1, 5])
    assert list(n.hashes) == [1, 3, 5, 7]
    assert list(n.downsample(num=2).hashes) == [1, 3]
    with pytest.raises(ValueError, match="higher than current sample num"):
        n.downsample(num=8)


def test_copy_pickle_frozen():
    mh = MinHash(0, 21, scaled=1, track_abundance=True)
    mh.set_abundances({3: 2, 9: 1})
    c = mh.copy()
    assert c == mh and c is not mh
    c.add_hash(11)
    assert c != mh
    f = mh.to_frozen()
    assert isinstance(f, FrozenMinHash) and f == mh and f.copy() is f
    for name, args in (("add_hash", (1,)), ("add_many", ([1],)), ("clear", ()), ("merge", (mh,)),
                       ("add_sequence", ("ACGT",)), ("remove_many", ([3],)), ("set_abundances", ({1: 1},))):
        with pytest.raises(TypeError):
            getattr(f, name)(*args)
    m = f.to_mutable()
    m.add_hash(4)
    assert len(m) == 3 and len(f) == 2
    g = pickle.loads(pickle.dumps(f))
    assert isinstance(g, FrozenMinHash) and g == f
    assert pickle.loads(pickle.dumps(mh)) == mh


def test_signature_container_and_json_roundtrip(golden, tmp_path):
    a = MinHash(0, 31, scaled=1000)
    a.add_many(golden["arrays"]["s47"])
    sig = SourmashSignature(a, name="forty-seven", filename="47.fa")
    assert sig.md5sum() == golden["meta"]["s47_md5"] and str(sig) == "forty-seven" and len(sig) == 1
    assert isinstance(sig.minhash, FrozenMinHash) and len(sig.minhash) == 5177
    b = MinHash(500, 31, track_abundance=True)
    b.set_abundances({int(h): i % 3 + 1 for i, h in enumerate(golden["arrays"]["demo0"])})
    sig2 = SourmashSignature(b, name="demo")
    text = smb.save_signatures_to_json([sig, sig2])
    back = list(smb.load_signatures_from_json(text))
    assert back[0] == sig and back[1] == sig2 and back[0].name == "forty-seven"
    assert back[1].minhash.track_abundance and back[1].minhash == b
    assert isinstance(text, bytes)                   # like the reference (ffi.string of the saved buffer)
    p = tmp_path / "x.sig"
    p.write_bytes(text)
    assert [s.md5sum() for s in smb.load_signatures(str(p))] == [sig.md5sum(), sig2.md5sum()]
    assert len(list(smb.load_signatures_from_json(text, ksize=21))) == 0


def test_compute_parameters_template():
    p = smb.ComputeParameters(ksizes=[21, 31, 51], scaled=1000, num_hashes=0)
    assert p.ksizes == [21, 31, 51] and p.scaled == 1000 and p.dna and not p.protein and p.seed == 42
    sig = SourmashSignature.from_params(p)
    assert len(sig) == 3 and [m.ksize for m in sig.sketches()] == [21, 31, 51]
    assert all(m.scaled == 1000 and m.num == 0 for m in sig.sketches())


def test_search_scoring_protocol():
    from sourmash_b200.search import (JaccardSearchBestOnly, calc_threshold_from_bp, make_jaccard_search_query)
    assert calc_threshold_from_bp(0, 1000, 50) == (0.0, 0)
    assert calc_threshold_from_bp(5000, 1000, 50) == (0.1, 5.0)
    with pytest.raises(ValueError):
        calc_threshold_from_bp(100000, 1000, 50)
    s = make_jaccard_search_query(do_containment=True, threshold=0.2)
    assert s.score_fn(10, 5, 20, 25) == 0.5 and s.passes(0.2) and not s.passes(0.1) and not s.passes(0)
    j = make_jaccard_search_query(threshold=0)
    assert j.score_fn(10, 5, 20, 25) == 0.2 and not j.passes(0.0)
    m = make_jaccard_search_query(do_max_containment=True)
    assert m.score_fn(10, 5, 20, 25) == 0.5
    b = JaccardSearchBestOnly(s.search_type, 0.1)
    b.collect(0.7, None)
    assert b.threshold == 0.7


def test_load_reference_written_sig_files(golden):
    import os
    from tests.conftest import GOLDEN
    sigs = list(smb.load_signatures(os.path.join(GOLDEN, "47.fa.sig")))
    assert len(sigs) == 1
    mh = sigs[0].minhash
    assert (mh.ksize, mh.scaled, mh.num, len(mh)) == (31, 1000, 0, 5177)
    assert sigs[0].md5sum() == golden["meta"]["s47_md5"] == "09a08691ce52952152f0e866a59f6261"
    assert mh._mins_array().tolist() == golden["arrays"]["s47"].tolist()
    # multi-sketch file with DNA and protein sketches (num=500)
    s10 = list(smb.load_signatures(os.path.join(GOLDEN, "genome-s10.fa.gz.sig")))
    dna = [s for s in s10 if s.minhash.is_dna]
    assert {s.minhash.ksize for s in dna} == {21, 30} and len(s10) > len(dna)
    for s in dna:
        assert s.md5sum() == golden["meta"]["genome_s10"][str(s.minhash.ksize)]["md5sum"]
    prot = [s for s in s10 if s.minhash.moltype == "protein"]
    assert prot and all(len(s.minhash) == 500 for s in prot)
    only21 = list(smb.load_signatures_from_json(os.path.join(GOLDEN, "genome-s10.fa.gz.sig"), ksize=21,
                                                select_moltype="dna"))
    assert len(only21) == 1
    # round trip through our writer keeps identity
    again = list(smb.load_signatures_from_json(smb.save_signatures_to_json(s10)))
    assert [a.md5sum() for a in again] == [a.md5sum() for a in s10]


def test_residue_encodings_and_protein_guards(golden):
    """Scalar parts of the protein path that need no GPU (encodings.rs:298-343)."""
    from sourmash_b200 import translate_codon
    from sourmash_b200._lowlevel import lib
    for codon, aa in golden["meta"]["kat"]["translate_codon"].items():
        assert translate_codon(codon) == aa
    assert translate_codon("tct") == "X" and translate_codon("TCN") == "S" and translate_codon("TTN") == "X"
    assert translate_codon("ATG") == "M" and translate_codon("TGA") == "*" and translate_codon("NNN") == "X"
    for bad in ("", "TCTA"):
        with pytest.raises(ValueError, match="Codon is invalid length"):
            translate_codon(bad)
    for aa in b"ACDEFGHIKLMNPQRSTVWY*XBZ":
        assert lib.sourmash_aa_to_dayhoff(bytes([aa])) == bytes([orc.lib.orc_aa_to_dayhoff(aa)])
        assert lib.sourmash_aa_to_hp(bytes([aa])) == bytes([orc.lib.orc_aa_to_hp(aa)])
    all_codons = ["".join(c) for c in __import__("itertools").product("ACGTNR", repeat=3)]
    for c in all_codons:
        assert translate_codon(c) == orc.translate_codon(c), c
    # shorter than k: nothing happens, even without a GPU and even on a DNA sketch (max_index == 0)
    smb.MinHash(10, 9, is_protein=True).add_protein("AG")
    smb.MinHash(0, 31, scaled=1).add_protein("AG")
    assert smb.MinHash(0, 2, dayhoff=True, scaled=1).seq_to_hashes("ACTGA") == []
    with pytest.raises(ValueError):
        smb.MinHash(0, 21, scaled=1).seq_to_hashes("ATGAGAGACGATAGACAGATGACC", is_protein=True)


def test_gather_and_prefetch_csv_layout():
    """CSV writers: the reference's column order (search.py:364-388, 480-523), None -> empty cell."""
    import csv
    import io
    from sourmash_b200.gather import (CI_COLUMNS, GATHER_COLUMNS, PREFETCH_COLUMNS, GatherRow, write_gather_csv,
                                      write_prefetch_csv)
    rows = [GatherRow(row=3, intersect_bp=5000, f_orig_query=0.5, f_match=0.25, f_unique_to_query=0.5,
                      f_unique_weighted=0.5, name="g3", md5="abc", gather_result_rank=0, remaining_bp=5000,
                      query_md5="12345678", query_bp=10000, ksize=31, scaled=1000, query_n_hashes=10,
                      query_containment_ani=0.97, total_weighted_hashes=10, sum_weighted_found=5),
            GatherRow(row=1, gather_result_rank=1, average_abund=2.5, median_abund=2.0, std_abund=0.5,
                      query_abundance=True, n_unique_weighted_found=7)]
    buf = io.StringIO()
    write_gather_csv(rows, buf)
    got = list(csv.reader(io.StringIO(buf.getvalue())))
    assert got[0] == GATHER_COLUMNS and len(got) == 3
    rec = dict(zip(got[0], got[1]))
    assert rec["intersect_bp"] == "5000" and rec["name"] == "g3" and rec["average_abund"] == "" and \
        rec["query_containment_ani"] == "0.97" and rec["match_containment_ani"] == ""
    assert dict(zip(got[0], got[2]))["median_abund"] == "2.0"
    buf = io.StringIO()
    write_gather_csv(rows, buf, estimate_ani_ci=True)
    assert next(csv.reader(io.StringIO(buf.getvalue()))) == GATHER_COLUMNS + CI_COLUMNS
    buf = io.StringIO()
    write_prefetch_csv([{"row": 0, "intersect_bp": 100, "jaccard": 0.5, "match_name": "m", "ksize": 21}], buf)
    got = list(csv.reader(io.StringIO(buf.getvalue())))
    assert got[0] == PREFETCH_COLUMNS and got[1][got[0].index("jaccard")] == "0.5" and "row" not in got[0]


def test_cli_plugin_protocol_and_argument_wiring(tmp_path):
    """The sourmash.cli_script plugin classes (reference protocol: src/sourmash/plugins.py:91-186):
    command / description attributes, parser construction, -q/-d, param strings; a command run
    without a GPU fails loudly instead of falling back."""
    import argparse
    from sourmash_b200 import plugin
    from sourmash_b200.exceptions import SourmashError
    parser, objs = plugin.build_parser()
    assert sorted(objs) == ["b200compare", "b200gather", "b200prefetch", "b200search", "b200sketch"]
    for cls in plugin.COMMANDS:
        assert cls.command and cls.description and issubclass(cls, plugin.CommandLinePlugin)
        sp = argparse.ArgumentParser()
        cls(sp)                                             # what sourmash's add_cli_scripts does
        assert {"quiet", "debug"} <= {a.dest for a in sp._actions}
    a = parser.parse_args(["b200sketch", "x.fa", "y.fa.gz", "-p", "k=21,k=31,scaled=1000,abund", "-o", "o.sig", "-q"])
    assert a.filenames == ["x.fa", "y.fa.gz"] and a.quiet and a.moltype == "dna"
    P = plugin.parse_param_string(a.param_string)
    assert P == {"ksizes": [21, 31], "scaled": 1000, "num": None, "seed": 42, "track_abundance": True}
    assert plugin.parse_param_string("k=7,num=500,seed=3,noabund")["num"] == 500
    with pytest.raises(ValueError):
        plugin.parse_param_string("k=31,scaled=100,num=5")
    with pytest.raises(ValueError):
        plugin.parse_param_string("k=31,bogus=1")
    a = parser.parse_args(["b200gather", "q.sig", "a.sig", "b.sig", "-k", "31", "--threshold-bp", "0", "-o", "g.csv"])
    assert a.query == "q.sig" and a.databases == ["a.sig", "b.sig"] and a.threshold_bp == 0 and a.ksize == 31
    a = parser.parse_args(["b200search", "q.sig", "db.zip", "--containment", "-t", "0.1", "-n", "0", "-o", "s.csv"])
    assert a.containment and not a.max_containment and a.threshold == 0.1 and a.num_results == 0 and a.databases == ["db.zip"]
    # pyproject registers exactly these classes under the reference's entry-point group
    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pyproject.toml")).read()
    assert '[project.entry-points."sourmash.cli_script"]' in text
    for cls in plugin.COMMANDS:
        assert f'{cls.command} = "sourmash_b200.plugin:{cls.__name__}"' in text
    if smb.batch.device_count() == 0:                       # no silent CPU fallback behind the CLI either
        fa = tmp_path / "g.fa"
        fa.write_text(">r\\n" + "ACGT" * 30 + "\\n")
        with pytest.raises(SourmashError):
            plugin.main(["b200sketch", str(fa), "-p", "k=21,scaled=1", "-o", str(tmp_path / "o.sig"), "-q"])
