"""Host-side scenarios of the reference's tests/test_signature.py (cited by line) that need no
GPU: signatures as containers, naming, hashing, pickling, freezing, and the .sig JSON round trips
through the library's native reader / writer."""
import copy
import gzip
import os
import pickle

import pytest

import sourmash_b200 as smb
from sourmash_b200 import FrozenMinHash, MinHash, SourmashSignature
from sourmash_b200.signature import (FrozenSourmashSignature, load_one_signature_from_json,
                                     load_signatures_from_json, save_signatures_to_json)
from tests.conftest import GOLDEN


@pytest.fixture(params=[True, False])
def track_abundance(request):
    return request.param


def _one_hash(track_abundance, **kw):
    e = MinHash(kw.pop("n", 1), 20, track_abundance=track_abundance, **kw)
    e.add_hash(5)
    return e


def test_copy_and_frozen_copies(track_abundance):                      # :20-56
    e = _one_hash(track_abundance)
    assert copy.copy(e) == e
    sig = SourmashSignature(e, name="foo", filename="bar.fa")
    sig2 = copy.copy(sig)
    assert sig == sig2 and sig2.name == "foo" and sig2.filename == "bar.fa"
    fz = sig.to_frozen()
    assert isinstance(fz, FrozenSourmashSignature) and fz.copy() is fz and fz.to_frozen() is fz and fz == sig
    mut = fz.to_mutable()
    assert type(mut) is SourmashSignature and mut == sig and mut is not fz
    mut.name = "changed"
    assert fz.name == "foo"


def test_compare_ne(track_abundance):                                  # :72-110
    a = SourmashSignature(_one_hash(track_abundance), name="a")
    e2 = MinHash(1, 30, track_abundance=track_abundance)
    e2.add_hash(5)
    b = SourmashSignature(e2, name="a")
    assert a != b and b != a
    e3 = MinHash(1, 20, track_abundance=track_abundance)
    e3.add_hash(6)
    assert a != SourmashSignature(e3, name="a")


def test_hashable_str_and_names(track_abundance):                      # :113-135, :258-286
    sig = SourmashSignature(_one_hash(track_abundance))
    assert sig.md5sum() == "eae27d77ca20db309e056e3d2dcd7d69"
    assert len({sig, SourmashSignature(_one_hash(track_abundance))}) == 1
    assert repr(sig) == "SourmashSignature('', eae27d77)" and str(sig) == sig.md5sum()[:8]
    sig._name = "fizbar"
    assert repr(sig) == "SourmashSignature('fizbar', eae27d77)"
    empty = MinHash(1, 20, track_abundance=track_abundance)
    assert str(SourmashSignature(empty, name="foo")) == "foo"
    assert str(SourmashSignature(empty, filename="foo.txt")) == "foo.txt"
    assert str(SourmashSignature(empty, name="foo", filename="foo.txt")) == "foo"


def test_roundtrips(track_abundance):                                  # :138-220
    sig = SourmashSignature(_one_hash(track_abundance))
    s = save_signatures_to_json([sig])
    sig2, = load_signatures_from_json(s)
    assert sig2 == sig and not isinstance(sig, FrozenSourmashSignature) and isinstance(sig2, FrozenSourmashSignature)
    assert isinstance(sig.minhash, FrozenMinHash) and isinstance(sig2.minhash, FrozenMinHash)
    assert sig2.minhash.track_abundance == track_abundance and sig2.minhash.hashes == sig.minhash.hashes
    sig.minhash = sig.minhash.to_mutable()                            # :157-165
    assert isinstance(sig.to_frozen().minhash, FrozenMinHash)
    assert list(load_signatures_from_json(s, ksize="20"))[0] == sig   # :168-177 non-int ksize
    empty = SourmashSignature(MinHash(1, 20, track_abundance=track_abundance))      # :180-190
    back, = load_signatures_from_json(save_signatures_to_json([empty]))
    assert len(back.minhash) == 0 and back == empty
    sc = MinHash(0, 20, track_abundance=track_abundance, max_hash=10)                # :193-205
    sc.add_hash(5)
    back, = load_signatures_from_json(save_signatures_to_json([SourmashSignature(sc)]))
    assert back.minhash.scaled == sc.scaled and back.minhash._max_hash == 10
    sd = _one_hash(track_abundance, seed=10)                                          # :208-220
    back, = load_signatures_from_json(save_signatures_to_json([SourmashSignature(sd)]))
    assert back.minhash.seed == 10


def test_multisig_one_sig_and_minified(track_abundance, tmp_path):    # :289-399
    sig1 = SourmashSignature(MinHash(1, 20, track_abundance=track_abundance), name="foo")
    sig2 = SourmashSignature(MinHash(1, 25, track_abundance=track_abundance), name="bar baz")
    x = save_signatures_to_json([sig1, sig2])
    assert isinstance(x, bytes) and b"\n" not in x
    y = list(load_signatures_from_json(x))
    assert len(y) == 2 and sig1 in y and sig2 in y and sig1 != sig2
    assert {s.name for s in y} == {"foo", "bar baz"}
    with pytest.raises(ValueError):
        load_one_signature_from_json(save_signatures_to_json([]))
    assert load_one_signature_from_json(save_signatures_to_json([sig1])) == sig1
    with pytest.raises(ValueError):
        load_one_signature_from_json(x)
    assert load_one_signature_from_json(save_signatures_to_json([sig1], compression=5)) == sig1
    with open(tmp_path / "1.sig", "wb") as fp:                         # :378-385 binary file object
        assert save_signatures_to_json([sig1], fp) is None
    with open(tmp_path / "1.sig", "rb") as fp:
        assert list(load_signatures_from_json(fp)) == [sig1]
    missing = tmp_path / "dne.sig"                                     # :388-399
    with pytest.raises(Exception):
        list(load_signatures_from_json(missing, do_raise=True))
    assert list(load_signatures_from_json(missing)) == []
    # reference-written file: minified copy is smaller and loads to the same sketches (:354-362)
    path = os.path.join(GOLDEN, "genome-s10.fa.gz.sig")
    sigs = list(load_signatures_from_json(path))
    minified = save_signatures_to_json(sigs)
    assert len(minified) <= os.path.getsize(path) and b"\n" not in minified
    assert [s.md5sum() for s in load_signatures_from_json(gzip.compress(minified))] == [s.md5sum() for s in sigs]


def test_frozen_signature_rules(track_abundance):                     # :652-683
    ss = SourmashSignature(_one_hash(track_abundance), name="foo").to_frozen()
    with pytest.raises(ValueError):
        ss.name = "foo2"
    with pytest.raises(ValueError):
        ss.minhash = ss.minhash.copy_and_clear()
    with pytest.raises(ValueError):
        ss.filename = "x"
    with pytest.raises(ValueError):
        ss.add_sequence("ACGT" * 10)
    with pytest.raises(ValueError):
        ss.add_protein("MVKV")
    with ss.update() as ss2:
        ss2.name = "foo2"
    assert ss2.name == "foo2" and ss.name == "foo" and isinstance(ss2, FrozenSourmashSignature)


def test_pickle(track_abundance):
    sig = SourmashSignature(_one_hash(track_abundance), name="p", filename="q.fa")
    back = pickle.loads(pickle.dumps(sig))
    assert back == sig and back.name == "p" and back.filename == "q.fa"
    fz = pickle.loads(pickle.dumps(sig.to_frozen()))
    assert fz == sig
