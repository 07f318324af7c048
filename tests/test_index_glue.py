"""LinearIndex search / prefetch / best_containment host logic without a GPU: the generic Index
scenarios of the reference's protocol suite (tests/test_index_protocol.py:206-520, cited) on its
own 2.fa / 47.fa / 63.fa fixtures, with the one-vs-many kernel replaced by the oracle (the GPU
counterparts are tests/test_gpu_api.py::test_index_search_prefetch and the CounterGather port)."""
import os

import numpy as np
import pytest

import oracle as orc
import sourmash_b200 as smb
from sourmash_b200 import batch as B
from sourmash_b200.index import LinearIndex
from tests.conftest import GOLDEN


class _FakeSet:
    def __init__(self, rows):
        self._rows = [np.asarray(r, dtype=np.uint64) for r in rows]

    @classmethod
    def from_rows(cls, rows, abund_rows=None):
        return cls(rows)

    def __len__(self):
        return len(self._rows)

    def rows(self):
        return self._rows

    def sizes(self):
        return np.array([len(r) for r in self._rows], dtype=np.int64)

    def downsample(self, max_hash):
        return _FakeSet([r[r <= np.uint64(max_hash)] for r in self._rows])


@pytest.fixture
def cpu_kernels(monkeypatch):
    monkeypatch.setattr(B, "SketchSet", _FakeSet)

    def one_vs_many(q, db):
        h, off = orc.to_csr(db.rows())
        return orc.one_vs_many(np.asarray(q, dtype=np.uint64), h, off).astype(np.uint32)
    monkeypatch.setattr(B, "one_vs_many", one_vs_many)

    def pairwise_common(a, b=None, num=0, want_usize=False):
        out = np.zeros((len(a), len(b)), dtype=np.uint32)
        us = np.zeros((len(a), len(b)), dtype=np.uint32)
        for i, x in enumerate(a.rows()):
            for j, y in enumerate(b.rows()):
                c, u = orc.intersection_size(x, y, num=num)
                out[i, j], us[i, j] = c, u
        return (out, us) if want_usize else out
    monkeypatch.setattr(B, "pairwise_common", pairwise_common)


@pytest.fixture
def three():
    ss2 = smb.signature.load_one_signature_from_json(os.path.join(GOLDEN, "2.fa.sig"), ksize=31)
    ss47 = smb.signature.load_one_signature_from_json(os.path.join(GOLDEN, "47.fa.sig"))
    ss63 = smb.signature.load_one_signature_from_json(os.path.join(GOLDEN, "63.fa.sig"))
    return ss2, ss47, ss63


@pytest.fixture
def index_obj(three):
    lidx = LinearIndex()
    for ss in three:
        lidx.insert(ss)
    return lidx


def test_search(cpu_kernels, index_obj, three):                         # :206-270
    ss2, ss47, ss63 = three
    sr = index_obj.search(ss2, threshold=1.0)
    assert len(sr) == 1 and sr[0].signature.minhash == ss2.minhash and sr[0].score == 1.0
    sr = index_obj.search(ss47, threshold=0.1)
    assert [s.signature.minhash for s in sr] == [ss47.minhash, ss63.minhash]
    assert sr[0].score == 1.0 and round(sr[1].score, 2) == 0.32
    sr = index_obj.search(ss63, threshold=0.1)
    assert [s.signature.minhash for s in sr] == [ss63.minhash, ss47.minhash] and round(sr[1].score, 2) == 0.32
    sr = index_obj.search(ss63, threshold=0.8)
    assert len(sr) == 1 and sr[0].signature.minhash == ss63.minhash and sr[0].score == 1.0
    sr = index_obj.search(ss63, do_containment=True, threshold=0.1)
    assert [s.signature.minhash for s in sr] == [ss63.minhash, ss47.minhash] and round(sr[1].score, 2) == 0.48
    with pytest.raises(TypeError):
        index_obj.search(ss2)                                           # threshold is required (:202-208)


def test_container_protocol(index_obj, three):                          # :272-312
    md5s = {ss.md5sum() for ss in index_obj.signatures()}
    assert md5s == {ss.md5sum() for ss in three} and len(index_obj) == 3 and bool(index_obj)
    assert {ss.md5sum() for ss, _ in index_obj.signatures_with_location()} == md5s
    assert not LinearIndex() and len(LinearIndex()) == 0


def test_prefetch_and_best_containment(cpu_kernels, index_obj, three):  # :396-520
    ss2, ss47, ss63 = three
    res = list(index_obj.prefetch(ss2, threshold_bp=0))
    assert len(res) == 1 and res[0].signature.minhash == ss2.minhash
    res = list(index_obj.prefetch(ss47, threshold_bp=0))
    assert [r.signature.minhash for r in res] == [ss47.minhash, ss63.minhash]
    for q in (ss2, ss47):
        m = index_obj.best_containment(q)
        assert m and m.score == 1.0 and m.signature.minhash == q.minhash
    mins = sorted(ss2.minhash.hashes)
    new_mh = ss2.minhash.copy_and_clear()
    with pytest.raises(ValueError):                                     # empty query
        index_obj.best_containment(smb.SourmashSignature(new_mh))
    new_mh.add_hash(mins.pop())
    score, match, _ = index_obj.best_containment(smb.SourmashSignature(new_mh))
    assert score == 1.0 and match.minhash == ss2.minhash
    with pytest.raises(ValueError):                                     # 1 hash cannot reach 5000 bp at scaled=1000
        index_obj.best_containment(smb.SourmashSignature(new_mh), threshold_bp=5000)
    for _ in range(3):
        new_mh.add_hash(mins.pop())
    score, match, _ = index_obj.best_containment(smb.SourmashSignature(new_mh))
    assert len(new_mh) == 4 and score == 1.0 and match.minhash == ss2.minhash
    with pytest.raises(ValueError):
        index_obj.best_containment(smb.SourmashSignature(new_mh), threshold_bp=5000)
    for _ in range(21):                                                 # 25 hashes in total (:479-520)
        new_mh.add_hash(mins.pop())
    for bp in (None, 5000):
        score, match, _ = index_obj.best_containment(smb.SourmashSignature(new_mh), threshold_bp=bp)
        assert score == 1.0 and match.minhash == ss2.minhash
    with pytest.raises(ValueError):
        list(LinearIndex().prefetch(ss2, threshold_bp=0))              # "no signatures to search"
