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


class _FakeSession:
    "batch.GatherSession with Python sets (same begin / peek / intersect / apply contract)."

    def __init__(self, query, db, min_count=1):
        self.rows = [set(int(x) for x in r) for r in db.rows()]
        self.remaining_set = set(int(x) for x in query)
        self.counts = [len(self.remaining_set & r) for r in self.rows]
        self.counts = [c if c >= min_count else 0 for c in self.counts]
        self.remaining = len(self.remaining_set)

    def peek(self):
        best = max(self.counts) if self.counts else 0
        return (best, self.counts.index(best)) if best else (0, 0)

    def intersect(self, row):
        return np.array(sorted(self.remaining_set & self.rows[row]), dtype=np.uint64)

    def apply(self, isect):
        s = set(int(x) for x in isect)
        self.counts = [c - len(s & r) if c else 0 for c, r in zip(self.counts, self.rows)]
        self.remaining_set -= s
        self.remaining = len(self.remaining_set)
        return self.remaining


def test_gather_report_on_the_reference_fixture(cpu_kernels, monkeypatch, golden):
    """gather_databases / prefetch_database report logic on the 12-genome fixture
    (tests/test_index_protocol.py:1057-1097), the session replaced by Python sets."""
    import glob
    from sourmash_b200 import gather as G
    from sourmash_b200.sigset import SignatureSet
    monkeypatch.setattr(B, "GatherSession", _FakeSession)
    d = os.path.join(GOLDEN, "gather")
    query = smb.signature.load_one_signature_from_json(os.path.join(d, "combined.sig"), ksize=21)
    ss = SignatureSet.from_files(sorted(glob.glob(os.path.join(d, "GCF*.sig"))))
    rows = ss.select(ksize=21)
    db = _FakeSet([ss.row(i) for i in rows])
    names = [ss.name(i) for i in rows]
    report = G.gather_databases(query.minhash, db, names=names, md5s=[ss.md5sum(i) for i in rows])
    want = golden["meta"]["gather_k21_expected"]
    assert [[g.name.split()[0], g.unique_intersect_bp // g.scaled] for g in report] == want
    q = set(query.minhash.hashes)
    covered = 0
    for rank, g in enumerate(report):
        m = set(int(x) for x in ss.row(rows[g.row]))
        assert g.gather_result_rank == rank and g.intersect_bp == len(q & m) * g.scaled
        assert g.f_orig_query == len(q & m) / len(q) and g.f_unique_to_query == (g.unique_intersect_bp // g.scaled) / len(q)
        assert g.f_unique_weighted == g.f_unique_to_query and g.average_abund is None and not g.query_abundance
        covered += g.unique_intersect_bp // g.scaled
        assert g.remaining_bp == (len(q) - covered) * g.scaled and g.sum_weighted_found == covered
        assert 0 < g.f_match <= g.f_match_orig <= 1
        assert g.query_containment_ani is None or 0.8 < g.query_containment_ani <= 1.0
    pre = G.prefetch_database(query.minhash, db, 0, names=names)
    assert len(pre) == 12 and [p["row"] for p in pre] == list(range(12))
    by_row = {g.row: g for g in report}
    for p in pre:
        assert p["intersect_bp"] == by_row[p["row"]].intersect_bp and p["f_query_match"] == by_row[p["row"]].f_match_orig
    assert [p["row"] for p in G.prefetch_database(query.minhash, db, 100000, names=names)] == \
        [p["row"] for p in pre if p["intersect_bp"] >= 100000]
    with_thr = G.gather_databases(query.minhash, db, threshold_bp=50000, names=names)
    scaled = query.minhash.scaled
    assert scaled == 10000
    assert [[g.name.split()[0], g.unique_intersect_bp // g.scaled] for g in with_thr] == [w for w in want if w[1] * scaled >= 50000]
    assert len(with_thr) == 11


@pytest.mark.parametrize("track", [False, True])
def test_gather_noident_keeps_denominators(cpu_kernels, monkeypatch, track):
    """noident hashes (search.py:803-818,910-944,553-590): searched query = query - noident, while
    query_bp / query_n_hashes / f_orig_query / f_unique_to_query / total_weighted_hashes keep counting
    them and remaining_bp carries them."""
    import glob
    from sourmash_b200 import gather as G
    from sourmash_b200.sigset import SignatureSet
    monkeypatch.setattr(B, "GatherSession", _FakeSession)
    d = os.path.join(GOLDEN, "gather")
    base = smb.signature.load_one_signature_from_json(os.path.join(d, "combined.sig"), ksize=21).minhash
    ss = SignatureSet.from_files(sorted(glob.glob(os.path.join(d, "GCF*.sig"))))
    rows = ss.select(ksize=21)
    db = _FakeSet([ss.row(i) for i in rows])
    in_db = set(int(x) for i in rows for x in ss.row(i))
    extra = [h for h in range(1000, 1060) if h not in in_db and h not in set(base.hashes)]
    assert len(extra) == 60
    hashes = sorted(base.hashes)
    abund = {h: 1 + (h % 7) for h in hashes + extra}
    plain = smb.MinHash(0, 21, scaled=base.scaled, track_abundance=track)
    padded = smb.MinHash(0, 21, scaled=base.scaled, track_abundance=track)
    if track:
        plain.set_abundances({h: abund[h] for h in hashes})
        padded.set_abundances(abund)
    else:
        plain.add_many(hashes)
        padded.add_many(hashes + extra)
    ref = G.gather_databases(plain, db)
    got = G.gather_databases(padded, db, noident_hashes=extra)
    n, n_pad = len(hashes), len(hashes) + 60
    w = sum(abund[h] for h in hashes) if track else n
    w_pad = w + (sum(abund[h] for h in extra) if track else 60)
    assert len(got) == len(ref) == 12
    for a, b in zip(got, ref):
        assert (a.row, a.intersect_bp, a.unique_intersect_bp, a.f_match, a.f_match_orig) == \
            (b.row, b.intersect_bp, b.unique_intersect_bp, b.f_match, b.f_match_orig)
        assert a.query_n_hashes == n_pad and a.query_bp == n_pad * a.scaled and b.query_n_hashes == n
        assert a.f_orig_query == (b.intersect_bp // b.scaled) / n_pad
        assert a.f_unique_to_query == (b.unique_intersect_bp // b.scaled) / n_pad
        assert a.remaining_bp == b.remaining_bp + 60 * a.scaled
        assert a.total_weighted_hashes == w_pad and b.total_weighted_hashes == w
        assert a.sum_weighted_found == b.sum_weighted_found                 # noident is never "found"
        if track:
            assert a.n_unique_weighted_found == b.n_unique_weighted_found
            assert a.f_unique_weighted == b.n_unique_weighted_found / w_pad
            assert (a.average_abund, a.median_abund, a.std_abund) == (b.average_abund, b.median_abund, b.std_abund)
        else:
            assert a.f_unique_weighted == a.f_unique_to_query
    # without the noident argument the padding is searched like any other hash: same picks, other threshold base
    assert [g.row for g in G.gather_databases(padded, db)] == [g.row for g in ref]
    with pytest.raises(KeyError):
        G.gather_databases(plain, db, noident_hashes=extra)                # not part of the query


def test_search_database_matches_index_search(cpu_kernels, three):
    """gather.search_database (one launch over a SketchSet) == LinearIndex.search object by object:
    scores, order, thresholds, best-only, one entry per md5; SearchResult CSV columns (search.py:283-355)."""
    import csv
    import glob
    import io
    from sourmash_b200 import distance_utils as DU
    from sourmash_b200 import gather as G
    from sourmash_b200.sigset import SignatureSet
    d = os.path.join(GOLDEN, "gather")
    query = smb.signature.load_one_signature_from_json(os.path.join(d, "combined.sig"), ksize=21)
    paths = sorted(glob.glob(os.path.join(d, "GCF*.sig")))
    ss = SignatureSet.from_files(paths + paths[:2])                    # two duplicates: one entry per md5
    rows = ss.select(ksize=21)
    db = _FakeSet([ss.row(i) for i in rows])
    meta = dict(names=[ss.name(i) for i in rows], md5s=[ss.md5sum(i) for i in rows],
                filenames=[ss.filename(i) for i in rows], query_name=query.name, query_filename=query.filename)
    lin = LinearIndex(ss.signatures(rows))
    qset = set(query.minhash.hashes)
    for kw in ({}, {"do_containment": True}, {"do_max_containment": True}):
        for thr in (0.0, 0.05, 0.2):
            got = G.search_database(query.minhash, db, threshold=thr, **kw, **meta)
            ref, seen = [], set()
            for r in lin.search(query, threshold=thr, **kw):
                if r.signature.md5sum() not in seen:
                    seen.add(r.signature.md5sum())
                    ref.append(r)
            assert [(g["similarity"], g["md5"]) for g in got] == [(r.score, r.signature.md5sum()) for r in ref], (kw, thr)
            assert len({g["md5"] for g in got}) == len(got) <= 12
            assert all(a["similarity"] >= b["similarity"] for a, b in zip(got, got[1:]))
    cont = G.search_database(query.minhash, db, threshold=0.0, do_containment=True, estimate_ani_ci=True, **meta)
    assert len(cont) == 12
    for g in cont:
        m = set(int(x) for x in ss.row(rows[g["row"]]))
        assert g["similarity"] == len(qset & m) / len(qset)
        want = DU.containment_to_distance(g["similarity"], 21, query.minhash.scaled,
                                          n_unique_kmers=len(qset) * query.minhash.scaled, estimate_ci=True)
        assert g["query_md5"] == query.md5sum()[:8] and g["name"] == meta["names"][g["row"]]
        assert g["ani"] in (None, want.ani) and (g["ani"] is None or (g["ani_low"] <= g["ani"] <= g["ani_high"]))
    jac = G.search_database(query.minhash, db, threshold=0.0, estimate_ani_ci=True, **meta)
    assert all("ani_low" not in g for g in jac)                            # no interval for Jaccard searches
    best = G.search_database(query.minhash, db, threshold=0.0, do_containment=True, best_only=True, **meta)
    assert best[0]["md5"] == cont[0]["md5"] and best[0]["similarity"] == max(g["similarity"] for g in cont)
    buf = io.StringIO()
    G.write_search_csv(cont, buf, estimate_ani_ci=True)
    table = list(csv.reader(io.StringIO(buf.getvalue())))
    assert table[0] == G.SEARCH_COLUMNS + G.SEARCH_CI_COLUMNS and len(table) == 13
    assert table[1][0] == str(cont[0]["similarity"]) and table[1][1] == cont[0]["md5"]
    with pytest.raises(TypeError):
        G.search_database(query.minhash, db, do_containment=True, do_max_containment=True)
    ss2, ss47, ss63 = three
    small = _FakeSet([s.minhash._mins_array() for s in (ss2, ss47, ss63)])
    sr = G.search_database(ss47.minhash, small, threshold=0.1, md5s=[s.md5sum() for s in three])    # test_index_protocol.py:206-230
    assert [g["md5"] for g in sr] == [ss47.md5sum(), ss63.md5sum()] and sr[0]["similarity"] == 1.0 and round(sr[1]["similarity"], 2) == 0.32


def test_find_refuses_incompatible_subjects(cpu_kernels, three):
    """Index.find scores subject by subject and `intersection_and_union_size` raises
    TypeError("incompatible MinHash objects") when ksize / molecule / seed differ from the query's
    (reference minhash.py:649-654, index/__init__.py:151-158); matches of earlier subjects are
    yielded first.  Round-1 advisor finding: a k=21 query 'matched' a k=31 subject."""
    ss2, ss47, ss63 = three
    k21 = smb.MinHash(0, 21, scaled=1000)
    k21.add_many(ss47.minhash.hashes)                                   # the same hashes under another ksize
    q21 = smb.SourmashSignature(k21, name="k21")
    lidx = LinearIndex([ss47, ss63])
    with pytest.raises(TypeError, match="incompatible MinHash objects"):
        lidx.search(q21, threshold=0.1)
    with pytest.raises(TypeError, match="incompatible MinHash objects"):
        list(lidx.prefetch(q21, threshold_bp=0))
    with pytest.raises(TypeError, match="incompatible MinHash objects"):
        lidx.counter_gather(q21, threshold_bp=0)
    seeded = smb.MinHash(0, 31, scaled=1000, seed=43)
    seeded.add_many(ss47.minhash.hashes)
    with pytest.raises(TypeError, match="incompatible MinHash objects"):
        lidx.search(smb.SourmashSignature(seeded, name="seed43"), threshold=0.1)
    # a mixed collection: subjects in front of the first incompatible one are still reported
    mixed = LinearIndex([ss47, smb.SourmashSignature(k21, name="k21"), ss63])
    from sourmash_b200.search import make_jaccard_search_query
    it = mixed.find(make_jaccard_search_query(threshold=0.1), ss47)
    first = next(it)
    assert first.signature.minhash == ss47.minhash and first.score == 1.0
    with pytest.raises(TypeError, match="incompatible MinHash objects"):
        next(it)
    # compatible queries are unaffected
    assert len(lidx.search(ss47, threshold=0.1)) == 2
