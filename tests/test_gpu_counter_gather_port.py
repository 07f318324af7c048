"""CounterGather / gather conformance: the scenarios of the reference's protocol suite
(tests/test_index_protocol.py:711-1310, each cited) run against sourmash_b200.index.CounterGather,
whose counting and decrementing happen in one-vs-many kernel launches, plus the 12-genome real-data
gather (:1057-1097) through CounterGather, LinearIndex and the batched gather_databases report."""
import glob
import os

import pytest

from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def smb():
    import sourmash_b200
    assert sourmash_b200.batch.device_count() > 0
    return sourmash_b200


def _sig(smb, values, name, *, scaled=1, track=False, like=None):
    mh = smb.MinHash(0, 31, scaled=scaled, track_abundance=track)
    mh.add_many(list(values))
    return smb.SourmashSignature(mh, name=name)


def _consume_all(query_mh, counter, threshold_bp=0):                  # test_index_protocol.py:742-763
    results, last = [], None
    query_mh = query_mh.to_mutable()
    while True:
        result = counter.peek(query_mh, threshold_bp=threshold_bp)
        if not result:
            break
        sr, intersect_mh = result
        if last:
            assert len(intersect_mh) <= last
        last = len(intersect_mh)
        counter.consume(intersect_mh)
        query_mh.remove_many(intersect_mh.hashes)
        results.append((sr, len(intersect_mh)))
    return results


def _names_sizes(results):
    return [[sr.signature.name.split()[0], n] for sr, n in results]


THREE = [["match1", 10], ["match2", 5], ["match3", 2]]


@pytest.mark.parametrize("ranges,scaleds,threshold_bp,expected", [
    ([range(0, 10), range(10, 15), range(15, 17)], (1, 1, 1), 0, THREE),          # :766-804 gather_1
    ([range(0, 10), range(7, 15), range(13, 17)], (1, 1, 1), 0, THREE),           # :807-848 gather_1_b
    ([range(0, 10), range(7, 15), range(13, 17)], (1, 1, 1), 3, THREE[:2]),       # :851-890 gather_1_c
    ([range(0, 10), range(7, 15), range(13, 17)], (10, 20, 30), 0, THREE),        # :893-930 gather_1_d
])
def test_counter_gather_contrived(smb, ranges, scaleds, threshold_bp, expected):
    from sourmash_b200.index import CounterGather
    query = _sig(smb, range(0, 20), "query")
    counter = CounterGather(query)
    for i, (r, s) in enumerate(zip(ranges, scaleds)):
        counter.add(_sig(smb, r, f"match{i + 1}", scaled=s))
    if scaleds == (1, 1, 1) and threshold_bp == 0:
        assert len(list(counter.signatures())) == 3                               # :711-739
    assert _names_sizes(_consume_all(query.minhash, counter, threshold_bp)) == expected


def test_counter_gather_diff_scaled_query_and_abundance(smb):
    from sourmash_b200.index import CounterGather
    ranges = [range(0, 10), range(7, 15), range(13, 17)]
    # :933-972 query coarser than every match
    q = _sig(smb, range(0, 20), "query", scaled=100)
    counter = CounterGather(q)
    for i, (r, s) in enumerate(zip(ranges, (10, 20, 30))):
        counter.add(_sig(smb, r, f"match{i + 1}", scaled=s))
    assert _names_sizes(_consume_all(q.minhash, counter)) == THREE
    # :975-1013 abundance query, flat matches
    q = _sig(smb, range(0, 20), "query", track=True)
    counter = CounterGather(q)
    for i, r in enumerate(ranges):
        counter.add(_sig(smb, r, f"match{i + 1}"))
    assert _names_sizes(_consume_all(q.minhash.flatten(), counter)) == THREE
    # :1016-1054 flat query, abundance matches
    q = _sig(smb, range(0, 20), "query")
    counter = CounterGather(q)
    for i, r in enumerate(ranges):
        counter.add(_sig(smb, r, f"match{i + 1}", track=True))
    assert _names_sizes(_consume_all(q.minhash.flatten(), counter)) == THREE


def test_counter_gather_edge_cases(smb):
    from sourmash_b200.index import CounterGather
    q = _sig(smb, range(0, 20), "query")
    # :1100-1116 exact match
    counter = CounterGather(q)
    counter.add(q, location="somewhere over the rainbow")
    (sr, n), = _consume_all(q.minhash, counter)
    assert sr.score == 1.0 and sr.signature == q and sr.location == "somewhere over the rainbow"
    # :1119-1144 identical matches collapse (md5 keyed)
    counter = CounterGather(q)
    for name in ("match1", "match2", "match3"):
        counter.add(_sig(smb, range(5, 15), name), location=name)
    (sr, n), = _consume_all(q.minhash, counter)
    assert sr.score == 0.5 and n == 10 and sr.location in ("match1", "match2", "match3")
    # :1147-1176 no add after peek / consume
    for poke in (lambda c: c.peek(q.minhash), lambda c: c.consume(q.minhash)):
        counter = CounterGather(q)
        counter.add(q, location="x")
        poke(counter)
        with pytest.raises(ValueError):
            counter.add(q, location="try again")
    # :1179-1190 consuming an empty intersection is fine
    counter = CounterGather(q)
    counter.add(q)
    counter.consume(q.minhash.copy_and_clear())
    # :1193-1206 empty initial query
    empty = _sig(smb, [], "query")
    counter = CounterGather(empty)
    counter.add(_sig(smb, range(0, 10), "match1"), require_overlap=False)
    assert counter.peek(empty.minhash) == []
    # :1209-1216 num query
    num = smb.MinHash(500, 31)
    num.add_many(range(0, 10))
    with pytest.raises(ValueError):
        CounterGather(smb.SourmashSignature(num, name="query"))
    # :1219-1231 empty current query
    counter = CounterGather(q)
    counter.add(q)
    assert _consume_all(q.minhash.copy_and_clear(), counter) == []
    # :1234-1247 num match
    nm = smb.MinHash(500, 31)
    nm.add_many(range(0, 20))
    counter = CounterGather(q)
    with pytest.raises(ValueError):
        counter.add(smb.SourmashSignature(nm, name="query"), location="x")
    # :1250-1263 current query not a subset of the original
    counter = CounterGather(q)
    counter.add(q)
    bad = q.minhash.copy_and_clear()
    bad.add_many(range(20, 30))
    with pytest.raises(ValueError):
        counter.peek(bad)
    # :1266-1281 no overlap
    q10 = _sig(smb, range(0, 10), "query")
    counter = CounterGather(q10)
    with pytest.raises(ValueError):
        counter.add(_sig(smb, range(10, 20), "match1"))
    assert counter.peek(q10.minhash) == []
    # :1284-1301 unattainable threshold
    counter = CounterGather(q)
    counter.add(_sig(smb, range(0, 10), "match1"))
    assert counter.peek(q.minhash, threshold_bp=30 * q.minhash.scaled) == []
    # :1304-1310 empty counter
    assert CounterGather(empty).peek(empty.minhash) == []


def test_gather_real_data_three_routes(smb, golden):
    """:1057-1097 -- 12 genomes vs their combined metagenome, k=21."""
    from sourmash_b200.gather import gather_databases
    from sourmash_b200.index import CounterGather, LinearIndex, gather
    from sourmash_b200.sigset import SignatureSet
    want = golden["meta"]["gather_k21_expected"]
    d = os.path.join(GOLDEN, "gather")
    paths = sorted(glob.glob(os.path.join(d, "GCF*.sig")))
    query = smb.signature.load_one_signature_from_json(os.path.join(d, "combined.sig"), ksize=21)
    subjects = [(smb.signature.load_one_signature_from_json(p, ksize=21), p) for p in paths]
    counter = CounterGather(query)
    for ss, loc in subjects:
        counter.add(ss, location=loc)
    assert _names_sizes(_consume_all(query.minhash, counter)) == want
    hits = list(gather(query, LinearIndex([ss for ss, _ in subjects]), threshold_bp=0))
    assert [[h.match.name.split()[0], h.intersect_size] for h in hits] == want
    sset = SignatureSet.from_files(paths)
    rows = sset.select(ksize=21)
    report = gather_databases(query.minhash, sset.to_sketchset(rows), names=[sset.name(i) for i in rows],
                              md5s=[sset.md5sum(i) for i in rows])
    assert [[g.name.split()[0], g.unique_intersect_bp // g.scaled] for g in report] == want
    assert report[0].md5 == subjects[[s.name.split()[0] for s, _ in subjects].index("NC_003198.1")][0].md5sum()
    assert report[-1].remaining_bp == (len(query.minhash) - sum(n for _, n in want)) * query.minhash.scaled
    assert abs(sum(g.f_unique_to_query for g in report) - sum(n for _, n in want) / len(query.minhash)) < 1e-12
