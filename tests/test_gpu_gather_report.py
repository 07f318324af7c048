"""Gather / prefetch reports (sourmash_b200.gather) against a brute-force restatement of the
reference loop with Python sets: GatherDatabases.__next__ (search.py:877-949), the column
definitions of GatherResult.build_gather_result (search.py:548-620) and CounterGather's tie
rule (index/__init__.py:841).  Integer columns exact, float columns bit-equal (same formulas)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bias(n, scaled):
    return 1.0 - (1.0 - 1.0 / scaled) ** float(n * scaled)


def _cont(c, n, scaled):
    if n == 0:
        return 0
    v = c / (n * _bias(n, scaled))
    return 1.0 if v >= 1 else 0.0 if v <= 0 else v


def _brute_gather(q_abund, rows, scaled, threshold_bp, track):
    orig = set(q_abund)
    total_w = sum(q_abund.values()) if track else len(orig)
    counters = {i: len(orig & r) for i, r in enumerate(rows) if len(orig & r)}
    remaining = set(orig)
    out = []
    while remaining and counters:
        n_thr = float(threshold_bp) / scaled if threshold_bp else 0
        if threshold_bp and n_thr / len(remaining) > 1.0:
            break
        best = max(counters.values())
        i = next(k for k, v in counters.items() if v == best)          # first inserted wins ties
        if best < n_thr:
            break
        isect = remaining & rows[i]
        ab = [q_abund[h] if track else 1 for h in sorted(isect)]
        for k in list(counters):
            counters[k] -= len(isect & rows[k])
            if counters[k] == 0:
                del counters[k]
        new_remaining = remaining - rows[i]
        row = {"row": i, "intersect_bp": len(orig & rows[i]) * scaled, "unique_intersect_bp": len(isect) * scaled,
               "f_orig_query": len(orig & rows[i]) / len(orig), "f_unique_to_query": len(isect) / len(orig),
               "f_match_orig": _cont(len(orig & rows[i]), len(rows[i]), scaled),
               "f_match": _cont(len(isect), len(rows[i]), scaled),
               "remaining_bp": (len(remaining) - len(isect)) * scaled,
               "sum_weighted_found": total_w - (sum(q_abund[h] for h in new_remaining) if track else len(new_remaining)),
               "gather_result_rank": len(out)}
        if track:
            row.update(n_unique_weighted_found=sum(ab), f_unique_weighted=sum(ab) / total_w,
                       average_abund=np.mean(ab), median_abund=np.median(ab), std_abund=np.std(ab))
        else:
            row["f_unique_weighted"] = row["f_unique_to_query"]
        out.append(row)
        remaining = new_remaining
    return out


@pytest.mark.parametrize("track", [False, True])
@pytest.mark.parametrize("threshold_bp", [0, 30000])
def test_gather_report_matches_bruteforce(track, threshold_bp):
    import sourmash_b200 as smb
    from sourmash_b200 import batch as B
    from sourmash_b200.gather import gather_databases
    rng = np.random.default_rng(5)
    scaled = 1000
    mx = B.max_hash_for_scaled(scaled)
    pool = np.unique(rng.integers(1, mx, size=4000, dtype=np.uint64))
    q = rng.choice(pool, size=2500, replace=False)
    rows = []
    for i in range(40):
        own = rng.choice(pool, size=int(rng.integers(20, 400)), replace=False)
        extra = np.unique(rng.integers(1, mx, size=int(rng.integers(10, 200)), dtype=np.uint64))
        rows.append(np.unique(np.concatenate([own, extra])))
    rows[7] = rows[3].copy()                                           # exact tie: first one must win
    rows.append(np.unique(rng.integers(1, mx, size=50, dtype=np.uint64)))   # no overlap at all
    q_abund = {int(h): int(a) for h, a in zip(q, rng.integers(1, 40, size=len(q)))}
    mh = smb.MinHash(0, 31, scaled=scaled, track_abundance=track)
    if track:
        mh.set_abundances(q_abund)
    else:
        mh.add_many(q)
    db = B.SketchSet.from_rows(rows)
    got = gather_databases(mh, db, threshold_bp=threshold_bp, names=[f"g{i}" for i in range(len(rows))])
    want = _brute_gather(q_abund, [set(int(x) for x in r) for r in rows], scaled, threshold_bp, track)
    assert [g.row for g in got] == [w["row"] for w in want] and len(got) > 5
    assert 7 not in [g.row for g in got]
    for g, w in zip(got, want):
        for k, v in w.items():
            assert getattr(g, k) == v, (g.row, k, getattr(g, k), v)
        assert g.name == f"g{g.row}" and g.query_bp == len(q) * scaled and g.scaled == scaled
        d = g.to_dict()
        assert ("average_abund" in d) == track
        # ANI columns = the per-object estimators on the same two sketches
        m = smb.MinHash(0, 31, scaled=scaled)
        m.add_many(rows[g.row])
        flat = mh.flatten()
        assert g.query_containment_ani == flat.containment_ani(m).ani
        assert g.match_containment_ani == m.containment_ani(flat).ani
        if g.query_containment_ani is not None and g.match_containment_ani is not None:
            assert g.average_containment_ani == flat.avg_containment_ani(m)
            assert g.max_containment_ani == max(g.query_containment_ani, g.match_containment_ani)
    assert got[-1].sum_weighted_found <= got[-1].total_weighted_hashes


def test_prefetch_report():
    import sourmash_b200 as smb
    from sourmash_b200 import batch as B
    from sourmash_b200.gather import prefetch_database
    rng = np.random.default_rng(9)
    scaled = 100
    mx = B.max_hash_for_scaled(scaled)
    q = np.unique(rng.integers(1, mx, size=3000, dtype=np.uint64))
    rows = [np.unique(np.concatenate([rng.choice(q, size=n, replace=False),
                                      rng.integers(1, mx, size=300, dtype=np.uint64)])) for n in (0, 5, 60, 900, 2999)]
    mh = smb.MinHash(0, 21, scaled=scaled)
    mh.add_many(q)
    res = prefetch_database(mh, B.SketchSet.from_rows(rows), 5000)
    assert [d["row"] for d in res] == [2, 3, 4]                         # >= 50 shared hashes
    for d in res:
        m = smb.MinHash(0, 21, scaled=scaled)
        m.add_many(rows[d["row"]])
        c = mh.count_common(m)
        assert d["intersect_bp"] == c * scaled and d["jaccard"] == mh.jaccard(m)
        assert d["f_match_query"] == mh.contained_by(m) and d["f_query_match"] == m.contained_by(mh)
        assert d["max_containment"] == mh.max_containment(m)
        assert d.get("query_containment_ani") == mh.containment_ani(m).ani
        assert d.get("match_containment_ani") == m.containment_ani(mh).ani
    with pytest.raises(ValueError):
        prefetch_database(mh, B.SketchSet.from_rows(rows), 10**9)
