"""Randomised differential tests (hypothesis): CUDA path vs oracle on arbitrary small inputs --
the pattern of the reference's proptest oracles (src/core/tests/minhash.rs:182-378) and
tests/test__minhash_hypothesis.py."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import oracle as orc

pytestmark = pytest.mark.gpu

COMMON = dict(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])


@pytest.fixture(scope="module")
def B():
    from sourmash_b200 import batch
    assert batch.device_count() > 0
    return batch


key_sets = st.lists(
    st.one_of(st.integers(0, 40), st.integers(0, 2**64 - 1), st.integers(2**64 - 4, 2**64 - 1),
              st.integers(0, 2**20).map(lambda x: x << 40)),
    max_size=60).map(lambda xs: np.unique(np.array(xs, dtype=np.uint64)))


@given(st.lists(key_sets, min_size=1, max_size=7))
@settings(**COMMON)
def test_pairwise_counts_and_jaccard(B, rows):
    h, off = orc.to_csr(rows)
    sset = B.SketchSet.from_host(h, off)
    assert np.array_equal(B.pairwise_common(sset), orc.pairwise_common(h, off))
    assert np.array_equal(B.compare_jaccard(sset), orc.compare_all_pairs(h, off))
    q = rows[0]
    assert np.array_equal(B.one_vs_many(q, sset), orc.one_vs_many(q, h, off).astype(np.uint32))


@given(st.lists(key_sets, min_size=2, max_size=6), st.integers(1, 25))
@settings(**COMMON)
def test_num_semantics(B, rows, num):
    rows = [r[:num] for r in rows]
    h, off = orc.to_csr(rows)
    got = B.compare_jaccard(B.SketchSet.from_host(h, off), num=num)
    assert np.array_equal(got, orc.compare_all_pairs(h, off, num=num))


dna = st.text(alphabet="ACGTacgtNnRX-", min_size=0, max_size=400)


@given(st.lists(dna, min_size=1, max_size=5), st.sampled_from([1, 3, 4, 5, 16, 17, 21, 31, 32, 33, 51, 64]),
       st.sampled_from([1, 2, 7, 1000]))
@settings(**COMMON)
def test_sketch_sequences_scaled(B, seqs, k, scaled):
    raw = [s.encode() for s in seqs]
    data = np.frombuffer(b"".join(raw), dtype=np.uint8) if any(raw) else np.zeros(0, np.uint8)
    offs = np.cumsum([0] + [len(r) for r in raw]).astype(np.uint64)
    sset, nk = B.sketch_sequences(data, offs, [k], scaled=scaled)
    mx = orc.max_hash_for_scaled(scaled)
    for row, r in zip(sset.rows(), raw):
        assert np.array_equal(row, orc.sketch_scaled(r, k, mx))
    assert nk == sum(max(len(r) - k + 1, 0) for r in raw)


@given(dna, st.sampled_from([3, 21, 31]), st.integers(1, 12), st.booleans())
@settings(**COMMON)
def test_sketch_num_and_abundance(B, seq, k, num, track):
    raw = (seq * 3).encode()
    data = np.frombuffer(raw, dtype=np.uint8) if raw else np.zeros(0, np.uint8)
    sset, _ = B.sketch_sequences(data, [0, len(raw)], [k], num=num, track_abundance=track)
    h, off, ab = sset.to_host(with_abunds=True)
    om = orc.OracleMinHash(scaled=0, ksize=k, num=num, track_abundance=track)
    om.add_sequence(raw, force=True)
    assert h.tolist() == om.mins().tolist()
    if track:
        assert ab.tolist() == om.abunds().tolist()


@given(st.lists(key_sets, min_size=1, max_size=8), key_sets, st.integers(1, 3))
@settings(**COMMON)
def test_gather(B, rows, query, threshold):
    db = B.SketchSet.from_rows(rows)
    ids, sizes = B.gather(query, db, threshold=threshold)
    q = np.array(query, dtype=np.uint64)
    counts = [orc.count_common(q, r) for r in rows]
    want = []
    while True:
        best = max(counts) if counts else 0
        if best < threshold or best == 0:
            break
        j = counts.index(best)
        isect = np.intersect1d(q, rows[j])
        want.append((j, len(isect)))
        counts = [c - orc.count_common(isect, r) for c, r in zip(counts, rows)]
        q = np.setdiff1d(q, isect)
        if not len(q):
            break
    assert list(zip(ids.tolist(), sizes.tolist())) == want


residues = st.text(alphabet="ACDEFGHIKLMNPQRSTVWYXBZ*acdefghiklmnpqrstvwy-", min_size=0, max_size=300)


@given(st.lists(dna, min_size=1, max_size=4), st.sampled_from(["protein", "dayhoff", "hp"]),
       st.sampled_from([1, 2, 5, 7, 10, 16, 17, 42]), st.sampled_from([1, 3, 200]), st.booleans())
@settings(**COMMON)
def test_sketch_translate(B, seqs, moltype, kaa, scaled, track):
    raw = [s.encode() for s in seqs]
    data = np.frombuffer(b"".join(raw), dtype=np.uint8) if any(raw) else np.zeros(0, np.uint8)
    offs = np.cumsum([0] + [len(r) for r in raw]).astype(np.uint64)
    sset, nk = B.sketch_sequences(data, offs, [kaa], scaled=scaled, moltype=moltype, track_abundance=track)
    h, off, ab = sset.to_host(with_abunds=True)
    for i, r in enumerate(raw):
        om = orc.OracleMinHash(scaled=scaled, ksize=3 * kaa, track_abundance=track)
        om.add_protein_family(r, moltype, False)
        assert h[int(off[i]):int(off[i + 1])].tolist() == om.mins().tolist()
        if track:
            assert ab[int(off[i]):int(off[i + 1])].tolist() == om.abunds().tolist()
    assert nk == sum(2 * max(len(r) - 3 * kaa + 1, 0) for r in raw)


@given(st.lists(residues, min_size=1, max_size=4), st.sampled_from(["protein", "dayhoff", "hp"]),
       st.sampled_from([1, 3, 7, 10, 16, 19, 42]), st.integers(0, 9))
@settings(**COMMON)
def test_sketch_protein_num(B, seqs, moltype, kaa, num):
    raw = [s.encode() for s in seqs]
    data = np.frombuffer(b"".join(raw), dtype=np.uint8) if any(raw) else np.zeros(0, np.uint8)
    offs = np.cumsum([0] + [len(r) for r in raw]).astype(np.uint64)
    scaled = 1 if num == 0 else 0
    sset, nk = B.sketch_sequences(data, offs, [kaa], scaled=scaled, num=num, moltype=moltype, input_is_protein=True)
    for row, r in zip(sset.rows(), raw):
        om = orc.OracleMinHash(scaled=scaled, ksize=3 * kaa, num=num)
        om.add_protein_family(r, moltype, True)
        assert row.tolist() == om.mins().tolist()
    assert nk == sum(max(len(r) - kaa + 1, 0) for r in raw)
