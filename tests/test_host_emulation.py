"""CPU-only check of the *device* per-thread k-mer code (csrc/kmer_roll.cuh compiled for the
host by tests/host_emul/roll_emul.cu) against the oracle: rolling forward/revcomp words,
canonical choice, murmur3 over register words, tile/lead/tail masking."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle as orc
from sourmash_b200.synth import synth_genome

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "roll_emul.cu")


@pytest.fixture(scope="module")
def emul():
    exe = os.path.join(tempfile.gettempdir(), "smb_roll_emul")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I/usr/local/cuda/include", "-x", "c++", SRC, "-o", exe])

    def run(seq, k, W, lead=0):
        with tempfile.TemporaryDirectory() as td:
            fin, fout = os.path.join(td, "in"), os.path.join(td, "out")
            with open(fin, "wb") as fh:
                fh.write(b"G" * lead + bytes(seq))
            subprocess.check_call([exe, str(k), str(W), str(lead), fin, fout])
            return np.fromfile(fout, dtype=np.uint64)
    return run


@pytest.mark.parametrize("k", [1, 3, 4, 5, 8, 15, 16, 17, 21, 24, 31, 32, 33, 47, 48, 51, 63, 64, 65])
def test_roll_matches_oracle(emul, k):
    g = synth_genome(5000 + k, seed=100 + k, n_every=89)
    g[1000:1500] = np.frombuffer(bytes(g[1000:1500]).lower(), dtype=np.uint8)
    g[2000] = ord("R")
    g[2001] = 0
    g[-1] = ord("n")
    want, err = orc.seq_to_hashes(bytes(g), k, force=True, keep_zeros=True)
    assert err is None
    for W, lead in ((16, 0), (64, 5), (128, 15)):
        got = emul(g, k, W, lead)
        assert np.array_equal(got, want), (k, W, lead)


def test_roll_fused_21_31_51_matches_oracle(emul):
    """The fused pass (one rolling 51-state, the 21- and 31-mers as prefixes; experimental
    SMB_SKETCH_FUSED): all three hash streams equal the oracle, invalid bases and stream ends included."""
    for n, seed in ((5000, 7), (51, 8), (52, 9), (50, 10), (31, 11), (30, 12), (21, 13), (20, 14), (137, 15)):
        g = synth_genome(max(n, 64), seed=seed, n_every=61)[:n]
        if n >= 200:
            g[100:160] = np.frombuffer(bytes(g[100:160]).lower(), dtype=np.uint8)
            g[1000] = ord("R"); g[1001] = 0; g[1050] = ord("n"); g[1071] = ord("N"); g[-25] = ord("N")
        want = []
        for k in (21, 31, 51):
            if n >= k:
                hs, err = orc.seq_to_hashes(bytes(g), k, force=True, keep_zeros=True)
                assert err is None
                want.append(np.asarray(hs, dtype=np.uint64))
            else:
                want.append(np.zeros(0, np.uint64))
        want = np.concatenate(want)
        for W, lead in ((16, 0), (64, 5), (128, 15)):
            got = emul(g, 0, W, lead)
            assert np.array_equal(got, want), (n, W, lead)


def test_roll_short_and_exact_lengths(emul):
    for k in (21, 31):
        for n in (k - 1, k, k + 1, 15, 16, 17, 47, 48, 49):
            g = synth_genome(max(n, 1), seed=n)[:n]
            if n < k:
                assert len(emul(g, k, 16, 3)) == 0
                continue
            want, _ = orc.seq_to_hashes(bytes(g), k, force=True, keep_zeros=True)
            assert np.array_equal(emul(g, k, 16, 3), want)


# ---------------------------------------------------------------------------------------------
# intersection kernel: table build + probes (csrc/split_table.cuh) driven like the CUDA kernel
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def tile_emul():
    exe = os.path.join(tempfile.gettempdir(), "smb_tile_emul")
    src = os.path.join(HERE, "host_emul", "tile_emul.cu")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I/usr/local/cuda/include", "-x", "c++", src, "-o", exe])

    def run(rows, max_entries=30000):
        hashes, offsets = orc.to_csr(rows)
        finite = hashes[hashes != np.uint64(2**64 - 1)]
        max_key = int(finite.max()) if len(finite) else 0
        if len(hashes) and int(hashes.max()) == 2**64 - 1:
            max_key = 2**64 - 1                               # the planner sees the raw maximum
        shift = 0
        while shift < 63 and (max_key >> shift) >= max_entries:
            shift += 1
        nb = (max_key >> shift) + 1
        with tempfile.TemporaryDirectory() as td:
            fin, fout = os.path.join(td, "in"), os.path.join(td, "out")
            with open(fin, "wb") as fh:
                fh.write(np.uint64(len(rows)).tobytes()); fh.write(offsets.tobytes()); fh.write(hashes.tobytes())
            subprocess.check_call([exe, str(shift), str(nb), fin, fout])
            n = len(rows)
            return np.fromfile(fout, dtype=np.uint32).reshape(n, n), (hashes, offsets)
    return run


def test_tile_table_logic_matches_oracle(tile_emul):
    from sourmash_b200.synth import rows_of, synth_sketches
    h, off = synth_sketches(24, mean=700, sd=150, lo=5, hi=1500, n_families=3, pool=900, seed=31)
    rows = rows_of(h, off)
    got, (hh, oo) = tile_emul(rows)
    assert np.array_equal(got, orc.pairwise_common(hh, oo))
    # coarse directory (few buckets => crowded buckets everywhere): the rare path becomes the main path
    got, _ = tile_emul(rows, max_entries=64)
    assert np.array_equal(got, orc.pairwise_common(hh, oo))


def test_tile_table_logic_edge_rows(tile_emul):
    rng = np.random.Generator(np.random.PCG64(3))
    big = 2**64 - 1
    rows = [np.zeros(0, np.uint64), np.array([0], np.uint64), np.array([0, 1, 2, 3, big], np.uint64),
            np.array([big], np.uint64), np.array([big - 1, big], np.uint64),
            np.arange(1, 700, dtype=np.uint64),                                       # dense: one bucket
            np.unique(rng.integers(0, 2**63, size=900, dtype=np.uint64)),
            # equal low words, different high words in neighbouring slots (low-word false positives)
            np.array([(7 << 32) | 5, (8 << 32) | 5, (9 << 32) | 5, (9 << 32) | 6], dtype=np.uint64),
            np.array([(8 << 32) | 5, (9 << 32) | 6, (10 << 32) | 5], dtype=np.uint64),
            np.unique(rng.integers(0, 1000, size=300, dtype=np.uint64))]
    for max_entries in (30000, 16, 2):
        got, (hh, oo) = tile_emul(rows, max_entries=max_entries)
        assert np.array_equal(got, orc.pairwise_common(hh, oo)), max_entries


# ---------------------------------------------------------------------------------------------
# protein-family kernel logic (csrc/aa_kmers.cuh): tables, staging, strided gathers, frame order
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def aa_emul():
    exe = os.path.join(tempfile.gettempdir(), "smb_aa_emul")
    src = os.path.join(HERE, "host_emul", "aa_emul.cu")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I/usr/local/cuda/include", "-x", "c++", src, "-o", exe])

    def run(seq, moltype, kaa, translate):
        with tempfile.TemporaryDirectory() as td:
            fin, fout = os.path.join(td, "in"), os.path.join(td, "out")
            with open(fin, "wb") as fh:
                fh.write(bytes(seq))
            subprocess.check_call([exe, str(orc.HASH_FUNCTIONS[moltype]), str(kaa), str(int(translate)), fin, fout])
            return np.fromfile(fout, dtype=np.uint64)
    return run


def _messy_dna(n, seed):
    g = synth_genome(n, seed=seed, n_every=53)
    g[100:160] = np.frombuffer(bytes(g[100:160]).lower(), dtype=np.uint8)
    g[200] = ord("R"); g[201] = 0; g[202] = ord("n"); g[-2] = ord("N")
    return g


@pytest.mark.parametrize("moltype", ["protein", "dayhoff", "hp"])
@pytest.mark.parametrize("kaa", [1, 2, 7, 8, 9, 16, 17, 33, 42])
def test_aa_translate_matches_oracle(aa_emul, moltype, kaa):
    for n in (3 * kaa - 1, 3 * kaa, 3 * kaa + 1, 3 * kaa + 2, 700, 1031):
        g = _messy_dna(max(n, 300), seed=n + kaa)[:n]
        want = orc.seq_to_hashes_translate(bytes(g), kaa, moltype, keep_zeros=True)
        got = aa_emul(g, moltype, kaa, True)
        if n < 3 * kaa:
            assert len(want) == 0 and len(got) == 0
        else:
            assert np.array_equal(got, want[1:-1]), (moltype, kaa, n)


@pytest.mark.parametrize("moltype", ["protein", "dayhoff", "hp"])
@pytest.mark.parametrize("kaa", [1, 3, 7, 10, 16, 19, 42])
def test_aa_protein_matches_oracle(aa_emul, moltype, kaa):
    rng = np.random.default_rng(kaa)
    alphabet = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYXBZJUO*acdefghiklmnpqrstvwy-\x00", dtype=np.uint8)
    for n in (kaa - 1, kaa, kaa + 1, 255, 256, 257, 900):
        seq = alphabet[rng.integers(0, len(alphabet), size=max(n, 0))]
        want = orc.seq_to_hashes_protein(bytes(seq), kaa, moltype, keep_zeros=True)
        got = aa_emul(seq, moltype, kaa, False)
        assert np.array_equal(got, want), (moltype, kaa, n)


# ---------------------------------------------------------------------------------------------
# inverted join (csrc/join_walk.cuh): group walk, group sizes, key-range shards
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def join_emul():
    exe = os.path.join(tempfile.gettempdir(), "smb_join_emul")
    src = os.path.join(HERE, "host_emul", "join_emul.cu")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I/usr/local/cuda/include", "-x", "c++", src, "-o", exe])

    def run(rows, n_shards):
        hashes, offsets = orc.to_csr(rows)
        n = len(rows)
        with tempfile.TemporaryDirectory() as td:
            fh, fo, fc, fp = (os.path.join(td, x) for x in ("h", "o", "c", "p"))
            hashes.tofile(fh); offsets.tofile(fo)
            subprocess.check_call([exe, str(n_shards), fh, fo, fc, fp])
            got = np.fromfile(fc, dtype=np.uint32).reshape(n, n)
            return got, int(np.fromfile(fp, dtype=np.uint64)[0])
    return run


def test_join_walk_matches_oracle(join_emul):
    from sourmash_b200.synth import synth_sketches
    rng = np.random.default_rng(3)
    h, off = synth_sketches(90, mean=300, sd=60, lo=100, hi=600, n_families=6, pool=400, seed=9)
    fam = [h[int(off[i]):int(off[i + 1])] for i in range(90)]
    big = np.uint64(2**64 - 1)
    edge = [np.unique(np.concatenate([rng.integers(0, 2**64 - 1, size=int(rng.integers(0, 30)), dtype=np.uint64),
                                      np.array([0, 5, big] if i % 3 == 0 else [5], dtype=np.uint64)])) for i in range(40)]
    edge[7] = np.zeros(0, np.uint64)
    edge[9] = edge[8].copy()
    dense = [np.arange(i % 4, 30, dtype=np.uint64) for i in range(25)]
    for rows in (fam, edge, dense):
        hh, oo = orc.to_csr(rows)
        want = orc.pairwise_common(hh, oo)
        iu = np.triu_indices(len(rows), 1)
        for shards in (1, 2, 5):
            got, pairs = join_emul(rows, shards)
            assert np.array_equal(got[iu], want[iu]), shards
            assert int(np.tril(got).sum()) == 0                      # only the upper triangle is touched
            assert pairs == int(want[iu].sum())                       # sum of C(m,2) == sum of all intersections


def test_join_stripe_helpers():
    "rows per CTA for the shared-memory budget the kernel uses (227 KB minus its 544-byte header): 10 000 columns -> 5 rows."
    exe = os.path.join(tempfile.gettempdir(), "smb_stripe_rows")
    src = os.path.join(tempfile.gettempdir(), "smb_stripe_rows.cu")
    with open(src, "w") as fh:
        fh.write('#include <stdio.h>\n#include "%s"\nint main() { int ns[] = {1, 64, 1024, 10000, 57900, 58000, 200000};'
                 ' for (int n : ns) printf("%%d ", smb::stripe_rows_per_block(227 * 1024, n)); return 0; }\n'
                 % os.path.join(os.path.dirname(HERE), "sourmash_b200", "csrc", "join_stripe.cuh"))
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-w", "-DSMB_SIMT_EMUL=1", "-include", os.path.join(HERE, "host_emul", "simt.h"),
                           "-I/usr/local/cuda/include", "-x", "c++", src, "-o", exe])
    got = [int(x) for x in subprocess.check_output([exe]).split()]
    assert got == [32, 32, 32, 5, 1, 0, 0]


# ---------------------------------------------------------------------------------------------
# inverted index over a resident set (csrc/db_index.cuh): smb_sketchset_build_index
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def index_emul():
    exe = os.path.join(tempfile.gettempdir(), "smb_index_emul")
    src = os.path.join(HERE, "host_emul", "index_emul.cu")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I/usr/local/cuda/include", "-x", "c++", src, "-o", exe])

    def run(query, rows):
        hashes, offsets = orc.to_csr(rows)
        with tempfile.TemporaryDirectory() as td:
            fq, fh, fo, fc = (os.path.join(td, x) for x in ("q", "h", "o", "c"))
            np.asarray(query, dtype=np.uint64).tofile(fq); hashes.tofile(fh); offsets.tofile(fo)
            subprocess.check_call([exe, fq, fh, fo, fc])
            return np.fromfile(fc, dtype=np.uint32)
    return run


def test_db_index_counts_match_oracle(index_emul):
    """Distinct keys / offsets / rows / directory of the inverted index and the per-query-hash lookup:
    the counts equal the oracle's one-vs-many (so search, prefetch and every gather round do)."""
    from sourmash_b200.synth import synth_sketches
    rng = np.random.default_rng(6)
    mx = orc.max_hash_for_scaled(1000)
    h, off = synth_sketches(80, mean=400, sd=80, lo=50, hi=800, n_families=5, pool=500, seed=43)
    fam = [h[int(off[i]):int(off[i + 1])] for i in range(80)]
    big = np.uint64(2**64 - 1)
    edge = [np.zeros(0, np.uint64), np.array([0], np.uint64), np.array([0, 1, 2, big - 1, big], np.uint64),
            np.array([big], np.uint64), np.arange(1, 300, dtype=np.uint64), np.arange(1, 300, dtype=np.uint64),
            np.unique(rng.integers(0, 2**64 - 1, size=500, dtype=np.uint64))]
    wide = [np.array([7, 1000 + i], dtype=np.uint64) for i in range(100)]                 # one hash in 100 rows
    queries = [np.unique(np.concatenate([fam[3], fam[17][:200], rng.integers(1, mx, size=3000, dtype=np.uint64),
                                         np.array([mx + 5, 2**63, 2**64 - 1], dtype=np.uint64)])),
               np.unique(np.concatenate([edge[2], edge[6][::3], np.array([1, 299, 300, 2**40], dtype=np.uint64)])),
               np.array([7], dtype=np.uint64), np.zeros(0, np.uint64), np.array([0, big], dtype=np.uint64)]
    for rows in (fam, edge, wide, [edge[0]], [edge[0], edge[0]]):
        hh, oo = orc.to_csr(rows)
        for query in queries:
            want = orc.one_vs_many(np.asarray(query, dtype=np.uint64), hh, oo).astype(np.uint32) if len(hh) \
                else np.zeros(len(rows), np.uint32)
            assert np.array_equal(index_emul(query, rows), want), (len(rows), len(query))


# ---------------------------------------------------------------------------------------------
# the experimental KERNELS themselves on the CPU (tests/host_emul/simt.h: CTAs as cooperative fibers with
# real __syncthreads / warp collectives / shared memory), csrc/join_kernels.cuh, join_stripe.cuh, range_kernels.cuh, db_index_kernels.cuh
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def simt():
    exe = os.path.join(tempfile.gettempdir(), "smb_simt_emul")
    src = os.path.join(HERE, "host_emul", "simt_emul.cu")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-w", "-I/usr/local/cuda/include", "-x", "c++", src, "-o", exe])

    def run(mode, rows, *args, query=None, dtype=np.uint32):
        hashes, offsets = orc.to_csr(rows)
        with tempfile.TemporaryDirectory() as td:
            fq, fh, fo, fc = (os.path.join(td, x) for x in ("q", "h", "o", "c"))
            hashes.tofile(fh); offsets.tofile(fo)
            cmd = [exe, mode] + [str(a) for a in args]
            if query is not None:
                np.asarray(query, dtype=np.uint64).tofile(fq)
                cmd.append(fq)
            subprocess.check_call(cmd + [fh, fo, fc])
            return np.fromfile(fc, dtype=dtype)
    return run


def test_simt_stripe_kernels_match_oracle(simt):
    """The stripe pipeline kernel by kernel as join_stripe_create / join_stripe_rows launch it: 32-bit keys +
    payloads, (host-sorted,) element-block row table, u16 / u32 tags + inverse permutation with descent detection, the
    warp-per-run repair of runs that mix hashes, the count kernel (dynamic (row, chunk) items, four group
    reads in flight, upper-only and two-direction modes, shared-memory stripe, fused float64 finalize), mirror,
    row chunks.  Sort keys shortened to 8 / 3 bits make mixed runs the rule rather than a 1-in-10^4 event."""
    from sourmash_b200.synth import synth_sketches
    rng = np.random.default_rng(31)
    h, off = synth_sketches(70, mean=120, sd=30, lo=20, hi=250, n_families=2, pool=160, seed=29)
    fam = [h[int(off[i]):int(off[i + 1])] for i in range(70)]
    wide = [np.unique(np.concatenate([rng.integers(1, 2**60, size=2, dtype=np.uint64), np.array([7], dtype=np.uint64),
                                      np.array([2**61] if i < 33 else [], dtype=np.uint64)])) for i in range(80)]
    wide[3] = np.zeros(0, np.uint64)
    big = np.uint64(2**64 - 1)
    edge = [np.unique(np.concatenate([rng.integers(0, 2**64 - 1, size=int(rng.integers(0, 20)), dtype=np.uint64),
                                      np.array([0, 5, big] if i % 3 == 0 else [5], dtype=np.uint64)])) for i in range(45)]
    edge[7] = np.zeros(0, np.uint64)
    edge[9] = edge[8].copy()
    small = [np.arange(1 + i % 3, 40 + i, dtype=np.uint64) for i in range(37)]          # dense small integers: no low bits at all
    small[0] = np.zeros(0, np.uint64)
    small[36] = np.zeros(0, np.uint64)
    # one hash shared by 150 rows (several 32-tag chunks in both directions), one by exactly 33 and one by 65 rows
    many = [np.unique(np.concatenate([rng.integers(1, 2**60, size=3, dtype=np.uint64), np.array([7], dtype=np.uint64),
                                      np.array([2**61] if i < 33 else [], dtype=np.uint64),
                                      np.array([2**62] if i >= 85 else [], dtype=np.uint64)])) for i in range(150)]
    tiny = [np.array([5], np.uint64)]
    pair = [np.array([1, 2, 3], np.uint64), np.array([2, 3, 4], np.uint64)]
    for rows in (fam, wide, edge, small, many, tiny, pair):
        hh, oo = orc.to_csr(rows)
        want = orc.compare_all_pairs(hh, oo, nthreads=2)
        n = len(rows)
        for R, upper, threads, tag_bits, sort_bits in ((5, 0, 64, 16, 32), (3, 1, 96, 32, 32), (32, 0, 32, 16, 8),
                                                       (7, 1, 128, 16, 3), (4, 1, 64, 32, 8)):
            got = simt("stripe", rows, R, upper, threads, tag_bits, sort_bits, dtype=np.float64).reshape(n, n)
            assert np.array_equal(got, want), (n, R, upper, threads, tag_bits, sort_bits)


def test_simt_stripe_repairs_runs_that_mix_hashes(simt, capfd):
    "hashes that agree in their top 32 bits and differ below (what a 32-bit sort key cannot separate), interleaved over rows"
    rng = np.random.default_rng(33)
    top = [np.uint64(v) << np.uint64(32) for v in (9, 3, 7)]
    lows = [np.uint64(0x1234abcd), np.uint64(0x00000001), np.uint64(0xffffffff)]
    rows = []
    for i in range(45):
        mine = [t | lo for ti, t in enumerate(top) for li, lo in enumerate(lows) if (i + ti + 2 * li) % 3 != 0]
        rows.append(np.unique(np.array(mine + rng.integers(1, 2**63, size=30, dtype=np.uint64).tolist(), dtype=np.uint64)))
    rows[6] = np.zeros(0, np.uint64)
    hh, oo = orc.to_csr(rows)
    want = orc.compare_all_pairs(hh, oo, nthreads=2)
    os.environ["SMB_EMUL_REPORT"] = "1"
    try:
        for R, upper, threads, tag_bits in ((5, 0, 64, 16), (4, 1, 96, 32)):
            got = simt("stripe", rows, R, upper, threads, tag_bits, 32, dtype=np.float64).reshape(45, 45)
            assert np.array_equal(got, want), (R, upper, threads)
    finally:
        del os.environ["SMB_EMUL_REPORT"]
    assert "mixed runs repaired: 3" in capfd.readouterr().err


def test_simt_join_kernels_match_oracle(simt):
    """The global-reduction join as launched (the stripe layout's fallback and the key-range-sharded form):
    row slices, gather, count, the estimate kernel's invariant, key-range shards."""
    from sourmash_b200.synth import synth_sketches
    rng = np.random.default_rng(3)
    h, off = synth_sketches(70, mean=200, sd=40, lo=50, hi=400, n_families=4, pool=260, seed=9)
    fam = [h[int(off[i]):int(off[i + 1])] for i in range(70)]
    big = np.uint64(2**64 - 1)
    edge = [np.unique(np.concatenate([rng.integers(0, 2**64 - 1, size=int(rng.integers(0, 20)), dtype=np.uint64),
                                      np.array([0, 5, big] if i % 3 == 0 else [5], dtype=np.uint64)])) for i in range(45)]
    edge[7] = np.zeros(0, np.uint64)
    edge[9] = edge[8].copy()
    for rows in (fam, edge):
        n = len(rows)
        hh, oo = orc.to_csr(rows)
        want = orc.pairwise_common(hh, oo)
        iu = np.triu_indices(n, 1)
        full = np.zeros_like(want)
        full[iu] = want[iu]
        for shards in (1, 3):
            got = simt("join", rows, shards).reshape(n, n)
            assert np.array_equal(got[iu], want[iu]) and int(np.tril(got).sum()) == 0, ("join", shards)


def test_simt_range_search_kernels_match_oracle(simt):
    """The range-major layout and its streaming pass as written: rm_bounds / rm_counts / rm_scatter_kernel (every element
    in exactly one part, rows in order inside a part) and one_vs_many_range_major_kernel (two-probe bitmap in shared
    memory, eight loads in flight, per-warp candidate queues and their drain: exact test, row attribution).  Bitmaps of
    2^6 .. 2^16 bits instead of 2^19 make false positives common; a 64-candidate burst overflows a queue mid-loop."""
    from sourmash_b200.synth import synth_sketches
    rng = np.random.default_rng(9)
    mx = orc.max_hash_for_scaled(1000)
    h, off = synth_sketches(75, mean=300, sd=60, lo=0, hi=700, n_families=3, pool=400, seed=47)
    fam = [h[int(off[i]):int(off[i + 1])] for i in range(75)]
    fam[11] = np.unique(rng.integers(1, mx, size=3000, dtype=np.uint64))             # a row with slices longer than 64
    big = np.uint64(2**64 - 1)
    edge = [np.zeros(0, np.uint64), np.array([0], np.uint64), np.array([0, 1, 2, big - 1, big], np.uint64),
            np.arange(1, 300, dtype=np.uint64), np.unique(rng.integers(0, 2**64 - 1, size=500, dtype=np.uint64))]
    q_fam = np.unique(np.concatenate([fam[3], fam[17][:100], fam[11][::2], rng.integers(1, mx, size=2000, dtype=np.uint64),
                                      np.array([mx + 5, 2**63], dtype=np.uint64)]))
    q_edge = np.unique(np.concatenate([edge[2], edge[4][::3], np.array([1, 299, 300], dtype=np.uint64)]))
    q_beyond = np.unique(np.concatenate([fam[3], np.array([mx + 5, 2**63, 2**64 - 1], dtype=np.uint64)]))   # keys beyond the database
    q_low = np.unique(rng.integers(0, 1000, size=200, dtype=np.uint64))                                       # query far below the db
    dense = [np.arange(1, 2000, dtype=np.uint64) for _ in range(40)]                                         # every element a match:
    q_dense = np.arange(1, 2000, dtype=np.uint64)                                                            # the queues overflow and drain
    many = [np.unique(rng.integers(1, mx, size=int(rng.integers(0, 25)), dtype=np.uint64)) for _ in range(330)]   # > 5 windows of
    q_many = np.unique(np.concatenate(many[::3] + [rng.integers(1, mx, size=500, dtype=np.uint64)]))               # the coarse row table
    for rows, query in ((fam, q_fam), (edge, q_edge), (fam, fam[5][:1]), (fam, q_beyond), (fam, q_low), (edge, q_low),
                        ([edge[0]], q_edge), (dense, q_dense), (many, q_many)):
        hh, oo = orc.to_csr(rows)
        want = orc.one_vs_many(np.asarray(query, dtype=np.uint64), hh, oo).astype(np.uint32)
        for P, bm_log2, threads in ((5, 16, 64), (2, 6, 96), (9, 12, 32), (3, 8, 64)):
            got = simt("ranges", rows, P, bm_log2, threads, query=query)
            assert np.array_equal(got, want), (len(rows), P, bm_log2, threads)


def test_simt_index_kernels_match_oracle(simt):
    "index_rowid_kernel + index_count_kernel as written (own-lane groups, warp-wide groups, length read on the device)."
    from sourmash_b200.synth import synth_sketches
    rng = np.random.default_rng(10)
    h, off = synth_sketches(60, mean=200, sd=40, lo=0, hi=400, n_families=3, pool=260, seed=53)
    fam = [h[int(off[i]):int(off[i + 1])] for i in range(60)]
    wide = [np.array([7, 1000 + i], dtype=np.uint64) for i in range(100)]            # one hash in 100 rows: the warp-wide path
    wide[4] = np.zeros(0, np.uint64)
    queries = {"fam": np.unique(np.concatenate([fam[3], fam[17][:100], rng.integers(1, 2**54, size=500, dtype=np.uint64)])),
               "wide": np.array([7, 1003, 1050, 5], dtype=np.uint64)}
    for rows, query in ((fam, queries["fam"]), (wide, queries["wide"]), (fam, queries["wide"])):
        hh, oo = orc.to_csr(rows)
        want = orc.one_vs_many(np.asarray(query, dtype=np.uint64), hh, oo).astype(np.uint32)
        for threads in (32, 96):
            got = simt("index", rows, threads, query=query).reshape(2, len(rows))
            assert np.array_equal(got[0], want) and np.array_equal(got[1], want), (len(rows), threads)


@pytest.fixture(scope="module")
def simt_sketch():
    exe = os.path.join(tempfile.gettempdir(), "smb_simt_sketch_emul")
    src = os.path.join(HERE, "host_emul", "simt_sketch_emul.cu")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-w", "-I/usr/local/cuda/include", "-x", "c++", src, "-o", exe])

    def run(k, W, max_hash, lead, genomes, device_sort=False):
        seqs = np.concatenate(genomes) if genomes else np.zeros(0, np.uint8)
        offs = np.cumsum([0] + [len(g) for g in genomes]).astype(np.uint64)
        with tempfile.TemporaryDirectory() as td:
            fs, fo, fc = (os.path.join(td, x) for x in ("s", "o", "c"))
            np.asarray(seqs, dtype=np.uint8).tofile(fs); offs.tofile(fo)
            env = dict(os.environ, SMB_EMUL_DEVICE_SORT="1") if device_sort else None
            subprocess.check_call([exe, str(k), str(W), str(max_hash), str(lead), fs, fo, fc], env=env)
            raw = np.fromfile(fc, dtype=np.uint64)
        rows, at = [], 0
        while at < len(raw):
            n = int(raw[at])
            if device_sort:
                rows.append((raw[at + 1:at + 1 + n], raw[at + 1 + n:at + 1 + 2 * n])); at += 1 + 2 * n
            else:
                rows.append(raw[at + 1:at + 1 + n]); at += 1 + n
        return rows
    return run


def test_simt_hash_kernels_match_oracle(simt_sketch):
    """hash_kmers_kernel<K> and the one-pass hash_kmers_fused_kernel as written: tiling over unaligned streams,
    per-thread windows, shared-memory survivor staging with its overflow path (every hash kept), flushes,
    launches over tile ranges -- the candidate sets equal the oracle's."""
    genomes = [synth_genome(2600, seed=61, n_every=83), synth_genome(900, seed=62), synth_genome(40, seed=63),
               synth_genome(20, seed=64), synth_genome(5000, seed=65)]
    genomes[0][300:380] = np.frombuffer(bytes(genomes[0][300:380]).lower(), dtype=np.uint8)
    genomes[4][1234] = ord("R"); genomes[4][-30] = ord("N")

    def expected(g, k, max_hash):
        if len(g) < k:
            return np.zeros(0, np.uint64)
        hs, err = orc.seq_to_hashes(bytes(g), k, force=True, keep_zeros=True)
        hs = np.asarray(hs, dtype=np.uint64)
        return np.unique(hs[(hs != 0) & (hs <= np.uint64(max_hash))])
    for max_hash in (2**64 - 1, orc.max_hash_for_scaled(20)):                 # keep all (staging overflows) / a sketch threshold
        for W, lead in ((16, 5), (32, 0)):
            for k in (21, 31, 51):
                got = simt_sketch(k, W, max_hash, lead, genomes)
                assert len(got) == len(genomes)
                for g, row in zip(genomes, got):
                    assert np.array_equal(row, expected(g, k, max_hash)), (k, W, lead, len(g))
            fused = simt_sketch(0, W, max_hash, lead, genomes)
            assert len(fused) == 3 * len(genomes)
            for gi, g in enumerate(genomes):
                for ki, k in enumerate((21, 31, 51)):
                    assert np.array_equal(fused[gi * 3 + ki], expected(g, k, max_hash)), ("fused", k, W, lead, len(g))


def test_simt_gather_loop_matches_oracle(simt):
    """The gather session as launched -- directory + bitmap kernels over the query, the global one-vs-many
    pass, counter update + argmax, live intersection, consumed flags -- and the same rounds with the inverted
    index doing the counts (directory of the index built by the device kernels): identical picks."""
    from sourmash_b200.synth import synth_sketches
    rng = np.random.default_rng(14)
    h, off = synth_sketches(50, mean=250, sd=50, lo=80, hi=500, n_families=3, pool=320, seed=71)
    rows = [h[int(off[i]):int(off[i + 1])] for i in range(50)]
    rows[17] = rows[4].copy()                                  # an exact tie: the lowest row wins
    rows[30] = np.zeros(0, np.uint64)
    query = np.unique(np.concatenate([rows[4], rows[9][:150], rows[25][50:250], rows[41][::2],
                                      rng.integers(1, 2**54, size=300, dtype=np.uint64)]))

    def oracle(threshold):
        q = query.copy()
        counts = np.array([orc.count_common(q, r) for r in rows], dtype=np.int64)
        out = []
        while True:
            j = int(np.argmax(counts))
            if counts[j] < threshold or counts[j] == 0:
                break
            isect = np.intersect1d(q, rows[j])
            out += [j, len(isect)]
            counts = counts - np.array([orc.count_common(isect, r) for r in rows], dtype=np.int64)
            q = np.setdiff1d(q, isect)
            if not len(q):
                break
        return out
    for threshold in (1, 40):
        want = oracle(threshold)
        assert len(want) >= 6
        for use_index in (0, 1):
            got = simt("gather", rows, use_index, threshold, query=query)
            assert got.tolist() == want, (threshold, use_index)


def test_simt_sketch_rows_with_abundances(simt_sketch):
    """hash kernel -> sort_unique_small_kernel (block-wide bitonic sort, unique, run lengths) and, for a row with
    more candidates than fit in shared memory, unique_sorted_row_kernel: the sketch rows and their abundances
    equal the oracle's (repeats planted so that abundances exceed one)."""
    unit = synth_genome(700, seed=81)
    genomes = [np.concatenate([unit, unit, unit[:300], synth_genome(500, seed=82)]),       # repeated k-mers
               synth_genome(1200, seed=83, n_every=101), synth_genome(25, seed=84),
               np.concatenate([synth_genome(9000, seed=85)] * 2)]                          # > SORT_MAX candidates when all are kept
    for max_hash in (2**64 - 1, orc.max_hash_for_scaled(10)):
        for k in (21, 0):
            got = simt_sketch(k, 16, max_hash, 3, genomes, device_sort=True)
            ks = (21, 31, 51) if k == 0 else (k,)
            assert len(got) == len(genomes) * len(ks)
            for gi, g in enumerate(genomes):
                for ki, kk in enumerate(ks):
                    hashes, abunds = got[gi * len(ks) + ki]
                    if len(g) < kk:
                        assert len(hashes) == 0
                        continue
                    hs, _ = orc.seq_to_hashes(bytes(g), kk, force=True, keep_zeros=True)
                    hs = np.asarray(hs, dtype=np.uint64)
                    keep = hs[(hs != 0) & (hs <= np.uint64(max_hash))]
                    want_h, want_a = np.unique(keep, return_counts=True)
                    assert np.array_equal(hashes, want_h) and np.array_equal(abunds, want_a.astype(np.uint64)), (max_hash, k, gi, kk)
    hs, _ = orc.seq_to_hashes(bytes(genomes[3]), 21, force=True, keep_zeros=True)
    assert int((np.asarray(hs, dtype=np.uint64) != 0).sum()) > 16384      # that row took the big-row path when all were kept


def test_simt_tile_kernels_match_oracle(simt):
    """pairwise_tile_split_kernel (the pair-by-pair intersection kernel) and its u64 predecessor as launched:
    table build in shared memory behind barriers, streamed rows, crowded buckets, UINT64_MAX handling, symmetric
    (upper triangle) and A x B modes, every TA, column blocks."""
    from sourmash_b200.synth import synth_sketches
    rng = np.random.default_rng(17)
    h, off = synth_sketches(26, mean=400, sd=100, lo=5, hi=900, n_families=3, pool=520, seed=91)
    fam = [h[int(off[i]):int(off[i + 1])] for i in range(26)]
    big = np.uint64(2**64 - 1)
    edge = [np.zeros(0, np.uint64), np.array([0], np.uint64), np.array([0, 1, 2, 3, big], np.uint64), np.array([big], np.uint64),
            np.array([big - 1, big], np.uint64), np.arange(1, 500, dtype=np.uint64),
            np.unique(rng.integers(0, 2**63, size=700, dtype=np.uint64)),
            np.array([(7 << 32) | 5, (8 << 32) | 5, (9 << 32) | 5, (9 << 32) | 6], dtype=np.uint64),
            np.array([(8 << 32) | 5, (9 << 32) | 6, (10 << 32) | 5], dtype=np.uint64),
            np.unique(rng.integers(0, 1000, size=300, dtype=np.uint64))]
    for rows in (fam, edge):
        n = len(rows)
        hh, oo = orc.to_csr(rows)
        want = orc.pairwise_common(hh, oo)
        iu = np.triu_indices(n, 1)
        for ta, variant, threads, cols in ((4, 1, 256, 512), (3, 1, 128, 7), (1, 1, 64, 512), (2, 0, 128, 512), (4, 2, 64, 5)):
            got = simt("tile", rows, ta, variant, threads, cols, 1).reshape(n, n)
            assert np.array_equal(got[iu], want[iu]), ("symmetric", n, ta, variant, threads, cols)
            nA = n // 2
            got = simt("tile", rows, ta, variant, threads, cols, 0).reshape(nA, n - nA)
            cross = np.array([[orc.count_common(rows[i], rows[nA + j]) for j in range(n - nA)] for i in range(nA)], dtype=np.uint32)
            assert np.array_equal(got, cross), ("AxB", n, ta, variant)


def test_simt_pair_kernels_match_oracle(simt):
    """generic pair kernel, bottom-k kernel, angular kernels and the finalize kernels as launched: the float64
    matrices of compare (scaled and num) bit for bit, angular within 1e-12."""
    from sourmash_b200.synth import synth_sketches
    rng = np.random.default_rng(19)
    h, off = synth_sketches(28, mean=150, sd=40, lo=0, hi=300, n_families=3, pool=200, seed=97)
    rows = [h[int(off[i]):int(off[i + 1])] for i in range(28)]
    rows[5] = np.zeros(0, np.uint64)
    rows[9] = rows[8].copy()
    rows[11] = np.unique(rng.integers(0, 2**64 - 1, size=400, dtype=np.uint64))
    n = len(rows)
    hh, oo = orc.to_csr(rows)
    for num in (50, 500):
        got = simt("pairs", rows, num, dtype=np.float64).reshape(3, n, n)
        assert np.array_equal(got[0], orc.compare_all_pairs(hh, oo, nthreads=2))
        if num == 500:                                                       # every row is a valid num=500 sketch
            assert np.array_equal(got[1], orc.compare_all_pairs(hh, oo, num=num, nthreads=2))
        ab = [1 + (r % np.uint64(5)) for r in rows]
        for i in range(n):
            for j in range(n):
                want = 1.0 if i == j else orc.angular_similarity(rows[i], ab[i], rows[j], ab[j])
                assert abs(got[2][i, j] - want) < 1e-12, (i, j)
