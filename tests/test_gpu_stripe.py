"""GPU parity tests of the stripe layout of the inverted join -- the default all-vs-all path of compare
(csrc/join_stripe.cuh): resident, host (row blocks + copies) and block-of-rows entry points, both tag
widths, both count modes, the global-reduction join behind it, edge rows, groups longer than a warp, and
hashes a 32-bit sort key cannot tell apart.  Everything is compared with the oracle bit for bit; the same
kernels run on the CPU in tests/test_host_emulation.py::test_simt_stripe_*."""
import numpy as np
import pytest

import oracle as orc
from sourmash_b200.synth import rows_of, synth_sketches

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    from sourmash_b200 import batch
    return batch


class _DeviceMatrix:
    """float64 device buffer for the *_device entry points: a torch CUDA tensor on a GPU box; plain host memory
    when the suite runs against the emulated library (tests/host_emul/run_gpu_tests_emulated.py)."""

    def __init__(self, shape):
        import torch
        if torch.cuda.is_available():
            self._t = torch.empty(shape, dtype=torch.float64, device="cuda")
            self.ptr = self._t.data_ptr()
        else:
            self._t = None
            self._a = np.full(shape, -1.0)
            self.ptr = self._a.ctypes.data

    def numpy(self):
        if self._t is None:
            return self._a
        import torch
        torch.cuda.synchronize()
        return self._t.cpu().numpy()


def _edge_rows():
    rng = np.random.Generator(np.random.PCG64(77))
    big = np.uint64(2**64 - 1)
    rows = [np.zeros(0, np.uint64), np.array([0], np.uint64), np.array([0, 1, 2, big - 1, big], np.uint64),
            np.array([big], np.uint64), np.arange(1, 300, dtype=np.uint64), np.arange(1, 300, dtype=np.uint64),
            np.unique(rng.integers(0, 2**64 - 1, size=500, dtype=np.uint64))]
    rows += [np.array([7, 1000 + i], dtype=np.uint64) for i in range(100)]      # one hash shared by 100 rows
    return rows


def _check_all_entry_points(B, h, off, want=None):
    n = len(off) - 1
    want = orc.compare_all_pairs(h, off, nthreads=8) if want is None else want
    sset = B.SketchSet.from_host(h, off)
    assert np.array_equal(B.compare_jaccard(sset), want)                     # host path: row blocks + copies
    d_out = _DeviceMatrix((n, n))
    B.compare_jaccard_device(sset, d_out.ptr)                                # resident path
    assert np.array_equal(d_out.numpy(), want)
    lo, hi = n // 3, min(n, n // 3 + 37)
    d_rows = _DeviceMatrix((hi - lo, n))
    B.compare_jaccard_rows_device(sset, lo, hi, d_rows.ptr)                 # a block of complete rows (multi-GPU unit)
    assert np.array_equal(d_rows.numpy(), want[lo:hi])


@pytest.mark.parametrize("layout,tags", [(None, None), (None, "u32"), ("stripe_full", None), ("plain", None)])
@pytest.mark.parametrize("n,fam", [(96, 6), (700, 9), (1500, 12)])
def test_stripe_layout_matches_oracle(B, monkeypatch, n, fam, layout, tags):
    monkeypatch.setenv("SMB_COMPARE_ALGO", "join")
    if layout:
        monkeypatch.setenv("SMB_JOIN_LAYOUT", layout)
    if tags:
        monkeypatch.setenv("SMB_STRIPE_TAGS", tags)
    h, off = synth_sketches(n, mean=400, sd=80, lo=0, hi=800, n_families=fam, pool=500, seed=n)
    _check_all_entry_points(B, h, off)


@pytest.mark.parametrize("layout", [None, "stripe_full"])
def test_stripe_layout_edge_rows(B, monkeypatch, layout):
    monkeypatch.setenv("SMB_COMPARE_ALGO", "join")
    if layout:
        monkeypatch.setenv("SMB_JOIN_LAYOUT", layout)
    rows = _edge_rows() * 12                                                  # > 1024 rows: the host path takes the join
    h, off = orc.to_csr(rows)
    _check_all_entry_points(B, h, off)


def test_stripe_repairs_runs_of_equal_sort_keys(B, monkeypatch):
    """The stream is sorted on the top 32 significant bits of the hashes; hashes that agree there and differ
    below form one run of the sort; stripe_tag_kernel notices such runs and stripe_fix_kernel redoes them in order.
    Planted: scaled=1000-sized keys (22 low bits) and full 64-bit keys (32 low bits), interleaved over rows."""
    monkeypatch.setenv("SMB_COMPARE_ALGO", "join")
    h, off = synth_sketches(1200, mean=300, sd=60, lo=0, hi=600, n_families=8, pool=400, seed=23)
    base = rows_of(h, off)
    for top, lows in ((np.uint64(0x001234ab) << np.uint64(32), (9, 3, 7, 1, 5)),                    # < max_hash(1000)
                      (np.uint64(0xfedc0001) << np.uint64(32), (0x1234abcd, 1, 0xffffffff, 77))):   # 64-bit keys
        rows = [r.copy() for r in base]
        for i in range(0, 1200, 3):
            extra = [top | np.uint64(v) for j, v in enumerate(lows) if (i + j) % 3 != 0]
            rows[i] = np.unique(np.concatenate([rows[i], np.array(extra, dtype=np.uint64)]))
        # a long mixed run: two hashes with one sort key, each in ~170 rows
        for i in range(0, 1200, 7):
            rows[i] = np.unique(np.concatenate([rows[i], np.array([top | np.uint64(0x3ff001 + (i // 7) % 2)], dtype=np.uint64)]))
        hh, oo = orc.to_csr(rows)
        want = orc.compare_all_pairs(hh, oo, nthreads=8)
        for layout in (None, "stripe_full"):
            if layout:
                monkeypatch.setenv("SMB_JOIN_LAYOUT", layout)
            else:
                monkeypatch.delenv("SMB_JOIN_LAYOUT", raising=False)
            assert np.array_equal(B.compare_jaccard(B.SketchSet.from_host(hh, oo)), want), (hex(int(top)), layout)


def test_stripe_groups_longer_than_a_warp_and_short_keys(B, monkeypatch):
    "one hash in 3300 rows (a hundred 32-tag chunks per element); keys below 2^32 (no low bits, fewer radix passes)"
    monkeypatch.setenv("SMB_COMPARE_ALGO", "join")
    rng = np.random.Generator(np.random.PCG64(5))
    rows = [np.unique(np.concatenate([rng.integers(1, 2**20, size=int(rng.integers(0, 60)), dtype=np.uint64),
                                      np.array([424242] if i < 3300 else [], dtype=np.uint64)]))
            for i in range(3400)]
    rows[17] = np.zeros(0, np.uint64)
    h, off = orc.to_csr(rows)
    want = orc.compare_all_pairs(h, off, nthreads=8)
    sset = B.SketchSet.from_host(h, off)
    assert np.array_equal(B.compare_jaccard(sset), want)
    monkeypatch.setenv("SMB_JOIN_LAYOUT", "stripe_full")
    assert np.array_equal(B.compare_jaccard(sset), want)


def test_rows_device_on_the_global_reduction_join(B, monkeypatch):
    "smb_compare_jaccard_rows_dev when the stripe layout is switched off (whole count matrix, then the rows)."
    monkeypatch.setenv("SMB_JOIN_LAYOUT", "plain")
    h, off = synth_sketches(300, mean=300, sd=60, lo=0, hi=600, n_families=5, pool=400, seed=4)
    want = orc.compare_all_pairs(h, off, nthreads=8)
    sset = B.SketchSet.from_host(h, off)
    d_rows = _DeviceMatrix((50, 300))
    B.compare_jaccard_rows_device(sset, 120, 170, d_rows.ptr)
    assert np.array_equal(d_rows.numpy(), want[120:170])


@pytest.mark.parametrize("algo,layout", [("join", None), ("join", "plain"), ("tile", None)])
def test_key_range_shards_add_up_to_whole_row_counters(B, monkeypatch, algo, layout):
    """The multi-GPU unit on one GPU: the whole-row partial counters of 1 / 3 / 8 shards (key ranges of the sorted stream
    for the stripe layout, upper-triangle shards mirrored for the others) sum to the oracle's counts, as 32- and 16-bit
    counters; 16-bit counters added two at a time as uint32 (what the reduce-scatter does) give the same; summed counters
    of a block of rows finalise to the oracle's float64 rows."""
    import torch
    monkeypatch.setenv("SMB_COMPARE_ALGO", algo)
    if layout:
        monkeypatch.setenv("SMB_JOIN_LAYOUT", layout)
    n = 600
    h, off = synth_sketches(n, mean=400, sd=80, lo=0, hi=800, n_families=9, pool=500, seed=77)
    want = orc.pairwise_common(h, off, nthreads=8)
    jac = orc.compare_all_pairs(h, off, nthreads=8)
    sset = B.SketchSet.from_host(h, off)
    off_diag = ~np.eye(n, dtype=bool)
    for bits, tdt, ndt in ((32, torch.int32, np.uint32), (16, torch.int16, np.uint16)):
        for shards in (1, 3, 8):
            total = torch.zeros((n, n), dtype=tdt, device="cuda")
            for r in range(shards):
                part = torch.full((n, n), -1, dtype=tdt, device="cuda")            # every cell must be written
                B.compare_counts_shard_device(sset, r, shards, part.data_ptr(), bits=bits)
                if bits == 32:
                    total += part
                else:                                              # pairs of 16-bit counters added as one int32, like ncclSum on the int32 view
                    total = (total.view(torch.int32) + part.view(torch.int32)).view(torch.int16)
            got = total.cpu().numpy().view(ndt)
            assert np.array_equal(got[off_diag], want[off_diag].astype(ndt)), (algo, layout, bits, shards)
        rows = torch.full((60, n), -1.0, dtype=torch.float64, device="cuda")
        block = total[70:130].contiguous()
        B.finalize_counts_rows_device(sset, block.data_ptr(), 70, 130, rows.data_ptr(), bits=bits)
        assert np.array_equal(rows.cpu().numpy(), jac[70:130]), (algo, layout, bits)


def test_take_rows_and_device_query_entry_points(B):
    import torch
    h, off = synth_sketches(300, mean=500, sd=80, lo=0, hi=900, n_families=5, pool=600, seed=5)
    sset = B.SketchSet.from_host(h, off)
    pick = np.array([7, 0, 299, 7, 150], dtype=np.uint32)
    sub = sset.take_rows(pick)
    hh, oo = sub.to_host()
    rows = [h[int(off[i]):int(off[i + 1])] for i in pick]
    assert np.array_equal(oo, np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.uint64))
    assert np.array_equal(hh, np.concatenate(rows))
    q = np.unique(np.concatenate([rows[0], rows[2][::2]]))
    d_q = torch.from_numpy(q.view(np.int64)).cuda()
    d_c = torch.full((300,), -1, dtype=torch.int32, device="cuda")
    B.one_vs_many_device(d_q.data_ptr(), len(q), sset, d_c.data_ptr())
    assert np.array_equal(d_c.cpu().numpy().astype(np.uint64), orc.one_vs_many(q, h, off))


def _jaccard_rows_from_counts(h, off, lo, hi):
    "rows [lo, hi) of compare_serial's matrix from the oracle's one-vs-many counts (n x n does not fit for n ~ 3e4)"
    sizes = np.diff(off).astype(np.uint64)
    want = np.empty((hi - lo, len(sizes)), dtype=np.float64)
    for i in range(lo, hi):
        common = orc.one_vs_many(h[int(off[i]):int(off[i + 1])], h, off, nthreads=8)
        union = sizes[i] + sizes - common
        want[i - lo] = common.astype(np.float64) / np.maximum(union, 1).astype(np.float64)
        want[i - lo, i] = 1.0
    return want


@pytest.mark.parametrize("n,ctas,swizzle", [(27000, None, None), (29000, None, None), (1500, "1", None), (1500, None, "0"), (1483, None, None)])
def test_stripe_row_blocks_for_one_and_two_ctas_per_sm(B, monkeypatch, n, ctas, swizzle):
    """The count kernel runs two CTAs per SM when a row block of >= 1 row fits half the shared memory (n <= 28 536 columns
    of u32 counters), else one; SMB_STRIPE_CTAS=1 forces one.  27 000 columns: one-row blocks, two CTAs; 29 000: two-row
    blocks, one CTA.  SMB_STRIPE_SWIZZLE=0: counters in column order instead of stripe_col's bank-spreading order (1 483
    columns: a last partial block of 11, left in place)."""
    monkeypatch.setenv("SMB_COMPARE_ALGO", "join")
    if ctas:
        monkeypatch.setenv("SMB_STRIPE_CTAS", ctas)
    if swizzle:
        monkeypatch.setenv("SMB_STRIPE_SWIZZLE", swizzle)
    rng = np.random.Generator(np.random.PCG64(n))
    pool = rng.integers(1, 2**54, size=4000, dtype=np.uint64)
    rows = [np.unique(rng.choice(pool, size=int(rng.integers(0, 30)))) for _ in range(n)]
    h, off = orc.to_csr(rows)
    sset = B.SketchSet.from_host(h, off)
    for lo, hi in ((0, 3), (n // 2 - 1, n // 2 + 6), (n - 5, n)):
        d_rows = _DeviceMatrix((hi - lo, n))
        B.compare_jaccard_rows_device(sset, lo, hi, d_rows.ptr)
        assert np.array_equal(d_rows.numpy(), _jaccard_rows_from_counts(h, off, lo, hi)), (n, lo, hi)
