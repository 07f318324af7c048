"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: shard bounds, the
variable-length all-gather of CSR shards and the bookkeeping around it.  The compute calls
themselves need GPUs and are exercised by bench.py --gpus N."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sourmash_b200.distributed import allgather_csr, shard_bounds
        from sourmash_b200.synth import synth_sketches
        h, off = synth_sketches(37, mean=200, sd=40, lo=0, hi=400, n_families=3, pool=250, seed=5)
        bounds = shard_bounds(37, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        local = torch.from_numpy(h[int(off[lo]):int(off[hi])].view(np.int64).copy())
        sizes = torch.from_numpy(np.diff(off.astype(np.int64))[lo:hi].copy())
        hashes, all_sizes = allgather_csr(torch, dist, local, sizes, torch.device("cpu"))
        ok = (np.array_equal(hashes.numpy().view(np.uint64), h)
              and np.array_equal(all_sizes, np.diff(off.astype(np.int64))))
        # an empty shard on one rank must also work
        e_local = local if rank == 0 else local[:0]
        e_sizes = sizes if rank == 0 else sizes[:0]
        eh, es = allgather_csr(torch, dist, e_local, e_sizes, torch.device("cpu"))
        ok = ok and eh.numel() == local.numel() if rank == 0 else ok
        q.put((rank, bool(ok), int(hashes.numel()), len(all_sizes)))
    finally:
        dist.destroy_process_group()


class _FakeSet:
    "host-only stand-in for batch.SketchSet (rows kept in numpy)"
    def __init__(self, rows):
        self._rows = rows
    def __len__(self):
        return len(self._rows)
    def sizes(self):
        return np.array([len(r) for r in self._rows], dtype=np.int64)
    def take_rows(self, rows):
        return _FakeSet([self._rows[int(i)] for i in rows])
    def to_host(self):
        import oracle as orc
        return orc.to_csr(self._rows)


class _FakeBatch:
    """oracle-backed stand-in for sourmash_b200.batch so that the collective logic of
    ShardedDatabase (tie-breaking, ownership, broadcasts) runs on CPU/gloo."""
    @staticmethod
    def one_vs_many(query, sset):
        import oracle as orc
        return np.array([orc.count_common(query, r) for r in sset._rows], dtype=np.uint32)

    class SketchSet:
        @staticmethod
        def from_host(h, off):
            return _FakeSet([h[int(off[i]):int(off[i + 1])] for i in range(len(off) - 1)])

    @staticmethod
    def gather(query, sset, threshold=1, max_rounds=None):
        "single-process CounterGather rounds: largest remaining overlap first, lowest row on ties"
        import oracle as orc
        rows, cur = sset._rows, np.array(query, dtype=np.uint64)
        cnt = np.array([orc.count_common(cur, r) for r in rows], dtype=np.int64)
        ids, sizes = [], []
        while len(cnt) and (max_rounds is None or len(ids) < max_rounds):
            j = int(np.argmax(cnt))
            if cnt[j] < max(threshold, 1):
                break
            isect = np.intersect1d(cur, rows[j])
            ids.append(j); sizes.append(len(isect))
            cnt = cnt - np.array([orc.count_common(isect, r) for r in rows], dtype=np.int64)
            cur = np.setdiff1d(cur, isect)
            if not len(cur):
                break
        return np.array(ids, dtype=np.uint32), np.array(sizes, dtype=np.uint32)

    class GatherSession:
        def __init__(self, query, sset, min_count=1):
            import oracle as orc
            self.q = np.array(query, dtype=np.uint64)
            self.rows = sset._rows
            self.counts = np.array([orc.count_common(self.q, r) for r in self.rows], dtype=np.int64)
        def peek(self):
            if not len(self.counts):
                return 0, 0
            j = int(np.argmax(self.counts))
            return int(self.counts[j]), j
        def intersect(self, row):
            return np.intersect1d(self.q, self.rows[row])
        def apply(self, isect):
            import oracle as orc
            self.counts -= np.array([orc.count_common(isect, r) for r in self.rows], dtype=np.int64)
            self.q = np.setdiff1d(self.q, isect)
            return len(self.q)


def _sharded_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle as orc
        from sourmash_b200.distributed import ShardedDatabase, shard_bounds
        from sourmash_b200.synth import rows_of, synth_sketches
        h, off = synth_sketches(31, mean=300, sd=50, lo=100, hi=500, n_families=3, pool=380, seed=8)
        rows = rows_of(h, off)
        rows[17] = rows[4].copy()                       # an exact tie across shards: lowest global row must win
        b = shard_bounds(31, world)
        db = ShardedDatabase(torch, dist, _FakeBatch, _FakeSet(rows[b[rank]:b[rank + 1]]), 31, b[rank])
        query = np.unique(np.concatenate([rows[4], rows[9][:150], rows[25][50:250], rows[30][::2]]))
        counts = db.search_counts(query)
        ok = np.array_equal(counts, np.array([orc.count_common(query, r) for r in rows], dtype=np.uint32))
        ids, sizes = db.gather(query, threshold=5)
        # single-process reference loop
        cur, cnt, want = query.copy(), np.array([orc.count_common(query, r) for r in rows]), []
        while True:
            j = int(np.argmax(cnt))
            if cnt[j] < 5:
                break
            isect = np.intersect1d(cur, rows[j])
            want.append((j, len(isect)))
            cnt = cnt - np.array([orc.count_common(isect, r) for r in rows])
            cur = np.setdiff1d(cur, isect)
            if not len(cur):
                break
        ok = ok and list(zip(ids.tolist(), sizes.tolist())) == want and ids[0] == 4
        ids2, sizes2 = db.gather_sharded_rounds(query, threshold=5)          # the loop with sharded counters: same picks
        ok = ok and list(zip(ids2.tolist(), sizes2.tolist())) == want
        q.put((rank, bool(ok), len(ids)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_search_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert all(r[1] for r in results), results
    assert results[0][2] == results[1][2] >= 3


def test_shard_bounds():
    sys.path.insert(0, ROOT)
    from sourmash_b200.distributed import shard_bounds
    assert shard_bounds(10, 1) == [0, 10]
    assert shard_bounds(10, 4) == [0, 2, 5, 7, 10]
    b = shard_bounds(10000, 8)
    assert b[0] == 0 and b[-1] == 10000 and all(b[i + 1] - b[i] == 1250 for i in range(8))
    assert shard_bounds(3, 8)[-1] == 3


@pytest.mark.timeout(120)
def test_allgather_csr_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == [0, 1]
    assert all(r[1] for r in results), results
    assert results[0][2] == results[1][2] and results[0][3] == 37
