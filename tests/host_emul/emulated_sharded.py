"""World-size-2 run (gloo, two CPU processes) of the sharded search / gather of sourmash_b200/distributed.py with the
REAL batch module on the EMULATED library: every rank holds a block of the database as a resident set (one of them
with the inverted index), queries are replicated, counts are all-gathered, gather rounds exchange (best count, row)
and the winner's intersection.  Results must equal the single-process oracle on both ranks.
Run by tests/test_emulated_library.py."""
import os
import socket
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def worker(rank, world, port, q):
    import emulated_boot
    emulated_boot.install()
    import numpy as np
    import torch
    import torch.distributed as dist

    import oracle as orc
    from sourmash_b200 import batch as B
    from sourmash_b200.distributed import ShardedDatabase, shard_bounds
    from sourmash_b200.synth import rows_of, synth_sketches
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h, off = synth_sketches(61, mean=250, sd=50, lo=50, hi=500, n_families=3, pool=320, seed=8)
        rows = rows_of(h, off)
        rows[37] = rows[4].copy()                        # an exact tie across the two shards: the lowest global row wins
        b = shard_bounds(61, world)
        local = B.SketchSet.from_rows(rows[b[rank]:b[rank + 1]])
        db = ShardedDatabase(torch, dist, B, local, 61, b[rank])
        if rank == 1:
            db.build_index()                             # mixed: one rank probes its index, the other streams its rows
        query = np.unique(np.concatenate([rows[4], rows[9][:150], rows[45][50:250], rows[60][::2]]))
        counts = db.search_counts(query)
        ok = np.array_equal(counts, np.array([orc.count_common(query, r) for r in rows], dtype=np.uint32))
        ids, sizes = db.gather(query, threshold=5)
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
        ok = ok and list(zip(ids.tolist(), sizes.tolist())) == want and ids[0] == 4 and len(want) >= 4
        q.put((rank, bool(ok), len(ids)))
    finally:
        dist.destroy_process_group()


def main():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in results), results
    assert results[0][2] == results[1][2]
    print("emulated sharded search / gather passed on 2 ranks (%d rounds)" % results[0][2])


if __name__ == "__main__":
    main()
