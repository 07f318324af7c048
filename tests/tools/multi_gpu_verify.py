#!/usr/bin/env python
"""Correctness of the N-GPU paths against the oracle (run under torchrun on N GPUs):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/tools/multi_gpu_verify.py

Every rank runs CompareShard on a 700-sketch problem and checks its block of rows of the
float64 matrix bit-for-bit against the oracle; the sketch shards are all-gathered and
compared with single-rank sketching."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle as orc  # noqa: E402
from sourmash_b200 import batch as B  # noqa: E402
from sourmash_b200.distributed import CompareShard, allgather_sketchset  # noqa: E402
from sourmash_b200.synth import synth_genome, synth_sketches  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B.set_device(local)
B.set_stream(torch.cuda.current_stream().cuda_stream)

# ---- compare
for n_cmp in (704, 700, 701):         # 704: equal row blocks (reduce-scatter, 16-bit counters); 700: uneven blocks at 8 ranks (all-reduce + slice); 701: odd n (32-bit counters)
    h, off = synth_sketches(n_cmp, mean=1500, sd=300, lo=100, hi=3000, n_families=7, pool=1800, seed=4)
    cs = CompareShard(torch, dist, B, h, off, rank, world)
    for _ in range(2):                # the second step reuses every buffer
        cs.step(e2e=True)
    lo, hi = cs.bounds[rank], cs.bounds[rank + 1]
    want = orc.compare_all_pairs(h, off, nthreads=8)
    got = cs.pin_out.numpy()
    assert got.shape == (hi - lo, n_cmp)
    assert np.array_equal(got, want[lo:hi]), f"rank {rank}: compare rows {lo}:{hi} of {n_cmp} differ"

# ---- sketch shards + all-gather
NG = 2 * world + 1
genomes = [synth_genome(100_000 + 1000 * g, seed=50 + g) for g in range(NG)]
mine = list(range(rank, NG, world))
seqs = np.concatenate([genomes[g] for g in mine])
offs = np.cumsum([0] + [len(genomes[g]) for g in mine]).astype(np.uint64)
sset, _ = B.sketch_sequences(seqs, offs, [21, 31], scaled=100)
full = allgather_sketchset(torch, dist, B, sset)
rows = full.rows()
order = [g for r in range(world) for g in range(r, NG, world)]          # rank-major gather order
mx = orc.max_hash_for_scaled(100)
for pos, g in enumerate(order):
    for ki, k in enumerate((21, 31)):
        assert np.array_equal(rows[pos * 2 + ki], orc.sketch_scaled(genomes[g], k, mx)), (rank, g, k)
# ---- sharded search / gather vs the single-GPU result
from sourmash_b200.distributed import ShardedDatabase, shard_bounds  # noqa: E402
from sourmash_b200.synth import rows_of  # noqa: E402
hh, oo = synth_sketches(400, mean=800, sd=100, lo=300, hi=1200, n_families=8, pool=1000, seed=12)
rws = rows_of(hh, oo)
bb = shard_bounds(400, world)
shard = B.SketchSet.from_rows(rws[bb[rank]:bb[rank + 1]])
sdb = ShardedDatabase(torch, dist, B, shard, 400, bb[rank])
query = np.unique(np.concatenate([rws[5], rws[123][:500], rws[250][100:700], rws[399][::2], rws[77][:30]]))
whole = B.SketchSet.from_host(hh, oo)
assert np.array_equal(sdb.search_counts(query), B.one_vs_many(query, whole))
big_q = np.unique(np.concatenate([query, np.random.Generator(np.random.PCG64(3)).integers(1, 2**54, size=300_000, dtype=np.uint64)]))
d_q = torch.from_numpy(big_q.view(np.int64)).cuda()
per = max(bb[i + 1] - bb[i] for i in range(world))
d_loc = torch.zeros(per, dtype=torch.int32, device="cuda")
d_all = torch.zeros(per * world, dtype=torch.int32, device="cuda")
sdb.search_counts_device(d_q, d_loc, d_all)                       # large query: range-major pass per shard, counters all-gathered on the device
got_all = d_all.cpu().numpy().reshape(world, per)
got_all = np.concatenate([got_all[r, : bb[r + 1] - bb[r]] for r in range(world)])
assert np.array_equal(got_all.astype(np.uint32), B.one_vs_many(big_q, whole))
ids, sizes = sdb.gather(query, threshold=3)
ids1, sizes1 = B.gather(query, whole, threshold=3)
assert np.array_equal(ids, ids1) and np.array_equal(sizes, sizes1) and len(ids) >= 4, (ids, ids1)
ids2, sizes2 = sdb.gather_sharded_rounds(query, threshold=3)        # the rounds with sharded counters (one collective per round)
assert np.array_equal(ids2, ids1) and np.array_equal(sizes2, sizes1)
dist.barrier()
if rank == 0:
    print(f"multi-GPU verify ok on {world} GPUs: compare rows bit-identical to the oracle, "
          f"{len(rows)} gathered sketch rows identical, sharded search/gather == single GPU ({len(ids)} rounds)")
dist.destroy_process_group()
