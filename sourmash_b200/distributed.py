"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

The reference has no distributed path at all; sharding follows SURVEY §8e:
  * sketching shards by genome, followed by ONE variable-length all-gather of the shards;
  * compare replicates the gathered sketches, deals row tiles cyclically to ranks, sums the
    partial count matrices with an all-reduce, and every rank finalises a block of rows.
Host-side logic (shard bounds, padded all-gather + compaction, offsets) is backend agnostic
and is what the gloo tests exercise; the compute calls need a GPU.
"""
import os

import numpy as np


def shard_bounds(n, world):
    """Contiguous row blocks: rank r owns rows [b[r], b[r+1])."""
    return [n * r // world for r in range(world + 1)]


def allgather_csr(torch, dist, local_hashes, local_sizes, device):
    """All-gather variable-length CSR shards (rank order == row order).

    local_hashes: int64 tensor (hashes viewed as int64) on `device`; local_sizes: int64 tensor.
    Returns (hashes int64 tensor [total], sizes int64 numpy [n_rows_total]).
    """
    world = dist.get_world_size()
    meta = torch.tensor([local_hashes.numel(), local_sizes.numel()], dtype=torch.int64, device=device)
    metas = torch.empty(2 * world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(metas, meta)
    metas = metas.cpu().numpy().reshape(world, 2)
    max_h, max_r = int(metas[:, 0].max()), int(metas[:, 1].max())
    pad_h = torch.zeros(max(max_h, 1), dtype=torch.int64, device=device)
    pad_h[: local_hashes.numel()] = local_hashes
    pad_s = torch.zeros(max(max_r, 1), dtype=torch.int64, device=device)
    pad_s[: local_sizes.numel()] = local_sizes
    all_h = torch.empty(world * pad_h.numel(), dtype=torch.int64, device=device)
    all_s = torch.empty(world * pad_s.numel(), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(all_h, pad_h)
    dist.all_gather_into_tensor(all_s, pad_s)
    total = int(metas[:, 0].sum())
    hashes = torch.empty(max(total, 1), dtype=torch.int64, device=device)
    sizes = []
    pos = 0
    all_s_host = all_s.cpu().numpy()
    for r in range(world):
        nh, nr = int(metas[r, 0]), int(metas[r, 1])
        hashes[pos:pos + nh] = all_h[r * pad_h.numel(): r * pad_h.numel() + nh]
        sizes.append(all_s_host[r * pad_s.numel(): r * pad_s.numel() + nr])
        pos += nh
    return hashes[:total], np.concatenate(sizes) if sizes else np.zeros(0, np.int64)


def allgather_sketchset(torch, dist, B, sset):
    """All-gather a per-rank SketchSet into a full SketchSet on every rank."""
    device = torch.device("cuda", torch.cuda.current_device())
    off = sset.offsets()
    local = torch.empty(max(int(off[-1]), 1), dtype=torch.int64, device=device)
    sset.copy_to_device(local.data_ptr())
    sizes = torch.from_numpy(np.diff(off.astype(np.int64))).to(device)
    hashes, all_sizes = allgather_csr(torch, dist, local[: int(off[-1])], sizes, device)
    h_off = np.zeros(len(all_sizes) + 1, dtype=np.uint64)
    h_off[1:] = np.cumsum(all_sizes)
    d_off = torch.from_numpy(h_off.view(np.int64)).to(device)
    return B.SketchSet.from_device(hashes.data_ptr(), d_off.data_ptr(), h_off, keepalive=(hashes, d_off))


class CompareShard:
    """One rank's part of an N-GPU all-vs-all compare (see module docstring)."""

    def __init__(self, torch, dist, B, hashes, offsets, rank, world):
        self.torch, self.dist, self.B, self.rank, self.world = torch, dist, B, rank, world
        n = len(offsets) - 1
        self.n = n
        self.bounds = shard_bounds(n, world)
        lo, hi = self.bounds[rank], self.bounds[rank + 1]
        self.device = torch.device("cuda", torch.cuda.current_device())
        local = hashes[int(offsets[lo]):int(offsets[hi])].view(np.int64)
        sizes = np.diff(offsets.astype(np.int64))[lo:hi]
        self.pin_h = torch.from_numpy(local.copy()).pin_memory()
        self.pin_s = torch.from_numpy(sizes.copy()).pin_memory()
        self.d_local = self.pin_h.to(self.device)
        self.d_sizes = self.pin_s.to(self.device)
        # experimental (SMB_JOIN_LAYOUT=stripe): every rank counts only its own block of rows, so there is
        # no partial count matrix and no all-reduce
        self.rows_direct = os.environ.get("SMB_JOIN_LAYOUT") == "stripe"
        self.d_common = None if self.rows_direct else torch.zeros((n, n), dtype=torch.int32, device=self.device)
        self.d_out = torch.empty((hi - lo, n), dtype=torch.float64, device=self.device)
        self.pin_out = torch.empty((hi - lo, n), dtype=torch.float64).pin_memory()
        self.h2d_bytes = int(local.nbytes + sizes.nbytes)
        self.d2h_bytes = int((hi - lo) * n * 8)

    def step(self, e2e):
        torch, dist, B = self.torch, self.dist, self.B
        if e2e:
            self.d_local.copy_(self.pin_h, non_blocking=True)
            self.d_sizes.copy_(self.pin_s, non_blocking=True)
        hashes, sizes = allgather_csr(torch, dist, self.d_local, self.d_sizes, self.device)
        h_off = np.zeros(self.n + 1, dtype=np.uint64)
        h_off[1:] = np.cumsum(sizes)
        d_off = torch.from_numpy(h_off.view(np.int64)).to(self.device)
        sset = B.SketchSet.from_device(hashes.data_ptr(), d_off.data_ptr(), h_off, keepalive=(hashes, d_off))
        lo, hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        if self.rows_direct:
            B.compare_jaccard_rows_device(sset, lo, hi, self.d_out.data_ptr())
        else:
            self.d_common.zero_()
            B.pairwise_counts_shard_device(sset, self.rank, self.world, self.d_common.data_ptr())
            dist.all_reduce(self.d_common)
            B.finalize_jaccard_rows_device(sset, self.d_common.data_ptr(), lo, hi, self.d_out.data_ptr())
        if e2e:
            self.pin_out.copy_(self.d_out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return float(self.pin_out[0, 1 if self.n > 1 else 0])
        return None


class ShardedDatabase:
    """A sketch database sharded by subject over the ranks: every rank keeps its
    block of rows in HBM; a query is replicated.

    * ``search_counts(query)``: local one-vs-many, then one all-gather of the u32 counts.
    * ``gather(query, threshold)``: the CounterGather rounds with two tiny collectives per
      round -- all-gather of (best count, global row), broadcast of the winner's intersection.
    Works with any backend for the collectives (NCCL on GPUs; tensors for gloo live on the CPU).
    """

    def __init__(self, torch, dist, B, local_rows_sset, n_rows_total, row_begin):
        self.torch, self.dist, self.B = torch, dist, B
        self.sset, self.n_total, self.row_begin = local_rows_sset, n_rows_total, row_begin
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" \
            else torch.device("cpu")

    def build_index(self):
        "Inverted index of this rank's block of rows (batch.SketchSet.build_index): counts and gather rounds probe it."
        return self.sset.build_index() if len(self.sset) else 0

    def search_counts(self, query):
        "|query ∩ S_j| for every row of the whole database, on every rank."
        torch, dist = self.torch, self.dist
        local = self.B.one_vs_many(query, self.sset).astype(np.int64)
        sizes = torch.tensor([len(local)], dtype=torch.int64, device=self.device)
        all_sizes = torch.empty(self.world, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(all_sizes, sizes)
        all_sizes = all_sizes.cpu().numpy()
        pad = torch.zeros(int(all_sizes.max()) if len(all_sizes) else 1, dtype=torch.int64, device=self.device)
        pad[: len(local)] = torch.from_numpy(local).to(self.device)
        out = torch.empty(self.world * pad.numel(), dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(out, pad)
        out = out.cpu().numpy().reshape(self.world, -1)
        return np.concatenate([out[r, : all_sizes[r]] for r in range(self.world)]).astype(np.uint32)

    def search_counts_device(self, d_query, d_local_counts, d_all_counts=None):
        """The same with the query (int64 view of the sorted u64 hashes) and the counters resident on the device: the
        local pass writes `d_local_counts` (int32 [local rows]); with `d_all_counts` (int32 [world * max local rows])
        one all-gather of equally sized blocks follows (shards of shard_bounds() differ by at most one row; a rank with
        fewer rows leaves its last slot unused).  No host round trip."""
        self.B.one_vs_many_device(d_query.data_ptr(), d_query.numel(), self.sset, d_local_counts.data_ptr())
        if d_all_counts is not None:
            self.dist.all_gather_into_tensor(d_all_counts, d_local_counts)
        return d_all_counts

    def gather(self, query, threshold=1, max_rounds=None):
        "Returns (global match rows, intersect sizes) in pick order -- identical on every rank."
        torch, dist = self.torch, self.dist
        threshold = max(int(threshold), 1)
        session = self.B.GatherSession(query, self.sset, threshold)
        ids, sizes = [], []
        max_rounds = self.n_total if max_rounds is None else max_rounds
        while len(ids) < max_rounds:
            cnt, row = session.peek() if len(self.sset) else (0, 0)
            mine = torch.tensor([cnt, self.row_begin + row], dtype=torch.int64, device=self.device)
            allv = torch.empty(2 * self.world, dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(allv, mine)
            allv = allv.cpu().numpy().reshape(self.world, 2)
            best = int(allv[:, 0].max())
            if best < threshold:
                break
            # ties: lowest global row (== first inserted in the reference's Counter)
            cand = allv[allv[:, 0] == best]
            grow = int(cand[:, 1].min())
            owner = int(np.nonzero((allv[:, 0] == best) & (allv[:, 1] == grow))[0][0])
            if owner == self.rank:
                isect = session.intersect(grow - self.row_begin)
                n = torch.tensor([len(isect)], dtype=torch.int64, device=self.device)
            else:
                isect, n = None, torch.zeros(1, dtype=torch.int64, device=self.device)
            dist.broadcast(n, src=owner)
            buf = torch.empty(int(n.item()), dtype=torch.int64, device=self.device)
            if owner == self.rank and len(isect):
                buf.copy_(torch.from_numpy(isect.view(np.int64)))
            if buf.numel():
                dist.broadcast(buf, src=owner)
            isect = buf.cpu().numpy().view(np.uint64)
            ids.append(grow)
            sizes.append(len(isect))
            if session.apply(isect) == 0:
                break
        return np.array(ids, dtype=np.uint32), np.array(sizes, dtype=np.uint32)
