// join_walk.cuh -- the per-element logic of the inverted join (compare_kernels.cu), shared with
// tests/host_emul/join_emul.cu so that the CPU-only suite can check it against the oracle.
//
// Input: the (hash, row) pairs of a sketch set sorted by hash, stably (rows ascend inside a group
// of equal hashes).  Element p pairs with every later element of its group; ids[p] < ids[b]
// because a row holds each hash once, so (ids[p], ids[b]) is a cell of the upper triangle.
#pragma once
#include "common.cuh"

namespace smb {

// calls emit(row_a, row_b) once for every pair formed by element p
template <class Emit>
__host__ __device__ __forceinline__ void join_walk(const u64* __restrict__ keys, const u32* __restrict__ ids,
                                                   u64 T, u64 p, Emit&& emit) {
    const u64 k = keys[p];
    u64 b = p + 1;
    if (b >= T || keys[b] != k) return;
    const u32 a = ids[p];
    do {
        emit(a, ids[b]);
        ++b;
    } while (b < T && keys[b] == k);
}

// Row-block passes (experimental, SMB_COMPARE_PASSES): the walk of element p only touches row
// ids[p] of the matrix, so a pass that walks the elements whose row lies in [r0, r1) completes
// exactly the cells (i, j > i) with i in [r0, r1).  After the passes for rows [0, r1) every cell
// with min(i, j) < r1 is final, i.e. the full rows [0, r1) of the symmetric result are known.
template <class Emit>
__host__ __device__ __forceinline__ void join_walk_rows(const u64* __restrict__ keys, const u32* __restrict__ ids,
                                                        u64 T, u64 p, u32 r0, u32 r1, Emit&& emit) {
    const u32 a = ids[p];
    if (a < r0 || a >= r1) return;
    join_walk(keys, ids, T, p, emit);
}

// size of the group that starts at p (0 if p is not the first element of its group)
__host__ __device__ __forceinline__ u64 join_group_size_at_head(const u64* __restrict__ keys, u64 T, u64 p) {
    const u64 k = keys[p];
    if (p > 0 && keys[p - 1] == k) return 0;
    u64 m = 1;
    while (p + m < T && keys[p + m] == k) ++m;
    return m;
}

// key range of shard `shard` of `n_shards` over [0, max_key]: [lo, lo + step), the last one unbounded
__host__ __device__ __forceinline__ void join_shard_range(u64 max_key, int shard, int n_shards, u64& lo, u64& hi,
                                                          bool& bounded) {
    const u64 step = max_key / (u64)n_shards + 1;
    lo = (u64)shard * step;
    hi = lo + step;
    bounded = shard + 1 < n_shards;
}

}  // namespace smb

// ---------------------------------------------------------------------------------------------
// Experimental "cluster" layout (SMB_JOIN_LAYOUT=cluster, off by default):
// rows are renumbered so that related rows get adjacent ranks, and one *warp* walks the group of
// an element, lane l taking the (l+1)-th later element of the current 32-element chunk.  The cells
// a warp instruction touches are then (rank_a, consecutive ranks): a few sectors instead of 32.
// ---------------------------------------------------------------------------------------------
namespace smb {

// true if element p belongs to a group of two or more rows (its hash is shared)
__host__ __device__ __forceinline__ bool join_is_shared(const u64* __restrict__ keys, u64 T, u64 p) {
    const u64 k = keys[p];
    return (p > 0 && keys[p - 1] == k) || (p + 1 < T && keys[p + 1] == k);
}

// lane `lane` of the warp that owns element p, chunk starting at b0 (= p + 1 + 32 * iteration):
// emits the pair if its element is still in p's group; returns whether it was.
template <class Emit>
__host__ __device__ __forceinline__ bool join_walk_lane(const u64* __restrict__ keys, const u32* __restrict__ ids,
                                                        u64 T, u64 p, u64 b0, u32 lane, Emit&& emit) {
    const u64 b = b0 + lane;
    if (b >= T || keys[b] != keys[p]) return false;
    emit(ids[p], ids[b]);
    return true;
}

// cell of the rank-space matrix that holds the count of rows (i, j), i != j
__host__ __device__ __forceinline__ void join_rank_cell(const u32* __restrict__ inv, u32 i, u32 j, u32& lo, u32& hi) {
    const u32 ri = inv[i], rj = inv[j];
    lo = ri < rj ? ri : rj;
    hi = ri < rj ? rj : ri;
}

}  // namespace smb
