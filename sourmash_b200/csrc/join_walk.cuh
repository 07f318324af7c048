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
