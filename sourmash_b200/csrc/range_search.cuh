// range_search.cuh -- the range-partitioned one-vs-many pass (SMB_SEARCH_LAYOUT=ranges; compare_kernels.cu
// launches it): helpers shared with tests/host_emul/ranges_emul.cu and the kernels themselves, which
// tests/host_emul/simt_emul.cu runs on the CPU against the oracle.
//
// The default pass (one_vs_many_global_kernel) answers every subject element with a random probe of
// a query bitmap that lives in L2: 1.5e9 single-sector L2 reads for a 12 GB database, several times
// the cost of streaming the database itself.  Hashes are uniform, so here the key space [0, max_key]
// is cut into P equal ranges; CTA p keeps the query bitmap of range p in shared memory and streams,
// for every subject row, the slice of the row that falls into its range (rows are sorted: a slice
// is contiguous, its bounds come from a table built once per resident set).  Probes hit shared
// memory; only bitmap hits go on to the key compare.
#pragma once
#include "common.cuh"

namespace smb {

// width of one of P >= 2 equal ranges over [0, max_key] (no overflow: max_key / 2 + 1 <= 2^63); every key
// <= max_key of the last range satisfies key - lo < width as well
__host__ __device__ __forceinline__ u64 range_width(u64 max_key, int P) { return max_key / (u64)P + 1; }

// first index i of the sorted row with row[i] >= key
__host__ __device__ __forceinline__ u64 range_lower_bound(const u64* __restrict__ row, u64 n, u64 key) {
    u64 lo = 0, hi = n;
    while (lo < hi) {
        const u64 mid = (lo + hi) >> 1;
        if (row[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// bound p (0..P) of a row: where range p starts; bound P is the row length (p * width may overflow)
__host__ __device__ __forceinline__ u64 range_bound(const u64* __restrict__ row, u64 n, u64 width, int p, int P) {
    if (p <= 0) return 0;
    if (p >= P) return n;
    return range_lower_bound(row, n, (u64)p * width);      // p < P: p * width <= max_key + P, no overflow for max_key < 2^64 - P
}

// smallest shift such that the bits of one range fit into `max_bits`
__host__ __device__ __forceinline__ u32 range_bitmap_shift(u64 width, u64 max_bits) {
    u32 s = 0;
    while (s < 63 && ((width - 1) >> s) + 1 > max_bits) ++s;
    return s;
}

// bit of key x inside the bitmap of the range that starts at lo (x >= lo)
__host__ __device__ __forceinline__ u64 range_bit(u64 x, u64 lo, u32 bm_shift) { return (x - lo) >> bm_shift; }

}  // namespace smb
