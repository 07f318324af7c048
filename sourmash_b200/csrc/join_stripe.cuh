// join_stripe.cuh -- lane-level logic of the "stripe" layout of the inverted join (experimental,
// SMB_JOIN_LAYOUT=stripe, off by default; compare_kernels.cu), shared with
// tests/host_emul/join_emul.cu so that the CPU-only suite checks it against the oracle.
//
// The plain join (join_walk.cuh) sends one global reduction per (shared hash, pair of rows) to a
// count matrix in HBM, and the rate of those reductions is its limit.  Here a CTA owns R complete
// rows of the result as a stripe of counters in shared memory: for every hash of its rows it visits
// the hash's whole group in the sorted stream and bumps stripe[row][other row] with shared-memory
// atomics, then turns the finished stripe into float64 Jaccard values and writes each output row
// once.  No count matrix in HBM, no global atomics, no separate zero / finalize passes, and rows
// finish in CTA order (row blocks can be downloaded while later ones are still being counted).
//
// Stream layout: the set's hashes sorted by value with their CSR element index as payload;
//   tags[q] = row of sorted element q, bit 31 set on the first element of a group of equal hashes;
//   pos[e]  = sorted position of CSR element e (row r owns e in [off[r], off[r + 1])).
// The group of element q = the run around q between two head flags; a warp scans it 32 tags at a
// time in both directions, the ballot of the "stop" predicate cuts the chunk at the group's end.
#pragma once
#include "common.cuh"

namespace smb {

static constexpr u32 STRIPE_HEAD = 0x80000000u;

// row that owns CSR element e: largest r with off[r] <= e (rows may be empty)
__host__ __device__ __forceinline__ u32 stripe_row_of(const u64* __restrict__ off, int n, u64 e) {
    int lo = 0, hi = n;                                   // invariant: off[lo] <= e < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= e) lo = mid; else hi = mid;
    }
    return (u32)lo;
}

__host__ __device__ __forceinline__ u32 stripe_make_tag(const u64* __restrict__ sorted_keys, u64 q, u32 row) {
    const bool head = q == 0 || sorted_keys[q] != sorted_keys[q - 1];
    return row | (head ? STRIPE_HEAD : 0u);
}

// Forward chunk `it` of element q, lane `lane`: looks at sorted element q + 1 + 32 it + lane.
// Returns the stop predicate (end of stream, or the first element of the next group).
__host__ __device__ __forceinline__ bool stripe_fwd_stop(const u32* __restrict__ tags, u64 T, u64 q, u32 it, u32 lane,
                                                         u32& tag) {
    const u64 b = q + 1 + 32ull * it + lane;
    if (b >= T) { tag = 0; return true; }
    tag = tags[b];
    return (tag & STRIPE_HEAD) != 0;
}
// Backward chunk: looks at q - 1 - 32 it - lane.  The head of the group stops the scan but belongs
// to the group; `valid` is false in front of the stream.
__host__ __device__ __forceinline__ bool stripe_bwd_stop(const u32* __restrict__ tags, u64 q, u32 it, u32 lane, u32& tag,
                                                         bool& valid) {
    const u64 d = 1 + 32ull * it + lane;
    if (d > q) { tag = 0; valid = false; return true; }
    tag = tags[q - d];
    valid = true;
    return (tag & STRIPE_HEAD) != 0;
}
// stop_mask = ballot of the stop predicate over the warp
__host__ __device__ __forceinline__ bool stripe_fwd_active(u32 stop_mask, u32 lane) {   // no stop at lanes <= lane
    return (stop_mask & ((2u << lane) - 1u)) == 0;
}
__host__ __device__ __forceinline__ bool stripe_bwd_active(u32 stop_mask, u32 lane, bool valid) {   // no stop at lanes < lane
    return valid && (stop_mask & ((1u << lane) - 1u)) == 0;
}
__host__ __device__ __forceinline__ bool stripe_continue(u32 stop_mask) { return stop_mask == 0; }

// local row of element e inside a block whose rows start at elements s_off[0..rows] (ascending)
__host__ __device__ __forceinline__ u32 stripe_local_row(const u64* __restrict__ s_off, int rows, u64 e) {
    u32 a = 0;
    for (int r = 1; r < rows; ++r) a += (s_off[r] <= e) ? 1u : 0u;
    return a;
}

// ---- sorting on the low 32 key bits only (SMB_JOIN_SORT=low32: 4 radix passes of 4-byte keys instead
// of 7 passes of 8-byte keys) ----
// Elements with equal low words form a run; nearly every run is one group of equal hashes.  The few
// runs that mix different hashes (expected: distinct^2 / 2^33) are re-sorted on the rotated key
// (low word first, then the high word), which keeps them in place as runs and orders them inside.
__host__ __device__ __forceinline__ u64 stripe_rotated_key(u64 k) { return (k << 32) | (k >> 32); }

// length of the run of equal low words that starts at q (0 if q is not its first element), and whether
// the run holds more than one distinct key
__host__ __device__ __forceinline__ u64 stripe_run_at_head(const u32* __restrict__ low_sorted, const u64* __restrict__ keys,
                                                           u64 T, u64 q, bool& mixed) {
    mixed = false;
    const u32 lw = low_sorted[q];
    if (q > 0 && low_sorted[q - 1] == lw) return 0;
    const u64 k0 = keys[q];
    u64 m = 1;
    while (q + m < T && low_sorted[q + m] == lw) { mixed = mixed || keys[q + m] != k0; ++m; }
    return m;
}

// the finalize step of compare (finalize_rows_kernel): ones on the diagonal, common / max(1, union)
__host__ __device__ __forceinline__ double stripe_jaccard(u32 common, u64 size_i, u64 size_j, bool diagonal) {
    if (diagonal) return 1.0;
    const u64 un = size_i + size_j - common;
    return (double)common / (double)(un > 1 ? un : 1);
}

// rows per CTA for a stripe of `ncols` u32 counters per row in `smem_bytes` of shared memory
__host__ __device__ __forceinline__ int stripe_rows_per_block(size_t smem_bytes, int ncols) {
    const size_t reserve = 40 * sizeof(u64);              // the block's row offsets
    if (smem_bytes <= reserve) return 0;
    const size_t r = (smem_bytes - reserve) / ((size_t)ncols * sizeof(u32));
    return (int)(r > 32 ? 32 : r);
}

}  // namespace smb
