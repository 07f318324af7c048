// experimental_kernels.cuh -- the kernels of the inverted join (the default compare path: row slices, gather,
// estimate, count) and of the paths that sit behind switches: cluster / row-pass /
// stripe layouts of the join, range-partitioned one-vs-many pass, inverted index of a resident set.  Kept in a header so that tests/host_emul/simt_emul.cu can compile the kernels THEMSELVES for the
// host (tests/host_emul/simt.h runs a CTA as cooperative fibers with __syncthreads / warp collectives)
// and run them against the oracle without a GPU; compare_kernels.cu includes it for the product.
// Shared memory is declared through SMB_SHARED / SMB_DYN_SHARED (common.cuh) for that reason.
#pragma once
#include "common.cuh"
#include "db_index.cuh"
#include "join_stripe.cuh"
#include "join_walk.cuh"
#include "range_search.cuh"

namespace smb {

__global__ void __launch_bounds__(256) range_bounds_kernel(const u64* __restrict__ h, const u64* __restrict__ off,
                                                          int n, u64 width, int P, u32* __restrict__ bounds) {
    const u64 total = (u64)n * (u64)(P + 1);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (u64)gridDim.x * blockDim.x) {
        const int p = (int)(i / (u64)n), r = (int)(i - (u64)p * n);
        bounds[i] = (u32)range_bound(h + off[r], off[r + 1] - off[r], width, p, P);
    }
}

struct RangeArgs {
    const u64* q; u64 nq;                 // the query, sorted
    const u32* dir; u32 shift; u64 nbk;   // directory over the query (launch_build_global_dir)
    const u64* hB; const u64* offB; int nB;
    const u32* bounds;                    // [P + 1][nB]
    u64 width; int P; u32 bm_shift, bm_words;
    u32* out;                             // zeroed by the caller; CTAs add their range's matches
};

// one CTA per key range
__global__ void __launch_bounds__(1024, 1) one_vs_many_ranges_kernel(RangeArgs a) {
    SMB_DYN_SHARED(u32, range_bm);
    SMB_SHARED u64 s_q[2];
    const int p = blockIdx.x;
    const u64 lo = (u64)p * a.width;
    if (threadIdx.x == 0) {
        // query keys of this range: lo <= k, k - lo < width (monotone predicate, no overflow)
        const u64 qlo = range_lower_bound(a.q, a.nq, lo);
        u64 l = qlo, hgh = a.nq;
        while (l < hgh) { const u64 mid = (l + hgh) >> 1; if (a.q[mid] - lo < a.width) l = mid + 1; else hgh = mid; }
        s_q[0] = qlo; s_q[1] = l;
    }
    for (u32 i = threadIdx.x; i < a.bm_words; i += blockDim.x) range_bm[i] = 0;
    __syncthreads();
    for (u64 i = s_q[0] + threadIdx.x; i < s_q[1]; i += blockDim.x) {
        const u64 bit = range_bit(a.q[i], lo, a.bm_shift);
        atomicOr(range_bm + (bit >> 5), 1u << (bit & 31));
    }
    __syncthreads();
    if (s_q[0] == s_q[1]) return;                          // no query key in this range: nothing can match
    const u32 lane = lane_id(), warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const u32* __restrict__ b0p = a.bounds + (size_t)p * a.nB;
    const u32* __restrict__ b1p = b0p + a.nB;
    constexpr int U = 4;                                   // row slices in flight per warp
    for (int rbase = (int)warp * 32; rbase < a.nB; rbase += (int)n_warps * 32) {
        const int r = rbase + (int)lane;
        u64 my_start = 0;
        u32 my_len = 0;
        if (r < a.nB) { const u32 b0 = b0p[r]; my_len = b1p[r] - b0; my_start = a.offB[r] + b0; }
        const int cnt = min(32, a.nB - rbase);
        for (int j = 0; j < cnt; j += U) {
            u64 x[U][2];
            u32 len[U];
            u64 start[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {                  // U x 2 independent loads in flight
                const int jj = min(j + u, cnt - 1);
                start[u] = __shfl_sync(0xffffffffu, my_start, jj);
                len[u] = (j + u < cnt) ? __shfl_sync(0xffffffffu, my_len, jj) : 0u;
                x[u][0] = lane < len[u] ? ld_nc_u64(a.hB + start[u] + lane) : 0;
                x[u][1] = lane + 32 < len[u] ? ld_nc_u64(a.hB + start[u] + lane + 32) : 0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (len[u] == 0) continue;                 // uniform in the warp
                u32 c = 0;
                for (u32 base = 0; base < len[u]; base += 64) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const u32 e = base + 32u * half + lane;
                        if (e >= len[u]) continue;
                        const u64 xv = base == 0 ? x[u][half] : ld_nc_u64(a.hB + start[u] + e);
                        const u64 bit = range_bit(xv, lo, a.bm_shift);
                        if (!((range_bm[bit >> 5] >> (bit & 31)) & 1u)) continue;
                        const u64 b = xv >> a.shift;       // bitmap hit: locate the key through the directory
                        if (b >= a.nbk) continue;
                        u64 pp = a.dir[b];
                        const u64 pe = a.dir[b + 1];
                        for (; pp < pe; ++pp) {
                            const u64 k = ld_nc_u64(a.q + pp);
                            if (k >= xv) { c += (k == xv); break; }
                        }
                    }
                }
                c = __reduce_add_sync(0xffffffffu, c);
                if (lane == 0 && c) atomicAdd(a.out + rbase + j + u, c);
            }
        }
    }
}

__global__ void __launch_bounds__(256) stripe_iota_kernel(u32* __restrict__ v, u64 T) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < T; i += (u64)gridDim.x * blockDim.x) v[i] = (u32)i;
}

// sorted (key, CSR element) stream -> tags (row | head flag) and the inverse permutation
__global__ void __launch_bounds__(256) stripe_tag_kernel(const u64* __restrict__ sorted_keys, const u32* __restrict__ src,
                                                        const u64* __restrict__ off, int n, u64 T,
                                                        u32* __restrict__ tags, u32* __restrict__ pos) {
    for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < T; q += (u64)gridDim.x * blockDim.x) {
        const u32 e = src[q];
        tags[q] = stripe_make_tag(sorted_keys, q, stripe_row_of(off, n, e));
        pos[e] = (u32)q;
    }
}

__global__ void __launch_bounds__(256) stripe_low32_kernel(const u64* __restrict__ h, u64 T, u32* __restrict__ low,
                                                          u32* __restrict__ vals) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < T; i += (u64)gridDim.x * blockDim.x) {
        low[i] = (u32)h[i];
        vals[i] = (u32)i;
    }
}
__global__ void __launch_bounds__(256) stripe_gather_keys_kernel(const u64* __restrict__ h, const u32* __restrict__ src,
                                                                u64 T, u64* __restrict__ keys) {
    for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < T; q += (u64)gridDim.x * blockDim.x) keys[q] = h[src[q]];
}
__global__ void __launch_bounds__(256) stripe_mixed_runs_kernel(const u32* __restrict__ low_sorted, const u64* __restrict__ keys,
                                                               u64 T, u8* __restrict__ flags) {
    for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < T; q += (u64)gridDim.x * blockDim.x) {
        bool mixed;
        const u64 m = stripe_run_at_head(low_sorted, keys, T, q, mixed);
        if (mixed) for (u64 j = 0; j < m; ++j) flags[q + j] = 1;
    }
}
__global__ void __launch_bounds__(256) stripe_repair_load_kernel(const u32* __restrict__ where, u64 n_sel, const u64* __restrict__ keys,
                                                                const u32* __restrict__ src, u64* __restrict__ rot,
                                                                u32* __restrict__ sel_src) {
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n_sel; j += (u64)gridDim.x * blockDim.x) {
        rot[j] = stripe_rotated_key(keys[where[j]]);
        sel_src[j] = src[where[j]];
    }
}
__global__ void __launch_bounds__(256) stripe_repair_store_kernel(const u32* __restrict__ where, u64 n_sel, const u64* __restrict__ rot_sorted,
                                                                 const u32* __restrict__ src_sorted, u64* __restrict__ keys,
                                                                 u32* __restrict__ src) {
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n_sel; j += (u64)gridDim.x * blockDim.x) {
        keys[where[j]] = stripe_rotated_key(rot_sorted[j]);          // rotating twice by 32 is the identity
        src[where[j]] = src_sorted[j];
    }
}

struct StripeArgs {
    const u32* tags;
    const u32* pos;
    const u64* off;          // CSR offsets of the set: element ranges and sizes of the rows
    u64 T;
    int n, rows_per_block, row_begin, row_end;
    double* out;             // row `row_begin` first, leading dimension n
    int upper_only;          // 1: count and write only cells (i, j >= i); stripe_mirror_kernel fills the rest
};

// One CTA = rows [r0, r1) of the result.  Warps take 32 consecutive elements of the block at a time;
// the group scans of two elements are in flight together (four independent tag loads per lane).
__global__ void __launch_bounds__(1024, 1) join_stripe_kernel(StripeArgs a) {
    SMB_DYN_SHARED(unsigned char, stripe_smem);
    u64* s_off = reinterpret_cast<u64*>(stripe_smem);                       // [rows + 1]
    u32* stripe = reinterpret_cast<u32*>(stripe_smem + 40 * sizeof(u64));   // [rows][n]
    const int r0 = a.row_begin + (int)blockIdx.x * a.rows_per_block;
    const int r1 = min(a.row_end, r0 + a.rows_per_block);
    const int rows = r1 - r0;
    const u32 n = (u32)a.n;
    for (int r = threadIdx.x; r <= rows; r += blockDim.x) s_off[r] = a.off[r0 + r];
    for (u32 i = threadIdx.x; i < (u32)rows * n; i += blockDim.x) stripe[i] = 0;
    __syncthreads();
    const u64 e_begin = s_off[0], e_end = s_off[rows];
    const u32 lane = lane_id(), warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const u32* __restrict__ tags = a.tags;

    // the first chunk in both directions of one element, then the rare continuations (groups > 32)
    auto finish = [&](u64 q, u32* row, u32 self, u32 tf, bool sf, u32 tb, bool sb, bool vb) {
        u32 m = __ballot_sync(0xffffffffu, sf);
        if (stripe_fwd_active(m, lane)) atomicAdd(row + (tf & ~STRIPE_HEAD), 1u);
        for (u32 it = 1; stripe_continue(m); ++it) {
            const bool st = stripe_fwd_stop(tags, a.T, q, it, lane, tf);
            m = __ballot_sync(0xffffffffu, st);
            if (stripe_fwd_active(m, lane)) atomicAdd(row + (tf & ~STRIPE_HEAD), 1u);
        }
        if (a.upper_only || (self & STRIPE_HEAD)) return; // (rows ascend inside a group: the elements in front of q
                                                          //  are the columns j < i) / q opens its group
        m = __ballot_sync(0xffffffffu, sb);
        if (stripe_bwd_active(m, lane, vb)) atomicAdd(row + (tb & ~STRIPE_HEAD), 1u);
        for (u32 it = 1; stripe_continue(m); ++it) {
            const bool st = stripe_bwd_stop(tags, q, it, lane, tb, vb);
            m = __ballot_sync(0xffffffffu, st);
            if (stripe_bwd_active(m, lane, vb)) atomicAdd(row + (tb & ~STRIPE_HEAD), 1u);
        }
    };

    for (u64 base = e_begin + (u64)warp * 32; base < e_end; base += (u64)n_warps * 32) {
        const u64 e = base + lane;
        const bool have = e < e_end;
        const u32 my_q = have ? a.pos[e] : 0u;
        const u32 my_row = have ? stripe_local_row(s_off, rows, e) : 0u;
        const u32 cnt = (u32)min((u64)32, e_end - base);
        for (u32 j = 0; j < cnt; j += 2) {
            const bool two = j + 1 < cnt;
            const u64 q0 = __shfl_sync(0xffffffffu, my_q, j);
            const u64 q1 = __shfl_sync(0xffffffffu, my_q, two ? j + 1 : j);
            u32* row0 = stripe + (size_t)__shfl_sync(0xffffffffu, my_row, j) * n;
            u32* row1 = stripe + (size_t)__shfl_sync(0xffffffffu, my_row, two ? j + 1 : j) * n;
            u32 tf0, tb0, tf1 = 0, tb1 = 0;
            bool vb0, vb1 = false, sf1 = true, sb1 = true;
            const u32 self0 = tags[q0];
            const bool sf0 = stripe_fwd_stop(tags, a.T, q0, 0, lane, tf0);
            const bool sb0 = stripe_bwd_stop(tags, q0, 0, lane, tb0, vb0);
            u32 self1 = STRIPE_HEAD;
            if (two) {
                self1 = tags[q1];
                sf1 = stripe_fwd_stop(tags, a.T, q1, 0, lane, tf1);
                sb1 = stripe_bwd_stop(tags, q1, 0, lane, tb1, vb1);
            }
            finish(q0, row0, self0, tf0, sf0, tb0, sb0, vb0);
            if (two) finish(q1, row1, self1, tf1, sf1, tb1, sb1, vb1);
        }
    }
    __syncthreads();
    // counts -> float64 rows, written once
    for (u32 i = threadIdx.x; i < (u32)rows * n; i += blockDim.x) {
        const u32 al = i / n, j = i - al * n;
        const int row = r0 + (int)al;
        if (a.upper_only && j < (u32)row) continue;
        const double v = stripe_jaccard(stripe[i], s_off[al + 1] - s_off[al], a.off[j + 1] - a.off[j], (u32)row == j);
        a.out[(size_t)(row - a.row_begin) * n + j] = v;
    }
}

// out[i][j] = out[j][i] for i in [row_begin, row_end), j < i: 32 x 32 tiles through shared memory, reads
// and writes both coalesced.  `full` points at row 0 of the whole matrix (rows < row_end are complete
// in their upper part).
__global__ void __launch_bounds__(1024) stripe_mirror_kernel(double* __restrict__ full, int n, int row_begin, int row_end) {
    SMB_SHARED double tile[32][33];
    const int ti = row_begin / 32 + (int)blockIdx.y;       // tile row (destination rows)
    const int tj = (int)blockIdx.x;                        // tile column (destination columns), tj <= ti
    if (tj > ti) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    // source tile = rows of tile column tj, columns of tile row ti (the upper part)
    const int sr = tj * 32 + ty, sc = ti * 32 + tx;
    tile[ty][tx] = (sr < n && sc < n) ? full[(size_t)sr * n + sc] : 0.0;
    __syncthreads();
    const int dr = ti * 32 + ty, dc = tj * 32 + tx;
    if (dr >= row_begin && dr < row_end && dc < dr && dc < n) full[(size_t)dr * n + dc] = tile[tx][ty];
}

__global__ void __launch_bounds__(256) index_rowid_kernel(const u64* __restrict__ off, int n_rows, u32* __restrict__ ids) {
    for (int r = blockIdx.x; r < n_rows; r += gridDim.x)
        for (u64 i = off[r] + threadIdx.x; i < off[r + 1]; i += blockDim.x) ids[i] = (u32)r;
}

// counts[row] += 1 for every (query hash, row) pair the index holds.  One lane per query hash; groups
// of up to 32 rows are walked by their lane, longer ones by the whole warp.
__global__ void __launch_bounds__(256) index_count_kernel(DbIndexView ix, const u64* __restrict__ q, u64 nq,
                                                         const u32* __restrict__ d_nq, u32* __restrict__ counts) {
    if (d_nq) nq = *d_nq;                                  // length produced on the device by an earlier kernel
    const u32 lane = lane_id();
    const u64 warp0 = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 n_warps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 base = warp0 * 32; base < nq; base += n_warps * 32) {
        const u64 i = base + lane;
        u32 b = 0, e = 0;
        if (i < nq) {
            const long long u = db_index_find(ix, q[i]);
            if (u >= 0) { b = ix.start[u]; e = ix.start[u + 1]; }
        }
        const bool wide = e - b > 32;
        if (!wide) for (u32 j = b; j < e; ++j) atomicAdd(counts + ix.rows[j], 1u);
        u32 todo = __ballot_sync(0xffffffffu, wide);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const u32 gb = __shfl_sync(0xffffffffu, b, src), ge = __shfl_sync(0xffffffffu, e, src);
            for (u32 j = gb + lane; j < ge; j += 32) atomicAdd(counts + ix.rows[j], 1u);
        }
    }
}

// ---- inverted join: default kernels, cluster layout, row-block passes (moved from compare_kernels.cu) ----
__global__ void __launch_bounds__(256) join_row_range_kernel(const u64* __restrict__ h, const u64* __restrict__ off,
                                                            int n_rows, u64 key_lo, u64 key_hi, int bounded_hi,
                                                            u64* __restrict__ beg, u64* __restrict__ cnt) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    if (r == n_rows) { cnt[r] = 0; return; }              // extra slot: the scan then yields the total
    const u64* row = h + off[r];
    const u64 len = off[r + 1] - off[r];
    u64 lo = 0, hi = len;
    while (lo < hi) { u64 mid = (lo + hi) >> 1; if (row[mid] < key_lo) lo = mid + 1; else hi = mid; }
    const u64 b = lo;
    u64 e = len;
    if (bounded_hi) {
        lo = b; hi = len;
        while (lo < hi) { u64 mid = (lo + hi) >> 1; if (row[mid] < key_hi) lo = mid + 1; else hi = mid; }
        e = lo;
    }
    beg[r] = b;
    cnt[r] = e - b;
}

__global__ void __launch_bounds__(256) join_gather_kernel(const u64* __restrict__ h, const u64* __restrict__ off,
                                                         const u64* __restrict__ beg, const u64* __restrict__ dst_off,
                                                         int n_rows, u64* __restrict__ keys, u32* __restrict__ ids) {
    for (int r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const u64 src = off[r] + beg[r], d0 = dst_off[r], n = dst_off[r + 1] - d0;
        for (u64 i = threadIdx.x; i < n; i += blockDim.x) { keys[d0 + i] = h[src + i]; ids[d0 + i] = (u32)r; }
    }
}

// out[0] += sum over groups of C(m,2); out[1] = max m
__global__ void __launch_bounds__(256) join_estimate_kernel(const u64* __restrict__ keys, u64 T,
                                                           unsigned long long* __restrict__ out) {
    unsigned long long pairs = 0, mmax = 0;
    for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p < T; p += (u64)gridDim.x * blockDim.x) {
        const u64 m = join_group_size_at_head(keys, T, p);
        pairs += m * (m - (m ? 1 : 0)) / 2;
        mmax = m > mmax ? m : mmax;
    }
    for (int d = 16; d; d >>= 1) {
        pairs += __shfl_xor_sync(0xffffffffu, pairs, d);
        unsigned long long o = __shfl_xor_sync(0xffffffffu, mmax, d);
        mmax = o > mmax ? o : mmax;
    }
    if (lane_id() == 0) { if (pairs) atomicAdd(out, pairs); if (mmax) atomicMax(out + 1, mmax); }
}

// Every element pairs with the later elements of its group: ids ascend inside a group, so
// (ids[p], ids[b]) is an upper-triangle cell.  Increments are fire-and-forget reductions (RED)
// resolved in L2, and their rate is what bounds the kernel: 1.455e9 reductions in 16.7 ms =
// 8.7e10 /s = 0.31 per clock per SM on the 10 000-sketch matrix, and three rewrites that attack
// everything else left the time unchanged to 0.1 % (profiles/r1q_join.txt, r1s_join.txt):
// evict-first stream loads + evict-last reductions (DRAM write-back 7.1 -> 1.4 GB, same time) and
// walking the groups from a shared-memory tile instead of dependent L2 loads (long-scoreboard
// stalls 105 -> 33 per issue, same time).  Going faster needs fewer reductions (accumulating in
// shared-memory tiles of the matrix), see DESIGN.md section 4.5.
__global__ void __launch_bounds__(256) join_count_kernel(const u64* __restrict__ keys, const u32* __restrict__ ids,
                                                        u64 T, u32* __restrict__ common, size_t ld) {
    const u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= T) return;
    join_walk(keys, ids, T, p, [&](u32 a, u32 b) { atomicAdd(common + (size_t)a * ld + b, 1u); });
}

__global__ void __launch_bounds__(256) join_rowkey_kernel(const u64* __restrict__ keys, const u32* __restrict__ ids,
                                                         u64 T, unsigned long long* __restrict__ rowkey) {
    const u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= T) return;
    if (join_is_shared(keys, T, p)) atomicMin(rowkey + ids[p], (unsigned long long)keys[p]);
}

__global__ void __launch_bounds__(256) join_iota_kernel(u32* __restrict__ v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (u32)i;
}

__global__ void __launch_bounds__(256) join_invert_kernel(const u32* __restrict__ order, int n, u32* __restrict__ inv) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) inv[order[r]] = (u32)r;
}

// per-rank element counts (rank r holds row order[r]); slot n is the scan's total
__global__ void __launch_bounds__(256) join_rank_counts_kernel(const u64* __restrict__ cnt, const u32* __restrict__ order,
                                                              int n, u64* __restrict__ cnt_rank) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    cnt_rank[r] = r < n ? cnt[order[r]] : 0;
}

// like join_gather_kernel, rows visited in rank order and labelled with their rank
__global__ void __launch_bounds__(256) join_gather_ranked_kernel(const u64* __restrict__ h, const u64* __restrict__ off,
                                                                const u64* __restrict__ beg, const u32* __restrict__ order,
                                                                const u64* __restrict__ dst_off, int n_rows,
                                                                u64* __restrict__ keys, u32* __restrict__ ids) {
    for (int r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const u32 row = order[r];
        const u64 src = off[row] + beg[row], d0 = dst_off[r], n = dst_off[r + 1] - d0;
        for (u64 i = threadIdx.x; i < n; i += blockDim.x) { keys[d0 + i] = h[src + i]; ids[d0 + i] = (u32)r; }
    }
}

// one warp per element: lanes take consecutive later elements of the group
__global__ void __launch_bounds__(256) join_count_warp_kernel(const u64* __restrict__ keys, const u32* __restrict__ ids,
                                                             u64 T, u32* __restrict__ common, size_t ld) {
    const u64 p = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= T) return;                                   // whole warps leave together
    const u32 lane = lane_id();
    for (u64 b0 = p + 1;; b0 += 32) {
        const bool hit = join_walk_lane(keys, ids, T, p, b0, lane,
                                        [&](u32 a, u32 b) { atomicAdd(common + (size_t)a * ld + b, 1u); });
        if (!__all_sync(0xffffffffu, hit)) break;
    }
}

// common[i][j] += rank-space count of (i, j), for i < j
__global__ void __launch_bounds__(256) join_unpermute_add_kernel(const u32* __restrict__ ranked, const u32* __restrict__ inv,
                                                                int n, size_t ld_ranked, u32* __restrict__ common, size_t ld) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    for (int i = blockIdx.y; i < j; i += gridDim.y) {       // every (i, j), i < j, is visited exactly once
        u32 lo, hi;
        join_rank_cell(inv, (u32)i, (u32)j, lo, hi);
        const u32 v = ranked[(size_t)lo * ld_ranked + hi];
        if (v) common[(size_t)i * ld + j] += v;
    }
}

__global__ void __launch_bounds__(256) join_count_rows_kernel(const u64* __restrict__ keys, const u32* __restrict__ ids,
                                                             u64 T, u32 r0, u32 r1, u32* __restrict__ common, size_t ld) {
    const u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= T) return;
    join_walk_rows(keys, ids, T, p, r0, r1, [&](u32 a, u32 b) { atomicAdd(common + (size_t)a * ld + b, 1u); });
}

}  // namespace smb
