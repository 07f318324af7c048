// range_kernels.cuh -- kernels of the range-partitioned one-vs-many pass (helpers: range_search.cuh);
// compare_kernels.cu launches them, tests/host_emul/simt_emul.cu runs them on the CPU against the oracle.
#pragma once
#include "common.cuh"
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

}  // namespace smb
