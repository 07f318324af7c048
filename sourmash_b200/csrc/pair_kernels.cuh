// pair_kernels.cuh -- the remaining kernels of the intersection path (compare_kernels.cu launches them): the
// largest-key reduction, the generic pair kernel for rows too large for shared memory, the bottom-k ("num")
// kernel, the abundance (angular) kernels and the counts -> float64 finalize kernels.  In a header so that
// tests/host_emul/simt_emul.cu can run the kernels themselves on the CPU (tests/host_emul/simt.h).
#pragma once
#include <math.h>

#include "common.cuh"
#include "search_kernels.cuh"

namespace smb {

__global__ void max_last_kernel(const u64* __restrict__ hA, const u64* __restrict__ offA, int nA,
                                const u64* __restrict__ hB, const u64* __restrict__ offB, int nB,
                                unsigned long long* __restrict__ d_max) {
    u64 m = 0;
    int total = nA + nB;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < total; r += gridDim.x * blockDim.x) {
        const u64* h = r < nA ? hA : hB;
        const u64* off = r < nA ? offA : offB;
        int i = r < nA ? r : r - nA;
        u64 beg = off[i], end = off[i + 1];
        if (end > beg) { u64 v = h[end - 1]; m = v > m ? v : m; }
    }
    for (int d = 16; d; d >>= 1) { u64 o = __shfl_xor_sync(0xffffffffu, m, d); m = o > m ? o : m; }
    if (lane_id() == 0 && m) atomicMax(d_max, (unsigned long long)m);
}

__global__ void __launch_bounds__(256) pairwise_generic_kernel(
    const u64* __restrict__ hA, const u64* __restrict__ offA, int nA, const u64* __restrict__ hB,
    const u64* __restrict__ offB, int nB, u32* __restrict__ out, size_t ldo, int symmetric,
    int shard, int n_shards) {
    // shard s of n_shards takes the rows i of A with i % n_shards == s (the partial matrices of the
    // shards are summed by the caller, so every pair must be counted by exactly one shard)
    const u64 npairs = (u64)nA * (u64)nB;
    const u64 wstride = (u64)gridDim.x * (blockDim.x >> 5);
    const int lane = lane_id();
    for (u64 w = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < npairs; w += wstride) {
        int i = (int)(w / (u64)nB), j = (int)(w % (u64)nB);
        if (symmetric && j <= i) continue;
        if (n_shards > 1 && i % n_shards != shard) continue;
        const u64* ra = hA + offA[i]; u64 na = offA[i + 1] - offA[i];
        const u64* rb = hB + offB[j]; u64 nb = offB[j + 1] - offB[j];
        if (na > nb) { const u64* tr = ra; ra = rb; rb = tr; u64 tn = na; na = nb; nb = tn; }
        u32 c = 0;
        for (u64 e = lane; e < na; e += 32) c += row_contains(rb, nb, ld_nc_u64(ra + e)) ? 1u : 0u;
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) out[(size_t)i * ldo + j] = c;
    }
}

__global__ void __launch_bounds__(256) pairwise_num_kernel(
    const u64* __restrict__ hA, const u64* __restrict__ offA, int nA, const u64* __restrict__ hB,
    const u64* __restrict__ offB, int nB, u32 num, u32* __restrict__ common,
    u32* __restrict__ usize, size_t ldo, int symmetric) {
    const u64 npairs = (u64)nA * (u64)nB;
    const u64 wstride = (u64)gridDim.x * (blockDim.x >> 5);
    const int lane = lane_id();
    for (u64 w = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < npairs; w += wstride) {
        int i = (int)(w / (u64)nB), j = (int)(w % (u64)nB);
        if (symmetric && j <= i) continue;
        const u64* ra = hA + offA[i]; u64 na = offA[i + 1] - offA[i];
        const u64* rb = hB + offB[j]; u64 nb = offB[j + 1] - offB[j];
        u64 matches_before = 0;     // matches among a_0 .. a_{base-1}
        u32 c_in_m = 0, c_total = 0;
        for (u64 base = 0; base < na; base += 32) {
            u64 e = base + lane;
            bool valid = e < na;
            u64 x = valid ? ld_nc_u64(ra + e) : 0;
            u64 lo = 0, hi = valid ? nb : 0;
            while (lo < hi) {
                u64 mid = (lo + hi) >> 1;
                if (ld_nc_u64(rb + mid) < x) lo = mid + 1; else hi = mid;
            }
            bool m = valid && lo < nb && ld_nc_u64(rb + lo) == x;
            u32 bal = __ballot_sync(0xffffffffu, m);
            u64 prior = matches_before + __popc(bal & ((1u << lane) - 1u));
            u64 rank = e + lo - prior;
            if (m) { ++c_total; if (num == 0 || rank < (u64)num) ++c_in_m; }
            matches_before += __popc(bal);
        }
        c_in_m = __reduce_add_sync(0xffffffffu, c_in_m);
        c_total = __reduce_add_sync(0xffffffffu, c_total);
        if (lane == 0) {
            u64 un = na + nb - c_total;
            if (num != 0 && un > num) un = num;
            common[(size_t)i * ldo + j] = c_in_m;
            if (usize) usize[(size_t)i * ldo + j] = (u32)un;
        }
    }
}

__global__ void __launch_bounds__(256) row_sumsq_kernel(const u64* __restrict__ ab, const u64* __restrict__ off,
                                                       int n, unsigned long long* __restrict__ out) {
    const int lane = lane_id();
    const int wstride = gridDim.x * (blockDim.x >> 5);
    for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += wstride) {
        unsigned long long acc = 0;
        for (u64 e = off[r] + lane; e < off[r + 1]; e += 32) { u64 v = ab[e]; acc += v * v; }
        for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if (lane == 0) out[r] = acc;
    }
}

__global__ void __launch_bounds__(256) pairwise_angular_kernel(
    const u64* __restrict__ h, const u64* __restrict__ ab, const u64* __restrict__ off, int n,
    const unsigned long long* __restrict__ sumsq, double* __restrict__ out) {
    const u64 npairs = (u64)n * (u64)n;
    const u64 wstride = (u64)gridDim.x * (blockDim.x >> 5);
    const int lane = lane_id();
    for (u64 w = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < npairs; w += wstride) {
        int i = (int)(w / (u64)n), j = (int)(w % (u64)n);
        if (j < i) continue;
        if (i == j) { if (lane == 0) out[(size_t)i * n + i] = 1.0; continue; }
        const u64 ao = off[i], na = off[i + 1] - ao, bo = off[j], nb = off[j + 1] - bo;
        unsigned long long dot = 0;
        for (u64 e = lane; e < na; e += 32) {
            u64 x = ld_nc_u64(h + ao + e);
            u64 lo = 0, hi = nb;
            while (lo < hi) { u64 mid = (lo + hi) >> 1; if (ld_nc_u64(h + bo + mid) < x) lo = mid + 1; else hi = mid; }
            if (lo < nb && ld_nc_u64(h + bo + lo) == x) dot += ab[ao + e] * ab[bo + lo];
        }
        for (int d = 16; d; d >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, d);
        if (lane == 0) {
            double na_ = sqrt((double)sumsq[i]), nb_ = sqrt((double)sumsq[j]);
            double v = 0.0;
            if (na_ != 0.0 && nb_ != 0.0) {
                double p = fmin((double)dot / (na_ * nb_), 1.0);
                v = 1.0 - 2.0 * acos(p) / 3.14159265358979323846264338327950288;
            }
            out[(size_t)i * n + j] = v;
            out[(size_t)j * n + i] = v;
        }
    }
}

__global__ void __launch_bounds__(256) finalize_matrix_kernel(
    const u32* __restrict__ common, const u32* __restrict__ usize, size_t ldo,
    const u64* __restrict__ offA, const u64* __restrict__ offB, int nA, int nB, int mode,
    int symmetric, double* __restrict__ out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y;
    if (j >= nB || i >= nA) return;
    double v;
    if (symmetric && i == j) {
        v = 1.0;                                  // compare.py:38 np.ones diagonal
    } else {
        int r = i, c = j;
        if (symmetric && j < i) { r = j; c = i; }
        u64 cm = common[(size_t)r * ldo + c];
        if (mode == 0) {
            u64 na = offA[r + 1] - offA[r], nbb = offB[c + 1] - offB[c];
            u64 un = na + nbb - cm;
            v = (double)cm / (double)(un > 1 ? un : 1);
        } else if (mode == 1) {
            u64 un = usize[(size_t)r * ldo + c];
            v = (double)cm / (double)(un > 1 ? un : 1);
        } else {
            v = (double)cm;
        }
    }
    out[(size_t)i * nB + j] = v;
}

__global__ void __launch_bounds__(256) finalize_rows_kernel(const u32* __restrict__ common, size_t n,
                                                           const u64* __restrict__ off, int row_begin,
                                                           int row_end, double* __restrict__ out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = row_begin + blockIdx.y;
    if (j >= (int)n || i >= row_end) return;
    double v = 1.0;
    if (i != j) {
        int r = i < j ? i : j, c = i < j ? j : i;
        u64 cm = common[(size_t)r * n + c];
        u64 un = (off[r + 1] - off[r]) + (off[c + 1] - off[c]) - cm;
        v = (double)cm / (double)(un > 1 ? un : 1);
    }
    out[(size_t)(i - row_begin) * n + j] = v;
}

}  // namespace smb
