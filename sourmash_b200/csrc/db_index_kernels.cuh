// db_index_kernels.cuh -- kernels of the inverted index over a resident set (layout and lookup: db_index.cuh);
// compare_kernels.cu launches them, tests/host_emul/simt_emul.cu runs them on the CPU against the oracle.
#pragma once
#include "common.cuh"
#include "db_index.cuh"

namespace smb {

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

}  // namespace smb
