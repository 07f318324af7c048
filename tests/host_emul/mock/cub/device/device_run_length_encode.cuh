// see device_radix_sort.cuh
#pragma once
#include <stddef.h>

#include <cuda_runtime.h>

namespace cub {
struct DeviceRunLengthEncode {
    template <class In, class U, class L, class R>
    static cudaError_t Encode(void* tmp, size_t& bytes, In in, U unique_out, L counts_out, R num_runs_out, int n, cudaStream_t = 0) {
        if (!tmp) { bytes = 16; return cudaSuccess; }
        int runs = 0;
        for (int i = 0; i < n;) {
            int j = i + 1;
            while (j < n && in[j] == in[i]) ++j;
            unique_out[runs] = in[i];
            counts_out[runs] = j - i;
            ++runs;
            i = j;
        }
        *num_runs_out = runs;
        return cudaSuccess;
    }
};
}  // namespace cub
