// see device_radix_sort.cuh
#pragma once
#include <stddef.h>

#include <type_traits>
#include <vector>

#include <cuda_runtime.h>

namespace cub {
struct DeviceScan {
    template <class In, class Out, class N>
    static cudaError_t ExclusiveSum(void* tmp, size_t& bytes, In in, Out out, N n, cudaStream_t = 0) {
        if (!tmp) { bytes = 16; return cudaSuccess; }
        typedef typename std::remove_reference<decltype(out[0])>::type T;
        T acc = 0;
        for (N i = 0; i < n; ++i) { const T v = (T)in[i]; out[i] = acc; acc = (T)(acc + v); }   // in place allowed
        return cudaSuccess;
    }
};
}  // namespace cub
