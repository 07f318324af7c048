// Host stand-ins for the cub device algorithms the library calls, for the emulated build of the CPU-only suite
// (tests/host_emul/emul_lib.py).  Same two-phase temp-storage protocol, same results (stable sorts on the key
// bits [begin_bit, end_bit)); the "device" pointers are host pointers there.  Test infrastructure only.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include <cuda_runtime.h>

namespace cub {

struct DeviceRadixSort {
    template <class K>
    static uint64_t bits_of(K k, int begin_bit, int end_bit) {
        const uint64_t v = (uint64_t)k >> begin_bit;
        const int w = end_bit - begin_bit;
        return w >= 64 ? v : (v & ((1ull << w) - 1ull));
    }
    template <class K, class V, class N>
    static cudaError_t SortPairs(void* tmp, size_t& bytes, const K* kin, K* kout, const V* vin, V* vout, N n, int begin_bit = 0,
                                 int end_bit = (int)sizeof(K) * 8, cudaStream_t = 0) {
        if (!tmp) { bytes = 16; return cudaSuccess; }
        std::vector<size_t> p((size_t)n);
        std::iota(p.begin(), p.end(), 0);
        std::stable_sort(p.begin(), p.end(), [&](size_t a, size_t b) { return bits_of(kin[a], begin_bit, end_bit) < bits_of(kin[b], begin_bit, end_bit); });
        std::vector<K> k2((size_t)n);
        std::vector<V> v2((size_t)n);
        for (size_t i = 0; i < (size_t)n; ++i) { k2[i] = kin[p[i]]; v2[i] = vin[p[i]]; }
        std::copy(k2.begin(), k2.end(), kout);
        std::copy(v2.begin(), v2.end(), vout);
        return cudaSuccess;
    }
    template <class K, class N>
    static cudaError_t SortKeys(void* tmp, size_t& bytes, const K* kin, K* kout, N n, int begin_bit = 0,
                                int end_bit = (int)sizeof(K) * 8, cudaStream_t = 0) {
        if (!tmp) { bytes = 16; return cudaSuccess; }
        std::vector<K> k2(kin, kin + (size_t)n);
        std::stable_sort(k2.begin(), k2.end(), [&](K a, K b) { return bits_of(a, begin_bit, end_bit) < bits_of(b, begin_bit, end_bit); });
        std::copy(k2.begin(), k2.end(), kout);
        return cudaSuccess;
    }
};

}  // namespace cub
