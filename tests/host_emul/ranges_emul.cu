// ranges_emul.cu -- the range-partitioned one-vs-many pass (sourmash_b200/csrc/range_search.cuh, compiled
// here for the host) driven like compare_kernels.cu's one_vs_many_ranges_kernel: equal key ranges over
// the database's key space, per-row slice bounds, the query bitmap of one range, probes of the slices.
// The key compare behind a bitmap hit (the directory walk of the CUDA kernel, GPU-tested) is a binary
// search here.  Test infrastructure for the CPU-only suite.
//   usage: ranges_emul <P> <max_bitmap_bits> <query.u64> <hashes.u64> <offsets.u64> <out.u32 (n counts)>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "../../sourmash_b200/csrc/range_search.cuh"

using namespace smb;

template <class T>
static std::vector<T> slurp(const char* path) {
    FILE* f = fopen(path, "rb");
    std::vector<T> v;
    if (!f) return v;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    v.resize((size_t)n / sizeof(T));
    if (n && fread(v.data(), 1, (size_t)n, f) != (size_t)n) v.clear();
    fclose(f);
    return v;
}

int main(int argc, char** argv) {
    if (argc != 7) return 2;
    const int P = atoi(argv[1]);
    const u64 max_bits = strtoull(argv[2], nullptr, 10);
    std::vector<u64> q = slurp<u64>(argv[3]), h = slurp<u64>(argv[4]), off = slurp<u64>(argv[5]);
    const int n = (int)off.size() - 1;
    u64 max_key = 0;
    for (u64 v : h) max_key = std::max(max_key, v);
    const u64 width = range_width(max_key, P);
    const u32 bm_shift = range_bitmap_shift(width, max_bits);
    const u64 bits = ((width - 1) >> bm_shift) + 1;
    if (bits > max_bits) return 3;
    // bounds table, [P + 1][n]
    std::vector<u32> bounds((size_t)(P + 1) * n);
    for (int p = 0; p <= P; ++p)
        for (int r = 0; r < n; ++r)
            bounds[(size_t)p * n + r] = (u32)range_bound(h.data() + off[r], off[r + 1] - off[r], width, p, P);
    for (int r = 0; r < n; ++r) {                           // the slices of a row tile it exactly
        if (bounds[r] != 0 || bounds[(size_t)P * n + r] != off[r + 1] - off[r]) return 4;
        for (int p = 0; p < P; ++p) if (bounds[(size_t)p * n + r] > bounds[(size_t)(p + 1) * n + r]) return 4;
    }
    std::vector<u32> out(n, 0);
    std::vector<u32> bm((bits + 31) / 32);
    for (int p = 0; p < P; ++p) {
        const u64 lo = (u64)p * width;
        const u64 qlo = range_lower_bound(q.data(), q.size(), lo);
        u64 l = qlo, hgh = q.size();
        while (l < hgh) { const u64 mid = (l + hgh) >> 1; if (q[mid] - lo < width) l = mid + 1; else hgh = mid; }
        const u64 qhi = l;
        std::fill(bm.begin(), bm.end(), 0u);
        for (u64 i = qlo; i < qhi; ++i) {
            const u64 bit = range_bit(q[i], lo, bm_shift);
            if (bit >= bits) return 5;
            bm[bit >> 5] |= 1u << (bit & 31);
        }
        if (qlo == qhi) continue;
        for (int r = 0; r < n; ++r) {
            const u64 b0 = bounds[(size_t)p * n + r], b1 = bounds[(size_t)(p + 1) * n + r];
            for (u64 e = off[r] + b0; e < off[r] + b1; ++e) {
                const u64 x = h[e];
                if (x < lo) return 6;                       // a slice only holds keys of its range
                const u64 bit = range_bit(x, lo, bm_shift);
                if (bit >= bits) return 6;
                if (!((bm[bit >> 5] >> (bit & 31)) & 1u)) continue;
                out[r] += std::binary_search(q.begin(), q.end(), x) ? 1u : 0u;
            }
        }
    }
    FILE* f = fopen(argv[6], "wb");
    fwrite(out.data(), 4, out.size(), f);
    fclose(f);
    return 0;
}
