// join_emul.cu -- runs the per-element logic of the inverted join (sourmash_b200/csrc/join_walk.cuh,
// compiled here for the host) over a CSR sketch set: slice rows to the shard's key range, sort the
// (hash, row) pairs stably by hash (std::stable_sort stands in for the radix sort), walk every
// element.  Test infrastructure for the CPU-only suite.
//   usage: join_emul <n_shards> <hashes.u64> <offsets.u64> <out.u32 (n*n, summed over shards)> <out_pairs.u64>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "../../sourmash_b200/csrc/join_walk.cuh"

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
    if (argc != 6) return 2;
    const int n_shards = atoi(argv[1]);
    std::vector<u64> h = slurp<u64>(argv[2]), off = slurp<u64>(argv[3]);
    const size_t n = off.size() - 1;
    u64 max_key = 0;
    for (u64 v : h) max_key = std::max(max_key, v);
    std::vector<u32> common(n * n, 0);
    u64 total_pairs = 0;
    for (int shard = 0; shard < n_shards; ++shard) {
        u64 lo, hi;
        bool bounded;
        join_shard_range(max_key, shard, n_shards, lo, hi, bounded);
        std::vector<u64> keys;
        std::vector<u32> ids;
        for (size_t r = 0; r < n; ++r)
            for (u64 i = off[r]; i < off[r + 1]; ++i)
                if (h[i] >= lo && (!bounded || h[i] < hi)) { keys.push_back(h[i]); ids.push_back((u32)r); }
        std::vector<size_t> order(keys.size());
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return keys[a] < keys[b]; });
        std::vector<u64> sk(keys.size());
        std::vector<u32> si(keys.size());
        for (size_t i = 0; i < order.size(); ++i) { sk[i] = keys[order[i]]; si[i] = ids[order[i]]; }
        const u64 T = sk.size();
        for (u64 p = 0; p < T; ++p) {
            join_walk(sk.data(), si.data(), T, p, [&](u32 a, u32 b) { common[(size_t)a * n + b] += 1; });
            const u64 m = join_group_size_at_head(sk.data(), T, p);
            total_pairs += m * (m - (m ? 1 : 0)) / 2;
        }
    }
    FILE* f = fopen(argv[4], "wb");
    fwrite(common.data(), 4, common.size(), f);
    fclose(f);
    f = fopen(argv[5], "wb");
    fwrite(&total_pairs, 8, 1, f);
    fclose(f);
    return 0;
}
