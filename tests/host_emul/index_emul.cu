// index_emul.cu -- the inverted index of a resident set (sourmash_b200/csrc/db_index.cuh, compiled here
// for the host): (hash, row) pairs sorted by hash (std::stable_sort stands in for the radix sort),
// distinct keys + offsets (run-length encode + exclusive sum), the bucket directory in the geometry
// db_index_dir_plan chooses, then one lookup per query hash and one increment per (hash, row) match,
// like compare_kernels.cu's index_count_kernel.  Test infrastructure for the CPU-only suite.
//   usage: index_emul <query.u64> <hashes.u64> <offsets.u64> <out.u32 (n counts)>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "../../sourmash_b200/csrc/db_index.cuh"

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
    if (argc != 5) return 2;
    std::vector<u64> q = slurp<u64>(argv[1]), h = slurp<u64>(argv[2]), off = slurp<u64>(argv[3]);
    const int n = (int)off.size() - 1;
    const u64 T = h.size();
    std::vector<u32> ids(T);
    for (int r = 0; r < n; ++r) for (u64 i = off[r]; i < off[r + 1]; ++i) ids[i] = (u32)r;
    std::vector<size_t> perm(T);
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](size_t a, size_t b) { return h[a] < h[b]; });
    std::vector<u64> keys;
    std::vector<u32> start, rows(T);
    for (u64 p = 0; p < T; ++p) {
        rows[p] = ids[perm[p]];
        if (p == 0 || h[perm[p]] != h[perm[p - 1]]) { keys.push_back(h[perm[p]]); start.push_back((u32)p); }
        else if (rows[p] <= rows[p - 1]) return 3;            // rows ascend inside a group
    }
    start.push_back((u32)T);
    u64 max_key = 0;
    for (u64 v : h) max_key = std::max(max_key, v);
    u32 shift;
    u64 nbk;
    db_index_dir_plan(keys.size(), max_key, shift, nbk);
    if ((max_key >> shift) + 1 != nbk || nbk > (1ull << 27) + 1) return 4;
    std::vector<u32> dir(nbk + 1);                            // dir[b] = number of keys with bucket < b
    for (u64 b = 0; b <= nbk; ++b)
        dir[b] = (u32)(std::lower_bound(keys.begin(), keys.end(), b, [&](u64 k, u64 bb) { return (k >> shift) < bb; }) - keys.begin());
    DbIndexView ix{keys.data(), keys.size(), start.data(), rows.data(), dir.data(), shift, nbk};
    for (u64 u = 0; u < keys.size(); ++u) if (db_index_find(ix, keys[u]) != (long long)u) return 5;
    std::vector<u32> out(n, 0);
    for (u64 x : q) {
        const long long u = db_index_find(ix, x);
        if (u < 0) continue;
        for (u32 j = start[u]; j < start[u + 1]; ++j) out[rows[j]] += 1;
    }
    FILE* f = fopen(argv[4], "wb");
    fwrite(out.data(), 4, out.size(), f);
    fclose(f);
    return 0;
}
