// db_index.cuh -- inverted index over a resident sketch set (hash -> rows that contain it) and its
// lookup, shared with tests/host_emul/index_emul.cu so that the CPU-only suite checks the logic
// against the oracle.
//
// search / prefetch / gather count |query ∩ row| for every row of a database.  The streaming passes
// (one_vs_many_*_kernel) read the whole database for that: 12 GB for 300 000 sketches, per query and,
// in gather, per round.  For a database that stays in HBM and is queried repeatedly the sorted
// (hash, row) stream of the set -- the one the inverted join of compare sorts -- can be kept instead:
//   keys[u]                   the distinct hashes, ascending
//   rows[start[u] .. start[u + 1])   the rows holding keys[u], ascending
//   dir[b]                    number of keys with (key >> shift) < b   (launch_build_global_dir)
// A query then costs one directory probe per *query* hash plus one increment per (hash, row) match:
// work proportional to the query and its matches, not to the database.  This is what the reference's
// RevIndex does on the CPU (src/core/src/index/revindex/); here it is the device-side structure
// behind the same counts, built on request (smb_sketchset_build_index).
#pragma once
#include "common.cuh"

namespace smb {

struct DbIndexView {
    const u64* keys;      // [n_keys]
    u64 n_keys;
    const u32* start;     // [n_keys + 1]
    const u32* rows;      // [start[n_keys]]
    const u32* dir;       // [nbk + 1]
    u32 shift;
    u64 nbk;
};

// position of x in keys, or -1
__host__ __device__ __forceinline__ long long db_index_find(const DbIndexView& ix, u64 x) {
    const u64 b = x >> ix.shift;
    if (b >= ix.nbk) return -1;                           // beyond the largest key
    u64 p = ix.dir[b];
    const u64 pe = ix.dir[b + 1];
    for (; p < pe; ++p) {
        const u64 k = ix.keys[p];
        if (k >= x) return k == x ? (long long)p : -1;
    }
    return -1;
}

// directory geometry for n_keys keys up to max_key: about two buckets per key, at most 2^27 buckets
__host__ __device__ __forceinline__ void db_index_dir_plan(u64 n_keys, u64 max_key, u32& shift, u64& nbk) {
    int nb_log2 = 8;
    while (nb_log2 < 27 && (1ull << nb_log2) < 2 * n_keys) ++nb_log2;
    u32 s = 0;
    while (s < 63 && (max_key >> s) >= (1ull << nb_log2)) ++s;
    shift = s;
    nbk = (max_key >> s) + 1;
}

}  // namespace smb
