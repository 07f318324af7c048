// join_emul.cu -- runs the per-element logic of the inverted join (sourmash_b200/csrc/join_walk.cuh,
// compiled here for the host) over a CSR sketch set: slice rows to the shard's key range, sort the
// (hash, row) pairs stably by hash (std::stable_sort stands in for the radix sort), walk every
// element.  Test infrastructure for the CPU-only suite.
//   usage: join_emul <n_shards> <hashes.u64> <offsets.u64> <out.u32 (n*n, summed over shards)> <out_pairs.u64> [cluster]
//          join_emul <R> <hashes.u64> <offsets.u64> <out.f64 (n*n jaccard)> <low32 | -> stripe <n_warps> <mirror_chunk_rows | 0>
// With "stripe" the experimental stripe layout (join_stripe.cuh) is emulated: CTAs of R rows, warps
// of 32 lanes with a host-side ballot, counters in a per-CTA stripe, float64 rows written directly.
// With "cluster" the experimental layout is emulated instead: row keys from the global sample, rows
// ranked by (key, id), gather in rank order, one 32-lane "warp" per element, un-permute.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "../../sourmash_b200/csrc/join_walk.cuh"
#include "../../sourmash_b200/csrc/join_stripe.cuh"

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

// the sorted (hash, id) stream of the rows restricted to [lo, hi) (bounded) / [lo, inf), rows visited
// in the order given and labelled with their position in it
static void sorted_stream(const std::vector<u64>& h, const std::vector<u64>& off, const std::vector<u32>& order,
                          u64 lo, u64 hi, bool bounded, std::vector<u64>& sk, std::vector<u32>& si) {
    std::vector<u64> keys;
    std::vector<u32> ids;
    for (size_t r = 0; r < order.size(); ++r)
        for (u64 i = off[order[r]]; i < off[order[r] + 1]; ++i)
            if (h[i] >= lo && (!bounded || h[i] < hi)) { keys.push_back(h[i]); ids.push_back((u32)r); }
    std::vector<size_t> perm(keys.size());
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](size_t a, size_t b) { return keys[a] < keys[b]; });
    sk.resize(keys.size()); si.resize(keys.size());
    for (size_t i = 0; i < perm.size(); ++i) { sk[i] = keys[perm[i]]; si[i] = ids[perm[i]]; }
}

static int cluster_main(int n_shards, const std::vector<u64>& h, const std::vector<u64>& off, const char* out_path) {
    const size_t n = off.size() - 1;
    u64 max_key = 0;
    for (u64 v : h) max_key = std::max(max_key, v);
    std::vector<u32> ident(n);
    std::iota(ident.begin(), ident.end(), 0);
    // row keys from the sample
    std::vector<u64> sk;
    std::vector<u32> si;
    sorted_stream(h, off, ident, 0, max_key / 64 + 1, true, sk, si);
    std::vector<u64> rowkey(n, ~0ull);
    for (u64 p = 0; p < sk.size(); ++p)
        if (join_is_shared(sk.data(), sk.size(), p)) rowkey[si[p]] = std::min(rowkey[si[p]], sk[p]);
    std::vector<u32> order(ident);
    std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return rowkey[a] < rowkey[b]; });
    std::vector<u32> inv(n);
    for (size_t r = 0; r < n; ++r) inv[order[r]] = (u32)r;
    std::vector<u32> common(n * n, 0);
    for (int shard = 0; shard < n_shards; ++shard) {
        u64 lo, hi;
        bool bounded;
        join_shard_range(max_key, shard, n_shards, lo, hi, bounded);
        sorted_stream(h, off, order, lo, hi, bounded, sk, si);
        const u64 T = sk.size();
        std::vector<u32> ranked(n * n, 0);
        for (u64 p = 0; p < T; ++p) {                       // one "warp" per element
            for (u64 b0 = p + 1;; b0 += 32) {
                bool all = true;
                for (u32 lane = 0; lane < 32; ++lane)
                    all &= join_walk_lane(sk.data(), si.data(), T, p, b0, lane,
                                          [&](u32 a, u32 b) { ranked[(size_t)a * n + b] += 1; });
                if (!all) break;
            }
        }
        for (size_t j = 0; j < n; ++j)
            for (size_t i = 0; i < j; ++i) {
                u32 rl, rh;
                join_rank_cell(inv.data(), (u32)i, (u32)j, rl, rh);
                common[i * n + j] += ranked[(size_t)rl * n + rh];
            }
        for (size_t a = 0; a < n; ++a)                      // nothing may land outside the upper triangle
            for (size_t b = 0; b <= a; ++b)
                if (ranked[a * n + b]) return 5;
    }
    FILE* f = fopen(out_path, "wb");
    fwrite(common.data(), 4, common.size(), f);
    fclose(f);
    return 0;
}

// row-block passes: pass k walks the elements whose row is in block k, then the *full* rows of block k
// (both triangles, read as c[min][max]) are captured; nothing captured is ever revisited
static int rows_main(int passes, const std::vector<u64>& h, const std::vector<u64>& off, const char* out_path) {
    const size_t n = off.size() - 1;
    std::vector<u32> ident(n);
    std::iota(ident.begin(), ident.end(), 0);
    std::vector<u64> sk;
    std::vector<u32> si;
    sorted_stream(h, off, ident, 0, 0, false, sk, si);
    const u64 T = sk.size();
    std::vector<u32> c(n * n, 0), captured(n * n, 0);
    const size_t per = (n + passes - 1) / passes;
    for (size_t r0 = 0; r0 < n; r0 += per) {
        const size_t r1 = std::min(n, r0 + per);
        for (u64 p = 0; p < T; ++p)
            join_walk_rows(sk.data(), si.data(), T, p, (u32)r0, (u32)r1, [&](u32 a, u32 b) { c[(size_t)a * n + b] += 1; });
        for (size_t i = r0; i < r1; ++i)
            for (size_t j = 0; j < n; ++j)
                if (i != j) captured[i * n + j] = c[std::min(i, j) * n + std::max(i, j)];
    }
    FILE* f = fopen(out_path, "wb");
    fwrite(captured.data(), 4, captured.size(), f);
    fclose(f);
    return 0;
}

// stripe layout: the kernel's loop structure (compare_kernels.cu join_stripe_kernel) with the lanes of a
// warp run one after the other and the ballot assembled on the host
static int stripe_main(int R, int n_warps, const std::vector<u64>& h, const std::vector<u64>& off, const char* out_path,
                       int upper_only, int chunk_rows, bool low32) {
    const int n = (int)off.size() - 1;
    const u64 T = h.size();
    std::vector<u32> src(T);
    std::iota(src.begin(), src.end(), 0);
    std::vector<u64> sk(T);
    if (!low32) {
        std::stable_sort(src.begin(), src.end(), [&](u32 a, u32 b) { return h[a] < h[b]; });
        for (u64 q = 0; q < T; ++q) sk[q] = h[src[q]];
    } else {
        // SMB_JOIN_SORT=low32 (compare_kernels.cu stripe_stream_low32): stable sort on the low words, gather the
        // keys, flag the runs that mix hashes, re-sort the flagged elements on the rotated key, put them back
        std::stable_sort(src.begin(), src.end(), [&](u32 a, u32 b) { return (u32)h[a] < (u32)h[b]; });
        std::vector<u32> low(T);
        for (u64 q = 0; q < T; ++q) { sk[q] = h[src[q]]; low[q] = (u32)sk[q]; }
        std::vector<char> flags(T, 0);
        for (u64 q = 0; q < T; ++q) {
            bool mixed;
            const u64 m = stripe_run_at_head(low.data(), sk.data(), T, q, mixed);
            if (mixed) for (u64 j = 0; j < m; ++j) flags[q + j] = 1;
        }
        std::vector<u32> where;
        for (u64 q = 0; q < T; ++q) if (flags[q]) where.push_back((u32)q);
        std::vector<size_t> order(where.size());
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
            return stripe_rotated_key(sk[where[a]]) < stripe_rotated_key(sk[where[b]]); });
        std::vector<u64> k2(where.size());
        std::vector<u32> s2(where.size());
        for (size_t j = 0; j < where.size(); ++j) { k2[j] = stripe_rotated_key(stripe_rotated_key(sk[where[order[j]]])); s2[j] = src[where[order[j]]]; }
        for (size_t j = 0; j < where.size(); ++j) { sk[where[j]] = k2[j]; src[where[j]] = s2[j]; }
        // what the stripe layout needs from the stream: equal hashes contiguous, rows ascending inside a group
        for (u64 q = 1; q < T; ++q) {
            if (sk[q] == sk[q - 1] && stripe_row_of(off.data(), n, src[q]) <= stripe_row_of(off.data(), n, src[q - 1])) return 8;
            if (sk[q] != sk[q - 1]) for (u64 p2 = q + 1; p2 < T && (u32)sk[p2] == (u32)sk[q - 1]; ++p2) if (sk[p2] == sk[q - 1]) return 8;
        }
        if (getenv("SMB_EMUL_REPORT")) fprintf(stderr, "repaired %zu elements\n", where.size());
    }
    std::vector<u32> tags(T), pos(T);
    for (u64 q = 0; q < T; ++q) {
        tags[q] = stripe_make_tag(sk.data(), q, stripe_row_of(off.data(), n, src[q]));
        pos[src[q]] = (u32)q;
    }
    std::vector<double> out((size_t)n * n, -1.0);
    for (int r0 = 0; r0 < n; r0 += R) {
        const int r1 = std::min(n, r0 + R), rows = r1 - r0;
        std::vector<u64> s_off(off.begin() + r0, off.begin() + r1 + 1);
        std::vector<u32> stripe((size_t)rows * n, 0);
        const u64 e_begin = s_off[0], e_end = s_off[rows];
        auto scan = [&](u64 q, u32* row) {
            u32 tag[32];
            bool valid[32];
            for (u32 it = 0;; ++it) {                                   // forward
                u32 m = 0;
                for (u32 l = 0; l < 32; ++l) m |= (stripe_fwd_stop(tags.data(), T, q, it, l, tag[l]) ? 1u : 0u) << l;
                for (u32 l = 0; l < 32; ++l) if (stripe_fwd_active(m, l)) row[tag[l] & ~STRIPE_HEAD] += 1;
                if (!stripe_continue(m)) break;
            }
            if (upper_only || (tags[q] & STRIPE_HEAD)) return;
            for (u32 it = 0;; ++it) {                                   // backward
                u32 m = 0;
                for (u32 l = 0; l < 32; ++l) m |= (stripe_bwd_stop(tags.data(), q, it, l, tag[l], valid[l]) ? 1u : 0u) << l;
                for (u32 l = 0; l < 32; ++l) if (stripe_bwd_active(m, l, valid[l])) row[tag[l] & ~STRIPE_HEAD] += 1;
                if (!stripe_continue(m)) break;
            }
        };
        for (int warp = 0; warp < n_warps; ++warp)
            for (u64 base = e_begin + (u64)warp * 32; base < e_end; base += (u64)n_warps * 32) {
                u32 my_q[32], my_row[32];
                for (u32 l = 0; l < 32; ++l) {
                    const u64 e = base + l;
                    const bool have = e < e_end;
                    my_q[l] = have ? pos[e] : 0u;
                    my_row[l] = have ? stripe_local_row(s_off.data(), rows, e) : 0u;
                }
                const u32 cnt = (u32)std::min<u64>(32, e_end - base);
                for (u32 j = 0; j < cnt; ++j) scan(my_q[j], stripe.data() + (size_t)my_row[j] * n);
            }
        for (u32 i = 0; i < (u32)rows * (u32)n; ++i) {
            const u32 al = i / (u32)n, j = i - al * (u32)n;
            const int row = r0 + (int)al;
            if (upper_only && j < (u32)row) continue;
            if (out[(size_t)row * n + j] != -1.0) return 6;              // every cell written once
            out[(size_t)row * n + j] = stripe_jaccard(stripe[i], s_off[al + 1] - s_off[al], off[j + 1] - off[j], (u32)row == j);
        }
    }
    if (upper_only) {
        // stripe_mirror_kernel, chunk of rows by chunk of rows like the host path: 32 x 32 tiles, tile row ti
        // (destination rows) x tile column tj <= ti, thread (tx, ty) reads (tj*32+ty, ti*32+tx), writes (ti*32+ty, tj*32+tx)
        for (int rb = 0; rb < n; rb += chunk_rows) {
            const int re = std::min(n, rb + chunk_rows);
            const int t0 = rb / 32, t1 = (re + 31) / 32;
            for (int ti = t0; ti < t1; ++ti)
                for (int tj = 0; tj < t1; ++tj) {
                    if (tj > ti) continue;
                    double tile[32][33];
                    for (int ty = 0; ty < 32; ++ty)
                        for (int tx = 0; tx < 32; ++tx) {
                            const int sr = tj * 32 + ty, sc = ti * 32 + tx;
                            tile[ty][tx] = (sr < n && sc < n) ? out[(size_t)sr * n + sc] : 0.0;
                        }
                    for (int ty = 0; ty < 32; ++ty)
                        for (int tx = 0; tx < 32; ++tx) {
                            const int dr = ti * 32 + ty, dc = tj * 32 + tx;
                            if (dr >= rb && dr < re && dc < dr && dc < n) {
                                if (out[(size_t)dr * n + dc] != -1.0) return 7;      // lower cells are written once
                                out[(size_t)dr * n + dc] = tile[tx][ty];
                            }
                        }
                }
        }
    }
    FILE* f = fopen(out_path, "wb");
    fwrite(out.data(), 8, out.size(), f);
    fclose(f);
    return 0;
}

int main(int argc, char** argv) {
    if (argc == 9) {                                        // <R> ... <out> <unused> stripe <n_warps> <unused>
        std::vector<u64> h = slurp<u64>(argv[2]), off = slurp<u64>(argv[3]);
        const int chunk = atoi(argv[8]);                    // 0: two directions; > 0: upper only, mirrored in chunks of rows
        return stripe_main(atoi(argv[1]), atoi(argv[7]), h, off, argv[4], chunk > 0, chunk, !strcmp(argv[5], "low32"));
    }
    if (argc == 8) {                                        // ... <out> <unused> rows <passes>
        std::vector<u64> h = slurp<u64>(argv[2]), off = slurp<u64>(argv[3]);
        return rows_main(atoi(argv[7]), h, off, argv[4]);
    }
    if (argc == 7) {
        std::vector<u64> h = slurp<u64>(argv[2]), off = slurp<u64>(argv[3]);
        return cluster_main(atoi(argv[1]), h, off, argv[4]);
    }
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
