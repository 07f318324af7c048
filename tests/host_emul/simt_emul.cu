// simt_emul.cu -- runs the KERNELS of sourmash_b200/csrc/*.cuh on the CPU through the SIMT
// emulator of simt.h (CTAs as cooperative fibers, real __syncthreads / warp collectives / shared memory) and
// writes their results for comparison with the oracle.  The launch geometry is shrunk (fewer threads, smaller
// shared memory) but the code is the code the GPU runs.  Test infrastructure for the CPU-only suite.
//   simt_emul stripe <R> <upper 0|1> <threads> <tag bits 16|32> <sort-key bits> <hashes.u64> <offsets.u64> <out.f64 n*n>
//   simt_emul ranges <P> <bitmap log2> <threads> <query.u64> <hashes.u64> <offsets.u64> <out.u32 n>
//   simt_emul join <key-range shards> <hashes.u64> <offsets.u64> <out.u32 n*n>              (upper-triangle counts)
//   simt_emul tile <TA 1..4> <variant 1 split | 0 u64 occ | 2 u64> <threads> <cols_per_cta> <symmetric 0|1> <hashes.u64> <offsets.u64> <out.u32>
//   simt_emul pairs <num> <hashes.u64> <offsets.u64> <out.f64 3*n*n: jaccard (generic kernel) | jaccard num | angular>
//   simt_emul gather <use_index 0|1> <threshold> <query.u64> <hashes.u64> <offsets.u64> <out.u32 (row, size) pairs>
//   simt_emul index  <threads> <query.u64> <hashes.u64> <offsets.u64> <out.u32 2n: direct | length on "device">
#define SMB_SIMT_EMUL 1
#include "simt.h"

#include <numeric>

#include "../../sourmash_b200/csrc/join_kernels.cuh"
#include "../../sourmash_b200/csrc/join_stripe.cuh"
#include "../../sourmash_b200/csrc/range_kernels.cuh"
#include "../../sourmash_b200/csrc/db_index_kernels.cuh"
#include "../../sourmash_b200/csrc/search_kernels.cuh"
#include "../../sourmash_b200/csrc/tile_kernels.cuh"
#include "../../sourmash_b200/csrc/pair_kernels.cuh"
#include <math.h>

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
template <class T>
static void dump(const char* path, const std::vector<T>& v) {
    FILE* f = fopen(path, "wb");
    fwrite(v.data(), sizeof(T), v.size(), f);
    fclose(f);
}

// The stripe pipeline as join_stripe_create / join_stripe_rows launch it (compare_kernels.cu), kernel by kernel;
// a stable host sort stands in for cub::DeviceRadixSort::SortPairs.  `sort_bits` < 64 shortens the sort key
// below the 32 bits the product uses, so that runs mixing different hashes -- rare with 32-bit keys -- occur
// in every small test set and the descent detection of the tag kernel / the fix kernel are exercised.
template <typename TagT>
static int stripe_run(int R, int upper, int threads, int sort_bits, std::vector<u64>& h, const std::vector<u64>& off,
                      const char* fout) {
    const int n = (int)off.size() - 1;
    const u64 T = h.size();
    u64 max_key = 0;
    for (u64 v : h) max_key = std::max(max_key, v);
    int key_bits = 1;
    while (key_bits < 64 && (max_key >> key_bits)) ++key_bits;
    int low_bits = stripe_low_bits(max_key);
    if (sort_bits < 32 && key_bits > sort_bits) low_bits = std::min(32, key_bits - sort_bits);   // test hook: coarser keys (the payload holds at most 32 low bits)
    h.push_back(0);
    std::vector<u32> key32(T + 1), key32s(T + 1), worklist(T + 1), pos(T + 1, 0xffffffffu), sizes(n + 1);
    std::vector<u64> pays(T + 1), payss(T + 1);
    smb_emu::launch(3, 64, 0, [&] { stripe_keys_kernel(h.data(), T, low_bits, key32.data(), pays.data()); });
    std::vector<u32> perm(T);
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](u32 a, u32 b) { return key32[a] < key32[b]; });
    for (u64 q = 0; q < T; ++q) { key32s[q] = key32[perm[q]]; payss[q] = pays[perm[q]]; }
    const u64 nblk = (T >> STRIPE_EBLK_LOG2) + 1;
    std::vector<u32> eblk(nblk + 1);
    smb_emu::launch(2, 64, 0, [&] { stripe_eblk_kernel(off.data(), n, T, eblk.data()); });
    smb_emu::launch((n + 63) / 64 + 1, 64, 0, [&] { stripe_sizes_kernel(off.data(), n, sizes.data()); });
    std::vector<TagT> tags(T + STRIPE_TAG_PAD + 1);
    u32 count = 0;
    u32* wl = low_bits ? worklist.data() : nullptr;
    const u32 swz = (R & 1) ? ((u32)n & ~31u) : 0u;          // both column orders over the test configurations
    smb_emu::launch(2, 96, 0, [&] { stripe_tag_kernel<TagT>(key32s.data(), payss.data(), off.data(), eblk.data(), T, tags.data(), pos.data(), wl, &count, swz); });
    if (low_bits)            // pays (the sort's input) is free: the ordered payloads of the repaired runs
        smb_emu::launch(2, 64, 0, [&] { stripe_fix_kernel<TagT>(key32s.data(), payss.data(), T, worklist.data(), &count, off.data(), n, pays.data(), tags.data(), pos.data(), swz); });
    if (getenv("SMB_EMUL_REPORT")) fprintf(stderr, "mixed runs repaired: %u\n", count);
    // the stream must now be the hashes in ascending order, rows ascending inside a group: pos[] a permutation, tags = owning
    // row, head flag exactly where the hash changes
    {
        std::vector<u32> inv(T, 0xffffffffu);
        for (u64 e = 0; e < T; ++e) {
            if (pos[e] >= T || inv[pos[e]] != 0xffffffffu) return 5;
            inv[pos[e]] = (u32)e;
        }
        constexpr u32 HEAD = StripeTag<TagT>::HEAD;
        for (u64 q = 0; q < T; ++q) {
            const u64 x = h[inv[q]];
            if (q + 1 < T) { const u64 y = h[inv[q + 1]]; if (x > y || (x == y && inv[q] > inv[q + 1])) return 5; }
            const bool head = q == 0 || h[inv[q - 1]] != x;
            if (((u32)tags[q] & HEAD ? true : false) != head) return 7;
            if (((u32)tags[q] & ~HEAD) != stripe_col(stripe_row_of(off.data(), n, inv[q]), swz)) return 8;
        }
    }
    std::vector<double> out((size_t)n * n, -1.0);
    const size_t smem = (size_t)STRIPE_HEADER + (size_t)R * n * sizeof(u32);
    // two launches over row chunks, like the host path of smb_compare_jaccard
    const int half = n / 2;
    for (int c = 0; c < 2; ++c) {
        const int r0 = c ? half : 0, r1 = c ? n : half;
        if (r1 <= r0) continue;
        StripeArgs a{tags.data(), pos.data(), off.data(), off.data() + 1, sizes.data(), T, n, R, r0, r1, out.data() + (size_t)r0 * n, nullptr, nullptr, swz};
        const int blocks = (r1 - r0 + R - 1) / R;
        if (upper) smb_emu::launch(blocks, threads, smem, [&] { join_stripe_kernel<TagT, true, 1>(a); });
        else smb_emu::launch(blocks, threads, smem, [&] { join_stripe_kernel<TagT, false, 2>(a); });
        if (upper) {
            const int t0 = r0 / STRIPE_MIRROR_TILE, t1 = (r1 + STRIPE_MIRROR_TILE - 1) / STRIPE_MIRROR_TILE;
            smb_emu::launch(smb_emu::Dim3(t1, t1 - t0), 1024, 0, [&] { stripe_mirror_kernel<double>(out.data(), n, r0, r1); });
        }
    }
    dump(fout, out);
    return 0;
}
static int stripe_main(int R, int upper, int threads, int tag_bits, int sort_bits, const char* fh, const char* fo, const char* fout) {
    std::vector<u64> h = slurp<u64>(fh), off = slurp<u64>(fo);
    if (R > STRIPE_MAX_ROWS) return 6;
    return tag_bits == 16 ? stripe_run<u16>(R, upper, threads, sort_bits, h, off, fout)
                          : stripe_run<u32>(R, upper, threads, sort_bits, h, off, fout);
}

// ---- global-reduction join (the stripe layout's fallback): row slices, gather, count, estimate (kernels as written; the
// exclusive scan and the radix sort of the product are host loops here) ----
struct Slice { std::vector<u64> keys; std::vector<u32> ids; };
static Slice sorted_slice(const std::vector<u64>& h, const std::vector<u64>& off, int n, u64 lo, u64 hi, bool bounded) {
    std::vector<u64> beg(n + 1), cnt(n + 1), doff(n + 2, 0);
    smb_emu::launch((n + 1 + 63) / 64, 64, 0, [&] { join_row_range_kernel(h.data(), off.data(), n, lo, hi, bounded ? 1 : 0, beg.data(), cnt.data()); });
    const std::vector<u64>& c = cnt;
    for (int r = 0; r <= n; ++r) doff[r + 1] = doff[r] + c[r];                     // cub::DeviceScan::ExclusiveSum
    const u64 T = doff[n];
    std::vector<u64> keys(T + 1);
    std::vector<u32> ids(T + 1);
    smb_emu::launch(7, 64, 0, [&] { join_gather_kernel(h.data(), off.data(), beg.data(), doff.data(), n, keys.data(), ids.data()); });
    std::vector<size_t> perm(T);
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](size_t a, size_t b) { return keys[a] < keys[b]; });   // cub::DeviceRadixSort::SortPairs
    Slice S;
    S.keys.resize(T + 1); S.ids.resize(T + 1);
    for (u64 i = 0; i < T; ++i) { S.keys[i] = keys[perm[i]]; S.ids[i] = ids[perm[i]]; }
    S.keys.resize(T); S.ids.resize(T);
    return S;
}

static int join_main(const char* mode, int arg, const char* fh, const char* fo, const char* fout) {
    std::vector<u64> h = slurp<u64>(fh), off = slurp<u64>(fo);
    const int n = (int)off.size() - 1;
    u64 max_key = 0;
    for (u64 v : h) max_key = std::max(max_key, v);
    h.push_back(0);
    std::vector<u32> common((size_t)n * n + 1, 0);
    if (!strcmp(mode, "join")) {                                             // arg = number of key-range shards
        unsigned long long est[2] = {0, 0}, total = 0;
        for (int shard = 0; shard < arg; ++shard) {
            u64 lo, hi;
            bool bounded;
            join_shard_range(max_key, shard, arg, lo, hi, bounded);
            Slice S = sorted_slice(h, off, n, lo, hi, bounded);
            const u64 T = S.keys.size();
            S.keys.push_back(0); S.ids.push_back(0);
            if (!T) continue;
            smb_emu::launch((T + 63) / 64, 64, 0, [&] { join_count_kernel(S.keys.data(), S.ids.data(), T, common.data(), (size_t)n); });
            smb_emu::launch(3, 96, 0, [&] { join_estimate_kernel(S.keys.data(), T, est); });
        }
        for (size_t i = 0; i < (size_t)n * n; ++i) total += common[i];
        if (est[0] != total) return 3;                                       // sum of C(m,2) == number of increments
    } else return 2;
    common.resize((size_t)n * n);
    dump(fout, common);
    return 0;
}

static std::vector<u32> host_dir(const std::vector<u64>& keys, u32 shift, u64 nbk) {
    std::vector<u32> dir(nbk + 2);
    for (u64 b = 0; b <= nbk; ++b)
        dir[b] = (u32)(std::lower_bound(keys.begin(), keys.end(), b, [&](u64 k, u64 bb) { return (k >> shift) < bb; }) - keys.begin());
    return dir;
}

// The range-major pass as range_major_build / launch_one_vs_many_range_major launch it (compare_kernels.cu): bounds,
// counts, (host) exclusive scan, scatter, then the streaming kernel.  `bm_log2` far below the product's 19 makes
// false positives of the bitmap -- and with them the exact test of the drain -- common.
static int ranges_main(int P, int bm_log2, int threads, const char* fq, const char* fh, const char* fo, const char* fout) {
    std::vector<u64> q = slurp<u64>(fq), h = slurp<u64>(fh), off = slurp<u64>(fo);
    const int n = (int)off.size() - 1;
    const u64 T = h.size();
    u64 max_key = 0;
    for (u64 v : h) max_key = std::max(max_key, v);
    const u64 width = range_width(max_key, P);
    const size_t cells = (size_t)n * P;
    std::vector<u32> bounds(cells + n + 1, 0xdeadbeefu), cnt(cells + 2), slice(cells + 2);
    h.push_back(0);                                             // keep h.data() valid for empty sets
    smb_emu::launch(3, 64, 0, [&] { rm_bounds_kernel(h.data(), off.data(), n, width, P, bounds.data()); });
    smb_emu::launch(2, 96, 0, [&] { rm_counts_kernel(bounds.data(), n, P, cnt.data()); });
    u32 run = 0;
    for (size_t i = 0; i <= cells; ++i) { slice[i] = run; run += cnt[i]; }       // cub::DeviceScan::ExclusiveSum
    if (slice[cells] != (u32)T) return 3;                                        // every element lies in exactly one range
    std::vector<u64> rm(T + 64, 0);
    smb_emu::launch(3, 64, 0, [&] { rm_scatter_kernel(h.data(), off.data(), bounds.data(), slice.data(), n, P, rm.data()); });
    std::vector<u32> out(n + 1, 0);
    const u64 nq = q.size();
    q.push_back(0);
    const int nc = (n >> RM_COARSE_LOG2) + 2;
    std::vector<u32> coarse((size_t)P * nc + 1);
    smb_emu::launch(2, 64, 0, [&] { rm_coarse_kernel(slice.data(), n, P, nc, coarse.data()); });
    RangeMajorArgs a{q.data(), nq, rm.data(), slice.data(), coarse.data(), nc, n, P, width, (u32)bm_log2,
                     range_bitmap_shift(width, 1ull << bm_log2), out.data()};
    const size_t smem = (bm_log2 > 5 ? ((size_t)1 << (bm_log2 - 3)) : 4) + (size_t)(threads / 32) * RM_QUEUE * sizeof(u32);
    if (nq && n) smb_emu::launch(P, threads, smem, [&] { one_vs_many_range_major_kernel(a); });
    out.resize(n);
    dump(fout, out);
    return 0;
}

static int index_main(int threads, const char* fq, const char* fh, const char* fo, const char* fout) {
    std::vector<u64> q = slurp<u64>(fq), h = slurp<u64>(fh), off = slurp<u64>(fo);
    const int n = (int)off.size() - 1;
    const u64 T = h.size();
    std::vector<u32> ids(T + 1);
    smb_emu::launch(5, 64, 0, [&] { index_rowid_kernel(off.data(), n, ids.data()); });
    std::vector<size_t> perm(T);
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](size_t a, size_t b) { return h[a] < h[b]; });
    std::vector<u64> keys;
    std::vector<u32> start, rows(T + 1);
    for (u64 p = 0; p < T; ++p) {
        rows[p] = ids[perm[p]];
        if (p == 0 || h[perm[p]] != h[perm[p - 1]]) { keys.push_back(h[perm[p]]); start.push_back((u32)p); }
    }
    start.push_back((u32)T);
    u64 max_key = 0;
    for (u64 v : h) max_key = std::max(max_key, v);
    u32 shift;
    u64 nbk;
    db_index_dir_plan(keys.size(), max_key, shift, nbk);
    std::vector<u32> dir = host_dir(keys, shift, nbk);
    keys.push_back(0);
    DbIndexView ix{keys.data(), (u64)keys.size() - 1, start.data(), rows.data(), dir.data(), shift, nbk};
    std::vector<u32> out(2 * (size_t)n + 1, 0);
    const u64 nq = q.size();
    q.push_back(0);
    if (nq && ix.n_keys) {
        smb_emu::launch(3, threads, 0, [&] { index_count_kernel(ix, q.data(), nq, nullptr, out.data()); });
        u32 d_nq = (u32)nq;                                      // the gather loop's variant: length read on the "device"
        smb_emu::launch(2, threads, 0, [&] { index_count_kernel(ix, q.data(), nq + 1000, &d_nq, out.data() + n); });
    }
    out.resize(2 * (size_t)n);
    dump(fout, out);
    return 0;
}

// ---- search / gather: the default kernels as launched (directory + bitmap over the query, global one-vs-many
// pass, and the rounds of the gather session), optionally with the inverted index for the per-round counts ----
static void device_dir(const std::vector<u64>& keys, u64 nk, u32 shift, u64 nb, std::vector<u32>& dir) {
    dir.assign(nb + 2, 0);
    smb_emu::launch(3, 64, 0, [&] { global_dir_fill_kernel(dir.data(), nb + 1, DIR_UNSET); });
    if (nk) smb_emu::launch(3, 64, 0, [&] { global_dir_heads_kernel(keys.data(), nk, shift, dir.data()); });
    smb_emu::launch(3, 64, 0, [&] { global_dir_resolve_kernel(keys.data(), nk, shift, nb, dir.data()); });
}
// one_vs_many_dev's large-query branch (capi.cu)
static void one_vs_many_global(const u64* q, u64 nq, const std::vector<u64>& h, const std::vector<u64>& off, int n, u32* counts) {
    if (!nq || !n) return;
    const u64 q_max = q[nq - 1];
    int nb_log2 = 12;
    while (nb_log2 < 26 && (1ull << nb_log2) < 2 * nq) ++nb_log2;
    u32 shift = 0;
    while (shift < 63 && (q_max >> shift) >= (1ull << nb_log2)) ++shift;
    const u64 nb = (q_max >> shift) + 1;
    std::vector<u64> qv(q, q + nq);
    qv.push_back(0);
    std::vector<u32> dir;
    device_dir(qv, nq, shift, nb, dir);
    const int fine_log2 = 3;
    std::vector<u32> bm((((size_t)nb << fine_log2) / 32) + 2, 0);
    smb_emu::launch(3, 64, 0, [&] { build_query_bitmap_kernel(qv.data(), nq, shift, fine_log2, bm.data()); });
    smb_emu::launch(4, 64, 0, [&] { one_vs_many_global_kernel(qv.data(), nq, dir.data(), shift, nb, bm.data(), fine_log2, h.data(), off.data(), n, counts); });
}

static int gather_main(int use_index, u32 threshold, const char* fq, const char* fh, const char* fo, const char* fout) {
    std::vector<u64> q = slurp<u64>(fq), h = slurp<u64>(fh), off = slurp<u64>(fo);
    const int n = (int)off.size() - 1;
    const u64 nq = q.size(), T = h.size();
    u64 max_len = 1;
    for (int r = 0; r < n; ++r) max_len = std::max<u64>(max_len, off[r + 1] - off[r]);
    h.push_back(0); q.push_back(0);
    // inverted index of the database, its directory built by the device kernels
    std::vector<u64> keys;
    std::vector<u32> start, rows(T + 1), idir;
    DbIndexView ix{};
    if (use_index && T) {
        std::vector<u32> ids(T + 1);
        smb_emu::launch(5, 64, 0, [&] { index_rowid_kernel(off.data(), n, ids.data()); });
        std::vector<size_t> perm(T);
        std::iota(perm.begin(), perm.end(), 0);
        std::stable_sort(perm.begin(), perm.end(), [&](size_t a, size_t b) { return h[a] < h[b]; });
        for (u64 p = 0; p < T; ++p) {
            rows[p] = ids[perm[p]];
            if (p == 0 || h[perm[p]] != h[perm[p - 1]]) { keys.push_back(h[perm[p]]); start.push_back((u32)p); }
        }
        start.push_back((u32)T);
        u64 max_key = 0;
        for (u64 i = 0; i < T; ++i) max_key = std::max(max_key, h[i]);
        u32 shift;
        u64 nbk;
        db_index_dir_plan(keys.size(), max_key, shift, nbk);
        const u64 nkeys = keys.size();
        keys.push_back(0);
        device_dir(keys, nkeys, shift, nbk, idir);
        ix = DbIndexView{keys.data(), nkeys, start.data(), rows.data(), idir.data(), shift, nbk};
    }
    auto counts_of = [&](const u64* x, u64 nx, const u32* d_nx, u32* out) {
        if (use_index && T) smb_emu::launch(2, 64, 0, [&] { index_count_kernel(ix, x, d_nx ? max_len : nx, d_nx, out); });
        else one_vs_many_global(x, d_nx ? *d_nx : nx, h, off, n, out);
    };
    std::vector<u32> counts(n + 1, 0), delta(n + 1, 0);
    std::vector<u8> alive(nq + 1, 1);
    std::vector<u64> isect(max_len + 1);
    u32 d_n = 0;
    unsigned long long best[2] = {0, 0};
    if (nq) counts_of(q.data(), nq, nullptr, counts.data());
    std::vector<u32> result;
    bool have_delta = false;
    u64 remaining = nq;
    for (int round = 0; round < n && nq; ++round) {
        smb_emu::launch(1, 64, 0, [&] { counter_update_argmax_kernel(counts.data(), have_delta ? delta.data() : nullptr, n, best); });
        if (best[0] < threshold || best[0] == 0) break;
        const u32 row = (u32)best[1];
        smb_emu::launch(1, 96, 0, [&] { intersect_alive_kernel(q.data(), nq, alive.data(), h.data() + off[row], off[row + 1] - off[row], isect.data(), &d_n); });
        if (d_n != best[0]) return 3;                        // the counter of the winner == its live intersection
        result.push_back(row); result.push_back(d_n);
        std::fill(delta.begin(), delta.end(), 0u);
        counts_of(isect.data(), d_n, &d_n, delta.data());
        smb_emu::launch(2, 64, 0, [&] { mark_dead_n_kernel(q.data(), nq, alive.data(), isect.data(), &d_n); });
        have_delta = true;
        remaining -= d_n;
        if (!remaining) break;
    }
    dump(fout, result);
    return 0;
}

// ---- the pair-by-pair intersection kernels as launched (launch_tile_ta of compare_kernels.cu) ----
template <int TA>
static void run_tile(const TileArgs& a, size_t smem, int threads, int variant) {
    const int tiles = (a.nA + TA - 1) / TA;
    smb_emu::Dim3 grid(tiles, (a.nB + a.cols_per_cta - 1) / a.cols_per_cta);
    if (variant == 1) smb_emu::launch(grid, threads, smem, [&] { pairwise_tile_split_kernel<TA, 4>(a); });
    else if (variant == 0) smb_emu::launch(grid, threads, smem, [&] { pairwise_tile_kernel<TA, 4, true>(a); });
    else smb_emu::launch(grid, threads, smem, [&] { pairwise_tile_kernel<TA, 4, false>(a); });
}
static int tile_main(int ta, int variant, int threads, int cols, int symmetric, const char* fh, const char* fo, const char* fout) {
    std::vector<u64> h = slurp<u64>(fh), off = slurp<u64>(fo);
    const int n = (int)off.size() - 1;
    u64 max_len = 0, max_key = 0;
    for (int r = 0; r < n; ++r) max_len = std::max<u64>(max_len, off[r + 1] - off[r]);
    for (u64 v : h) if (v != SMB_U64_MAX) max_key = std::max(max_key, v);
    if (!h.empty() && h.back() == SMB_U64_MAX && max_key == 0) max_key = SMB_U64_MAX;
    for (u64 v : h) max_key = std::max(max_key, v);               // the planner sees the raw maximum (set_max_key)
    PairwisePlan plan = plan_pairwise_impl(max_len, max_key, cols);
    if (plan.tables_per_cta == 0) return 3;
    if (ta > plan.tables_per_cta) return 4;
    const size_t key_bytes = ((size_t)plan.cap + 2) * 8;
    const size_t smem = (size_t)ta * (key_bytes + ((size_t)plan.nb + 2) * 2);
    h.push_back(0);
    // symmetric: A == B, upper triangle; otherwise the first half of the rows against the second half
    const int nA = symmetric ? n : n / 2, nB = symmetric ? n : n - n / 2;
    const u64* offB = symmetric ? off.data() : off.data() + nA;
    std::vector<u32> out((size_t)nA * nB + 1, 0xdeadbeefu);
    TileArgs a{h.data(), off.data(), nA, h.data(), offB, nB, out.data(), (size_t)nB, plan.shift, plan.nb, plan.cap,
               plan.cols_per_cta, symmetric, 0, 1, -1};
    switch (ta) {
        case 1: run_tile<1>(a, smem, threads, variant); break;
        case 2: run_tile<2>(a, smem, threads, variant); break;
        case 3: run_tile<3>(a, smem, threads, variant); break;
        default: run_tile<4>(a, smem, threads, variant); break;
    }
    out.resize((size_t)nA * nB);
    dump(fout, out);
    return 0;
}

// ---- the remaining pair kernels: generic, bottom-k, angular, finalize (compare_jaccard_impl / smb_compare_angular) ----
static int pairs_main(u32 num, const char* fh, const char* fo, const char* fout) {
    std::vector<u64> h = slurp<u64>(fh), off = slurp<u64>(fo);
    const int n = (int)off.size() - 1;
    const size_t nn = (size_t)n * n;
    h.push_back(0);
    unsigned long long d_max = 0, want_max = 0;
    smb_emu::launch(2, 64, 0, [&] { max_last_kernel(h.data(), off.data(), n, nullptr, nullptr, 0, &d_max); });
    for (int r = 0; r < n; ++r) if (off[r + 1] > off[r]) want_max = std::max<unsigned long long>(want_max, h[off[r + 1] - 1]);
    if (d_max != want_max) return 3;
    std::vector<u32> common(nn + 1, 0), common_num(nn + 1, 0), usize(nn + 1, 0);
    std::vector<double> out(3 * nn + 1, -1.0);
    // scaled: generic kernel (rows of any size) + finalize, whole matrix and a block of rows
    smb_emu::launch(3, 64, 0, [&] { pairwise_generic_kernel(h.data(), off.data(), n, h.data(), off.data(), n, common.data(), (size_t)n, 1, 0, 1); });
    smb_emu::launch(smb_emu::Dim3((n + 63) / 64, n), 64, 0,
                    [&] { finalize_matrix_kernel(common.data(), nullptr, (size_t)n, off.data(), off.data(), n, n, 0, 1, out.data()); });
    {
        const int r0 = n / 3, r1 = std::min(n, r0 + 7);
        std::vector<double> rows((size_t)(r1 - r0) * n + 1, -1.0);
        smb_emu::launch(smb_emu::Dim3((n + 63) / 64, r1 - r0), 64, 0, [&] { finalize_rows_kernel(common.data(), (size_t)n, off.data(), r0, r1, rows.data()); });
        for (int i = r0; i < r1; ++i) for (int j = 0; j < n; ++j) if (rows[(size_t)(i - r0) * n + j] != out[(size_t)i * n + j]) return 4;
    }
    // bottom-k: num kernel + finalize mode 1
    smb_emu::launch(3, 64, 0, [&] { pairwise_num_kernel(h.data(), off.data(), n, h.data(), off.data(), n, num, common_num.data(), usize.data(), (size_t)n, 1); });
    smb_emu::launch(smb_emu::Dim3((n + 63) / 64, n), 64, 0,
                    [&] { finalize_matrix_kernel(common_num.data(), usize.data(), (size_t)n, off.data(), off.data(), n, n, 1, 1, out.data() + nn); });
    // abundances 1 + hash % 5: angular similarity
    std::vector<u64> ab(h.size());
    for (size_t i = 0; i < h.size(); ++i) ab[i] = 1 + h[i] % 5;
    std::vector<unsigned long long> sumsq(n + 1, 0);
    smb_emu::launch(2, 64, 0, [&] { row_sumsq_kernel(ab.data(), off.data(), n, sumsq.data()); });
    smb_emu::launch(3, 64, 0, [&] { pairwise_angular_kernel(h.data(), ab.data(), off.data(), n, sumsq.data(), out.data() + 2 * nn); });
    out.resize(3 * nn);
    dump(fout, out);
    return 0;
}

int main(int argc, char** argv) {
    if (argc == 6 && !strcmp(argv[1], "pairs")) return pairs_main((u32)atoi(argv[2]), argv[3], argv[4], argv[5]);
    if (argc == 10 && !strcmp(argv[1], "tile"))
        return tile_main(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), argv[7], argv[8], argv[9]);
    if (argc == 8 && !strcmp(argv[1], "gather")) return gather_main(atoi(argv[2]), (u32)atoi(argv[3]), argv[4], argv[5], argv[6], argv[7]);
    if (argc == 10 && !strcmp(argv[1], "stripe"))
        return stripe_main(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), argv[7], argv[8], argv[9]);
    if (argc == 9 && !strcmp(argv[1], "ranges"))
        return ranges_main(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), argv[5], argv[6], argv[7], argv[8]);
    if (argc == 7 && !strcmp(argv[1], "index")) return index_main(atoi(argv[2]), argv[3], argv[4], argv[5], argv[6]);
    if (argc == 6 && !strcmp(argv[1], "join"))
        return join_main(argv[1], atoi(argv[2]), argv[3], argv[4], argv[5]);
    fprintf(stderr, "usage: see the head of simt_emul.cu\n");
    return 2;
}
