// simt_sketch_emul.cu -- runs the rolled hashing KERNELS of sourmash_b200/csrc/sketch_device.cuh
// (hash_kmers_kernel<K> and the experimental one-pass hash_kmers_fused_kernel) on the CPU through simt.h:
// the tiling over streams, per-thread windows, survivor staging in shared memory (and its overflow path),
// the flush to the candidate rows.  Output: per row the sorted distinct candidates, for comparison with the
// oracle.  Test infrastructure for the CPU-only suite.
//   simt_sketch_emul <k: 21|31|51, or 0 = fused 21+31+51> <W> <max_hash> <lead> <seqs.u8> <offsets.u64> <out.u64>
// `lead` junk bytes precede the first stream, so that streams start at unaligned addresses.
// out: for every row (stream-major, then k ascending for the fused pass): count, then the hashes.
// With SMB_EMUL_DEVICE_SORT set the rows are materialised by sort_unique_small_kernel / unique_sorted_row_kernel
// and every row is followed by its abundances (run lengths).
#define SMB_SIMT_EMUL 1
#include "simt.h"

#include "../../sourmash_b200/csrc/sketch_device.cuh"

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
    if (argc != 8) return 2;
    const int k = atoi(argv[1]), W = atoi(argv[2]);
    const u64 max_hash = strtoull(argv[3], nullptr, 10);
    const u32 lead0 = (u32)atoi(argv[4]);
    std::vector<u8> seqs = slurp<u8>(argv[5]);
    std::vector<u64> offs = slurp<u64>(argv[6]);
    const int ns = (int)offs.size() - 1;
    const int nk = k ? 1 : 3;
    // device-like buffers: 16-byte aligned base, streams shifted by `lead0`, readable up to the next line
    std::vector<u8> raw(lead0 + seqs.size() + 64 + 16, 'G');
    u8* base = raw.data() + ((16 - ((uintptr_t)raw.data() & 15)) & 15);
    memcpy(base + lead0, seqs.data(), seqs.size());
    std::vector<u64> s_off(ns), s_len(ns);
    std::vector<u32> tile_start(ns + 1, 0);
    for (int i = 0; i < ns; ++i) {
        s_off[i] = lead0 + offs[i]; s_len[i] = offs[i + 1] - offs[i];
        const u64 lp = (s_off[i] & 15) + s_len[i], per = (u64)HASH_THREADS * W;
        tile_start[i + 1] = tile_start[i] + (u32)((lp + per - 1) / per);
    }
    const int n_rows = ns * nk;
    std::vector<u64> cand_off(n_rows + 1, 0);
    for (int i = 0; i < ns; ++i) for (int j = 0; j < nk; ++j) cand_off[i * nk + j + 1] = cand_off[i * nk + j] + s_len[i] + 1;
    std::vector<u64> cand(cand_off[n_rows] + 1, 0);
    std::vector<u32> cnt(n_rows + 1, 0);
    HashArgs a{};
    a.bases = base; a.stream_off = s_off.data(); a.stream_len = s_len.data(); a.stream_row = nullptr;
    a.tile_start = tile_start.data(); a.n_streams = ns; a.W = W; a.seed = 42; a.max_hash = max_hash;
    a.cand = cand.data(); a.cand_off = cand_off.data(); a.cand_cnt = cnt.data();
    a.tile_base = 0; a.row_stride = nk; a.row_index = 0; a.raw_out = nullptr;
    const unsigned tiles = tile_start[ns];
    if (tiles) {
        if (k == 0) {
            FusedArgs f{};
            f.a = a;
            for (int i = 0; i < 3; ++i) { f.row_index[i] = i; f.max_hash[i] = max_hash; }
            // two launches over tile ranges, like the upload-overlapped path of sketch_streams
            const unsigned mid = tiles / 2;
            f.a.tile_base = 0;
            if (mid) smb_emu::launch(mid, HASH_THREADS, 0, [&] { hash_kmers_fused_kernel(f); });
            f.a.tile_base = mid;
            smb_emu::launch(tiles - mid, HASH_THREADS, 0, [&] { hash_kmers_fused_kernel(f); });
        } else if (k == 21) smb_emu::launch(tiles, HASH_THREADS, 0, [&] { hash_kmers_kernel<21, false>(a); });
        else if (k == 31) smb_emu::launch(tiles, HASH_THREADS, 0, [&] { hash_kmers_kernel<31, false>(a); });
        else if (k == 51) smb_emu::launch(tiles, HASH_THREADS, 0, [&] { hash_kmers_kernel<51, false>(a); });
        else return 4;
    }
    for (int r = 0; r < n_rows; ++r) if (cnt[r] > cand_off[r + 1] - cand_off[r]) return 5;   // more candidates than windows
    FILE* f = fopen(argv[7], "wb");
    if (getenv("SMB_EMUL_DEVICE_SORT")) {
        // row materialisation on the "device": sort_unique_small_kernel (rows up to SORT_MAX candidates) and, for a
        // longer row, a host sort standing in for cub + unique_sorted_row_kernel; abundances = run lengths
        std::vector<u32> ucnt(n_rows + 1, 0);
        std::vector<u64> abund(cand.size() + 1, 0);
        const size_t smem = (size_t)SORT_MAX * 8 + (size_t)SORT_MAX * 4;
        smb_emu::launch(n_rows, SORT_THREADS, smem, [&] { sort_unique_small_kernel(cand.data(), cand_off.data(), cnt.data(), ucnt.data(), abund.data()); });
        for (int r = 0; r < n_rows; ++r) {
            if (cnt[r] > (u32)SORT_MAX) {
                std::vector<u64> sorted(cand.begin() + cand_off[r], cand.begin() + cand_off[r] + cnt[r]), heads(cnt[r] + 1);
                std::sort(sorted.begin(), sorted.end());
                smb_emu::launch(1, SORT_THREADS, 0, [&] { unique_sorted_row_kernel(sorted.data(), (u64)cnt[r], cand.data() + cand_off[r], abund.data() + cand_off[r], heads.data(), ucnt.data() + r); });
            }
            const u64 n = ucnt[r];
            fwrite(&n, 8, 1, f);
            fwrite(cand.data() + cand_off[r], 8, n, f);
            fwrite(abund.data() + cand_off[r], 8, n, f);
        }
        fclose(f);
        return 0;
    }
    for (int r = 0; r < n_rows; ++r) {
        std::vector<u64> v(cand.begin() + cand_off[r], cand.begin() + cand_off[r] + cnt[r]);
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        const u64 n = v.size();
        fwrite(&n, 8, 1, f);
        fwrite(v.data(), 8, v.size(), f);
    }
    fclose(f);
    return 0;
}
