// tile_emul.cu -- runs the intersection kernel's table code (sourmash_b200/csrc/split_table.cuh,
// compiled for the host) exactly as pairwise_tile_split_kernel drives it: three build phases over
// `nthreads` emulated threads, then warps of 32 lanes streaming the other rows in batches of
// 32*U elements with the warp-uniform verify branch and the per-lane crowded-bucket branch.
// Test infrastructure (no GPU needed).
//   usage: tile_emul <shift> <nb> <csr_in> <counts_out>
// csr_in: u64 n_rows, u64 offsets[n_rows+1], u64 hashes[...]; output: u32 counts[n_rows*n_rows]
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../sourmash_b200/csrc/split_table.cuh"

using namespace smb;

int main(int argc, char** argv) {
    if (argc != 5) { fprintf(stderr, "usage: tile_emul shift nb in out\n"); return 2; }
    const u32 shift = (u32)atoi(argv[1]);
    const int nb = atoi(argv[2]);
    FILE* f = fopen(argv[3], "rb");
    if (!f) return 3;
    u64 n = 0;
    if (fread(&n, 8, 1, f) != 1) return 3;
    std::vector<u64> off(n + 1);
    if (fread(off.data(), 8, n + 1, f) != n + 1) return 3;
    std::vector<u64> h(off[n]);
    if (off[n] && fread(h.data(), 8, off[n], f) != off[n]) return 3;
    fclose(f);
    const int nthreads = 96, U = 4;
    std::vector<u32> out(n * n, 0);
    for (u64 i = 0; i < n; ++i) {
        // table of row i (trailing 2^64-1 key stripped, like the kernel)
        int na = (int)(off[i + 1] - off[i]);
        int hm = 0;
        if (na > 0 && h[off[i] + na - 1] == SMB_U64_MAX) { --na; hm = 1; }
        std::vector<u32> lo(na + 2), hi(na + 2);
        std::vector<u16> dir(nb + 2);
        for (int t = 0; t < nthreads; ++t) split_table_load(lo.data(), hi.data(), dir.data(), h.data() + off[i], na, nb, t, nthreads);
        for (int t = 0; t < nthreads; ++t) split_table_heads(lo.data(), hi.data(), dir.data(), na, shift, t, nthreads);
        for (int t = 0; t < nthreads; ++t) split_table_flags(dir.data(), nb, t, nthreads);
        SplitTable tab{lo.data(), hi.data(), dir.data()};
        for (u64 j = 0; j < n; ++j) {
            int nbj = (int)(off[j + 1] - off[j]);
            int bmax = 0;
            if (nbj > 0 && h[off[j] + nbj - 1] == SMB_U64_MAX) { --nbj; bmax = 1; }
            const u64* row = h.data() + off[j];
            u32 cnt = 0;
            const int full = nbj - (nbj % (32 * U));
            int base = 0;
            for (; base < full; base += 32 * U) {
                bool hit[32]; u32 ov[32][U]; bool any = false;
                for (int lane = 0; lane < 32; ++lane) {
                    hit[lane] = false;
                    for (int u = 0; u < U; ++u) {
                        ov[lane][u] = 0;
                        hit[lane] |= split_probe_low(tab, row[base + u * 32 + lane], shift, ov[lane][u]);
                    }
                    any |= hit[lane];
                }
                for (int lane = 0; lane < 32; ++lane)
                    for (int u = 0; u < U; ++u) {
                        const u64 q = row[base + u * 32 + lane];
                        if (any) cnt += split_probe_verify(tab, q, shift);
                        if (ov[lane][u] & 1u) cnt += split_probe_rest(tab, q, shift);
                    }
            }
            for (; base < nbj; ++base) cnt += split_probe_verify(tab, row[base], shift) + split_probe_rest(tab, row[base], shift);
            out[i * n + j] = cnt + (u32)(hm & bmax);
        }
    }
    f = fopen(argv[4], "wb");
    fwrite(out.data(), 4, out.size(), f);
    fclose(f);
    return 0;
}
