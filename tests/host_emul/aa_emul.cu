// aa_emul.cu -- runs the *device* protein-family k-mer code (sourmash_b200/csrc/aa_kmers.cuh,
// compiled here for the host) over a sequence, tile by tile exactly like hash_aa_kernel:
// stage the residues of a tile, then hash every window start of the tile from the staged arrays.
// Test infrastructure (CPU-only check of the kernel's logic against the oracle).
//   usage: aa_emul <hash_function 2|3|4> <kaa> <translate 0|1> <infile> <outfile>
// output = u64 per window in the reference's seq_to_hashes order.
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../sourmash_b200/csrc/aa_kmers.cuh"

using namespace smb;

int main(int argc, char** argv) {
    if (argc != 6) { fprintf(stderr, "usage: aa_emul hf kaa translate in out\n"); return 2; }
    const int hf = atoi(argv[1]);
    const u32 kaa = (u32)atoi(argv[2]);
    const bool translate = atoi(argv[3]) != 0;
    FILE* f = fopen(argv[4], "rb");
    if (!f) return 3;
    std::vector<u8> seq;
    u8 tmp[65536]; size_t n;
    while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) seq.insert(seq.end(), tmp, tmp + n);
    fclose(f);
    const u64 L = seq.size();
    const AaTables T = build_aa_tables(hf);
    const u32 stride = translate ? 3u : 1u;
    const u64 span = (u64)kaa * stride;
    std::vector<u64> out;
    if (kaa > 0 && L >= span) {
        const AaFrames F = aa_frames(L, kaa);
        out.assign((L - span + 1) * (translate ? 2 : 1), 0xdeadbeefULL);
        const u32 nst = aa_stage_len(kaa, stride);
        std::vector<u8> sf(nst), sr(nst);
        for (u64 t0 = 0; t0 + span <= L; t0 += AA_TILE) {
            for (u32 i = 0; i < nst; ++i) aa_stage(T, translate, seq.data(), L, t0, i, sf.data(), sr.data());
            for (u32 tid = 0; tid < (u32)AA_TILE; ++tid) {
                const u64 p = t0 + tid;
                if (p + span > L) continue;
                const u64 hf_ = aa_hash_fwd(sf.data(), tid, kaa, stride, 42);
                if (translate) {
                    out[aa_raw_index(F, p, false)] = hf_;
                    out[aa_raw_index(F, L - p - span, true)] = aa_hash_rev(sr.data(), tid, kaa, 42);
                } else {
                    out[p] = hf_;
                }
            }
        }
    }
    f = fopen(argv[5], "wb");
    fwrite(out.data(), 8, out.size(), f);
    fclose(f);
    return 0;
}
