// roll_emul.cu -- runs the *device* k-mer rolling/hash code (sourmash_b200/csrc/kmer_roll.cuh,
// compiled here for the host) over a sequence, tiling it exactly like hash_kmers_kernel does.
// Test infrastructure: lets the CPU-only test suite check the kernel's per-thread logic
// against the oracle without a GPU.
//   usage: roll_emul <k> <W> <lead> <infile> <outfile>
// infile holds `lead` junk bytes followed by the sequence; output = u64 per window (0 = invalid).
// k = 0 runs the fused k = 21, 31, 51 pass (hash_thread_windows_fused): the three outputs back to back.
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../sourmash_b200/csrc/kmer_roll.cuh"

using namespace smb;

template <int K>
static void run(const std::vector<u8>& buf, u32 lead, int W, u64 seed, std::vector<u64>& out) {
    const u64 Lp = buf.size();
    const u64 L = Lp - lead;
    out.assign(L >= (u64)K ? L - K + 1 : 0, 0xdeadbeefULL);
    std::vector<u8> padded(buf);
    padded.resize(((Lp + 15) & ~15ULL) + 16, 'A');    // readable up to the next 16-byte line
    const u64 nwin = Lp >= (u64)K ? Lp - K + 1 : 0;
    for (u64 w0 = 0; w0 < nwin; w0 += (u64)W) {
        hash_thread_windows<K>(padded.data(), Lp, lead, w0, W, seed, [&](u64 w, bool valid, u64 h) {
            if (w >= lead && w + K <= Lp) out[w - lead] = valid ? h : 0ULL;
        });
    }
}

static void run_fused(const std::vector<u8>& buf, u32 lead, int W, u64 seed, std::vector<u64>& out) {
    const u64 Lp = buf.size();
    const u64 L = Lp - lead;
    const int ks[3] = {21, 31, 51};
    u64 at[4] = {0, 0, 0, 0};
    for (int i = 0; i < 3; ++i) at[i + 1] = at[i] + (L >= (u64)ks[i] ? L - ks[i] + 1 : 0);
    out.assign(at[3], 0xdeadbeefULL);
    std::vector<u8> padded(buf);
    padded.resize(((Lp + 15) & ~15ULL) + 16, 'A');
    const u64 nwin = Lp >= 21 ? Lp - 21 + 1 : 0;
    for (u64 w0 = 0; w0 < nwin; w0 += (u64)W) {
        hash_thread_windows_fused(padded.data(), Lp, lead, w0, W, seed, [&](u64 w, int which, bool valid, u64 h) {
            if (w >= lead && w + ks[which] <= Lp) {
                u64& slot = out[at[which] + (w - lead)];
                if (slot != 0xdeadbeefULL) { fprintf(stderr, "window emitted twice\n"); exit(5); }
                slot = valid ? h : 0ULL;
            }
        });
    }
}

int main(int argc, char** argv) {
    if (argc != 6) { fprintf(stderr, "usage: roll_emul k W lead in out\n"); return 2; }
    int k = atoi(argv[1]), W = atoi(argv[2]);
    u32 lead = (u32)atoi(argv[3]);
    FILE* f = fopen(argv[4], "rb");
    if (!f) return 3;
    std::vector<u8> buf;
    u8 tmp[65536]; size_t n;
    while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    fclose(f);
    std::vector<u64> out;
    if (k == 0) {
        run_fused(buf, lead, W, 42, out);
        f = fopen(argv[5], "wb");
        fwrite(out.data(), 8, out.size(), f);
        fclose(f);
        return 0;
    }
    switch (k) {
#define CASE(KK) case KK: run<KK>(buf, lead, W, 42, out); break;
        CASE(1) CASE(3) CASE(4) CASE(5) CASE(8) CASE(15) CASE(16) CASE(17) CASE(21) CASE(24) CASE(31) CASE(32)
        CASE(33) CASE(47) CASE(48) CASE(51) CASE(63) CASE(64) CASE(65)
        default: fprintf(stderr, "k not instantiated\n"); return 4;
    }
    f = fopen(argv[5], "wb");
    fwrite(out.data(), 8, out.size(), f);
    fclose(f);
    return 0;
}
