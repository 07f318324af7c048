// aa_kmers.cuh -- protein-family k-mer hashing (protein / dayhoff / hp sketches), either from
// residues (add_protein) or from DNA by six-frame translation (add_sequence on such a sketch).
// __host__ __device__ so that tests/host_emul can run exactly this code on the CPU.
//
// Replaces   SeqToHashes::next, translate + protein branches   src/core/src/signature.rs:307-392
//            translate_codon / aa_to_dayhoff / aa_to_hp / to_aa  src/core/src/encodings.rs:103-365
//
// Observation that shapes the kernel: the reference walks three forward frames and three frames
// of the reverse complement, each cut into codons.  Frame f, residue window j covers DNA
// [f + 3j, f + 3j + 3k) -- so *every* DNA window of 3k bases is hashed exactly twice: once
// translated as it stands and once translated from its reverse complement.  The six-frame walk is
// therefore one pass over DNA positions p with two hashes per position, and the residue of the
// codon starting at p (forward: bases p,p+1,p+2; reverse: comp(p+2),comp(p+1),comp(p)) is computed
// once per position into shared memory and reused by the k windows that contain it.
#pragma once
#include <string.h>

#include "common.cuh"

namespace smb {

static constexpr int AA_TILE = 256;          // window start positions per CTA (= tile_start_generic)

// Tables built on the host from the sketch's hash function (capi.cu build_aa_tables):
struct AaTables {
    u8 base_code[256];    // raw byte -> A0 C1 G2 T3 N4 other5 (lower case folded, signature.rs:214)
    u8 codon[216];        // [c0*36 + c1*6 + c2] -> residue, already dayhoff/hp re-encoded
    u8 residue[256];      // raw residue byte -> upper-cased and re-encoded residue (protein input)
};

// complement on codes: A<->T, C<->G, N->N; anything else has no complement (COMPLEMENT -> 0,
// encodings.rs:85-95) and stays "other"
__host__ __device__ __forceinline__ u32 aa_comp_code(u32 c) { return c < 4u ? 3u - c : c; }

// residues of the forward and reverse-complement codons that start at DNA position `pos`
// (undefined windows -- fewer than 3 bases left -- are never read by a valid window)
__host__ __device__ __forceinline__ void aa_translate_at(const AaTables& T, const u8* __restrict__ seq, u64 L,
                                                         u64 pos, u8& fwd, u8& rev) {
    const u32 c0 = pos < L ? T.base_code[seq[pos]] : 5u;
    const u32 c1 = pos + 1 < L ? T.base_code[seq[pos + 1]] : 5u;
    const u32 c2 = pos + 2 < L ? T.base_code[seq[pos + 2]] : 5u;
    fwd = T.codon[c0 * 36u + c1 * 6u + c2];
    rev = T.codon[aa_comp_code(c2) * 36u + aa_comp_code(c1) * 6u + aa_comp_code(c0)];
}

// murmur3 x64_128 (first word) of K bytes delivered by at(0..K-1); same arithmetic as
// kmer_roll.cuh murmur_words / oracle orc_hash_murmur
template <class At>
__host__ __device__ __forceinline__ u64 murmur_at(u32 K, u64 seed, At&& at) {
    u64 h1 = seed, h2 = seed;
    const u32 nblk = K / 16;
    for (u32 b = 0; b < nblk; ++b) {
        u64 k1 = 0, k2 = 0;
        for (int i = 7; i >= 0; --i) { k1 = (k1 << 8) | (u64)at(16 * b + i); k2 = (k2 << 8) | (u64)at(16 * b + 8 + i); }
        k1 *= SMB_C1; k1 = smb_rotl64(k1, 31); k1 *= SMB_C2; h1 ^= k1;
        h1 = smb_rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729ULL;
        k2 *= SMB_C2; k2 = smb_rotl64(k2, 33); k2 *= SMB_C1; h2 ^= k2;
        h2 = smb_rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5ULL;
    }
    const u32 tail = K & 15u, tb = 16 * nblk;
    if (tail > 8) {
        u64 k2 = 0;
        for (u32 i = tail; i > 8; --i) k2 = (k2 << 8) | (u64)at(tb + i - 1);
        k2 *= SMB_C2; k2 = smb_rotl64(k2, 33); k2 *= SMB_C1; h2 ^= k2;
    }
    if (tail > 0) {
        u64 k1 = 0;
        for (u32 i = tail < 8 ? tail : 8; i > 0; --i) k1 = (k1 << 8) | (u64)at(tb + i - 1);
        k1 *= SMB_C1; k1 = smb_rotl64(k1, 31); k1 *= SMB_C2; h1 ^= k1;
    }
    h1 ^= (u64)K; h2 ^= (u64)K;
    h1 += h2; h2 += h1;
    h1 = smb_fmix64(h1); h2 = smb_fmix64(h2);
    return h1 + h2;
}

// Number of staged positions a tile needs: AA_TILE window starts, each reading kaa residues at
// `stride` (3 for translated DNA, 1 for residues).
__host__ __device__ __forceinline__ u32 aa_stage_len(u32 kaa, u32 stride) { return AA_TILE + (kaa - 1) * stride; }

// Stage one position of a tile: sf/sr are the tile's shared-memory arrays, `i` the local index.
__host__ __device__ __forceinline__ void aa_stage(const AaTables& T, bool translate, const u8* __restrict__ seq,
                                                  u64 L, u64 t0, u32 i, u8* sf, u8* sr) {
    const u64 pos = t0 + i;
    if (translate) {
        u8 f, r;
        aa_translate_at(T, seq, L, pos, f, r);
        sf[i] = f; sr[i] = r;
    } else {
        sf[i] = pos < L ? T.residue[seq[pos]] : (u8)0;
    }
}

// hashes of the window that starts at local index `i` of a staged tile
__host__ __device__ __forceinline__ u64 aa_hash_fwd(const u8* sf, u32 i, u32 kaa, u32 stride, u64 seed) {
    return murmur_at(kaa, seed, [&](u32 t) -> u32 { return sf[i + t * stride]; });
}
// reverse-complement strand: its residues run backwards over the forward coordinates
__host__ __device__ __forceinline__ u64 aa_hash_rev(const u8* sr, u32 i, u32 kaa, u64 seed) {
    return murmur_at(kaa, seed, [&](u32 t) -> u32 { return sr[i + (kaa - 1u - t) * 3u]; });
}

// seq_to_hashes order for translated input (signature.rs:310-345): frame 0 forward, frame 0
// reverse, frame 1 forward, ...  nwin[f] = residue windows of frame f (same for both strands).
struct AaFrames { u64 nwin[3]; };
__host__ __device__ __forceinline__ AaFrames aa_frames(u64 L, u32 kaa) {
    AaFrames F;
    for (u32 f = 0; f < 3; ++f) {
        const u64 na = L >= f ? (L - f) / 3 : 0;
        F.nwin[f] = na >= kaa ? na - kaa + 1 : 0;
    }
    return F;
}
__host__ __device__ __forceinline__ u64 aa_raw_index(const AaFrames& F, u64 q, bool reverse) {
    // q = start of the window in the strand's own coordinates (forward: p; reverse: L - p - 3k)
    const u32 f = (u32)(q % 3);
    u64 base = 0;
    for (u32 g = 0; g < f; ++g) base += 2 * F.nwin[g];
    return base + (reverse ? F.nwin[f] : 0) + q / 3;
}

// ---------------------------------------------------------------------------------------
// hf: 2 protein, 3 dayhoff, 4 hp (include/sourmash.h:11-17)
// Host side: tables built from the reference's maps:
// CODONTABLE / DAYHOFFTABLE / HPTABLE, src/core/src/encodings.rs:103-296.
inline uint8_t aa_reencode(uint8_t aa, int hf) {
    auto in = [&](const char* set) { return aa != 0 && strchr(set, (int)aa) != nullptr; };
    if (hf == 3) {          // encodings.rs:218-252, 328-333
        if (aa == 'C') return 'a';
        if (in("AGPST")) return 'b';
        if (in("DENQ")) return 'c';
        if (in("HKR")) return 'd';
        if (in("ILMV")) return 'e';
        if (in("FWY")) return 'f';
        return aa == '*' ? '*' : 'X';
    }
    if (hf == 4) {               // encodings.rs:262-296, 335-340
        if (in("AFGILMPVWY")) return 'h';
        if (in("NCSTDERHKQ")) return 'p';
        return aa == '*' ? '*' : 'X';
    }
    return aa;
}
// standard genetic code for codes A0 C1 G2 T3; index c0*16 + c1*4 + c2 (encodings.rs:103-201)
static const char kGeneticCode[65] =
    "KNKNTTTTRSRSIIMI" "QHQHPPPPRRRRLLLL" "EDEDAAAAGGGGVVVV" "*Y*YSSSS*CWCLFLF";
inline uint8_t codon_residue(uint32_t c0, uint32_t c1, uint32_t c2) {
    if (c0 > 3 || c1 > 3 || c2 > 4) return 'X';          // not in CODONTABLE
    if (c2 < 4) return (uint8_t)kGeneticCode[c0 * 16 + c1 * 4 + c2];
    // "xyN": only the four-fold degenerate families have an entry (TCN CTN CCN CGN ACN GTN GCN GGN)
    const char r = kGeneticCode[c0 * 16 + c1 * 4];
    for (uint32_t t = 1; t < 4; ++t) if (kGeneticCode[c0 * 16 + c1 * 4 + t] != r) return 'X';
    return (uint8_t)r;
}
inline AaTables build_aa_tables(int hf) {
    AaTables T;
    for (int b = 0; b < 256; ++b) {
        int u = (b >= 'a' && b <= 'z') ? b - 32 : b;       // signature.rs:214
        T.base_code[b] = u == 'A' ? 0 : u == 'C' ? 1 : u == 'G' ? 2 : u == 'T' ? 3 : u == 'N' ? 4 : 5;
        T.residue[b] = aa_reencode((uint8_t)u, hf);
    }
    for (uint32_t c0 = 0; c0 < 6; ++c0)
        for (uint32_t c1 = 0; c1 < 6; ++c1)
            for (uint32_t c2 = 0; c2 < 6; ++c2)
                T.codon[c0 * 36 + c1 * 6 + c2] = aa_reencode(codon_residue(c0, c1, c2), hf);
    return T;
}

}  // namespace smb
