// kmer_roll.cuh -- rolling canonical k-mer state + murmur3 over register-resident words.
// __host__ __device__ so that tests/host_emul can run exactly this code on the CPU.
#pragma once
#include <string.h>

#include "common.cuh"

namespace smb {

#ifdef __CUDA_ARCH__
__device__ __forceinline__ u32 fshr(u32 lo, u32 hi, u32 s) { return __funnelshift_r(lo, hi, s); }
__device__ __forceinline__ u32 fshl(u32 lo, u32 hi, u32 s) { return __funnelshift_l(lo, hi, s); }
__device__ __forceinline__ u32 pick_byte(u32 table, u32 idx) { return __byte_perm(table, 0u, idx) & 0xffu; }
#else
inline u32 fshr(u32 lo, u32 hi, u32 s) { return (u32)((((u64)hi << 32) | lo) >> (s & 31)); }
inline u32 fshl(u32 lo, u32 hi, u32 s) { return (u32)(((((u64)hi << 32) | lo) << (s & 31)) >> 32); }
inline u32 pick_byte(u32 table, u32 idx) { return (table >> (8 * (idx & 3))) & 0xffu; }
#endif

// ---------------------------------------------------------------------------------------
// murmur3 x64_128 (first word) over K ASCII bytes held as little-endian 32-bit words.
// Bytes >= K in the top word are zero.  Matches oracle/oracle.c orc_hash_murmur.
// ---------------------------------------------------------------------------------------
template <int K>
__host__ __device__ __forceinline__ u64 murmur_words(const u32 (&w)[(K + 3) / 4], u64 seed) {
    constexpr int N32 = (K + 3) / 4;
    constexpr int NBLK = K / 16;
    constexpr int TAIL = K % 16;
    u64 h1 = seed, h2 = seed;
    auto word64 = [&](int i) -> u64 {          // i-th little-endian u64 of the k-mer (zero padded)
        u32 lo = (2 * i < N32) ? w[2 * i] : 0u;
        u32 hi = (2 * i + 1 < N32) ? w[2 * i + 1] : 0u;
        return ((u64)hi << 32) | lo;
    };
#pragma unroll
    for (int b = 0; b < NBLK; ++b) {
        u64 k1 = word64(2 * b), k2 = word64(2 * b + 1);
        k1 *= SMB_C1; k1 = smb_rotl64(k1, 31); k1 *= SMB_C2; h1 ^= k1;
        h1 = smb_rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729ULL;
        k2 *= SMB_C2; k2 = smb_rotl64(k2, 33); k2 *= SMB_C1; h2 ^= k2;
        h2 = smb_rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5ULL;
    }
    if (TAIL > 8) {
        u64 k2 = word64(2 * NBLK + 1);
        k2 *= SMB_C2; k2 = smb_rotl64(k2, 33); k2 *= SMB_C1; h2 ^= k2;
    }
    if (TAIL > 0) {
        u64 k1 = word64(2 * NBLK);
        k1 *= SMB_C1; k1 = smb_rotl64(k1, 31); k1 *= SMB_C2; h1 ^= k1;
    }
    h1 ^= (u64)K; h2 ^= (u64)K;
    h1 += h2; h2 += h1;
    h1 = smb_fmix64(h1); h2 = smb_fmix64(h2);
    return h1 + h2;
}

// ---------------------------------------------------------------------------------------
// Rolling k-mer state for compile-time K.
//   fw[]  forward k-mer, byte t of the k-mer at byte t (little endian words)
//   rc[]  reverse complement, same layout
//   cf/cr 2-bit codes (A0 C1 G2 T3), first base most significant -> integer order ==
//         byte-lexicographic order of the ASCII strings (signature.rs:304 std::cmp::min)
// ---------------------------------------------------------------------------------------
template <int K>
struct Roll {
    static constexpr int N32 = (K + 3) / 4;
    static constexpr int NC = (2 * K + 31) / 32;       // 32-bit words of 2-bit codes
    u32 fw[N32], rc[N32];
    u32 cf[NC], cr[NC];
    u32 run;                                            // consecutive valid bases so far

    __host__ __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < N32; ++i) { fw[i] = 0; rc[i] = 0; }
#pragma unroll
        for (int i = 0; i < NC; ++i) { cf[i] = 0; cr[i] = 0; }
        run = 0;
    }

    // push one raw input byte; returns whether it is a valid base
    __host__ __device__ __forceinline__ bool push(u32 x) {
        const u32 up = x & 0xDFu;                        // to_ascii_uppercase for letters
        const u32 c2 = (up >> 1) & 3u;                   // A0 C1 T2 G3
        const u32 expect = pick_byte(0x47544341u, c2);   // "ACTG"[c2]
        const bool ok = (expect == up);
        const u32 comp = pick_byte(0x43414754u, c2);     // complement: "TGAC"[c2]
        const u32 code = c2 ^ (c2 >> 1);                 // A0 C1 G2 T3
        run = ok ? run + 1u : 0u;
        // forward: drop byte 0, append `up` at byte K-1
#pragma unroll
        for (int i = 0; i < N32 - 1; ++i) fw[i] = fshr(fw[i], fw[i + 1], 8);
        fw[N32 - 1] >>= 8;
        fw[(K - 1) / 4] |= up << (8 * ((K - 1) % 4));
        // reverse complement: prepend `comp` at byte 0, drop byte K
#pragma unroll
        for (int i = N32 - 1; i > 0; --i) rc[i] = fshl(rc[i - 1], rc[i], 8);
        rc[0] = (rc[0] << 8) | comp;
        if (K % 4 != 0) rc[N32 - 1] &= (1u << (8 * (K % 4))) - 1u;
        // 2-bit forward: shift left by 2, insert code at the bottom, keep 2K bits
#pragma unroll
        for (int i = NC - 1; i > 0; --i) cf[i] = fshl(cf[i - 1], cf[i], 2);
        cf[0] = (cf[0] << 2) | code;
        if ((2 * K) % 32 != 0) cf[NC - 1] &= (1u << ((2 * K) % 32)) - 1u;
        // 2-bit revcomp: shift right by 2, insert (3-code) at the top (bit 2K-2)
#pragma unroll
        for (int i = 0; i < NC - 1; ++i) cr[i] = fshr(cr[i], cr[i + 1], 2);
        cr[NC - 1] >>= 2;
        cr[(2 * K - 2) / 32] |= (code ^ 3u) << ((2 * K - 2) % 32);
        return ok;
    }

    __host__ __device__ __forceinline__ bool fwd_is_canonical() const {
        // multiword compare, most significant word first; tie -> forward (identical strings)
        bool lt = false, decided = false;
#pragma unroll
        for (int i = NC - 1; i >= 0; --i) {
            if (!decided && cf[i] != cr[i]) { lt = cf[i] < cr[i]; decided = true; }
        }
        return decided ? lt : true;
    }

    __host__ __device__ __forceinline__ u64 hash(u64 seed) const {
        const bool f = fwd_is_canonical();
        u32 sel[N32];
#pragma unroll
        for (int i = 0; i < N32; ++i) sel[i] = f ? fw[i] : rc[i];
        return murmur_words<K>(sel, seed);
    }
};


struct Bytes16 { u32 w[4]; };
#ifdef __CUDA_ARCH__
__device__ __forceinline__ Bytes16 load16(const u8* p) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    Bytes16 b; b.w[0] = v.x; b.w[1] = v.y; b.w[2] = v.z; b.w[3] = v.w; return b;
}
#else
inline Bytes16 load16(const u8* p) { Bytes16 b; memcpy(b.w, p, 16); return b; }
#endif

// Sequential 32-bit word reader over a thread's byte range, in "aligned coordinates":
//   base   16-byte aligned pointer to the line holding the stream's first byte
//   lead   bytes of that line that precede the stream; Lp = lead + stream length
// Bytes outside [lead, Lp) read as 0 (an invalid base), so windows that touch them never emit.
struct WordStream {
    const u8* base;
    u64 Lp, pos;
    u32 lead, r0, r1, r2, r3, i;
    __host__ __device__ __forceinline__ WordStream(const u8* b, u64 lp, u32 ld, u64 start)
        : base(b), Lp(lp), pos(start), lead(ld), r0(0), r1(0), r2(0), r3(0), i(0) {}
    __host__ __device__ __forceinline__ void refill() {
        r0 = r1 = r2 = r3 = 0u;
        if (pos < Lp) {
            Bytes16 v = load16(base + pos);
            r0 = v.w[0]; r1 = v.w[1]; r2 = v.w[2]; r3 = v.w[3];
            const u64 rem = Lp - pos;
            if (rem < 16 || pos < lead) {              // partial line: zero the bytes outside the stream
                const u32 hi = rem < 16 ? (u32)rem : 16u;
                u32 lo = 0;
                if (pos < lead) { const u64 d = lead - pos; lo = d < 16 ? (u32)d : 16u; }
                const u32 vm = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
                u32* r[4] = {&r0, &r1, &r2, &r3};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    u32 m = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b) if ((vm >> (4 * q + b)) & 1u) m |= 0xffu << (8 * b);
                    *r[q] &= m;
                }
            }
        }
        pos += 16;
    }
    __host__ __device__ __forceinline__ u32 next() {
        if ((i & 3u) == 0u) refill();
        ++i;
        const u32 w = r0;
        r0 = r1; r1 = r2; r2 = r3;
        return w;
    }
};

// One thread's share of a stream: windows [w0, w0 + W), W a multiple of 4.  emit(w, valid, h)
// is called for each of them in order (w in aligned coordinates, i.e. including `lead`; windows
// past the end of the stream come out with valid == false).
template <int K, class Emit>
__host__ __device__ __forceinline__ void hash_thread_windows(const u8* __restrict__ base, u64 Lp, u32 lead,
                                                             u64 w0, int W, u64 seed, Emit&& emit) {
    const u64 nwin = Lp >= (u64)K ? Lp - K + 1 : 0;
    if (w0 >= nwin) return;
    constexpr int Q = (K - 1) / 4, R = (K - 1) % 4;
    WordStream ws(base, Lp, lead, w0);
    Roll<K> st;
    st.init();
    // warm-up: the first K-1 bases only fill the rolling state
#pragma unroll 1
    for (int q = 0; q < Q; ++q) {
        const u32 word = ws.next();
#pragma unroll
        for (int b = 0; b < 4; ++b) st.push((word >> (8 * b)) & 0xffu);
    }
    u32 cur = ws.next();
#pragma unroll
    for (int b = 0; b < R; ++b) st.push((cur >> (8 * b)) & 0xffu);
    // main phase: one window per base; bases come 4 at a time, re-aligned by a funnel shift
    u64 w = w0;
#pragma unroll 1
    for (int q = 0; q < (W >> 2); ++q) {
        const u32 nxt = ws.next();
        const u32 word = R ? fshr(cur, nxt, 8 * R) : cur;
        cur = nxt;
#pragma unroll
        for (int b = 0; b < 4; ++b, ++w) {
            st.push((word >> (8 * b)) & 0xffu);
            emit(w, st.run >= (u32)K, st.hash(seed));
        }
    }
}

// ---------------------------------------------------------------------------------------
// k = 21, 31 and 51 in one pass (experimental, SMB_SKETCH_FUSED; sketch_kernels.cu).  The three
// k-mers that START at the same base are prefixes of the 51-mer, so one rolling 51-state serves
// all of them:
//   forward   bytes 0..k-1 of fw51: whole words plus a masked last word (21 = 5 x 4 + 1, 31 = 7 x 4 + 3)
//   revcomp   the last k bytes of rc51: bytes 20..50 for k = 31 (word aligned), bytes 30..50 for
//             k = 21 (a 16-bit funnel shift per word)
//   order     2-bit codes: forward = the top 2k bits of cf51, revcomp = the low 2k bits of cr51
//   validity  a 64-bit shift register of "valid base" flags; the k-prefix needs bits 50 .. 51-k
// emit(w, which, valid, h): which = 0, 1, 2 for k = 21, 31, 51; windows in the order of the start w.
// ---------------------------------------------------------------------------------------
template <class Emit>
__host__ __device__ __forceinline__ void hash_thread_windows_fused(const u8* __restrict__ base, u64 Lp, u32 lead,
                                                                   u64 w0, int W, u64 seed, Emit&& emit) {
    constexpr int K = 51;
    const u64 nwin = Lp >= 21 ? Lp - 21 + 1 : 0;              // starts that have at least a 21-mer
    if (w0 >= nwin) return;
    constexpr int Q = (K - 1) / 4, R = (K - 1) % 4;
    WordStream ws(base, Lp, lead, w0);
    Roll<K> st;
    st.init();
    u64 vmask = 0;                                            // bit i: the base i positions before the newest is valid
#pragma unroll 1
    for (int q = 0; q < Q; ++q) {
        const u32 word = ws.next();
#pragma unroll
        for (int b = 0; b < 4; ++b) vmask = (vmask << 1) | (st.push((word >> (8 * b)) & 0xffu) ? 1ull : 0ull);
    }
    u32 cur = ws.next();
#pragma unroll
    for (int b = 0; b < R; ++b) vmask = (vmask << 1) | (st.push((cur >> (8 * b)) & 0xffu) ? 1ull : 0ull);
    constexpr u64 M51 = (1ull << 51) - 1, M31 = M51 & ~((1ull << 20) - 1), M21 = M51 & ~((1ull << 30) - 1);
    u64 w = w0;
#pragma unroll 1
    for (int q = 0; q < (W >> 2); ++q) {
        const u32 nxt = ws.next();
        const u32 word = R ? fshr(cur, nxt, 8 * R) : cur;
        cur = nxt;
#pragma unroll 1
        for (int b = 0; b < 4; ++b, ++w) {
            vmask = (vmask << 1) | (st.push((word >> (8 * b)) & 0xffu) ? 1ull : 0ull);
            // k = 21
            {
                const u32 f0 = fshr(st.cf[1], st.cf[2], 28), f1 = fshr(st.cf[2], st.cf[3], 28);     // cf51 >> 60
                const u32 r0 = st.cr[0], r1 = st.cr[1] & 0x3ffu;                                   // low 42 bits of cr51
                const bool fwd = f1 != r1 ? f1 < r1 : f0 <= r0;
                u32 sel[6];
#pragma unroll
                for (int i = 0; i < 5; ++i) sel[i] = fwd ? st.fw[i] : fshr(st.rc[7 + i], st.rc[8 + i], 16);
                sel[5] = fwd ? (st.fw[5] & 0xffu) : (st.rc[12] >> 16);
                emit(w, 0, (vmask & M21) == M21, murmur_words<21>(sel, seed));
            }
            // k = 31
            {
                const u32 f0 = fshr(st.cf[1], st.cf[2], 8), f1 = fshr(st.cf[2], st.cf[3], 8);       // cf51 >> 40
                const u32 r0 = st.cr[0], r1 = st.cr[1] & 0x3fffffffu;                              // low 62 bits of cr51
                const bool fwd = f1 != r1 ? f1 < r1 : f0 <= r0;
                u32 sel[8];
#pragma unroll
                for (int i = 0; i < 7; ++i) sel[i] = fwd ? st.fw[i] : st.rc[5 + i];
                sel[7] = fwd ? (st.fw[7] & 0xffffffu) : st.rc[12];
                emit(w, 1, (vmask & M31) == M31, murmur_words<31>(sel, seed));
            }
            emit(w, 2, (vmask & M51) == M51, st.hash(seed));
        }
    }
}

}  // namespace smb
