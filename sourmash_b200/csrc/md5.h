// md5.h -- RFC 1321 MD5 for signature identity (KmerMinHash::md5sum,
// src/core/src/sketch/minhash.rs:290-307).  Host only; not on the hot path.
#pragma once
#include <stdint.h>
#include <string.h>

#include <string>

namespace smb {

class Md5 {
  public:
    Md5() : a_(0x67452301u), b_(0xefcdab89u), c_(0x98badcfeu), d_(0x10325476u), bytes_(0), fill_(0) {}

    void update(const uint8_t* p, size_t n) {
        bytes_ += n;
        while (n > 0) {
            size_t take = 64 - fill_;
            if (take > n) take = n;
            memcpy(block_ + fill_, p, take);
            fill_ += take; p += take; n -= take;
            if (fill_ == 64) { transform(block_); fill_ = 0; }
        }
    }

    std::string hexdigest() {
        uint64_t bits = bytes_ * 8;
        static const uint8_t pad[64] = {0x80};
        size_t padlen = (fill_ < 56) ? (56 - fill_) : (120 - fill_);
        update(pad, padlen);
        uint8_t lenb[8];
        for (int i = 0; i < 8; ++i) lenb[i] = (uint8_t)(bits >> (8 * i));
        update(lenb, 8);
        const uint32_t st[4] = {a_, b_, c_, d_};
        static const char* hex = "0123456789abcdef";
        std::string out(32, '0');
        for (int w = 0; w < 4; ++w)
            for (int j = 0; j < 4; ++j) {
                uint8_t byte = (uint8_t)(st[w] >> (8 * j));
                out[8 * w + 2 * j] = hex[byte >> 4];
                out[8 * w + 2 * j + 1] = hex[byte & 15];
            }
        return out;
    }

  private:
    static uint32_t rol(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }

    void transform(const uint8_t* p) {
        static const uint32_t T[64] = {
            0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
            0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
            0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
            0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
            0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
            0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
            0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
            0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
        static const int S[4][4] = {{7, 12, 17, 22}, {5, 9, 14, 20}, {4, 11, 16, 23}, {6, 10, 15, 21}};
        uint32_t x[16];
        for (int i = 0; i < 16; ++i)
            x[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) |
                   ((uint32_t)p[4 * i + 3] << 24);
        uint32_t a = a_, b = b_, c = c_, d = d_;
        for (int i = 0; i < 64; ++i) {
            const int round = i >> 4;
            uint32_t f;
            int g;
            switch (round) {
                case 0: f = (b & c) | (~b & d); g = i; break;
                case 1: f = (b & d) | (c & ~d); g = (5 * i + 1) & 15; break;
                case 2: f = b ^ c ^ d; g = (3 * i + 5) & 15; break;
                default: f = c ^ (b | ~d); g = (7 * i) & 15; break;
            }
            uint32_t tmp = d;
            d = c; c = b;
            b = b + rol(a + f + T[i] + x[g], S[round][i & 3]);
            a = tmp;
        }
        a_ += a; b_ += b; c_ += c; d_ += d;
    }

    uint32_t a_, b_, c_, d_;
    uint64_t bytes_;
    uint8_t block_[64];
    size_t fill_;
};

}  // namespace smb
