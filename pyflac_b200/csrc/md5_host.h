// md5_host.h -- RFC 1321 MD5, incremental, host side.
//
// FLAC's STREAMINFO carries the MD5 of the unencoded PCM ((bps+7)/8 little-endian bytes per sample,
// interleaved; up: md5.c FLAC__MD5Accumulate, scope row E1).  The hash is a serial chain per stream, so it is
// computed where the PCM lives: by md5_kernel (enc_pack.cu) when the PCM is resident in HBM, and by this code
// on host threads when the caller hands over host memory (the bytes are already there; DESIGN.md "MD5 placement").
#pragma once
#include <stdint.h>
#include <string.h>

namespace fb {

struct Md5 {
    uint32_t h[4]; uint64_t len; uint8_t buf[64]; uint32_t fill;
    void init() { h[0] = 0x67452301u; h[1] = 0xefcdab89u; h[2] = 0x98badcfeu; h[3] = 0x10325476u; len = 0; fill = 0; }
    static uint32_t rol(uint32_t v, int s) { return (v << s) | (v >> (32 - s)); }
    void block(const uint8_t* p) {
        static const uint32_t K[64] = {
            0xd76aa478,0xe8c7b756,0x242070db,0xc1bdceee,0xf57c0faf,0x4787c62a,0xa8304613,0xfd469501,0x698098d8,0x8b44f7af,0xffff5bb1,0x895cd7be,
            0x6b901122,0xfd987193,0xa679438e,0x49b40821,0xf61e2562,0xc040b340,0x265e5a51,0xe9b6c7aa,0xd62f105d,0x02441453,0xd8a1e681,0xe7d3fbc8,
            0x21e1cde6,0xc33707d6,0xf4d50d87,0x455a14ed,0xa9e3e905,0xfcefa3f8,0x676f02d9,0x8d2a4c8a,0xfffa3942,0x8771f681,0x6d9d6122,0xfde5380c,
            0xa4beea44,0x4bdecfa9,0xf6bb4b60,0xbebfbc70,0x289b7ec6,0xeaa127fa,0xd4ef3085,0x04881d05,0xd9d4d039,0xe6db99e5,0x1fa27cf8,0xc4ac5665,
            0xf4292244,0x432aff97,0xab9423a7,0xfc93a039,0x655b59c3,0x8f0ccc92,0xffeff47d,0x85845dd1,0x6fa87e4f,0xfe2ce6e0,0xa3014314,0x4e0811a1,
            0xf7537e82,0xbd3af235,0x2ad7d2bb,0xeb86d391};
        uint32_t w[16], a = h[0], b = h[1], c = h[2], d = h[3];
        memcpy(w, p, 64);                                   // little-endian host
#define FB_R(f, g, s, i) { const uint32_t t = a + (f) + K[i] + w[g]; a = d; d = c; c = b; b = b + rol(t, s); }
        for (int i = 0; i < 16; i += 4) { FB_R((b & c) | (~b & d), i, 7, i) FB_R((b & c) | (~b & d), i + 1, 12, i + 1) FB_R((b & c) | (~b & d), i + 2, 17, i + 2) FB_R((b & c) | (~b & d), i + 3, 22, i + 3) }
        for (int i = 16; i < 32; i += 4) { FB_R((d & b) | (~d & c), (5 * i + 1) & 15, 5, i) FB_R((d & b) | (~d & c), (5 * (i + 1) + 1) & 15, 9, i + 1) FB_R((d & b) | (~d & c), (5 * (i + 2) + 1) & 15, 14, i + 2) FB_R((d & b) | (~d & c), (5 * (i + 3) + 1) & 15, 20, i + 3) }
        for (int i = 32; i < 48; i += 4) { FB_R(b ^ c ^ d, (3 * i + 5) & 15, 4, i) FB_R(b ^ c ^ d, (3 * (i + 1) + 5) & 15, 11, i + 1) FB_R(b ^ c ^ d, (3 * (i + 2) + 5) & 15, 16, i + 2) FB_R(b ^ c ^ d, (3 * (i + 3) + 5) & 15, 23, i + 3) }
        for (int i = 48; i < 64; i += 4) { FB_R(c ^ (b | ~d), (7 * i) & 15, 6, i) FB_R(c ^ (b | ~d), (7 * (i + 1)) & 15, 10, i + 1) FB_R(c ^ (b | ~d), (7 * (i + 2)) & 15, 15, i + 2) FB_R(c ^ (b | ~d), (7 * (i + 3)) & 15, 21, i + 3) }
#undef FB_R
        h[0] += a; h[1] += b; h[2] += c; h[3] += d;
    }
    void update(const uint8_t* p, size_t n) {
        len += n;
        if (fill) {
            size_t k = 64 - fill; if (k > n) k = n;
            memcpy(buf + fill, p, k); fill += (uint32_t)k; p += k; n -= k;
            if (fill == 64) { block(buf); fill = 0; }
        }
        while (n >= 64) { block(p); p += 64; n -= 64; }
        if (n) { memcpy(buf, p, n); fill = (uint32_t)n; }
    }
    void final(uint8_t out[16]) {
        const uint64_t bits = len * 8; uint8_t pad[72]; memset(pad, 0, sizeof pad); pad[0] = 0x80;
        const size_t padlen = (fill < 56) ? (56 - fill) : (120 - fill);
        update(pad, padlen);
        uint8_t lb[8];
        for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (8 * i));
        update(lb, 8);
        for (int i = 0; i < 16; i++) out[i] = (uint8_t)(h[i >> 2] >> (8 * (i & 3)));
    }
    // MD5 of `nvals` samples stored in `cont`-byte little-endian containers, hashing the low `bytes` bytes of each
    void update_samples(const void* pcm, size_t nvals, uint32_t cont, uint32_t bytes) {
        if (cont == bytes) { update((const uint8_t*)pcm, nvals * cont); return; }
        uint8_t tmp[4096 * 3];
        const uint8_t* p = (const uint8_t*)pcm;
        while (nvals) {
            const size_t m = nvals < 4096 ? nvals : 4096;
            size_t k = 0;
            for (size_t i = 0; i < m; i++) for (uint32_t b = 0; b < bytes; b++) tmp[k++] = p[i * cont + b];
            update(tmp, k);
            p += m * cont; nvals -= m;
        }
    }
};

}  // namespace fb
