// md5_mb.h -- multi-buffer MD5 on the host: 8 independent streams per AVX2 register lane, 16 per AVX-512 register.
//
// One MD5 is a serial chain, but a batch holds hundreds of independent streams, so the host side hashes eight of
// them at a time in SIMD lanes (about 8x the per-core throughput of the scalar chain).  Used by the host->host
// encode path (engine.cu) when the container bytes are the hashed bytes (int16/16-bit, int32/32-bit);
// everything else, and CPUs without AVX2, use the scalar fb::Md5 (md5_host.h).
#pragma once
#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>

#include "md5_host.h"

namespace fb {

inline bool md5_mb_available() { return __builtin_cpu_supports("avx2"); }

#define FB_ROL(v, s) _mm256_or_si256(_mm256_slli_epi32((v), (s)), _mm256_srli_epi32((v), 32 - (s)))
#define FB_STEP(f, wv, k, s) { __m256i t = _mm256_add_epi32(_mm256_add_epi32(a, (f)), _mm256_add_epi32((wv), _mm256_set1_epi32((int)(k)))); \
                               a = d; d = c; c = b; b = _mm256_add_epi32(b, FB_ROL(t, s)); }
#define FB_F1 _mm256_xor_si256(d, _mm256_and_si256(b, _mm256_xor_si256(c, d)))
#define FB_F2 _mm256_xor_si256(c, _mm256_and_si256(d, _mm256_xor_si256(b, c)))
#define FB_F3 _mm256_xor_si256(b, _mm256_xor_si256(c, d))
#define FB_F4 _mm256_xor_si256(c, _mm256_or_si256(b, _mm256_xor_si256(d, ones)))

__attribute__((target("avx2"))) static inline void md5_mb_transpose8(__m256i r[8]) {
    const __m256i t0 = _mm256_unpacklo_epi32(r[0], r[1]), t1 = _mm256_unpackhi_epi32(r[0], r[1]);
    const __m256i t2 = _mm256_unpacklo_epi32(r[2], r[3]), t3 = _mm256_unpackhi_epi32(r[2], r[3]);
    const __m256i t4 = _mm256_unpacklo_epi32(r[4], r[5]), t5 = _mm256_unpackhi_epi32(r[4], r[5]);
    const __m256i t6 = _mm256_unpacklo_epi32(r[6], r[7]), t7 = _mm256_unpackhi_epi32(r[6], r[7]);
    const __m256i u0 = _mm256_unpacklo_epi64(t0, t2), u1 = _mm256_unpackhi_epi64(t0, t2);
    const __m256i u2 = _mm256_unpacklo_epi64(t1, t3), u3 = _mm256_unpackhi_epi64(t1, t3);
    const __m256i u4 = _mm256_unpacklo_epi64(t4, t6), u5 = _mm256_unpackhi_epi64(t4, t6);
    const __m256i u6 = _mm256_unpacklo_epi64(t5, t7), u7 = _mm256_unpackhi_epi64(t5, t7);
    r[0] = _mm256_permute2x128_si256(u0, u4, 0x20); r[4] = _mm256_permute2x128_si256(u0, u4, 0x31);
    r[1] = _mm256_permute2x128_si256(u1, u5, 0x20); r[5] = _mm256_permute2x128_si256(u1, u5, 0x31);
    r[2] = _mm256_permute2x128_si256(u2, u6, 0x20); r[6] = _mm256_permute2x128_si256(u2, u6, 0x31);
    r[3] = _mm256_permute2x128_si256(u3, u7, 0x20); r[7] = _mm256_permute2x128_si256(u3, u7, 0x31);
}

// Advance 8 MD5 states (h[lane][0..3]) over `nblocks` 64-byte blocks read from ptr[lane].
__attribute__((target("avx2"))) static inline void md5_x8_avx2(uint32_t h[8][4], const uint8_t* const ptr[8], size_t nblocks) {
    static const uint32_t K[64] = {
        0xd76aa478,0xe8c7b756,0x242070db,0xc1bdceee,0xf57c0faf,0x4787c62a,0xa8304613,0xfd469501,0x698098d8,0x8b44f7af,0xffff5bb1,0x895cd7be,
        0x6b901122,0xfd987193,0xa679438e,0x49b40821,0xf61e2562,0xc040b340,0x265e5a51,0xe9b6c7aa,0xd62f105d,0x02441453,0xd8a1e681,0xe7d3fbc8,
        0x21e1cde6,0xc33707d6,0xf4d50d87,0x455a14ed,0xa9e3e905,0xfcefa3f8,0x676f02d9,0x8d2a4c8a,0xfffa3942,0x8771f681,0x6d9d6122,0xfde5380c,
        0xa4beea44,0x4bdecfa9,0xf6bb4b60,0xbebfbc70,0x289b7ec6,0xeaa127fa,0xd4ef3085,0x04881d05,0xd9d4d039,0xe6db99e5,0x1fa27cf8,0xc4ac5665,
        0xf4292244,0x432aff97,0xab9423a7,0xfc93a039,0x655b59c3,0x8f0ccc92,0xffeff47d,0x85845dd1,0x6fa87e4f,0xfe2ce6e0,0xa3014314,0x4e0811a1,
        0xf7537e82,0xbd3af235,0x2ad7d2bb,0xeb86d391};
    const __m256i ones = _mm256_set1_epi32(-1);
    __m256i st[8];
    for (int l = 0; l < 8; l++) st[l] = _mm256_set_epi32(0, 0, 0, 0, (int)h[l][3], (int)h[l][2], (int)h[l][1], (int)h[l][0]);
    md5_mb_transpose8(st);                         // st[0..3] = a, b, c, d across the 8 lanes
    __m256i A = st[0], B = st[1], Cc = st[2], D = st[3];
    for (size_t blk = 0; blk < nblocks; blk++) {
        __m256i w[16];
        for (int l = 0; l < 8; l++) {
            w[l] = _mm256_loadu_si256((const __m256i*)(ptr[l] + 64 * blk));
            w[8 + l] = _mm256_loadu_si256((const __m256i*)(ptr[l] + 64 * blk + 32));
        }
        md5_mb_transpose8(w);
        md5_mb_transpose8(w + 8);
        __m256i a = A, b = B, c = Cc, d = D;
        FB_STEP(FB_F1, w[0], K[0], 7) FB_STEP(FB_F1, w[1], K[1], 12) FB_STEP(FB_F1, w[2], K[2], 17) FB_STEP(FB_F1, w[3], K[3], 22)
        FB_STEP(FB_F1, w[4], K[4], 7) FB_STEP(FB_F1, w[5], K[5], 12) FB_STEP(FB_F1, w[6], K[6], 17) FB_STEP(FB_F1, w[7], K[7], 22)
        FB_STEP(FB_F1, w[8], K[8], 7) FB_STEP(FB_F1, w[9], K[9], 12) FB_STEP(FB_F1, w[10], K[10], 17) FB_STEP(FB_F1, w[11], K[11], 22)
        FB_STEP(FB_F1, w[12], K[12], 7) FB_STEP(FB_F1, w[13], K[13], 12) FB_STEP(FB_F1, w[14], K[14], 17) FB_STEP(FB_F1, w[15], K[15], 22)
        FB_STEP(FB_F2, w[1], K[16], 5) FB_STEP(FB_F2, w[6], K[17], 9) FB_STEP(FB_F2, w[11], K[18], 14) FB_STEP(FB_F2, w[0], K[19], 20)
        FB_STEP(FB_F2, w[5], K[20], 5) FB_STEP(FB_F2, w[10], K[21], 9) FB_STEP(FB_F2, w[15], K[22], 14) FB_STEP(FB_F2, w[4], K[23], 20)
        FB_STEP(FB_F2, w[9], K[24], 5) FB_STEP(FB_F2, w[14], K[25], 9) FB_STEP(FB_F2, w[3], K[26], 14) FB_STEP(FB_F2, w[8], K[27], 20)
        FB_STEP(FB_F2, w[13], K[28], 5) FB_STEP(FB_F2, w[2], K[29], 9) FB_STEP(FB_F2, w[7], K[30], 14) FB_STEP(FB_F2, w[12], K[31], 20)
        FB_STEP(FB_F3, w[5], K[32], 4) FB_STEP(FB_F3, w[8], K[33], 11) FB_STEP(FB_F3, w[11], K[34], 16) FB_STEP(FB_F3, w[14], K[35], 23)
        FB_STEP(FB_F3, w[1], K[36], 4) FB_STEP(FB_F3, w[4], K[37], 11) FB_STEP(FB_F3, w[7], K[38], 16) FB_STEP(FB_F3, w[10], K[39], 23)
        FB_STEP(FB_F3, w[13], K[40], 4) FB_STEP(FB_F3, w[0], K[41], 11) FB_STEP(FB_F3, w[3], K[42], 16) FB_STEP(FB_F3, w[6], K[43], 23)
        FB_STEP(FB_F3, w[9], K[44], 4) FB_STEP(FB_F3, w[12], K[45], 11) FB_STEP(FB_F3, w[15], K[46], 16) FB_STEP(FB_F3, w[2], K[47], 23)
        FB_STEP(FB_F4, w[0], K[48], 6) FB_STEP(FB_F4, w[7], K[49], 10) FB_STEP(FB_F4, w[14], K[50], 15) FB_STEP(FB_F4, w[5], K[51], 21)
        FB_STEP(FB_F4, w[12], K[52], 6) FB_STEP(FB_F4, w[3], K[53], 10) FB_STEP(FB_F4, w[10], K[54], 15) FB_STEP(FB_F4, w[1], K[55], 21)
        FB_STEP(FB_F4, w[8], K[56], 6) FB_STEP(FB_F4, w[15], K[57], 10) FB_STEP(FB_F4, w[6], K[58], 15) FB_STEP(FB_F4, w[13], K[59], 21)
        FB_STEP(FB_F4, w[4], K[60], 6) FB_STEP(FB_F4, w[11], K[61], 10) FB_STEP(FB_F4, w[2], K[62], 15) FB_STEP(FB_F4, w[9], K[63], 21)
        A = _mm256_add_epi32(A, a); B = _mm256_add_epi32(B, b); Cc = _mm256_add_epi32(Cc, c); D = _mm256_add_epi32(D, d);
    }
    st[0] = A; st[1] = B; st[2] = Cc; st[3] = D;
    st[4] = st[5] = st[6] = st[7] = _mm256_setzero_si256();
    md5_mb_transpose8(st);
    for (int l = 0; l < 8; l++) { alignas(32) uint32_t tmp[8]; _mm256_store_si256((__m256i*)tmp, st[l]); for (int i = 0; i < 4; i++) h[l][i] = tmp[i]; }
}
#undef FB_ROL
#undef FB_STEP
#undef FB_F1
#undef FB_F2
#undef FB_F3
#undef FB_F4

// ---- AVX-512: 16 streams per register, native rotate (vprold) and 3-input logic (vpternlogd) ----
inline bool md5_mb16_available() { return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw"); }

__attribute__((target("avx512f"))) static inline void md5_mb_transpose16(const __m512i r[16], __m512i w[16]) {
    __m512i t[16], u[16];
    for (int k = 0; k < 8; k++) { t[2 * k] = _mm512_unpacklo_epi32(r[2 * k], r[2 * k + 1]); t[2 * k + 1] = _mm512_unpackhi_epi32(r[2 * k], r[2 * k + 1]); }
    for (int k = 0; k < 4; k++) {
        u[4 * k + 0] = _mm512_unpacklo_epi64(t[4 * k], t[4 * k + 2]); u[4 * k + 1] = _mm512_unpackhi_epi64(t[4 * k], t[4 * k + 2]);
        u[4 * k + 2] = _mm512_unpacklo_epi64(t[4 * k + 1], t[4 * k + 3]); u[4 * k + 3] = _mm512_unpackhi_epi64(t[4 * k + 1], t[4 * k + 3]);
    }
    // u[4k+m], 128-bit lane L: word 4L+m of rows 4k..4k+3
    for (int m = 0; m < 4; m++) {
        const __m512i a = _mm512_shuffle_i32x4(u[m], u[4 + m], 0x88), b = _mm512_shuffle_i32x4(u[m], u[4 + m], 0xdd);
        const __m512i c = _mm512_shuffle_i32x4(u[8 + m], u[12 + m], 0x88), d = _mm512_shuffle_i32x4(u[8 + m], u[12 + m], 0xdd);
        w[m] = _mm512_shuffle_i32x4(a, c, 0x88); w[8 + m] = _mm512_shuffle_i32x4(a, c, 0xdd);
        w[4 + m] = _mm512_shuffle_i32x4(b, d, 0x88); w[12 + m] = _mm512_shuffle_i32x4(b, d, 0xdd);
    }
}

#define FB_STEP16(imm, wv, k, s) { __m512i t = _mm512_add_epi32(_mm512_add_epi32(a, _mm512_ternarylogic_epi32(b, c, d, imm)), \
                                                              _mm512_add_epi32((wv), _mm512_set1_epi32((int)(k)))); \
                                   a = d; d = c; c = b; b = _mm512_add_epi32(b, _mm512_rol_epi32(t, s)); }
// truth tables over (b, c, d): F = (b&c)|(~b&d) = 0xCA, G = (d&b)|(~d&c) = 0xE4, H = b^c^d = 0x96, I = c^(b|~d) = 0x39
__attribute__((target("avx512f"))) static inline void md5_x16_avx512(uint32_t h[16][4], const uint8_t* const ptr[16], size_t nblocks) {
    static const uint32_t K[64] = {
        0xd76aa478,0xe8c7b756,0x242070db,0xc1bdceee,0xf57c0faf,0x4787c62a,0xa8304613,0xfd469501,0x698098d8,0x8b44f7af,0xffff5bb1,0x895cd7be,
        0x6b901122,0xfd987193,0xa679438e,0x49b40821,0xf61e2562,0xc040b340,0x265e5a51,0xe9b6c7aa,0xd62f105d,0x02441453,0xd8a1e681,0xe7d3fbc8,
        0x21e1cde6,0xc33707d6,0xf4d50d87,0x455a14ed,0xa9e3e905,0xfcefa3f8,0x676f02d9,0x8d2a4c8a,0xfffa3942,0x8771f681,0x6d9d6122,0xfde5380c,
        0xa4beea44,0x4bdecfa9,0xf6bb4b60,0xbebfbc70,0x289b7ec6,0xeaa127fa,0xd4ef3085,0x04881d05,0xd9d4d039,0xe6db99e5,0x1fa27cf8,0xc4ac5665,
        0xf4292244,0x432aff97,0xab9423a7,0xfc93a039,0x655b59c3,0x8f0ccc92,0xffeff47d,0x85845dd1,0x6fa87e4f,0xfe2ce6e0,0xa3014314,0x4e0811a1,
        0xf7537e82,0xbd3af235,0x2ad7d2bb,0xeb86d391};
    alignas(64) uint32_t lanes[4][16];
    for (int l = 0; l < 16; l++) for (int i = 0; i < 4; i++) lanes[i][l] = h[l][i];
    __m512i A = _mm512_load_si512(lanes[0]), B = _mm512_load_si512(lanes[1]), Cc = _mm512_load_si512(lanes[2]), D = _mm512_load_si512(lanes[3]);
    for (size_t blk = 0; blk < nblocks; blk++) {
        __m512i r[16], w[16];
        for (int l = 0; l < 16; l++) r[l] = _mm512_loadu_si512((const void*)(ptr[l] + 64 * blk));
        md5_mb_transpose16(r, w);
        __m512i a = A, b = B, c = Cc, d = D;
        FB_STEP16(0xCA, w[0], K[0], 7) FB_STEP16(0xCA, w[1], K[1], 12) FB_STEP16(0xCA, w[2], K[2], 17) FB_STEP16(0xCA, w[3], K[3], 22)
        FB_STEP16(0xCA, w[4], K[4], 7) FB_STEP16(0xCA, w[5], K[5], 12) FB_STEP16(0xCA, w[6], K[6], 17) FB_STEP16(0xCA, w[7], K[7], 22)
        FB_STEP16(0xCA, w[8], K[8], 7) FB_STEP16(0xCA, w[9], K[9], 12) FB_STEP16(0xCA, w[10], K[10], 17) FB_STEP16(0xCA, w[11], K[11], 22)
        FB_STEP16(0xCA, w[12], K[12], 7) FB_STEP16(0xCA, w[13], K[13], 12) FB_STEP16(0xCA, w[14], K[14], 17) FB_STEP16(0xCA, w[15], K[15], 22)
        FB_STEP16(0xE4, w[1], K[16], 5) FB_STEP16(0xE4, w[6], K[17], 9) FB_STEP16(0xE4, w[11], K[18], 14) FB_STEP16(0xE4, w[0], K[19], 20)
        FB_STEP16(0xE4, w[5], K[20], 5) FB_STEP16(0xE4, w[10], K[21], 9) FB_STEP16(0xE4, w[15], K[22], 14) FB_STEP16(0xE4, w[4], K[23], 20)
        FB_STEP16(0xE4, w[9], K[24], 5) FB_STEP16(0xE4, w[14], K[25], 9) FB_STEP16(0xE4, w[3], K[26], 14) FB_STEP16(0xE4, w[8], K[27], 20)
        FB_STEP16(0xE4, w[13], K[28], 5) FB_STEP16(0xE4, w[2], K[29], 9) FB_STEP16(0xE4, w[7], K[30], 14) FB_STEP16(0xE4, w[12], K[31], 20)
        FB_STEP16(0x96, w[5], K[32], 4) FB_STEP16(0x96, w[8], K[33], 11) FB_STEP16(0x96, w[11], K[34], 16) FB_STEP16(0x96, w[14], K[35], 23)
        FB_STEP16(0x96, w[1], K[36], 4) FB_STEP16(0x96, w[4], K[37], 11) FB_STEP16(0x96, w[7], K[38], 16) FB_STEP16(0x96, w[10], K[39], 23)
        FB_STEP16(0x96, w[13], K[40], 4) FB_STEP16(0x96, w[0], K[41], 11) FB_STEP16(0x96, w[3], K[42], 16) FB_STEP16(0x96, w[6], K[43], 23)
        FB_STEP16(0x96, w[9], K[44], 4) FB_STEP16(0x96, w[12], K[45], 11) FB_STEP16(0x96, w[15], K[46], 16) FB_STEP16(0x96, w[2], K[47], 23)
        FB_STEP16(0x39, w[0], K[48], 6) FB_STEP16(0x39, w[7], K[49], 10) FB_STEP16(0x39, w[14], K[50], 15) FB_STEP16(0x39, w[5], K[51], 21)
        FB_STEP16(0x39, w[12], K[52], 6) FB_STEP16(0x39, w[3], K[53], 10) FB_STEP16(0x39, w[10], K[54], 15) FB_STEP16(0x39, w[1], K[55], 21)
        FB_STEP16(0x39, w[8], K[56], 6) FB_STEP16(0x39, w[15], K[57], 10) FB_STEP16(0x39, w[6], K[58], 15) FB_STEP16(0x39, w[13], K[59], 21)
        FB_STEP16(0x39, w[4], K[60], 6) FB_STEP16(0x39, w[11], K[61], 10) FB_STEP16(0x39, w[2], K[62], 15) FB_STEP16(0x39, w[9], K[63], 21)
        A = _mm512_add_epi32(A, a); B = _mm512_add_epi32(B, b); Cc = _mm512_add_epi32(Cc, c); D = _mm512_add_epi32(D, d);
    }
    _mm512_store_si512(lanes[0], A); _mm512_store_si512(lanes[1], B); _mm512_store_si512(lanes[2], Cc); _mm512_store_si512(lanes[3], D);
    for (int l = 0; l < 16; l++) for (int i = 0; i < 4; i++) h[l][i] = lanes[i][l];
}
#undef FB_STEP16

// MD5 digests of up to 8 byte strings (lane l: data[l], len[l]); the common prefix of whole blocks goes through the
// 8-lane kernel, the ragged remainders through the scalar chain.
inline void md5_group8(const uint8_t* const data[8], const size_t len[8], int n, uint8_t digests[8][16]) {
    Md5 m[8];
    for (int l = 0; l < n; l++) m[l].init();
    size_t common = (size_t)-1;
    for (int l = 0; l < n; l++) common = len[l] / 64 < common ? len[l] / 64 : common;
    if (n == 8 && common > 0 && md5_mb_available()) {
        uint32_t h[8][4];
        for (int l = 0; l < 8; l++) for (int i = 0; i < 4; i++) h[l][i] = m[l].h[i];
        md5_x8_avx2(h, data, common);
        for (int l = 0; l < 8; l++) { for (int i = 0; i < 4; i++) m[l].h[i] = h[l][i]; m[l].len = common * 64; }
    } else common = 0;
    for (int l = 0; l < n; l++) { m[l].update(data[l] + common * 64, len[l] - common * 64); m[l].final(digests[l]); }
}

// MD5 digests of up to 16 byte strings: whole-block common prefix through the 16-lane kernel (lanes beyond n repeat
// lane 0 and are discarded), ragged remainders through the scalar chain.
inline void md5_group16(const uint8_t* const data[16], const size_t len[16], int n, uint8_t digests[16][16]) {
    Md5 m[16];
    for (int l = 0; l < n; l++) m[l].init();
    size_t common = (size_t)-1;
    for (int l = 0; l < n; l++) common = len[l] / 64 < common ? len[l] / 64 : common;
    if (n >= 9 && common > 0 && md5_mb16_available()) {
        uint32_t h[16][4]; const uint8_t* p[16];
        for (int l = 0; l < 16; l++) { const int sl = l < n ? l : 0; p[l] = data[sl]; for (int i = 0; i < 4; i++) h[l][i] = m[sl].h[i]; }
        md5_x16_avx512(h, p, common);
        for (int l = 0; l < n; l++) { for (int i = 0; i < 4; i++) m[l].h[i] = h[l][i]; m[l].len = common * 64; }
        for (int l = 0; l < n; l++) { m[l].update(data[l] + common * 64, len[l] - common * 64); m[l].final(digests[l]); }
        return;
    }
    // fall back to two groups of eight
    md5_group8(data, len, n < 8 ? n : 8, digests);
    if (n > 8) md5_group8(data + 8, len + 8, n - 8, digests + 8);
}

}  // namespace fb
