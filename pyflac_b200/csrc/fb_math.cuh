// fb_math.cuh -- the scalar decision arithmetic of the encode path, written once for device code
// (and compilable for the host so tests/ can exercise it without a GPU).
//
// Bit-exactness rules: every double operation rounds on its own -- on the device that
// means explicit __d*_rn intrinsics (nvcc would otherwise contract a*b+c into DFMA); the host build
// of this header is compiled with -ffp-contract=off.  What each function mirrors is cited as
// "up:" = upstream xiph/flac 1.4.3 src/libFLAC (the binary pyFLAC binds, build_args.py:49-51).
#pragma once
#include <stdint.h>
#include <math.h>
#include "fb_common.cuh"

#if defined(__CUDACC__)
#define FB_HD __host__ __device__ __forceinline__
#else
#define FB_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define FB_DMUL(a, b) __dmul_rn((a), (b))
#define FB_DADD(a, b) __dadd_rn((a), (b))
#define FB_DSUB(a, b) __dsub_rn((a), (b))
#define FB_DDIV(a, b) __ddiv_rn((a), (b))
#define FB_FMUL(a, b) __fmul_rn((a), (b))
#else
#define FB_DMUL(a, b) ((a) * (b))
#define FB_DADD(a, b) ((a) + (b))
#define FB_DSUB(a, b) ((a) - (b))
#define FB_DDIV(a, b) ((a) / (b))
#define FB_FMUL(a, b) ((a) * (b))
#endif

namespace fb {

constexpr double kLn2 = 0.69314718055994530942;

FB_HD uint32_t ilog2_u32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return 31u - (uint32_t)__clz((int)v);
#else
    uint32_t l = 0; while (v >>= 1) l++; return l;
#endif
}
FB_HD uint32_t ilog2_u64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return 63u - (uint32_t)__clzll((long long)v);
#else
    uint32_t l = 0; while (v >>= 1) l++; return l;
#endif
}
// up: bitmath.c FLAC__bitmath_silog2
FB_HD uint32_t silog2(int64_t v) {
    if (v == 0) return 0;
    if (v == -1) return 2;
    v = (v < 0) ? (-(v + 1)) : v;
    return ilog2_u64((uint64_t)v) + 2;
}

// up: lpc.c FLAC__lpc_compute_lp_coefficients (Levinson-Durbin).
// lp is [kMaxOrder][kMaxOrder] floats, err[kMaxOrder] doubles; lpc is kMaxOrder doubles of scratch.
// Returns the (possibly reduced) maximum order.
FB_HD int levinson(const double* ac, int max_order, float* lp, double* err_out, double* lpc) {
    double err = ac[0];
    for (int i = 0; i < max_order; i++) {
        double r = -ac[i + 1];
        for (int j = 0; j < i; j++) r = FB_DSUB(r, FB_DMUL(lpc[j], ac[i - j]));
        r = FB_DDIV(r, err);
        lpc[i] = r;
        int j = 0;
        for (; j < (i >> 1); j++) {
            double tmp = lpc[j];
            lpc[j] = FB_DADD(lpc[j], FB_DMUL(r, lpc[i - 1 - j]));
            lpc[i - 1 - j] = FB_DADD(lpc[i - 1 - j], FB_DMUL(r, tmp));
        }
        if (i & 1) lpc[j] = FB_DADD(lpc[j], FB_DMUL(lpc[j], r));
        err = FB_DMUL(err, FB_DSUB(1.0, FB_DMUL(r, r)));
        for (j = 0; j <= i; j++) lp[i * kMaxOrder + j] = (float)(-lpc[j]);
        err_out[i] = err;
        if (err == 0.0) return i + 1;
    }
    return max_order;
}

// up: lpc.c FLAC__lpc_compute_expected_bits_per_residual_sample_with_error_scale.
// *used_log reports whether the value depends on log() (for the ambiguity guard).
FB_HD double expected_bits_per_sample(double lpc_error, double error_scale, bool* used_log) {
    *used_log = false;
    if (lpc_error > 0.0) {
        double bps = FB_DDIV(FB_DMUL(0.5, log(FB_DMUL(error_scale, lpc_error))), kLn2);
        if (bps >= 0.0) { *used_log = true; return bps; }
        return 0.0;
    } else if (lpc_error < 0.0) {
        return 1e32;
    }
    return 0.0;
}

// up: lpc.c FLAC__lpc_quantize_coefficients. 0 = ok, 1/2 = cannot quantize.
FB_HD int quantize_coefficients(const float* lp, int order, int precision, int32_t* q_out, int* shift_out) {
    double cmax = 0.0;
    precision--;
    int32_t qmax = 1 << precision, qmin = -qmax;
    qmax--;
    for (int i = 0; i < order; i++) {
        const double d = fabs((double)lp[i]);
        if (d > cmax) cmax = d;
    }
    if (cmax <= 0.0) return 2;
    int log2cmax;
    (void)frexp(cmax, &log2cmax);
    log2cmax--;
    int shift = precision - log2cmax - 1;
    if (shift > 15) shift = 15;
    else if (shift < -16) return 1;
    if (shift >= 0) {
        double error = 0.0;
        const float scale = (float)(1 << shift);
        for (int i = 0; i < order; i++) {
            error = FB_DADD(error, (double)FB_FMUL(lp[i], scale));
            long long q = llround(error);
            if (q > qmax) q = qmax; else if (q < qmin) q = qmin;
            error = FB_DSUB(error, (double)q);
            q_out[i] = (int32_t)q;
        }
    } else {
        double error = 0.0;
        const float scale = (float)(1 << (-shift));
        for (int i = 0; i < order; i++) {
            error = FB_DADD(error, (double)(lp[i] / scale));
            long long q = llround(error);
            if (q > qmax) q = qmax; else if (q < qmin) q = qmin;
            error = FB_DSUB(error, (double)q);
            q_out[i] = (int32_t)q;
        }
        shift = 0;
    }
    *shift_out = shift;
    return 0;
}

// up: stream_encoder.c set_partitioned_rice_: Rice parameter from the partition's sum of |residual|
// through the truncated 18-bit fixed-point reciprocal.
FB_HD uint32_t rice_parameter(uint64_t sum, uint32_t n, uint32_t rice_limit) {
    const uint32_t div = 0x40000u / n;
    uint32_t k;
    if (sum < 2) k = 0;
    else {
        const uint64_t m = ((sum - 1) * div) >> 18;
        k = (m == 0) ? 0 : ilog2_u64(m) + 1;
    }
    if (k >= rice_limit) k = rice_limit - 1;
    return k;
}
// up: stream_encoder.c count_rice_bits_in_partition_ (estimate; saturates at UINT32_MAX)
FB_HD uint32_t rice_partition_bits(uint32_t k, uint32_t n, uint64_t sum) {
    const uint64_t v = (uint64_t)4 + (uint64_t)((1 + k) * n) + (k ? (sum >> (k - 1)) : (sum << 1)) - (uint64_t)(n >> 1);
    return v < 0xffffffffull ? (uint32_t)v : 0xffffffffu;
}

// ---------------------------------------------------------------- CRC ----
// ref: format.h:456-475 -- CRC-8 poly 0x07 over the frame header, CRC-16 poly 0x8005 over the frame; init 0.
FB_HD uint8_t crc8_byte(uint8_t crc, uint8_t b) {
    crc ^= b;
#pragma unroll
    for (int i = 0; i < 8; i++) crc = (uint8_t)((crc & 0x80) ? ((crc << 1) ^ 0x07) : (crc << 1));
    return crc;
}
FB_HD uint16_t crc16_byte(uint16_t crc, uint8_t b) {
    crc ^= (uint16_t)((uint16_t)b << 8);
#pragma unroll
    for (int i = 0; i < 8; i++) crc = (uint16_t)((crc & 0x8000) ? ((crc << 1) ^ 0x8005) : (crc << 1));
    return crc;
}
// multiply two residues mod the CRC-16 polynomial (carry-less, 16 steps): used to append zero bytes,
// i.e. to combine CRCs of adjacent chunks computed in parallel.
FB_HD uint16_t crc16_mulmod(uint16_t a, uint16_t b) {
    uint32_t r = 0;
#pragma unroll
    for (int i = 15; i >= 0; i--) {
        r <<= 1;
        if (r & 0x10000u) r ^= 0x18005u;
        if ((b >> i) & 1) r ^= a;
    }
    return (uint16_t)r;
}

// Positional weights for a CTA-parallel CRC-16: the frame is cut into 64-byte chunks counted from its END,
// chunk j's CRC is multiplied by x^(512 j) mod P (so that it stands where the chunk stands) and everything is
// XORed: crc(A||B) = crc(A) * x^(8|B|) + crc(B).  lo[j] = x^(512 j), j < 256; hi[h] = x^(512*256*h), h < 16.
// Built at compile time; one mulmod per chunk instead of a log-depth tree of them.
struct CrcPosTable { uint16_t lo[256]; uint16_t hi[16]; };
constexpr uint16_t crc16_mulmod_c(uint16_t a, uint16_t b) {
    uint32_t r = 0;
    for (int i = 15; i >= 0; i--) { r <<= 1; if (r & 0x10000u) r ^= 0x18005u; if ((b >> i) & 1) r ^= a; }
    return (uint16_t)r;
}
constexpr CrcPosTable make_crc_pos_table() {
    CrcPosTable t{};
    uint16_t x512 = 2;
    for (int i = 0; i < 9; i++) x512 = crc16_mulmod_c(x512, x512);
    t.lo[0] = 1;
    for (int j = 1; j < 256; j++) t.lo[j] = crc16_mulmod_c(t.lo[j - 1], x512);
    const uint16_t step_hi = crc16_mulmod_c(t.lo[255], x512);
    t.hi[0] = 1;
    for (int j = 1; j < 16; j++) t.hi[j] = crc16_mulmod_c(t.hi[j - 1], step_hi);
    return t;
}
// Weight of chunk j for ANY j: the tables cover j < 4096 (frames under 256 KiB -- everything the encoder can produce, whose frames
// must fit shared memory); a decoder can meet bigger frames (blocksize 65535, or many wide channels), and there the remaining factor
// x^(512 * 4096 * (j >> 12)) comes from square-and-multiply.  Host and device (tests/test_abi_cpu.py checks it against a bytewise CRC).
FB_HD uint16_t crc16_weigh_chunk(uint16_t c16, uint32_t j, const CrcPosTable& t) {
    if (j) c16 = crc16_mulmod(c16, t.lo[j & 255u]);
    if (j >> 8) c16 = crc16_mulmod(c16, t.hi[(j >> 8) & 15u]);
    uint32_t h = j >> 12;
    if (h) {
        uint16_t step = crc16_mulmod(t.hi[15], t.hi[1]);                // x^(512 * 4096)
        for (; h; h >>= 1) { if (h & 1u) c16 = crc16_mulmod(c16, step); step = crc16_mulmod(step, step); }
    }
    return c16;
}

#if defined(__CUDACC__)
static __device__ const CrcPosTable g_crc_pos = make_crc_pos_table();

// Slice-by-4 tables for the same CRC: t[0] is the plain byte table, t[k][x] = CRC of byte x followed by k zero bytes, so
// one 32-bit step is four INDEPENDENT lookups instead of four dependent ones.  Built at compile time.
struct __align__(16) Crc16Slice { uint16_t t[4][256]; };
constexpr Crc16Slice make_crc16_slice() {
    Crc16Slice s{};
    for (int x = 0; x < 256; x++) {
        uint16_t v = (uint16_t)(x << 8);
        for (int i = 0; i < 8; i++) v = (uint16_t)((v & 0x8000) ? ((v << 1) ^ 0x8005) : (v << 1));
        s.t[0][x] = v;
    }
    for (int k = 1; k < 4; k++)
        for (int x = 0; x < 256; x++) s.t[k][x] = (uint16_t)((s.t[k - 1][x] << 8) ^ s.t[0][s.t[k - 1][x] >> 8]);
    return s;
}
static __device__ const Crc16Slice g_crc16_slice = make_crc16_slice();

// As cta_crc16 below (but without a CTA-wide wait: only warp 0 blocks), for input that can be fetched four bytes at a time: word_at(j) returns bytes j..j+3 (byte j in the
// low bits; j is arbitrary, reading up to 3 bytes past nb must be harmless); tabs = the four tables in shared memory.
template <int THREADS, typename WordAt>
__device__ __forceinline__ uint16_t cta_crc16_words(WordAt word_at, uint32_t nb, const uint16_t (*tabs)[256], uint32_t* warp_x, int tid) {
    const uint32_t nchunks = (nb + 63u) >> 6;
    uint32_t acc = 0;
    for (uint32_t j = (uint32_t)tid; j < nchunks; j += THREADS) {          // j counts chunks from the end of the frame
        const uint32_t end = nb - (j << 6);
        uint32_t crc = 0;
        if (end >= 64u) {
            const uint32_t beg = end - 64u;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const uint32_t w = word_at(beg + 4u * (uint32_t)i);
                crc = (uint32_t)tabs[3][(crc >> 8) ^ (w & 0xffu)] ^ tabs[2][(crc & 0xffu) ^ ((w >> 8) & 0xffu)] ^ tabs[1][(w >> 16) & 0xffu] ^ tabs[0][w >> 24];
            }
        } else {
            for (uint32_t b = 0; b < end; b++) crc = ((crc << 8) & 0xffffu) ^ tabs[0][(crc >> 8) ^ (word_at(b) & 0xffu)];
        }
        uint16_t c16 = (uint16_t)crc;
        if (j) c16 = crc16_mulmod(c16, g_crc_pos.lo[j & 255u]);
        if (j >> 8) c16 = crc16_mulmod(c16, g_crc_pos.hi[(j >> 8) & 15u]);
        acc ^= c16;
    }
    acc = __reduce_xor_sync(0xffffffffu, acc);
    if ((tid & 31) == 0) warp_x[tid >> 5] = acc;
    // only warp 0 needs the partial results: the other warps signal and go on with their next job (named barrier 1)
    uint32_t r = 0;
    if (tid < 32) {
        asm volatile("bar.sync 1, %0;" :: "n"(THREADS) : "memory");
        r = (tid < THREADS / 32) ? warp_x[tid] : 0u;
        r = __reduce_xor_sync(0xffffffffu, r);
    } else {
        __threadfence_block();
        asm volatile("bar.arrive 1, %0;" :: "n"(THREADS) : "memory");
    }
    return (uint16_t)r;
}

// CRC-16 (poly 0x8005, init 0) of nb bytes by a whole CTA of THREADS threads.  byte_at(j) returns byte j;
// crc_tab is the 256-entry byte table in shared memory; warp_x is THREADS/32 words of shared scratch.
// The result is valid on warp 0 (all lanes) after the call; contains one __syncthreads().
template <int THREADS, typename ByteAt>
__device__ __forceinline__ uint16_t cta_crc16(ByteAt byte_at, uint32_t nb, const uint16_t* crc_tab, uint32_t* warp_x, int tid) {
    const uint32_t nchunks = (nb + 63u) >> 6;
    uint32_t acc = 0;
    for (uint32_t j = (uint32_t)tid; j < nchunks; j += THREADS) {          // j counts chunks from the end of the frame
        const uint32_t end = nb - (j << 6);
        const uint32_t beg = end >= 64u ? end - 64u : 0u;
        uint16_t crc = 0;
        for (uint32_t b = beg; b < end; b++) crc = (uint16_t)((crc << 8) ^ crc_tab[(crc >> 8) ^ byte_at(b)]);
        if (j) crc = crc16_mulmod(crc, g_crc_pos.lo[j & 255u]);
        if (j >> 8) crc = crc16_mulmod(crc, g_crc_pos.hi[(j >> 8) & 15u]);
        acc ^= crc;
    }
    acc = __reduce_xor_sync(0xffffffffu, acc);
    if ((tid & 31) == 0) warp_x[tid >> 5] = acc;
    __syncthreads();
    uint32_t r = 0;
    if (tid < 32) {
        r = (tid < THREADS / 32) ? warp_x[tid] : 0u;
        r = __reduce_xor_sync(0xffffffffu, r);
    }
    return (uint16_t)r;
}
#endif

// ---------------------------------------------------------------- frame header ----
// up: stream_encoder_framing.c FLAC__frame_add_header (ref: format.h:416-462).
// Writes at most 16 bytes (incl. CRC-8) into out[], returns the byte count.
// ca: 0 independent, 1 left/side, 2 right/side, 3 mid/side.
FB_HD int build_frame_header(uint8_t* out, uint32_t channels, uint32_t bps, uint32_t sample_rate,
                             uint32_t N, uint32_t frame_number, int ca) {
    int n = 0;
    uint32_t u, bs_hint = 0, sr_hint = 0;
    out[n++] = 0xFF;
    out[n++] = 0xF8;
    switch (N) {
        case 192: u = 1; break; case 576: u = 2; break; case 1152: u = 3; break; case 2304: u = 4; break;
        case 4608: u = 5; break; case 256: u = 8; break; case 512: u = 9; break; case 1024: u = 10; break;
        case 2048: u = 11; break; case 4096: u = 12; break; case 8192: u = 13; break; case 16384: u = 14; break;
        case 32768: u = 15; break;
        default: bs_hint = u = (N <= 0x100) ? 6 : 7; break;
    }
    uint32_t b2 = u << 4;
    switch (sample_rate) {
        case 88200: u = 1; break; case 176400: u = 2; break; case 192000: u = 3; break; case 8000: u = 4; break;
        case 16000: u = 5; break; case 22050: u = 6; break; case 24000: u = 7; break; case 32000: u = 8; break;
        case 44100: u = 9; break; case 48000: u = 10; break; case 96000: u = 11; break;
        default:
            if (sample_rate <= 255000 && sample_rate % 1000 == 0) sr_hint = u = 12;
            else if (sample_rate <= 655350 && sample_rate % 10 == 0) sr_hint = u = 14;
            else if (sample_rate <= 0xffff) sr_hint = u = 13;
            else u = 0;
            break;
    }
    out[n++] = (uint8_t)(b2 | u);
    switch (ca) { case 0: u = channels - 1; break; case 1: u = 8; break; case 2: u = 9; break; default: u = 10; break; }
    uint32_t b3 = u << 4;
    switch (bps) { case 8: u = 1; break; case 12: u = 2; break; case 16: u = 4; break; case 20: u = 5; break;
                   case 24: u = 6; break; case 32: u = 7; break; default: u = 0; break; }
    out[n++] = (uint8_t)(b3 | (u << 1));
    // UTF-8 style coded frame number
    const uint32_t v = frame_number;
    if (v < 0x80) out[n++] = (uint8_t)v;
    else if (v < 0x800) { out[n++] = (uint8_t)(0xC0 | (v >> 6)); out[n++] = (uint8_t)(0x80 | (v & 0x3F)); }
    else if (v < 0x10000) { out[n++] = (uint8_t)(0xE0 | (v >> 12)); out[n++] = (uint8_t)(0x80 | ((v >> 6) & 0x3F)); out[n++] = (uint8_t)(0x80 | (v & 0x3F)); }
    else if (v < 0x200000) { out[n++] = (uint8_t)(0xF0 | (v >> 18)); out[n++] = (uint8_t)(0x80 | ((v >> 12) & 0x3F)); out[n++] = (uint8_t)(0x80 | ((v >> 6) & 0x3F)); out[n++] = (uint8_t)(0x80 | (v & 0x3F)); }
    else if (v < 0x4000000) { out[n++] = (uint8_t)(0xF8 | (v >> 24)); out[n++] = (uint8_t)(0x80 | ((v >> 18) & 0x3F)); out[n++] = (uint8_t)(0x80 | ((v >> 12) & 0x3F)); out[n++] = (uint8_t)(0x80 | ((v >> 6) & 0x3F)); out[n++] = (uint8_t)(0x80 | (v & 0x3F)); }
    else { out[n++] = (uint8_t)(0xFC | (v >> 30)); out[n++] = (uint8_t)(0x80 | ((v >> 24) & 0x3F)); out[n++] = (uint8_t)(0x80 | ((v >> 18) & 0x3F)); out[n++] = (uint8_t)(0x80 | ((v >> 12) & 0x3F)); out[n++] = (uint8_t)(0x80 | ((v >> 6) & 0x3F)); out[n++] = (uint8_t)(0x80 | (v & 0x3F)); }
    if (bs_hint == 6) out[n++] = (uint8_t)(N - 1);
    else if (bs_hint == 7) { out[n++] = (uint8_t)((N - 1) >> 8); out[n++] = (uint8_t)(N - 1); }
    if (sr_hint == 12) out[n++] = (uint8_t)(sample_rate / 1000);
    else if (sr_hint == 13) { out[n++] = (uint8_t)(sample_rate >> 8); out[n++] = (uint8_t)sample_rate; }
    else if (sr_hint == 14) { out[n++] = (uint8_t)((sample_rate / 10) >> 8); out[n++] = (uint8_t)(sample_rate / 10); }
    uint8_t crc = 0;
    for (int i = 0; i < n; i++) crc = crc8_byte(crc, out[i]);
    out[n++] = crc;
    return n;
}

}  // namespace fb
