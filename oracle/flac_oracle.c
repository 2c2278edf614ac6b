/*
 * flac_oracle.c -- TEST INFRASTRUCTURE ONLY (see flac_oracle.h).
 *
 * Plain-C restatement of the libFLAC 1.4.3 encode decisions + bitstream writer and of the decode
 * path.  Each function cites what it follows: "up:" = upstream xiph/flac 1.4.3 src/libFLAC (pinned
 * by /root/reference/scripts/linux.sh:6, .github/workflows/build.yml:9, CHANGELOG.rst:14 -- source
 * not in the reference tree), "ref:" = a file under /root/reference, "SV" = SURVEY.md section.
 * Compile with -ffp-contract=off: every double operation must round on its own (SV A.6/A.7).
 */
#include "flac_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#ifndef M_LN2
#define M_LN2 0.69314718055994530942
#endif
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ref: pyflac/include/FLAC/format.h:93-151 (limits), :168-257, :416-475 (field widths) */
#define MAX_FIXED_ORDER 4
#define MAX_LPC_ORDER 32
#define MIN_QLP_PREC 5
#define MAX_QLP_PREC 15
#define RICE_ESC 15
#define RICE2_ESC 31
#define MAX_EXTRA_RESIDUAL_BPS 4

static const char VENDOR[] = "reference libFLAC 1.4.3 20230623"; /* SV 0.2, Appendix B */

/* ------------------------------------------------------------------ CRC, MD5 ---- */

/* up: crc.c -- CRC-8 poly 0x07, CRC-16 poly 0x8005, init 0, MSB first (ref: format.h:456-475) */
uint8_t fo_crc8(const uint8_t *d, size_t n)
{
    uint8_t c = 0;
    while (n--) {
        c ^= *d++;
        for (int i = 0; i < 8; i++) c = (uint8_t)((c & 0x80) ? (c << 1) ^ 0x07 : (c << 1));
    }
    return c;
}
uint16_t fo_crc16(const uint8_t *d, size_t n)
{
    uint16_t c = 0;
    while (n--) {
        c ^= (uint16_t)(*d++ << 8);
        for (int i = 0; i < 8; i++) c = (uint16_t)((c & 0x8000) ? (c << 1) ^ 0x8005 : (c << 1));
    }
    return c;
}

/* RFC 1321; up: md5.c (FLAC__MD5Accumulate feeds (bps+7)/8 LE bytes per sample, interleaved; SV E1) */
typedef struct { uint32_t a, b, c, d; uint64_t len; uint8_t buf[64]; uint32_t fill; } md5_t;
static const uint32_t MD5K[64] = {
    0xd76aa478,0xe8c7b756,0x242070db,0xc1bdceee,0xf57c0faf,0x4787c62a,0xa8304613,0xfd469501,
    0x698098d8,0x8b44f7af,0xffff5bb1,0x895cd7be,0x6b901122,0xfd987193,0xa679438e,0x49b40821,
    0xf61e2562,0xc040b340,0x265e5a51,0xe9b6c7aa,0xd62f105d,0x02441453,0xd8a1e681,0xe7d3fbc8,
    0x21e1cde6,0xc33707d6,0xf4d50d87,0x455a14ed,0xa9e3e905,0xfcefa3f8,0x676f02d9,0x8d2a4c8a,
    0xfffa3942,0x8771f681,0x6d9d6122,0xfde5380c,0xa4beea44,0x4bdecfa9,0xf6bb4b60,0xbebfbc70,
    0x289b7ec6,0xeaa127fa,0xd4ef3085,0x04881d05,0xd9d4d039,0xe6db99e5,0x1fa27cf8,0xc4ac5665,
    0xf4292244,0x432aff97,0xab9423a7,0xfc93a039,0x655b59c3,0x8f0ccc92,0xffeff47d,0x85845dd1,
    0x6fa87e4f,0xfe2ce6e0,0xa3014314,0x4e0811a1,0xf7537e82,0xbd3af235,0x2ad7d2bb,0xeb86d391 };
static const uint8_t MD5S[64] = { 7,12,17,22,7,12,17,22,7,12,17,22,7,12,17,22, 5,9,14,20,5,9,14,20,5,9,14,20,5,9,14,20,
    4,11,16,23,4,11,16,23,4,11,16,23,4,11,16,23, 6,10,15,21,6,10,15,21,6,10,15,21,6,10,15,21 };
static void md5_block(md5_t *m, const uint8_t *p)
{
    uint32_t w[16], a = m->a, b = m->b, c = m->c, d = m->d;
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4*i] | (uint32_t)p[4*i+1] << 8 | (uint32_t)p[4*i+2] << 16 | (uint32_t)p[4*i+3] << 24;
    for (int i = 0; i < 64; i++) {
        uint32_t f; int g;
        if (i < 16) { f = (b & c) | (~b & d); g = i; }
        else if (i < 32) { f = (d & b) | (~d & c); g = (5*i + 1) & 15; }
        else if (i < 48) { f = b ^ c ^ d; g = (3*i + 5) & 15; }
        else { f = c ^ (b | ~d); g = (7*i) & 15; }
        f += a + MD5K[i] + w[g];
        a = d; d = c; c = b;
        b += (f << MD5S[i]) | (f >> (32 - MD5S[i]));
    }
    m->a += a; m->b += b; m->c += c; m->d += d;
}
static void md5_init(md5_t *m) { m->a = 0x67452301; m->b = 0xefcdab89; m->c = 0x98badcfe; m->d = 0x10325476; m->len = 0; m->fill = 0; }
static void md5_update(md5_t *m, const uint8_t *p, size_t n)
{
    m->len += n;
    while (n) {
        size_t k = 64 - m->fill; if (k > n) k = n;
        memcpy(m->buf + m->fill, p, k); m->fill += (uint32_t)k; p += k; n -= k;
        if (m->fill == 64) { md5_block(m, m->buf); m->fill = 0; }
    }
}
static void md5_final(md5_t *m, uint8_t out[16])
{
    uint64_t bits = m->len * 8; uint8_t pad = 0x80, z = 0, lenb[8];
    md5_update(m, &pad, 1);
    while (m->fill != 56) md5_update(m, &z, 1);
    for (int i = 0; i < 8; i++) lenb[i] = (uint8_t)(bits >> (8*i));
    md5_update(m, lenb, 8);
    uint32_t v[4] = { m->a, m->b, m->c, m->d };
    for (int i = 0; i < 16; i++) out[i] = (uint8_t)(v[i/4] >> (8*(i&3)));
}
void fo_md5(const uint8_t *d, size_t n, uint8_t out[16]) { md5_t m; md5_init(&m); md5_update(&m, d, n); md5_final(&m, out); }

/* ------------------------------------------------------------------ bit writer ---- */

/* up: bitwriter.c -- MSB-first; ref: format.h "all numbers big-endian" */
typedef struct { uint8_t *buf; size_t cap; size_t bits; int overflow; } bw_t;
static void bw_put(bw_t *w, uint64_t v, unsigned n)
{
    while (n) {
        size_t byte = w->bits >> 3; unsigned used = (unsigned)(w->bits & 7), room = 8 - used;
        unsigned k = n < room ? n : room;
        if (byte >= w->cap) { w->overflow = 1; return; }
        uint8_t chunk = (uint8_t)((v >> (n - k)) & ((1u << k) - 1));
        if (used == 0) w->buf[byte] = 0;
        w->buf[byte] |= (uint8_t)(chunk << (room - k));
        w->bits += k; n -= k;
    }
}
static void bw_put_signed(bw_t *w, int64_t v, unsigned n) { bw_put(w, n >= 64 ? (uint64_t)v : ((uint64_t)v & ((1ull << n) - 1)), n); }
/* up: FLAC__bitwriter_write_unary_unsigned: val zeros then a one */
static void bw_unary(bw_t *w, uint32_t val) { while (val >= 32) { bw_put(w, 0, 32); val -= 32; } bw_put(w, 1, val + 1); }
/* up: FLAC__bitwriter_write_utf8_uint32 (frame number; ref: format.h:441-449) */
static void bw_utf8(bw_t *w, uint32_t v)
{
    if (v < 0x80) bw_put(w, v, 8);
    else if (v < 0x800) { bw_put(w, 0xC0 | (v >> 6), 8); bw_put(w, 0x80 | (v & 0x3F), 8); }
    else if (v < 0x10000) { bw_put(w, 0xE0 | (v >> 12), 8); bw_put(w, 0x80 | ((v >> 6) & 0x3F), 8); bw_put(w, 0x80 | (v & 0x3F), 8); }
    else if (v < 0x200000) { bw_put(w, 0xF0 | (v >> 18), 8); bw_put(w, 0x80 | ((v >> 12) & 0x3F), 8); bw_put(w, 0x80 | ((v >> 6) & 0x3F), 8); bw_put(w, 0x80 | (v & 0x3F), 8); }
    else if (v < 0x4000000) { bw_put(w, 0xF8 | (v >> 24), 8); bw_put(w, 0x80 | ((v >> 18) & 0x3F), 8); bw_put(w, 0x80 | ((v >> 12) & 0x3F), 8); bw_put(w, 0x80 | ((v >> 6) & 0x3F), 8); bw_put(w, 0x80 | (v & 0x3F), 8); }
    else { bw_put(w, 0xFC | (v >> 30), 8); bw_put(w, 0x80 | ((v >> 24) & 0x3F), 8); bw_put(w, 0x80 | ((v >> 18) & 0x3F), 8); bw_put(w, 0x80 | ((v >> 12) & 0x3F), 8); bw_put(w, 0x80 | ((v >> 6) & 0x3F), 8); bw_put(w, 0x80 | (v & 0x3F), 8); }
}
/* up: FLAC__bitwriter_write_rice_signed_block: zig-zag fold, q zeros, stop bit, k LSBs (SV Appendix B) */
static void bw_rice(bw_t *w, int32_t val, unsigned k)
{
    uint32_t u = ((uint32_t)val << 1) ^ (uint32_t)(val >> 31);
    bw_unary(w, u >> k);
    if (k) bw_put(w, u & ((1u << k) - 1), k);
}

/* ------------------------------------------------------------------ settings ---- */

/* up: stream_encoder.c compression_levels_[] ; ref: stream_encoder.h:845-853 ; SV A.1 */
static const struct { int ms, loose; uint32_t max_lpc; uint32_t max_po; int parts; } LEVELS[9] = {
    {0,0,0,3,1}, {1,1,0,3,1}, {1,0,0,3,1}, {0,0,6,4,1}, {1,1,8,4,1}, {1,0,8,5,1}, {1,0,8,6,2}, {1,0,12,6,2}, {1,0,12,6,3} };

typedef struct {
    uint32_t sample_rate, channels, bps, blocksize;
    int do_ms, loose; uint32_t loose_frames;
    uint32_t max_lpc_order, qlp_precision, max_part_order;
    int apod_parts;            /* 1 = tukey(0.5); 2,3 = subdivide_tukey(parts) */
    float apod_p;
    int limit_min_bitrate;
} settings_t;

static uint32_t ilog2_u32(uint32_t v) { uint32_t l = 0; while (v >>= 1) l++; return l; }           /* up: FLAC__bitmath_ilog2 */
static uint32_t ilog2_u64(uint64_t v) { uint32_t l = 0; while (v >>= 1) l++; return l; }           /* up: FLAC__bitmath_ilog2_wide */
static uint32_t silog2(int64_t v)                                                                  /* up: FLAC__bitmath_silog2 */
{
    if (v == 0) return 0;
    if (v == -1) return 2;
    v = (v < 0) ? (-(v + 1)) : v;
    return ilog2_u64((uint64_t)v) + 2;
}

/* up: stream_encoder.c init_stream_internal_ validation order, verified by 26 probes (SV A.1).
 * Status values: ref: pyflac/builder/encoder.py:65-80 */
int fo_encoder_init_status(const fo_enc_cfg *c, int has_write, int has_seek, int has_tell)
{
    uint32_t bs = c->blocksize, maxlpc;
    if (!has_write || (has_seek && !has_tell)) return 3;            /* INVALID_CALLBACKS */
    if (c->channels == 0 || c->channels > 8) return 4;              /* INVALID_NUMBER_OF_CHANNELS */
    if (c->bps < 4 || c->bps > 32) return 5;                        /* INVALID_BITS_PER_SAMPLE */
    if (c->sample_rate > 1048575u) return 6;                        /* INVALID_SAMPLE_RATE */
    if (c->level > 8) maxlpc = LEVELS[8].max_lpc; else maxlpc = LEVELS[c->level].max_lpc;
    if (bs == 0) bs = maxlpc == 0 ? 1152 : 4096;
    if (bs < 16 || bs > 65535) return 7;                            /* INVALID_BLOCK_SIZE */
    if (maxlpc > MAX_LPC_ORDER) return 8;                           /* INVALID_MAX_LPC_ORDER */
    if (bs < maxlpc) return 10;                                     /* BLOCK_SIZE_TOO_SMALL_FOR_LPC_ORDER */
    if (c->streamable_subset) {
        if (!(c->bps == 8 || c->bps == 12 || c->bps == 16 || c->bps == 20 || c->bps == 24 || c->bps == 32)) return 11; /* NOT_STREAMABLE */
        if (c->sample_rate <= 48000 && (bs > 4608 || maxlpc > 12)) return 11;
        if (bs > 16384) return 11;
        /* up: FLAC__format_sample_rate_is_subset: above 16 bits the frame header can only carry multiples of 10 Hz */
        if (c->sample_rate >= (1u << 16) && c->sample_rate % 10u != 0u) return 11;
    }
    return 0;
}

static int resolve_settings(const fo_enc_cfg *c, settings_t *s)
{
    uint32_t lvl = c->level > 8 ? 8 : c->level;
    memset(s, 0, sizeof *s);
    s->sample_rate = c->sample_rate; s->channels = c->channels; s->bps = c->bps;
    s->do_ms = LEVELS[lvl].ms; s->loose = LEVELS[lvl].loose;
    s->max_lpc_order = LEVELS[lvl].max_lpc; s->max_part_order = LEVELS[lvl].max_po;
    s->apod_parts = LEVELS[lvl].parts;
    s->apod_p = s->apod_parts == 1 ? 0.5f : 0.5f / s->apod_parts;  /* up: set_apodization "subdivide_tukey(": p/parts in float (SV A.5) */
    s->limit_min_bitrate = c->limit_min_bitrate;
    if (s->channels != 2) { s->do_ms = 0; s->loose = 0; } else if (!s->do_ms) s->loose = 0;
    s->blocksize = c->blocksize ? c->blocksize : (s->max_lpc_order == 0 ? 1152 : 4096);
    /* auto qlp precision: up: init_stream_internal_ ; verified at 13 blocksizes (SV A.1) */
    if (s->bps < 16) { uint32_t p = 2 + s->bps / 2; s->qlp_precision = p < MIN_QLP_PREC ? MIN_QLP_PREC : p; }
    else if (s->bps == 16) {
        uint32_t b = s->blocksize;
        s->qlp_precision = b <= 192 ? 7 : b <= 384 ? 8 : b <= 576 ? 9 : b <= 1152 ? 10 : b <= 2304 ? 11 : b <= 4608 ? 12 : 13;
    } else {
        uint32_t b = s->blocksize;
        s->qlp_precision = b <= 384 ? MAX_QLP_PREC - 2 : b <= 1152 ? MAX_QLP_PREC - 1 : MAX_QLP_PREC;
    }
    if (s->loose) {
        s->loose_frames = (uint32_t)((double)s->sample_rate * 0.4 / (double)s->blocksize + 0.5);
        if (s->loose_frames == 0) s->loose_frames = 1;
    }
    return 0;
}

/* ------------------------------------------------------------------ window ---- */

/* up: window.c FLAC__window_tukey (SV A.5): rectangle with raised-cosine ends, Np computed in float */
void fo_window_tukey(float *w, int32_t L, float p)
{
    int32_t n;
    if (p <= 0.0f) { for (n = 0; n < L; n++) w[n] = 1.0f; return; }
    if (p >= 1.0f) { /* hann */
        const int32_t N = L - 1;
        for (n = 0; n < L; n++) w[n] = (float)(0.5f - 0.5f * cosf(2.0f * (float)M_PI * n / N));
        return;
    }
    {
        const int32_t Np = (int32_t)(p / 2.0f * L) - 1;
        for (n = 0; n < L; n++) w[n] = 1.0f;
        if (Np > 0) {
            for (n = 0; n <= Np; n++) {
                w[n] = (float)(0.5f - 0.5f * cosf((float)(M_PI * n / Np)));
                w[L - Np - 1 + n] = (float)(0.5f - 0.5f * cosf((float)(M_PI * (n + Np) / Np)));
            }
        }
    }
}

/* ------------------------------------------------------------------ LPC analysis ---- */

/* up: lpc.c FLAC__lpc_compute_autocorrelation: one double accumulator per lag, ascending i (SV A.6, E5).
 * d[] are floats, each product is formed in double (exact), each add rounds once. */
static void autocorrelation(const float *d, uint32_t len, uint32_t lags, double *autoc)
{
    for (uint32_t j = 0; j < lags; j++) {
        double acc = 0.0;
        for (uint32_t i = j; i < len; i++) acc += (double)d[i] * (double)d[i - j];
        autoc[j] = acc;
    }
}

/* up: lpc.c FLAC__lpc_compute_lp_coefficients (Levinson-Durbin, SV A.6). Returns possibly reduced max order. */
static uint32_t lp_coefficients(const double *autoc, uint32_t max_order, float lp[][MAX_LPC_ORDER], double *error)
{
    double r, err, lpc[MAX_LPC_ORDER];
    uint32_t i, j;
    err = autoc[0];
    for (i = 0; i < max_order; i++) {
        r = -autoc[i + 1];
        for (j = 0; j < i; j++) r -= lpc[j] * autoc[i - j];
        r /= err;
        lpc[i] = r;
        for (j = 0; j < (i >> 1); j++) {
            double tmp = lpc[j];
            lpc[j] += r * lpc[i - 1 - j];
            lpc[i - 1 - j] += r * tmp;
        }
        if (i & 1) lpc[j] += lpc[j] * r;
        err *= (1.0 - r * r);
        for (j = 0; j <= i; j++) lp[i][j] = (float)(-lpc[j]);
        error[i] = err;
        if (err == 0.0) return i + 1;
    }
    return max_order;
}

/* up: lpc.c FLAC__lpc_compute_expected_bits_per_residual_sample_with_error_scale (SV A.7) */
static double expected_bits_per_sample(double lpc_error, double error_scale)
{
    if (lpc_error > 0.0) {
        double bps = (double)0.5 * log(error_scale * lpc_error) / M_LN2;
        return bps >= 0.0 ? bps : 0.0;
    } else if (lpc_error < 0.0) return 1e32;
    return 0.0;
}
/* up: lpc.c FLAC__lpc_compute_best_order (first strict minimum, initial best = (uint32_t)-1) */
static uint32_t best_lpc_order(const double *lpc_error, uint32_t max_order, uint32_t total_samples, uint32_t overhead_bits_per_order)
{
    uint32_t order, indx, best_index = 0;
    double bits, best_bits = (double)(uint32_t)(-1), error_scale = 0.5 / (double)total_samples;
    for (indx = 0, order = 1; indx < max_order; indx++, order++) {
        bits = expected_bits_per_sample(lpc_error[indx], error_scale) * (double)(total_samples - order) + (double)(order * overhead_bits_per_order);
        if (bits < best_bits) { best_index = indx; best_bits = bits; }
    }
    return best_index + 1;
}

/* up: lpc.c FLAC__lpc_quantize_coefficients (SV A.7). Returns 0 ok, 1/2 failure. */
static int quantize_coefficients(const float *lp, uint32_t order, uint32_t precision, int32_t *q_out, int *shift)
{
    uint32_t i; double cmax = 0.0; int32_t qmax, qmin;
    precision--;
    qmax = 1 << precision; qmin = -qmax; qmax--;
    for (i = 0; i < order; i++) { const double d = fabs(lp[i]); if (d > cmax) cmax = d; }
    if (cmax <= 0.0) return 2;
    {
        const int max_shiftlimit = (1 << 4) - 1, min_shiftlimit = -max_shiftlimit - 1;
        int log2cmax;
        (void)frexp(cmax, &log2cmax);
        log2cmax--;
        *shift = (int)precision - log2cmax - 1;
        if (*shift > max_shiftlimit) *shift = max_shiftlimit;
        else if (*shift < min_shiftlimit) return 1;
    }
    if (*shift >= 0) {
        double error = 0.0; int32_t q;
        for (i = 0; i < order; i++) {
            error += lp[i] * (1 << *shift);   /* float * int -> float product (exact, power of two), then double add */
            q = (int32_t)lround(error);
            if (q > qmax) q = qmax; else if (q < qmin) q = qmin;
            error -= q;
            q_out[i] = q;
        }
    } else {
        const int nshift = -(*shift);
        double error = 0.0; int32_t q;
        for (i = 0; i < order; i++) {
            error += lp[i] / (1 << nshift);
            q = (int32_t)lround(error);
            if (q > qmax) q = qmax; else if (q < qmin) q = qmin;
            error -= q;
            q_out[i] = q;
        }
        *shift = 0;
    }
    return 0;
}

/* up: lpc.c FLAC__lpc_max_prediction_before_shift_bps / FLAC__lpc_max_residual_bps */
static uint32_t max_prediction_before_shift_bps(uint32_t sbps, const int32_t *q, uint32_t order)
{
    int32_t s = 0; for (uint32_t i = 0; i < order; i++) s += abs(q[i]);
    if (s == 0) s = 1;
    return sbps + silog2(s);
}
static uint32_t max_residual_bps(uint32_t sbps, const int32_t *q, uint32_t order, int shift)
{
    int32_t p = (int32_t)max_prediction_before_shift_bps(sbps, q, order) - shift;
    return ((int32_t)sbps > p) ? sbps + 1 : (uint32_t)p + 1;
}

/* up: lpc.c FLAC__lpc_compute_residual_from_qlp_coefficients[_wide|_limit_residual] (SV E9):
 * r[i] = x[i] - ((sum_j q[j]*x[i-1-j]) >> shift); 64-bit accumulate is exact for every variant that
 * libFLAC would pick; `limit` reproduces the _limit_residual rejection. data points at sample `order`. */
static int lpc_residual(const int64_t *data, uint32_t n, const int32_t *q, uint32_t order, int shift, int limit, int32_t *res)
{
    for (uint32_t i = 0; i < n; i++) {
        int64_t sum = 0;
        for (uint32_t j = 0; j < order; j++) sum += (int64_t)q[j] * (int64_t)data[(int64_t)i - 1 - j];
        int64_t r = (int64_t)data[i] - (sum >> shift);
        if (limit && (r <= INT32_MIN || r > INT32_MAX)) return 0;
        res[i] = (int32_t)r;
    }
    return 1;
}

/* ------------------------------------------------------------------ fixed predictors ---- */

/* up: fixed.c FLAC__fixed_compute_best_predictor[_wide] (SV A.4).  data points at sample 4, len = N-4.
 * `narrow`: libFLAC used 32-bit accumulators (sums wrap mod 2^32) -- reproduced for exactness. */
static uint32_t fixed_best_predictor(const int64_t *data, uint32_t len, int narrow, uint64_t err_out[5])
{
    uint64_t e0 = 0, e1 = 0, e2 = 0, e3 = 0, e4 = 0;
    for (uint32_t i = 0; i < len; i++) {
        int64_t x0 = data[i], x1 = data[(int64_t)i - 1], x2 = data[(int64_t)i - 2], x3 = data[(int64_t)i - 3], x4 = data[(int64_t)i - 4];
        int64_t d1 = x0 - x1, d2 = d1 - (x1 - x2), d3 = d2 - ((x1 - x2) - (x2 - x3));
        int64_t d4 = d3 - (((x1 - x2) - (x2 - x3)) - ((x2 - x3) - (x3 - x4)));
        e0 += (uint64_t)llabs(x0); e1 += (uint64_t)llabs(d1); e2 += (uint64_t)llabs(d2); e3 += (uint64_t)llabs(d3); e4 += (uint64_t)llabs(d4);
    }
    if (narrow) { e0 &= 0xffffffffu; e1 &= 0xffffffffu; e2 &= 0xffffffffu; e3 &= 0xffffffffu; e4 &= 0xffffffffu; }
    err_out[0] = e0; err_out[1] = e1; err_out[2] = e2; err_out[3] = e3; err_out[4] = e4;
#define MIN2(a,b) ((a) < (b) ? (a) : (b))
    if (e0 <= MIN2(MIN2(MIN2(e1, e2), e3), e4)) return 0;
    if (e1 <= MIN2(MIN2(e2, e3), e4)) return 1;
    if (e2 <= MIN2(e3, e4)) return 2;
    if (e3 <= e4) return 3;
    return 4;
}

/* up: fixed.c FLAC__fixed_compute_best_predictor_limit_residual (subframe_bps >= 28), AS THE SHIPPED x86-64 BINARY RUNS IT
 * (its AVX2 build of the routine; verified with FLAC__stream_encoder_disable_instruction_set):
 *   - sums and order choice agree with the C source for every block length that is 0 or 1 mod 4 (all standard
 *     blocksizes); for lengths 2 or 3 mod 4 (only a short last block can have one) the AVX2 loop reads past the block
 *     and the binary's choice depends on stale memory -- this restatement (and the GPU path) stays with the C sums;
 *   - the C source stores 34.0 as the estimate of every order that does not take the lead, which makes an all-zero
 *     block FIXED order 0; the AVX2 build gives every valid order the total_error_0 estimate, so an all-zero block is
 *     CONSTANT.  The binary is what pyFLAC users run, so that is what is restated here.
 * Unlike the plain variant the
 * sums run over ALL samples (order k from sample k on), an order whose residual does not fit int32 anywhere
 * (|r| > INT32_MAX) is invalid, the winner is the first strictly smallest valid total, and -- 1.4.3's
 * CHECK_ORDER_IS_VALID macro -- the bits-per-sample estimate stored for an order that takes the lead is computed from
 * total_error_0.  sig points at sample 0. */
static uint32_t fixed_best_predictor_limit(const int64_t *sig, uint32_t N, int c_source, uint64_t err_out[5], float rbps[5])
{
    uint64_t tot[5] = {0, 0, 0, 0, 0}, smallest = UINT64_MAX;
    int valid[5] = {1, 1, 1, 1, 1};
    uint32_t order = 0;
    const double len = (double)(N - MAX_FIXED_ORDER);
    for (uint32_t n = 0; n < N; n++) {
        int64_t x0 = sig[n], x1 = n >= 1 ? sig[n - 1] : 0, x2 = n >= 2 ? sig[n - 2] : 0, x3 = n >= 3 ? sig[n - 3] : 0, x4 = n >= 4 ? sig[n - 4] : 0;
        uint64_t e[5];
        e[0] = (uint64_t)llabs(x0);
        e[1] = n >= 1 ? (uint64_t)llabs(x0 - x1) : 0;
        e[2] = n >= 2 ? (uint64_t)llabs(x0 - 2 * x1 + x2) : 0;
        e[3] = n >= 3 ? (uint64_t)llabs(x0 - 3 * x1 + 3 * x2 - x3) : 0;
        e[4] = n >= 4 ? (uint64_t)llabs(x0 - 4 * x1 + 6 * x2 - 4 * x3 + x4) : 0;
        for (int k = 0; k < 5; k++) { tot[k] += e[k]; if (e[k] > (uint64_t)INT32_MAX) valid[k] = 0; }
    }
    for (int k = 0; k < 5; k++) {
        if (valid[k] && tot[k] < smallest) { order = (uint32_t)k; smallest = tot[k]; }
        /* pinned against the binary: every VALID order gets the estimate of total_error_0 (so [1] == 0.0, the
         * constant-subframe trigger, happens exactly for an all-zero block: a non-zero constant block comes out as
         * FIXED order 1), an invalid order gets 34.0 (>= any subframe_bps: "don't even try") */
        if (c_source)   /* C source (the 33-bit routine has no SIMD build): only an order that takes the lead gets the estimate */
            rbps[k] = (smallest == tot[k] && order == (uint32_t)k && valid[k]) ? (float)((tot[0] > 0) ? log(M_LN2 * (double)tot[0] / len) / M_LN2 : 0.0) : 34.0f;
        else
            rbps[k] = valid[k] ? (float)((tot[0] > 0) ? log(M_LN2 * (double)tot[0] / len) / M_LN2 : 0.0) : 34.0f;
        err_out[k] = tot[k];
    }
    return order;
}

/* up: fixed.c FLAC__fixed_compute_residual. data points at sample `order`. */
static void fixed_residual(const int64_t *data, uint32_t n, uint32_t order, int32_t *res)
{
    for (uint32_t i = 0; i < n; i++) {
        const int64_t *d = data + i;
        int64_t r;
        switch (order) {
            case 0: r = d[0]; break;
            case 1: r = (int64_t)d[0] - d[-1]; break;
            case 2: r = (int64_t)d[0] - 2 * (int64_t)d[-1] + d[-2]; break;
            case 3: r = (int64_t)d[0] - 3 * (int64_t)d[-1] + 3 * (int64_t)d[-2] - d[-3]; break;
            default: r = (int64_t)d[0] - 4 * (int64_t)d[-1] + 6 * (int64_t)d[-2] - 4 * (int64_t)d[-3] + d[-4]; break;
        }
        res[i] = (int32_t)r;
    }
}

/* ------------------------------------------------------------------ Rice partition search ---- */

/* up: stream_encoder.c precompute_partition_info_sums_ + set_partitioned_rice_ + count_rice_bits_in_partition_
 * + find_best_partition_order_ (SV A.8, E10).  Returns the estimated residual bits. */
static uint32_t find_best_partition_order(const int32_t *res, uint32_t residual_samples, uint32_t pred_order,
                                          uint32_t rice_limit, uint32_t max_po, uint32_t bps, fo_subframe *sf)
{
    const uint32_t blocksize = residual_samples + pred_order;
    uint64_t *sums;
    uint32_t best_bits = 0, best_po = 0;
    uint32_t params[2][1u << FO_MAX_PART_ORDER];
    int best_idx = 0;

    /* up: format.c FLAC__format_get_max_rice_partition_order_from_blocksize_limited_max_and_predictor_order */
    while (max_po > 0 && (blocksize >> max_po) <= pred_order) max_po--;

    sums = (uint64_t *)malloc(sizeof(uint64_t) * (2u << max_po));
    {   /* sums at max order (32-bit accumulator when libFLAC's threshold test says it cannot overflow) */
        const uint32_t dps = blocksize >> max_po; uint32_t parts = 1u << max_po;
        const uint32_t threshold = 32 - ilog2_u32(dps);
        const int narrow = bps + MAX_EXTRA_RESIDUAL_BPS < threshold;
        uint32_t p, rs = 0, end = (uint32_t)(-(int)pred_order);
        for (p = 0; p < parts; p++) {
            uint64_t s = 0;
            end += dps;
            for (; rs < end; rs++) s += (uint64_t)llabs((int64_t)res[rs]);
            sums[p] = narrow ? (s & 0xffffffffu) : s;
        }
        /* merge downwards */
        uint32_t from = 0, to = parts; int po;
        for (po = (int)max_po - 1; po >= 0; po--) {
            parts >>= 1;
            for (uint32_t i = 0; i < parts; i++) { sums[to++] = sums[from] + sums[from + 1]; from += 2; }
        }
    }
    {
        int po; uint32_t sum_off = 0;
        for (po = (int)max_po; po >= 0; po--) {
            /* set_partitioned_rice_ */
            const uint32_t parts = 1u << po;
            const uint32_t psb = blocksize >> po;
            const uint32_t div_base = 0x40000 / psb;
            uint32_t bits_ = 2 + 4, best_rice = 0;
            uint32_t *par = params[!best_idx];
            int ok = 1;
            for (uint32_t p = 0; p < parts; p++) {
                uint32_t ps = psb, div, k, pbits, best_pbits = UINT32_MAX;
                uint64_t mean = sums[sum_off + p], v;
                if (p > 0) div = div_base;
                else {
                    if (ps <= pred_order) { ok = 0; break; }
                    ps -= pred_order;
                    div = 0x40000 / ps;
                }
                if (mean < 2 || (((mean - 1) * div) >> 18) == 0) k = 0;
                else k = ilog2_u64(((mean - 1) * div) >> 18) + 1;
                if (k >= rice_limit) k = rice_limit - 1;
                v = (uint64_t)4 + (uint64_t)((1 + k) * ps) + (k ? (mean >> (k - 1)) : (mean << 1)) - (uint64_t)(ps >> 1);
                pbits = v < UINT32_MAX ? (uint32_t)v : UINT32_MAX;
                if (pbits < best_pbits) { best_rice = k; best_pbits = pbits; }
                par[p] = best_rice;
                if (best_pbits < UINT32_MAX - bits_) bits_ += best_pbits; else bits_ = UINT32_MAX;
            }
            if (!ok) break;
            sum_off += parts;
            if (best_bits == 0 || bits_ < best_bits) { best_bits = bits_; best_idx = !best_idx; best_po = (uint32_t)po; }
        }
    }
    sf->partition_order = (int32_t)best_po;
    sf->rice2 = 0;
    for (uint32_t p = 0; p < (1u << best_po); p++) {
        sf->rice[p] = params[best_idx][p];
        if (sf->rice[p] >= RICE_ESC) sf->rice2 = 1;
    }
    free(sums);
    return best_bits;
}

/* ------------------------------------------------------------------ subframe search ---- */

typedef struct {
    settings_t s;
    uint32_t N;                        /* current blocksize */
    float *window;                     /* N floats */
    float *windowed;                   /* N floats */
    int32_t *res[2];                   /* residual workspaces */
    uint32_t max_part_order;           /* min(level max, ctz(N)) */
    /* loose mid/side state */
    uint32_t loose_count; int last_ca;
} enc_t;

static uint32_t add_sat(uint32_t est, uint32_t bits) { return bits < UINT32_MAX - est ? est + bits : UINT32_MAX; }

/* up: stream_encoder.c process_subframe_ + evaluate_{verbatim,constant,fixed,lpc}_subframe_ + apply_apodization_
 * (SV A.4-A.9).  sig is modified only by the caller (wasted bits).  Returns the best residual workspace index. */
static int process_subframe(enc_t *e, const int64_t *sig, uint32_t sbps, uint32_t wasted, int disable_constant,
                            fo_subframe *best, int32_t **best_res, fo_signal_trace *tr)
{
    const uint32_t N = e->N;
    const uint32_t rice_limit = e->s.bps > 16 ? RICE2_ESC : RICE_ESC;
    fo_subframe cand;
    int cur = 0; /* best residual lives in e->res[cur] */
    uint32_t best_bits;

    memset(best, 0, sizeof *best);
    best->type = 1; best->wasted = (int32_t)wasted; best->sbps = (int32_t)sbps;
    best_bits = 8 + wasted + N * sbps;                                   /* evaluate_verbatim_subframe_ */
    best->bits_est = best_bits;

    if (N > MAX_FIXED_ORDER) {
        uint64_t ferr[5];
        /* accumulator width: up: process_subframe_ "subframe_bps + ilog2(blocksize-4)+1 < 32" */
        const int narrow = sbps + ilog2_u32(N - MAX_FIXED_ORDER) + 1 < 32;
        float rbps_hi[5] = {0, 0, 0, 0, 0};
        const int hi = sbps >= 28;                                       /* up: process_subframe_ picks the _limit_residual variant */
        uint32_t forder = hi ? fixed_best_predictor_limit(sig, N, 0, ferr, rbps_hi)   /* the binary treats the 33-bit side channel the same way (pinned) */
                             : fixed_best_predictor(sig + MAX_FIXED_ORDER, N - MAX_FIXED_ORDER, narrow, ferr);
        int constant = 0;
        if (tr) { memcpy(tr->fixed_err, ferr, sizeof ferr); tr->fixed_order = (int32_t)forder; }
        /* fixed_residual_bits_per_sample[1] == 0.0  <=>  E1 == 0 (the log term cannot hit exactly 0 and still be constant);
         * the _limit_residual variant yields 0.0 for order 1 exactly when total_error_0 == 0 (all-zero block) */
        if (!disable_constant && (hi ? rbps_hi[1] == 0.0f : ferr[1] == 0)) {
            constant = 1;
            for (uint32_t i = 1; i < N; i++) if (sig[0] != sig[i]) { constant = 0; break; }
        }
        if (tr) tr->is_constant = constant;
        if (constant) {
            uint32_t bits = 8 + wasted + sbps;                           /* evaluate_constant_subframe_ */
            if (bits < best_bits) { best->type = 0; best->bits_est = best_bits = bits; }
        } else {
            /* ---- fixed (guess order only; "don't even try" test is dead for the guessed order) ---- */
            {
                float fbits = hi ? rbps_hi[forder]
                                 : (float)(ferr[forder] > 0 ? log(M_LN2 * (double)ferr[forder] / (double)(N - MAX_FIXED_ORDER)) / M_LN2 : 0.0);
                uint32_t fo = forder;
                if (fo >= N) fo = N - 1;
                if (!(fbits >= (float)sbps)) {
                    memset(&cand, 0, sizeof cand);
                    cand.type = 2; cand.order = (int32_t)fo; cand.wasted = (int32_t)wasted; cand.sbps = (int32_t)sbps;
                    fixed_residual(sig + fo, N - fo, fo, e->res[!cur]);
                    uint32_t rb = find_best_partition_order(e->res[!cur], N - fo, fo, rice_limit, e->max_part_order, sbps, &cand);
                    cand.bits_est = add_sat(8 + wasted + fo * sbps, rb);
                    if (tr) tr->fixed_bits = cand.bits_est;
                    if (cand.bits_est < best_bits) { *best = cand; best_bits = cand.bits_est; cur = !cur; }
                }
            }
            /* ---- LPC ---- */
            if (e->s.max_lpc_order > 0) {
                uint32_t max_lpc = e->s.max_lpc_order >= N ? N - 1 : e->s.max_lpc_order;
                if (max_lpc > 0) {
                    /* apply_apodization_ state machine: a = apodization index (only one), b = depth, c = part */
                    uint32_t a = 0, b = 1, c = 0; int step = 0;
                    double autoc[MAX_LPC_ORDER + 1], autoc_root[MAX_LPC_ORDER + 1], lpc_error[MAX_LPC_ORDER];
                    float lp[MAX_LPC_ORDER][MAX_LPC_ORDER];
                    const int subdivide = e->s.apod_parts > 1;
                    memset(autoc, 0, sizeof autoc); memset(autoc_root, 0, sizeof autoc_root);
                    while (a < 1) {
                        uint32_t max_this = max_lpc, guess;
                        int have = 1;
                        if (b == 1) {
                            for (uint32_t i = 0; i < N; i++) e->windowed[i] = (float)sig[i] * e->window[i];   /* FLAC__lpc_window_data */
                            autocorrelation(e->windowed, N, max_this + 1, autoc);
                            if (subdivide) { memcpy(autoc_root, autoc, max_this * sizeof(double)); b++; }
                            else a++;
                        } else {
                            if (N / b <= MAX_LPC_ORDER) have = 0;
                            else if (!(c % 2)) {
                                /* FLAC__lpc_window_data_partial(in, window, out, N, part_size = N/b/2, data_shift = (c/2*N)/b) */
                                const uint32_t part = N / b / 2, dshift = (c / 2 * N) / b;
                                if (part + dshift < N) {
                                    uint32_t i, j;
                                    for (i = 0; i < part; i++) e->windowed[i] = (float)sig[dshift + i] * e->window[i];
                                    if (i > N - part - dshift) i = N - part - dshift;
                                    for (j = N - part; j < N; i++, j++) e->windowed[i] = (float)sig[dshift + i] * e->window[j];
                                    if (i < N) e->windowed[i] = 0.0f;
                                }
                                autocorrelation(e->windowed, N / b, max_this + 1, autoc);
                            } else {
                                /* punch-out: root minus previous partial, i < max order only (1.4.3 off-by-one, SV A.5) */
                                for (uint32_t i = 0; i < max_this; i++) autoc[i] = autoc_root[i] - autoc[i];
                            }
                            /* set_next_subdivide_tukey */
                            if (b == 2) { if (c == 0) c = 2; else { c = 0; b++; } }
                            else if (c < 2 * b - 1) c++;
                            else { c = 0; b++; }
                            if (b > (uint32_t)e->s.apod_parts) { a++; b = 1; c = 0; }
                        }
                        if (tr && step < FO_MAX_APOD_STEPS) { tr->lpc_order[step] = 0; tr->lpc_bits[step] = 0; }
                        if (!have) continue;
                        if (tr && step < FO_MAX_APOD_STEPS) memcpy(tr->autoc[step], autoc, sizeof(double) * (max_this + 1));
                        if (autoc[0] == 0.0) { step++; continue; }
                        max_this = lp_coefficients(autoc, max_this, lp, lpc_error);
                        guess = best_lpc_order(lpc_error, max_this, N, sbps + e->s.qlp_precision);
                        if (tr && step < FO_MAX_APOD_STEPS) { memcpy(tr->lpc_err[step], lpc_error, sizeof(double) * max_this); tr->lpc_order[step] = (int32_t)guess; }
                        {
                            const uint32_t order = guess;
                            double rbps = expected_bits_per_sample(lpc_error[order - 1], 0.5 / (double)(N - order));
                            if (!(rbps >= (double)sbps)) {
                                /* evaluate_lpc_subframe_ */
                                uint32_t prec = e->s.qlp_precision; int shift;
                                memset(&cand, 0, sizeof cand);
                                if (sbps <= 17) { uint32_t lim = 32 - sbps - ilog2_u32(order); if (lim < prec) prec = lim; }
                                if (quantize_coefficients(lp[order - 1], order, prec, cand.qlp, &shift) == 0) {
                                    int okres;
                                    if (max_residual_bps(sbps, cand.qlp, order, shift) > 32)
                                        okres = lpc_residual(sig + order, N - order, cand.qlp, order, shift, 1, e->res[!cur]);
                                    else
                                        okres = lpc_residual(sig + order, N - order, cand.qlp, order, shift, 0, e->res[!cur]);
                                    if (okres) {
                                        cand.type = 3; cand.order = (int32_t)order; cand.wasted = (int32_t)wasted; cand.sbps = (int32_t)sbps;
                                        cand.precision = (int32_t)prec; cand.shift = shift;
                                        uint32_t rb = find_best_partition_order(e->res[!cur], N - order, order, rice_limit, e->max_part_order, sbps, &cand);
                                        cand.bits_est = add_sat(8 + wasted + 4 + 5 + order * (prec + sbps), rb);
                                        if (tr && step < FO_MAX_APOD_STEPS) tr->lpc_bits[step] = cand.bits_est;
                                        if (cand.bits_est > 0 && cand.bits_est < best_bits) { *best = cand; best_bits = cand.bits_est; cur = !cur; }
                                    }
                                }
                            }
                        }
                        step++;
                    }
                    if (tr) tr->n_apod = step;
                }
            }
        }
    }
    *best_res = e->res[cur];
    /* keep the winner's residual out of the way of the next call: swap so res[0] is free again */
    if (cur == 1) { int32_t *t = e->res[0]; e->res[0] = e->res[1]; e->res[1] = t; }
    return 0;
}

/* up: stream_encoder.c get_wasted_bits_ (SV A.3) */
static uint32_t wasted_bits(int64_t *sig, uint32_t n)
{
    uint32_t i, shift; int64_t x = 0;
    for (i = 0; i < n && !(x & 1); i++) x |= sig[i];
    if (x == 0) shift = 0;
    else for (shift = 0; !(x & 1); shift++) x >>= 1;
    if (shift > 0) for (i = 0; i < n; i++) sig[i] >>= shift;
    return shift;
}

/* ------------------------------------------------------------------ framing ---- */

/* up: stream_encoder_framing.c FLAC__frame_add_header (SV Appendix B; ref: format.h:416-462) */
static void write_frame_header(bw_t *w, const settings_t *s, uint32_t N, uint32_t frame_number, int ca)
{
    uint32_t u, bs_hint = 0, sr_hint = 0;
    size_t start = w->bits >> 3;
    bw_put(w, 0x3ffe, 14); bw_put(w, 0, 1); bw_put(w, 0, 1);
    switch (N) {
        case 192: u = 1; break; case 576: u = 2; break; case 1152: u = 3; break; case 2304: u = 4; break; case 4608: u = 5; break;
        case 256: u = 8; break; case 512: u = 9; break; case 1024: u = 10; break; case 2048: u = 11; break; case 4096: u = 12; break;
        case 8192: u = 13; break; case 16384: u = 14; break; case 32768: u = 15; break;
        default: bs_hint = u = (N <= 0x100) ? 6 : 7; break;
    }
    bw_put(w, u, 4);
    switch (s->sample_rate) {
        case 88200: u = 1; break; case 176400: u = 2; break; case 192000: u = 3; break; case 8000: u = 4; break;
        case 16000: u = 5; break; case 22050: u = 6; break; case 24000: u = 7; break; case 32000: u = 8; break;
        case 44100: u = 9; break; case 48000: u = 10; break; case 96000: u = 11; break;
        default:
            if (s->sample_rate <= 255000 && s->sample_rate % 1000 == 0) sr_hint = u = 12;
            else if (s->sample_rate <= 655350 && s->sample_rate % 10 == 0) sr_hint = u = 14;
            else if (s->sample_rate <= 0xffff) sr_hint = u = 13;
            else u = 0;
            break;
    }
    bw_put(w, u, 4);
    switch (ca) { case 0: u = s->channels - 1; break; case 1: u = 8; break; case 2: u = 9; break; default: u = 10; break; }
    bw_put(w, u, 4);
    switch (s->bps) { case 8: u = 1; break; case 12: u = 2; break; case 16: u = 4; break; case 20: u = 5; break; case 24: u = 6; break; case 32: u = 7; break; default: u = 0; break; }
    bw_put(w, u, 3); bw_put(w, 0, 1);
    bw_utf8(w, frame_number);
    if (bs_hint) bw_put(w, N - 1, bs_hint == 6 ? 8 : 16);
    switch (sr_hint) {
        case 12: bw_put(w, s->sample_rate / 1000, 8); break;
        case 13: bw_put(w, s->sample_rate, 16); break;
        case 14: bw_put(w, s->sample_rate / 10, 16); break;
        default: break;
    }
    bw_put(w, fo_crc8(w->buf + start, (w->bits >> 3) - start), 8);
}

/* up: stream_encoder_framing.c FLAC__subframe_add_{constant,verbatim,fixed,lpc} + add_residual_partitioned_rice_ */
static void write_subframe(bw_t *w, const fo_subframe *sf, const int64_t *sig, const int32_t *res, uint32_t N)
{
    const uint32_t wflag = sf->wasted ? 1 : 0, sbps = (uint32_t)sf->sbps, order = (uint32_t)sf->order;
    switch (sf->type) {
        case 0: bw_put(w, 0x00 | wflag, 8); break;
        case 1: bw_put(w, 0x02 | wflag, 8); break;
        case 2: bw_put(w, 0x10 | (order << 1) | wflag, 8); break;
        default: bw_put(w, 0x40 | ((order - 1) << 1) | wflag, 8); break;
    }
    if (sf->wasted) bw_unary(w, (uint32_t)sf->wasted - 1);
    if (sf->type == 0) { bw_put_signed(w, sig[0], sbps); return; }
    if (sf->type == 1) { for (uint32_t i = 0; i < N; i++) bw_put_signed(w, sig[i], sbps); return; }
    for (uint32_t i = 0; i < order; i++) bw_put_signed(w, sig[i], sbps);
    if (sf->type == 3) {
        bw_put(w, (uint32_t)sf->precision - 1, 4);
        bw_put_signed(w, sf->shift, 5);
        for (uint32_t i = 0; i < order; i++) bw_put_signed(w, sf->qlp[i], (uint32_t)sf->precision);
    }
    bw_put(w, sf->rice2 ? 1 : 0, 2);
    bw_put(w, (uint32_t)sf->partition_order, 4);
    {
        const uint32_t plen = sf->rice2 ? 5 : 4, parts = 1u << sf->partition_order, dps = N >> sf->partition_order;
        uint32_t k = 0;
        for (uint32_t p = 0; p < parts; p++) {
            uint32_t n = dps; if (p == 0) n -= order;
            bw_put(w, sf->rice[p], plen);
            for (uint32_t i = 0; i < n; i++) bw_rice(w, res[k + i], sf->rice[p]);
            k += n;
        }
    }
}

/* up: stream_encoder.c process_subframes_ (SV A.3, A.9, E2) + process_frame_ tail (zero pad, CRC-16) */
static int encode_frame(enc_t *e, int64_t *sigs[], uint32_t frame_number, bw_t *w, fo_frame_trace *tr)
{
    const settings_t *s = &e->s;
    const uint32_t N = e->N, ch = s->channels;
    fo_subframe best[FO_MAX_CHANNELS + 2];
    int32_t *bres[FO_MAX_CHANNELS + 2];
    int32_t *resbuf[FO_MAX_CHANNELS + 2];
    uint32_t sbps[FO_MAX_CHANNELS + 2], wst[FO_MAX_CHANNELS + 2];
    int do_indep, do_ms, ca = 0, all_const = 1, disable_const = 0;
    int64_t *mid = sigs[ch], *side = sigs[ch + 1];

    if (s->do_ms) {
        if (s->loose) {
            if (e->loose_count == 0) { do_indep = 1; do_ms = 1; }
            else { do_indep = (e->last_ca == 0); do_ms = !do_indep; }
        } else { do_indep = 1; do_ms = 1; }
    } else { do_indep = 1; do_ms = 0; }

    if (do_ms) for (uint32_t i = 0; i < N; i++) { side[i] = sigs[0][i] - sigs[1][i]; mid[i] = (sigs[0][i] + sigs[1][i]) >> 1; }
    if (do_indep) for (uint32_t c = 0; c < ch; c++) { uint32_t wb = wasted_bits(sigs[c], N); if (wb > s->bps) wb = s->bps; wst[c] = wb; sbps[c] = s->bps - wb; }
    if (do_ms) for (uint32_t c = 0; c < 2; c++) { uint32_t wb = wasted_bits(sigs[ch + c], N);
        /* up: get_wasted_bits_wide_ (33-bit side of 32-bit input): an all-zero side reports ONE wasted bit, which moves it
         * onto the 32-bit paths (pinned: the binary writes CONSTANT, wasted = 1, 32-bit zero) */
        if (c == 1 && s->bps == 32 && wb == 0) { int allz = 1; for (uint32_t i = 0; i < N; i++) if (sigs[ch + 1][i]) { allz = 0; break; } if (allz) wb = 1; }
        if (wb > s->bps) wb = s->bps;
        wst[ch + c] = wb; sbps[ch + c] = s->bps - wb + (c ? 1 : 0); }

    /* every signal gets its own residual buffer so the winner survives the next call */
    for (uint32_t c = 0; c < ch + 2; c++) resbuf[c] = 0;
    if (do_indep) for (uint32_t c = 0; c < ch; c++) {
        if (s->limit_min_bitrate && all_const && c + 1 == ch) disable_const = 1;
        process_subframe(e, sigs[c], sbps[c], wst[c], disable_const, &best[c], &bres[c], tr ? &tr->sig[c] : 0);
        resbuf[c] = (int32_t *)malloc(sizeof(int32_t) * N); memcpy(resbuf[c], bres[c], sizeof(int32_t) * N); bres[c] = resbuf[c];
        if (best[c].type != 0) all_const = 0;
        if (tr) { tr->sig[c].wasted = (int32_t)wst[c]; tr->sig[c].sbps = (int32_t)sbps[c]; tr->sig[c].best = best[c]; }
    }
    if (do_ms) for (uint32_t c = ch; c < ch + 2; c++) {
        process_subframe(e, sigs[c], sbps[c], wst[c], disable_const, &best[c], &bres[c], tr ? &tr->sig[c] : 0);
        resbuf[c] = (int32_t *)malloc(sizeof(int32_t) * N); memcpy(resbuf[c], bres[c], sizeof(int32_t) * N); bres[c] = resbuf[c];
        if (tr) { tr->sig[c].wasted = (int32_t)wst[c]; tr->sig[c].sbps = (int32_t)sbps[c]; tr->sig[c].best = best[c]; }
    }

    if (do_ms) {
        if (s->loose && e->loose_count > 0) ca = (e->last_ca == 0) ? 0 : 3;
        else {
            uint32_t bits[4], minb; int k;
            bits[0] = best[0].bits_est + best[1].bits_est;
            bits[1] = best[0].bits_est + best[ch + 1].bits_est;
            bits[2] = best[1].bits_est + best[ch + 1].bits_est;
            bits[3] = best[ch].bits_est + best[ch + 1].bits_est;
            ca = 0; minb = bits[0];
            for (k = s->loose ? 3 : 1; k <= 3; k++) if (bits[k] < minb) { minb = bits[k]; ca = k; }
        }
    }
    if (tr) { tr->blocksize = N; tr->frame_number = frame_number; tr->channel_assignment = ca; tr->n_signals = (int32_t)(ch + (do_ms ? 2 : 0)); }

    write_frame_header(w, s, N, frame_number, ca);
    if (do_ms) {
        int l, r;
        switch (ca) { case 0: l = 0; r = 1; break; case 1: l = 0; r = (int)ch + 1; break; case 2: l = (int)ch + 1; r = 1; break; default: l = (int)ch; r = (int)ch + 1; break; }
        write_subframe(w, &best[l], sigs[l], bres[l] , N);
        write_subframe(w, &best[r], sigs[r], bres[r], N);
    } else {
        for (uint32_t c = 0; c < ch; c++) write_subframe(w, &best[c], sigs[c], bres[c], N);
    }
    for (uint32_t c = 0; c < ch + 2; c++) free(resbuf[c]);
    if (w->bits & 7) bw_put(w, 0, 8 - (unsigned)(w->bits & 7));

    if (s->loose) { e->loose_count++; if (e->loose_count >= s->loose_frames) e->loose_count = 0; }
    e->last_ca = ca;
    return 0;
}

/* ------------------------------------------------------------------ stream ---- */

static void put_streaminfo(uint8_t *p, const settings_t *s, uint32_t min_fs, uint32_t max_fs, uint64_t total, const uint8_t md5[16])
{
    /* ref: format.h:546-557 ; SV Appendix B */
    bw_t w = { p, 34, 0, 0 };
    bw_put(&w, s->blocksize, 16); bw_put(&w, s->blocksize, 16);
    bw_put(&w, min_fs, 24); bw_put(&w, max_fs, 24);
    bw_put(&w, s->sample_rate, 20); bw_put(&w, s->channels - 1, 3); bw_put(&w, s->bps - 1, 5);
    bw_put(&w, total, 36);
    for (int i = 0; i < 16; i++) bw_put(&w, md5[i], 8);
}

long fo_encode_stream(const fo_enc_cfg *cfg, const int32_t *pcm, uint64_t nsamples,
                      uint8_t *out, size_t out_cap,
                      uint64_t *frame_off, uint32_t *frame_len, uint32_t frames_cap, uint32_t *nframes_out,
                      fo_frame_trace *traces, uint32_t traces_cap)
{
    enc_t e; settings_t *s = &e.s;
    uint8_t md5zero[16] = {0}, digest[16];
    md5_t md5;
    size_t pos = 0;
    uint32_t nframes = 0, min_fs = 0, max_fs = 0, frame_number = 0, cur_window_N = 0;
    int64_t *sigs[FO_MAX_CHANNELS + 2];
    uint8_t *fbuf; size_t fcap;

    if (fo_encoder_init_status(cfg, 1, 0, 0) != 0) return -1;
    if (cfg->bps > 32) return -3;
    memset(&e, 0, sizeof e);
    resolve_settings(cfg, s);
    {
        const uint32_t B = s->blocksize;
        for (uint32_t c = 0; c < s->channels + 2; c++) sigs[c] = (int64_t *)malloc(sizeof(int64_t) * (B + 8));
        e.window = (float *)malloc(sizeof(float) * B); e.windowed = (float *)malloc(sizeof(float) * (B + 8));
        e.res[0] = (int32_t *)malloc(sizeof(int32_t) * (B + 8)); e.res[1] = (int32_t *)malloc(sizeof(int32_t) * (B + 8));
        fcap = (size_t)B * (s->channels) * 5 + 1024; fbuf = (uint8_t *)malloc(fcap);
    }

    /* stream prologue: "fLaC" + STREAMINFO + VORBIS_COMMENT(vendor) -- 4 + 38 + 44 bytes (SV 3.1, Appendix B) */
    if (out_cap < 86) return -2;
    memcpy(out, "fLaC", 4);
    out[4] = 0x00; out[5] = 0; out[6] = 0; out[7] = 34;
    put_streaminfo(out + 8, s, 0, 0, 0, md5zero);
    {
        const uint32_t vl = (uint32_t)strlen(VENDOR), len = 4 + vl + 4;
        uint8_t *p = out + 42;
        p[0] = 0x84; p[1] = (uint8_t)(len >> 16); p[2] = (uint8_t)(len >> 8); p[3] = (uint8_t)len;
        p[4] = (uint8_t)vl; p[5] = (uint8_t)(vl >> 8); p[6] = (uint8_t)(vl >> 16); p[7] = (uint8_t)(vl >> 24);
        memcpy(p + 8, VENDOR, vl);
        memset(p + 8 + vl, 0, 4);
        pos = 42 + 4 + len;
    }

    md5_init(&md5);
    for (uint64_t done = 0; done < nsamples;) {
        uint32_t N = (nsamples - done >= s->blocksize) ? s->blocksize : (uint32_t)(nsamples - done);
        const uint32_t ch = s->channels, bytes = (s->bps + 7) / 8;
        bw_t w;
        /* MD5 over (bps+7)/8 little-endian bytes per sample, interleaved (up: md5.c FLAC__MD5Accumulate) */
        for (uint32_t i = 0; i < N; i++) for (uint32_t c = 0; c < ch; c++) {
            int32_t v = pcm[(done + i) * ch + c]; uint8_t b4[4] = { (uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24) };
            md5_update(&md5, b4, bytes);
            sigs[c][i] = v;
        }
        /* short last block: finish() sets blocksize = remainder, windows are rebuilt, max partition order follows (SV A.2) */
        e.N = N;
        if (s->max_lpc_order > 0 && cur_window_N != N) { fo_window_tukey(e.window, (int32_t)N, s->apod_p); cur_window_N = N; }
        { uint32_t po = 0, b = N; while (!(b & 1)) { po++; b >>= 1; } if (po > FO_MAX_PART_ORDER) po = FO_MAX_PART_ORDER; e.max_part_order = po < s->max_part_order ? po : s->max_part_order; }
        w.buf = fbuf; w.cap = fcap; w.bits = 0; w.overflow = 0;
        encode_frame(&e, sigs, frame_number, &w, (traces && nframes < traces_cap) ? &traces[nframes] : 0);
        {
            size_t n = w.bits >> 3; uint16_t crc = fo_crc16(fbuf, n);
            if (w.overflow || pos + n + 2 > out_cap) { pos = 0; goto fail; }
            memcpy(out + pos, fbuf, n); out[pos + n] = (uint8_t)(crc >> 8); out[pos + n + 1] = (uint8_t)crc;
            n += 2;
            if (frame_off && nframes < frames_cap) { frame_off[nframes] = pos; frame_len[nframes] = (uint32_t)n; }
            if (min_fs == 0 || n < min_fs) min_fs = (uint32_t)n;
            if (n > max_fs) max_fs = (uint32_t)n;
            pos += n;
        }
        nframes++; frame_number++; done += N;
    }
    md5_final(&md5, digest);
    if (cfg->seekable) put_streaminfo(out + 8, s, min_fs, max_fs, nsamples, digest);
fail:
    for (uint32_t c = 0; c < s->channels + 2; c++) free(sigs[c]);
    free(e.window); free(e.windowed); free(e.res[0]); free(e.res[1]); free(fbuf);
    if (nframes_out) *nframes_out = nframes;
    return pos ? (long)pos : -2;
}

/* ================================================================== decoder ==== */

/* up: bitreader.c / stream_decoder.c (SV D1-D4; ref: format.h:209-475, stream_decoder.h:1440-1513) */
typedef struct { const uint8_t *p; size_t len; size_t bit; int err; } br_t;
static uint64_t br_get(br_t *r, unsigned n)
{
    uint64_t v = 0;
    while (n) {
        size_t byte = r->bit >> 3; unsigned off = (unsigned)(r->bit & 7), room = 8 - off, k = n < room ? n : room;
        if (byte >= r->len) { r->err = 1; return 0; }
        v = (v << k) | ((r->p[byte] >> (room - k)) & ((1u << k) - 1));
        r->bit += k; n -= k;
    }
    return v;
}
static int64_t br_get_signed(br_t *r, unsigned n)
{
    uint64_t v = br_get(r, n);
    if (n == 0) return 0;
    if (n < 64 && (v >> (n - 1))) v |= ~((1ull << n) - 1);
    return (int64_t)v;
}
static uint32_t br_unary(br_t *r) { uint32_t q = 0; while (!r->err && br_get(r, 1) == 0) q++; return q; }

static int decode_residual(br_t *r, int32_t *res, uint32_t N, uint32_t order)
{
    uint32_t method = (uint32_t)br_get(r, 2), po, parts, k = 0;
    if (method > 1) return -3;
    po = (uint32_t)br_get(r, 4); parts = 1u << po;
    if ((N >> po) < order || (po > 0 && (N & (parts - 1)))) return -3;
    for (uint32_t p = 0; p < parts; p++) {
        uint32_t n = (N >> po) - (p == 0 ? order : 0);
        uint32_t par = (uint32_t)br_get(r, method ? 5 : 4);
        if (par == (method ? 31u : 15u)) {
            uint32_t raw = (uint32_t)br_get(r, 5);
            for (uint32_t i = 0; i < n; i++) res[k + i] = raw ? (int32_t)br_get_signed(r, raw) : 0;
        } else {
            for (uint32_t i = 0; i < n; i++) {
                uint32_t q = br_unary(r), lo = par ? (uint32_t)br_get(r, par) : 0;
                uint32_t u = (q << par) | lo;
                res[k + i] = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
            }
        }
        k += n;
        if (r->err) return -3;
    }
    return 0;
}

static int decode_subframe(br_t *r, int64_t *out, uint32_t N, uint32_t bps, int32_t *res)
{
    uint32_t hdr = (uint32_t)br_get(r, 8), wasted = 0, type;
    if (hdr & 0x80) return -3;
    if (hdr & 1) { wasted = br_unary(r) + 1; if (wasted >= bps) return -3; bps -= wasted; }
    type = (hdr >> 1) & 0x3f;
    if (type == 0) { int64_t v = br_get_signed(r, bps); for (uint32_t i = 0; i < N; i++) out[i] = v; }
    else if (type == 1) { for (uint32_t i = 0; i < N; i++) out[i] = br_get_signed(r, bps); }
    else if (type >= 8 && type <= 12) {
        uint32_t order = type - 8;
        if (order > N) return -3;
        for (uint32_t i = 0; i < order; i++) out[i] = br_get_signed(r, bps);
        int rc = decode_residual(r, res, N, order); if (rc) return rc;
        for (uint32_t i = order; i < N; i++) {
            int64_t p;
            switch (order) {
                case 0: p = 0; break;
                case 1: p = out[i-1]; break;
                case 2: p = 2*out[i-1] - out[i-2]; break;
                case 3: p = 3*out[i-1] - 3*out[i-2] + out[i-3]; break;
                default: p = 4*out[i-1] - 6*out[i-2] + 4*out[i-3] - out[i-4]; break;
            }
            out[i] = res[i - order] + p;
        }
    } else if (type >= 32) {
        uint32_t order = type - 31, prec; int shift; int32_t q[32];
        if (order > N) return -3;
        for (uint32_t i = 0; i < order; i++) out[i] = br_get_signed(r, bps);
        prec = (uint32_t)br_get(r, 4) + 1; if (prec == 16) return -3;
        shift = (int)br_get_signed(r, 5); if (shift < 0) return -3;
        for (uint32_t i = 0; i < order; i++) q[i] = (int32_t)br_get_signed(r, prec);
        int rc = decode_residual(r, res, N, order); if (rc) return rc;
        for (uint32_t i = order; i < N; i++) {
            int64_t sum = 0;
            for (uint32_t j = 0; j < order; j++) sum += (int64_t)q[j] * out[i - 1 - j];
            out[i] = res[i - order] + (sum >> shift);
        }
    } else return -3;
    if (wasted) for (uint32_t i = 0; i < N; i++) out[i] = (int64_t)((uint64_t)out[i] << wasted);
    return r->err ? -3 : 0;
}

long fo_decode_stream(const uint8_t *in, size_t in_len, int32_t *out, uint64_t out_cap, uint32_t *info)
{
    size_t pos = 4; int last = 0;
    uint32_t si_sr = 0, si_ch = 0, si_bps = 0, si_minbs = 0, si_maxbs = 0; uint64_t si_total = 0; uint8_t si_md5[16] = {0};
    uint64_t total = 0; md5_t md5; int first = 1;
    int64_t *chbuf[FO_MAX_CHANNELS]; int32_t *res = 0; uint32_t bufN = 0;
    long rc = 0;
    if (in_len < 4 || memcmp(in, "fLaC", 4)) return -1;
    while (!last) {
        if (pos + 4 > in_len) return -3;
        uint32_t type = in[pos] & 0x7f, len = (uint32_t)in[pos+1] << 16 | (uint32_t)in[pos+2] << 8 | in[pos+3];
        last = in[pos] >> 7; pos += 4;
        if (pos + len > in_len) return -3;
        if (type == 0 && len >= 34) {
            br_t r = { in + pos, len, 0, 0 };
            si_minbs = (uint32_t)br_get(&r, 16); si_maxbs = (uint32_t)br_get(&r, 16); br_get(&r, 24); br_get(&r, 24);
            si_sr = (uint32_t)br_get(&r, 20); si_ch = (uint32_t)br_get(&r, 3) + 1; si_bps = (uint32_t)br_get(&r, 5) + 1;
            si_total = br_get(&r, 36); memcpy(si_md5, in + pos + 18, 16);
        }
        pos += len;
    }
    (void)si_minbs;
    memset(chbuf, 0, sizeof chbuf);
    md5_init(&md5);
    while (pos < in_len) {
        br_t r = { in + pos, in_len - pos, 0, 0 };
        uint32_t sync = (uint32_t)br_get(&r, 14), N, sr, ch, ca, bps, bs_code, sr_code, ca_code, bps_code;
        if (sync != 0x3ffe) { rc = -3; break; }
        br_get(&r, 1); uint32_t variable = (uint32_t)br_get(&r, 1);
        bs_code = (uint32_t)br_get(&r, 4); sr_code = (uint32_t)br_get(&r, 4); ca_code = (uint32_t)br_get(&r, 4); bps_code = (uint32_t)br_get(&r, 3);
        if (br_get(&r, 1)) { rc = -3; break; }
        { /* UTF-8 coded number */
            uint32_t b0 = (uint32_t)br_get(&r, 8), extra = 0;
            if (b0 & 0x80) { uint32_t m = 0x40; while (b0 & m) { extra++; m >>= 1; } if (extra == 0 || extra > 6) { rc = -3; break; } }
            for (uint32_t i = 0; i < extra; i++) br_get(&r, 8);
            (void)variable;
        }
        switch (bs_code) {
            case 0: N = 0; break; case 1: N = 192; break;
            case 2: case 3: case 4: case 5: N = 576u << (bs_code - 2); break;
            case 6: N = (uint32_t)br_get(&r, 8) + 1; break; case 7: N = (uint32_t)br_get(&r, 16) + 1; break;
            default: N = 256u << (bs_code - 8); break;
        }
        switch (sr_code) {
            case 0: sr = si_sr; break; case 1: sr = 88200; break; case 2: sr = 176400; break; case 3: sr = 192000; break;
            case 4: sr = 8000; break; case 5: sr = 16000; break; case 6: sr = 22050; break; case 7: sr = 24000; break;
            case 8: sr = 32000; break; case 9: sr = 44100; break; case 10: sr = 48000; break; case 11: sr = 96000; break;
            case 12: sr = (uint32_t)br_get(&r, 8) * 1000; break; case 13: sr = (uint32_t)br_get(&r, 16); break;
            case 14: sr = (uint32_t)br_get(&r, 16) * 10; break; default: sr = 0; break;
        }
        if (N == 0 || sr_code == 15) { rc = -3; break; }
        if (ca_code < 8) { ch = ca_code + 1; ca = 0; } else if (ca_code <= 10) { ch = 2; ca = ca_code - 7; } else { rc = -3; break; }
        switch (bps_code) { case 0: bps = si_bps; break; case 1: bps = 8; break; case 2: bps = 12; break; case 4: bps = 16; break; case 5: bps = 20; break; case 6: bps = 24; break; case 7: bps = 32; break; default: bps = 0; break; }
        if (bps == 0) { rc = -3; break; }
        { uint8_t c8 = (uint8_t)br_get(&r, 8); if (r.err || fo_crc8(in + pos, (r.bit >> 3) - 1) != c8) { rc = -4; break; } }
        if (N > bufN) {
            for (uint32_t c = 0; c < FO_MAX_CHANNELS; c++) { free(chbuf[c]); chbuf[c] = (int64_t *)malloc(sizeof(int64_t) * N); }
            free(res); res = (int32_t *)malloc(sizeof(int32_t) * N); bufN = N;
        }
        for (uint32_t c = 0; c < ch && !rc; c++) {
            uint32_t sb = bps + ((ca == 1 && c == 1) || (ca == 2 && c == 0) || (ca == 3 && c == 1) ? 1 : 0);
            rc = decode_subframe(&r, chbuf[c], N, sb, res);
        }
        if (rc) break;
        if (r.bit & 7) { if (br_get(&r, 8 - (unsigned)(r.bit & 7)) != 0) { rc = -3; break; } }
        { size_t n = r.bit >> 3; uint16_t c16 = (uint16_t)br_get(&r, 16); if (r.err || fo_crc16(in + pos, n) != c16) { rc = -4; break; } }
        /* undo inter-channel decorrelation (ref: format.h:388-393 ; SV D4) */
        for (uint32_t i = 0; i < N; i++) {
            int64_t a = chbuf[0][i], b = ch > 1 ? chbuf[1][i] : 0;
            if (ca == 1) chbuf[1][i] = a - b;
            else if (ca == 2) chbuf[0][i] = a + b;
            else if (ca == 3) { int64_t m = (int64_t)(((uint64_t)a << 1) | (uint64_t)(b & 1)); chbuf[0][i] = (m + b) >> 1; chbuf[1][i] = (m - b) >> 1; }
        }
        if (first) { if (info) { info[0] = ch; info[1] = bps; info[2] = sr; info[3] = N; } first = 0; }
        if (out) {
            if (total + N > out_cap) { rc = -2; break; }
            for (uint32_t i = 0; i < N; i++) for (uint32_t c = 0; c < ch; c++) out[(total + i) * ch + c] = (int32_t)chbuf[c][i];
        }
        { const uint32_t bytes = (bps + 7) / 8;
          for (uint32_t i = 0; i < N; i++) for (uint32_t c = 0; c < ch; c++) {
              int32_t v = (int32_t)chbuf[c][i]; uint8_t b4[4] = { (uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24) };
              md5_update(&md5, b4, bytes);
          } }
        total += N;
        pos += r.bit >> 3;
    }
    for (uint32_t c = 0; c < FO_MAX_CHANNELS; c++) free(chbuf[c]);
    free(res);
    if (rc) return rc;
    (void)si_maxbs; (void)si_ch; (void)si_total;
    {
        uint8_t d[16]; static const uint8_t z[16] = {0};
        md5_final(&md5, d);
        if (memcmp(si_md5, z, 16) && memcmp(si_md5, d, 16)) return -5;
    }
    return (long)total;
}
