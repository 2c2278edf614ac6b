// Host build of pyflac_b200/csrc/fb_math.cuh so the scalar decision arithmetic of the CUDA path can be
// unit-tested on CPU-only machines (tests/test_fb_math_cpu.py). Compiled with -ffp-contract=off.
#include "../../pyflac_b200/csrc/fb_math.cuh"
using namespace fb;
extern "C" {
int t_levinson(const double* ac, int max_order, float* lp, double* err) { double lpc[kMaxOrder]; return levinson(ac, max_order, lp, err, lpc); }
double t_expected_bits(double e, double scale) { bool u; return expected_bits_per_sample(e, scale, &u); }
int t_quantize(const float* lp, int order, int precision, int32_t* q, int* shift) { return quantize_coefficients(lp, order, precision, q, shift); }
uint32_t t_rice_parameter(uint64_t sum, uint32_t n, uint32_t limit) { return rice_parameter(sum, n, limit); }
uint32_t t_rice_bits(uint32_t k, uint32_t n, uint64_t sum) { return rice_partition_bits(k, n, sum); }
int t_frame_header(uint8_t* out, uint32_t ch, uint32_t bps, uint32_t sr, uint32_t N, uint32_t fn, int ca) { return build_frame_header(out, ch, bps, sr, N, fn, ca); }
uint16_t t_crc16_mulmod(uint16_t a, uint16_t b) { return crc16_mulmod(a, b); }
uint16_t t_crc16_byte(uint16_t c, uint8_t b) { return crc16_byte(c, b); }
uint32_t t_silog2(int64_t v) { return silog2(v); }
// CRC-16 of n bytes the way the decoder's dec_crc_kernel combines it: 64-byte chunks counted from the END, each weighted by x^(512 j)
uint16_t t_crc16_chunked(const uint8_t* p, uint32_t n) {
    static const CrcPosTable tab = make_crc_pos_table();
    const uint32_t nchunks = (n + 63u) >> 6;
    uint16_t acc = 0;
    for (uint32_t j = 0; j < nchunks; j++) {
        const uint32_t end = n - (j << 6), beg = end >= 64u ? end - 64u : 0u;
        uint16_t crc = 0;
        for (uint32_t b = beg; b < end; b++) crc = crc16_byte(crc, p[b]);
        acc ^= crc16_weigh_chunk(crc, j, tab);
    }
    return acc;
}
uint16_t t_crc16_plain(const uint8_t* p, uint32_t n) { uint16_t c = 0; for (uint32_t b = 0; b < n; b++) c = crc16_byte(c, p[b]); return c; }
}
