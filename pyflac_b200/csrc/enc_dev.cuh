// enc_dev.cuh -- device helpers shared by the encode kernels (enc_analyze.cu, enc_pack.cu, enc_fused.cu).
#pragma once
#include "fb_common.cuh"
#include "fb_math.cuh"

namespace fb {

constexpr int kMaxSteps = kMaxApodSteps;
constexpr int kAcStoreStride = 14;     // lags kept per (signal, window): 13 + the unused odd partner

struct __align__(16) WarpScratch {
    double   ac[16];                   // autocorrelation of the current apodization step
    double   lperr[kMaxOrder];         // Levinson error per order
    double   lpc[kMaxOrder];           // Levinson recursion state
    float    lp[kMaxOrder * kMaxOrder];
    int32_t  q[16];                    // quantised coefficients of the current candidate
    int32_t  misc[8];
};

// ---- libm-log guard: lookup of a host override / logging of a decision inside the band (warp-uniform calls) ----
__device__ __forceinline__ const LogGuardOverride* guard_find(const EncParams& P, const FrameDesc& fd, int s, int step) {
    for (uint32_t i = 0; i < P.guard_n_ovr; i++) {
        const LogGuardOverride* o = P.guard_ovr + i;
        if (o->stream == fd.stream && o->frame_number == fd.frame_number && o->signal == (uint32_t)s && o->step == (uint32_t)step) return o;
    }
    return nullptr;
}
__device__ __forceinline__ void guard_record(const EncParams& P, EncStats* stats, const FrameDesc& fd, int s, int step, int N, int sbps, int max_order,
                                             uint32_t overhead, const double* lperr, int guess, bool skip, uint32_t kind, int lane) {
    if (!stats) return;
    unsigned long long slot = 0;
    if (lane == 0) slot = atomicAdd(&stats->log_ambiguous, 1ull);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (!P.guard_log || slot >= P.guard_cap) return;
    LogGuardEntry* e = P.guard_log + slot;
    if (lane == 0) {
        e->stream = fd.stream; e->frame_number = fd.frame_number; e->signal = (uint32_t)s; e->step = (uint32_t)step;
        e->N = (uint32_t)N; e->sbps = (uint32_t)sbps; e->max_order = (uint32_t)max_order; e->overhead = overhead;
        e->guess = guess; e->skip = skip ? 1 : 0; e->kind = kind; e->pad = 0;
    }
    if (lane < kMaxOrder) e->lperr[lane] = lane < max_order ? lperr[lane] : 0.0;
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// up: stream_encoder.c find_best_partition_order_ / set_partitioned_rice_: given the sums at
// the maximum order in psum[0 .. 2^omax), search orders omax..0 (first strict minimum), merging pairwise.
// Lane p owns partitions p and p+32.  Returns estimated residual bits; best parameters land in k0/k1.
__device__ __forceinline__ uint32_t rice_search(unsigned long long* psum, int N, int pred_order, int omax, bool narrow_sums,
                                                uint32_t rice_limit, int lane, int* best_order_out,
                                                uint32_t* k0_out, uint32_t* k1_out) {
    if (narrow_sums) {      // libFLAC's 32-bit partition accumulators wrap
        for (int p = lane; p < (1 << omax); p += 32) psum[p] &= 0xffffffffull;
        __syncwarp();
    }
    uint32_t best_bits = 0, bk0 = 0, bk1 = 0;
    int best_o = 0, off = 0;
    for (int o = omax; o >= 0; o--) {
        const int parts = 1 << o;
        const uint32_t psb = (uint32_t)N >> o;
        unsigned long long lane_bits = 0;
        uint32_t k0 = 0, k1 = 0;
#pragma unroll
        for (int t = 0; t < 2; t++) {
            const int p = lane + 32 * t;
            if (p < parts) {
                const uint32_t n = psb - (p == 0 ? (uint32_t)pred_order : 0u);
                const unsigned long long s = psum[off + p];
                const uint32_t k = rice_parameter(s, n, rice_limit);
                lane_bits += rice_partition_bits(k, n, s);
                if (t == 0) k0 = k; else k1 = k;
            }
        }
        unsigned long long tot = warp_sum_u64(lane_bits) + 6ull;
        const uint32_t bits = tot < 0xffffffffull ? (uint32_t)tot : 0xffffffffu;
        if (best_bits == 0 || bits < best_bits) { best_bits = bits; best_o = o; bk0 = k0; bk1 = k1; }
        if (o > 0) {
            const int half = parts >> 1;
            for (int p2 = lane; p2 < half; p2 += 32) psum[off + parts + p2] = psum[off + 2 * p2] + psum[off + 2 * p2 + 1];
            off += parts;
        }
        __syncwarp();
    }
    *best_order_out = best_o; *k0_out = bk0; *k1_out = bk1;
    return best_bits;
}

__device__ __forceinline__ uint32_t add_sat(uint32_t est, uint32_t bits) {
    return bits < 0xffffffffu - est ? est + bits : 0xffffffffu;
}

// side33: the signal is the 33-bit side channel of 32-bit stereo (up: get_wasted_bits_wide_): an all-zero side reports ONE
// wasted bit, which moves it onto the 32-bit paths (pinned against the binary, oracle/flac_oracle.c:encode_frame)
__device__ __forceinline__ int wasted_from_or(uint32_t o, int bps, bool side33 = false) {
    const int w = o ? (__ffs((int)o) - 1) : (side33 ? 1 : 0);
    return w > bps ? bps : w;
}

__device__ __forceinline__ void put_bits(uint32_t* buf, uint32_t pos, uint32_t val, uint32_t n) {
    if (n == 0) return;
    if (n < 32) val &= (1u << n) - 1u;
    const uint32_t w = pos >> 5, o = pos & 31;
    if (o + n <= 32) atomicOr(&buf[w], val << (32 - o - n));
    else {
        const uint32_t r = o + n - 32;
        atomicOr(&buf[w], val >> r);
        atomicOr(&buf[w + 1], val << (32 - r));
    }
}

// up to 33 bits (the side channel of 32-bit stereo): top bit, then the low 32
__device__ __forceinline__ void put_bits64(uint32_t* buf, uint32_t pos, long long v, uint32_t n) {
    if (n > 32) { put_bits(buf, pos, (uint32_t)((unsigned long long)v >> 32), n - 32); put_bits(buf, pos + n - 32, (uint32_t)v, 32); }
    else put_bits(buf, pos, (uint32_t)v, n);
}


}  // namespace fb
