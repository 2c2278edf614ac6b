// enc_analyze.cu -- encode analysis kernel (scope rows E2-E11): one CTA of four warps per frame.
//
// For every signal of the frame (a channel, or mid / side) the kernel reproduces libFLAC 1.4.3's
// process_subframe_ decision sequence: wasted bits, fixed-predictor error sums,
// window + sequential-double autocorrelation, Levinson-Durbin, order guess, coefficient quantisation,
// residual -> partition sums -> Rice parameter / partition-order search, and keeps the candidate with the
// smallest libFLAC bit estimate.  Output: one 128-byte SubframePlan per signal + one channel-assignment
// byte per frame; the pack kernel (enc_pack.cu) turns plans into bits.
//
// Three kernels per batch (B200: this path is latency/issue bound, so each kernel keeps its warps doing one kind of
// work, minimises instructions and keeps many frames resident per SM):
//  * frame_bits_kernel -- OR / AND of every signal of every frame (wasted bits, constant detection): one warp per
//    frame, pure streaming.
//  * autoc_kernel -- the autocorrelations of all (frame, signal, window) jobs.  One DFMA chain per lag in ascending
//    sample order (products of two floats are exact in double, so fma == libFLAC's mul+add).  A lane owns four
//    consecutive lags of one job, a warp carries up to eight jobs (two stereo frames), jobs are independent warps:
//    no barriers, PCM read straight from HBM/L2, a 640-byte ring of windowed doubles per job in shared memory.
//  * analyze_kernel -- one CTA of four warps per frame.  The frame is staged ONCE into shared memory in its
//    container form -- for 16-bit stereo the raw interleaved words (L | R << 16), 16 KiB per 4096-sample frame,
//    mid/side derived on the fly; for every other shape one int32 row block per signal.  Rows of B0 = ceil(N/32)
//    samples with an odd row stride: lane p streams the contiguous samples [p*B0, (p+1)*B0) bank-conflict free, one
//    shared load per sample, predictor history in registers (statically rotated window), partition sums leave the
//    lane through a handful of shared atomics.  Work items (per-signal fixed analysis, then one LPC evaluation per
//    (signal, apodization step)) are handed to the four warps through a shared-memory queue.
//
// Exactness: integer work is exact; floating point follows fb_math.cuh (unfused, RN).  No tensor cores:
// this is integer / bit-serial work, not a dense contraction.
#include <type_traits>
#include "fb_common.cuh"
#include "fb_math.cuh"
#include "enc_dev.cuh"

namespace fb {

constexpr int kAnThreads = 128;
constexpr int kAnWarps = kAnThreads / 32;
constexpr int kAcJobsMax = 8;          // (frame, signal, window) jobs carried by one warp of the autocorrelation kernel
constexpr int kAcRing = 82;            // doubles per job: 16 mirror + 2 slots of 32 (+2: jobs land in different bank groups)
constexpr int kAcThreads = 128;        // autocorrelation kernel: four independent warps per CTA

// Signals of the packed layout (16-bit stereo staged as raw words L | R << 16); plain layouts index their own row block.
enum SigKind : int { kLo16 = 0, kHi16 = 1, kMid16 = 2, kSide16 = 3, kPlain = 4 };

// How a warp reads one signal out of the staged frame.  Packed: value = (lo * ca + hi * cb) >> sh with
// (ca, cb, sh) = (1,0,w) left, (0,1,w) right, (1,1,1+w) mid, (1,-1,w) side, w = wasted bits -- one code path for
// all four signals keeps the instruction working set of a CTA (whose warps run different signals) small.
struct SigView {
    const int32_t* base;     // packed: the frame's words; plain: the signal's own row block (wasted bits already removed)
    int ca, cb, sh;
    int cab;                 // packed: (ca & 0xff) | (cb & 0xff) << 8 for the dot-product instruction
    // 33-bit side channel of 32-bit stereo (no wasted bits): derived from the left (base) and right (base2) rows,
    // value = (left << shl) - (right << shr) in 64-bit arithmetic (the rows hold the channels without THEIR wasted bits)
    const int32_t* base2;
    int shl, shr, s33;
};
__device__ __forceinline__ long long sig_word64(const SigView& V, int p) {
    return V.s33 ? ((long long)V.base[p] << V.shl) - ((long long)V.base2[p] << V.shr) : (long long)V.base[p];
}

struct FrameGeo {
    int N, B0, RS, pad;      // samples, samples per row, row stride in words, RS - B0
    uint32_t magic;          // ceil(2^32 / B0) (pad != 0 only)
};

__device__ __forceinline__ int pidx(const FrameGeo& G, int i) {
    return G.pad ? i + (int)__umulhi((uint32_t)i, G.magic) : i;
}

template <bool PACKED>
__device__ __forceinline__ int sig_word(int w, const SigView& V) {
    if constexpr (PACKED) {
        // one IDP.2A: (low half) * ca + (high half) * cb, the two int16 halves of w against the two int8 coefficients
        return __dp2a_lo(w, V.cab, 0) >> V.sh;
    } else {
        return w;
    }
}

// ------------------------------------------------------------------------------------------------
// r[i] = x[i] - ((sum_j q[j] * x[i-1-j]) >> shift), sum of |r| per partition at the maximum partition order.
// Fixed predictors are the same formula with binomial coefficients and shift 0 (up: fixed.c
// FLAC__fixed_compute_residual == lpc residual with q = {1},{2,-1},{3,-3,1},{4,-6,4,-1}).
// WIDE: 64-bit accumulate (up: FLAC__lpc_compute_residual_from_qlp_coefficients_wide); the 32-bit form is
// used exactly when libFLAC proves it cannot overflow.  check_limit reproduces the _limit_residual rejection.
//
// Lane p walks its own contiguous samples; h[] is the predictor history, rotated statically: inside a group
// of C samples, sample u reads h[(u-1-j) mod C] and then overwrites h[u] (the tap that just expired).
// C is the order CLASS (4, 8 or 12 taps; coefficients beyond the real order are zero): three code bodies instead of
// thirteen keep the instruction working set of an SM -- whose warps run different orders at the same time -- inside
// the instruction cache (the per-order version spent half its stall samples waiting for instruction fetch).
// psum must be zeroed by the caller; partition totals arrive through shared atomics (<= 3 per lane).
template <int C, bool WIDE, bool PACKED, bool S33 = false>
__device__ __noinline__ bool residual_partition_sums(SigView V, FrameGeo G, const int32_t* __restrict__ qs, int order, int shift,
                                                     int psize, bool check_limit, unsigned long long* psum, int lane) {
    using HistT = typename std::conditional<S33, long long, int32_t>::type;      // S33: 33-bit samples (always WIDE, plain layout)
    int32_t q[C];
#pragma unroll
    for (int j = 0; j < C; j++) q[j] = (j < order) ? qs[j] : 0;
    bool bad = false;
    const int blk_lo = lane * G.B0, blk_hi = min(G.N, blk_lo + G.B0);
    int lo = max(blk_lo, order);
    const int32_t* rowp = V.base + lane * G.RS - blk_lo;      // rowp[i] is sample i for blk_lo <= i < blk_hi
    while (lo < blk_hi) {
        const int part = lo / psize;
        const int hi = min(blk_hi, (part + 1) * psize);
        HistT h[C];
#pragma unroll
        for (int k = 0; k < C; k++) {       // taps before sample 0 carry zero coefficients
            if (S33) h[k] = (HistT)sig_word64(V, pidx(G, max(lo - C + k, 0)));
            else h[k] = (HistT)sig_word<PACKED>(V.base[pidx(G, max(lo - C + k, 0))], V);
        }
        unsigned long long acc = 0;
        for (int g = lo; g < hi; g += C) {
#pragma unroll
            for (int u = 0; u < C; u++) {
                if (g + u < hi) {
                    HistT xv;
                    if (S33) xv = (HistT)sig_word64(V, lane * G.RS - blk_lo + g + u);
                    else xv = (HistT)sig_word<PACKED>(rowp[g + u], V);
                    long long r;
                    if (WIDE) {
                        long long s = 0;
#pragma unroll
                        for (int j = 0; j < C; j++) s += (long long)q[j] * (long long)h[(u - 1 - j + 2 * C) % C];
                        r = (long long)xv - (s >> shift);
                        if (check_limit && (r <= (long long)INT32_MIN || r > (long long)INT32_MAX)) bad = true;
                    } else {
                        int s = 0;
#pragma unroll
                        for (int j = 0; j < C; j++) s += q[j] * (int)h[(u - 1 - j + 2 * C) % C];
                        r = (long long)((int)xv - (s >> shift));
                    }
                    acc += (unsigned long long)(r < 0 ? -r : r);
                    h[u] = xv;
                }
            }
        }
        atomicAdd(&psum[part], acc);
        lo = hi;
    }
    __syncwarp();
    return __any_sync(0xffffffffu, bad);
}

template <bool WIDE, bool PACKED>
__device__ __forceinline__ bool residual_order(int order, const SigView& V, const FrameGeo& G, const int32_t* q, int shift,
                                               int psize, bool limit, unsigned long long* psum, int lane) {
    if (order <= 4) return residual_partition_sums<4, WIDE, PACKED>(V, G, q, order, shift, psize, limit, psum, lane);
    if (order <= 8) return residual_partition_sums<8, WIDE, PACKED>(V, G, q, order, shift, psize, limit, psum, lane);
    return residual_partition_sums<12, WIDE, PACKED>(V, G, q, order, shift, psize, limit, psum, lane);
}

template <bool PACKED>
__device__ __forceinline__ bool residual_dispatch(bool wide, int order, const SigView& V, const FrameGeo& G, const int32_t* q,
                                                  int shift, int psize, int nparts, bool limit, unsigned long long* psum, int lane) {
    for (int p = lane; p < nparts; p += 32) psum[p] = 0ull;
    __syncwarp();
    if (!PACKED && V.s33) {         // 33-bit side channel: 64-bit samples, always the wide accumulator
        if (order <= 4) return residual_partition_sums<4, true, false, true>(V, G, q, order, shift, psize, limit, psum, lane);
        if (order <= 8) return residual_partition_sums<8, true, false, true>(V, G, q, order, shift, psize, limit, psum, lane);
        return residual_partition_sums<12, true, false, true>(V, G, q, order, shift, psize, limit, psum, lane);
    }
    return wide ? residual_order<true, PACKED>(order, V, G, q, shift, psize, limit, psum, lane)
                : residual_order<false, PACKED>(order, V, G, q, shift, psize, limit, psum, lane);
}

// ------------------------------------------------------------------------------------------------
// Fixed-predictor error sums (up: fixed.c FLAC__fixed_compute_best_predictor[_wide]):
// sum |k-th difference| over samples 4..N-1, k = 0..4.  Lane p streams its own samples; the running
// differences live in registers; 32-bit lane partials are flushed into 64-bit totals every `flush` samples
// (flush * 2^(sbps+4) < 2^32).  Returns the warp totals in e[0..4] (every lane).
template <bool PACKED>
__device__ __noinline__ void fixed_error_sums(SigView V, FrameGeo G, int flush, int lane, unsigned long long* e) {
    const int blk_lo = lane * G.B0, blk_hi = min(G.N, blk_lo + G.B0);
    const int lo = max(blk_lo, 4);
    unsigned long long e0 = 0, e1 = 0, e2 = 0, e3 = 0, e4 = 0;
    if (lo < blk_hi) {
        const int32_t* rowp = V.base + lane * G.RS - blk_lo;
        int x1 = sig_word<PACKED>(V.base[pidx(G, lo - 1)], V), x2 = sig_word<PACKED>(V.base[pidx(G, lo - 2)], V);
        const int x3 = sig_word<PACKED>(V.base[pidx(G, lo - 3)], V), x4 = sig_word<PACKED>(V.base[pidx(G, lo - 4)], V);
        int d1 = x1 - x2, d2 = d1 - (x2 - x3), d3 = d2 - ((x2 - x3) - (x3 - x4));
        for (int g = lo; g < blk_hi; g += flush) {
            uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
            const int ge = min(blk_hi, g + flush);
#pragma unroll 4
            for (int i = g; i < ge; i++) {
                const int x0 = sig_word<PACKED>(rowp[i], V);
                const int a1 = x0 - x1, a2 = a1 - d1, a3 = a2 - d2, a4 = a3 - d3;
                s0 += (uint32_t)abs(x0); s1 += (uint32_t)abs(a1); s2 += (uint32_t)abs(a2); s3 += (uint32_t)abs(a3); s4 += (uint32_t)abs(a4);
                x1 = x0; d1 = a1; d2 = a2; d3 = a3;
            }
            e0 += s0; e1 += s1; e2 += s2; e3 += s3; e4 += s4;
        }
    }
    e[0] = warp_sum_u64(e0); e[1] = warp_sum_u64(e1); e[2] = warp_sum_u64(e2); e[3] = warp_sum_u64(e3); e[4] = warp_sum_u64(e4);
}

// up: fixed.c FLAC__fixed_compute_best_predictor_limit_residual (subframe_bps >= 28), as the shipped x86-64 binary runs
// it (oracle/flac_oracle.c:fixed_best_predictor_limit documents the pinning): sums of |k-th difference| over ALL
// samples (order k from sample k on) in 64-bit arithmetic, an order is invalid when any of its residuals exceeds
// INT32_MAX in magnitude.  Returns totals in e[0..4] and a validity bit mask (every lane).  Plain layout only.
__device__ __noinline__ uint32_t fixed_error_sums_limit(SigView V, FrameGeo G, int lane, unsigned long long* e) {
    const int blk_lo = lane * G.B0, blk_hi = min(G.N, blk_lo + G.B0);
    unsigned long long e0 = 0, e1 = 0, e2 = 0, e3 = 0, e4 = 0;
    uint32_t invalid = 0;
    if (blk_lo < blk_hi) {
        auto at = [&](int i) { return i >= 0 ? sig_word64(V, pidx(G, i)) : 0ll; };
        long long x1 = at(blk_lo - 1), x2 = at(blk_lo - 2), x3 = at(blk_lo - 3), x4 = at(blk_lo - 4);
        for (int i = blk_lo; i < blk_hi; i++) {
            const long long x0 = sig_word64(V, lane * G.RS - blk_lo + i);
            const unsigned long long a0 = (unsigned long long)llabs(x0);
            const unsigned long long a1 = i >= 1 ? (unsigned long long)llabs(x0 - x1) : 0ull;
            const unsigned long long a2 = i >= 2 ? (unsigned long long)llabs(x0 - 2 * x1 + x2) : 0ull;
            const unsigned long long a3 = i >= 3 ? (unsigned long long)llabs(x0 - 3 * x1 + 3 * x2 - x3) : 0ull;
            const unsigned long long a4 = i >= 4 ? (unsigned long long)llabs(x0 - 4 * x1 + 6 * x2 - 4 * x3 + x4) : 0ull;
            e0 += a0; e1 += a1; e2 += a2; e3 += a3; e4 += a4;
            invalid |= (a0 > 0x7fffffffull ? 1u : 0u) | (a1 > 0x7fffffffull ? 2u : 0u) | (a2 > 0x7fffffffull ? 4u : 0u) |
                       (a3 > 0x7fffffffull ? 8u : 0u) | (a4 > 0x7fffffffull ? 16u : 0u);
            x4 = x3; x3 = x2; x2 = x1; x1 = x0;
        }
    }
    e[0] = warp_sum_u64(e0); e[1] = warp_sum_u64(e1); e[2] = warp_sum_u64(e2); e[3] = warp_sum_u64(e3); e[4] = warp_sum_u64(e4);
    return __reduce_or_sync(0xffffffffu, invalid);
}

template <bool PACKED>
__device__ __forceinline__ void fixed_error_sums_rt(const SigView& V, const FrameGeo& G, int sbps, int lane, unsigned long long* e) {
    const int room = 32 - (sbps + 4);
    const int flush = room >= 6 ? 64 : (room >= 1 ? (1 << room) : 1);
    fixed_error_sums<PACKED>(V, G, flush, lane, e);
}

// ------------------------------------------------------------------------------------------------
// Per-frame record shared by the three kernels: OR / AND of every signal, then the autocorrelations
// [signal][window][kAcStoreStride] (doubles).
struct FrameBits { uint32_t or_[kMaxSignals], and_[kMaxSignals]; };
__device__ __forceinline__ const FrameBits* frame_bits(const unsigned char* work, size_t stride, int f) {
    return reinterpret_cast<const FrameBits*>(work + (size_t)f * stride);
}
__device__ __forceinline__ double* frame_ac(unsigned char* work, size_t stride, int f) {
    return reinterpret_cast<double*>(work + (size_t)f * stride + sizeof(FrameBits));
}
// up: process_subframes_ + get_wasted_bits_: OR of all samples gives the wasted bits, OR == AND means
// every sample is equal (constant subframe).  One warp per frame.
template <typename PcmT, bool PACKED>
__global__ void __launch_bounds__(128)
frame_bits_kernel(const PcmT* __restrict__ pcm, const FrameDesc* __restrict__ frames, int n_frames, EncParams P,
                  unsigned char* __restrict__ work, size_t work_stride) {
    const int lane = threadIdx.x & 31, f = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (f >= n_frames) return;
    const FrameDesc fd = frames[f];
    const int N = (int)fd.blocksize, ch = (int)P.channels, nsig = (int)P.n_signals;
    const PcmT* base = pcm + fd.pcm_off;
    FrameBits* out = reinterpret_cast<FrameBits*>(work + (size_t)f * work_stride);
    if (PACKED) {
        uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0, a0 = ~0u, a1 = ~0u, a2 = ~0u, a3 = ~0u;
        const bool aligned = ((reinterpret_cast<uintptr_t>(base) & 3u) == 0);
        for (int i = lane; i < N; i += 32) {
            int wd;
            if (aligned) wd = __ldg(reinterpret_cast<const int*>(base) + i);
            else wd = (int)((uint32_t)(uint16_t)__ldg(base + 2 * i) | ((uint32_t)(uint16_t)__ldg(base + 2 * i + 1) << 16));
            const int lo = (int)(short)wd, hi = wd >> 16, m = (lo + hi) >> 1, sd = lo - hi;
            o0 |= (uint32_t)lo; o1 |= (uint32_t)hi; o2 |= (uint32_t)m; o3 |= (uint32_t)sd;
            a0 &= (uint32_t)lo; a1 &= (uint32_t)hi; a2 &= (uint32_t)m; a3 &= (uint32_t)sd;
        }
        o0 = __reduce_or_sync(0xffffffffu, o0); o1 = __reduce_or_sync(0xffffffffu, o1);
        o2 = __reduce_or_sync(0xffffffffu, o2); o3 = __reduce_or_sync(0xffffffffu, o3);
        a0 = __reduce_and_sync(0xffffffffu, a0); a1 = __reduce_and_sync(0xffffffffu, a1);
        a2 = __reduce_and_sync(0xffffffffu, a2); a3 = __reduce_and_sync(0xffffffffu, a3);
        if (lane == 0) {
            out->or_[0] = o0; out->or_[1] = o1; out->or_[2] = o2; out->or_[3] = o3;
            out->and_[0] = a0; out->and_[1] = a1; out->and_[2] = a2; out->and_[3] = a3;
        }
    } else {
        for (int s = 0; s < nsig; s++) {
            uint32_t o = 0, a = ~0u;
            for (int i = lane; i < N; i += 32) {
                int v;
                if (s < ch) v = (int)__ldg(base + (uint64_t)i * ch + s);
                else {      // 64-bit: the sum / difference of two 32-bit samples needs 33 bits (only the low 32 matter for OR / AND)
                    const long long l = (long long)__ldg(base + (uint64_t)i * ch), r = (long long)__ldg(base + (uint64_t)i * ch + 1);
                    v = (s == ch) ? (int)((l + r) >> 1) : (int)(l - r);
                }
                o |= (uint32_t)v; a &= (uint32_t)v;
            }
            o = __reduce_or_sync(0xffffffffu, o); a = __reduce_and_sync(0xffffffffu, a);
            if (lane == 0) { out->or_[s] = o; out->and_[s] = a; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Sequential double autocorrelation (up: lpc.c FLAC__lpc_compute_autocorrelation) of
// windowed segments (up: FLAC__lpc_window_data / _partial).
//
// A job = one (frame, signal, window).  Jobs of one window depth b are numbered frame-major and dealt to warps in
// runs of kAcJobsMax; the warp runs max(len) steps (shorter jobs continue with zero samples, which leave their
// accumulators unchanged).  Lane = (job, quad): it owns the chains of lags 4q .. 4q+3; lags beyond the first reuse
// the lagged operands of the previous steps from registers, so two steps cost one 16-byte shared load for the
// current values, one for the lagged values and eight DFMAs.  Each chain is strictly sequential in i (one rounding
// per add, ascending i).
//
// Per job a ring of two 32-sample slots of doubles (plus a 16-entry mirror of slot 1's tail in front of slot 0, so
// "i - lag" is always a plain negative offset).  The inputs of chunk c+1 are fetched from HBM/L2 before the chains of
// chunk c run and are windowed (f32 multiply, widen to f64) into the other slot after them.
//   segment sample i:  i <  part          : x[off+i] * w[i]
//                      part <= i < 2*part : x[off+i] * w[N-2*part+i]
//                      i == 2*part        : 0            (full window: part = N)
struct __align__(16) AcJob {
    const void* base;        // first container element of the segment (frame base + off * channels)
    int32_t  sel;            // c0 | c1 << 8 | (ca & 0xff) << 16 | (cb & 0xff) << 24: value = (x[c0]*ca + x[c1]*cb) >> sh
    int32_t  sh;
    int32_t  len, part, N;
    uint32_t woff;           // window table offset of the frame's blocksize
};

template <typename PcmT, bool PACKED>
__device__ __forceinline__ float ac_fetch(const AcJob& J, const float* __restrict__ windows, int ch, int i) {
    const bool in = (i < J.len) && (i < 2 * J.part);
    if (!in) return 0.0f;
    const float wv = __ldg(windows + J.woff + (i < J.part ? i : J.N - 2 * J.part + i));
    const int ca = (int)(signed char)(J.sel >> 16), cb = (int)(signed char)(J.sel >> 24);
    int v = 0;
    if (PACKED) {
        const PcmT* b = reinterpret_cast<const PcmT*>(J.base);
        int wd;
        if ((reinterpret_cast<uintptr_t>(b) & 3u) == 0) wd = __ldg(reinterpret_cast<const int*>(b) + i);
        else wd = (int)((uint32_t)(uint16_t)__ldg(b + 2 * i) | ((uint32_t)(uint16_t)__ldg(b + 2 * i + 1) << 16));
        const int lo = (int)(short)wd, hi = wd >> 16;
        v = (lo * ca + hi * cb) >> J.sh;
    } else {
        // up: FLAC__lpc_window_data[_wide]: 64-bit so that mid / side of 32-bit input (33 bits) convert exactly like libFLAC's
        // (float)int64; values that fit int32 give the same float as (float)int32
        const PcmT* b = reinterpret_cast<const PcmT*>(J.base) + (size_t)i * ch;
        long long v64 = (long long)__ldg(b + (J.sel & 0xff)) * ca;
        if (cb) v64 += (long long)__ldg(b + ((J.sel >> 8) & 0xff)) * cb;
        v64 >>= J.sh;
        return FB_FMUL(__ll2float_rn(v64), wv);
    }
    return FB_FMUL(__int2float_rn(v), wv);
}

struct AcSource {            // packed layout: what the nsig jobs of one (frame, window) share
    const int16_t* base;
    int len, part;
    int wofs, wtail;         // window index = wofs + i (i < part) or wtail + i (part <= i < 2*part)
    uint32_t shpack;         // right shift of signal s in byte s (wasted bits; +1 for mid)
};

// What a lane holds for the next chunk while the chains of the current one run: raw loads only (packed layout: one word
// and one window value per source), so nothing waits for HBM/L2 until ac_finish() turns them into windowed floats AFTER
// the chains.  Other layouts: the finished floats (their fetch is per job).
struct AcRaw { int wd[2]; float wv[2]; float dv[kAcJobsMax]; };

template <typename PcmT, bool PACKED>
__device__ __forceinline__ void ac_load(AcRaw& R, const AcSource (&src)[2], const AcJob* __restrict__ jobs,
                                        int jobs_per_warp, const float* __restrict__ windows, int ch, int i) {
    if (PACKED) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
            R.wv[u] = 0.0f; R.wd[u] = 0;
            if (i < src[u].len && i < 2 * src[u].part) {
                R.wv[u] = __ldg(windows + (i < src[u].part ? src[u].wofs : src[u].wtail) + i);
                if ((reinterpret_cast<uintptr_t>(src[u].base) & 3u) == 0) R.wd[u] = __ldg(reinterpret_cast<const int*>(src[u].base) + i);
                else R.wd[u] = (int)((uint32_t)(uint16_t)__ldg(src[u].base + 2 * i) | ((uint32_t)(uint16_t)__ldg(src[u].base + 2 * i + 1) << 16));
            }
        }
    } else {
#pragma unroll
        for (int q = 0; q < kAcJobsMax; q++) R.dv[q] = (q < jobs_per_warp) ? ac_fetch<PcmT, PACKED>(jobs[q], windows, ch, i) : 0.0f;
    }
}

template <bool PACKED>
__device__ __forceinline__ void ac_finish(float (&dv)[kAcJobsMax], const AcRaw& R, const AcSource (&src)[2], int nsig) {
    if (PACKED) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int lo = (int)(short)R.wd[u], hi = R.wd[u] >> 16;
            const float wv = R.wv[u];
            const uint32_t sp = src[u].shpack;
            dv[4 * u + 0] = FB_FMUL(__int2float_rn(lo >> (sp & 0xff)), wv);
            dv[4 * u + 1] = FB_FMUL(__int2float_rn(hi >> ((sp >> 8) & 0xff)), wv);
            if (nsig > 2) {
                dv[4 * u + 2] = FB_FMUL(__int2float_rn((lo + hi) >> ((sp >> 16) & 0xff)), wv);
                dv[4 * u + 3] = FB_FMUL(__int2float_rn((lo - hi) >> (sp >> 24)), wv);
            } else {        // two signals per source: jobs 0,1 | 2,3
                dv[4 * u + 2] = 0.0f; dv[4 * u + 3] = 0.0f;
            }
        }
        if (nsig == 2) { dv[2] = dv[4]; dv[3] = dv[5]; dv[4] = dv[5] = 0.0f; }
    } else {
#pragma unroll
        for (int q = 0; q < kAcJobsMax; q++) dv[q] = R.dv[q];
    }
}

#ifndef FB_AC_MIN_CTAS
#define FB_AC_MIN_CTAS 6      // measured: 85 registers (no spills) at 24 warps/SM beats 64 registers at 32 warps/SM by ~6 %
#endif
template <typename PcmT, bool PACKED>
__global__ void __launch_bounds__(kAcThreads, FB_AC_MIN_CTAS)
autoc_kernel(const PcmT* __restrict__ pcm, const FrameDesc* __restrict__ frames, int n_frames, const float* __restrict__ windows,
             EncParams P, unsigned char* __restrict__ work, size_t work_stride, int lanes_per_job, int jobs_per_warp, int unshifted) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nsig = (int)P.n_signals, ch = (int)P.channels;
    const int nwin = (int)(P.apod_parts * (P.apod_parts + 1) / 2);
    double* ring = reinterpret_cast<double*>(smem_raw) + (size_t)warp * kAcJobsMax * kAcRing;
    AcJob* jobs = reinterpret_cast<AcJob*>(smem_raw + (size_t)(kAcThreads / 32) * kAcJobsMax * kAcRing * 8) + warp * kAcJobsMax;

    // ---- which jobs does this warp own?  depth b = 1..parts, each depth padded to whole warps ----
    long long W = (long long)blockIdx.x * (kAcThreads / 32) + warp;
    int b = 1;
    for (;; b++) {
        if (b > (int)P.apod_parts) return;
        const long long nj = (long long)n_frames * nsig * b, nw = (nj + jobs_per_warp - 1) / jobs_per_warp;
        if (W < nw) break;
        W -= nw;
    }
    const long long r0 = W * jobs_per_warp, nj = (long long)n_frames * nsig * b;
    if (lane < jobs_per_warp) {
        AcJob J;
        J.base = nullptr; J.sel = 0; J.sh = 0; J.len = 0; J.part = 0; J.N = 0; J.woff = 0;
        const long long r = r0 + lane;
        if (r < nj) {
            const int f = (int)(r / (nsig * b)), rem = (int)(r - (long long)f * (nsig * b)), k = rem / nsig, s = rem - k * nsig;
            const FrameDesc fd = frames[f];
            const int N = (int)fd.blocksize;
            if (N > 4 && !(b > 1 && N / b <= 32)) {
                const FrameBits* fb = frame_bits(work, work_stride, f);
                // un-shifted mode (enc_fused.cu): the wasted bits are not known yet; the consumer rescales by the exact factor 4^-wasted
                const int wst = unshifted ? 0 : wasted_from_or(fb->or_[s], (int)P.bps, P.bps == 32 && P.do_mid_side && s == ch + 1);
                const int off = (k * N) / b;
                J.base = pcm + fd.pcm_off + (size_t)off * ch;
                int c0, c1, ca, cb, sh = wst;
                if (s < ch) { c0 = s; c1 = s; ca = 1; cb = 0; if (PACKED && s == 1) { ca = 0; cb = 1; } }
                else if (s == ch) { c0 = 0; c1 = 1; ca = 1; cb = 1; sh = wst + 1; }
                else { c0 = 0; c1 = 1; ca = 1; cb = -1; }
                J.sel = c0 | (c1 << 8) | ((ca & 0xff) << 16) | ((cb & 0xff) << 24);
                J.sh = sh; J.len = N / b; J.part = (b == 1) ? N : N / b / 2; J.N = N; J.woff = fd.window_off;
            }
        }
        jobs[lane] = J;
    }
    for (int idx = lane; idx < jobs_per_warp * 16; idx += 32) ring[(idx >> 4) * kAcRing + (idx & 15)] = 0.0;
    __syncwarp();
    int maxlen = 0;
    for (int q = 0; q < jobs_per_warp; q++) maxlen = max(maxlen, jobs[q].len);
    if (maxlen == 0) return;

    // Packed 16-bit stereo: the nsig jobs of one (frame, window) read the same words and window values, so the warp
    // fetches per SOURCE (two of them) and derives left / right / mid / side from the one word; their geometry lives
    // in registers.  Other layouts fetch per job from the descriptors in shared memory.
    AcSource src[2];
    if (PACKED) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const AcJob& J0 = jobs[u * nsig < jobs_per_warp ? u * nsig : 0];
            src[u].base = reinterpret_cast<const int16_t*>(J0.base); src[u].len = (u * nsig < jobs_per_warp) ? J0.len : 0;
            src[u].part = J0.part; src[u].wofs = (int)J0.woff; src[u].wtail = (int)J0.woff + J0.N - 2 * J0.part;
            uint32_t shp = 0;
            for (int s2 = 0; s2 < nsig; s2++) shp |= (uint32_t)(jobs[(u * nsig + s2) < jobs_per_warp ? u * nsig + s2 : 0].sh & 0xff) << (8 * s2);
            src[u].shpack = shp;
        }
    }
    float dv[kAcJobsMax];
    AcRaw raw;
    ac_load<PcmT, PACKED>(raw, src, jobs, jobs_per_warp, windows, ch, lane);
    ac_finish<PACKED>(dv, raw, src, nsig);
#pragma unroll
    for (int q = 0; q < kAcJobsMax; q++) if (q < jobs_per_warp) ring[q * kAcRing + 16 + lane] = (double)dv[q];
    __syncwarp();

    const int jb = lane / lanes_per_job, qd = lane - jb * lanes_per_job;
    const bool active = jb < jobs_per_warp;
    const double* jobring = ring + (active ? jb : 0) * kAcRing;
    const int lag0 = 4 * qd;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;       // chains of lags lag0 .. lag0+3
    double p1 = 0.0, p2 = 0.0, p3 = 0.0;                 // lagged operands of the three previous steps
    const int nchunks = (maxlen + 31) >> 5;
    for (int c = 0; c < nchunks; c++) {
        const int slot = c & 1;
        ac_load<PcmT, PACKED>(raw, src, jobs, jobs_per_warp, windows, ch, (c + 1) * 32 + lane);     // in flight while the chains run
        if (active) {
            const double* curp = jobring + 16 + slot * 32;
            const double* lagp = curp - lag0;
#pragma unroll
            for (int s = 0; s < 32; s += 2) {
                const double2 c2 = *reinterpret_cast<const double2*>(curp + s);
                const double2 l2 = *reinterpret_cast<const double2*>(lagp + s);
                a0 = fma(c2.x, l2.x, a0); a1 = fma(c2.x, p1, a1); a2 = fma(c2.x, p2, a2); a3 = fma(c2.x, p3, a3);
                a0 = fma(c2.y, l2.y, a0); a1 = fma(c2.y, l2.x, a1); a2 = fma(c2.y, p1, a2); a3 = fma(c2.y, p2, a3);
                p3 = p1; p2 = l2.x; p1 = l2.y;
            }
        }
        __syncwarp();
        ac_finish<PACKED>(dv, raw, src, nsig);
#pragma unroll
        for (int q = 0; q < kAcJobsMax; q++) {
            if (q < jobs_per_warp) {
                const double d = (double)dv[q];
                ring[q * kAcRing + 16 + (slot ^ 1) * 32 + lane] = d;
                if (slot == 0 && lane >= 16) ring[q * kAcRing + lane - 16] = d;      // slot 1's tail mirrored in front of slot 0
            }
        }
        __syncwarp();
    }
    if (active) {
        const long long r = r0 + jb;
        if (r < nj && jobs[jb].len > 0) {
            const int f = (int)(r / (nsig * b)), rem = (int)(r - (long long)f * (nsig * b)), k = rem / nsig, s = rem - k * nsig;
            double* dst = frame_ac(work, work_stride, f) + ((size_t)s * nwin + (b - 1) * b / 2 + k) * kAcStoreStride + lag0;
            if (lag0 + 0 < kAcStoreStride) dst[0] = a0;
            if (lag0 + 1 < kAcStoreStride) dst[1] = a1;
            if (lag0 + 2 < kAcStoreStride) dst[2] = a2;
            if (lag0 + 3 < kAcStoreStride) dst[3] = a3;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The same autocorrelation for the headline layout (16-bit stereo staged as packed words, L / R / mid / side, at most nine
// lags: levels 3-7), with the chains dealt differently: TWO lanes per job, five chains per lane (lags 0-4 and 4-8; lag 4 is
// computed twice), SIXTEEN jobs = four (frame, window) sources per warp, all 32 lanes busy.  autoc_kernel spends three lanes
// x four chains = twelve chains on a nine-lag job and keeps 24 lanes busy: per job and step 0.5 FP64 warp instructions and
// 0.31 shared-memory wavefronts against 0.31 and 0.19 here -- and these two are what bounds the kernel (dependent DFMA
// chains fed from shared memory).  Each chain is still one strictly sequential sum in ascending sample order.
constexpr int kAc5Jobs = 16, kAc5Src = 4;
#ifndef FB_AC5_MIN_CTAS
#define FB_AC5_MIN_CTAS 5
#endif
struct AcRaw4 { int wd[kAc5Src]; float wv[kAc5Src]; };

__device__ __forceinline__ void ac5_load(AcRaw4& R, const AcSource (&src)[kAc5Src], const float* __restrict__ windows, int i) {
#pragma unroll
    for (int u = 0; u < kAc5Src; u++) {
        R.wv[u] = 0.0f; R.wd[u] = 0;
        if (i < src[u].len && i < 2 * src[u].part) {
            R.wv[u] = __ldg(windows + (i < src[u].part ? src[u].wofs : src[u].wtail) + i);
            if ((reinterpret_cast<uintptr_t>(src[u].base) & 3u) == 0) R.wd[u] = __ldg(reinterpret_cast<const int*>(src[u].base) + i);
            else R.wd[u] = (int)((uint32_t)(uint16_t)__ldg(src[u].base + 2 * i) | ((uint32_t)(uint16_t)__ldg(src[u].base + 2 * i + 1) << 16));
        }
    }
}
// windowed sample of the four signals of every source -> the jobs' ring slots (as doubles)
__device__ __forceinline__ void ac5_store(double* __restrict__ ring, const AcRaw4& R, const AcSource (&src)[kAc5Src], int pos, bool mirror, int mpos) {
#pragma unroll
    for (int u = 0; u < kAc5Src; u++) {
        const int lo = (int)(short)R.wd[u], hi = R.wd[u] >> 16;
        const float wv = R.wv[u];
        const uint32_t sp = src[u].shpack;
        double d[4];
        d[0] = (double)FB_FMUL(__int2float_rn(lo >> (sp & 0xff)), wv);
        d[1] = (double)FB_FMUL(__int2float_rn(hi >> ((sp >> 8) & 0xff)), wv);
        d[2] = (double)FB_FMUL(__int2float_rn((lo + hi) >> ((sp >> 16) & 0xff)), wv);
        d[3] = (double)FB_FMUL(__int2float_rn((lo - hi) >> (sp >> 24)), wv);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            ring[(4 * u + q) * kAcRing + pos] = d[q];
            if (mirror) ring[(4 * u + q) * kAcRing + mpos] = d[q];
        }
    }
}

__global__ void __launch_bounds__(kAcThreads, FB_AC5_MIN_CTAS)
autoc5_kernel(const int16_t* __restrict__ pcm, const FrameDesc* __restrict__ frames, int n_frames, const float* __restrict__ windows,
              EncParams P, unsigned char* __restrict__ work, size_t work_stride, int unshifted) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int nsig = 4, ch = 2;
    const int nwin = (int)(P.apod_parts * (P.apod_parts + 1) / 2);
    double* ring = reinterpret_cast<double*>(smem_raw) + (size_t)warp * kAc5Jobs * kAcRing;
    AcJob* jobs = reinterpret_cast<AcJob*>(smem_raw + (size_t)(kAcThreads / 32) * kAc5Jobs * kAcRing * 8) + warp * kAc5Jobs;

    // ---- which jobs does this warp own?  depth b = 1..parts, each depth padded to whole warps (as autoc_kernel) ----
    long long W = (long long)blockIdx.x * (kAcThreads / 32) + warp;
    int b = 1;
    for (;; b++) {
        if (b > (int)P.apod_parts) return;
        const long long nj = (long long)n_frames * nsig * b, nw = (nj + kAc5Jobs - 1) / kAc5Jobs;
        if (W < nw) break;
        W -= nw;
    }
    const long long r0 = W * kAc5Jobs, nj = (long long)n_frames * nsig * b;
    if (lane < kAc5Jobs) {
        AcJob J;
        J.base = nullptr; J.sel = 0; J.sh = 0; J.len = 0; J.part = 0; J.N = 0; J.woff = 0;
        const long long r = r0 + lane;
        if (r < nj) {
            const int f = (int)(r / (nsig * b)), rem = (int)(r - (long long)f * (nsig * b)), k = rem / nsig, s = rem - k * nsig;
            const FrameDesc fd = frames[f];
            const int N = (int)fd.blocksize;
            if (N > 4 && !(b > 1 && N / b <= 32)) {
                const FrameBits* fb = frame_bits(work, work_stride, f);
                const int wst = unshifted ? 0 : wasted_from_or(fb->or_[s], (int)P.bps, false);
                const int off = (k * N) / b;
                J.base = pcm + fd.pcm_off + (size_t)off * ch;
                J.sh = (s == ch) ? wst + 1 : wst;                        // mid: (L + R) >> 1
                J.len = N / b; J.part = (b == 1) ? N : N / b / 2; J.N = N; J.woff = fd.window_off;
            }
        }
        jobs[lane] = J;
    }
    for (int idx = lane; idx < kAc5Jobs * 16; idx += 32) ring[(idx >> 4) * kAcRing + (idx & 15)] = 0.0;
    __syncwarp();
    int maxlen = 0;
    for (int q = 0; q < kAc5Jobs; q++) maxlen = max(maxlen, jobs[q].len);
    if (maxlen == 0) return;

    AcSource src[kAc5Src];                  // the four jobs of one (frame, window) read the same words and window values
#pragma unroll
    for (int u = 0; u < kAc5Src; u++) {
        const AcJob& J0 = jobs[u * nsig];
        src[u].base = reinterpret_cast<const int16_t*>(J0.base); src[u].len = J0.len;
        src[u].part = J0.part; src[u].wofs = (int)J0.woff; src[u].wtail = (int)J0.woff + J0.N - 2 * J0.part;
        uint32_t shp = 0;
        for (int s2 = 0; s2 < nsig; s2++) shp |= (uint32_t)(jobs[u * nsig + s2].sh & 0xff) << (8 * s2);
        src[u].shpack = shp;
    }
    AcRaw4 raw;
    ac5_load(raw, src, windows, lane);
    ac5_store(ring, raw, src, 16 + lane, false, 0);
    __syncwarp();

    // lanes 0-15 take the low halves (lags 0-4) of the sixteen jobs, lanes 16-31 the high halves: a quarter warp then reads eight
    // different jobs' rings (16 bytes each, 4 banks apart: conflict free); with the two halves of a job in neighbouring lanes the
    // high half's -32 bytes put jobs j and j + 2 into the same banks
#ifndef FB_AC5_PAIRED
#define FB_AC5_PAIRED 0
#endif
    const int jb = FB_AC5_PAIRED ? lane >> 1 : lane & 15, qd = FB_AC5_PAIRED ? lane & 1 : lane >> 4;
    const double* jobring = ring + jb * kAcRing;
    const int lag0 = 4 * qd;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0;      // chains of lags lag0 .. lag0+4
    double p1 = 0.0, p2 = 0.0, p3 = 0.0, p4 = 0.0;                 // lagged operands of the four previous steps
    const int nchunks = (maxlen + 31) >> 5;
    for (int c = 0; c < nchunks; c++) {
        const int slot = c & 1;
        ac5_load(raw, src, windows, (c + 1) * 32 + lane);          // in flight while the chains run
        {
            const double* curp = jobring + 16 + slot * 32;
            const double* lagp = curp - lag0;
#pragma unroll
            for (int s = 0; s < 32; s += 2) {
                const double2 c2 = *reinterpret_cast<const double2*>(curp + s);
                // (the low halves' lagged operands are the current ones; loading them only in lanes 16-31 measured slower: 0.75 against 0.72 ms)
                const double2 l2 = *reinterpret_cast<const double2*>(lagp + s);
                a0 = fma(c2.x, l2.x, a0); a1 = fma(c2.x, p1, a1); a2 = fma(c2.x, p2, a2); a3 = fma(c2.x, p3, a3); a4 = fma(c2.x, p4, a4);
                a0 = fma(c2.y, l2.y, a0); a1 = fma(c2.y, l2.x, a1); a2 = fma(c2.y, p1, a2); a3 = fma(c2.y, p2, a3); a4 = fma(c2.y, p3, a4);
                p4 = p2; p3 = p1; p2 = l2.x; p1 = l2.y;
            }
        }
        __syncwarp();
        // the next chunk into the other slot; slot 1's tail is mirrored in front of slot 0
        ac5_store(ring, raw, src, 16 + (slot ^ 1) * 32 + lane, slot == 0 && lane >= 16, lane - 16);
        __syncwarp();
    }
    {
        const long long r = r0 + jb;
        if (r < nj && jobs[jb].len > 0) {
            const int f = (int)(r / (nsig * b)), rem = (int)(r - (long long)f * (nsig * b)), k = rem / nsig, s = rem - k * nsig;
            double* dst = frame_ac(work, work_stride, f) + ((size_t)s * nwin + (b - 1) * b / 2 + k) * kAcStoreStride + lag0;
            dst[0] = a0; dst[1] = a1; dst[2] = a2; dst[3] = a3;
            if (qd == 1) dst[4] = a4;                              // lag 8 (lane 0's fifth chain is lag 4 again)
        }
    }
}

// ------------------------------------------------------------------------------------------------
struct AnShared {
    uint32_t sig_or[kMaxSignals], sig_and[kMaxSignals];
    uint32_t best_bits[kMaxSignals];                 // winner of the fixed task, then of the whole signal
    uint32_t step_bits[kMaxSignals][kMaxSteps];      // LPC candidate estimate per apodization step (0xffffffff = none)
    int      need_list[kMaxSignals];
    int      nneed;
    int      queue_a, queue_b;
    int32_t  fixed_q[kAnWarps][4];                   // binomial coefficients of the fixed candidate a warp is evaluating
};

#ifndef FB_AN_MIN_CTAS
#define FB_AN_MIN_CTAS 8
#endif
template <typename PcmT, bool PACKED>
#ifndef FB_AN_MIN_CTAS_WIDE
#define FB_AN_MIN_CTAS_WIDE 4      // measured on the 24-bit mono level-8 shape (configs[2]): 150 registers at 3 CTAs/SM 4.74 ms, 114 at 4: 4.33, 96 (spills) at 5: 4.28
#endif
__global__ void __launch_bounds__(kAnThreads, PACKED ? FB_AN_MIN_CTAS : FB_AN_MIN_CTAS_WIDE)
analyze_kernel(const PcmT* __restrict__ pcm, const FrameDesc* __restrict__ frames, const float* __restrict__ windows,
               EncParams P, SubframePlan* __restrict__ plans, uint8_t* __restrict__ frame_ca,
               SignalDebug* __restrict__ dbg, EncStats* __restrict__ stats, int pass,
               unsigned char* __restrict__ work, size_t work_stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nsig = (int)P.n_signals;
    const FrameDesc fd = frames[blockIdx.x];
    const int N = (int)fd.blocksize;
    const int ch = (int)P.channels;
    // up: process_subframes_ loose_mid_side_stereo -- decision frames (pass 0) analyse L, R, M, S and choose between
    // independent and mid/side only; the frames that follow (pass 1) analyse just the pair that decision picked.
    int mode = 0;                           // 0 = all signals, 1 = independent channels only, 2 = mid/side only
    if (P.loose_frames) {
        if (fd.lead == 0u) { if (pass != 0) return; }
        else {
            if (pass == 0) return;
            const int pca = (fd.lead & kLeadForced) ? (int)(fd.lead & 3u) : (int)frame_ca[blockIdx.x - fd.lead];
            mode = (pca == 0) ? 1 : 2;
        }
    }
    auto sig_active = [&](int s) { return mode == 0 || (mode == 1 ? s < ch : s >= ch); };

    FrameGeo G;
    G.N = N; G.B0 = (N + 31) >> 5; G.pad = (G.B0 & 1) ? 0 : 1; G.RS = G.B0 + G.pad;
    G.magic = G.pad ? (uint32_t)((0x100000000ull + (uint32_t)G.B0 - 1ull) / (uint32_t)G.B0) : 0u;

    // ---- shared memory carve-up (sizes mirrored by analyze_smem_bytes) ----
    const int sig_words = (int)P.an_stride;                                // words per staged signal (or per packed frame)
    const int n_steps = (int)P.apod_steps;                                     // apodization steps per signal
    const int nwin = (int)(P.apod_parts * (P.apod_parts + 1) / 2);
    int32_t* xall = reinterpret_cast<int32_t*>(smem_raw);
    unsigned char* cur = smem_raw + (size_t)(PACKED ? 1 : nsig) * sig_words * 4;
    WarpScratch* wsall = reinterpret_cast<WarpScratch*>(cur);                    cur += (size_t)kAnWarps * sizeof(WarpScratch);
    const double* acstore = frame_ac(work, work_stride, (int)blockIdx.x);        // [nsig][nwin][kAcStoreStride], written by autoc_kernel
    unsigned long long* psum_all = reinterpret_cast<unsigned long long*>(cur);   cur += (size_t)kAnWarps * 2 * kMaxParts * 8;
    SubframePlan* base_plan = reinterpret_cast<SubframePlan*>(cur);              cur += (size_t)nsig * sizeof(SubframePlan);
    SubframePlan* step_plan = reinterpret_cast<SubframePlan*>(cur);              cur += (size_t)nsig * n_steps * sizeof(SubframePlan);
    AnShared& S = *reinterpret_cast<AnShared*>(cur);

    unsigned long long* psum = psum_all + (size_t)warp * 2 * kMaxParts;
    WarpScratch& ws = wsall[warp];

    if (tid < kMaxSignals) {
        const FrameBits* fb = frame_bits(work, work_stride, (int)blockIdx.x);
        S.sig_or[tid] = tid < nsig ? fb->or_[tid] : 0u; S.sig_and[tid] = tid < nsig ? fb->and_[tid] : 0xffffffffu; S.best_bits[tid] = 0u;
    }
    if (tid == 0) { S.queue_a = 0; S.queue_b = 0; S.nneed = 0; }
    for (int i = tid; i < kMaxSignals * kMaxSteps; i += kAnThreads) (&S.step_bits[0][0])[i] = 0xffffffffu;

    // =================== stage the frame (container form) ===================
    {
        const PcmT* base = pcm + fd.pcm_off;
        if (PACKED) {
            const bool aligned = ((reinterpret_cast<uintptr_t>(base) & 3u) == 0);
            for (int i = tid; i < N; i += kAnThreads) {
                int wd;
                if (aligned) wd = __ldg(reinterpret_cast<const int*>(base) + i);
                else wd = (int)((uint32_t)(uint16_t)__ldg(base + 2 * i) | ((uint32_t)(uint16_t)__ldg(base + 2 * i + 1) << 16));
                xall[pidx(G, i)] = wd;
            }
        } else {
            const bool need_lr = P.bps == 32 && P.do_mid_side;       // the 33-bit side channel is read through the left / right rows
            for (int s = 0; s < nsig; s++) {
                if (!sig_active(s) && !(need_lr && s < ch)) continue;
                int32_t* x = xall + (size_t)s * sig_words;
                const int wst = wasted_from_or(frame_bits(work, work_stride, (int)blockIdx.x)->or_[s], (int)P.bps, P.bps == 32 && P.do_mid_side && s == ch + 1);   // rows hold the signal with wasted bits removed
                for (int i = tid; i < N; i += kAnThreads) {
                    int v;
                    if (s < ch) v = (int)__ldg(base + (uint64_t)i * ch + s) >> wst;
                    else {      // 64-bit: 32-bit input needs 33 bits here; a side channel that keeps all 33 is read through L / R instead
                        const long long l = (long long)__ldg(base + (uint64_t)i * ch), r = (long long)__ldg(base + (uint64_t)i * ch + 1);
                        v = (s == ch) ? (int)(((l + r) >> 1) >> wst) : (int)((l - r) >> wst);
                    }
                    x[pidx(G, i)] = v;
                }
            }
        }
    }
    __syncthreads();

    // per-signal facts every thread can derive on its own
    auto sig_wasted = [&](int s) { return wasted_from_or(S.sig_or[s], (int)P.bps, P.bps == 32 && P.do_mid_side && s == ch + 1); };
    auto sig_sbps = [&](int s) { return (int)P.bps - sig_wasted(s) + ((P.do_mid_side && s == ch + 1) ? 1 : 0); };
    auto sig_view = [&](int s) {
        SigView V;
        V.base2 = nullptr; V.shl = 0; V.shr = 0; V.s33 = 0; V.cab = 1;
        if (PACKED) { V.base = xall; V.ca = (s != 1); V.cb = (s == 0) ? 0 : (s == 3 ? -1 : 1); V.sh = sig_wasted(s) + (s == 2 ? 1 : 0); V.cab = (V.ca & 0xff) | ((V.cb & 0xff) << 8); }
        else {
            V.base = xall + (size_t)s * sig_words; V.ca = 1; V.cb = 0; V.sh = 0;
            if (P.bps == 32 && P.do_mid_side && s == ch + 1 && sig_wasted(s) == 0) {      // 33-bit side: read through the left / right rows
                V.base = xall; V.base2 = xall + (size_t)sig_words; V.shl = sig_wasted(0); V.shr = sig_wasted(1); V.s33 = 1;
            }
        }
        return V;
    };
    // up: process_subframe_ constant test + process_subframes_ limit_min_bitrate: when every earlier channel is
    // constant, the last channel (and mid/side after it) may not use a constant subframe
    // subframe_bps >= 28 (the _limit_residual predictor search): only an all-zero block comes out as CONSTANT there
    // (oracle/flac_oracle.c:fixed_best_predictor_limit), a non-zero constant block becomes FIXED order 1
    auto sig_const = [&](int s) { return N > 4 && (sig_sbps(s) >= 28 ? S.sig_or[s] == 0u : S.sig_or[s] == S.sig_and[s]); };
    auto sig_disable_const = [&](int s) {
        if (!(P.limit_min_bitrate && mode != 2 && s >= ch - 1)) return false;
        for (int c2 = 0; c2 < ch - 1; c2++) if (!sig_const(c2)) return false;
        return true;
    };
    const int omax_frame = min((int)P.max_part_order, N ? (__ffs(N) - 1) : 0);
    const int max_lpc = (N > 4 && P.max_lpc_order > 0) ? (((int)P.max_lpc_order >= N) ? N - 1 : (int)P.max_lpc_order) : 0;
    if (tid == 0) {
        int n = 0;
        for (int s = 0; s < nsig; s++)
            if (sig_active(s) && N > 4 && max_lpc > 0 && !(sig_const(s) && !sig_disable_const(s))) S.need_list[n++] = s;
        S.nneed = n;
    }
    __syncthreads();
    const int nneed = S.nneed;

    // =================== phase A: per-signal fixed analysis, from a queue ===================
    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&S.queue_a, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= nsig) break;
        {
            // ---- fixed analysis of signal s: verbatim baseline, constant, or the guessed fixed order ----
            const int s = t;
            SubframePlan& pl = base_plan[s];
            reinterpret_cast<uint32_t*>(&pl)[lane] = 0u;
            __syncwarp();
            if (!sig_active(s)) continue;
            const int wasted = sig_wasted(s), sbps = sig_sbps(s);
            SignalDebug* dg = dbg ? dbg + (size_t)blockIdx.x * nsig + s : nullptr;
            uint32_t best_bits = 8u + (uint32_t)wasted + (uint32_t)N * (uint32_t)sbps;      // up: evaluate_verbatim_subframe_
            if (lane == 0) { pl.type = kVerbatim; pl.wasted = (uint8_t)wasted; pl.sbps = (uint8_t)sbps; pl.bits_est = best_bits; }
            if (dg && lane == 0) { dg->n_apod = 0; dg->fixed_bits = 0; dg->is_constant = 0; dg->fixed_order = 0; for (int k = 0; k < 5; k++) dg->fixed_err[k] = 0; }
            if (N > 4) {
                const bool constant = sig_const(s);
                if (constant && !sig_disable_const(s)) {
                    const uint32_t bits = 8u + (uint32_t)wasted + (uint32_t)sbps;     // up: evaluate_constant_subframe_
                    if (dg && lane == 0) dg->is_constant = 1;
                    if (bits < best_bits) { best_bits = bits; if (lane == 0) { pl.type = kConstant; pl.bits_est = bits; } }
                } else {
                    const SigView V = sig_view(s);
                    unsigned long long e[5];
                    int forder;
                    bool try_fixed = true;
                    if (!PACKED && sbps >= 28) {
                        const uint32_t invalid = fixed_error_sums_limit(V, G, lane, e);
                        unsigned long long smallest = ~0ull;
                        forder = 0;
#pragma unroll
                        for (int k = 0; k < 5; k++) if (!((invalid >> k) & 1u) && e[k] < smallest) { forder = k; smallest = e[k]; }
                        try_fixed = ((invalid >> forder) & 1u) == 0u;      // no valid order: estimate 34.0 >= subframe_bps, "don't even try"
                    } else {
                        fixed_error_sums_rt<PACKED>(V, G, sbps, lane, e);
                        if ((uint32_t)sbps + ilog2_u32((uint32_t)N - 4u) + 1u < 32u) {   // libFLAC's 32-bit accumulators wrap
#pragma unroll
                            for (int k = 0; k < 5; k++) e[k] &= 0xffffffffull;
                        }
                        const unsigned long long m34 = min(e[3], e[4]), m234 = min(e[2], m34), m1234 = min(e[1], m234);
                        if (e[0] <= m1234) forder = 0; else if (e[1] <= m234) forder = 1; else if (e[2] <= m34) forder = 2; else if (e[3] <= e[4]) forder = 3; else forder = 4;
                    }
                    if (dg && lane == 0) { for (int k = 0; k < 5; k++) dg->fixed_err[k] = e[k]; dg->fixed_order = forder; dg->is_constant = constant; }
                    // fixed candidate at the guessed order (up: evaluate_fixed_subframe_)
                    int fo = forder; if (fo >= N) fo = N - 1;
                    if (try_fixed) {
                    int omax = omax_frame;
                    while (omax > 0 && (N >> omax) <= fo) omax--;
                    const int nparts = 1 << omax, psize = N >> omax;
                    const bool narrow = (uint32_t)sbps + 4u < 32u - ilog2_u32((uint32_t)psize);
                    if (lane == 0) {
                        const int32_t c[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};   // x[i] - sum q_j x[i-1-j]
                        for (int j = 0; j < 4; j++) S.fixed_q[warp][j] = c[fo][j];
                    }
                    __syncwarp();
                    // fixed residual of <=24-bit input fits 32-bit arithmetic (|4th difference| < 2^(sbps+4))
                    residual_dispatch<PACKED>(sbps + 4 > 31, fo, V, G, S.fixed_q[warp], 0, psize, nparts, false, psum, lane);
                    int po; uint32_t k0, k1;
                    const uint32_t rb = rice_search(psum, N, fo, omax, narrow, P.rice_limit, lane, &po, &k0, &k1);
                    const uint32_t est = add_sat(8u + (uint32_t)wasted + (uint32_t)fo * (uint32_t)sbps, rb);
                    if (dg && lane == 0) dg->fixed_bits = est;
                    if (est < best_bits) {
                        best_bits = est;
                        if (lane < (1 << po)) pl.rice[lane] = (uint8_t)k0;
                        if (lane + 32 < (1 << po)) pl.rice[lane + 32] = (uint8_t)k1;
                        const bool r2 = __any_sync(0xffffffffu, (lane < (1 << po) && k0 >= 15u) || (lane + 32 < (1 << po) && k1 >= 15u));
                        if (lane == 0) { pl.type = kFixed; pl.order = (uint8_t)fo; pl.part_order = (uint8_t)po; pl.rice2 = r2; pl.bits_est = est; pl.shift = 0; pl.precision = 0; }
                    }
                    }
                    __syncwarp();
                }
            }
            if (lane == 0) S.best_bits[s] = best_bits;
            __syncwarp();
        }
    }
    __syncthreads();

    // =================== phase B: one LPC candidate per (signal, apodization step), from a queue ===================
    // up: apply_apodization_ + evaluate_lpc_subframe_.  Step list of set_next_subdivide_tukey:
    // full window, then for depth b = 2..parts: partial windows c = 0,2,.. interleaved with their punch-outs.
    const int n_tasks_b = nneed * n_steps;
    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&S.queue_b, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_tasks_b) break;
        const int s = S.need_list[t / n_steps], step = t - (t / n_steps) * n_steps;
        // ---- step -> (b, c): depth and position in libFLAC's walk; b == 1 is the full window ----
        int b = 1, c = 0;
        {
            int k = step;
            if (k > 0) {
                k -= 1; b = 2;
                for (;;) { const int cnt = (b == 2) ? 2 : 2 * b; if (k < cnt) break; k -= cnt; b++; }
                c = (b == 2) ? 2 * k : k;        // depth 2 visits c = 0 and c = 2 only (its punch-outs equal the other half)
            }
        }
        if (b > 1 && N / b <= 32) continue;      // window too short: libFLAC skips the step
        const int wasted = sig_wasted(s), sbps = sig_sbps(s);
        SignalDebug* dg = dbg ? dbg + (size_t)blockIdx.x * nsig + s : nullptr;
        const double* myac = acstore + (size_t)s * nwin * kAcStoreStride;
        int max_this = max_lpc;
        double ac_cur = 0.0;                     // lane j holds lag j
        if (b == 1) { if (lane <= max_this) ac_cur = myac[lane]; }
        else if (!(c & 1)) { if (lane <= max_this) ac_cur = myac[((b - 1) * b / 2 + c / 2) * kAcStoreStride + lane]; }
        else if (lane <= max_this) {
            // punch-out: root minus the partial window before it, for lags < max order only (1.4.3 off-by-one)
            const double partial = myac[((b - 1) * b / 2 + c / 2) * kAcStoreStride + lane];
            ac_cur = (lane < max_this) ? FB_DSUB(myac[lane], partial) : partial;
        }
        if (lane <= max_this) ws.ac[lane] = ac_cur;
        __syncwarp();
        if (dg && step < kMaxApodSteps) { if (lane <= max_this) dg->autoc[step][lane] = ac_cur; if (lane == 0) { dg->lpc_order[step] = 0; dg->lpc_bits[step] = 0; } }
        if (ws.ac[0] == 0.0) { __syncwarp(); continue; }

        if (lane == 0) ws.misc[0] = levinson(ws.ac, max_this, ws.lp, ws.lperr, ws.lpc);
        __syncwarp();
        max_this = ws.misc[0];

        // up: lpc.c FLAC__lpc_compute_best_order -- first strict minimum, initial best (uint32_t)-1
        int guess;
        uint32_t guard_kind = 0;
        const uint32_t guard_overhead = (uint32_t)sbps + P.qlp_precision;
        {
            const double escale = FB_DDIV(0.5, (double)N);
            const uint32_t overhead = (uint32_t)sbps + P.qlp_precision;
            double bits = 1.7976931348623157e308; bool ul = false;
            if (lane >= 1 && lane <= max_this) {
                const double e = expected_bits_per_sample(ws.lperr[lane - 1], escale, &ul);
                bits = FB_DADD(FB_DMUL(e, (double)(N - lane)), (double)((uint32_t)lane * overhead));
            }
            double bb = bits; int bi = lane;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, bb, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob < bb || (ob == bb && oi < bi)) { bb = ob; bi = oi; }
            }
            guess = (bb < 4294967295.0) ? bi : 1;
            // guard band: a runner-up within guard_rel (1e-12 relative: several thousand ulps of the log) of the winner could flip under a libm
            // log that differs in the last ulp: the decision is logged and the host repeats it with the reference's libm (DESIGN.md "log guard")
            const int ul_best = __shfl_sync(0xffffffffu, (int)ul, guess & 31);
            const bool amb = (lane >= 1 && lane <= max_this && lane != guess) && (ul || ul_best) && fabs(bits - bb) <= P.guard_rel * fabs(bb);
            const unsigned amb_mask = __ballot_sync(0xffffffffu, amb);
            guard_kind = amb_mask ? 1u : 0u;
            if (amb_mask && P.guard_flip) guess = __ffs((int)amb_mask) - 1;
        }
        if (dg && step < kMaxApodSteps) { if (lane < max_this) dg->lpc_err[step][lane] = ws.lperr[lane]; if (lane == 0) dg->lpc_order[step] = guess; }

        const LogGuardOverride* g_ov = P.guard_n_ovr ? guard_find(P, fd, s, step) : nullptr;
        if (g_ov) guess = g_ov->guess;
        const int order = guess;
        bool ul2;
        const double rbps = expected_bits_per_sample(ws.lperr[order - 1], FB_DDIV(0.5, (double)(N - order)), &ul2);
        bool g_skip = rbps >= (double)sbps;
        if (ul2 && fabs(rbps - (double)sbps) <= P.guard_rel * (double)sbps) { guard_kind |= 2u; if (P.guard_flip) g_skip = !g_skip; }
        if (g_ov) g_skip = g_ov->skip != 0;
        else if (guard_kind) guard_record(P, stats, fd, s, step, N, sbps, max_this, guard_overhead, ws.lperr, guess, g_skip, guard_kind, lane);
        if (!g_skip) {
            int prec = (int)P.qlp_precision;
            if (sbps <= 17) prec = min(prec, 32 - sbps - (int)ilog2_u32((uint32_t)order));
            if (lane == 0) {
                int sh = 0;
                const int rc = quantize_coefficients(ws.lp + (order - 1) * kMaxOrder, order, prec, ws.q, &sh);
                int32_t asum = 0;
                for (int j = 0; j < order; j++) asum += abs(ws.q[j]);
                if (asum == 0) asum = 1;
                ws.misc[1] = rc; ws.misc[2] = sh; ws.misc[3] = (int)silog2((int64_t)asum);
            }
            __syncwarp();
            if (ws.misc[1] == 0) {
                const int shift = ws.misc[2];
                // up: lpc.c FLAC__lpc_max_prediction_before_shift_bps / FLAC__lpc_max_residual_bps
                const int pred_bps = sbps + ws.misc[3];
                const int resid_bps = ((sbps > pred_bps - shift) ? sbps : pred_bps - shift) + 1;
                const bool limit = resid_bps > 32;
                int omax = omax_frame;
                while (omax > 0 && (N >> omax) <= order) omax--;
                const int nparts = 1 << omax, psize = N >> omax;
                const bool narrow = (uint32_t)sbps + 4u < 32u - ilog2_u32((uint32_t)psize);
                const SigView V = sig_view(s);
                const bool rejected = residual_dispatch<PACKED>(limit || pred_bps > 32, order, V, G, ws.q, shift, psize, nparts, limit, psum, lane);
                if (!rejected) {
                    int po; uint32_t k0, k1;
                    const uint32_t rb = rice_search(psum, N, order, omax, narrow, P.rice_limit, lane, &po, &k0, &k1);
                    const uint32_t est = add_sat(8u + (uint32_t)wasted + 4u + 5u + (uint32_t)order * (uint32_t)(prec + sbps), rb);
                    if (dg && lane == 0 && step < kMaxApodSteps) dg->lpc_bits[step] = est;
                    SubframePlan& pl = step_plan[(size_t)s * n_steps + step];
                    reinterpret_cast<uint32_t*>(&pl)[lane] = 0u;
                    __syncwarp();
                    if (lane < (1 << po)) pl.rice[lane] = (uint8_t)k0;
                    if (lane + 32 < (1 << po)) pl.rice[lane + 32] = (uint8_t)k1;
                    const bool r2 = __any_sync(0xffffffffu, (lane < (1 << po) && k0 >= 15u) || (lane + 32 < (1 << po) && k1 >= 15u));
                    if (lane < order) pl.qlp[lane] = ws.q[lane];
                    if (lane == 0) {
                        pl.type = kLpc; pl.order = (uint8_t)order; pl.part_order = (uint8_t)po; pl.rice2 = r2; pl.bits_est = est; pl.shift = shift;
                        pl.precision = (uint8_t)prec; pl.wasted = (uint8_t)wasted; pl.sbps = (uint8_t)sbps;
                        S.step_bits[s][step] = est;
                    }
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();

    // =================== selection: candidates in libFLAC's order, replace only on strict < ===================
    for (int s = warp; s < nsig; s += kAnWarps) {
        uint32_t best = S.best_bits[s];
        int pick = -1;
        if (sig_active(s))
            for (int k = 0; k < n_steps; k++) { const uint32_t e = S.step_bits[s][k]; if (e < best) { best = e; pick = k; } }
        const uint32_t* src = reinterpret_cast<const uint32_t*>(pick < 0 ? &base_plan[s] : &step_plan[(size_t)s * n_steps + pick]);
        uint32_t* dst = reinterpret_cast<uint32_t*>(plans + (size_t)blockIdx.x * nsig + s);
        dst[lane] = src[lane];     // 128 bytes = 32 words
        if (lane == 0) {
            if (dbg) {
                bool need = false;
                for (int k = 0; k < nneed; k++) need = need || (S.need_list[k] == s);
                (dbg + (size_t)blockIdx.x * nsig + s)->n_apod = need ? n_steps : 0;
            }
            S.best_bits[s] = best;
        }
    }
    __syncthreads();

    // ---- channel assignment (up: process_subframes_): first minimum of {L+R, L+S, R+S, M+S} ----
    if (tid == 0) {
        int ca = 0;
        if (P.loose_frames) {
            if (mode == 0) ca = (S.best_bits[2] + S.best_bits[3] < S.best_bits[0] + S.best_bits[1]) ? 3 : 0;
            else ca = (mode == 1) ? 0 : 3;
        } else if (P.do_mid_side) {
            const uint32_t bL = S.best_bits[0], bR = S.best_bits[1], bM = S.best_bits[2], bS = S.best_bits[3];
            uint32_t minb = bL + bR;
            if (bL + bS < minb) { minb = bL + bS; ca = 1; }
            if (bR + bS < minb) { minb = bR + bS; ca = 2; }
            if (bM + bS < minb) { minb = bM + bS; ca = 3; }
        }
        frame_ca[blockIdx.x] = (uint8_t)ca;
    }
}

// ------------------------------------------------------------------------------------------------ host side
static bool packed_layout(const EncParams& P) { return P.container_bytes == 2 && P.channels == 2; }

static uint32_t apod_steps(const EncParams& P) {
    // up: set_next_subdivide_tukey: 1 (full) + 2 (depth 2) + 2b (depth b >= 3)
    uint32_t n = 1;
    for (uint32_t b = 2; b <= P.apod_parts; b++) n += (b == 2) ? 2u : 2u * b;
    return P.max_lpc_order ? n : 0u;
}

// bytes of per-frame scratch in HBM: OR/AND record + autocorrelations of every (signal, window)
size_t analyze_work_stride(const EncParams& P) {
    const size_t nwin = P.apod_parts * (P.apod_parts + 1) / 2;
    return sizeof(FrameBits) + (size_t)P.n_signals * nwin * kAcStoreStride * sizeof(double);
}

template <typename PcmT, bool PACKED>
static void launch_all(const PcmT* pcm, const FrameDesc* frames, const float* windows, const EncParams& P, int n_frames,
                       SubframePlan* plans, uint8_t* frame_ca, SignalDebug* dbg, EncStats* stats, size_t smem_bytes,
                       unsigned char* work, cudaStream_t stream, cudaEvent_t* ev) {
    const size_t stride = analyze_work_stride(P);
    frame_bits_kernel<PcmT, PACKED><<<(n_frames + 3) / 4, 128, 0, stream>>>(pcm, frames, n_frames, P, work, stride);
    if (ev) cudaEventRecord(ev[0], stream);
    if (P.max_lpc_order > 0) {
        const int lags = (int)P.max_lpc_order + 1, lpj = (lags + 3) / 4;
        int jpw = (32 / lpj) < kAcJobsMax ? (32 / lpj) : kAcJobsMax;
        if (PACKED && jpw > 2 * (int)P.n_signals) jpw = 2 * (int)P.n_signals;     // packed layout: a warp fetches for two (frame, window) sources
        if (PACKED) jpw = jpw / (int)P.n_signals * (int)P.n_signals;
        long long warps = 0;
        for (uint32_t b = 1; b <= P.apod_parts; b++) warps += ((long long)n_frames * P.n_signals * b + jpw - 1) / jpw;
        const size_t ac_smem = (size_t)(kAcThreads / 32) * kAcJobsMax * (kAcRing * 8 + sizeof(AcJob));
        const unsigned blocks = (unsigned)((warps + kAcThreads / 32 - 1) / (kAcThreads / 32));
        autoc_kernel<PcmT, PACKED><<<blocks, kAcThreads, ac_smem, stream>>>(pcm, frames, n_frames, windows, P, work, stride, lpj, jpw, 0);
    }
    if (ev) cudaEventRecord(ev[1], stream);
    cudaFuncSetAttribute(analyze_kernel<PcmT, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    // loose mid/side: decision frames first, then the frames that follow them (they read the decision from frame_ca)
    for (int pass = 0; pass < (P.loose_frames ? 2 : 1); pass++)
        analyze_kernel<PcmT, PACKED><<<(unsigned)n_frames, kAnThreads, smem_bytes, stream>>>(pcm, frames, windows, P, plans, frame_ca, dbg, stats,
                                                                                            pass, work, stride);
}

// The autocorrelation kernel alone, on the un-shifted signals of 16-bit stereo frames (first kernel of the enc_fused.cu path:
// it needs no OR/AND pass in front of it).  Same work records as launch_analyze.
void launch_autoc_unshifted(const void* pcm, const FrameDesc* frames, const float* windows, const EncParams& P, int n_frames, void* work, cudaStream_t stream) {
    if (P.max_lpc_order == 0 || n_frames == 0) return;
    const size_t stride = analyze_work_stride(P);
    const int lags = (int)P.max_lpc_order + 1, lpj = (lags + 3) / 4;
#ifndef FB_AC5
#define FB_AC5 1
#endif
    if (FB_AC5 && P.n_signals == 4 && lags > 4 && lags <= 9) {                 // two lanes x five chains per job, sixteen jobs per warp
        long long warps5 = 0;
        for (uint32_t b = 1; b <= P.apod_parts; b++) warps5 += ((long long)n_frames * 4 * b + kAc5Jobs - 1) / kAc5Jobs;
        const size_t smem5 = (size_t)(kAcThreads / 32) * kAc5Jobs * (kAcRing * 8 + sizeof(AcJob));
        static bool attr_set = false;       // (an attribute of the function, the same value on every device)
        if (!attr_set) { cudaFuncSetAttribute(autoc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem5); attr_set = true; }
        const unsigned blocks5 = (unsigned)((warps5 + kAcThreads / 32 - 1) / (kAcThreads / 32));
        autoc5_kernel<<<blocks5, kAcThreads, smem5, stream>>>((const int16_t*)pcm, frames, n_frames, windows, P, (unsigned char*)work, stride, 1);
        return;
    }
    int jpw = (32 / lpj) < kAcJobsMax ? (32 / lpj) : kAcJobsMax;
    if (jpw > 2 * (int)P.n_signals) jpw = 2 * (int)P.n_signals;
    jpw = jpw / (int)P.n_signals * (int)P.n_signals;
    long long warps = 0;
    for (uint32_t b = 1; b <= P.apod_parts; b++) warps += ((long long)n_frames * P.n_signals * b + jpw - 1) / jpw;
    const size_t ac_smem = (size_t)(kAcThreads / 32) * kAcJobsMax * (kAcRing * 8 + sizeof(AcJob));
    const unsigned blocks = (unsigned)((warps + kAcThreads / 32 - 1) / (kAcThreads / 32));
    autoc_kernel<int16_t, true><<<blocks, kAcThreads, ac_smem, stream>>>((const int16_t*)pcm, frames, n_frames, windows, P, (unsigned char*)work, stride, lpj, jpw, 1);
}

// host-visible launcher (called from engine.cu); `work` = n_frames * analyze_work_stride(P) bytes of device scratch.
// ev (optional): two events recorded after the OR/AND kernel and after the autocorrelation kernel.
// Returns the number of kernels launched.
int launch_analyze(const void* pcm, const FrameDesc* frames, const float* windows, const EncParams& P, int n_frames,
                   SubframePlan* plans, uint8_t* frame_ca, SignalDebug* dbg, EncStats* stats, size_t smem_bytes,
                   void* work, cudaStream_t stream, cudaEvent_t* ev) {
    if (packed_layout(P)) launch_all<int16_t, true>((const int16_t*)pcm, frames, windows, P, n_frames, plans, frame_ca, dbg, stats, smem_bytes, (unsigned char*)work, stream, ev);
    else if (P.container_bytes == 2) launch_all<int16_t, false>((const int16_t*)pcm, frames, windows, P, n_frames, plans, frame_ca, dbg, stats, smem_bytes, (unsigned char*)work, stream, ev);
    else launch_all<int32_t, false>((const int32_t*)pcm, frames, windows, P, n_frames, plans, frame_ca, dbg, stats, smem_bytes, (unsigned char*)work, stream, ev);
    return 1 + (P.max_lpc_order > 0 ? 1 : 0) + (P.loose_frames ? 2 : 1);
}

// Shared-memory plan of the analysis kernel; fills P.apod_steps (apodization steps per signal) and P.an_stride (staged
// words per signal).  Called by the host before launching.
void analyze_layout(EncParams& P) {
    P.apod_steps = apod_steps(P);
    // 32 rows of ceil(N/32) samples with an odd row stride
    const uint32_t b0 = (P.blocksize + 31) / 32, rs = b0 | 1u;
    P.an_stride = ((32u * rs + 3u) / 4u) * 4u;
}

size_t analyze_smem_bytes(const EncParams& P) {
    const size_t nsig = P.n_signals;
    return (size_t)(packed_layout(P) ? 1 : nsig) * P.an_stride * 4 + (size_t)kAnWarps * sizeof(WarpScratch) +
           (size_t)kAnWarps * 2 * kMaxParts * 8 + nsig * sizeof(SubframePlan) +
           nsig * P.apod_steps * sizeof(SubframePlan) + sizeof(AnShared) + 64;
}

}  // namespace fb
