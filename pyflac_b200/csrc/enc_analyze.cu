// enc_analyze.cu -- encode analysis kernel: one CTA per frame, one warp per signal.
//
// For every signal (channel, or mid / side) the warp reproduces libFLAC 1.4.3's process_subframe_
// decision sequence (SURVEY A.3-A.9, rows E2-E11 of the scope table): wasted bits, fixed-predictor
// error sums, window + sequential-double autocorrelation, Levinson-Durbin, order guess, coefficient
// quantisation, residual -> partition sums -> Rice parameter / partition-order search, and keeps the
// candidate with the smallest libFLAC bit estimate.  Thread 0 then picks the channel assignment.
// Output: one 128-byte SubframePlan per signal + one channel-assignment byte per frame.  The pack
// kernel (enc_pack.cu) turns plans into bits.
//
// Exactness: integer work is exact; floating point follows fb_math.cuh (unfused, RN); the
// autocorrelation is one DFMA chain per lag in ascending sample order (products of two floats are
// exact in double, so fma == mul+add as libFLAC computes it).  No tensor cores: this is integer /
// bit-serial work bounded by dependent-issue latency, not by a dense contraction.
#include "fb_common.cuh"
#include "fb_math.cuh"

namespace fb {

constexpr int kAcJobsMax = 3;          // (signal, window) chains packed into one warp: floor(32 / lags)
constexpr int kAcBufStride = 112;      // doubles per job: 16 mirror + 3 slots of 32
constexpr int kMaxWindows = 6;         // full + 2 halves + 3 thirds (subdivide_tukey(3))
constexpr int kAcStoreStride = 13;     // lags kept per (signal, window)

struct __align__(16) WarpScratch {
    double   ac[16];                   // autocorrelation of the current apodization step
    double   lperr[kMaxOrder];         // Levinson error per order
    double   lpc[kMaxOrder];           // Levinson recursion state
    float    lp[kMaxOrder * kMaxOrder];
    int32_t  q[16];                    // quantised coefficients of the current candidate
    int32_t  misc[8];
    SubframePlan plan;                 // best candidate so far
};

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T>
__device__ __forceinline__ int load_pcm(const T* p, uint64_t idx) { return (int)__ldg(p + idx); }

// r[i] = x[i] - ((sum_j q[j] * x[i-1-j]) >> shift) with partition sums of |r| at the maximum
// partition order.  Fixed predictors are the same formula with binomial coefficients and shift 0
// (up: fixed.c FLAC__fixed_compute_residual == lpc residual with q = {1},{2,-1},{3,-3,1},{4,-6,4,-1}).
// WIDE: 64-bit accumulate (up: FLAC__lpc_compute_residual_from_qlp_coefficients_wide); the 32-bit
// form is used exactly when libFLAC proves it cannot overflow (max_prediction_before_shift_bps <= 32).
// check_limit reproduces the _limit_residual rejection (residual must fit int32, INT32_MIN excluded).
template <int ORDER, bool WIDE>
__device__ __forceinline__ bool residual_partition_sums(const int32_t* __restrict__ x, int N, const int32_t* qs,
                                                        int shift, int psize, int nparts, bool narrow_sums,
                                                        bool check_limit, unsigned long long* psum, int lane) {
    int32_t q[ORDER > 0 ? ORDER : 1];
#pragma unroll
    for (int j = 0; j < ORDER; j++) q[j] = qs[j];
    bool bad = false;
    for (int p = 0; p < nparts; p++) {
        int lo = p * psize;
        const int hi = lo + psize;
        if (p == 0) lo = ORDER;
        unsigned long long acc = 0;
        for (int i = lo + lane; i < hi; i += 32) {
            long long r;
            if (WIDE) {
                long long s = 0;
#pragma unroll
                for (int j = 0; j < ORDER; j++) s += (long long)q[j] * (long long)x[i - 1 - j];
                r = (long long)x[i] - (s >> shift);
                if (check_limit && (r <= (long long)INT32_MIN || r > (long long)INT32_MAX)) bad = true;
            } else {
                int s = 0;
#pragma unroll
                for (int j = 0; j < ORDER; j++) s += q[j] * x[i - 1 - j];
                r = (long long)(x[i] - (s >> shift));
            }
            acc += (unsigned long long)(r < 0 ? -r : r);
        }
        unsigned long long tot;
        if (__reduce_or_sync(0xffffffffu, (unsigned)(acc >> 27)) == 0u) tot = __reduce_add_sync(0xffffffffu, (unsigned)acc);
        else tot = warp_sum_u64(acc);
        if (lane == 0) psum[p] = narrow_sums ? (tot & 0xffffffffull) : tot;
    }
    return __any_sync(0xffffffffu, bad);
}

template <bool WIDE>
__device__ bool residual_dispatch(int order, const int32_t* x, int N, const int32_t* q, int shift, int psize,
                                  int nparts, bool narrow, bool limit, unsigned long long* psum, int lane) {
    switch (order) {
        case 0: return residual_partition_sums<0, WIDE>(x, N, q, shift, psize, nparts, narrow, limit, psum, lane);
        case 1: return residual_partition_sums<1, WIDE>(x, N, q, shift, psize, nparts, narrow, limit, psum, lane);
        case 2: return residual_partition_sums<2, WIDE>(x, N, q, shift, psize, nparts, narrow, limit, psum, lane);
        case 3: return residual_partition_sums<3, WIDE>(x, N, q, shift, psize, nparts, narrow, limit, psum, lane);
        case 4: return residual_partition_sums<4, WIDE>(x, N, q, shift, psize, nparts, narrow, limit, psum, lane);
        case 5: return residual_partition_sums<5, WIDE>(x, N, q, shift, psize, nparts, narrow, limit, psum, lane);
        case 6: return residual_partition_sums<6, WIDE>(x, N, q, shift, psize, nparts, narrow, limit, psum, lane);
        case 7: return residual_partition_sums<7, WIDE>(x, N, q, shift, psize, nparts, narrow, limit, psum, lane);
        case 8: return residual_partition_sums<8, WIDE>(x, N, q, shift, psize, nparts, narrow, limit, psum, lane);
        case 9: return residual_partition_sums<9, WIDE>(x, N, q, shift, psize, nparts, narrow, limit, psum, lane);
        case 10: return residual_partition_sums<10, WIDE>(x, N, q, shift, psize, nparts, narrow, limit, psum, lane);
        case 11: return residual_partition_sums<11, WIDE>(x, N, q, shift, psize, nparts, narrow, limit, psum, lane);
        default: return residual_partition_sums<12, WIDE>(x, N, q, shift, psize, nparts, narrow, limit, psum, lane);
    }
}

// up: stream_encoder.c find_best_partition_order_ / set_partitioned_rice_: given the sums at
// the maximum order in psum[0 .. 2^omax), search orders omax..0 (first strict minimum), merging pairwise.
// Lane p owns partitions p and p+32.  Returns estimated residual bits; best parameters land in k0/k1.
__device__ __forceinline__ uint32_t rice_search(unsigned long long* psum, int N, int pred_order, int omax,
                                                uint32_t rice_limit, int lane, int* best_order_out,
                                                uint32_t* k0_out, uint32_t* k1_out) {
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

// Sequential double autocorrelation (up: lpc.c FLAC__lpc_compute_autocorrelation) of
// windowed segments (up: FLAC__lpc_window_data / _partial), for every (signal, window) "job"
// of the frame.  Jobs of equal length are packed into lane blocks of `L` lanes (lane = lag), up to
// floor(32 / L) jobs per warp, so that one DFMA instruction advances several chains: this phase is
// bound by dependent-issue latency, not by HBM or FP64 throughput.  Each chain is strictly sequential in
// i (one rounding per add, ascending i) -- parallelism comes from lags x signals x windows only.
//
// Per job a ring of three 32-sample slots of doubles (plus a 16-entry mirror of slot 2's tail in front of
// slot 0, so "i - lag" is always a plain negative offset).  While the 32 fused steps of chunk c run
//      acc = fma(slot[s], slot[s - lag], acc)          (immediate-offset shared loads, no index math)
// the same warp windows chunk c+1 (f32 multiply, widen to f64) into the next slot: the loads and
// conversions sit in the same basic block as the DFMA chain, so they fill its issue gaps.
// Terms with i < lag multiply zero history and leave the accumulator unchanged, as do zero-padded tail terms.
//   segment sample i:  i <  part          : x[dshift+i] * w[i]
//                      part <= i < 2*part : x[dshift+i] * w[N-2*part+i]
//                      i == 2*part        : 0            (full window: part = N)
struct AcJobs { const int32_t* xs[kAcJobsMax]; int cnt; };

__device__ __forceinline__ void ac_window_store(double* __restrict__ buf, const AcJobs& J, const float* __restrict__ w,
                                                int N, int part, int len, int base, int slot, int lane) {
    const int i = base + lane;
    const bool in = (i < len) && (i < 2 * part);
    float wv = 0.0f;
    if (in) wv = __ldg(w + (i < part ? i : N - 2 * part + i));
#pragma unroll
    for (int b2 = 0; b2 < kAcJobsMax; b2++) {
        if (b2 < J.cnt) {
            const double d = in ? (double)FB_FMUL(__int2float_rn(J.xs[b2][i]), wv) : 0.0;
            buf[b2 * kAcBufStride + 16 + slot * 32 + lane] = d;
            if (slot == 2 && lane >= 16) buf[b2 * kAcBufStride + lane - 16] = d;
        }
    }
}

template <int SLOT>
__device__ __forceinline__ void ac_chunk(double& acc, double* __restrict__ buf, const double* __restrict__ jobbuf, int lag,
                                         const AcJobs& J, const float* __restrict__ w, int N, int part, int len,
                                         int next_base, int lane) {
    constexpr int NEXT = (SLOT + 1) % 3;
    // inputs of the next chunk first (their latency hides under the DFMA chain below)
    const int i = next_base + lane;
    const bool in = (i < len) && (i < 2 * part);
    float wv = 0.0f;
    int xv[kAcJobsMax];
    if (in) wv = __ldg(w + (i < part ? i : N - 2 * part + i));
#pragma unroll
    for (int b2 = 0; b2 < kAcJobsMax; b2++) xv[b2] = (in && b2 < J.cnt) ? J.xs[b2][i] : 0;
    const double* curp = jobbuf + 16 + SLOT * 32;
    const double* lagp = curp - lag;
#pragma unroll
    for (int s = 0; s < 32; s += 2) {
        const double2 c2 = *reinterpret_cast<const double2*>(curp + s);
        acc = fma(c2.x, lagp[s], acc);
        acc = fma(c2.y, lagp[s + 1], acc);
    }
#pragma unroll
    for (int b2 = 0; b2 < kAcJobsMax; b2++) {
        if (b2 < J.cnt) {
            const double d = (double)FB_FMUL(__int2float_rn(xv[b2]), wv);
            buf[b2 * kAcBufStride + 16 + NEXT * 32 + lane] = d;
            if (NEXT == 2 && lane >= 16) buf[b2 * kAcBufStride + lane - 16] = d;
        }
    }
    __syncwarp();
}

__device__ __forceinline__ void autoc_phase(const int32_t* __restrict__ xall, int smem_stride, const float* __restrict__ w,
                                            int N, int L, int parts, int nwin, const int* __restrict__ need_list, int nneed,
                                            double* __restrict__ buf, double* __restrict__ acstore, int warp, int nwarps,
                                            int lane) {
    const int gmax = min(kAcJobsMax, 32 / L);
    const int jb = lane / L, lag = lane - jb * L;
    int g = 0;
    for (int b = 1; b <= parts; b++) {
        if (b > 1 && N / b <= 32) continue;
        const int len = N / b, part = (b == 1) ? N : N / b / 2;
        const int nj = nneed * b;
        const int ngrp = (nj + gmax - 1) / gmax, gsz = (nj + ngrp - 1) / ngrp;      // balanced lane blocks
        for (int j0 = 0; j0 < nj; j0 += gsz, g++) {
            if (g % nwarps != warp) continue;
            AcJobs J;
            J.cnt = min(gsz, nj - j0);
#pragma unroll
            for (int b2 = 0; b2 < kAcJobsMax; b2++) {
                const int j = min(j0 + b2, nj - 1), sidx = need_list[j / b], k = j - (j / b) * b;
                J.xs[b2] = xall + (size_t)sidx * smem_stride + (k * N) / b;
            }
            const bool active = jb < J.cnt;
            const double* jobbuf = buf + (active ? jb : 0) * kAcBufStride;
            const int lag_eff = active ? lag : 0;
            for (int idx = lane; idx < J.cnt * 16; idx += 32) buf[(idx >> 4) * kAcBufStride + (idx & 15)] = 0.0;
            ac_window_store(buf, J, w, N, part, len, 0, 0, lane);
            __syncwarp();
            double acc = 0.0;
            const int nchunks = (len + 31) >> 5;
            for (int c = 0; c < nchunks; c += 3) {
                ac_chunk<0>(acc, buf, jobbuf, lag_eff, J, w, N, part, len, (c + 1) * 32, lane);
                if (c + 1 < nchunks) ac_chunk<1>(acc, buf, jobbuf, lag_eff, J, w, N, part, len, (c + 2) * 32, lane);
                if (c + 2 < nchunks) ac_chunk<2>(acc, buf, jobbuf, lag_eff, J, w, N, part, len, (c + 3) * 32, lane);
            }
            if (active) {
                const int j = j0 + jb, sidx = need_list[j / b], k = j - (j / b) * b;
                acstore[((size_t)sidx * nwin + (b - 1) * b / 2 + k) * kAcStoreStride + lag] = acc;
            }
            __syncwarp();
        }
    }
}

template <typename PcmT>
__global__ void __launch_bounds__(32 * kMaxSignals, 1)
analyze_kernel(const PcmT* __restrict__ pcm, const FrameDesc* __restrict__ frames, const float* __restrict__ windows,
               EncParams P, SubframePlan* __restrict__ plans, uint8_t* __restrict__ frame_ca,
               SignalDebug* __restrict__ dbg, EncStats* __restrict__ stats, int pass) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, sig = threadIdx.x >> 5;
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
    const bool active = mode == 0 || (mode == 1 ? sig < ch : sig >= ch);

    int32_t* xall = reinterpret_cast<int32_t*>(smem_raw);
    // [signals | pool: partition sums (phases 1,3) aliased with the autocorrelation rings (phase 2) | per-warp scratch | ...]
    unsigned char* pool = smem_raw + (size_t)nsig * P.smem_stride * 4;
    WarpScratch* wsall = reinterpret_cast<WarpScratch*>(pool + P.pool_bytes);
    unsigned long long* psum = reinterpret_cast<unsigned long long*>(pool) + (size_t)(threadIdx.x >> 5) * 2 * kMaxParts;
    double* acbuf = reinterpret_cast<double*>(pool) + (size_t)(threadIdx.x >> 5) * P.ac_gsz * kAcBufStride;
    const int nwin = (int)(P.apod_parts * (P.apod_parts + 1) / 2);
    double* acstore = reinterpret_cast<double*>(wsall + nsig);                 // [nsig][nwin][kAcStoreStride]
    uint32_t* sig_bits = reinterpret_cast<uint32_t*>(acstore + (size_t)nsig * nwin * kAcStoreStride);
    int* need_list = reinterpret_cast<int*>(sig_bits + kMaxSignals);           // signals that go through LPC analysis
    int* need_flag = need_list + kMaxSignals;
    int* nneed_p = need_flag + kMaxSignals;
    int* const_flag = nneed_p + 4;

    int32_t* x = xall + (size_t)sig * P.smem_stride;
    WarpScratch& ws = wsall[sig];
    SignalDebug* dg = dbg ? dbg + (size_t)blockIdx.x * nsig + sig : nullptr;

    // =================== phase 1 (one warp per signal): load, wasted bits, fixed predictors ===================
    // up: process_subframes_ + get_wasted_bits_
    uint32_t orv = 0;
    if (active) {
        const PcmT* base = pcm + fd.pcm_off;
        for (int i = lane; i < N; i += 32) {
            int v;
            if (sig < ch) v = load_pcm(base, (uint64_t)i * ch + sig);
            else {
                const int l = load_pcm(base, (uint64_t)i * ch), r = load_pcm(base, (uint64_t)i * ch + 1);
                v = (sig == ch) ? ((l + r) >> 1) : (l - r);
            }
            x[i] = v;
            orv |= (uint32_t)v;
        }
    }
    orv = __reduce_or_sync(0xffffffffu, orv);
    int wasted = orv ? (__ffs((int)orv) - 1) : 0;
    if (wasted > (int)P.bps) wasted = (int)P.bps;
    const int sbps = (int)P.bps - wasted + ((P.do_mid_side && sig == ch + 1) ? 1 : 0);
    __syncwarp();
    if (wasted) {
        for (int i = lane; i < N; i += 32) x[i] >>= wasted;
        __syncwarp();
    }

    // baseline: verbatim (up: evaluate_verbatim_subframe_)
    reinterpret_cast<uint32_t*>(&ws.plan)[lane] = 0u;
    __syncwarp();
    uint32_t best_bits = 8u + (uint32_t)wasted + (uint32_t)N * (uint32_t)sbps;
    if (lane == 0) {
        SubframePlan& pl = ws.plan;
        pl.type = kVerbatim; pl.wasted = (uint8_t)wasted; pl.sbps = (uint8_t)sbps; pl.bits_est = best_bits;
    }
    if (dg && lane == 0) { dg->n_apod = 0; dg->fixed_bits = 0; dg->is_constant = 0; dg->fixed_order = 0; for (int k = 0; k < 5; k++) dg->fixed_err[k] = 0; }

    // per-frame partition geometry (up: process_subframes_ max_partition_order = min(level max, ctz(N)))
    const int omax_frame = min((int)P.max_part_order, N ? (__ffs(N) - 1) : 0);
    const int max_lpc = (N > 4 && P.max_lpc_order > 0) ? (((int)P.max_lpc_order >= N) ? N - 1 : (int)P.max_lpc_order) : 0;
    bool want_lpc = false;
    bool constant = false;
    int forder = 0;

    if (active && N > 4) {
        // fixed predictor error sums (up: fixed.c FLAC__fixed_compute_best_predictor[_wide])
        unsigned long long e0 = 0, e1 = 0, e2 = 0, e3 = 0, e4 = 0;
        for (int i = 4 + lane; i < N; i += 32) {
            const int x0 = x[i], x1 = x[i - 1], x2 = x[i - 2], x3 = x[i - 3], x4 = x[i - 4];
            const int a1 = x0 - x1, b1 = x1 - x2, c1 = x2 - x3, d1 = x3 - x4;
            const int a2 = a1 - b1, b2 = b1 - c1, c2 = c1 - d1;
            const int a3 = a2 - b2, b3 = b2 - c2;
            const int a4 = a3 - b3;
            e0 += (unsigned)abs(x0); e1 += (unsigned)abs(a1); e2 += (unsigned)abs(a2); e3 += (unsigned)abs(a3); e4 += (unsigned)abs(a4);
        }
        e0 = warp_sum_u64(e0); e1 = warp_sum_u64(e1); e2 = warp_sum_u64(e2); e3 = warp_sum_u64(e3); e4 = warp_sum_u64(e4);
        const unsigned long long e1_true = e1;   // constant detection must not be fooled by the 32-bit wrap below
        if ((uint32_t)sbps + ilog2_u32((uint32_t)N - 4u) + 1u < 32u) {   // libFLAC's 32-bit accumulators wrap
            e0 &= 0xffffffffull; e1 &= 0xffffffffull; e2 &= 0xffffffffull; e3 &= 0xffffffffull; e4 &= 0xffffffffull;
        }
        {
            const unsigned long long m34 = min(e3, e4), m234 = min(e2, m34), m1234 = min(e1, m234);
            if (e0 <= m1234) forder = 0; else if (e1 <= m234) forder = 1; else if (e2 <= m34) forder = 2; else if (e3 <= e4) forder = 3; else forder = 4;
        }
        if (e1_true == 0) constant = (x[0] == x[1]) && (x[1] == x[2]) && (x[2] == x[3]) && (x[3] == x[4]);   // samples 3..N-1 are equal already
        if (dg && lane == 0) { dg->fixed_err[0] = e0; dg->fixed_err[1] = e1; dg->fixed_err[2] = e2; dg->fixed_err[3] = e3; dg->fixed_err[4] = e4; dg->fixed_order = forder; dg->is_constant = constant; }

    }
    // up: process_subframes_ limit_min_bitrate -- when every earlier channel came out constant, the last channel
    // (and mid/side after it) may not use a constant subframe, so the frame never shrinks to headers only
    if (lane == 0) const_flag[sig] = constant ? 1 : 0;
    __syncthreads();
    bool disable_const = false;
    if (P.limit_min_bitrate && mode != 2 && sig >= ch - 1) {
        disable_const = true;
        for (int c2 = 0; c2 < ch - 1; c2++) if (!const_flag[c2]) disable_const = false;
    }
    if (active && N > 4) {
        if (constant && !disable_const) {
            const uint32_t bits = 8u + (uint32_t)wasted + (uint32_t)sbps;     // up: evaluate_constant_subframe_
            if (bits < best_bits) { best_bits = bits; if (lane == 0) { ws.plan.type = kConstant; ws.plan.bits_est = bits; } }
        } else {
            // fixed candidate at the guessed order (up: evaluate_fixed_subframe_)
            int fo = forder; if (fo >= N) fo = N - 1;
            int omax = omax_frame;
            while (omax > 0 && (N >> omax) <= fo) omax--;
            const int nparts = 1 << omax, psize = N >> omax;
            const bool narrow = (uint32_t)sbps + 4u < 32u - ilog2_u32((uint32_t)psize);
            if (lane == 0) {
                const int32_t c[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};   // x[i] - sum q_j x[i-1-j]
                for (int j = 0; j < 4; j++) ws.q[j] = c[fo][j];
            }
            __syncwarp();
            // fixed residual of <=24-bit input fits 32-bit arithmetic (|4th difference| < 2^(sbps+4))
            if (sbps + 4 <= 31) residual_dispatch<false>(fo, x, N, ws.q, 0, psize, nparts, narrow, false, psum, lane);
            else residual_dispatch<true>(fo, x, N, ws.q, 0, psize, nparts, narrow, false, psum, lane);
            __syncwarp();
            int po; uint32_t k0, k1;
            const uint32_t rb = rice_search(psum, N, fo, omax, P.rice_limit, lane, &po, &k0, &k1);
            const uint32_t est = add_sat(8u + (uint32_t)wasted + (uint32_t)fo * (uint32_t)sbps, rb);
            if (dg && lane == 0) dg->fixed_bits = est;
            if (est < best_bits) {
                best_bits = est;
                SubframePlan& pl = ws.plan;
                if (lane < (1 << po)) pl.rice[lane] = (uint8_t)k0;
                if (lane + 32 < (1 << po)) pl.rice[lane + 32] = (uint8_t)k1;
                const bool r2 = __any_sync(0xffffffffu, (lane < (1 << po) && k0 >= 15u) || (lane + 32 < (1 << po) && k1 >= 15u));
                if (lane == 0) { pl.type = kFixed; pl.order = (uint8_t)fo; pl.part_order = (uint8_t)po; pl.rice2 = r2; pl.bits_est = est; pl.shift = 0; pl.precision = 0; }
            }
            __syncwarp();
            want_lpc = max_lpc > 0;
        }
    }
    if (lane == 0) need_flag[sig] = want_lpc ? 1 : 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0;
        for (int s2 = 0; s2 < nsig; s2++) if (need_flag[s2]) need_list[n++] = s2;
        *nneed_p = n;
    }
    __syncthreads();

    // =================== phase 2 (packed lane blocks): all autocorrelations of the frame ===================
    const int nneed = *nneed_p;
    if (nneed > 0)
        autoc_phase(xall, (int)P.smem_stride, windows + fd.window_off, N, max_lpc + 1, (int)P.apod_parts, nwin, need_list, nneed,
                    acbuf, acstore, sig, nsig, lane);
    __syncthreads();

    // =================== phase 3 (one warp per signal): LPC candidates, one per apodization step ===================
    // up: apply_apodization_ + evaluate_lpc_subframe_
    if (want_lpc) {
        const double* myac = acstore + (size_t)sig * nwin * kAcStoreStride;
        double ac_root = 0.0, ac_cur = 0.0;     // lane j holds lag j
        int step = 0;
        int b = 1, c = 0;                        // (depth, part) walk of set_next_subdivide_tukey; b == 1 is the full window
        bool done = false;
        while (!done) {
            int max_this = max_lpc;
            bool have = true;
            if (b == 1) {
                if (lane <= max_this) ac_cur = myac[lane];
                if (P.apod_parts > 1) { ac_root = ac_cur; b = 2; c = 0; } else done = true;
            } else {
                if (N / b <= 32) have = false;
                else if (!(c & 1)) {
                    if (lane <= max_this) ac_cur = myac[((b - 1) * b / 2 + c / 2) * kAcStoreStride + lane];
                } else {
                    // punch-out: root minus previous partial for lags < max order only (1.4.3 off-by-one)
                    if (lane < max_this) ac_cur = FB_DSUB(ac_root, ac_cur);
                }
                if (b == 2) { if (c == 0) c = 2; else { c = 0; b++; } }
                else if (c < 2 * b - 1) c++;
                else { c = 0; b++; }
                if (b > (int)P.apod_parts) done = true;
            }
            if (!have) continue;
            if (lane <= max_this) ws.ac[lane] = ac_cur;
            __syncwarp();
            if (dg && step < kMaxApodSteps) { if (lane <= max_this) dg->autoc[step][lane] = ac_cur; if (lane == 0) { dg->lpc_order[step] = 0; dg->lpc_bits[step] = 0; } }
            if (ws.ac[0] == 0.0) { step++; __syncwarp(); continue; }

            if (lane == 0) ws.misc[0] = levinson(ws.ac, max_this, ws.lp, ws.lperr, ws.lpc);
            __syncwarp();
            max_this = ws.misc[0];

            // up: lpc.c FLAC__lpc_compute_best_order -- first strict minimum, initial best (uint32_t)-1
            int guess;
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
                // guard band: a runner-up within 1e-9 relative of the winner could flip under a libm log that
                // differs in the last ulp (DESIGN.md "log guard"); counted, never silently ignored
                const int ul_best = __shfl_sync(0xffffffffu, (int)ul, guess & 31);
                const bool amb = (lane >= 1 && lane <= max_this && lane != guess) && (ul || ul_best) &&
                                 fabs(bits - bb) <= 1e-9 * fabs(bb);
                if (__any_sync(0xffffffffu, amb) && lane == 0 && stats) atomicAdd(&stats->log_ambiguous, 1ull);
            }
            if (dg && step < kMaxApodSteps) { if (lane < max_this) dg->lpc_err[step][lane] = ws.lperr[lane]; if (lane == 0) dg->lpc_order[step] = guess; }

            const int order = guess;
            bool ul2;
            const double rbps = expected_bits_per_sample(ws.lperr[order - 1], FB_DDIV(0.5, (double)(N - order)), &ul2);
            if (ul2 && fabs(rbps - (double)sbps) <= 1e-9 * (double)sbps && lane == 0 && stats) atomicAdd(&stats->log_ambiguous, 1ull);
            if (!(rbps >= (double)sbps)) {
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
                    bool rejected;
                    if (!limit && pred_bps <= 32) rejected = residual_dispatch<false>(order, x, N, ws.q, shift, psize, nparts, narrow, false, psum, lane);
                    else rejected = residual_dispatch<true>(order, x, N, ws.q, shift, psize, nparts, narrow, limit, psum, lane);
                    __syncwarp();
                    if (!rejected) {
                        int po; uint32_t k0, k1;
                        const uint32_t rb = rice_search(psum, N, order, omax, P.rice_limit, lane, &po, &k0, &k1);
                        const uint32_t est = add_sat(8u + (uint32_t)wasted + 4u + 5u + (uint32_t)order * (uint32_t)(prec + sbps), rb);
                        if (dg && lane == 0 && step < kMaxApodSteps) dg->lpc_bits[step] = est;
                        if (est < best_bits) {
                            best_bits = est;
                            SubframePlan& pl = ws.plan;
                            if (lane < (1 << po)) pl.rice[lane] = (uint8_t)k0;
                            if (lane + 32 < (1 << po)) pl.rice[lane + 32] = (uint8_t)k1;
                            const bool r2 = __any_sync(0xffffffffu, (lane < (1 << po) && k0 >= 15u) || (lane + 32 < (1 << po) && k1 >= 15u));
                            if (lane < order) pl.qlp[lane] = ws.q[lane];
                            if (lane == 0) { pl.type = kLpc; pl.order = (uint8_t)order; pl.part_order = (uint8_t)po; pl.rice2 = r2; pl.bits_est = est; pl.shift = shift; pl.precision = (uint8_t)prec; }
                        }
                    }
                }
                __syncwarp();
            }
            step++;
        }
        if (dg && lane == 0) dg->n_apod = step;
    }
    __syncwarp();

    // ---- publish the plan ----
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&ws.plan);
        uint32_t* dst = reinterpret_cast<uint32_t*>(plans + (size_t)blockIdx.x * nsig + sig);
        dst[lane] = src[lane];     // 128 bytes = 32 words
        if (lane == 0) sig_bits[sig] = best_bits;
    }
    __syncthreads();

    // ---- channel assignment (up: process_subframes_): first minimum of {L+R, L+S, R+S, M+S} ----
    if (threadIdx.x == 0) {
        int ca = 0;
        if (P.loose_frames) {
            if (mode == 0) ca = (sig_bits[2] + sig_bits[3] < sig_bits[0] + sig_bits[1]) ? 3 : 0;
            else ca = (mode == 1) ? 0 : 3;
        } else if (P.do_mid_side) {
            const uint32_t bL = sig_bits[0], bR = sig_bits[1], bM = sig_bits[2], bS = sig_bits[3];
            uint32_t minb = bL + bR;
            if (bL + bS < minb) { minb = bL + bS; ca = 1; }
            if (bR + bS < minb) { minb = bR + bS; ca = 2; }
            if (bM + bS < minb) { minb = bM + bS; ca = 3; }
        }
        frame_ca[blockIdx.x] = (uint8_t)ca;
    }
}

// host-visible launcher (called from engine.cu)
void launch_analyze(const void* pcm, const FrameDesc* frames, const float* windows, const EncParams& P, int n_frames,
                    SubframePlan* plans, uint8_t* frame_ca, SignalDebug* dbg, EncStats* stats, size_t smem_bytes,
                    cudaStream_t stream) {
    const dim3 grid((unsigned)n_frames), block(32u * P.n_signals);
    // loose mid/side: decision frames first, then the frames that follow them (they read the decision from frame_ca)
    for (int pass = 0; pass < (P.loose_frames ? 2 : 1); pass++) {
        if (P.container_bytes == 2) {
            cudaFuncSetAttribute(analyze_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
            analyze_kernel<int16_t><<<grid, block, smem_bytes, stream>>>((const int16_t*)pcm, frames, windows, P, plans, frame_ca, dbg, stats, pass);
        } else {
            cudaFuncSetAttribute(analyze_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
            analyze_kernel<int32_t><<<grid, block, smem_bytes, stream>>>((const int32_t*)pcm, frames, windows, P, plans, frame_ca, dbg, stats, pass);
        }
    }
}

// Shared-memory plan of the analysis kernel; fills P.pool_bytes / P.ac_gsz (called by the host before launching).
void analyze_layout(EncParams& P) {
    const uint32_t nsig = P.n_signals, L = (P.max_lpc_order ? P.max_lpc_order : 1) + 1;
    const uint32_t gmax = (32 / L) < (uint32_t)kAcJobsMax ? (32 / L) : (uint32_t)kAcJobsMax;
    uint32_t groups = 0, gsz_max = 1;
    for (uint32_t b = 1; b <= P.apod_parts; b++) {
        const uint32_t nj = nsig * b, ngrp = (nj + gmax - 1) / gmax, gsz = (nj + ngrp - 1) / ngrp;
        groups += ngrp;
        if (gsz > gsz_max) gsz_max = gsz;
    }
    const uint32_t active = groups < nsig ? groups : nsig;
    const uint32_t ac_bytes = P.max_lpc_order ? active * gsz_max * kAcBufStride * 8u : 0u;
    const uint32_t psum_bytes = nsig * 2u * kMaxParts * 8u;
    P.ac_gsz = gsz_max;
    P.pool_bytes = ((ac_bytes > psum_bytes ? ac_bytes : psum_bytes) + 15u) / 16u * 16u;
}

size_t analyze_smem_bytes(const EncParams& P) {
    return (size_t)P.n_signals * P.smem_stride * 4 + P.pool_bytes + (size_t)P.n_signals * sizeof(WarpScratch) +
           (size_t)P.n_signals * (P.apod_parts * (P.apod_parts + 1) / 2) * kAcStoreStride * sizeof(double) + kMaxSignals * 4 * 4 + 64;
}

}  // namespace fb
