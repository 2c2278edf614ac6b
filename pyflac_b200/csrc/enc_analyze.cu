// enc_analyze.cu -- encode analysis kernel (scope rows E2-E11): one CTA of four warps per frame.
//
// For every signal of the frame (a channel, or mid / side) the kernel reproduces libFLAC 1.4.3's
// process_subframe_ decision sequence: wasted bits, fixed-predictor error sums,
// window + sequential-double autocorrelation, Levinson-Durbin, order guess, coefficient quantisation,
// residual -> partition sums -> Rice parameter / partition-order search, and keeps the candidate with the
// smallest libFLAC bit estimate.  Output: one 128-byte SubframePlan per signal + one channel-assignment
// byte per frame; the pack kernel (enc_pack.cu) turns plans into bits.
//
// Layout of the work (B200: latency/issue bound, so the design minimises instructions and keeps many
// frames resident per SM):
//  * The frame is staged ONCE into shared memory in its container form -- for 16-bit stereo the raw
//    interleaved words (L | R << 16), 16 KiB per 4096-sample frame, mid/side derived on the fly; for every
//    other shape one int32 row block per signal.  Rows of B0 = ceil(N/32) samples, row stride odd, so both
//    access patterns below are bank-conflict free.
//  * Streaming passes (fixed-predictor sums, residual + partition sums) give lane p the contiguous samples
//    [p*B0, (p+1)*B0): one shared load per sample, the predictor history lives in registers (statically
//    rotated window), partition sums leave the lane through a handful of shared atomics.
//  * The autocorrelation is one DFMA chain per lag in ascending sample order (products of two floats are
//    exact in double, so fma == libFLAC's mul+add).  A lane owns the two chains (2p, 2p+1) of one
//    (signal, window) job, up to four jobs per warp: one warp carries all four signals of a stereo frame.
//  * Work items (autocorrelation groups, per-signal fixed analysis, per (signal, apodization step) LPC
//    evaluation) are handed to the four warps through a shared-memory queue, long items first.
//
// Exactness: integer work is exact; floating point follows fb_math.cuh (unfused, RN).  No tensor cores:
// this is integer / bit-serial work, not a dense contraction.
#include "fb_common.cuh"
#include "fb_math.cuh"

namespace fb {

constexpr int kAnThreads = 128;
constexpr int kAnWarps = kAnThreads / 32;
constexpr int kAcJobsMax = 4;          // (signal, window) jobs carried by one warp
constexpr int kAcRing = 112;           // doubles per job: 16 mirror + 3 slots of 32
constexpr int kAcStoreStride = 14;     // lags kept per (signal, window): 13 + the unused odd partner
constexpr int kMaxSteps = kMaxApodSteps;

// Signals of the packed layout (16-bit stereo staged as raw words L | R << 16); plain layouts index their own row block.
enum SigKind : int { kLo16 = 0, kHi16 = 1, kMid16 = 2, kSide16 = 3, kPlain = 4 };

// How a warp reads one signal out of the staged frame.  Packed: value = (lo * ca + hi * cb) >> sh with
// (ca, cb, sh) = (1,0,w) left, (0,1,w) right, (1,1,1+w) mid, (1,-1,w) side, w = wasted bits -- one code path for
// all four signals keeps the instruction working set of a CTA (whose warps run different signals) small.
struct SigView {
    const int32_t* base;     // packed: the frame's words; plain: the signal's own row block (wasted bits already removed)
    int ca, cb, sh;
};

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
        const int lo = (int)(short)w, hi = w >> 16;
        return (lo * V.ca + hi * V.cb) >> V.sh;
    } else {
        return w;
    }
}

struct __align__(16) WarpScratch {
    double   ac[16];                   // autocorrelation of the current apodization step
    double   lperr[kMaxOrder];         // Levinson error per order
    double   lpc[kMaxOrder];           // Levinson recursion state
    float    lp[kMaxOrder * kMaxOrder];
    int32_t  q[16];                    // quantised coefficients of the current candidate
    int32_t  misc[8];
};

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// r[i] = x[i] - ((sum_j q[j] * x[i-1-j]) >> shift), sum of |r| per partition at the maximum partition order.
// Fixed predictors are the same formula with binomial coefficients and shift 0 (up: fixed.c
// FLAC__fixed_compute_residual == lpc residual with q = {1},{2,-1},{3,-3,1},{4,-6,4,-1}).
// WIDE: 64-bit accumulate (up: FLAC__lpc_compute_residual_from_qlp_coefficients_wide); the 32-bit form is
// used exactly when libFLAC proves it cannot overflow.  check_limit reproduces the _limit_residual rejection.
//
// Lane p walks its own contiguous samples; h[] is the predictor history, rotated statically: inside a group
// of ORDER samples, sample u reads h[(u-1-j) mod ORDER] and then overwrites h[u] (the tap that just expired).
// psum must be zeroed by the caller; partition totals arrive through shared atomics (<= 3 per lane).
template <int ORDER, bool WIDE, bool PACKED>
__device__ __noinline__ bool residual_partition_sums(SigView V, FrameGeo G, const int32_t* __restrict__ qs, int shift,
                                                     int psize, bool check_limit, unsigned long long* psum, int lane) {
    constexpr int O = ORDER > 0 ? ORDER : 1;
    int32_t q[O];
#pragma unroll
    for (int j = 0; j < ORDER; j++) q[j] = qs[j];
    bool bad = false;
    const int blk_lo = lane * G.B0, blk_hi = min(G.N, blk_lo + G.B0);
    int lo = max(blk_lo, ORDER);
    const int32_t* rowp = V.base + lane * G.RS - blk_lo;      // rowp[i] is sample i for blk_lo <= i < blk_hi
    while (lo < blk_hi) {
        const int part = lo / psize;
        const int hi = min(blk_hi, (part + 1) * psize);
        int32_t h[O];
#pragma unroll
        for (int k = 0; k < ORDER; k++) h[k] = sig_word<PACKED>(V.base[pidx(G, lo - ORDER + k)], V);
        unsigned long long acc = 0;
        for (int g = lo; g < hi; g += O) {
#pragma unroll
            for (int u = 0; u < O; u++) {
                if (g + u < hi) {
                    const int xv = sig_word<PACKED>(rowp[g + u], V);
                    long long r;
                    if (WIDE) {
                        long long s = 0;
#pragma unroll
                        for (int j = 0; j < ORDER; j++) s += (long long)q[j] * (long long)h[(u - 1 - j + 2 * O) % O];
                        r = (long long)xv - (s >> shift);
                        if (check_limit && (r <= (long long)INT32_MIN || r > (long long)INT32_MAX)) bad = true;
                    } else {
                        int s = 0;
#pragma unroll
                        for (int j = 0; j < ORDER; j++) s += q[j] * h[(u - 1 - j + 2 * O) % O];
                        r = (long long)(xv - (s >> shift));
                    }
                    acc += (unsigned long long)(r < 0 ? -r : r);
                    if (ORDER > 0) h[u] = xv;
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
    switch (order) {
        case 0: return residual_partition_sums<0, WIDE, PACKED>(V, G, q, shift, psize, limit, psum, lane);
        case 1: return residual_partition_sums<1, WIDE, PACKED>(V, G, q, shift, psize, limit, psum, lane);
        case 2: return residual_partition_sums<2, WIDE, PACKED>(V, G, q, shift, psize, limit, psum, lane);
        case 3: return residual_partition_sums<3, WIDE, PACKED>(V, G, q, shift, psize, limit, psum, lane);
        case 4: return residual_partition_sums<4, WIDE, PACKED>(V, G, q, shift, psize, limit, psum, lane);
        case 5: return residual_partition_sums<5, WIDE, PACKED>(V, G, q, shift, psize, limit, psum, lane);
        case 6: return residual_partition_sums<6, WIDE, PACKED>(V, G, q, shift, psize, limit, psum, lane);
        case 7: return residual_partition_sums<7, WIDE, PACKED>(V, G, q, shift, psize, limit, psum, lane);
        case 8: return residual_partition_sums<8, WIDE, PACKED>(V, G, q, shift, psize, limit, psum, lane);
        case 9: return residual_partition_sums<9, WIDE, PACKED>(V, G, q, shift, psize, limit, psum, lane);
        case 10: return residual_partition_sums<10, WIDE, PACKED>(V, G, q, shift, psize, limit, psum, lane);
        case 11: return residual_partition_sums<11, WIDE, PACKED>(V, G, q, shift, psize, limit, psum, lane);
        default: return residual_partition_sums<12, WIDE, PACKED>(V, G, q, shift, psize, limit, psum, lane);
    }
}

template <bool PACKED>
__device__ __forceinline__ bool residual_dispatch(bool wide, int order, const SigView& V, const FrameGeo& G, const int32_t* q,
                                                  int shift, int psize, int nparts, bool limit, unsigned long long* psum, int lane) {
    for (int p = lane; p < nparts; p += 32) psum[p] = 0ull;
    __syncwarp();
    return wide ? residual_order<true, PACKED>(order, V, G, q, shift, psize, limit, psum, lane)
                : residual_order<false, PACKED>(order, V, G, q, shift, psize, limit, psum, lane);
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

template <bool PACKED>
__device__ __forceinline__ void fixed_error_sums_rt(const SigView& V, const FrameGeo& G, int sbps, int lane, unsigned long long* e) {
    const int room = 32 - (sbps + 4);
    const int flush = room >= 6 ? 64 : (room >= 1 ? (1 << room) : 1);
    fixed_error_sums<PACKED>(V, G, flush, lane, e);
}

// ------------------------------------------------------------------------------------------------
// Sequential double autocorrelation (up: lpc.c FLAC__lpc_compute_autocorrelation) of
// windowed segments (up: FLAC__lpc_window_data / _partial) for a group of up to four
// (signal, window) jobs of equal length.  Lane = (job, pair): it owns the chains of lags 2*pair and
// 2*pair+1; the second chain reuses the first one's lagged operand of the previous step, so two steps cost
// one 16-byte shared load for the current values, one for the lagged values and four DFMAs.
// Each chain is strictly sequential in i (one rounding per add, ascending i).
//
// Per job a ring of three 32-sample slots of doubles (plus a 16-entry mirror of slot 2's tail in front of
// slot 0, so "i - lag" is always a plain negative offset).  While the 32 steps of chunk c run, the same warp
// windows chunk c+1 (f32 multiply, widen to f64) into the next slot.
//   segment sample i:  i <  part          : x[off+i] * w[i]
//                      part <= i < 2*part : x[off+i] * w[N-2*part+i]
//                      i == 2*part        : 0            (full window: part = N)
struct AcGroup {
    SigView v[kAcJobsMax];
    int off[kAcJobsMax];
    int cnt;
};

template <bool PACKED>
__device__ __forceinline__ void ac_fetch(const AcGroup& J, const FrameGeo& G, const float* __restrict__ w, int part, int len,
                                         int base, int lane, float& wv, int (&xv)[kAcJobsMax]) {
    const int i = base + lane;
    const bool in = (i < len) && (i < 2 * part);
    wv = 0.0f;
#pragma unroll
    for (int b = 0; b < kAcJobsMax; b++) xv[b] = 0;
    if (in) {
        wv = __ldg(w + (i < part ? i : G.N - 2 * part + i));
#pragma unroll
        for (int b = 0; b < kAcJobsMax; b++) if (b < J.cnt) xv[b] = sig_word<PACKED>(J.v[b].base[pidx(G, J.off[b] + i)], J.v[b]);
    }
}

// windowed samples of one chunk -> ring slot `slot` (slot 2 also feeds the mirror in front of slot 0)
__device__ __forceinline__ void ac_store(double* __restrict__ buf, int cnt, int slot, float wv, const int (&xv)[kAcJobsMax], int lane) {
#pragma unroll
    for (int b = 0; b < kAcJobsMax; b++) {
        if (b < cnt) {
            const double d = (double)FB_FMUL(__int2float_rn(xv[b]), wv);
            buf[b * kAcRing + 16 + slot * 32 + lane] = d;
            if (slot == 2 && lane >= 16) buf[b * kAcRing + lane - 16] = d;
        }
    }
}

template <bool PACKED>
__device__ __noinline__ void autoc_group(const AcGroup& J, FrameGeo G, const float* __restrict__ w, int part, int len, int pairs,
                                         double* __restrict__ buf, int lane, double& out_a, double& out_b) {
    const int jb = lane / pairs, pr = lane - jb * pairs;
    const bool active = jb < J.cnt;
    const double* jobbuf = buf + (active ? jb : 0) * kAcRing;
    const int lag2 = active ? 2 * pr : 0;
    for (int idx = lane; idx < J.cnt * 16; idx += 32) buf[(idx >> 4) * kAcRing + (idx & 15)] = 0.0;
    {
        float wv; int xv[kAcJobsMax];
        ac_fetch<PACKED>(J, G, w, part, len, 0, lane, wv, xv);
        ac_store(buf, J.cnt, 0, wv, xv, lane);
    }
    __syncwarp();
    double acc_a = 0.0, acc_b = 0.0;
    const int nchunks = (len + 31) >> 5;
    int slot = 0;
    for (int c = 0; c < nchunks; c++) {
        float wv; int xv[kAcJobsMax];
        ac_fetch<PACKED>(J, G, w, part, len, (c + 1) * 32, lane, wv, xv);   // inputs of the next chunk: latency hides under the chains
        const double* curp = jobbuf + 16 + slot * 32;
        const double* lagp = curp - lag2;
        double prev = lagp[-1];
#pragma unroll
        for (int s = 0; s < 32; s += 2) {
            const double2 c2 = *reinterpret_cast<const double2*>(curp + s);
            const double2 l2 = *reinterpret_cast<const double2*>(lagp + s);
            acc_a = fma(c2.x, l2.x, acc_a);
            acc_b = fma(c2.x, prev, acc_b);
            acc_a = fma(c2.y, l2.y, acc_a);
            acc_b = fma(c2.y, l2.x, acc_b);
            prev = l2.y;
        }
        slot = (slot == 2) ? 0 : slot + 1;
        ac_store(buf, J.cnt, slot, wv, xv, lane);
        __syncwarp();
    }
    out_a = acc_a; out_b = acc_b;
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

template <typename PcmT, bool PACKED>
__global__ void __launch_bounds__(kAnThreads, PACKED ? 8 : 3)
analyze_kernel(const PcmT* __restrict__ pcm, const FrameDesc* __restrict__ frames, const float* __restrict__ windows,
               EncParams P, SubframePlan* __restrict__ plans, uint8_t* __restrict__ frame_ca,
               SignalDebug* __restrict__ dbg, EncStats* __restrict__ stats, int pass) {
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
    const int n_steps = (int)P.ac_gsz;                                     // apodization steps per signal
    const int nwin = (int)(P.apod_parts * (P.apod_parts + 1) / 2);
    int32_t* xall = reinterpret_cast<int32_t*>(smem_raw);
    unsigned char* cur = smem_raw + (size_t)(PACKED ? 1 : nsig) * sig_words * 4;
    double* acbuf_all = reinterpret_cast<double*>(cur);                          // phase A: autocorrelation rings ...
    WarpScratch* wsall = reinterpret_cast<WarpScratch*>(cur);                    cur += (size_t)P.pool_bytes;   // ... phase B: LPC scratch
    double* acstore = reinterpret_cast<double*>(cur);                            cur += (size_t)nsig * nwin * kAcStoreStride * sizeof(double);
    unsigned long long* psum_all = reinterpret_cast<unsigned long long*>(cur);   cur += (size_t)kAnWarps * 2 * kMaxParts * 8;
    SubframePlan* base_plan = reinterpret_cast<SubframePlan*>(cur);              cur += (size_t)nsig * sizeof(SubframePlan);
    SubframePlan* step_plan = reinterpret_cast<SubframePlan*>(cur);              cur += (size_t)nsig * n_steps * sizeof(SubframePlan);
    AnShared& S = *reinterpret_cast<AnShared*>(cur);

    unsigned long long* psum = psum_all + (size_t)warp * 2 * kMaxParts;
    WarpScratch& ws = wsall[warp];

    if (tid < kMaxSignals) { S.sig_or[tid] = 0u; S.sig_and[tid] = 0xffffffffu; S.best_bits[tid] = 0u; }
    if (tid == 0) { S.queue_a = 0; S.queue_b = 0; S.nneed = 0; }
    for (int i = tid; i < kMaxSignals * kMaxSteps; i += kAnThreads) (&S.step_bits[0][0])[i] = 0xffffffffu;
    __syncthreads();

    // =================== stage the frame; OR / AND of every signal (wasted bits, constant detection) ===================
    // up: process_subframes_ + get_wasted_bits_
    {
        const PcmT* base = pcm + fd.pcm_off;
        if (PACKED) {
            uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0, a0 = ~0u, a1 = ~0u, a2 = ~0u, a3 = ~0u;
            const bool aligned = ((reinterpret_cast<uintptr_t>(base) & 3u) == 0);
            for (int i = tid; i < N; i += kAnThreads) {
                int wd;
                if (aligned) wd = __ldg(reinterpret_cast<const int*>(base) + i);
                else wd = (int)((uint32_t)(uint16_t)__ldg(base + 2 * i) | ((uint32_t)(uint16_t)__ldg(base + 2 * i + 1) << 16));
                xall[pidx(G, i)] = wd;
                const int lo = (int)(short)wd, hi = wd >> 16, m = (lo + hi) >> 1, sd = lo - hi;
                o0 |= (uint32_t)lo; o1 |= (uint32_t)hi; o2 |= (uint32_t)m; o3 |= (uint32_t)sd;
                a0 &= (uint32_t)lo; a1 &= (uint32_t)hi; a2 &= (uint32_t)m; a3 &= (uint32_t)sd;
            }
            o0 = __reduce_or_sync(0xffffffffu, o0); o1 = __reduce_or_sync(0xffffffffu, o1);
            o2 = __reduce_or_sync(0xffffffffu, o2); o3 = __reduce_or_sync(0xffffffffu, o3);
            a0 = __reduce_and_sync(0xffffffffu, a0); a1 = __reduce_and_sync(0xffffffffu, a1);
            a2 = __reduce_and_sync(0xffffffffu, a2); a3 = __reduce_and_sync(0xffffffffu, a3);
            if (lane == 0) {
                atomicOr(&S.sig_or[0], o0); atomicOr(&S.sig_or[1], o1); atomicAnd(&S.sig_and[0], a0); atomicAnd(&S.sig_and[1], a1);
                if (nsig > 2) { atomicOr(&S.sig_or[2], o2); atomicOr(&S.sig_or[3], o3); atomicAnd(&S.sig_and[2], a2); atomicAnd(&S.sig_and[3], a3); }
            }
        } else {
            for (int s = 0; s < nsig; s++) {
                if (!sig_active(s)) continue;
                int32_t* x = xall + (size_t)s * sig_words;
                uint32_t o = 0, a = ~0u;
                for (int i = tid; i < N; i += kAnThreads) {
                    int v;
                    if (s < ch) v = (int)__ldg(base + (uint64_t)i * ch + s);
                    else {
                        const int l = (int)__ldg(base + (uint64_t)i * ch), r = (int)__ldg(base + (uint64_t)i * ch + 1);
                        v = (s == ch) ? ((l + r) >> 1) : (l - r);
                    }
                    x[pidx(G, i)] = v;
                    o |= (uint32_t)v; a &= (uint32_t)v;
                }
                o = __reduce_or_sync(0xffffffffu, o); a = __reduce_and_sync(0xffffffffu, a);
                if (lane == 0) { atomicOr(&S.sig_or[s], o); atomicAnd(&S.sig_and[s], a); }
            }
        }
    }
    __syncthreads();

    // per-signal facts every thread can derive on its own
    auto sig_wasted = [&](int s) { const uint32_t o = S.sig_or[s]; const int wst = o ? (__ffs((int)o) - 1) : 0; return wst > (int)P.bps ? (int)P.bps : wst; };
    auto sig_sbps = [&](int s) { return (int)P.bps - sig_wasted(s) + ((P.do_mid_side && s == ch + 1) ? 1 : 0); };
    auto sig_view = [&](int s) {
        SigView V;
        if (PACKED) { V.base = xall; V.ca = (s != 1); V.cb = (s == 0) ? 0 : (s == 3 ? -1 : 1); V.sh = sig_wasted(s) + (s == 2 ? 1 : 0); }
        else { V.base = xall + (size_t)s * sig_words; V.ca = 1; V.cb = 0; V.sh = 0; }
        return V;
    };
    if (!PACKED) {      // plain rows are stored with the wasted bits already removed
        for (int s = 0; s < nsig; s++) {
            const int wst = sig_active(s) ? sig_wasted(s) : 0;
            if (wst) { int32_t* x = xall + (size_t)s * sig_words; for (int i = tid; i < sig_words; i += kAnThreads) x[i] >>= wst; }
        }
    }
    // up: process_subframe_ constant test + process_subframes_ limit_min_bitrate: when every earlier channel is
    // constant, the last channel (and mid/side after it) may not use a constant subframe
    auto sig_const = [&](int s) { return N > 4 && S.sig_or[s] == S.sig_and[s]; };
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

    // =================== phase A: autocorrelation groups (long) and per-signal fixed analysis, from a queue ===================
    const int L = max_lpc + 1, pairs = (L + 1) >> 1;
    const int gmax = min(kAcJobsMax, 32 / pairs);
    // group enumeration (identical on every warp): for each window depth b, the nneed*b jobs in balanced groups
    int n_groups = 0;
    if (nneed > 0)
        for (int b = 1; b <= (int)P.apod_parts; b++) {
            if (b > 1 && N / b <= 32) continue;
            n_groups += (nneed * b + gmax - 1) / gmax;
        }
    const int n_tasks_a = n_groups + nsig;
    const float* wtab = windows + fd.window_off;
    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&S.queue_a, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_tasks_a) break;
        if (t < n_groups) {
            // ---- locate group t ----
            int b = 1, g0 = 0, j0 = 0, gsz = 1, nj = 0;
            for (;; b++) {
                if (b > 1 && N / b <= 32) continue;
                nj = nneed * b;
                const int ngrp = (nj + gmax - 1) / gmax;
                gsz = (nj + ngrp - 1) / ngrp;
                if (t < g0 + ngrp) { j0 = (t - g0) * gsz; break; }
                g0 += ngrp;
            }
            const int len = N / b, part = (b == 1) ? N : N / b / 2;
            AcGroup J;
            J.cnt = min(gsz, nj - j0);
#pragma unroll
            for (int q = 0; q < kAcJobsMax; q++) {
                const int j = min(j0 + q, nj - 1), sidx = S.need_list[j / b], k = j - (j / b) * b;
                J.v[q] = sig_view(sidx); J.off[q] = (k * N) / b;
            }
            double* buf = acbuf_all + (size_t)(n_groups < kAnWarps ? t : warp) * kAcJobsMax * kAcRing;
            double ra, rb;
            autoc_group<PACKED>(J, G, wtab, part, len, pairs, buf, lane, ra, rb);
            const int jb = lane / pairs, pr = lane - jb * pairs;
            if (jb < J.cnt) {
                const int j = j0 + jb, sidx = S.need_list[j / b], k = j - (j / b) * b;
                double* dst = acstore + ((size_t)sidx * nwin + (b - 1) * b / 2 + k) * kAcStoreStride;
                dst[2 * pr] = ra; dst[2 * pr + 1] = rb;
            }
            __syncwarp();
        } else {
            // ---- fixed analysis of signal s: verbatim baseline, constant, or the guessed fixed order ----
            const int s = t - n_groups;
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
                    fixed_error_sums_rt<PACKED>(V, G, sbps, lane, e);
                    if ((uint32_t)sbps + ilog2_u32((uint32_t)N - 4u) + 1u < 32u) {   // libFLAC's 32-bit accumulators wrap
#pragma unroll
                        for (int k = 0; k < 5; k++) e[k] &= 0xffffffffull;
                    }
                    int forder;
                    {
                        const unsigned long long m34 = min(e[3], e[4]), m234 = min(e[2], m34), m1234 = min(e[1], m234);
                        if (e[0] <= m1234) forder = 0; else if (e[1] <= m234) forder = 1; else if (e[2] <= m34) forder = 2; else if (e[3] <= e[4]) forder = 3; else forder = 4;
                    }
                    if (dg && lane == 0) { for (int k = 0; k < 5; k++) dg->fixed_err[k] = e[k]; dg->fixed_order = forder; dg->is_constant = constant; }
                    // fixed candidate at the guessed order (up: evaluate_fixed_subframe_)
                    int fo = forder; if (fo >= N) fo = N - 1;
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

// host-visible launcher (called from engine.cu)
void launch_analyze(const void* pcm, const FrameDesc* frames, const float* windows, const EncParams& P, int n_frames,
                    SubframePlan* plans, uint8_t* frame_ca, SignalDebug* dbg, EncStats* stats, size_t smem_bytes,
                    cudaStream_t stream) {
    const dim3 grid((unsigned)n_frames), block(kAnThreads);
    // loose mid/side: decision frames first, then the frames that follow them (they read the decision from frame_ca)
    for (int pass = 0; pass < (P.loose_frames ? 2 : 1); pass++) {
        if (packed_layout(P)) {
            cudaFuncSetAttribute(analyze_kernel<int16_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
            analyze_kernel<int16_t, true><<<grid, block, smem_bytes, stream>>>((const int16_t*)pcm, frames, windows, P, plans, frame_ca, dbg, stats, pass);
        } else if (P.container_bytes == 2) {
            cudaFuncSetAttribute(analyze_kernel<int16_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
            analyze_kernel<int16_t, false><<<grid, block, smem_bytes, stream>>>((const int16_t*)pcm, frames, windows, P, plans, frame_ca, dbg, stats, pass);
        } else {
            cudaFuncSetAttribute(analyze_kernel<int32_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
            analyze_kernel<int32_t, false><<<grid, block, smem_bytes, stream>>>((const int32_t*)pcm, frames, windows, P, plans, frame_ca, dbg, stats, pass);
        }
    }
}

// Shared-memory plan of the analysis kernel; fills P.pool_bytes (autocorrelation rings), P.ac_gsz (apodization
// steps per signal) and P.an_stride (staged words per signal).  Called by the host before launching.
void analyze_layout(EncParams& P) {
    const uint32_t nsig = P.n_signals, L = (P.max_lpc_order ? P.max_lpc_order : 1) + 1, pairs = (L + 1) / 2;
    const uint32_t gmax = (32 / pairs) < (uint32_t)kAcJobsMax ? (32 / pairs) : (uint32_t)kAcJobsMax;
    uint32_t groups = 0;
    for (uint32_t b = 1; b <= P.apod_parts; b++) groups += (nsig * b + gmax - 1) / gmax;
    const uint32_t areas = groups < (uint32_t)kAnWarps ? groups : (uint32_t)kAnWarps;
    const uint32_t ring_bytes = P.max_lpc_order ? areas * kAcJobsMax * kAcRing * 8u : 0u, ws_bytes = (uint32_t)(kAnWarps * sizeof(WarpScratch));
    P.pool_bytes = ring_bytes > ws_bytes ? ring_bytes : ws_bytes;     // rings (phase A) and LPC scratch (phase B) share the pool
    P.ac_gsz = apod_steps(P);
    // 32 rows of ceil(N/32) samples with an odd row stride
    const uint32_t b0 = (P.blocksize + 31) / 32, rs = b0 | 1u;
    P.an_stride = ((32u * rs + 3u) / 4u) * 4u;
}

size_t analyze_smem_bytes(const EncParams& P) {
    const size_t nsig = P.n_signals, nwin = P.apod_parts * (P.apod_parts + 1) / 2;
    return (size_t)(packed_layout(P) ? 1 : nsig) * P.an_stride * 4 + P.pool_bytes + nsig * nwin * kAcStoreStride * sizeof(double) +
           (size_t)kAnWarps * 2 * kMaxParts * 8 + nsig * sizeof(SubframePlan) +
           nsig * P.ac_gsz * sizeof(SubframePlan) + sizeof(AnShared) + 64;
}

}  // namespace fb
