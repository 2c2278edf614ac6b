// enc_fused.cu -- the whole per-frame encode path in ONE kernel for 16-bit stereo (the headline layout):
//
//   stage (TMA) -> OR/AND -> [autocorrelation || fixed predictors] -> LPC candidates -> choose -> pack -> CRC-16 -> store
//
// One CTA of eight warps per frame, several CTAs resident per SM.  The frame -- 16 KiB of interleaved int16 pairs for a
// 4096-sample block -- crosses HBM exactly once: warp 0 issues one cp.async.bulk (TMA, SASS UBLKCP) per tile row, the
// copies complete on an mbarrier, and every later phase (scope rows E2-E12) works on the shared-memory tile.  Nothing
// but the finished frame bytes and 4 bytes of length goes back to HBM (the multi-kernel path of enc_analyze.cu /
// enc_pack.cu re-reads the PCM four times and round-trips plans and autocorrelations; it remains the path for the other
// layouts, loose mid/side and debug traces).
//
// Tile layout: 32 rows of B0w = roundup4(ceil(N/32)) packed words (L | R << 16), row stride RS = B0w (+4 so that RS/4
// is odd).  Lane p of an analysis warp streams row p with 16-byte shared loads: eight consecutive lanes hit eight
// different 16-byte bank groups, so the quarter-warp wavefronts of LDS.128 are conflict free, and the row starts stay
// 16-byte aligned, which is what lets TMA write them.
//
// Phases that cannot fill a CTA overlap: the autocorrelation (a handful of strictly sequential f64 chains, latency
// bound) runs on one or a few warps while the other warps do the integer fixed-predictor analysis; the other CTAs of
// the SM are in different phases and fill the issue slots.
//
// Instruction diet against the multi-kernel path (which was issue bound):
//  * fixed predictors: ONE pass yields the five error sums AND the per-partition sums of all five orders; the
//    residual pass of the chosen order is gone (|k-th difference| is the order-k residual);
//  * autocorrelation: a lane owns two lags (not four) of one signal, so a frame's 36 chains use 20 lanes of one warp;
//  * pack: a lane codes 16 consecutive samples; its bits form one contiguous run that is assembled in a register
//    and stored word by word -- only the first and last word of a run are shared with a neighbour and need atomicOr.
//
// Exactness: integer work is exact; floating point follows fb_math.cuh (unfused, RN); each autocorrelation chain
// accumulates in ascending sample order.  No tensor cores: this is integer / bit-serial work.
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <vector>
#include <stdio.h>
#include <stdlib.h>
#include "fb_common.cuh"
#include "fb_math.cuh"
#include "enc_dev.cuh"

namespace fb {

constexpr int kFuThreads = 256;
constexpr int kFuWarps = kFuThreads / 32;
constexpr int kFuSig = 4;                  // L, R, mid, side
constexpr int kWinSlots = 16;              // 32-float slots of the window-value ring (cp.async destination)
constexpr int kWinAhead = 8;               // chunks a window value is requested ahead of its use
constexpr int kFuSlots = 4;                // 32-sample slots in the autocorrelation ring (power of two)
constexpr int kFuRing = 16 + 32 * kFuSlots + 2;   // doubles per autocorrelation job: 16 mirror + the slots (+2: jobs land in different banks)
constexpr int kFuRun = 16;                 // consecutive samples a lane codes in the pack phase
constexpr int kFuChunk = 32 * kFuRun;      // samples a warp codes per round

// Shared-memory plan, computed on the host (fused_layout) and passed by value.  The `ovl` region is used twice:
// analysis scratch first, the frame image afterwards.
struct FuLayout {
    uint32_t tile_words;
    uint32_t ovl_off, ovl_bytes;
    uint32_t ring_off, wring_off, acstore_off, ws_off, psum_off, fixsum_off, baseplan_off, stepplan_off;   // inside ovl
    uint32_t shared_off, crctab_off, total_bytes;
    uint32_t obuf_words, n_win, n_steps, pad;
};

struct FuShared {
    unsigned long long mbar;
    uint32_t sig_or[kFuSig], sig_and[kFuSig];
    uint32_t best_bits[kFuSig];
    uint32_t step_bits[kFuSig][kMaxSteps];
    int      need_list[kFuSig];
    int      nneed, queue_a, queue_b, ca;
    int      ac_warp, pad2;                // the warp that runs the first autocorrelation item (rotates per SM, see g_fu_ticket)
    SubframePlan plan[2];                  // the two coded subframes
    int32_t  sigidx[2];
    uint32_t segtot[kFuWarps];
    uint32_t crc_warp[kFuWarps];
    uint8_t  hdr[16];
    uint32_t hdr_len, pad;
};

struct FuGeo { int N, B0w, RS, gap; uint32_t magic; };
// word index of sample i in the tile: row * RS + col
__device__ __forceinline__ int tix(const FuGeo& G, int i) { return i + (int)__umulhi((uint32_t)i, G.magic) * G.gap; }

// one signal of the packed frame: value = (lo * ca + hi * cb) >> sh (one IDP.2A + one shift)
struct Sig { int cab, sh; };
__device__ __forceinline__ Sig make_sig(int s, int wasted) {
    const int ca = (s != 1), cb = (s == 0) ? 0 : (s == 3 ? -1 : 1);
    Sig g; g.cab = (ca & 0xff) | ((cb & 0xff) << 8); g.sh = wasted + (s == 2 ? 1 : 0);
    return g;
}
__device__ __forceinline__ int sv(int w, const Sig& g) { return __dp2a_lo(w, g.cab, 0) >> g.sh; }

// ------------------------------------------------------------------------------------------------ TMA / mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP.S.G)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------------ autocorrelation
// up: lpc.c FLAC__lpc_window_data[_partial] + FLAC__lpc_compute_autocorrelation.
// One item = one window (depth b, position k) of the frame, nsig jobs (signals) that share the sample word and the
// window value.  Lane = (job, lag pair): chains of lags 2q and 2q+1, each strictly sequential in i (fma of two exact
// float products == libFLAC's mul + add).  Per job a ring of four 32-sample slots of doubles with a 16-entry mirror of
// the last slot's tail in front of slot 0, so "i - lag" is a plain negative offset.  Window values arrive through a
// cp.async ring eight chunks ahead; chunk c+1 is converted before the chains of chunk c run.
// ROLE 0: one warp converts and runs the chains (levels with many windows).  ROLE 1 / 2: a pair of warps -- the chain warp
// (1) only runs the DFMA chains, its partner (2) windows and converts one chunk ahead into the ring; they meet once per
// 32-sample chunk at a named barrier (bar_id, 64 threads).  A warp issues in order, so a lone warp pays the conversion's
// dependent steps (shared load -> int -> float -> multiply -> double -> store, ~150 cycles) in front of every chunk's
// 260 cycles of chain latency; the pair hides them (measured: 1050 -> ~300 cycles per chunk).
template <int ROLE>
__device__ __noinline__ void fu_autoc_item(const int32_t* __restrict__ tile, const FuGeo G, const float* __restrict__ win_tab,
                                           int b, int k, int nsig, int lags, const uint32_t* __restrict__ sig_or, int bps,
                                           double* __restrict__ ring, float* __restrict__ wring, double* __restrict__ acstore, int n_win,
                                           int lane, int bar_id) {
    const int N = G.N;
    const int len = N / b, part = (b == 1) ? N : N / b / 2, off = (k * N) / b;
    const int wtail = N - 2 * part;                   // window index = i (i < part) or wtail + i (part <= i < 2*part)
    const int LJ = (lags + 1) >> 1;                   // lanes per job
    const int nchunks = (len + 31) >> 5;
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" :: "r"(bar_id) : "memory"); };

    if (ROLE != 1) {
        int shp[kFuSig];
#pragma unroll
        for (int s = 0; s < kFuSig; s++) shp[s] = (s < nsig) ? wasted_from_or(sig_or[s], bps) + (s == 2 ? 1 : 0) : 0;
        for (int idx = lane; idx < nsig * 16; idx += 32) ring[(idx >> 4) * kFuRing + (idx & 15)] = 0.0;
        // The window table lives in global memory (16 KiB per blocksize, shared by every frame; the SM's L1 is mostly carved
        // into shared memory here, so a read is an L2 round trip of several hundred cycles).  Its values travel straight into
        // a small shared ring with cp.async -- no destination register, so nothing waits on them -- kWinAhead chunks ahead
        // of their use; samples outside the window read as zero (src-size 0 zero-fills).
        const uint32_t wr_base = smem_u32(wring) + (uint32_t)lane * 4u;
        auto request = [&](int c) {
            const int i = c * 32 + lane;
            const bool in = (i < len && i < 2 * part);
            const float* src = win_tab + (in ? (i < part ? i : wtail + i) : 0);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(wr_base + (uint32_t)(c & (kWinSlots - 1)) * 128u), "l"(src), "r"(in ? 4u : 0u) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto convert = [&](int c) {
            const int slot = c & (kFuSlots - 1);
            const int i = c * 32 + lane;
            float rwv;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(rwv) : "r"(wr_base + (uint32_t)(c & (kWinSlots - 1)) * 128u) : "memory");
            const int rw = (i < len && i < 2 * part) ? tile[tix(G, off + i)] : 0;
            const int lo = (int)(short)rw, hi = rw >> 16;
            float dv[kFuSig];
            dv[0] = FB_FMUL(__int2float_rn(lo >> shp[0]), rwv);
            dv[1] = FB_FMUL(__int2float_rn(hi >> shp[1]), rwv);
            dv[2] = FB_FMUL(__int2float_rn((lo + hi) >> shp[2]), rwv);
            dv[3] = FB_FMUL(__int2float_rn((lo - hi) >> shp[3]), rwv);
#pragma unroll
            for (int s = 0; s < kFuSig; s++) {
                if (s < nsig) {
                    const double d = (double)dv[s];
                    ring[s * kFuRing + 16 + slot * 32 + lane] = d;
                    if (slot == kFuSlots - 1 && lane >= 16) ring[s * kFuRing + lane - 16] = d;      // the last slot's tail mirrored in front of slot 0
                }
            }
        };
#pragma unroll 1
        for (int c = 0; c < kWinAhead; c++) request(c);
        asm volatile("cp.async.wait_group %0;" :: "n"(kWinAhead - 1) : "memory");       // chunk 0 has landed
        convert(0);
        if (ROLE == 2) {
            // converter of a pair: chunk c+1 is ready before the chain warp passes barrier c.  While the chains of chunk c
            // read slot c (and the tail of slot c-1, the mirror only when slot == 0) this warp fills slot c+2.
            request(kWinAhead);
            asm volatile("cp.async.wait_group %0;" :: "n"(kWinAhead - 1) : "memory");
            convert(1);
#pragma unroll 1
            for (int c = 0; c < nchunks; c++) {
                pair_sync();
                request(c + 1 + kWinAhead);
                asm volatile("cp.async.wait_group %0;" :: "n"(kWinAhead - 1) : "memory");
                convert(c + 2);
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            return;
        }
        __syncwarp();
        // ROLE 0 continues below with the chains, converting chunk c+1 in front of the chains of chunk c
        const int jb = lane / LJ, qd = lane - jb * LJ;
        const bool active = jb < nsig;
        const int lag0 = 2 * qd;
        const double* jobring = ring + (active ? jb : 0) * kFuRing;
        double a0 = 0.0, a1 = 0.0, p1 = 0.0;
#pragma unroll 1
        for (int c = 0; c < nchunks; c++) {
            const int slot = c & (kFuSlots - 1);
            request(c + kWinAhead);
            asm volatile("cp.async.wait_group %0;" :: "n"(kWinAhead - 1) : "memory");   // chunk c+1's window values have landed
            convert(c + 1);
            if (active) {
                const double* curp = jobring + 16 + slot * 32;
                const double* lagp = curp - lag0;
#pragma unroll
                for (int s = 0; s < 32; s += 2) {
                    const double2 c2 = *reinterpret_cast<const double2*>(curp + s);
                    const double2 l2 = *reinterpret_cast<const double2*>(lagp + s);
                    a0 = fma(c2.x, l2.x, a0); a1 = fma(c2.x, p1, a1);
                    a0 = fma(c2.y, l2.y, a0); a1 = fma(c2.y, l2.x, a1);
                    p1 = l2.y;
                }
            }
            __syncwarp();
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (active) {
            double* dst = acstore + ((size_t)jb * n_win + (b - 1) * b / 2 + k) * kAcStoreStride + lag0;
            dst[0] = a0;
            if (lag0 + 1 < kAcStoreStride) dst[1] = a1;
        }
        __syncwarp();
        return;
    }
    // ---- ROLE 1: the chain warp of a pair ----
    {
        const int jb = lane / LJ, qd = lane - jb * LJ;
        const bool active = jb < nsig;
        const int lag0 = 2 * qd;
        const double* jobring = ring + (active ? jb : 0) * kFuRing;
        double a0 = 0.0, a1 = 0.0, p1 = 0.0;
#pragma unroll 1
        for (int c = 0; c < nchunks; c++) {
            const int slot = c & (kFuSlots - 1);
            pair_sync();                                                   // chunk c (and c+1) converted
            if (active) {
                const double* curp = jobring + 16 + slot * 32;
                const double* lagp = curp - lag0;
#pragma unroll
                for (int s = 0; s < 32; s += 2) {
                    const double2 c2 = *reinterpret_cast<const double2*>(curp + s);
                    const double2 l2 = *reinterpret_cast<const double2*>(lagp + s);
                    a0 = fma(c2.x, l2.x, a0); a1 = fma(c2.x, p1, a1);
                    a0 = fma(c2.y, l2.y, a0); a1 = fma(c2.y, l2.x, a1);
                    p1 = l2.y;
                }
            }
        }
        if (active) {
            double* dst = acstore + ((size_t)jb * n_win + (b - 1) * b / 2 + k) * kAcStoreStride + lag0;
            dst[0] = a0;
            if (lag0 + 1 < kAcStoreStride) dst[1] = a1;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ fixed predictors
// up: fixed.c FLAC__fixed_compute_best_predictor[_wide] and, in the same pass, what
// precompute_partition_info_sums_ would compute from the order-k fixed residual: |k-th difference| IS that residual.
// Lane p streams row p; per (row x partition) segment the five 32-bit partial sums go to fixsum[k][partition]
// (samples 4.. only, exactly the range of libFLAC's error sums; the caller adds the samples k..3 of the chosen order
// to partition 0).  Returns the five totals in e[] (every lane).
template <bool VEC>
__device__ __noinline__ void fu_fixed_sums(const int32_t* __restrict__ tile, const FuGeo G, const Sig sg, int psize,
                                           unsigned long long* __restrict__ fixsum, int lane, unsigned long long* e) {
    const int blk_lo = lane * G.B0w, blk_hi = min(G.N, blk_lo + G.B0w);
    unsigned long long e0 = 0, e1 = 0, e2 = 0, e3 = 0, e4 = 0;
    int lo = max(blk_lo, 4);
    if (lo < blk_hi) {
        const int32_t* rowp = tile + lane * G.RS - blk_lo;               // rowp[i] = word of sample i (this lane's row)
        int x1 = sv(tile[tix(G, lo - 1)], sg), x2 = sv(tile[tix(G, lo - 2)], sg);
        const int x3 = sv(tile[tix(G, lo - 3)], sg), x4 = sv(tile[tix(G, lo - 4)], sg);
        int d1 = x1 - x2, d2 = d1 - (x2 - x3), d3 = d2 - ((x2 - x3) - (x3 - x4));
        while (lo < blk_hi) {
            const int part = lo / psize;
            const int hi = min(blk_hi, (part + 1) * psize);
            uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;             // < 2^(17+4) per term, <= 2048 terms
#define FU_FIXED_STEP(word) { const int x0 = sv((word), sg); const int a1 = x0 - x1, a2 = a1 - d1, a3 = a2 - d2, a4 = a3 - d3; \
            s0 += (uint32_t)abs(x0); s1 += (uint32_t)abs(a1); s2 += (uint32_t)abs(a2); s3 += (uint32_t)abs(a3); s4 += (uint32_t)abs(a4); \
            x1 = x0; d1 = a1; d2 = a2; d3 = a3; }
            if (VEC) {
#pragma unroll 2
                for (int i = lo; i < hi; i += 4) {
                    const int4 w = *reinterpret_cast<const int4*>(rowp + i);
                    FU_FIXED_STEP(w.x) FU_FIXED_STEP(w.y) FU_FIXED_STEP(w.z) FU_FIXED_STEP(w.w)
                }
            } else {
                for (int i = lo; i < hi; i++) FU_FIXED_STEP(rowp[i])
            }
#undef FU_FIXED_STEP
            atomicAdd(&fixsum[0 * kMaxParts + part], (unsigned long long)s0);
            atomicAdd(&fixsum[1 * kMaxParts + part], (unsigned long long)s1);
            atomicAdd(&fixsum[2 * kMaxParts + part], (unsigned long long)s2);
            atomicAdd(&fixsum[3 * kMaxParts + part], (unsigned long long)s3);
            atomicAdd(&fixsum[4 * kMaxParts + part], (unsigned long long)s4);
            e0 += s0; e1 += s1; e2 += s2; e3 += s3; e4 += s4;
            lo = hi;
        }
    }
    e[0] = warp_sum_u64(e0); e[1] = warp_sum_u64(e1); e[2] = warp_sum_u64(e2); e[3] = warp_sum_u64(e3); e[4] = warp_sum_u64(e4);
}

// ------------------------------------------------------------------------------------------------ LPC residual
// r[i] = x[i] - ((sum_j q[j] x[i-1-j]) >> shift); sum of |r| per partition at the maximum partition order
// (up: FLAC__lpc_compute_residual_from_qlp_coefficients + precompute_partition_info_sums_, rows E9 / E10).
// Fast path: 32-bit accumulate (chosen exactly when libFLAC proves it cannot overflow), partitions that are multiples
// of four samples.  Lane p walks row p in groups of C samples (C = order class 4 / 8 / 12, surplus taps carry zero
// coefficients: three code bodies keep the instruction working set small), sixteen bytes per shared load, history
// rotated by register renaming.  psum must be zeroed by the caller.
template <int C>
__device__ __noinline__ void fu_lpc_psums_vec(const int32_t* __restrict__ tile, const FuGeo G, const Sig sg, const int32_t* __restrict__ qs,
                                              int order, int shift, int psize, unsigned long long* __restrict__ psum, int lane) {
    int32_t q[C];
#pragma unroll
    for (int j = 0; j < C; j++) q[j] = (j < order) ? qs[j] : 0;
    const int blk_lo = lane * G.B0w, blk_hi = min(G.N, blk_lo + G.B0w);
    const int32_t* rowp = tile + lane * G.RS - blk_lo;
    int lo = blk_lo;
    if (lane == 0) {
        // the first `order` samples are warm-up; samples up to the next multiple of four are done one by one (they lie in
        // partition 0: a partition is longer than the order and a multiple of four here)
        lo = (order + 3) & ~3;
        unsigned long long head = 0;
        for (int i = order; i < min(lo, G.N); i++) {
            int s = 0;
            for (int j = 0; j < order; j++) s += qs[j] * sv(tile[tix(G, i - 1 - j)], sg);
            const long long r = (long long)(sv(tile[tix(G, i)], sg) - (s >> shift));
            head += (unsigned long long)(r < 0 ? -r : r);
        }
        if (head) atomicAdd(&psum[0], head);
    }
    while (lo < blk_hi) {
        const int part = lo / psize;
        const int hi = min(blk_hi, (part + 1) * psize);
        int32_t h[C];
#pragma unroll
        for (int v = 0; v < C / 4; v++) {                                // the C samples before lo (zero taps before sample 0)
            const int idx = lo - C + 4 * v;
            int4 w = make_int4(0, 0, 0, 0);
            if (idx >= 0) w = *reinterpret_cast<const int4*>(tile + tix(G, idx));
            h[4 * v + 0] = sv(w.x, sg); h[4 * v + 1] = sv(w.y, sg); h[4 * v + 2] = sv(w.z, sg); h[4 * v + 3] = sv(w.w, sg);
        }
        unsigned long long acc = 0;
        for (int g = lo; g < hi; g += C) {
#pragma unroll
            for (int v = 0; v < C / 4; v++) {
                if (g + 4 * v < hi) {
                    const int4 w = *reinterpret_cast<const int4*>(rowp + g + 4 * v);
                    const int xw[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const int u = 4 * v + t;
                        const int xv = sv(xw[t], sg);
                        int s = 0;
#pragma unroll
                        for (int j = 0; j < C; j++) s += q[j] * h[(u - 1 - j + 2 * C) % C];
                        const long long r = (long long)(xv - (s >> shift));
                        acc += (unsigned long long)(r < 0 ? -r : r);
                        h[u] = xv;
                    }
                }
            }
        }
        atomicAdd(&psum[part], acc);
        lo = hi;
    }
    __syncwarp();
}

// Every other case (64-bit accumulate, the int32 residual check, partitions that are not multiples of four): one
// sample per shared load, same statically rotated history.
template <int C, bool WIDE>
__device__ __noinline__ bool fu_lpc_psums_scalar(const int32_t* __restrict__ tile, const FuGeo G, const Sig sg, const int32_t* __restrict__ qs,
                                                 int order, int shift, int psize, bool check_limit, unsigned long long* __restrict__ psum, int lane) {
    int32_t q[C];
#pragma unroll
    for (int j = 0; j < C; j++) q[j] = (j < order) ? qs[j] : 0;
    bool bad = false;
    const int blk_lo = lane * G.B0w, blk_hi = min(G.N, blk_lo + G.B0w);
    int lo = max(blk_lo, order);
    const int32_t* rowp = tile + lane * G.RS - blk_lo;
    while (lo < blk_hi) {
        const int part = lo / psize;
        const int hi = min(blk_hi, (part + 1) * psize);
        int32_t h[C];
#pragma unroll
        for (int k = 0; k < C; k++) h[k] = sv(tile[tix(G, max(lo - C + k, 0))], sg);   // taps before sample 0 carry zero coefficients
        unsigned long long acc = 0;
        for (int g = lo; g < hi; g += C) {
#pragma unroll
            for (int u = 0; u < C; u++) {
                if (g + u < hi) {
                    const int xv = sv(rowp[g + u], sg);
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
                        for (int j = 0; j < C; j++) s += q[j] * h[(u - 1 - j + 2 * C) % C];
                        r = (long long)(xv - (s >> shift));
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

__device__ __forceinline__ bool fu_lpc_psums(const int32_t* tile, const FuGeo& G, const Sig& sg, const int32_t* q, int order, int shift,
                                             int psize, int nparts, bool wide, bool limit, unsigned long long* psum, int lane) {
    for (int p = lane; p < nparts; p += 32) psum[p] = 0ull;
    __syncwarp();
    if (!wide && !limit && (psize & 3) == 0) {
        if (order <= 4) fu_lpc_psums_vec<4>(tile, G, sg, q, order, shift, psize, psum, lane);
        else if (order <= 8) fu_lpc_psums_vec<8>(tile, G, sg, q, order, shift, psize, psum, lane);
        else fu_lpc_psums_vec<12>(tile, G, sg, q, order, shift, psize, psum, lane);
        return false;
    }
    if (wide || limit) {
        if (order <= 4) return fu_lpc_psums_scalar<4, true>(tile, G, sg, q, order, shift, psize, limit, psum, lane);
        if (order <= 8) return fu_lpc_psums_scalar<8, true>(tile, G, sg, q, order, shift, psize, limit, psum, lane);
        return fu_lpc_psums_scalar<12, true>(tile, G, sg, q, order, shift, psize, limit, psum, lane);
    }
    if (order <= 4) return fu_lpc_psums_scalar<4, false>(tile, G, sg, q, order, shift, psize, false, psum, lane);
    if (order <= 8) return fu_lpc_psums_scalar<8, false>(tile, G, sg, q, order, shift, psize, false, psum, lane);
    return fu_lpc_psums_scalar<12, false>(tile, G, sg, q, order, shift, psize, false, psum, lane);
}

// ------------------------------------------------------------------------------------------------ pack
// i / psize through the precomputed reciprocal; magic 0 marks one-sample partitions (the 32-bit reciprocal of 1 does not exist)
__device__ __forceinline__ uint32_t pdiv(uint32_t i, uint32_t magic) { return magic ? __umulhi(i, magic) : i; }

// val < 2^n, 1 <= n <= 32
__device__ __forceinline__ void put_code(uint32_t* buf, uint32_t pos, uint32_t val, uint32_t n) {
    const uint32_t w = pos >> 5, o = pos & 31u;
    if (o + n <= 32u) atomicOr(&buf[w], val << (32u - o - n));
    else { const uint32_t r = o + n - 32u; atomicOr(&buf[w], val >> r); atomicOr(&buf[w + 1u], val << (32u - r)); }
}

// Rice-coded body of one subframe (up: add_residual_partitioned_rice_ + FLAC__bitwriter_write_rice_signed_block).
// Round r: warp w codes samples [(r * 8 + w) * 512, +512), lane l the 16 consecutive samples at + 16 l, four at a time
// (one 16-byte shared load; the predictor history slides through registers).  Pass 1 counts the lane's bits (codes + the
// parameter field of every partition that starts inside its run); an exclusive warp scan and the eight warp totals (one
// CTA barrier per round) give the lane's absolute bit position; pass 2 computes the residuals again and ORs the codes
// into the zeroed image: neighbouring lanes are ~160 bits apart, so the shared atomics rarely meet in one word.  The
// residuals are computed twice instead of parked in registers: the loops stay rolled and small (this phase was
// instruction-fetch bound when it was unrolled over the run).  Returns the body length in bits.
// C = order class, WIDE = 64-bit accumulate (chosen exactly as the analysis does).
template <int C, bool WIDE, bool EMIT>
__device__ __forceinline__ uint32_t fu_pack_run(const SubframePlan& pl, const int32_t (&q)[C], int order, int shift, const int32_t* __restrict__ tile,
                                                const FuGeo& G, const Sig& sg, int i0, uint32_t plen, uint32_t psize, uint32_t pmagic, bool runpart,
                                                uint32_t pos, uint32_t* __restrict__ obuf) {
    const int N = G.N;
    int32_t xw[C];                                     // the C samples before the current quad
#pragma unroll
    for (int v = 0; v < C / 4; v++) {
        const int idx = i0 - C + 4 * v;
        int4 w = make_int4(0, 0, 0, 0);
        if (idx >= 0) w = *reinterpret_cast<const int4*>(tile + tix(G, idx));
        xw[4 * v + 0] = sv(w.x, sg); xw[4 * v + 1] = sv(w.y, sg); xw[4 * v + 2] = sv(w.z, sg); xw[4 * v + 3] = sv(w.w, sg);
    }
    const uint32_t part0 = pdiv((uint32_t)i0, pmagic);
    uint32_t kk = pl.rice[part0];
    uint32_t bits = 0;
#pragma unroll 1
    for (int ib = i0; ib < i0 + kFuRun && ib < N; ib += 4) {
        const int4 w = *reinterpret_cast<const int4*>(tile + tix(G, ib));
        int32_t xq[4];
        xq[0] = sv(w.x, sg); xq[1] = sv(w.y, sg); xq[2] = sv(w.z, sg); xq[3] = sv(w.w, sg);
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int i = ib + t;
            int32_t r;
            if (WIDE) {
                long long s = 0;
#pragma unroll
                for (int j = 0; j < C; j++) s += (long long)q[j] * (long long)((t - 1 - j >= 0) ? xq[(t - 1 - j) & 3] : xw[(C + t - 1 - j) % C]);
                r = (int32_t)((long long)xq[t] - (s >> shift));
            } else {
                int s = 0;
#pragma unroll
                for (int j = 0; j < C; j++) s += q[j] * ((t - 1 - j >= 0) ? xq[(t - 1 - j) & 3] : xw[(C + t - 1 - j) % C]);
                r = xq[t] - (s >> shift);
            }
            const uint32_t uu = ((uint32_t)r << 1) ^ (uint32_t)(r >> 31);
            bool head;
            if (runpart) head = (i == i0 && (uint32_t)i0 == part0 * psize && i0 >= order) || (i == order);
            else {
                const uint32_t part = pdiv((uint32_t)i, pmagic);
                kk = pl.rice[part];
                head = (i == order) || ((uint32_t)i == part * psize && i > order);
            }
            if (i >= order && i < N) {
                if (EMIT) {
                    if (head) { put_code(obuf, pos, kk, plen); pos += plen; }
                    pos += uu >> kk;                                                 // unary zeros: the image is already zero
                    put_code(obuf, pos, (1u << kk) | (uu & ((1u << kk) - 1u)), kk + 1u);
                    pos += kk + 1u;
                } else {
                    bits += (uu >> kk) + 1u + kk + (head ? plen : 0u);
                }
            }
        }
#pragma unroll
        for (int k2 = 0; k2 < C - 4; k2++) xw[k2] = xw[k2 + 4];
#pragma unroll
        for (int k2 = 0; k2 < 4; k2++) xw[C - 4 + k2] = xq[k2];
    }
    return bits;
}

template <int C, bool WIDE>
__device__ __noinline__ uint32_t fu_pack_body(const SubframePlan& pl, const int32_t* __restrict__ qs, int order, int shift,
                                              const int32_t* __restrict__ tile, const FuGeo G, const Sig sg, uint32_t body,
                                              uint32_t* __restrict__ obuf, FuShared& S, int warp, int lane) {
    const int N = G.N;
    int32_t q[C];
#pragma unroll
    for (int j = 0; j < C; j++) q[j] = (j < order) ? qs[j] : 0;
    const uint32_t plen = pl.rice2 ? 5u : 4u;
    const uint32_t psize = (uint32_t)N >> pl.part_order;
    const uint32_t pmagic = psize > 1u ? (uint32_t)((0x100000000ull + psize - 1u) / psize) : 0u;   // i / psize == umulhi(i, pmagic) for i, psize < 2^16 (psize 1: pdiv)
    const bool runpart = (psize % (uint32_t)kFuRun) == 0u;                        // a run never straddles a partition boundary
    uint32_t done_bits = 0;
    for (int r0 = 0; r0 < N; r0 += kFuWarps * kFuChunk) {
        const int i0 = r0 + warp * kFuChunk + lane * kFuRun;
        uint32_t mybits = 0;
        if (i0 < N) mybits = fu_pack_run<C, WIDE, false>(pl, q, order, shift, tile, G, sg, i0, plen, psize, pmagic, runpart, 0u, obuf);
        // exclusive scan of the lanes' bit counts; warp totals through shared memory
        uint32_t incl = mybits;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t2 = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t2; }
        if (lane == 31) S.segtot[warp] = incl;
        __syncthreads();
        uint32_t pos = body + done_bits + (incl - mybits), round_bits = 0;
#pragma unroll
        for (int w2 = 0; w2 < kFuWarps; w2++) { const uint32_t t2 = S.segtot[w2]; if (w2 < warp) pos += t2; round_bits += t2; }
        if (mybits) fu_pack_run<C, WIDE, true>(pl, q, order, shift, tile, G, sg, i0, plen, psize, pmagic, runpart, pos, obuf);
        done_bits += round_bits;
        __syncthreads();                              // segtot is reused by the next round / subframe
    }
    return done_bits;
}

template <bool WIDE>
__device__ __forceinline__ uint32_t fu_pack_dispatch(int order, const SubframePlan& pl, const int32_t* q, int shift, const int32_t* tile,
                                                     const FuGeo& G, const Sig& sg, uint32_t body, uint32_t* obuf, FuShared& S, int warp, int lane) {
    if (order <= 4) return fu_pack_body<4, WIDE>(pl, q, order, shift, tile, G, sg, body, obuf, S, warp, lane);
    if (order <= 8) return fu_pack_body<8, WIDE>(pl, q, order, shift, tile, G, sg, body, obuf, S, warp, lane);
    return fu_pack_body<12, WIDE>(pl, q, order, shift, tile, G, sg, body, obuf, S, warp, lane);
}

// ------------------------------------------------------------------------------------------------ the kernel
// The autocorrelation warp of a CTA keeps one SM sub-partition's FP64 pipe busy (a warp-wide DFMA occupies it for two
// cycles whatever the number of active lanes).  If the four resident CTAs all gave that job to the same warp index, their
// chains would share ONE sub-partition's pipe (measured: 27 cycles per step instead of the 8.1-cycle DFMA latency).  A
// ticket per SM rotates the warp index, so co-resident CTAs land on different sub-partitions.
__device__ unsigned int g_fu_ticket[256];

#ifndef FB_FU_MIN_CTAS
#define FB_FU_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(kFuThreads, FB_FU_MIN_CTAS)
fused_encode_kernel(const int16_t* __restrict__ pcm, const FrameDesc* __restrict__ frames, const float* __restrict__ windows,
                    EncParams P, FuLayout L, uint8_t* __restrict__ frame_ca, EncStats* __restrict__ stats,
                    uint8_t* __restrict__ scratch, uint32_t scratch_stride, uint32_t* __restrict__ frame_len,
                    unsigned long long* __restrict__ tl) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // optional per-CTA timeline (FLACB200_FU_TIMELINE=<file>; tools/fu_timeline.py): clock64 at the phase boundaries
#define FU_MARK(k) do { if (tl && tid == 0) tl[(size_t)blockIdx.x * 16 + (k)] = (unsigned long long)clock64(); } while (0)
    FU_MARK(0);
    const int nsig = (int)P.n_signals;                    // 2 or 4
    const FrameDesc fd = frames[blockIdx.x];
    const int N = (int)fd.blocksize;

    FuGeo G;
    G.N = N;
    G.B0w = (((N + 31) >> 5) + 3) & ~3;
    if (G.B0w < 4) G.B0w = 4;
    G.gap = ((G.B0w >> 2) & 1) ? 0 : 4;
    G.RS = G.B0w + G.gap;
    G.magic = (uint32_t)((0x100000000ull + (uint32_t)G.B0w - 1ull) / (uint32_t)G.B0w);

    int32_t* tile = reinterpret_cast<int32_t*>(smem_raw);
    unsigned char* ovl = smem_raw + L.ovl_off;
    double* ring_all = reinterpret_cast<double*>(ovl + L.ring_off);
    float* wring_all = reinterpret_cast<float*>(ovl + L.wring_off);
    double* acstore = reinterpret_cast<double*>(ovl + L.acstore_off);
    WarpScratch* wsall = reinterpret_cast<WarpScratch*>(ovl + L.ws_off);
    unsigned long long* psum_all = reinterpret_cast<unsigned long long*>(ovl + L.psum_off);
    unsigned long long* fixsum_all = reinterpret_cast<unsigned long long*>(ovl + L.fixsum_off);
    SubframePlan* base_plan = reinterpret_cast<SubframePlan*>(ovl + L.baseplan_off);
    SubframePlan* step_plan = reinterpret_cast<SubframePlan*>(ovl + L.stepplan_off);
    uint32_t* obuf = reinterpret_cast<uint32_t*>(ovl);
    FuShared& S = *reinterpret_cast<FuShared*>(smem_raw + L.shared_off);
    uint16_t (*crc_tabs)[256] = reinterpret_cast<uint16_t (*)[256]>(smem_raw + L.crctab_off);
    unsigned long long* psum = psum_all + (size_t)warp * 2 * kMaxParts;
    WarpScratch& ws = wsall[warp];
    const int n_steps = (int)L.n_steps, nwin = (int)L.n_win;

    // =================== stage: the frame crosses HBM once ===================
    const int16_t* base = pcm + fd.pcm_off;
    const bool use_tma = ((reinterpret_cast<uintptr_t>(base) & 15u) == 0u);
    if (tid == 0) {
        mbar_init(&S.mbar, 1); S.queue_a = 0; S.queue_b = 0; S.nneed = 0; S.ca = 0;
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        S.ac_warp = (int)(atomicAdd(&g_fu_ticket[smid & 255u], 1u) & (unsigned)(kFuWarps - 1));
        if (tl) tl[(size_t)blockIdx.x * 16 + 11] = smid;
    }
    if (tid < kFuSig) { S.sig_or[tid] = 0u; S.sig_and[tid] = 0xffffffffu; S.best_bits[tid] = 0u; }
    for (int i = tid; i < kFuSig * kMaxSteps; i += kFuThreads) (&S.step_bits[0][0])[i] = 0xffffffffu;
    reinterpret_cast<uint2*>(&crc_tabs[0][0])[tid] = reinterpret_cast<const uint2*>(&g_crc16_slice.t[0][0])[tid];   // 256 threads x 8 bytes = the four tables
    __syncthreads();
    if (use_tma) {
        if (warp == 0) {
            // one bulk copy per tile row (row r = samples [r*B0w, (r+1)*B0w)), whole 16-byte units; completion in bytes on the mbarrier
            const int n_r = max(0, min(G.B0w, N - lane * G.B0w));
            const uint32_t bytes = (uint32_t)(n_r * 4) & ~15u;
            const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
            if (lane == 0) mbar_expect_tx(&S.mbar, total);
            __syncwarp();
            int32_t* dst = tile + lane * G.RS;
            const int32_t* src = reinterpret_cast<const int32_t*>(base) + lane * G.B0w;
            if (bytes) tma_load_1d(dst, src, bytes, &S.mbar);
            // the last words of a row whose length is not a multiple of four; zero up to the next quad
            for (int w = (int)(bytes >> 2); w < ((n_r + 3) & ~3); w++) dst[w] = (w < n_r) ? __ldg(src + w) : 0;
            mbar_wait(&S.mbar, 0);          // the other warps wait at the CTA barrier below without spending issue slots
        }
    } else {
        const bool al4 = ((reinterpret_cast<uintptr_t>(base) & 3u) == 0u);
        const int Nq = (N + 3) & ~3;
        for (int i = tid; i < Nq; i += kFuThreads) {
            int wd = 0;
            if (i < N) {
                if (al4) wd = __ldg(reinterpret_cast<const int*>(base) + i);
                else wd = (int)((uint32_t)(uint16_t)__ldg(base + 2 * i) | ((uint32_t)(uint16_t)__ldg(base + 2 * i + 1) << 16));
            }
            tile[tix(G, i)] = wd;
        }
    }
    __syncthreads();
    FU_MARK(1);

    // =================== OR / AND of every signal (up: get_wasted_bits_) ===================
    {
        uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0, a0 = ~0u, a1 = ~0u, a2 = ~0u, a3 = ~0u;
        const int row = tid >> 3, sub = tid & 7;
        const int row_lo = row * G.B0w;
        for (int c = 4 * sub; c < G.B0w && row_lo + c < N; c += 32) {
            const int4 w = *reinterpret_cast<const int4*>(tile + row * G.RS + c);
            const int xw[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int t = 0; t < 4; t++) {
                if (row_lo + c + t < N) {
                    const int lo = (int)(short)xw[t], hi = xw[t] >> 16, m = (lo + hi) >> 1, sd = lo - hi;
                    o0 |= (uint32_t)lo; o1 |= (uint32_t)hi; o2 |= (uint32_t)m; o3 |= (uint32_t)sd;
                    a0 &= (uint32_t)lo; a1 &= (uint32_t)hi; a2 &= (uint32_t)m; a3 &= (uint32_t)sd;
                }
            }
        }
        o0 = __reduce_or_sync(0xffffffffu, o0); o1 = __reduce_or_sync(0xffffffffu, o1);
        o2 = __reduce_or_sync(0xffffffffu, o2); o3 = __reduce_or_sync(0xffffffffu, o3);
        a0 = __reduce_and_sync(0xffffffffu, a0); a1 = __reduce_and_sync(0xffffffffu, a1);
        a2 = __reduce_and_sync(0xffffffffu, a2); a3 = __reduce_and_sync(0xffffffffu, a3);
        if (lane == 0) {
            atomicOr(&S.sig_or[0], o0); atomicOr(&S.sig_or[1], o1); atomicAnd(&S.sig_and[0], a0); atomicAnd(&S.sig_and[1], a1);
            if (nsig > 2) { atomicOr(&S.sig_or[2], o2); atomicOr(&S.sig_or[3], o3); atomicAnd(&S.sig_and[2], a2); atomicAnd(&S.sig_and[3], a3); }
        }
    }
    __syncthreads();
    FU_MARK(2);

    const int bps = (int)P.bps, ch = 2;
    auto sig_wasted = [&](int s) { return wasted_from_or(S.sig_or[s], bps); };
    auto sig_sbps = [&](int s) { return bps - sig_wasted(s) + ((P.do_mid_side && s == ch + 1) ? 1 : 0); };
    auto sig_const = [&](int s) { return N > 4 && S.sig_or[s] == S.sig_and[s]; };
    // up: process_subframes_ limit_min_bitrate: when every earlier channel is constant, the last channel (and mid/side after
    // it) may not use a constant subframe
    auto sig_disable_const = [&](int s) {
        if (!(P.limit_min_bitrate && s >= ch - 1)) return false;
        for (int c2 = 0; c2 < ch - 1; c2++) if (!sig_const(c2)) return false;
        return true;
    };
    const int omax_frame = min((int)P.max_part_order, N ? (__ffs(N) - 1) : 0);
    const int max_lpc = (N > 4 && P.max_lpc_order > 0) ? (((int)P.max_lpc_order >= N) ? N - 1 : (int)P.max_lpc_order) : 0;
    if (tid == 0) {
        int n = 0;
        for (int s = 0; s < nsig; s++)
            if (N > 4 && max_lpc > 0 && !(sig_const(s) && !sig_disable_const(s))) S.need_list[n++] = s;
        S.nneed = n;
    }
    __syncthreads();
    const int nneed = S.nneed;
    FU_MARK(3);

    // =================== queue A: autocorrelation items first (long, latency bound), then the fixed analyses ===================
    const int n_ac = (max_lpc > 0 && nneed > 0) ? nwin : 0;
    {
        // autocorrelation item t (window t = depth b, position k) belongs to warp (ac_warp + t) mod 8: at most six windows.
        // Up to three windows (levels 3-7) get a second warp each, (ac_warp + t + 4) mod 8, that converts for the chain warp.
        const int wrel = (warp - S.ac_warp) & (kFuWarps - 1);
        const bool paired = n_ac * 2 + 2 <= kFuWarps;
        const int t = paired ? (wrel & 3) : wrel;
        const bool mine = paired ? (wrel < 4 ? wrel < n_ac : (wrel - 4) < n_ac) : wrel < n_ac;
        if (mine) {
            int b = 1, k = t;
            while (k >= b) { k -= b; b++; }
            if (!(b > 1 && N / b <= 32)) {                                   // libFLAC skips windows this short
                const bool mark = tl && lane == 0 && wrel == 0;
                if (mark) { tl[(size_t)blockIdx.x * 16 + 12] = (unsigned long long)clock64(); tl[(size_t)blockIdx.x * 16 + 14] = (unsigned long long)warp; }
                double* rg = ring_all + (size_t)t * kFuSig * kFuRing;
                float* wr = wring_all + (size_t)t * kWinSlots * 32;
                const float* wt = windows + fd.window_off;
                const int lags = (int)P.max_lpc_order + 1;
                if (!paired) fu_autoc_item<0>(tile, G, wt, b, k, nsig, lags, S.sig_or, bps, rg, wr, acstore, nwin, lane, 0);
                else if (wrel < 4) fu_autoc_item<1>(tile, G, wt, b, k, nsig, lags, S.sig_or, bps, rg, wr, acstore, nwin, lane, 2 + t);
                else fu_autoc_item<2>(tile, G, wt, b, k, nsig, lags, S.sig_or, bps, rg, wr, acstore, nwin, lane, 2 + t);
                if (mark) tl[(size_t)blockIdx.x * 16 + 13] = (unsigned long long)clock64();
            }
        }
    }
    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&S.queue_a, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= nsig) break;
        // ---- fixed analysis of signal s: verbatim baseline, constant, or the guessed fixed order (rows E3, E10, E11) ----
        const int s = t;
        SubframePlan& pl = base_plan[s];
        reinterpret_cast<uint32_t*>(&pl)[lane] = 0u;
        __syncwarp();
        const int wasted = sig_wasted(s), sbps = sig_sbps(s);
        uint32_t best_bits = 8u + (uint32_t)wasted + (uint32_t)N * (uint32_t)sbps;          // up: evaluate_verbatim_subframe_
        if (lane == 0) { pl.type = kVerbatim; pl.wasted = (uint8_t)wasted; pl.sbps = (uint8_t)sbps; pl.bits_est = best_bits; }
        if (N > 4) {
            if (sig_const(s) && !sig_disable_const(s)) {
                const uint32_t bits = 8u + (uint32_t)wasted + (uint32_t)sbps;                 // up: evaluate_constant_subframe_
                if (bits < best_bits) { best_bits = bits; if (lane == 0) { pl.type = kConstant; pl.bits_est = bits; } }
            } else {
                const Sig sg = make_sig(s, wasted);
                unsigned long long* fixsum = fixsum_all + (size_t)s * 5 * kMaxParts;
                const int nparts0 = 1 << omax_frame, psize0 = N >> omax_frame;
                for (int i = lane; i < 5 * kMaxParts; i += 32) fixsum[i] = 0ull;
                __syncwarp();
                unsigned long long e[5];
                if ((psize0 & 3) == 0) fu_fixed_sums<true>(tile, G, sg, psize0, fixsum, lane, e);
                else fu_fixed_sums<false>(tile, G, sg, psize0, fixsum, lane, e);
                __syncwarp();
                if ((uint32_t)sbps + ilog2_u32((uint32_t)N - 4u) + 1u < 32u) {                 // libFLAC's 32-bit accumulators wrap
#pragma unroll
                    for (int k = 0; k < 5; k++) e[k] &= 0xffffffffull;
                }
                int forder;
                const unsigned long long m34 = min(e[3], e[4]), m234 = min(e[2], m34), m1234 = min(e[1], m234);
                if (e[0] <= m1234) forder = 0; else if (e[1] <= m234) forder = 1; else if (e[2] <= m34) forder = 2; else if (e[3] <= e[4]) forder = 3; else forder = 4;
                int fo = forder; if (fo >= N) fo = N - 1;
                int omax = omax_frame;
                while (omax > 0 && (N >> omax) <= fo) omax--;
                const int nparts = 1 << omax, psize = N >> omax, ratio = nparts0 >> omax;
                const bool narrow = (uint32_t)sbps + 4u < 32u - ilog2_u32((uint32_t)psize);
                // partition sums of the chosen order: the pass above covers samples 4..N-1; samples fo..3 are added here, each to
                // the partition it lies in (all in partition 0 unless the partitions are shorter than four samples)
                if (lane == 0) {
                    for (int i = fo; i < 4; i++) {
                        const int x0 = sv(tile[tix(G, i)], sg);
                        const int xa = i >= 1 ? sv(tile[tix(G, i - 1)], sg) : 0, xb = i >= 2 ? sv(tile[tix(G, i - 2)], sg) : 0, xc = i >= 3 ? sv(tile[tix(G, i - 3)], sg) : 0;
                        int d;
                        if (fo == 0) d = x0; else if (fo == 1) d = x0 - xa; else if (fo == 2) d = x0 - 2 * xa + xb; else d = x0 - 3 * xa + 3 * xb - xc;
                        fixsum[fo * kMaxParts + i / psize0] += (unsigned long long)(uint32_t)abs(d);
                    }
                }
                __syncwarp();
                for (int p = lane; p < nparts; p += 32) {
                    unsigned long long v = 0;
                    for (int j = 0; j < ratio; j++) v += fixsum[fo * kMaxParts + p * ratio + j];
                    psum[p] = v;
                }
                __syncwarp();
                int po; uint32_t k0, k1;
                const uint32_t rb = rice_search(psum, N, fo, omax, narrow, P.rice_limit, lane, &po, &k0, &k1);
                const uint32_t est = add_sat(8u + (uint32_t)wasted + (uint32_t)fo * (uint32_t)sbps, rb);
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
    __syncthreads();
    FU_MARK(4);

    // =================== queue B: one LPC candidate per (signal, apodization step) ===================
    // up: apply_apodization_ + evaluate_lpc_subframe_.  Step list of set_next_subdivide_tukey:
    // full window, then for depth b = 2..parts: partial windows c = 0,2,.. interleaved with their punch-outs.
    const int n_tasks_b = nneed * n_steps;
    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&S.queue_b, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_tasks_b) break;
        const int s = S.need_list[t / n_steps], step = t - (t / n_steps) * n_steps;
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
            // guard band (DESIGN.md "log guard"): counted, never silently ignored
            const int ul_best = __shfl_sync(0xffffffffu, (int)ul, guess & 31);
            const bool amb = (lane >= 1 && lane <= max_this && lane != guess) && (ul || ul_best) && fabs(bits - bb) <= 1e-12 * fabs(bb);
            if (__any_sync(0xffffffffu, amb) && lane == 0 && stats) atomicAdd(&stats->log_ambiguous, 1ull);
        }
        const int order = guess;
        bool ul2;
        const double rbps = expected_bits_per_sample(ws.lperr[order - 1], FB_DDIV(0.5, (double)(N - order)), &ul2);
        if (ul2 && fabs(rbps - (double)sbps) <= 1e-12 * (double)sbps && lane == 0 && stats) atomicAdd(&stats->log_ambiguous, 1ull);
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
                const Sig sg = make_sig(s, wasted);
                const bool rejected = fu_lpc_psums(tile, G, sg, ws.q, order, shift, psize, nparts, limit || pred_bps > 32, limit, psum, lane);
                if (!rejected) {
                    int po; uint32_t k0, k1;
                    const uint32_t rb = rice_search(psum, N, order, omax, narrow, P.rice_limit, lane, &po, &k0, &k1);
                    const uint32_t est = add_sat(8u + (uint32_t)wasted + 4u + 5u + (uint32_t)order * (uint32_t)(prec + sbps), rb);
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
    FU_MARK(5);

    // =================== selection: candidates in libFLAC's order, replace only on strict <; channel assignment ===================
    if (warp == 0) {
        uint32_t bestv = 0xffffffffu; int pick = -1;
        if (lane < nsig) {
            bestv = S.best_bits[lane];
            for (int k = 0; k < n_steps; k++) { const uint32_t e = S.step_bits[lane][k]; if (e < bestv) { bestv = e; pick = k; } }
        }
        const uint32_t bL = __shfl_sync(0xffffffffu, bestv, 0), bR = __shfl_sync(0xffffffffu, bestv, 1);
        const uint32_t bM = __shfl_sync(0xffffffffu, bestv, 2), bS = __shfl_sync(0xffffffffu, bestv, 3);
        int ca = 0;                                                       // up: process_subframes_, first minimum of {L+R, L+S, R+S, M+S}
        if (P.do_mid_side) {
            uint32_t minb = bL + bR;
            if (bL + bS < minb) { minb = bL + bS; ca = 1; }
            if (bR + bS < minb) { minb = bR + bS; ca = 2; }
            if (bM + bS < minb) { minb = bM + bS; ca = 3; }
        }
        const int si0 = (ca == 2) ? 3 : (ca == 3 ? 2 : 0), si1 = (ca == 0 || ca == 2) ? 1 : 3;
        const int pick0 = __shfl_sync(0xffffffffu, pick, si0), pick1 = __shfl_sync(0xffffffffu, pick, si1);
        const uint32_t* src0 = reinterpret_cast<const uint32_t*>(pick0 < 0 ? &base_plan[si0] : &step_plan[(size_t)si0 * n_steps + pick0]);
        const uint32_t* src1 = reinterpret_cast<const uint32_t*>(pick1 < 0 ? &base_plan[si1] : &step_plan[(size_t)si1 * n_steps + pick1]);
        reinterpret_cast<uint32_t*>(&S.plan[0])[lane] = src0[lane];       // 128 bytes = 32 words
        reinterpret_cast<uint32_t*>(&S.plan[1])[lane] = src1[lane];
        if (lane == 0) {
            S.sigidx[0] = si0; S.sigidx[1] = si1; S.ca = ca;
            frame_ca[blockIdx.x] = (uint8_t)ca;
            S.hdr_len = (uint32_t)build_frame_header(S.hdr, P.channels, P.bps, P.sample_rate, (uint32_t)N, fd.frame_number, ca);
        }
    }
    __syncthreads();

    // =================== pack (row E12): the analysis scratch becomes the frame image ===================
    FU_MARK(6);
    for (uint32_t i = tid; i < L.obuf_words; i += kFuThreads) obuf[i] = 0u;
    __syncthreads();
    FU_MARK(7);
    if (tid < (int)S.hdr_len) put_bits(obuf, (uint32_t)tid * 8u, S.hdr[tid], 8);
    uint32_t pos = S.hdr_len * 8u;
    for (int c = 0; c < ch; c++) {
        const SubframePlan& pl = S.plan[c];
        const Sig sg = make_sig(S.sigidx[c], pl.wasted);
        const uint32_t sbps = pl.sbps, order = pl.order, wf = pl.wasted ? 1u : 0u;
        auto X = [&](int i) -> uint32_t { return (uint32_t)sv(tile[tix(G, i)], sg); };
        const uint32_t after_hdr = pos + 8u + pl.wasted;
        if (warp == 0) {   // subframe header, warm-up, predictor description: one lane per field
            if (lane == 0) {
                uint32_t tb;
                switch (pl.type) {
                    case kConstant: tb = 0x00u; break;
                    case kVerbatim: tb = 0x02u; break;
                    case kFixed: tb = 0x10u | (order << 1); break;
                    default: tb = 0x40u | ((order - 1u) << 1); break;
                }
                put_bits(obuf, pos, tb | wf, 8);
                if (pl.wasted) put_bits(obuf, pos + 8u, 1u, pl.wasted);        // unary: wasted-1 zeros, then 1
                if (pl.type == kConstant) put_bits(obuf, after_hdr, X(0), sbps);
            }
            if (pl.type == kFixed || pl.type == kLpc) {
                if ((uint32_t)lane < order) put_bits(obuf, after_hdr + (uint32_t)lane * sbps, X(lane), sbps);
                uint32_t p2 = after_hdr + order * sbps;
                if (pl.type == kLpc) {
                    if (lane == 12) put_bits(obuf, p2, (((uint32_t)pl.precision - 1u) << 5) | ((uint32_t)pl.shift & 31u), 9);
                    if (lane >= 16 && (uint32_t)(lane - 16) < order) put_bits(obuf, p2 + 9u + (uint32_t)(lane - 16) * pl.precision, (uint32_t)pl.qlp[lane - 16], pl.precision);
                    p2 += 9u + order * pl.precision;
                }
                if (lane == 31) put_bits(obuf, p2, ((pl.rice2 ? 1u : 0u) << 4) | pl.part_order, 6);
            }
        }
        if (pl.type == kConstant) pos = after_hdr + sbps;
        else if (pl.type == kVerbatim) {
            for (int i = tid; i < N; i += kFuThreads) put_bits(obuf, after_hdr + (uint32_t)i * sbps, X(i), sbps);
            pos = after_hdr + (uint32_t)N * sbps;
        } else {
            uint32_t body = after_hdr + order * sbps + 6u, blen;
            if (pl.type == kLpc) {
                body += 9u + order * pl.precision;
                int32_t asum = 0;
                for (uint32_t j = 0; j < order; j++) asum += abs(pl.qlp[j]);
                if (asum == 0) asum = 1;
                // same accumulator-width rule as the analysis (up: FLAC__lpc_max_prediction_before_shift_bps)
                if ((int)sbps + (int)silog2((int64_t)asum) <= 32) blen = fu_pack_dispatch<false>((int)order, pl, pl.qlp, pl.shift, tile, G, sg, body, obuf, S, warp, lane);
                else blen = fu_pack_dispatch<true>((int)order, pl, pl.qlp, pl.shift, tile, G, sg, body, obuf, S, warp, lane);
            } else {
                const int32_t cfix[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};
                blen = fu_pack_body<4, false>(pl, cfix[order], (int)order, 0, tile, G, sg, body, obuf, S, warp, lane);
            }
            pos = body + blen;
        }
    }
    __syncthreads();

    // =================== CRC-16 over the byte-padded frame, append, store ===================
    FU_MARK(8);
    const uint32_t nb = (pos + 7u) >> 3;
    {
        const uint16_t c2 = cta_crc16_words<kFuThreads>([&](uint32_t j) {
            const uint32_t w0 = __byte_perm(obuf[j >> 2], 0u, 0x0123), w1 = __byte_perm(obuf[(j >> 2) + 1u], 0u, 0x0123);
            return __funnelshift_r(w0, w1, (j & 3u) * 8u);
        }, nb, crc_tabs, S.crc_warp, tid);
        if (tid == 0) put_bits(obuf, nb * 8u, c2, 16);
    }
    __syncthreads();
    FU_MARK(9);
    {
        const uint32_t total = nb + 2u;
        uint32_t* dst = reinterpret_cast<uint32_t*>(scratch + (size_t)blockIdx.x * scratch_stride);
        for (uint32_t wd = tid; wd < (total + 3u) / 4u; wd += kFuThreads) dst[wd] = __byte_perm(obuf[wd], 0u, 0x0123);
        if (tid == 0) frame_len[blockIdx.x] = total;
    }
    FU_MARK(10);
#undef FU_MARK
}

// ------------------------------------------------------------------------------------------------ host side
static uint32_t fu_apod_steps(const EncParams& P) {
    uint32_t n = 1;
    for (uint32_t b = 2; b <= P.apod_parts; b++) n += (b == 2) ? 2u : 2u * b;
    return P.max_lpc_order ? n : 0u;
}

static FuLayout fused_layout(const EncParams& P, uint32_t scratch_stride) {
    FuLayout L{};
    const uint32_t b0w = std::max(4u, (((P.blocksize + 31u) / 32u) + 3u) & ~3u), rs = b0w + (((b0w >> 2) & 1u) ? 0u : 4u);
    L.tile_words = 32u * rs;
    L.n_win = P.apod_parts * (P.apod_parts + 1u) / 2u;
    L.n_steps = fu_apod_steps(P);
    L.obuf_words = scratch_stride / 4u + 4u;
    auto up = [](uint32_t v, uint32_t a) { return (v + a - 1u) / a * a; };
    L.ovl_off = up(L.tile_words * 4u, 128u);
    uint32_t o = 0;
    L.ring_off = o;      o += up((P.max_lpc_order ? L.n_win : 0u) * kFuSig * kFuRing * 8u, 16u);
    L.wring_off = o;     o += (P.max_lpc_order ? L.n_win : 0u) * kWinSlots * 32u * 4u;
    L.acstore_off = o;   o += up(kFuSig * L.n_win * kAcStoreStride * 8u, 16u);
    L.ws_off = o;        o += up(kFuWarps * (uint32_t)sizeof(WarpScratch), 16u);
    L.psum_off = o;      o += kFuWarps * 2u * kMaxParts * 8u;
    L.fixsum_off = o;    o += kFuSig * 5u * kMaxParts * 8u;
    L.baseplan_off = o;  o += kFuSig * (uint32_t)sizeof(SubframePlan);
    L.stepplan_off = o;  o += kFuSig * std::max(1u, L.n_steps) * (uint32_t)sizeof(SubframePlan);
    L.ovl_bytes = up(std::max(o, L.obuf_words * 4u + 16u), 128u);
    L.shared_off = L.ovl_off + L.ovl_bytes;
    L.crctab_off = up(L.shared_off + (uint32_t)sizeof(FuShared), 16u);
    L.total_bytes = L.crctab_off + 4u * 256u * 2u;
    return L;
}

// Can this batch take the fused kernel?  16-bit stereo in an int16 container, no loose mid/side (its followers wait for
// a decision made in another frame), tile + scratch small enough for at least two CTAs per SM.
bool fused_eligible(const EncParams& P, uint32_t scratch_stride, int max_smem_optin) {
    if (!(P.container_bytes == 2 && P.channels == 2) || P.loose_frames) return false;
    if (P.max_lpc_order > kMaxOrder) return false;
    const FuLayout L = fused_layout(P, scratch_stride);
    return (int)L.total_bytes <= max_smem_optin && L.total_bytes <= 110u * 1024u;
}

void launch_fused(const void* pcm, const FrameDesc* frames, const float* windows, const EncParams& P, int n_frames, uint8_t* frame_ca,
                  EncStats* stats, uint8_t* scratch, uint32_t scratch_stride, uint32_t* frame_len, cudaStream_t stream) {
    const FuLayout L = fused_layout(P, scratch_stride);
    cudaFuncSetAttribute(fused_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total_bytes);
    // debugging aid: FLACB200_FU_TIMELINE=<file> dumps 16 clock64 marks per CTA of every launch (synchronous; tools/fu_timeline.py)
    unsigned long long* tl = nullptr;
    const char* tl_path = getenv("FLACB200_FU_TIMELINE");
    if (tl_path && cudaMalloc(&tl, (size_t)n_frames * 16 * 8) != cudaSuccess) tl = nullptr;
    if (tl) cudaMemsetAsync(tl, 0, (size_t)n_frames * 16 * 8, stream);
    fused_encode_kernel<<<(unsigned)n_frames, kFuThreads, L.total_bytes, stream>>>((const int16_t*)pcm, frames, windows, P, L, frame_ca, stats,
                                                                                  scratch, scratch_stride, frame_len, tl);
    if (tl) {
        std::vector<unsigned long long> h((size_t)n_frames * 16);
        cudaStreamSynchronize(stream);
        cudaMemcpy(h.data(), tl, h.size() * 8, cudaMemcpyDeviceToHost);
        if (FILE* f = fopen(tl_path, "wb")) { fwrite(h.data(), 8, h.size(), f); fclose(f); }
        cudaFree(tl);
    }
}

}  // namespace fb
