// enc_fused.cu -- the per-frame encode path for 16-bit stereo (the headline layout) on TMA-staged frame tiles:
//
//   autoc_kernel (enc_analyze.cu, un-shifted mode: window + f64 autocorrelation chains, no OR/AND pass in front)
//   -> fused_analyze_kernel: stage (TMA) -> OR/AND -> fixed predictors -> LPC candidates -> choose        (rows E2-E11)
//   -> fused_pack_kernel:    stage (TMA) -> Rice bodies -> frame image -> CRC-16 -> store                 (row E12)
//
// The frame -- 16 KiB of interleaved int16 pairs for a 4096-sample block -- is staged once per kernel: warp 0 issues one
// cp.async.bulk (TMA, SASS UBLKCP) per tile row, the copies complete on an mbarrier, and every phase works on the
// shared-memory tile.  Between the kernels travel the autocorrelations (HBM scratch, 512 B per frame), two 128-byte plans
// and the channel assignment per frame.  A first version ran everything in ONE kernel per frame; phases of very different
// parallelism (a few sequential f64 chains next to thousands of independent samples) idled at its barriers, and splitting at
// the two natural boundaries was faster although the tile is staged twice.  The multi-kernel path of enc_analyze.cu /
// enc_pack.cu stages with plain loads and remains the path for the other layouts, loose mid/side and debug traces.
//
// Tile layout: 32 rows of B0w = roundup4(ceil(N/32)) packed words (L | R << 16), row stride RS = B0w (+4 so that RS/4
// is odd).  Lane p of an analysis warp streams row p with 16-byte shared loads: eight consecutive lanes hit eight
// different 16-byte bank groups, so the quarter-warp wavefronts of LDS.128 are conflict free, and the row starts stay
// 16-byte aligned, which is what lets TMA write them.
//
// Instruction diet against the multi-kernel path (which was issue bound):
//  * fixed predictors: ONE pass yields the five error sums AND the per-partition sums of all five orders; the
//    residual pass of the chosen order is gone (|k-th difference| is the order-k residual);
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
constexpr int kFuAnThreads = 128;          // analysis kernel: four warps, one work item each
constexpr int kFuAnWarps = kFuAnThreads / 32;
constexpr int kFuSig = 4;                  // L, R, mid, side
constexpr int kFuRun = 20;                 // consecutive samples a lane codes in the pack phase (five 16-byte quads: lanes 80 bytes apart, conflict free)
constexpr int kFuChunk = 32 * kFuRun;      // samples a warp codes per round

// Shared-memory plan, computed on the host (fused_layout) and passed by value.  The `ovl` region is used twice:
// analysis scratch first, the frame image afterwards.
struct FuLayout {
    uint32_t tile_words;
    uint32_t ovl_off, ovl_bytes;
    uint32_t acstore_off, ws_off, psum_off, fixsum_off, baseplan_off, stepplan_off;   // inside ovl
    uint32_t pk_obuf_off, pk_shared_off, pk_crctab_off, pk_total_bytes;   // pack kernel: tile | image | FuShared | CRC tables
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
    int      frame, pad2;                  // the frame this CTA encodes (its ticket)
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

// ------------------------------------------------------------------------------------------------ fixed predictors
// up: fixed.c FLAC__fixed_compute_best_predictor[_wide] and, in the same pass, what
// precompute_partition_info_sums_ would compute from the order-k fixed residual: |k-th difference| IS that residual.
// Lane p streams row p; per (row x partition) segment the five 32-bit partial sums go to fixsum[k][partition]
// (samples 4.. only, exactly the range of libFLAC's error sums; the caller adds the samples k..3 of the chosen order
// to partition 0).  Returns the five totals in e[] (every lane).
template <bool VEC>
__device__ __noinline__ void fu_fixed_sums(const int32_t* __restrict__ tile, const FuGeo G, const Sig sg, int psize,
                                           unsigned long long* __restrict__ fixsum, int fstride, int lane, unsigned long long* e) {
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
            atomicAdd(&fixsum[0 * fstride + part], (unsigned long long)s0);
            atomicAdd(&fixsum[1 * fstride + part], (unsigned long long)s1);
            atomicAdd(&fixsum[2 * fstride + part], (unsigned long long)s2);
            atomicAdd(&fixsum[3 * fstride + part], (unsigned long long)s3);
            atomicAdd(&fixsum[4 * fstride + part], (unsigned long long)s4);
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
// Round r: warp w codes samples [(r * 8 + w) * 640, +640), lane l the 20 consecutive samples at + 20 l, four at a time
// (one 16-byte shared load; the predictor history slides through registers).  Pass 1 counts the lane's bits (codes + the
// parameter field of every partition that starts inside its run); an exclusive warp scan and the eight warp totals (one
// CTA barrier per round) give the lane's absolute bit position; pass 2 computes the residuals again and ORs the codes
// into the zeroed image: neighbouring lanes are ~200 bits apart, so the shared atomics rarely meet in one word.  The
// residuals are computed twice instead of parked in registers: the loops stay rolled and small (this phase was
// instruction-fetch bound when it was unrolled over the run).  Returns the body length in bits.
// C = order class, WIDE = 64-bit accumulate (chosen exactly as the analysis does).
template <int C, bool WIDE, bool EMIT>
__device__ __forceinline__ uint32_t fu_pack_run(const SubframePlan& pl, const int32_t (&q)[C], int order, int shift, const int32_t* __restrict__ tile,
                                                const FuGeo& G, const Sig& sg, int i0, uint32_t plen, uint32_t psize, uint32_t pmagic, bool quadpart,
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
    uint32_t kk = 0;
    uint32_t bits = 0;
#pragma unroll 1
    for (int ib = i0; ib < i0 + kFuRun && ib < N; ib += 4) {
        const int4 w = *reinterpret_cast<const int4*>(tile + tix(G, ib));
        int32_t xq[4];
        xq[0] = sv(w.x, sg); xq[1] = sv(w.y, sg); xq[2] = sv(w.z, sg); xq[3] = sv(w.w, sg);
        bool qhead = false;                            // quadpart: a partition can only start at the first sample of a quad
        if (quadpart) {
            const uint32_t part = pdiv((uint32_t)ib, pmagic);
            kk = pl.rice[part];
            qhead = ((uint32_t)ib == part * psize) && ib > order;
        }
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
            if (quadpart) head = (t == 0 && qhead) || (i == order);
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
    const bool quadpart = (psize & 3u) == 0u;                                      // a 4-sample quad never straddles a partition boundary
    uint32_t done_bits = 0;
    for (int r0 = 0; r0 < N; r0 += kFuWarps * kFuChunk) {
        const int i0 = r0 + warp * kFuChunk + lane * kFuRun;
        uint32_t mybits = 0;
        if (i0 < N) mybits = fu_pack_run<C, WIDE, false>(pl, q, order, shift, tile, G, sg, i0, plen, psize, pmagic, quadpart, 0u, obuf);
        // exclusive scan of the lanes' bit counts; warp totals through shared memory
        uint32_t incl = mybits;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t2 = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t2; }
        if (lane == 31) S.segtot[warp] = incl;
        __syncthreads();
        uint32_t pos = body + done_bits + (incl - mybits), round_bits = 0;
#pragma unroll
        for (int w2 = 0; w2 < kFuWarps; w2++) { const uint32_t t2 = S.segtot[w2]; if (w2 < warp) pos += t2; round_bits += t2; }
        if (mybits) fu_pack_run<C, WIDE, true>(pl, q, order, shift, tile, G, sg, i0, plen, psize, pmagic, quadpart, pos, obuf);
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

// ------------------------------------------------------------------------------------------------ staging (both kernels)
// The frame crosses the SM boundary once per kernel: warp 0 issues one cp.async.bulk (TMA, SASS UBLKCP) per tile row, whole
// 16-byte units, completion counted in bytes on an mbarrier; frames that do not start on a 16-byte boundary are staged
// with plain loads.  Ends with a CTA barrier: the tile is complete for every thread.
template <int NT>
__device__ __forceinline__ void fu_stage(const int16_t* __restrict__ base, int N, const FuGeo& G, int32_t* __restrict__ tile,
                                         unsigned long long* mbar, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    const bool use_tma = ((reinterpret_cast<uintptr_t>(base) & 15u) == 0u);
    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (use_tma) {
        if (warp == 0) {
            const int n_r = max(0, min(G.B0w, N - lane * G.B0w));            // row r = samples [r*B0w, (r+1)*B0w)
            const uint32_t bytes = (uint32_t)(n_r * 4) & ~15u;
            const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
            if (lane == 0) mbar_expect_tx(mbar, total);
            __syncwarp();
            int32_t* dst = tile + lane * G.RS;
            const int32_t* src = reinterpret_cast<const int32_t*>(base) + lane * G.B0w;
            if (bytes) tma_load_1d(dst, src, bytes, mbar);
            // the last words of a row whose length is not a multiple of four; zero up to the next quad
            for (int w = (int)(bytes >> 2); w < ((n_r + 3) & ~3); w++) dst[w] = (w < n_r) ? __ldg(src + w) : 0;
            mbar_wait(mbar, 0);             // the other warps wait at the barrier below without spending issue slots
        }
    } else {
        const bool al4 = ((reinterpret_cast<uintptr_t>(base) & 3u) == 0u);
        const int Nq = (N + 3) & ~3;
        for (int i = tid; i < Nq; i += NT) {
            int wd = 0;
            if (i < N) {
                if (al4) wd = __ldg(reinterpret_cast<const int*>(base) + i);
                else wd = (int)((uint32_t)(uint16_t)__ldg(base + 2 * i) | ((uint32_t)(uint16_t)__ldg(base + 2 * i + 1) << 16));
            }
            tile[tix(G, i)] = wd;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ FuGeo fu_geo(int N) {
    FuGeo G;
    G.N = N;
    G.B0w = (((N + 31) >> 5) + 3) & ~3;
    if (G.B0w < 4) G.B0w = 4;
    G.gap = ((G.B0w >> 2) & 1) ? 0 : 4;
    G.RS = G.B0w + G.gap;
    G.magic = (uint32_t)((0x100000000ull + (uint32_t)G.B0w - 1ull) / (uint32_t)G.B0w);
    return G;
}

// ------------------------------------------------------------------------------------------------ analysis kernel
// One CTA of four warps per frame (rows E2-E11): stage, OR/AND, then one warp per work item -- the fixed analysis of a
// signal, then one LPC candidate per (signal, apodization step) -- and the choice.  Everything the pack kernel needs
// leaves as two 128-byte plans + the channel assignment.  Seven CTAs per SM: the kernel is issue / latency bound and
// lives on the number of warps that have work at the same time, so every warp of a CTA always has an item.
#ifndef FB_FU_AN_CTAS
#define FB_FU_AN_CTAS 7
#endif
__global__ void __launch_bounds__(kFuAnThreads, FB_FU_AN_CTAS)
fused_analyze_kernel(const int16_t* __restrict__ pcm, const FrameDesc* __restrict__ frames, EncParams P, FuLayout L,
                     const double* __restrict__ g_ac, size_t ac_frame_stride, SubframePlan* __restrict__ plans, uint8_t* __restrict__ frame_ca,
                     EncStats* __restrict__ stats) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.x;
    FuShared& S = *reinterpret_cast<FuShared*>(smem_raw + L.shared_off);
    const int nsig = (int)P.n_signals;                    // 2 or 4
    const int n_steps = (int)L.n_steps, nwin = (int)L.n_win;
    const FrameDesc fd = frames[f];
    const int N = (int)fd.blocksize;
    const FuGeo G = fu_geo(N);
    int32_t* tile = reinterpret_cast<int32_t*>(smem_raw);
    unsigned char* ovl = smem_raw + L.ovl_off;
    double* acstore = reinterpret_cast<double*>(ovl + L.acstore_off);
    WarpScratch* wsall = reinterpret_cast<WarpScratch*>(ovl + L.ws_off);
    unsigned long long* psum_all = reinterpret_cast<unsigned long long*>(ovl + L.psum_off);
    unsigned long long* fixsum_all = reinterpret_cast<unsigned long long*>(ovl + L.fixsum_off);
    SubframePlan* base_plan = reinterpret_cast<SubframePlan*>(ovl + L.baseplan_off);
    SubframePlan* step_plan = reinterpret_cast<SubframePlan*>(ovl + L.stepplan_off);
    unsigned long long* psum = psum_all + (size_t)warp * 2 * kMaxParts;
    WarpScratch& ws = wsall[warp];

    if (tid == 0) { S.queue_a = 0; S.queue_b = 0; S.nneed = 0; S.ca = 0; }
    if (tid < kFuSig) { S.sig_or[tid] = 0u; S.sig_and[tid] = 0xffffffffu; S.best_bits[tid] = 0u; }
    for (int i = tid; i < kFuSig * kMaxSteps; i += kFuAnThreads) (&S.step_bits[0][0])[i] = 0xffffffffu;
    fu_stage<kFuAnThreads>(pcm + fd.pcm_off, N, G, tile, &S.mbar, tid);

    // =================== OR / AND of every signal (up: get_wasted_bits_) ===================
    {
        uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0, a0 = ~0u, a1 = ~0u, a2 = ~0u, a3 = ~0u;
        for (int rs = tid; rs < 256; rs += kFuAnThreads) {               // (row, 8 threads per row)
            const int row = rs >> 3, sub = rs & 7;
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

    // =================== queue A: the fixed analyses (the autocorrelation arrives from the chain warp of an earlier CTA) ===================
    const bool want_lpc = (max_lpc > 0 && nneed > 0);
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
                const int fstride = 1 << (int)P.max_part_order;                          // >= the frame's partitions
                unsigned long long* fixsum = fixsum_all + (size_t)s * 5 * fstride;
                const int nparts0 = 1 << omax_frame, psize0 = N >> omax_frame;
                for (int i = lane; i < 5 * fstride; i += 32) fixsum[i] = 0ull;
                __syncwarp();
                unsigned long long e[5];
                if ((psize0 & 3) == 0) fu_fixed_sums<true>(tile, G, sg, psize0, fixsum, fstride, lane, e);
                else fu_fixed_sums<false>(tile, G, sg, psize0, fixsum, fstride, lane, e);
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
                        fixsum[fo * fstride + i / psize0] += (unsigned long long)(uint32_t)abs(d);
                    }
                }
                __syncwarp();
                for (int p = lane; p < nparts; p += 32) {
                    unsigned long long v = 0;
                    for (int j = 0; j < ratio; j++) v += fixsum[fo * fstride + p * ratio + j];
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

    // =================== the frame's autocorrelations (fused_autoc_kernel, earlier on the stream) ===================
    if (want_lpc) {
        // into shared memory, scaled by 4^-wasted: the chain kernel windowed the un-shifted signal (exact power-of-two scaling)
        for (int i = tid; i < nsig * nwin * kAcStoreStride; i += kFuAnThreads) {
            const int s = i / (nwin * kAcStoreStride);
            const int w2 = 2 * sig_wasted(s);
            acstore[i] = __dmul_rn(__ldcg(reinterpret_cast<const double*>(reinterpret_cast<const unsigned char*>(g_ac) + (size_t)f * ac_frame_stride) + i), __longlong_as_double((long long)(1023 - w2) << 52));
        }
        __syncthreads();
    }

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
        // the two coded subframes' plans (128 bytes each) are all the pack kernel needs besides the PCM
        uint32_t* dst = reinterpret_cast<uint32_t*>(plans + (size_t)f * 2);
        dst[lane] = src0[lane];
        dst[32 + lane] = src1[lane];
        if (lane == 0) frame_ca[f] = (uint8_t)ca;
    }

}

// ------------------------------------------------------------------------------------------------ pack kernel
// One CTA of eight warps per frame (row E12): stage the frame again (TMA), code the two chosen subframes into a zeroed
// frame image in shared memory, CRC-16, store.
#ifndef FB_FU_PK_CTAS
#define FB_FU_PK_CTAS 4
#endif
__global__ void __launch_bounds__(kFuThreads, FB_FU_PK_CTAS)
fused_pack_kernel(const int16_t* __restrict__ pcm, const FrameDesc* __restrict__ frames, EncParams P, FuLayout L,
                  const SubframePlan* __restrict__ plans, const uint8_t* __restrict__ frame_ca,
                  uint8_t* __restrict__ scratch, uint32_t scratch_stride, uint32_t* __restrict__ frame_len) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.x;
    const int ch = 2;
    const FrameDesc fd = frames[f];
    const int N = (int)fd.blocksize;
    const FuGeo G = fu_geo(N);
    int32_t* tile = reinterpret_cast<int32_t*>(smem_raw);
    uint32_t* obuf = reinterpret_cast<uint32_t*>(smem_raw + L.pk_obuf_off);
    FuShared& S = *reinterpret_cast<FuShared*>(smem_raw + L.pk_shared_off);
    uint16_t (*crc_tabs)[256] = reinterpret_cast<uint16_t (*)[256]>(smem_raw + L.pk_crctab_off);

    reinterpret_cast<uint2*>(&crc_tabs[0][0])[tid] = reinterpret_cast<const uint2*>(&g_crc16_slice.t[0][0])[tid];   // 256 threads x 8 bytes = the four tables
    if (warp == 0) {
        const int ca = (int)frame_ca[f];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(plans + (size_t)f * 2);
        reinterpret_cast<uint32_t*>(&S.plan[0])[lane] = src[lane];           // 2 x 128 bytes
        reinterpret_cast<uint32_t*>(&S.plan[1])[lane] = src[32 + lane];
        if (lane == 0) {
            S.sigidx[0] = (ca == 2) ? 3 : (ca == 3 ? 2 : 0); S.sigidx[1] = (ca == 0 || ca == 2) ? 1 : 3; S.ca = ca;
            S.hdr_len = (uint32_t)build_frame_header(S.hdr, P.channels, P.bps, P.sample_rate, (uint32_t)N, fd.frame_number, ca);
        }
    }
    fu_stage<kFuThreads>(pcm + fd.pcm_off, N, G, tile, &S.mbar, tid);

    // =================== pack (row E12): the analysis scratch becomes the frame image ===================
    for (uint32_t i = tid; i < L.obuf_words; i += kFuThreads) obuf[i] = 0u;
    __syncthreads();
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
    const uint32_t nb = (pos + 7u) >> 3;
    {
        const uint16_t c2 = cta_crc16_words<kFuThreads>([&](uint32_t j) {
            const uint32_t w0 = __byte_perm(obuf[j >> 2], 0u, 0x0123), w1 = __byte_perm(obuf[(j >> 2) + 1u], 0u, 0x0123);
            return __funnelshift_r(w0, w1, (j & 3u) * 8u);
        }, nb, crc_tabs, S.crc_warp, tid);
        if (tid == 0) put_bits(obuf, nb * 8u, c2, 16);
    }
    __syncthreads();
    {
        const uint32_t total = nb + 2u;
        uint32_t* dst = reinterpret_cast<uint32_t*>(scratch + (size_t)f * scratch_stride);
        for (uint32_t wd = tid; wd < (total + 3u) / 4u; wd += kFuThreads) dst[wd] = __byte_perm(obuf[wd], 0u, 0x0123);
        if (tid == 0) frame_len[f] = total;
    }
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
    // analysis kernel: tile | scratch | FuShared
    L.ovl_off = up(L.tile_words * 4u, 128u);
    uint32_t o = 0;
    L.acstore_off = o;   o += up(kFuSig * L.n_win * kAcStoreStride * 8u, 16u);
    L.ws_off = o;        o += up(kFuAnWarps * (uint32_t)sizeof(WarpScratch), 16u);
    L.psum_off = o;      o += kFuAnWarps * 2u * kMaxParts * 8u;
    L.fixsum_off = o;    o += kFuSig * 5u * (1u << P.max_part_order) * 8u;
    L.baseplan_off = o;  o += kFuSig * (uint32_t)sizeof(SubframePlan);
    L.stepplan_off = o;  o += kFuSig * std::max(1u, L.n_steps) * (uint32_t)sizeof(SubframePlan);
    L.ovl_bytes = up(o, 128u);
    L.shared_off = L.ovl_off + L.ovl_bytes;
    L.crctab_off = 0;
    L.total_bytes = up(L.shared_off + (uint32_t)sizeof(FuShared), 16u);
    // pack kernel: tile | frame image | FuShared | CRC tables
    L.pk_obuf_off = up(L.tile_words * 4u, 128u);
    L.pk_shared_off = L.pk_obuf_off + up(L.obuf_words * 4u + 16u, 128u);
    L.pk_crctab_off = up(L.pk_shared_off + (uint32_t)sizeof(FuShared), 16u);
    L.pk_total_bytes = L.pk_crctab_off + 4u * 256u * 2u;
    return L;
}

// Can this batch take the TMA-staged kernels of this file?  16-bit stereo in an int16 container, no loose mid/side (its
// followers wait for a decision made in another frame), tiles small enough for at least two CTAs per SM.
bool fused_eligible(const EncParams& P, uint32_t scratch_stride, int max_smem_optin) {
    if (!(P.container_bytes == 2 && P.channels == 2) || P.loose_frames) return false;
    if (P.max_lpc_order > kMaxOrder) return false;
    const FuLayout L = fused_layout(P, scratch_stride);
    const uint32_t m = std::max(L.total_bytes, L.pk_total_bytes);
    return (int)m <= max_smem_optin && m <= 110u * 1024u;
}

// ac = the autocorrelations autoc_kernel (enc_analyze.cu, un-shifted mode) left in the per-frame work records:
// frame f's [signal][window][kAcStoreStride] doubles start at ac + f * ac_frame_stride bytes.  plans: 2 per frame.
void launch_fused(const void* pcm, const FrameDesc* frames, const EncParams& P, int n_frames, const void* ac, size_t ac_frame_stride,
                  SubframePlan* plans, uint8_t* frame_ca, EncStats* stats, uint8_t* scratch, uint32_t scratch_stride, uint32_t* frame_len,
                  cudaStream_t stream, cudaEvent_t ev_after_analyze) {
    const FuLayout L = fused_layout(P, scratch_stride);
    cudaFuncSetAttribute(fused_analyze_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total_bytes);
    fused_analyze_kernel<<<(unsigned)n_frames, kFuAnThreads, L.total_bytes, stream>>>((const int16_t*)pcm, frames, P, L, (const double*)ac, ac_frame_stride,
                                                                                    plans, frame_ca, stats);
    if (ev_after_analyze) cudaEventRecord(ev_after_analyze, stream);
    cudaFuncSetAttribute(fused_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.pk_total_bytes);
    fused_pack_kernel<<<(unsigned)n_frames, kFuThreads, L.pk_total_bytes, stream>>>((const int16_t*)pcm, frames, P, L, plans, frame_ca, scratch, scratch_stride, frame_len);
}

}  // namespace fb
