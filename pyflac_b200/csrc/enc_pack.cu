// enc_pack.cu -- bitstream assembly: plans -> FLAC frame bytes (scope rows E12/E13).
//
// One CTA (8 warps) per frame.  The coded signals are rebuilt in shared memory, residuals are
// recomputed from the plan (nothing but the 128-byte plan crosses HBM between analysis and packing),
// every Rice partition gets its exact bit length (warp reduction), an exclusive scan over
// {frame header, subframe headers, partitions} gives each its absolute bit position, then all warps
// emit codes with a warp-shuffle prefix scan of code lengths and atomicOr into a zeroed frame image.
// CRC-16 is computed over 256 parallel chunks and combined in GF(2)[x]/(x^16+x^15+x^2+1).
// up: stream_encoder_framing.c FLAC__frame_add_header / FLAC__subframe_add_* ,
//     bitwriter.c FLAC__bitwriter_write_rice_signed_block (ref: format.h:209-475).
#include <stdlib.h>
#include <algorithm>
#include "fb_common.cuh"
#include "fb_math.cuh"
#include "enc_dev.cuh"

namespace fb {

constexpr int kPackThreads = 256;
constexpr int kPackWarps = kPackThreads / 32;
constexpr int kTileK = 4;                                   // residuals kept in registers per thread and tile
constexpr int kTileSamples = kPackThreads * kTileK;         // 4096 samples per tile
constexpr int kSegSamples = 32 * kTileK;                    // contiguous samples owned by one warp within a tile

struct PackShared {
    SubframePlan plan[kMaxChannels];
    uint32_t segtot[kPackWarps];         // code bits produced by each warp in the current tile
    int32_t  sigidx[kMaxChannels];       // which analysed signal is coded as channel c
    uint16_t crc_tab[256];
    uint32_t crc_warp[kPackWarps];
    uint8_t  hdr[16];
    uint32_t hdr_len, x1;
};

// Rice-coded body of one subframe (up: add_residual_partitioned_rice_ + FLAC__bitwriter_write_rice_signed_block).
// The whole CTA works on one subframe at a time, tile by tile: thread (warp w, lane l) owns samples
//   i = t0 + w*512 + k*32 + l,  k = 0..15
// computes each residual ONCE (kept zig-zag folded in registers), the warp sums its code lengths, a
// CTA barrier publishes the 8 segment totals, then every 32-sample group runs a warp-shuffle prefix
// scan of code lengths to get absolute bit positions and ORs its codes into the frame image.
//   bit position of sample i = body + (code bits of all earlier samples) + plen * (partition(i) + 1)
// (a partition's parameter field sits right before its first sample's code).  Returns the body length in bits.
// Residual arithmetic: r = x[i] - ((sum q_j x[i-1-j]) >> shift); fixed predictors are the same formula with
// binomial coefficients and shift 0.  WIDE = 64-bit accumulate, chosen exactly as the analysis kernel does.
// S33: 33-bit samples (side channel of 32-bit stereo): x[] holds sample >> 1, lsb[] the dropped bits (always WIDE).
template <int ORDER, bool WIDE, bool S33 = false>
__device__ __noinline__ uint32_t pack_rice_body(const SubframePlan& pl, const int32_t* __restrict__ qs, int order, int shift,
                                                const int32_t* __restrict__ x, const uint32_t* __restrict__ lsb, int N, uint32_t body, uint32_t* obuf,
                                                PackShared& S, int warp, int lane) {
    auto X = [&](int i) -> long long {
        if (S33) return ((long long)x[i] << 1) | (long long)((lsb[i >> 5] >> (i & 31)) & 1u);
        return (long long)x[i];
    };
    // ORDER is the order CLASS (4, 8 or 12 taps, coefficients beyond the real order are zero): three code bodies
    // instead of thirteen keep the kernel inside the instruction cache
    int32_t q[ORDER];
#pragma unroll
    for (int j = 0; j < ORDER; j++) q[j] = (j < order) ? qs[j] : 0;
    const uint32_t plen = pl.rice2 ? 5u : 4u;
    const uint32_t psize = (uint32_t)N >> pl.part_order;
    const uint32_t magic = psize > 1u ? (uint32_t)((0x100000000ull + psize - 1u) / psize) : 0u;   // i / psize == umulhi(i, magic) for i, psize < 2^16; 0 = one-sample partitions
    auto pdiv = [&](uint32_t i) { return magic ? __umulhi(i, magic) : i; };
    uint32_t done_bits = 0;                                                      // code bits of all previous tiles
    for (int t0 = 0; t0 < N; t0 += kTileSamples) {
        uint32_t u[kTileK];
        uint32_t mybits = 0;
        const int seg0 = t0 + warp * kSegSamples;
#pragma unroll
        for (int k = 0; k < kTileK; k++) {
            const int i = seg0 + k * 32 + lane;
            u[k] = 0xffffffffu;                                  // marks "no sample"
            if (i >= order && i < N) {
                int32_t r;
                if (WIDE) {
                    long long sacc = 0;
#pragma unroll
                    for (int j = 0; j < ORDER; j++) sacc += (long long)q[j] * X(max(i - 1 - j, 0));
                    r = (int32_t)(X(i) - (sacc >> shift));
                } else {
                    int sacc = 0;
#pragma unroll
                    for (int j = 0; j < ORDER; j++) sacc += q[j] * x[max(i - 1 - j, 0)];
                    r = x[i] - (sacc >> shift);
                }
                const uint32_t uu = ((uint32_t)r << 1) ^ (uint32_t)(r >> 31);
                const uint32_t kk = pl.rice[pdiv((uint32_t)i)];
                u[k] = uu;
                mybits += (uu >> kk) + 1u + kk;
            }
        }
        mybits = __reduce_add_sync(0xffffffffu, mybits);
        if (lane == 0) S.segtot[warp] = mybits;
        __syncthreads();
        uint32_t wpos = body + done_bits, tile_bits = 0;
#pragma unroll
        for (int w2 = 0; w2 < kPackWarps; w2++) { const uint32_t t = S.segtot[w2]; if (w2 < warp) wpos += t; tile_bits += t; }
#pragma unroll
        for (int k = 0; k < kTileK; k++) {
            const int i = seg0 + k * 32 + lane;
            const bool valid = (i >= order && i < N);
            uint32_t kk = 0, part = 0, len = 0;
            if (valid) {
                part = pdiv((uint32_t)i);
                kk = pl.rice[part];
                len = (u[k] >> kk) + 1u + kk;
            }
            uint32_t incl = len;                                  // warp-shuffle inclusive prefix scan of code lengths
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (valid) {
                const uint32_t pos = wpos + (incl - len) + plen * (part + 1u);
                put_bits(obuf, pos + (u[k] >> kk), (1u << kk) | (u[k] & ((1u << kk) - 1u)), kk + 1u);
                if ((uint32_t)i == part * psize || i == order) put_bits(obuf, pos - plen, kk, plen);   // first sample of its partition
            }
            wpos += __shfl_sync(0xffffffffu, incl, 31);
        }
        done_bits += tile_bits;
        __syncthreads();                                          // segtot is reused by the next tile / subframe
    }
    return done_bits + plen * (1u << pl.part_order);
}

template <bool WIDE>
__device__ __forceinline__ uint32_t pack_rice_dispatch(int order, const SubframePlan& pl, const int32_t* q, int shift, const int32_t* x,
                                                       const uint32_t* lsb, bool s33, int N, uint32_t body, uint32_t* obuf, PackShared& S, int warp, int lane) {
    if (s33) {
        if (order <= 4) return pack_rice_body<4, true, true>(pl, q, order, shift, x, lsb, N, body, obuf, S, warp, lane);
        if (order <= 8) return pack_rice_body<8, true, true>(pl, q, order, shift, x, lsb, N, body, obuf, S, warp, lane);
        return pack_rice_body<12, true, true>(pl, q, order, shift, x, lsb, N, body, obuf, S, warp, lane);
    }
    if (order <= 4) return pack_rice_body<4, WIDE>(pl, q, order, shift, x, lsb, N, body, obuf, S, warp, lane);
    if (order <= 8) return pack_rice_body<8, WIDE>(pl, q, order, shift, x, lsb, N, body, obuf, S, warp, lane);
    return pack_rice_body<12, WIDE>(pl, q, order, shift, x, lsb, N, body, obuf, S, warp, lane);
}

// B32: 32-bit input (its side channel can need 33 bits); every other depth compiles without that machinery.
template <typename PcmT, bool B32>
__global__ void __launch_bounds__(kPackThreads, 4)
pack_kernel(const PcmT* __restrict__ pcm, const FrameDesc* __restrict__ frames, EncParams P,
            const SubframePlan* __restrict__ plans, const uint8_t* __restrict__ frame_ca,
            uint8_t* __restrict__ scratch, uint32_t scratch_stride, uint32_t* __restrict__ frame_len, uint32_t obuf_words) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const FrameDesc fd = frames[blockIdx.x];
    const int N = (int)fd.blocksize, ch = (int)P.channels, nsig = (int)P.n_signals;
    const int ca = frame_ca[blockIdx.x];

    int32_t* xall = reinterpret_cast<int32_t*>(smem_raw);
    uint32_t* obuf = reinterpret_cast<uint32_t*>(smem_raw + (size_t)ch * P.smem_stride * 4);
    PackShared& S = *reinterpret_cast<PackShared*>(smem_raw + (size_t)ch * P.smem_stride * 4 + (size_t)obuf_words * 4);
    uint32_t* lsb = reinterpret_cast<uint32_t*>(smem_raw + (size_t)ch * P.smem_stride * 4 + (size_t)obuf_words * 4 + ((sizeof(PackShared) + 15) / 16) * 16);   // 33-bit side: dropped low bits

    // ---- zero the frame image, CRC table, coded-signal map, frame header ----
    for (uint32_t i = tid; i < obuf_words; i += kPackThreads) obuf[i] = 0u;
    {
        uint16_t c = (uint16_t)(tid << 8);
#pragma unroll
        for (int i = 0; i < 8; i++) c = (uint16_t)((c & 0x8000) ? ((c << 1) ^ 0x8005) : (c << 1));
        S.crc_tab[tid] = c;
    }
    if (tid < ch) {
        int si = tid;
        if (P.do_mid_side) {
            if (ca == 1) si = (tid == 0) ? 0 : 3;
            else if (ca == 2) si = (tid == 0) ? 3 : 1;
            else if (ca == 3) si = (tid == 0) ? 2 : 3;
        }
        S.sigidx[tid] = si;
    }
    if (tid == 0) S.hdr_len = (uint32_t)build_frame_header(S.hdr, P.channels, P.bps, P.sample_rate, (uint32_t)N, fd.frame_number, ca);
    __syncthreads();
    for (int i = tid; i < ch * 32; i += kPackThreads) {
        const int c = i >> 5, wd = i & 31;
        reinterpret_cast<uint32_t*>(&S.plan[c])[wd] =
            reinterpret_cast<const uint32_t*>(plans + (size_t)blockIdx.x * nsig + S.sigidx[c])[wd];
    }
    __syncthreads();

    // ---- rebuild the coded signals (wasted bits removed) ----
    {
        const PcmT* base = pcm + fd.pcm_off;
        if (ch == 2) {
            // stereo: one pass, both channels of a sample loaded once (one 32-bit word for an aligned int16 container); four
            // samples per thread are in flight before the first one is used
            const int si0 = S.sigidx[0], si1 = S.sigidx[1], w0 = S.plan[0].wasted, w1 = S.plan[1].wasted;
            const bool s33_0 = B32 && S.plan[0].sbps > 32, s33_1 = B32 && S.plan[1].sbps > 32;
            const bool word_ok = sizeof(PcmT) == 2 && ((reinterpret_cast<uintptr_t>(base) & 3u) == 0);
            int32_t* x0 = xall; int32_t* x1 = xall + P.smem_stride;
            for (int i0 = 0; i0 < N; i0 += 4 * kPackThreads) {
                long long l[4], r[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int i = i0 + u * kPackThreads + tid;
                    l[u] = 0; r[u] = 0;
                    if (i < N) {
                        if (word_ok) { const int wd = __ldg(reinterpret_cast<const int*>(base) + i); l[u] = (long long)(short)wd; r[u] = (long long)(wd >> 16); }
                        else { l[u] = (long long)__ldg(base + 2 * (uint64_t)i); r[u] = (long long)__ldg(base + 2 * (uint64_t)i + 1); }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int i = i0 + u * kPackThreads + tid;
                    if (i0 + u * kPackThreads >= N) break;                      // uniform: whole CTA past the end
                    const long long m = (l[u] + r[u]) >> 1, sd = l[u] - r[u];  // 64-bit: 32-bit input needs 33 bits here
                    long long v0 = (si0 == 0 ? l[u] : si0 == 1 ? r[u] : si0 == 2 ? m : sd) >> w0;
                    long long v1 = (si1 == 0 ? l[u] : si1 == 1 ? r[u] : si1 == 2 ? m : sd) >> w1;
                    if (s33_0) { const uint32_t word = __ballot_sync(0xffffffffu, (v0 & 1ll) != 0); if (lane == 0 && i < N) lsb[i >> 5] = word; v0 >>= 1; }
                    if (s33_1) { const uint32_t word = __ballot_sync(0xffffffffu, (v1 & 1ll) != 0); if (lane == 0 && i < N) lsb[i >> 5] = word; v1 >>= 1; }
                    if (i < N) { x0[i] = (int32_t)v0; x1[i] = (int32_t)v1; }
                }
            }
        } else {
            for (int c = 0; c < ch; c++) {
                const int si = S.sigidx[c], wasted = S.plan[c].wasted;
                int32_t* x = xall + (size_t)c * P.smem_stride;
                for (int i = tid; i < N; i += kPackThreads) x[i] = (int)__ldg(base + (uint64_t)i * ch + si) >> wasted;
            }
        }
    }
    if (tid < (int)S.hdr_len) put_bits(obuf, (uint32_t)tid * 8u, S.hdr[tid], 8);
    __syncthreads();

    // ---- subframes, in channel order (each one starts where the previous one ended) ----
    uint32_t pos = S.hdr_len * 8u;
    for (int c = 0; c < ch; c++) {
        const SubframePlan& pl = S.plan[c];
        const int32_t* x = xall + (size_t)c * P.smem_stride;
        const uint32_t sbps = pl.sbps, order = pl.order, wf = pl.wasted ? 1u : 0u;
        const bool s33 = B32 && sbps > 32;
        auto X = [&](int i) -> long long { return s33 ? (((long long)x[i] << 1) | (long long)((lsb[i >> 5] >> (i & 31)) & 1u)) : (long long)x[i]; };
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
                if (pl.type == kConstant) put_bits64(obuf, after_hdr, X(0), sbps);
            }
            if (pl.type == kFixed || pl.type == kLpc) {
                if ((uint32_t)lane < order) put_bits64(obuf, after_hdr + (uint32_t)lane * sbps, X(lane), sbps);
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
            for (int i = tid; i < N; i += kPackThreads) put_bits64(obuf, after_hdr + (uint32_t)i * sbps, X(i), sbps);
            pos = after_hdr + (uint32_t)N * sbps;
        } else {
            uint32_t body = after_hdr + order * sbps + 6u, blen;
            if (pl.type == kLpc) {
                body += 9u + order * pl.precision;
                int32_t asum = 0;
                for (uint32_t j = 0; j < order; j++) asum += abs(pl.qlp[j]);
                if (asum == 0) asum = 1;
                // same accumulator-width rule as the analysis kernel (up: FLAC__lpc_max_prediction_before_shift_bps)
                if ((int)sbps + (int)silog2((int64_t)asum) <= 32) blen = pack_rice_dispatch<false>((int)order, pl, pl.qlp, pl.shift, x, lsb, s33, N, body, obuf, S, warp, lane);
                else blen = pack_rice_dispatch<true>((int)order, pl, pl.qlp, pl.shift, x, lsb, s33, N, body, obuf, S, warp, lane);
            } else {
                const int32_t cfix[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};
                if (s33) blen = pack_rice_body<4, true, true>(pl, cfix[order], (int)order, 0, x, lsb, N, body, obuf, S, warp, lane);
                else blen = pack_rice_body<4, false>(pl, cfix[order], (int)order, 0, x, lsb, N, body, obuf, S, warp, lane);
            }
            pos = body + blen;
        }
    }
    __syncthreads();

    // ---- CRC-16 over the byte-padded frame, append, store ----
    const uint32_t nb = (pos + 7u) >> 3;
    {
        const uint16_t c2 = cta_crc16<kPackThreads>([&](uint32_t j) { return (uint8_t)(obuf[j >> 2] >> (24 - 8 * (int)(j & 3u))); },
                                                    nb, S.crc_tab, S.crc_warp, tid);
        if (tid == 0) put_bits(obuf, nb * 8u, c2, 16);
    }
    __syncthreads();
    {
        const uint32_t total = nb + 2u;
        uint32_t* dst = reinterpret_cast<uint32_t*>(scratch + (size_t)blockIdx.x * scratch_stride);
        for (uint32_t wd = tid; wd < (total + 3u) / 4u; wd += kPackThreads) dst[wd] = __byte_perm(obuf[wd], 0u, 0x0123);
        if (tid == 0) frame_len[blockIdx.x] = total;
    }
}

void launch_pack(const void* pcm, const FrameDesc* frames, const EncParams& P, int n_frames, const SubframePlan* plans,
                 const uint8_t* frame_ca, uint8_t* scratch, uint32_t scratch_stride, uint32_t* frame_len,
                 cudaStream_t stream) {
    const uint32_t obuf_words = scratch_stride / 4 + 4;
    const size_t smem = (size_t)P.channels * P.smem_stride * 4 + (size_t)obuf_words * 4 + ((sizeof(PackShared) + 15) / 16) * 16 + ((size_t)P.blocksize + 31) / 32 * 4 + 16;
    if (P.container_bytes == 2) {
        cudaFuncSetAttribute(pack_kernel<int16_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        pack_kernel<int16_t, false><<<n_frames, kPackThreads, smem, stream>>>((const int16_t*)pcm, frames, P, plans, frame_ca, scratch, scratch_stride, frame_len, obuf_words);
    } else if (P.bps < 32) {
        cudaFuncSetAttribute(pack_kernel<int32_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        pack_kernel<int32_t, false><<<n_frames, kPackThreads, smem, stream>>>((const int32_t*)pcm, frames, P, plans, frame_ca, scratch, scratch_stride, frame_len, obuf_words);
    } else {
        cudaFuncSetAttribute(pack_kernel<int32_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        pack_kernel<int32_t, true><<<n_frames, kPackThreads, smem, stream>>>((const int32_t*)pcm, frames, P, plans, frame_ca, scratch, scratch_stride, frame_len, obuf_words);
    }
}

size_t pack_smem_bytes(const EncParams& P, uint32_t scratch_stride) {
    return (size_t)P.channels * P.smem_stride * 4 + (size_t)(scratch_stride / 4 + 4) * 4 + ((sizeof(PackShared) + 15) / 16) * 16 + ((size_t)P.blocksize + 31) / 32 * 4 + 16;
}

// ------------------------------------------------------------------ layout: scan, compaction, stream prologue ----

// Exclusive scan of frame byte lengths (frames are ordered by stream, then frame number) with a
// per-stream prologue gap, so that every stream's .flac image is contiguous in the output arena.
__global__ void __launch_bounds__(1024)
scan_kernel(const uint32_t* __restrict__ frame_len, const FrameDesc* __restrict__ frames, int n_frames,
            uint32_t prologue_bytes, uint64_t base, uint64_t* __restrict__ frame_off, uint64_t* __restrict__ total_bytes) {
    // tiles of 4096 frames: four consecutive frames per thread, warp-shuffle scan, one pass over the 32 warp totals per tile
    // (was: every thread walked its own run of frames with dependent, uncoalesced loads, then 20 barriers of Hillis-Steele: 80 us)
    __shared__ unsigned long long warp_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long carry = base;
    for (int t0 = 0; t0 < n_frames; t0 += 4096) {
        const int f0 = t0 + 4 * tid;
        uint32_t len[4]; unsigned long long v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int f = f0 + k;
            len[k] = 0; v[k] = 0;
            if (f < n_frames) {
                const bool first = (f == 0) || (frames[f].stream != frames[f - 1].stream);
                len[k] = frame_len[f];
                v[k] = (unsigned long long)len[k] + (first ? prologue_bytes : 0u);
            }
        }
        v[1] += v[0]; v[2] += v[1]; v[3] += v[2];                       // inclusive within the thread
        unsigned long long x = v[3];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const unsigned long long before = carry + (warp ? warp_tot[warp - 1] : 0ull) + (x - v[3]);   // everything in front of this thread's four frames
#pragma unroll
        for (int k = 0; k < 4; k++) if (f0 + k < n_frames) frame_off[f0 + k] = before + v[k] - len[k];
        carry += warp_tot[31];
        __syncthreads();
    }
    if (tid == 0) *total_bytes = carry - base;
}

__global__ void __launch_bounds__(256)
compact_kernel(const uint8_t* __restrict__ scratch, uint32_t scratch_stride, const uint32_t* __restrict__ frame_len,
               const uint64_t* __restrict__ frame_off, uint8_t* __restrict__ arena) {
    const uint8_t* src = scratch + (size_t)blockIdx.x * scratch_stride;
    uint8_t* dst = arena + frame_off[blockIdx.x];
    const uint32_t n = frame_len[blockIdx.x];
    // head bytes up to 4-byte alignment of dst, then 32-bit stores assembled from two aligned source words
    const uint32_t head = min(n, (uint32_t)((4u - ((uintptr_t)dst & 3u)) & 3u));
    if (threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
    const uint32_t nwords = (n - head) / 4u;
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst + head);
    const uint32_t sh = head * 8u;
    for (uint32_t wd = threadIdx.x; wd < nwords; wd += 256) {
        const uint32_t a = s32[wd], b = (sh ? s32[wd + 1] : 0u);
        d32[wd] = sh ? __funnelshift_r(a, b, sh) : a;
    }
    const uint32_t tail0 = head + nwords * 4u;
    if (threadIdx.x < n - tail0) dst[tail0 + threadIdx.x] = src[tail0 + threadIdx.x];
}

// One thread per stream: min/max frame size, total length, and the stream prologue
// "fLaC" + STREAMINFO + VORBIS_COMMENT(vendor) (ref: format.h:546-557, :631-647).
__global__ void finalize_kernel(const uint32_t* __restrict__ frame_len, const uint64_t* __restrict__ frame_off,
                                const uint32_t* __restrict__ stream_first, const uint32_t* __restrict__ stream_nframes,
                                const uint64_t* __restrict__ stream_samples, const uint8_t* __restrict__ md5,
                                int n_streams, EncParams P, uint32_t write_prologue,
                                uint8_t* __restrict__ arena, StreamInfoOut* __restrict__ info) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    const uint32_t f0 = stream_first[s], nf = stream_nframes[s];
    uint32_t mn = 0, mx = 0; uint64_t bytes = 0;
    for (uint32_t f = f0; f < f0 + nf; f++) {
        const uint32_t l = frame_len[f];
        if (mn == 0 || l < mn) mn = l;
        if (l > mx) mx = l;
        bytes += l;
    }
    StreamInfoOut o;
    o.total_samples = stream_samples[s];
    const uint64_t pro = write_prologue ? (uint64_t)kStreamPrologueBytes : 0ull;
    o.byte_off = nf ? frame_off[f0] - pro : 0ull;
    o.byte_len = bytes + pro;
    o.min_framesize = mn; o.max_framesize = mx; o.n_frames = nf; o.pad = 0;
    for (int i = 0; i < 16; i++) o.md5[i] = md5 ? md5[(size_t)s * 16 + i] : 0;
    info[s] = o;
    if (write_prologue && nf) {
        uint8_t* p = arena + o.byte_off;
        p[0] = 'f'; p[1] = 'L'; p[2] = 'a'; p[3] = 'C';
        p[4] = 0x00; p[5] = 0; p[6] = 0; p[7] = 34;
        const uint32_t bs = P.blocksize;
        p[8] = (uint8_t)(bs >> 8); p[9] = (uint8_t)bs; p[10] = (uint8_t)(bs >> 8); p[11] = (uint8_t)bs;
        p[12] = (uint8_t)(mn >> 16); p[13] = (uint8_t)(mn >> 8); p[14] = (uint8_t)mn;
        p[15] = (uint8_t)(mx >> 16); p[16] = (uint8_t)(mx >> 8); p[17] = (uint8_t)mx;
        const uint64_t ts = o.total_samples & 0xFFFFFFFFFull;
        const uint64_t v = ((uint64_t)P.sample_rate << 44) | ((uint64_t)(P.channels - 1) << 41) | ((uint64_t)(P.bps - 1) << 36) | ts;
        for (int i = 0; i < 8; i++) p[18 + i] = (uint8_t)(v >> (56 - 8 * i));
        for (int i = 0; i < 16; i++) p[26 + i] = o.md5[i];
        const char vendor[] = "reference libFLAC 1.4.3 20230623";   // must equal the oracle's for byte-identical files
        const uint32_t vl = 32, len = 4 + vl + 4;
        p[42] = 0x84; p[43] = 0; p[44] = 0; p[45] = (uint8_t)len;
        p[46] = (uint8_t)vl; p[47] = 0; p[48] = 0; p[49] = 0;
        for (uint32_t i = 0; i < vl; i++) p[50 + i] = (uint8_t)vendor[i];
        p[82] = 0; p[83] = 0; p[84] = 0; p[85] = 0;
    }
}

// ------------------------------------------------------------------ MD5 (scope row E1) ----
// up: md5.c FLAC__MD5Accumulate -- MD5 over (bps+7)/8 little-endian bytes per sample, interleaved.
// Serial per stream by construction (Merkle-Damgard chain); one thread per stream, runs on a side
// CUDA stream concurrently with analysis/packing.  `state` carries (a,b,c,d,len,buffered bytes)
// across calls so that streaming callers can feed a stream in pieces.

__device__ __forceinline__ uint32_t rotl32(uint32_t v, int s) { return __funnelshift_l(v, v, s); }

__device__ void md5_block(uint32_t h[4], const uint32_t w[16]) {
    const uint32_t K[64] = {
        0xd76aa478,0xe8c7b756,0x242070db,0xc1bdceee,0xf57c0faf,0x4787c62a,0xa8304613,0xfd469501,
        0x698098d8,0x8b44f7af,0xffff5bb1,0x895cd7be,0x6b901122,0xfd987193,0xa679438e,0x49b40821,
        0xf61e2562,0xc040b340,0x265e5a51,0xe9b6c7aa,0xd62f105d,0x02441453,0xd8a1e681,0xe7d3fbc8,
        0x21e1cde6,0xc33707d6,0xf4d50d87,0x455a14ed,0xa9e3e905,0xfcefa3f8,0x676f02d9,0x8d2a4c8a,
        0xfffa3942,0x8771f681,0x6d9d6122,0xfde5380c,0xa4beea44,0x4bdecfa9,0xf6bb4b60,0xbebfbc70,
        0x289b7ec6,0xeaa127fa,0xd4ef3085,0x04881d05,0xd9d4d039,0xe6db99e5,0x1fa27cf8,0xc4ac5665,
        0xf4292244,0x432aff97,0xab9423a7,0xfc93a039,0x655b59c3,0x8f0ccc92,0xffeff47d,0x85845dd1,
        0x6fa87e4f,0xfe2ce6e0,0xa3014314,0x4e0811a1,0xf7537e82,0xbd3af235,0x2ad7d2bb,0xeb86d391 };
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3];
// critical path per step: LOP3 (f depends on the fresh b) -> IADD3 -> LEA.HI (rotate + add).  K[i] + w[g] is formed by
// an opaque add so that it stays OFF the chain (left to itself the compiler adds K after the 3-input add, a 4th
// dependent instruction: 23 instead of ~15 cycles per step).
#define FB_MD5_STEP(f, g, s, i) { uint32_t kw; asm volatile("add.u32 %0, %1, %2;" : "=r"(kw) : "r"(w[g]), "r"(K[i])); \
                                  const uint32_t t = (a + kw) + (f); a = d; d = c; c = b; b = b + rotl32(t, s); }
#pragma unroll
    for (int i = 0; i < 16; i++) { const int sh[4] = {7, 12, 17, 22}; FB_MD5_STEP((b & c) | (~b & d), i, sh[i & 3], i) }
#pragma unroll
    for (int i = 16; i < 32; i++) { const int sh[4] = {5, 9, 14, 20}; FB_MD5_STEP((d & b) | (~d & c), (5 * i + 1) & 15, sh[i & 3], i) }
#pragma unroll
    for (int i = 32; i < 48; i++) { const int sh[4] = {4, 11, 16, 23}; FB_MD5_STEP(b ^ c ^ d, (3 * i + 5) & 15, sh[i & 3], i) }
#pragma unroll
    for (int i = 48; i < 64; i++) { const int sh[4] = {6, 10, 15, 21}; FB_MD5_STEP(c ^ (b | ~d), (7 * i) & 15, sh[i & 3], i) }
#undef FB_MD5_STEP
    h[0] += a; h[1] += b; h[2] += c; h[3] += d;
}

__device__ __forceinline__ void md5_block4(uint32_t h[4], const uint4& n0, const uint4& n1, const uint4& n2, const uint4& n3) {
    const uint32_t w[16] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, n3.x, n3.y, n3.z, n3.w};
    md5_block(h, w);
}

// Generic feeder: packs (bps+7)/8 little-endian bytes of each sample into 32-bit words, one sample at a time.
struct Md5Feeder {
    uint32_t h[4]; uint32_t w[16]; uint64_t acc; uint32_t accbits, widx;
    __device__ void init() { h[0] = 0x67452301u; h[1] = 0xefcdab89u; h[2] = 0x98badcfeu; h[3] = 0x10325476u; acc = 0; accbits = 0; widx = 0; }
    __device__ void flush_words() {
        while (accbits >= 32) {
            // static indexing keeps w[] in registers
#pragma unroll
            for (int t = 0; t < 16; t++) if ((int)widx == t) w[t] = (uint32_t)acc;
            widx++; acc >>= 32; accbits -= 32;
            if (widx == 16) { md5_block(h, w); widx = 0; }
        }
    }
    __device__ void push(uint32_t v, uint32_t nbytes) {
        acc |= (uint64_t)(nbytes == 4 ? v : (v & ((1u << (8 * nbytes)) - 1u))) << accbits;
        accbits += 8 * nbytes;
        flush_words();
    }
    __device__ void finish(uint64_t total_bytes, uint8_t* out) {
        push(0x80u, 1);
        while (!(accbits == 0 && widx == 14)) push(0u, 1);
        w[14] = (uint32_t)(total_bytes * 8ull); w[15] = (uint32_t)((total_bytes * 8ull) >> 32);
        md5_block(h, w);
        for (int i = 0; i < 16; i++) out[i] = (uint8_t)(h[i >> 2] >> (8 * (i & 3)));
    }
};

// One thread per stream, one warp (32 streams) per CTA.  Fast paths: the container bytes ARE the hashed bytes (int16
// container with 16-bit samples, int32 with 32-bit), or 24-bit samples in an int32 container (three of every four bytes,
// gathered with one PRMT per hashed word).  The chain's static schedule is 12 cycles per step (LEA.HI -> LOP3 -> IADD3),
// 768 cycles per 64-byte block -- about one DRAM round trip, and the instructions that feed it sit in the same in-order
// issue stream, so: the bytes travel global -> shared with cp.async in 256-byte pieces per stream (half a warp copies one
// stream's piece: whole 128-byte lines instead of 32 scattered 16-byte sectors per request), kMd5Ring pieces deep (two: the next piece lands while this one is hashed; shared memory is what a chain CTA takes from the
// encode CTAs next to it), from
// register-resident source addresses (three instructions per copy); each lane reads its own row back 16 bytes at a time
// (row stride 272 bytes: the eight lanes of an LDS.128 phase fall in distinct banks).  Measured on the bench batch
// (256 streams of 1.92 MB): 21.1 ms with per-lane loads one block ahead -> 13.3 ms alone.
#ifndef FB_MD5_RING
#define FB_MD5_RING 4
#endif
constexpr int kMd5Piece = 256, kMd5Row = kMd5Piece + 16, kMd5Ring = FB_MD5_RING;
// Host -> host path: the kernel is launched before the PCM has arrived; the copy stream sets flag c when chunk c (streams
// cs[c] .. cs[c+1]-1) is in HBM.  The warp waits for the chunk of its last stream (chunks land in order), so every chain starts
// the moment its bytes are there and all chains of a batch run side by side in ONE launch.
__device__ __noinline__ void md5_wait_for_chunk(const Md5Gate& gate, int n_streams) {
    const int last = min(((int)blockIdx.x * (int)(blockDim.x >> 5) + (int)(threadIdx.x >> 5)) * 32 + 31, n_streams - 1);
    int c = 0;
    while (c + 1 < gate.nchunks && last >= gate.cs[c + 1]) c++;
    // every lane polls (one broadcast load per try): a single polling lane left the warp split behind the wait -- lane 0 and
    // lanes 1..31 then ran the whole hash as two passes, 2x the time -- whatever __syncwarp() followed
    const volatile uint32_t* f = gate.flags + c;
    unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (__any_sync(0xffffffffu, *f == 0u)) {
        __nanosleep(500);
        // the flag is set by the copy stream; if that stream cannot make progress while this kernel runs (a profiler that
        // serialises all work) give up loudly after 30 s instead of hanging the GPU
        unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 30000000000ull) __trap();
    }
    __threadfence();
}
template <typename PcmT, bool K24>
__global__ void __launch_bounds__(128) md5_kernel(const PcmT* __restrict__ pcm, const uint64_t* __restrict__ stream_pcm_off,
                           const uint64_t* __restrict__ stream_samples, int n_streams, uint32_t channels, uint32_t bps,
                           uint8_t* __restrict__ digest_out, Md5Gate gate) {
    // one warp = 32 chains; the warps of a CTA never meet (no CTA barrier): a CTA of four warps puts one chain warp on each SM
    // sub-partition, so a batch of 256 streams occupies 2 SMs instead of 8 (see launch_md5_gated)
    extern __shared__ __align__(16) uint8_t md5_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int w0 = (blockIdx.x * (blockDim.x >> 5) + wib) * 32, s = w0 + lane;
    if (w0 >= n_streams) return;                                           // whole warp past the end
    uint8_t (*ring)[32][kMd5Row] = reinterpret_cast<uint8_t (*)[32][kMd5Row]>(md5_smem + (size_t)wib * (kMd5Ring * 32 * kMd5Row));
    const bool live = s < n_streams;
    if (gate.flags) md5_wait_for_chunk(gate, n_streams);               // host -> host path: launched ahead of its input
    const uint32_t bytes_per = (bps + 7) / 8;
    // lanes past the last stream mirror the warp's first stream for the copies (valid addresses, nothing hashed)
    const PcmT* p = pcm + stream_pcm_off[live ? s : w0];
    const uint64_t nvals = live ? stream_samples[s] * channels : 0ull;
    Md5Feeder f; f.init();
    const bool fast = live && bytes_per == (K24 ? 3u : (uint32_t)sizeof(PcmT)) && (((uintptr_t)p) & 15u) == 0;
    const uint64_t np_own = fast ? (nvals * sizeof(PcmT)) / kMd5Piece : 0ull;     // whole pieces of this lane's stream
    uint64_t np_max = np_own, np_all = live ? np_own : ~0ull;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const uint64_t v = __shfl_xor_sync(0xffffffffu, np_max, o), u = __shfl_xor_sync(0xffffffffu, np_all, o);
        np_max = v > np_max ? v : np_max; np_all = u < np_all ? u : np_all;
    }
    if (np_max) {                                                          // warp-uniform
        const int half = lane >> 4, l16 = lane & 15;
        // Copy duty of a lane: 16 bytes of the piece of streams 2j + half, j = 0..15.  Pieces that every stream of the warp
        // has (q < np_all) are requested with no per-stream test.
        uint64_t src[16];
#pragma unroll
        for (int j = 0; j < 16; j++) src[j] = __shfl_sync(0xffffffffu, (uint64_t)(uintptr_t)p, 2 * j + half) + (uint64_t)l16 * 16;
        auto request = [&](uint64_t q) {
            const int slot = (int)(q % kMd5Ring);
            const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(&ring[slot][half][l16 * 16]);
            if (q < np_all) {
#pragma unroll
                for (int j = 0; j < 16; j++)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst0 + (uint32_t)(2 * j * kMd5Row)), "l"(src[j] + q * kMd5Piece) : "memory");
            } else if (q < np_max) {
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const uint64_t np = __shfl_sync(0xffffffffu, np_own, 2 * j + half);
                    if (q < np)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst0 + (uint32_t)(2 * j * kMd5Row)), "l"(src[j] + q * kMd5Piece) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");            // requests past the end commit empty groups
        };
        for (int q = 0; q < kMd5Ring - 1; q++) request((uint64_t)q);
        for (uint64_t q = 0; q < np_max; q++) {
            asm volatile("cp.async.wait_group %0;" :: "n"(kMd5Ring - 2) : "memory");
            __syncwarp();                                                  // piece q is there for every lane, and every lane is done with piece q-1
            request(q + kMd5Ring - 1);                                     // into the slot of piece q-1: a whole piece (3000 cycles of hashing) ahead even with two slots
            const uint4* row = reinterpret_cast<const uint4*>(&ring[q % kMd5Ring][lane][0]);
            if (q < np_own) {                                              // uniform while q < np_all
                if (K24) {
                    // 64 samples = 192 hashed bytes = three blocks; four samples (one 16-byte read) give three words
                    uint32_t w[48];
#pragma unroll
                    for (int g = 0; g < 16; g++) {
                        const uint4 x = row[g];
                        w[3 * g] = __byte_perm(x.x, x.y, 0x4210); w[3 * g + 1] = __byte_perm(x.y, x.z, 0x5421); w[3 * g + 2] = __byte_perm(x.z, x.w, 0x6542);
                    }
                    md5_block(f.h, w); md5_block(f.h, w + 16); md5_block(f.h, w + 32);
                } else {
                    // four blocks; the words of the next block are read while this one is hashed, two blocks per trip so that the
                    // two register sets alternate without moves
                    uint4 a0 = row[0], a1 = row[1], a2 = row[2], a3 = row[3];
#pragma unroll 1
                    for (int k = 0; k < 4; k += 2) {
                        const uint4 b0 = row[4 * k + 4], b1 = row[4 * k + 5], b2 = row[4 * k + 6], b3 = row[4 * k + 7];
                        md5_block4(f.h, a0, a1, a2, a3);
                        const int kn = (k + 2) & 3;
                        a0 = row[4 * kn]; a1 = row[4 * kn + 1]; a2 = row[4 * kn + 2]; a3 = row[4 * kn + 3];
                        md5_block4(f.h, b0, b1, b2, b3);
                    }
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    if (!live) return;
    // what the pieces did not cover (everything, for other container / sample-size pairs) goes through the generic feeder
    for (uint64_t i = np_own * (kMd5Piece / sizeof(PcmT)); i < nvals; i++) f.push((uint32_t)(int)__ldg(p + i), bytes_per);
    f.finish(nvals * bytes_per, digest_out + (size_t)s * 16);
}

void launch_md5_gated(const void* pcm, uint32_t container_bytes, const uint64_t* stream_pcm_off, const uint64_t* stream_samples,
                      int n_streams, uint32_t channels, uint32_t bps, uint8_t* digest_out, const Md5Gate& gate, cudaStream_t stream) {
    // Chains are latency bound: one warp (32 streams) per SM sub-partition is as fast as it gets (two warps on one sub-partition share
    // its ALU pipe: 8 of every 12 cycles each).  Four warps per CTA = one per sub-partition.  An MD5 CTA sits on its SM for 13-16 ms,
    // the lifetime of hundreds of encode CTAs, next to which it costs the encode kernels far more than its issue slots (measured:
    // 24 one-warp CTAs in flight on 24 SMs slowed the encode step by 0.7 ms); packed four to a CTA the chains of a 256-stream batch
    // touch 2 SMs.  The SM cannot change its L1 / shared-memory split while a CTA is resident, hence the largest carve-out.
    // (eight warps on one SM with a two-slot ring: 4.81 instead of 4.87 ms per step, at twice the latency of a single batch's digests)
    const int wpc = getenv("FLACB200_MD5_WARPS") ? std::max(1, std::min(4, atoi(getenv("FLACB200_MD5_WARPS")))) : 4;
    const int threads = 32 * wpc, nwarps = (n_streams + 31) / 32, blocks = (nwarps + wpc - 1) / wpc;
    // ... and it asks for (nearly) the whole shared memory of its SM, although the rings take 139 KB: encode CTAs that squeeze in next to
    // a chain CTA run at a fraction of their speed and hold up the end of every encode kernel (measured, 20 steps of the bench batch:
    // 5.62 ms per step with 70 KB chain CTAs, 5.13 with 139 KB, 4.88 with 190-226 KB; 4.64 without MD5).  Only while the chain CTAs are
    // few: a 4096-stream batch would take 32 SMs out of the encode kernels' hands.
    size_t smem = (size_t)wpc * kMd5Ring * 32 * kMd5Row;
    if (wpc >= 4 && blocks <= 16) smem = std::max<size_t>(smem, (size_t)200 * 1024);
    if (const char* ev = getenv("FLACB200_MD5_SMEM_KB")) smem = std::max<size_t>((size_t)wpc * kMd5Ring * 32 * kMd5Row, (size_t)atoi(ev) * 1024);
    static bool attr_set = false;
    if (!attr_set) {
        attr_set = true;
        const int mx = 227 * 1024;
        cudaFuncSetAttribute(md5_kernel<int16_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(md5_kernel<int32_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(md5_kernel<int32_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(md5_kernel<int16_t, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(md5_kernel<int32_t, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(md5_kernel<int32_t, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    if (container_bytes == 2) md5_kernel<int16_t, false><<<blocks, threads, smem, stream>>>((const int16_t*)pcm, stream_pcm_off, stream_samples, n_streams, channels, bps, digest_out, gate);
    else if ((bps + 7) / 8 == 3) md5_kernel<int32_t, true><<<blocks, threads, smem, stream>>>((const int32_t*)pcm, stream_pcm_off, stream_samples, n_streams, channels, bps, digest_out, gate);
    else md5_kernel<int32_t, false><<<blocks, threads, smem, stream>>>((const int32_t*)pcm, stream_pcm_off, stream_samples, n_streams, channels, bps, digest_out, gate);
}
void launch_md5(const void* pcm, uint32_t container_bytes, const uint64_t* stream_pcm_off, const uint64_t* stream_samples,
                int n_streams, uint32_t channels, uint32_t bps, uint8_t* digest_out, cudaStream_t stream) {
    Md5Gate none; none.flags = nullptr; none.nchunks = 0;
    launch_md5_gated(pcm, container_bytes, stream_pcm_off, stream_samples, n_streams, channels, bps, digest_out, none, stream);
}

// The MD5 chain of a batch ends long after its frames are final: the digests are patched into the finished stream
// images (STREAMINFO bytes 26..41 of each stream) and the per-stream info by this kernel on the batch's side stream.
__global__ void md5_patch_kernel(const uint8_t* __restrict__ md5, const uint32_t* __restrict__ stream_nframes, int n_streams,
                                 uint32_t write_prologue, uint8_t* __restrict__ arena, StreamInfoOut* __restrict__ info) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    for (int i = 0; i < 16; i++) info[s].md5[i] = md5[(size_t)s * 16 + i];
    if (write_prologue && stream_nframes[s]) {
        uint8_t* p = arena + info[s].byte_off + 26;
        for (int i = 0; i < 16; i++) p[i] = md5[(size_t)s * 16 + i];
    }
}
void launch_md5_patch(const uint8_t* md5, const uint32_t* stream_nframes, int n_streams, uint32_t write_prologue, uint8_t* arena,
                      StreamInfoOut* info, cudaStream_t stream) {
    md5_patch_kernel<<<(n_streams + 127) / 128, 128, 0, stream>>>(md5, stream_nframes, n_streams, write_prologue, arena, info);
}

void launch_layout(const uint32_t* frame_len, const FrameDesc* frames, int n_frames, uint32_t prologue_bytes, uint64_t base,
                   uint64_t* frame_off, uint64_t* total_bytes, cudaStream_t stream) {
    scan_kernel<<<1, 1024, 0, stream>>>(frame_len, frames, n_frames, prologue_bytes, base, frame_off, total_bytes);
}
void launch_compact(const uint8_t* scratch, uint32_t stride, const uint32_t* frame_len, const uint64_t* frame_off,
                    uint8_t* arena, int n_frames, cudaStream_t stream) {
    compact_kernel<<<n_frames, 256, 0, stream>>>(scratch, stride, frame_len, frame_off, arena);
}
void launch_finalize(const uint32_t* frame_len, const uint64_t* frame_off, const uint32_t* stream_first,
                     const uint32_t* stream_nframes, const uint64_t* stream_samples, const uint8_t* md5, int n_streams,
                     const EncParams& P, uint32_t write_prologue, uint8_t* arena, StreamInfoOut* info, cudaStream_t stream) {
    finalize_kernel<<<(n_streams + 127) / 128, 128, 0, stream>>>(frame_len, frame_off, stream_first, stream_nframes, stream_samples,
                                                                  md5, n_streams, P, write_prologue, arena, info);
}

}  // namespace fb
