// dec_kernels.cu -- decode path kernels (scope rows D1-D5).
//
// FLAC frames carry no length field and Rice codes are only self-delimiting sequentially, so the batch decoder is
// organised as: (1) find every byte position that COULD start a frame (sync code + plausible header + CRC-8),
// in parallel over all bytes of all streams; (2) decode every candidate independently, one thread per candidate
// frame (bit-serial Rice unpack + fixed/LPC synthesis -- a serial recurrence per subframe); (3) per stream, walk
// the chain "next frame starts where this one ended" from the first frame: true frames are always candidates,
// false candidates are never reached by the chain, so the result equals libFLAC's sequential parse;
// (4) post-process the chained frames with one CTA each: CRC-16, wasted bits are already applied, undo
// left/side, right/side, mid/side, narrow and interleave to the caller's container with coalesced stores.
// up: stream_decoder.c frame_sync_ / read_frame_header_ / read_subframe_* / read_residual_partitioned_rice_,
//     bitreader.c FLAC__bitreader_read_rice_signed_block, lpc.c FLAC__lpc_restore_signal[_wide],
//     fixed.c FLAC__fixed_restore_signal (ref: format.h:209-475, stream_decoder.h:1440-1513).
#include "fb_common.cuh"
#include "fb_math.cuh"
#include "dec_common.cuh"

namespace fb {

// ------------------------------------------------------------------ metadata walk (one thread per stream) ----
__global__ void dec_meta_kernel(const uint8_t* __restrict__ blob, const uint64_t* __restrict__ stream_off,
                                const uint64_t* __restrict__ stream_len, int n_streams, int raw_frames, DecStreamMeta raw_params,
                                DecStreamMeta* __restrict__ meta) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    DecStreamMeta m;
    if (raw_frames) { m = raw_params; m.first_frame = 0; m.status = kDecOk; meta[s] = m; return; }
    m.first_frame = 0; m.sample_rate = 0; m.channels = 0; m.bps = 0; m.min_blocksize = 0; m.max_blocksize = 0; m.total_samples = 0; m.status = kDecOk;
    m.have_last = 0; m.last_blocksize = 0; m.next_sample = 0;
    for (int i = 0; i < 16; i++) m.md5[i] = 0;
    const uint8_t* p = blob + stream_off[s];
    const uint64_t len = stream_len[s];
    if (len < 8 || p[0] != 'f' || p[1] != 'L' || p[2] != 'a' || p[3] != 'C') { m.status = kDecNotFlac; meta[s] = m; return; }
    uint64_t pos = 4; bool last = false, have_si = false;
    while (!last) {
        if (pos + 4 > len) { m.status = kDecBadMetadata; break; }
        const uint32_t type = p[pos] & 0x7f, blen = (uint32_t)p[pos + 1] << 16 | (uint32_t)p[pos + 2] << 8 | p[pos + 3];
        last = (p[pos] >> 7) != 0; pos += 4;
        if (pos + blen > len) { m.status = kDecBadMetadata; break; }
        if (type == 0 && blen >= 34) {   // STREAMINFO (ref: format.h:546-557)
            const uint8_t* q = p + pos;
            m.min_blocksize = (uint32_t)q[0] << 8 | q[1]; m.max_blocksize = (uint32_t)q[2] << 8 | q[3];
            m.sample_rate = (uint32_t)q[10] << 12 | (uint32_t)q[11] << 4 | (q[12] >> 4);
            m.channels = ((q[12] >> 1) & 7) + 1;
            m.bps = (((uint32_t)q[12] & 1) << 4 | (q[13] >> 4)) + 1;
            m.total_samples = ((uint64_t)(q[13] & 0xF) << 32) | (uint64_t)q[14] << 24 | (uint64_t)q[15] << 16 | (uint64_t)q[16] << 8 | q[17];
            for (int i = 0; i < 16; i++) m.md5[i] = q[18 + i];
            have_si = true;
        }
        pos += blen;
    }
    if (m.status == kDecOk && !have_si) m.status = kDecBadMetadata;
    m.first_frame = (uint32_t)pos;
    meta[s] = m;
}

// ------------------------------------------------------------------ frame-start candidates ----
// up: stream_decoder.c frame_sync_ + read_frame_header_.  Header layout: SURVEY Appendix B.
__device__ __forceinline__ bool parse_header(const uint8_t* __restrict__ p, uint64_t avail, const DecStreamMeta& m, DecCand& c) {
    if (avail < 5) return false;
    if (p[0] != 0xFF || (p[1] & 0xFE) != 0xF8) return false;            // sync 0x3ffe + reserved 0
    const uint32_t variable = p[1] & 1u;
    const uint32_t bs_code = p[2] >> 4, sr_code = p[2] & 0xF, ca_code = p[3] >> 4, bps_code = (p[3] >> 1) & 7;
    if (p[3] & 1) return false;
    if (bs_code == 0 || sr_code == 15) return false;
    if (ca_code > 10) return false;
    if (bps_code == 3) return false;
    // UTF-8 style coded number
    uint32_t n = 4; uint64_t number;
    {
        const uint32_t b0 = p[n++];
        uint32_t extra;
        if (!(b0 & 0x80)) { number = b0; extra = 0; }
        else if ((b0 & 0xE0) == 0xC0) { number = b0 & 0x1F; extra = 1; }
        else if ((b0 & 0xF0) == 0xE0) { number = b0 & 0x0F; extra = 2; }
        else if ((b0 & 0xF8) == 0xF0) { number = b0 & 0x07; extra = 3; }
        else if ((b0 & 0xFC) == 0xF8) { number = b0 & 0x03; extra = 4; }
        else if ((b0 & 0xFE) == 0xFC) { number = b0 & 0x01; extra = 5; }
        else if (b0 == 0xFE && variable) { number = 0; extra = 6; }
        else return false;
        if (n + extra + 1 > avail) return false;
        for (uint32_t i = 0; i < extra; i++) { const uint32_t b = p[n++]; if ((b & 0xC0) != 0x80) return false; number = (number << 6) | (b & 0x3F); }
    }
    uint32_t N;
    switch (bs_code) {
        case 1: N = 192; break;
        case 2: case 3: case 4: case 5: N = 576u << (bs_code - 2); break;
        case 6: if (n + 2 > avail) return false; N = (uint32_t)p[n] + 1; n += 1; break;
        case 7: if (n + 3 > avail) return false; N = ((uint32_t)p[n] << 8 | p[n + 1]) + 1; n += 2; break;
        default: N = 256u << (bs_code - 8); break;
    }
    uint32_t sr;
    switch (sr_code) {
        case 0: sr = m.sample_rate; break; case 1: sr = 88200; break; case 2: sr = 176400; break; case 3: sr = 192000; break;
        case 4: sr = 8000; break; case 5: sr = 16000; break; case 6: sr = 22050; break; case 7: sr = 24000; break;
        case 8: sr = 32000; break; case 9: sr = 44100; break; case 10: sr = 48000; break; case 11: sr = 96000; break;
        case 12: if (n + 2 > avail) return false; sr = (uint32_t)p[n] * 1000u; n += 1; break;
        case 13: if (n + 3 > avail) return false; sr = (uint32_t)p[n] << 8 | p[n + 1]; n += 2; break;
        default: if (n + 3 > avail) return false; sr = ((uint32_t)p[n] << 8 | p[n + 1]) * 10u; n += 2; break;
    }
    uint32_t bps;
    switch (bps_code) { case 0: bps = m.bps; break; case 1: bps = 8; break; case 2: bps = 12; break; case 4: bps = 16; break;
                        case 5: bps = 20; break; case 6: bps = 24; break; default: bps = 32; break; }
    if (bps == 0) return false;
    if (n + 1 > avail) return false;
    uint8_t crc = 0;
    for (uint32_t i = 0; i < n; i++) crc = crc8_byte(crc, p[i]);
    if (crc != p[n]) return false;
    c.blocksize = N; c.hdr_bytes = n + 1; c.sample_rate = sr; c.bps = (uint8_t)bps;
    c.channels = (uint8_t)(ca_code < 8 ? ca_code + 1 : 2); c.ca = (uint8_t)(ca_code < 8 ? 0 : ca_code - 7);
    c.variable = (uint8_t)variable; c.number = number;
    return true;
}

// One warp per 4096-byte segment; pass 0 counts candidates, pass 1 writes them in position order.
// The scan itself is an HBM stream: every lane loads 16 aligned bytes, a SIMD-in-register test marks positions that
// hold 0xFF followed by 0xF8 / 0xF9, and only those (about one per 30 KiB of random data plus the real frame starts)
// go through the header parser and its CRC-8.
__global__ void __launch_bounds__(128)
dec_sync_kernel(const uint8_t* __restrict__ blob, const uint64_t* __restrict__ stream_off, const uint64_t* __restrict__ stream_len,
                const DecStreamMeta* __restrict__ meta, const DecSegment* __restrict__ segs, int n_segs, int pass,
                uint32_t* __restrict__ seg_count, const uint32_t* __restrict__ seg_base, DecCand* __restrict__ cands) {
    const int seg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (seg >= n_segs) return;
    const DecSegment sg = segs[seg];
    const DecStreamMeta m = meta[sg.stream];
    const uint8_t* sbase = blob + stream_off[sg.stream];
    const uint64_t slen = stream_len[sg.stream];
    if (pass && seg_count[seg] == 0u) return;                                       // nothing to write for this segment
    uint32_t count = 0;
    const uint32_t out0 = pass ? seg_base[seg] : 0u;
    if (m.status == kDecOk) {
        const uint8_t* seg_lo = sbase + sg.start;                                   // first byte of the segment
        const uint8_t* a0 = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(seg_lo) & ~(uintptr_t)15);
        // 32-bit offsets relative to seg_lo keep the per-row bookkeeping short (the scan is instruction bound, not HBM bound)
        const int nbytes = (int)sg.bytes;                                           // segment = [0, nbytes)
        const uint64_t end64 = slen - sg.start;
        const int end_rel = end64 > 0x7fffffffull ? 0x7fffffff : (int)end64;        // stream end
        const uint64_t first64 = m.first_frame > sg.start ? m.first_frame - sg.start : 0ull;
        const int first_rel = first64 > 0x7fffffffull ? 0x7fffffff : (int)first64;
        const int row0 = (int)(a0 - seg_lo) + 16 * lane;                            // this lane's chunk in row 0 (-15 .. 496)
        // four 512-byte rows per trip, all loads issued before any is looked at
        for (int rb = 0; rb + row0 - 16 * lane < nbytes; rb += 2048) {
            uint4 v[4]; uint32_t nb31[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int rel = row0 + rb + 512 * u;
                // a chunk is needed when it overlaps the segment or supplies the byte after the segment's last one
                const bool need = rel + 16 > 0 && rel < nbytes + 16 && rel < end_rel;
                v[u] = need ? __ldg(reinterpret_cast<const uint4*>(seg_lo + rel)) : make_uint4(0u, 0u, 0u, 0u);
                nb31[u] = (lane == 31 && rel < nbytes && rel + 16 < end_rel) ? (uint32_t)__ldg(seg_lo + rel + 16) : 0u;   // first byte of the next row
            }
            uint32_t hits[4];                                                       // bit k: 0xFF then 0xF8/0xF9 at byte k of the chunk
#pragma unroll
            for (int u = 0; u < 4; u++) {
                uint32_t nb = __shfl_down_sync(0xffffffffu, v[u].x & 0xffu, 1);      // first byte of the next lane's chunk
                if (lane == 31) nb = nb31[u];
                const uint32_t w[5] = {v[u].x, v[u].y, v[u].z, v[u].w, nb};
                uint32_t h = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t nx = __funnelshift_r(w[q], w[q + 1], 8);                             // the following bytes
                    // a byte of c is zero exactly where this byte is 0xFF and the next one 0xF8 / 0xF9; exact zero-byte test
                    // (7-bit adds cannot carry across bytes), 0x80 in every matching byte
                    const uint32_t c = ~w[q] | ((nx & 0xFEFEFEFEu) ^ 0xF8F8F8F8u);
                    const uint32_t both = ~((((c & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | c) | 0x7F7F7F7Fu);
                    if (both) h |= (((both >> 7) & 1u) | ((both >> 14) & 2u) | ((both >> 21) & 4u) | ((both >> 28) & 8u)) << (4 * q);
                }
                hits[u] = h;
            }
            if (!__any_sync(0xffffffffu, (hits[0] | hits[1] | hits[2] | hits[3]) != 0u)) continue;
            // rare (about one row in 60 on random data, plus the real frame starts): rows, lanes and bits in position order
#pragma unroll 1
            for (int u = 0; u < 4; u++) {
                uint32_t hu = u == 0 ? hits[0] : (u == 1 ? hits[1] : (u == 2 ? hits[2] : hits[3]));
                const int rel = row0 + rb + 512 * u;
                // keep positions inside this segment, inside the stream's audio part, with a second byte present
                for (uint32_t t = hu; t; t &= t - 1) {
                    const int k = __ffs((int)t) - 1, cr = rel + k;
                    if (cr < 0 || cr >= nbytes || cr + 1 >= end_rel || cr < first_rel) hu &= ~(1u << k);
                }
                uint32_t lanes = __ballot_sync(0xffffffffu, hu != 0);
                while (lanes) {
                    const int src = __ffs((int)lanes) - 1;
                    lanes &= lanes - 1;
                    uint32_t found = 0;
                    if (lane == src) {
                        for (uint32_t t = hu; t; t &= t - 1) {
                            const uint8_t* c = seg_lo + rel + (__ffs((int)t) - 1);
                            const uint64_t pos = (uint64_t)(c - sbase);
                            DecCand cd;
                            bool hit = parse_header(c, slen - pos, m, cd);
                            // frames of one stream keep channels / bps (cheap plausibility filter; the chain decides anyway)
                            if (hit && (cd.channels != m.channels || cd.bps != m.bps) && m.channels) hit = false;
                            if (hit) {
                                if (pass) {
                                    cd.stream = sg.stream; cd.pos = (uint32_t)pos;
                                    cd.status = kDecPending; cd.end_pos = 0; cd.sample_slot = 0; cd.valid = 0; cd.sample_off = 0;
                                    cands[out0 + count + found] = cd;
                                }
                                found++;
                            }
                        }
                    }
                    count += __shfl_sync(0xffffffffu, found, src);
                }
            }
        }
    }
    if (!pass && lane == 0) seg_count[seg] = count;
}

// Generic single-CTA exclusive scan of uint32 -> uint32 / uint64 (n up to a few million; 64-bit sums throughout: the
// per-stream element counts can approach 2^32).  Tiles of 4096 elements: coalesced loads (the next tile is requested
// before the current one is scanned), per-thread prefix of 4, warp shuffles, one pass over the data.
__global__ void __launch_bounds__(1024)
dec_scan_u32_kernel(const uint32_t* __restrict__ in, int n, uint32_t* __restrict__ out32, uint64_t* __restrict__ out64, uint64_t* __restrict__ total) {
    typedef unsigned long long u64;
    __shared__ u64 warp_tot[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u64 carry = 0;
    auto load4 = [&](int base, uint32_t (&v)[4]) {
        const int i0 = base + 4 * tid;
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = (i0 + k < n) ? in[i0 + k] : 0u;
    };
    uint32_t cur[4], nxt[4];
    load4(0, cur);
    int par = 0;
    for (int base = 0; base < n; base += 4096, par ^= 1) {
        load4(base + 4096, nxt);
        const u64 s1 = (u64)cur[0] + cur[1], s2 = s1 + cur[2], s3 = s2 + cur[3];
        u64 inc = s3;                                                        // inclusive scan of the threads' sums within the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u64 t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) warp_tot[par][warp] = inc;
        __syncthreads();                                                     // (the other parity's totals are free again by now)
        const u64 wt = warp_tot[par][lane];
        u64 winc = wt;                                                       // every warp scans the 32 warp totals itself
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u64 t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
        const u64 warp_excl = __shfl_sync(0xffffffffu, winc - wt, warp);
        const u64 tile_total = __shfl_sync(0xffffffffu, winc, 31);
        const u64 e0 = carry + warp_excl + (inc - s3);
        const int i0 = base + 4 * tid;
        const u64 e[4] = {e0, e0 + cur[0], e0 + s1, e0 + s2};
#pragma unroll
        for (int k = 0; k < 4; k++) if (i0 + k < n) { if (out32) out32[i0 + k] = (uint32_t)e[k]; if (out64) out64[i0 + k] = e[k]; }
        carry += tile_total;
#pragma unroll
        for (int k = 0; k < 4; k++) cur[k] = nxt[k];
    }
    if (tid == 0 && total) *total = carry;
}

__global__ void dec_cand_size_kernel(const DecCand* __restrict__ cands, int n, uint32_t* __restrict__ sizes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    // 32-bit stereo: a side subframe may hold 33-bit samples; they are kept as (sample >> 1) plus a bitmap of the dropped bits
    // slots are padded so that every frame's planes start 32-byte aligned
    // (eight elements: dec_frame_kernel stores 32 bytes at a time where the blocksize allows)
    if (i < n) sizes[i] = (cands[i].blocksize * cands[i].channels + ((cands[i].bps == 32 && cands[i].channels == 2) ? (cands[i].blocksize + 31) / 32 : 0u) + 7u) & ~7u;
}

// ------------------------------------------------------------------ frame decode: one thread per candidate ----
// MSB-first bit reader.  The 32 lanes of a warp each decode their own frame, so anything a lane does "now and then" is
// done by SOME lane at practically every symbol and a branch around it buys nothing: the warp runs the body every time and
// pays for the divergence on top (the first version branched around a 27-39 instruction refill: 85 instructions per sample).
// The reader is therefore split by how often a step is due:
//  * fill_fast(), once per symbol, has no branch and no bookkeeping.  The only state is the bit position: with
//    j = word holding the next bit and r = its offset in that word, the window is  ahi = bits [pos, pos+32)  and
//    alo = word[j+1] << r.  After any consume of up to 32 bits ahi still holds every bit of words j' and (if the position
//    stayed in the same word) j'+1 that belong in it, so "OR in word[j'+1] >> (32 - r'), set alo = word[j'+1] << r'" is
//    right whether or not a word boundary was crossed (OR-ing bits that are already there changes nothing):
//    two address instructions, one shared load, a byte swap, two shifts and an OR;
//  * top_up_hot(), once per four symbols in the hot loop, keeps the ring ahead of the reader: a lane asks for the next
//    16-byte chunk with cp.async whenever fewer than kDecRing chunks from its current one are requested (a predicated copy, no
//    branch), and EVERY lane closes one cp.async group per iteration.  Groups belong to threads, but the hardware counts
//    them per warp: when only the lanes that had requested something closed a group and waited for "all but my newest",
//    the warp as a whole waited for nearly every lane's newest request -- a full memory latency per iteration (51 % of all
//    stall samples, profiles/r02d).  With one group per iteration for the whole warp, "all but the newest" means "everything
//    asked for in an earlier iteration", and that is enough: outside the generic path an iteration takes at most
//    2 x 10 + 4 x 32 = 148 bits, and a chunk that is needed now lay kDecRing chunk numbers (at least 3 x 128 + 1 bits) ahead
//    of the position at the top-up BEFORE the one that requested it -- more than two iterations' worth, so the request is at
//    least one iteration old (kDecHotGroups below; a ring of eight allows four groups in flight and measured slower).
//    Anything that can take more than 32 bits at a time (long unary runs, the scalar loops, subframe headers) goes through
//    top_up(), which requests what is missing and waits for all of it.
// The bytes travel global -> shared with cp.async and are read back a word at a time; loads into registers would not do:
// the lanes reach their chunk boundaries at different symbols but share one register scoreboard, so every lane's refill
// would wait for the load another lane issued an iteration earlier (measured: 62 % of all stall samples).
//
// What bounds the kernel once the instruction count is down (85 -> 48 per sample) is neither issue slots nor bytes but the
// number of memory REQUESTS an SM can send: every lane reads its own stream and writes its own plane, so a warp-wide store
// or cp.async is 32 separate requests, about one per four cycles per SM all told, and shared loads queue behind them
// (400-cycle LDS latencies in profiles/r02d).  Hence the 32-byte plane stores in the hot loop (STG.256: 4.57 -> 2.44 ms).
// Compile-time knobs below: the measured alternatives (profiles/r02d_ab_logs.txt).
constexpr int kDecFrameThreads = 64;
#ifndef FB_DEC_RING
#define FB_DEC_RING 4
#endif
#ifndef FB_DEC_CARVE
#define FB_DEC_CARVE 0
#endif
#ifndef FB_DEC_CG
#define FB_DEC_CG 0
#endif
#ifndef FB_DEC_ST256
#define FB_DEC_ST256 1
#endif
constexpr int kDecRing = FB_DEC_RING;   // 16-byte chunks per thread in the shared-memory ring (8: 128 contiguous bytes per thread)
// top_up_hot(): cp.async groups (= loop iterations) that may still be in flight.  A chunk that is needed now lay at least kDecRing
// chunk numbers ahead of the position at the top-up before the one that requested it, i.e. (kDecRing - 1) * 128 + 1 bits, and an
// iteration takes at most 148 bits outside the generic path: ring 8 -> requested >= 5 iterations ago, ring 4 -> >= 1.
constexpr int kDecHotGroups = kDecRing == 8 ? 4 : 1;
static_assert(kDecRing == 8 || kDecRing == 4, "ring depth");

struct BitReader {
    // Bank conflicts: a thread's ring is contiguous, so word w of lanes l and l + 2 (ring of four chunks) sits in the same bank,
    // and the lanes of a warp run within a few words of each other (same blocksize, similar bit rates).  Lane l therefore counts
    // its chunks from l mod kDecRing: positions, chunk numbers and the end mark below are all BIASED by that many chunks
    // (16 bytes = 128 bits each), c16 is moved back by as many chunks, and nothing in the hot path knows: the low five bits of
    // a bit position and the source address c16 + chunk are unchanged, while chunk c of lane l lands in ring slot (c + l) mod kDecRing.
    const uint4* c16;                   // 16-byte aligned address at or below the first byte read, minus the bias
    const uint4* safe;                  // some valid address for copies that are switched off (zero fill)
    uint32_t bq;                        // 32 + bit offset (from c16) of the next unread bit: bq >> 5 is the word fill_fast() merges
    uint32_t wendb;                     // byte offset of the first word past the stream (aligned up); 0 once a parse has given up
    uint32_t req;                       // next chunk to request
    uint32_t ahi, alo;                  // the 32 bits at the position; word[bq >> 5] << (bq & 31)
    uint32_t ring;                      // shared-memory address of this thread's ring (kDecRing * 16 bytes, aligned to its size)
    // chunk ci -> its ring slot when on is set (a predicated copy: no branch); chunks that start past the stream read as zero
    __device__ __forceinline__ void request(uint32_t ci, bool on) const {
        const uint32_t dst = ring | ((ci & (uint32_t)(kDecRing - 1)) << 4);
        const bool in = (ci << 4) < wendb;
#if FB_DEC_CG
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16, %2;\n\t}"
#else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p cp.async.ca.shared.global [%0], [%1], 16, %2;\n\t}"
#endif
                     :: "r"(dst), "l"(in ? c16 + ci : safe), "r"(in ? 16u : 0u), "r"((uint32_t)on) : "memory");   // size 0: zero fill
    }
    // generic: chunks cur .. cur + kDecRing - 1 requested and resident, cur = the chunk of the word the next fill merges
    __device__ __forceinline__ void top_up() {
        const uint32_t want = (bq >> 7) + (uint32_t)kDecRing;
        if (req < want) {
#pragma unroll 1
            do { request(req, true); req++; } while (req < want);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
    }
    // hot loop, once per iteration by every lane (see the header comment)
    __device__ __forceinline__ void top_up_hot() {
        const uint32_t want = (bq >> 7) + (uint32_t)kDecRing;
        const bool need = req < want;
        request(req, need);
        req += need ? 1u : 0u;
#pragma unroll 1
        while (req < want) { request(req, true); req++; }               // (an iteration that took more than 128 bits can cross two chunk boundaries)
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" :: "n"(kDecHotGroups) : "memory");
    }
    __device__ __forceinline__ uint32_t ring_word(uint32_t byte_off) const {
        uint32_t raw;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(raw) : "r"(ring | (byte_off & (uint32_t)(kDecRing * 16 - 4))));
        return __byte_perm(raw, 0, 0x0123);
    }
    __device__ __forceinline__ void init(const uint8_t* base, uint64_t start, uint64_t slen, uint32_t ring_addr, const void* safe_addr) {
        const uint8_t* p = base + start;
        const uintptr_t a = (uintptr_t)p, a16 = a & ~(uintptr_t)15;
        const uint32_t lb = threadIdx.x & (uint32_t)(kDecRing - 1);     // the bias, in chunks
        { const uint4* biased = reinterpret_cast<const uint4*>(a16) - lb; asm volatile("mov.u64 %0, %1;" : "=l"(c16) : "l"(biased)); }   // (kept as one value: see ring below)
        safe = reinterpret_cast<const uint4*>(safe_addr);
        const uint64_t end_rel = slen - start + (uint64_t)(a - a16);    // stream end relative to the unbiased c16 (bytes)
        wendb = ((uint32_t)(end_rel < 0xFFFFFF00ull ? end_rel + 3u : 0xFFFFFF00ull) & ~3u) + 16u * lb;   // (streams are shorter than 4 GiB)
        // (through a volatile asm: left to itself the compiler rebuilds the address from %tid at every use, five instructions per symbol)
        asm volatile("mov.u32 %0, %1;" : "=r"(ring) : "r"(ring_addr));
#pragma unroll
        for (int i = 0; i < kDecRing; i++) request((uint32_t)i + lb, true);
        req = (uint32_t)kDecRing + lb;
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        bq = (uint32_t)(a - a16) * 8u + 32u + 128u * lb;                // the first bit is a multiple of 8
        ahi = ring_word((bq - 32u) >> 3) << (bq & 31u);                  // the part of the first word that belongs to the stream
        alo = 0u;
        fill();
    }
    // Needs word bq >> 5 resident (top_up).  See the header comment: right with and without a word boundary behind us.
    __device__ __forceinline__ void fill_fast() {
        const uint32_t w = ring_word(bq >> 3);
        ahi |= __funnelshift_l(w, 0u, bq);                               // w >> (32 - r), 0 when r == 0
        alo = __funnelshift_l(0u, w, bq);                                // w << r
    }
    __device__ __forceinline__ void fill() { top_up(); fill_fast(); }
    __device__ __forceinline__ void consume(uint32_t len) {             // len <= 32, straight after a fill
        ahi = __funnelshift_lc(alo, ahi, len);
        alo = __funnelshift_lc(0u, alo, len);
        bq += len;
    }
    // bits consumed beyond the end of the stream (aligned up to a word)
    __device__ __forceinline__ bool overrun() const { return ((bq - 32u) >> 3) >= wendb + 4u; }
    __device__ __forceinline__ void give_up() { wendb = 0u; bq |= 0x100u; }     // overrun() from now on; later chunks read as zero
    __device__ __forceinline__ uint32_t get(uint32_t k) {          // k <= 32
        fill();
        const uint32_t v = __funnelshift_rc(ahi, 0u, 32u - k);      // ahi >> (32 - k), 0 when k == 0
        consume(k);
        return v;
    }
    __device__ __forceinline__ int32_t get_signed(uint32_t k) {
        if (k == 0) return 0;
        const uint32_t v = get(k);
        return (int32_t)(v << (32 - k)) >> (32 - k);
    }
    // the same two for the hot loop (k <= 32; covered by top_up_hot() like the fast-path Rice codes)
    __device__ __forceinline__ uint32_t get_hot(uint32_t k) {
        fill_fast();
        const uint32_t v = __funnelshift_rc(ahi, 0u, 32u - k);
        consume(k);
        return v;
    }
    __device__ __forceinline__ int32_t get_signed_hot(uint32_t k) {
        if (k == 0) return 0;
        const uint32_t v = get_hot(k);
        return (int32_t)(v << (32 - k)) >> (32 - k);
    }
    __device__ __forceinline__ uint32_t unary() {
        fill();
        uint32_t q = 0;
        while (ahi == 0u) {                                        // rare: more than 31 zeros in a row
            q += 32; consume(32);
            fill();
            if (overrun() || q > (1u << 26)) { give_up(); return q; }
        }
        const uint32_t z = (uint32_t)__clz((int)ahi);
        q += z;
        consume(z + 1u);
        return q;
    }
    // one Rice code with parameter k (< 31): when the unary zeros, the stop bit and the k low bits all lie in the 32 bits
    // of ahi this is one branch-free refill and one extraction; anything longer takes the generic path.
    // FAST: the caller has called top_up_hot() at most four fast-path symbols ago.
    template <bool FAST>
    __device__ __forceinline__ int32_t rice(uint32_t k) {
        if (FAST) fill_fast(); else fill();
        const uint32_t z = (uint32_t)__clz((int)ahi);              // 32 when ahi == 0
        const uint32_t len = z + 1u + k;
        uint32_t u;
        if (len <= 32u) {
            const uint32_t t = __funnelshift_l(alo, ahi, z + 1u);   // the bits after the stop bit (z + 1 <= 32 - k < 32 ... or k == 0)
            u = (z << k) | __funnelshift_rc(t, 0u, 32u - k);
            ahi = __funnelshift_lc(alo, ahi, len);                  // consume; alo is rebuilt by the next fill
            bq += len;
        } else {
            const uint32_t qv = unary();
            u = (qv << k) | get(k);
        }
        return (int32_t)(u >> 1) ^ -(int32_t)(u & 1u);
    }
    // bit offset (from the stream start) of the next unread bit; start = byte offset handed to init()
    __device__ __forceinline__ uint64_t bit_position(const uint8_t* base, uint64_t start) const {
        const int64_t off0 = (int64_t)start - (int64_t)((uintptr_t)(base + start) & 15u);
        return (uint64_t)(off0 * 8ll + (int64_t)bq - 32ll - 128ll * (long long)(threadIdx.x & (uint32_t)(kDecRing - 1)));
    }
};

#ifndef FB_DEC_MIN_CTAS
#define FB_DEC_MIN_CTAS 14     // 72 registers: the stash of the 32-byte stores stays in registers (16 CTAs / 64 registers spill it: 2.62 ms against 2.44 for 131 072 frames)
#endif

constexpr int kDecFastOrder = 12;      // predictor orders up to this keep history and coefficients in registers

// four samples of the prediction recurrence with TAPS taps (ref: FLAC__lpc_restore_signal / _wide, lpc.c; the 32-bit
// accumulate is exact whenever libFLAC picks it, so WIDE only has to be set when it would pick the 64-bit one)
template <int TAPS, bool WIDE>
__device__ __forceinline__ void restore4(int32_t (&h)[kDecFastOrder], const int32_t (&q)[kDecFastOrder], const int32_t (&r)[4], int shift, int4& v4) {
    int32_t v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        if (WIDE) {
            long long s = 0;
#pragma unroll
            for (int j = TAPS - 1; j >= 0; j--) s += (long long)q[j] * (long long)h[j];     // the newest sample's product last: one multiply-add on the sample-to-sample chain
            v[u] = (int32_t)((long long)r[u] + (s >> shift));
        } else {
            int s = 0;
#pragma unroll
            for (int j = TAPS - 1; j >= 0; j--) s += q[j] * h[j];
            v[u] = r[u] + (s >> shift);
        }
#pragma unroll
        for (int j = TAPS - 1; j > 0; j--) h[j] = h[j - 1];
        h[0] = v[u];
    }
    v4 = make_int4(v[0], v[1], v[2], v[3]);
}

__global__ void __launch_bounds__(kDecFrameThreads, FB_DEC_MIN_CTAS)
dec_frame_kernel(const uint8_t* __restrict__ blob, const uint64_t* __restrict__ stream_off, const uint64_t* __restrict__ stream_len,
                 DecCand* __restrict__ cands, int n_cands, const uint64_t* __restrict__ slot_off, int32_t* __restrict__ samples) {
    __shared__ __align__(128) uint4 ring_buf[kDecFrameThreads][kDecRing];
    const int ci = blockIdx.x * kDecFrameThreads + threadIdx.x;
    if (ci >= n_cands) return;
    DecCand c = cands[ci];
    const uint8_t* sbase = blob + stream_off[c.stream];
    const uint64_t slen = stream_len[c.stream];
    const uint64_t body = (uint64_t)c.pos + c.hdr_bytes;
    BitReader br; br.init(sbase, body, slen, (uint32_t)__cvta_generic_to_shared(&ring_buf[threadIdx.x][0]), cands);
    const uint32_t N = c.blocksize;
    int32_t* out = samples + slot_off[ci];                                  // 16-byte aligned (dec_cand_size_kernel pads the slots)
    int status = kDecOk;
    for (uint32_t chn = 0; chn < c.channels && status == kDecOk; chn++) {
        int32_t* o = out + (size_t)chn * N;
        const bool is_side = (c.ca == 1 && chn == 1) || (c.ca == 2 && chn == 0) || (c.ca == 3 && chn == 1);
        uint32_t bps = c.bps + (is_side ? 1u : 0u);
        // the side channel of a 32-bit stream spans 33 bits even when its subframe drops wasted bits: its plane always holds
        // sample >> 1 and the bitmap behind the planes the low bits (zero when there are wasted bits)
        const bool side33 = is_side && c.bps == 32;
        const uint32_t hdr = br.get(8);
        if (hdr & 0x80) { status = kDecBadFrame; break; }
        uint32_t wasted = 0;
        if (hdr & 1) { wasted = br.unary() + 1; if (br.overrun()) { status = kDecIncomplete; break; } if (wasted >= bps) { status = kDecBadFrame; break; } bps -= wasted; }
        const uint32_t wsh = (side33 && wasted) ? wasted - 1u : wasted;
        if (side33 && bps <= 32) { uint32_t* bm = reinterpret_cast<uint32_t*>(out + (size_t)c.channels * N); for (uint32_t w2 = 0; w2 < (N + 31u) / 32u; w2++) bm[w2] = 0u; }
        const uint32_t type = (hdr >> 1) & 0x3f;
        if (bps > 33) { status = kDecBadFrame; break; }
        if (type == 0 && bps <= 32) {                                           // CONSTANT
            const int32_t v = br.get_signed(bps) << wsh;
            for (uint32_t i = 0; i < N; i++) o[i] = v;
            if (br.overrun()) status = kDecIncomplete;
            continue;
        }
        if (type == 1 && bps <= 32) {                                           // VERBATIM
            for (uint32_t i = 0; i < N; i++) o[i] = br.get_signed(bps) << wsh;
            if (br.overrun()) status = kDecIncomplete;
            continue;
        }
        uint32_t order = 0; bool lpc = false;
        if (type >= 8 && type <= 12) order = type - 8;
        else if (type >= 32) { order = type - 31; lpc = true; }
        else if (type >= 2) { status = kDecUnparseable; break; }            // reserved subframe types
        if (order > N) { status = kDecBadFrame; break; }
        if (bps == 33 || order > (uint32_t)kDecFastOrder) {
            // ---- rare, generic code: the 33-bit side channel of a 32-bit stereo stream (the plane receives sample >> 1 and a
            // bitmap behind the planes the dropped low bits; dec_post_kernel puts them back) and predictor orders above
            // kDecFastOrder.  64-bit history and accumulate in local memory. ----
            const bool w33 = bps == 33;
            uint32_t* bitmap = reinterpret_cast<uint32_t*>(out + (size_t)c.channels * N);
            uint32_t word = 0;
            auto emit = [&](uint32_t i, long long v) {
                if (!w33) { o[i] = (int32_t)v << wsh; return; }
                o[i] = (int32_t)(v >> 1);
                word |= (uint32_t)(v & 1ll) << (i & 31u);
                if ((i & 31u) == 31u || i + 1 == N) { bitmap[i >> 5] = word; word = 0; }
            };
            auto getw = [&]() -> long long {
                if (!w33) return (long long)br.get_signed(bps);
                const uint32_t hi = br.get(1); const uint32_t lo = br.get(32); return (long long)lo - (hi ? 0x100000000ll : 0ll);
            };
            if (type == 0) { const long long v = getw(); for (uint32_t i = 0; i < N; i++) emit(i, v); }
            else if (type == 1) { for (uint32_t i = 0; i < N; i++) emit(i, getw()); }
            else {
                int shift = 0;
                long long hh[32]; int qq[32];
                for (uint32_t i = 0; i < order; i++) { const long long v = getw(); hh[i & 31u] = v; emit(i, v); }
                if (lpc) {
                    const uint32_t prec = br.get(4) + 1; if (prec == 16) { status = kDecBadFrame; break; }
                    shift = br.get_signed(5); if (shift < 0) { status = kDecBadFrame; break; }
                    for (uint32_t j = 0; j < order; j++) qq[j] = br.get_signed(prec);
                } else {
                    const int cf[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};
                    for (uint32_t j = 0; j < order; j++) qq[j] = cf[order][j];
                }
                const uint32_t method = br.get(2);
                if (method > 1) { status = kDecUnparseable; break; }
                const uint32_t po = br.get(4), plen = method ? 5u : 4u, pesc = method ? 31u : 15u;
                if ((N >> po) < order || (po > 0 && (N & ((1u << po) - 1)))) { status = kDecBadFrame; break; }
                uint32_t left = 0, k = 0, raw = 0;
                bool first = true;
                for (uint32_t i = order; i < N; i++) {
                    while (left == 0) { left = (N >> po) - (first ? order : 0u); first = false; k = br.get(plen); raw = (k == pesc) ? br.get(5) : 0u; }
                    left--;
                    const int32_t r = (k == pesc) ? br.get_signed(raw) : br.rice<false>(k);
                    long long sacc = 0;
                    for (uint32_t j = 0; j < order; j++) sacc += (long long)qq[j] * hh[(i - 1 - j) & 31u];
                    const long long v = (long long)r + (sacc >> shift);
                    hh[i & 31u] = v;
                    emit(i, v);
                    if (br.overrun()) break;
                }
            }
            if (br.overrun()) status = kDecIncomplete;
            continue;
        }
        // ---- FIXED / LPC up to kDecFastOrder taps: history (newest first: h[j] = sample i-1-j) and coefficients in registers ----
        int32_t h[kDecFastOrder], q[kDecFastOrder];
#pragma unroll
        for (int j = 0; j < kDecFastOrder; j++) { h[j] = 0; q[j] = 0; }
        for (uint32_t i = 0; i < order; i++) {
            const int32_t v = br.get_signed(bps);
            o[i] = v << wsh;
#pragma unroll
            for (int j = kDecFastOrder - 1; j > 0; j--) h[j] = h[j - 1];
            h[0] = v;
        }
        int shift = 0;
        uint32_t prec = 0;
        if (lpc) {
            prec = br.get(4) + 1; if (prec == 16) { status = kDecBadFrame; break; }
            shift = br.get_signed(5); if (shift < 0) { status = kDecBadFrame; break; }
            for (uint32_t j = 0; j < order; j++) {
                const int32_t cv = br.get_signed(prec);
#pragma unroll
                for (int jj = 0; jj < kDecFastOrder; jj++) if ((uint32_t)jj == j) q[jj] = cv;
            }
        } else {
            const int32_t cf[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};
#pragma unroll
            for (int jj = 0; jj < 4; jj++) q[jj] = cf[order][jj];
        }
        // ref: lpc.c FLAC__lpc_restore_signal vs _wide: 32-bit accumulate is exact iff bps + precision + ilog2(order) <= 32
        const bool wide = lpc ? (bps + prec + ilog2_u32(order ? order : 1) > 32) : (bps + order > 31);
        // residual: method, partition order, then per partition a parameter and its symbols
        const uint32_t method = br.get(2);
        if (method > 1) { status = kDecUnparseable; break; }                  // reserved residual coding methods
        const uint32_t po = br.get(4), plen = method ? 5u : 4u, pesc = method ? 31u : 15u;
        const uint32_t psize = N >> po;
        if (psize < order || (po > 0 && (N & ((1u << po) - 1)))) { status = kDecBadFrame; break; }
        uint32_t left = 0, k = 0, raw = 0;
        bool first = true;
        // Blocks and partitions whose lengths are multiples of four (every stream libFLAC writes apart from a short last
        // block) run four samples per iteration: history rotates by register renaming, one 16-byte store per four
        // samples, one partition test per four.  Scalar iterations lead up to the first multiple of four; everything
        // else is scalar throughout.  Lanes of a warp keep their own counters, so mixed orders stay converged.
        const bool quads = ((N | psize) & 3u) == 0u;
        // Stores: what bounds this kernel once the instruction count is down is the number of memory REQUESTS an SM can send
        // (every lane writes its own plane: 32 separate requests per warp store, about one per four cycles per SM), so two
        // quads leave as one 32-byte store (STG.256) where the blocksize is a multiple of eight.
        const bool octs = FB_DEC_ST256 && (N & 7u) == 0u;
        const uint32_t scalar_end = quads ? min(N, octs ? (order + 7u) & ~7u : (order + 3u) & ~3u) : N;
        uint32_t i = order;
        for (; i < scalar_end; i++) {
            while (left == 0) { left = psize - (first ? order : 0u); first = false; k = br.get(plen); raw = (k == pesc) ? br.get(5) : 0u; }
            left--;
            const int32_t r = (k == pesc) ? br.get_signed(raw) : br.rice<false>(k);
            long long s = 0;
#pragma unroll
            for (int j = 0; j < kDecFastOrder; j++) s += (long long)q[j] * (long long)h[j];
            const int32_t v = wide ? (int32_t)((long long)r + (s >> shift)) : r + ((int32_t)s >> shift);
#pragma unroll
            for (int j = kDecFastOrder - 1; j > 0; j--) h[j] = h[j - 1];
            h[0] = v;
            o[i] = v << wsh;
            if (br.overrun()) break;
        }
        if (quads && !br.overrun()) {
            // the lanes that arrive here together all take the widest class any of them needs (surplus taps multiply zero
            // coefficients): one body per warp and iteration instead of up to three run one after the other
            const int cls = (int)__reduce_max_sync(__activemask(), order <= 4u ? 0u : (order <= 8u ? 1u : 2u));
            int4 stash = make_int4(0, 0, 0, 0);
            for (; i < N; i += 4) {
                br.top_up_hot();                                                 // covers this iteration: at most 2 x 10 + 4 x 32 bits outside the generic path
                while (left == 0) { left = psize - (first ? order : 0u); first = false; k = br.get_hot(plen); raw = (k == pesc) ? br.get_hot(5) : 0u; }
                left -= 4;
                int32_t r[4];
                if (k == pesc) {
#pragma unroll
                    for (int u = 0; u < 4; u++) r[u] = br.get_signed_hot(raw);
                } else {
#pragma unroll
                    for (int u = 0; u < 4; u++) r[u] = br.rice<true>(k);
                }
                int4 v4;
                if (!wide) {
                    if (cls == 0) restore4<4, false>(h, q, r, shift, v4);
                    else if (cls == 1) restore4<8, false>(h, q, r, shift, v4);
                    else restore4<12, false>(h, q, r, shift, v4);
                } else {
                    if (cls == 0) restore4<4, true>(h, q, r, shift, v4);
                    else if (cls == 1) restore4<8, true>(h, q, r, shift, v4);
                    else restore4<12, true>(h, q, r, shift, v4);
                }
                v4.x <<= wsh; v4.y <<= wsh; v4.z <<= wsh; v4.w <<= wsh;
                if (!octs) *reinterpret_cast<int4*>(o + i) = v4;
                else if (i & 4u) asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(o + i - 4), "r"(stash.x), "r"(stash.y), "r"(stash.z), "r"(stash.w),
                                              "r"(v4.x), "r"(v4.y), "r"(v4.z), "r"(v4.w) : "memory");
                else stash = v4;
                if (br.wendb == 0u) break;                                       // a parse that ran off the end (zeros: the generic path gives up)
            }
        }
        if (br.overrun()) status = kDecIncomplete;
    }
    uint64_t endbit = br.bit_position(sbase, body);
    if (status == kDecOk) {
        // zero padding to the byte boundary, then CRC-16 (checked by the post kernel)
        const uint32_t padbits = (uint32_t)((8 - (endbit & 7)) & 7);
        if (padbits && br.get(padbits) != 0) status = kDecBadFrame;
        endbit += padbits;
        const uint64_t endbyte = endbit >> 3;
        if (endbyte + 2 > slen) status = kDecIncomplete;
        c.end_pos = (uint32_t)(endbyte + 2);
    }
    if (status != kDecOk) {                         // where the parse stopped: the resynchronisation scan goes on from here
        const uint64_t stop = br.wendb == 0u ? slen : (br.bit_position(sbase, body) + 7ull) >> 3;      // (a parse that gave up: the whole rest)
        c.end_pos = (uint32_t)(stop < slen ? stop : slen);
    }
    cands[ci].status = status;
    cands[ci].end_pos = c.end_pos;
}

// ------------------------------------------------------------------ CRC-16 of every decodable candidate: one warp each ----
// (ref: format.h:467-475.)  Runs before the chain walk so that the walk can treat a frame with a bad checksum exactly as
// libFLAC does: report it, drop it and search on from just behind its sync code.
#ifndef FB_DEC_CRC_STAGED
#define FB_DEC_CRC_STAGED 1
#endif
#if FB_DEC_CRC_STAGED
// The frame is cut into 64-byte chunks counted from its END (so every chunk but the first is whole and the positional
// weights are x^(512 j)); lane l of tile t takes chunk 32 t + l.  A tile's 2 KiB travel global -> shared as whole 16-byte
// pieces, coalesced (four per lane: 64 sectors per tile; a lane fetching its own chunk word by word touched 1024), and
// each lane then reads its chunk back from shared memory a word at a time.  Lanes are 64 bytes apart, so the tile is kept
// skewed by one word per 64 bytes (byte b at word (b >> 2) + (b >> 6)): lane l's k-th word falls into bank 17 l + const.
constexpr int kCrcTileBytes = 2048;
__global__ void __launch_bounds__(128)
dec_crc_kernel(const uint8_t* __restrict__ blob, const uint64_t* __restrict__ stream_off, DecCand* __restrict__ cands, int n_cands) {
    __shared__ __align__(16) uint16_t tabs[4][256];
    __shared__ uint32_t tile_buf[4][(kCrcTileBytes + 32) / 4 + (kCrcTileBytes + 32) / 64 + 1];
    for (int i = threadIdx.x; i < 256; i += 128) reinterpret_cast<uint2*>(&tabs[0][0])[i] = reinterpret_cast<const uint2*>(&g_crc16_slice.t[0][0])[i];
    __syncthreads();
    const int ci = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (ci >= n_cands) return;
    const DecCand c = cands[ci];
    if (c.status != kDecOk) return;
    const uint8_t* fb = blob + stream_off[c.stream] + c.pos;
    const uint32_t nb = c.end_pos - c.pos - 2u;
    uint32_t* tile = tile_buf[threadIdx.x >> 5];
    auto tword = [&](uint32_t b) -> uint32_t& { return tile[(b >> 2) + (b >> 6)]; };           // b = byte offset in the tile, multiple of 4
    const uint32_t nchunks = (nb + 63u) >> 6;
    uint32_t acc = 0;
    for (uint32_t t0 = 0; t0 < nchunks; t0 += 32) {
        // tile = frame bytes [lo, hi), hi = nb - 64 t0; staged from the 16-byte boundary at or below fb + lo
        const uint32_t hi = nb - (t0 << 6), lo = hi > (uint32_t)kCrcTileBytes ? hi - (uint32_t)kCrcTileBytes : 0u;
        const uint8_t* g0 = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(fb + lo) & ~(uintptr_t)15);
        const uint32_t shift = (uint32_t)((fb + lo) - g0);                      // tile byte b of the frame sits at tile[shift + b - lo]
        const uint32_t pieces = (shift + (hi - lo) + 15u) >> 4;                 // <= 129
        __syncwarp();
        for (uint32_t pi = (uint32_t)lane; pi < pieces; pi += 32) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(g0) + pi);
            uint32_t* d = &tword(pi << 4);                                      // a 16-byte piece never straddles a 64-byte block
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
        __syncwarp();
        const uint32_t j = t0 + (uint32_t)lane;
        if (j < nchunks) {
            const uint32_t end = nb - (j << 6);                                  // chunk = frame bytes [end - 64, end) or [0, end)
            uint32_t crc = 0;
            if (end >= 64u) {
                const uint32_t o = shift + (end - 64u - lo), sh8 = (o & 3u) * 8u;
                const uint32_t ob = o & ~3u;
                uint32_t prev = tword(ob);
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const uint32_t nxt = tword(ob + 4u * (uint32_t)(i + 1));
                    const uint32_t w = __funnelshift_r(prev, nxt, sh8);
                    prev = nxt;
                    crc = (uint32_t)tabs[3][(crc >> 8) ^ (w & 0xffu)] ^ tabs[2][(crc & 0xffu) ^ ((w >> 8) & 0xffu)] ^ tabs[1][(w >> 16) & 0xffu] ^ tabs[0][w >> 24];
                }
            } else {
                for (uint32_t b = 0; b < end; b++) {
                    const uint32_t tb = shift + b - lo;
                    crc = ((crc << 8) & 0xffffu) ^ tabs[0][(crc >> 8) ^ ((tword(tb & ~3u) >> ((tb & 3u) * 8u)) & 0xffu)];
                }
            }
            acc ^= crc16_weigh_chunk((uint16_t)crc, j, g_crc_pos);          // (any frame size: a decoder can meet frames of 256 KiB and more)
        }
    }
    acc = __reduce_xor_sync(0xffffffffu, acc);
    if (lane == 0) {
        const uint16_t stored = (uint16_t)((uint16_t)fb[nb] << 8 | fb[nb + 1]);
        if ((uint16_t)acc != stored) cands[ci].status = kDecCrcMismatch;
    }
}
#else
__global__ void __launch_bounds__(128)
dec_crc_kernel(const uint8_t* __restrict__ blob, const uint64_t* __restrict__ stream_off, DecCand* __restrict__ cands, int n_cands) {
    __shared__ __align__(16) uint16_t tabs[4][256];
    for (int i = threadIdx.x; i < 256; i += 128) reinterpret_cast<uint2*>(&tabs[0][0])[i] = reinterpret_cast<const uint2*>(&g_crc16_slice.t[0][0])[i];
    __syncthreads();
    const int ci = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (ci >= n_cands) return;
    const DecCand c = cands[ci];
    if (c.status != kDecOk) return;
    const uint8_t* fb = blob + stream_off[c.stream] + c.pos;
    const uint32_t nb = c.end_pos - c.pos - 2u;
    const uint32_t mis = (uint32_t)((uintptr_t)fb & 3u);
    const uint32_t* fw = reinterpret_cast<const uint32_t*>(fb - mis);       // aligned words; byte j of the frame is byte j + mis of fw
    auto word_at = [&](uint32_t j) { const uint32_t o = j + mis; return __funnelshift_r(__ldg(fw + (o >> 2)), __ldg(fw + (o >> 2) + 1), (o & 3u) * 8u); };
    const uint32_t nchunks = (nb + 63u) >> 6;
    uint32_t acc = 0;
    for (uint32_t j = (uint32_t)lane; j < nchunks; j += 32) {               // 64-byte chunks counted from the end of the frame
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
        acc ^= crc16_weigh_chunk((uint16_t)crc, j, g_crc_pos);
    }
    acc = __reduce_xor_sync(0xffffffffu, acc);
    if (lane == 0) {
        const uint16_t stored = (uint16_t)((uint16_t)fb[nb] << 8 | fb[nb + 1]);
        if ((uint16_t)acc != stored) cands[ci].status = kDecCrcMismatch;
    }
}
#endif

// ------------------------------------------------------------------ chain walk: one thread per stream ----
// The sequential part of libFLAC's stream_decoder.c (frame_sync_ / read_frame_) over the candidate table, incl. what
// 1.4.3 does when something is wrong (pinned against the reference binary, tests/test_gpu_decode.py::test_decode_error_
// recovery_matches_libflac):
//  * bytes that are not a frame where one should start: LOST_SYNC once, the search goes on to the next candidate;
//  * a frame whose CRC-16 does not match: FRAME_CRC_MISMATCH, the frame is dropped and the search resumes two bytes
//    behind its sync code (so LOST_SYNC follows while the scan runs through the frame's body);
//  * a frame that cannot be parsed: UNPARSEABLE_STREAM (reserved values) or LOST_SYNC, the search resumes where the parse stopped;
//  * a good frame whose number says that frames are missing in front of it (and something was decoded before): the gap is
//    filled with silence in units of the previous frame's blocksize; missing frames at the very start or end are not filled.
// Errors behind the last good frame are reported only when no more input can follow (eof): a streaming caller sees them
// once, when the next good frame or the end of the stream closes the gap.
__global__ void dec_chain_kernel(DecCand* __restrict__ cands, const uint32_t* __restrict__ cand_first, const DecStreamMeta* __restrict__ meta,
                                 const uint64_t* __restrict__ stream_len, int n_streams, int eof, DecStreamResult* __restrict__ res,
                                 uint32_t* __restrict__ stream_samples32, unsigned int* __restrict__ any_gap) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    DecStreamResult r;
    r.total_samples = 0; r.n_frames = 0; r.status = meta[s].status; r.consumed = meta[s].first_frame; r.max_blocksize = 0; r.pcm_off = 0;
    r.sample_rate = meta[s].sample_rate; r.channels = meta[s].channels; r.bps = meta[s].bps; r.n_events = 0; r.gap_samples = 0;
    r.next_sample = meta[s].next_sample; r.last_blocksize = meta[s].last_blocksize; r.have_last = meta[s].have_last;
    for (int k = 0; k < kDecMaxEvents; k++) { r.ev_frame[k] = 0; r.ev_status[k] = 0; }
    if (r.status == kDecOk) {
        uint64_t expect = meta[s].first_frame;
        uint32_t i = cand_first[s];
        const uint32_t iend = cand_first[s + 1];
        const uint64_t slen = stream_len[s];
        // events since the last good frame (committed when the next good frame or the end of input closes the gap)
        uint8_t pend[kDecMaxEvents]; uint32_t npend = 0; int pend_first = kDecOk;
        auto push = [&](int flac_status, int own_status) { if (npend < (uint32_t)kDecMaxEvents) pend[npend] = (uint8_t)flac_status; npend++; if (pend_first == kDecOk) pend_first = own_status; };
        auto commit = [&]() {
            for (uint32_t k = 0; k < npend; k++) { if (r.n_events < (uint32_t)kDecMaxEvents && k < (uint32_t)kDecMaxEvents) { r.ev_frame[r.n_events] = r.n_frames; r.ev_status[r.n_events] = pend[k]; } r.n_events++; }
            if (r.status == kDecOk) r.status = pend_first;
            npend = 0; pend_first = kDecOk;
        };
        bool have_last = meta[s].have_last != 0, in_gap = false, incomplete_tail = false;
        uint64_t next_sample = meta[s].next_sample;      // first sample behind the last delivered frame, by the frame headers' numbering
        uint32_t last_bs = meta[s].last_blocksize, fixed_bs = meta[s].max_blocksize && meta[s].min_blocksize == meta[s].max_blocksize ? meta[s].max_blocksize : 0u;
        while (expect < slen) {
            while (i < iend && cands[i].pos < expect) i++;
            if (i >= iend) break;
            if (cands[i].pos != expect && !in_gap) { push(0, kDecLostSync); in_gap = true; }             // LOST_SYNC, once per search
            const DecCand c = cands[i];
            if (c.status == kDecOk) {
                if (r.n_frames == 0) { r.sample_rate = c.sample_rate; r.channels = c.channels; r.bps = c.bps; }
                if (fixed_bs == 0 && !c.variable) fixed_bs = c.blocksize;                                   // up: fixed_block_size falls back to the first frame's
                const uint64_t start = c.variable ? c.number : c.number * (uint64_t)fixed_bs;
                uint64_t gap = (have_last && start > next_sample) ? start - next_sample : 0ull;
                if (gap > (1ull << 28)) gap = 0;                                                           // a corrupt number, not a hole in the stream
                const uint64_t chn = (uint64_t)(r.channels ? r.channels : 1);
                if ((r.total_samples + gap + c.blocksize) * chn > 0xFFFFFFF0ull) { if (r.status == kDecOk) r.status = kDecUnsupported; break; }
                commit();
                in_gap = false;
                r.total_samples += gap; r.gap_samples += (uint32_t)gap;
                cands[i].valid = 1; cands[i].sample_off = r.total_samples;
                r.total_samples += c.blocksize; r.n_frames++;
                if (c.blocksize > r.max_blocksize) r.max_blocksize = c.blocksize;
                have_last = true; next_sample = start + c.blocksize; last_bs = c.blocksize;
                expect = c.end_pos;
                r.consumed = expect;
                i++;
            } else if (c.status == kDecIncomplete) {
                incomplete_tail = true;                                     // the frame runs past the input: wait for more, or lost at the end
                break;
            } else {
                if (c.status == kDecCrcMismatch) { push(2, kDecCrcMismatch); expect = (uint64_t)c.pos + 2; }
                else { push(c.status == kDecUnparseable ? 3 : 0, c.status); expect = c.end_pos > c.pos + 2 ? c.end_pos : (uint64_t)c.pos + 2; }
                in_gap = false;                                             // the next search reports its own LOST_SYNC
                i++;
            }
        }
        r.next_sample = next_sample; r.last_blocksize = last_bs; r.have_last = have_last ? 1u : 0u;
        if (eof) {
            // nothing follows: what is left behind the last good frame is lost
            if (r.consumed < slen && !in_gap && (incomplete_tail || expect < slen || npend == 0)) push(0, incomplete_tail ? kDecIncomplete : kDecLostSync);
            if (r.consumed < slen || npend) commit();
            if (r.consumed < slen && r.status == kDecOk) r.status = kDecLostSync;
        } else if (r.status == kDecOk && r.consumed < slen) {
            r.status = incomplete_tail ? kDecIncomplete : kDecLostSync;    // informational for streaming callers: bytes are still pending
        }
    }
    if (r.gap_samples) atomicOr(any_gap, 1u);
    res[s] = r;
    stream_samples32[s] = (uint32_t)(r.total_samples * (r.channels ? r.channels : 1));   // elements; < 2^32 (checked above)
}

__global__ void dec_assign_kernel(DecStreamResult* __restrict__ res, const uint64_t* __restrict__ pcm_off, int n_streams) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_streams) res[s].pcm_off = pcm_off[s];
}

// ------------------------------------------------------------------ post: one CTA per chained frame ----
// (CRC-16 is checked per candidate by dec_crc_kernel before the chain walk.)  Undo channel decorrelation
// (ref: format.h:388-393), narrow to the caller's container, interleave [sample][channel], coalesced stores.
// undo one inter-channel decorrelation (ref: format.h:388-393); a = plane 0, b = plane 1
__device__ __forceinline__ void undo_stereo(int ca, int32_t a, int32_t b, int32_t& l, int32_t& r) {
    if (ca == 1) { l = a; r = a - b; }
    else if (ca == 2) { l = a + b; r = b; }
    else { const int32_t m2 = (int32_t)(((uint32_t)a << 1) | ((uint32_t)b & 1u)); l = (m2 + b) >> 1; r = (m2 - b) >> 1; }
}

template <typename OutT>
__global__ void __launch_bounds__(256, 4)
dec_post_kernel(const uint8_t* __restrict__ blob, const uint64_t* __restrict__ stream_off, DecCand* __restrict__ cands,
                const uint64_t* __restrict__ slot_off, const int32_t* __restrict__ samples, DecStreamResult* __restrict__ res,
                OutT* __restrict__ pcm_out) {
    const int tid = threadIdx.x;
    const DecCand c = cands[blockIdx.x];
    if (!c.valid) return;
    const uint32_t N = c.blocksize, ch = c.channels;
    const int32_t* in = samples + slot_off[blockIdx.x];                     // 16-byte aligned (padded slots)
    OutT* out = pcm_out + res[c.stream].pcm_off + c.sample_off * ch;
    // Stereo frames whose planes can be read four samples at a time: the first 1024 sample quads of both planes are requested up front.
    const bool side33 = c.bps == 32;                 // side plane = side >> 1, low bits in the bitmap behind the planes (dec_frame_kernel)
    const bool quads = c.ca != 0 && !side33 && (N & 3u) == 0u;
    const uint32_t nq = N >> 2;
    const int4* pa = reinterpret_cast<const int4*>(in);
    const int4* pb = reinterpret_cast<const int4*>(in + N);
    int4 A[4], B[4];
    if (quads) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t qi = (uint32_t)tid + 256u * (uint32_t)u;
            if (qi < nq) { A[u] = __ldcs(pa + qi); B[u] = __ldcs(pb + qi); }
        }
    }
    if (c.ca == 0) {
        for (uint32_t e = tid; e < N * ch; e += 256) { const uint32_t i = e / ch, k = e - i * ch; out[e] = (OutT)in[(size_t)k * N + i]; }
    } else if (quads) {
        const bool al16 = ((uintptr_t)out & 15u) == 0u, alpair = ((uintptr_t)out & (2u * sizeof(OutT) - 1u)) == 0u;
        for (uint32_t q0 = 0; q0 < nq; q0 += 1024u) {
            if (q0) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t qi = q0 + (uint32_t)tid + 256u * (uint32_t)u;
                    if (qi < nq) { A[u] = __ldcs(pa + qi); B[u] = __ldcs(pb + qi); }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t qi = q0 + (uint32_t)tid + 256u * (uint32_t)u;
                if (qi >= nq) continue;
                int32_t l[4], r[4];
                undo_stereo(c.ca, A[u].x, B[u].x, l[0], r[0]); undo_stereo(c.ca, A[u].y, B[u].y, l[1], r[1]);
                undo_stereo(c.ca, A[u].z, B[u].z, l[2], r[2]); undo_stereo(c.ca, A[u].w, B[u].w, l[3], r[3]);
                OutT* po = out + (size_t)qi * 8u;
                if (sizeof(OutT) == 2) {
                    uint32_t w[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) w[k] = ((uint32_t)l[k] & 0xffffu) | ((uint32_t)r[k] << 16);
                    if (al16) __stcs(reinterpret_cast<uint4*>(po), make_uint4(w[0], w[1], w[2], w[3]));
                    else if (alpair) {
#pragma unroll
                        for (int k = 0; k < 4; k++) reinterpret_cast<uint32_t*>(po)[k] = w[k];
                    } else {                                                   // a batch that mixes channel counts can leave a stream 2-byte aligned
#pragma unroll
                        for (int k = 0; k < 4; k++) { po[2 * k] = (OutT)l[k]; po[2 * k + 1] = (OutT)r[k]; }
                    }
                } else {
                    if (al16) {
                        __stcs(reinterpret_cast<int4*>(po), make_int4(l[0], r[0], l[1], r[1]));
                        __stcs(reinterpret_cast<int4*>(po) + 1, make_int4(l[2], r[2], l[3], r[3]));
                    } else if (alpair) {
#pragma unroll
                        for (int k = 0; k < 4; k++) reinterpret_cast<int2*>(po)[k] = make_int2(l[k], r[k]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; k++) { po[2 * k] = (OutT)l[k]; po[2 * k + 1] = (OutT)r[k]; }
                    }
                }
            }
        }
    } else {
        const uint32_t* bitmap = reinterpret_cast<const uint32_t*>(in + (size_t)2 * N);
        for (uint32_t i = tid; i < N; i += 256) {
            const int32_t a = in[i], b = in[N + i];
            int32_t l, r;
            if (side33) {
                // modulo-2^32 arithmetic is exact because the results fit 32 bits.  side = 2*sh + lsb:
                //   left/side: R = L - side;  side/right: L = side + R;  mid/side: L = mid + sh + lsb, R = mid - sh
                const uint32_t lsb = (bitmap[i >> 5] >> (i & 31u)) & 1u;
                if (c.ca == 1) { l = a; r = (int32_t)((uint32_t)a - (((uint32_t)b << 1) | lsb)); }
                else if (c.ca == 2) { r = b; l = (int32_t)((((uint32_t)a << 1) | lsb) + (uint32_t)b); }
                else { l = (int32_t)((uint32_t)a + (uint32_t)b + lsb); r = (int32_t)((uint32_t)a - (uint32_t)b); }
            } else undo_stereo(c.ca, a, b, l, r);
            out[2 * i] = (OutT)l; out[2 * i + 1] = (OutT)r;
        }
    }
}

// first candidate of every stream = scan value at the stream's first segment (sentinel: all candidates)
__global__ void dec_cand_first_kernel(const uint32_t* __restrict__ seg_base, const uint32_t* __restrict__ first_seg, int n_streams, int n_segs,
                                      const uint64_t* __restrict__ total, uint32_t* __restrict__ cand_first) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_streams) return;
    const uint32_t fs = first_seg[s];
    cand_first[s] = (s < n_streams && (int)fs < n_segs) ? seg_base[fs] : (uint32_t)*total;
}
// copy a few words to mapped host memory (readbacks that must not queue behind bulk D2H copies on the copy engine)
__global__ void dec_mirror_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int n_words) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_words) dst[i] = src[i];
}

// ------------------------------------------------------------------ launchers ----
void launch_dec_meta(const uint8_t* blob, const uint64_t* soff, const uint64_t* slen, int ns, int raw, const DecStreamMeta& rawp, DecStreamMeta* meta, cudaStream_t st) {
    dec_meta_kernel<<<(ns + 127) / 128, 128, 0, st>>>(blob, soff, slen, ns, raw, rawp, meta);
}
void launch_dec_sync(const uint8_t* blob, const uint64_t* soff, const uint64_t* slen, const DecStreamMeta* meta, const DecSegment* segs, int nsegs, int pass,
                     uint32_t* seg_count, const uint32_t* seg_base, DecCand* cands, cudaStream_t st) {
    dec_sync_kernel<<<(nsegs + 3) / 4, 128, 0, st>>>(blob, soff, slen, meta, segs, nsegs, pass, seg_count, seg_base, cands);
}
void launch_dec_scan(const uint32_t* in, int n, uint32_t* out32, uint64_t* out64, uint64_t* total, cudaStream_t st) {
    dec_scan_u32_kernel<<<1, 1024, 0, st>>>(in, n, out32, out64, total);
}
void launch_dec_cand_size(const DecCand* cands, int n, uint32_t* sizes, cudaStream_t st) {
    if (n) dec_cand_size_kernel<<<(n + 255) / 256, 256, 0, st>>>(cands, n, sizes);
}
void launch_dec_frames(const uint8_t* blob, const uint64_t* soff, const uint64_t* slen, DecCand* cands, int n, const uint64_t* slot_off, int32_t* samples, cudaStream_t st) {
    if (!n) return;
    // sixteen CTAs per SM need 16 x (8 KiB of rings + 1 KiB reserved): ask for the large shared-memory carve-out once per device
    static unsigned long long carved = 0ull;
    int dev = 0; cudaGetDevice(&dev);
    if (FB_DEC_CARVE && dev >= 0 && dev < 64 && !((carved >> dev) & 1ull)) {
        cudaFuncSetAttribute(dec_frame_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        carved |= 1ull << dev;
    }
    dec_frame_kernel<<<(n + kDecFrameThreads - 1) / kDecFrameThreads, kDecFrameThreads, 0, st>>>(blob, soff, slen, cands, n, slot_off, samples);
}
void launch_dec_chain(DecCand* cands, const uint32_t* cand_first, const DecStreamMeta* meta, const uint64_t* slen, int ns, int eof, DecStreamResult* res,
                      uint32_t* ss32, unsigned int* any_gap, cudaStream_t st) {
    dec_chain_kernel<<<(ns + 127) / 128, 128, 0, st>>>(cands, cand_first, meta, slen, ns, eof, res, ss32, any_gap);
}
void launch_dec_crc(const uint8_t* blob, const uint64_t* soff, DecCand* cands, int n, cudaStream_t st) {
    if (n) dec_crc_kernel<<<(n + 3) / 4, 128, 0, st>>>(blob, soff, cands, n);
}
void launch_dec_assign(DecStreamResult* res, const uint64_t* pcm_off, int ns, cudaStream_t st) {
    dec_assign_kernel<<<(ns + 127) / 128, 128, 0, st>>>(res, pcm_off, ns);
}
void launch_dec_cand_first(const uint32_t* seg_base, const uint32_t* first_seg, int ns, int nsegs, const uint64_t* total, uint32_t* cand_first, cudaStream_t st) {
    dec_cand_first_kernel<<<(ns + 1 + 127) / 128, 128, 0, st>>>(seg_base, first_seg, ns, nsegs, total, cand_first);
}
void launch_dec_mirror(const void* src, void* dst, size_t bytes, cudaStream_t st) {
    const int n = (int)((bytes + 3) / 4);
    if (n) dec_mirror_kernel<<<(n + 255) / 256, 256, 0, st>>>((const uint32_t*)src, (uint32_t*)dst, n);
}
void launch_dec_post(const uint8_t* blob, const uint64_t* soff, DecCand* cands, int n, const uint64_t* slot_off, const int32_t* samples, DecStreamResult* res,
                     void* pcm_out, int out_bytes, cudaStream_t st) {
    if (!n) return;
    if (out_bytes == 2) dec_post_kernel<int16_t><<<n, 256, 0, st>>>(blob, soff, cands, slot_off, samples, res, (int16_t*)pcm_out);
    else dec_post_kernel<int32_t><<<n, 256, 0, st>>>(blob, soff, cands, slot_off, samples, res, (int32_t*)pcm_out);
}

}  // namespace fb
