// dec_engine.cu -- host side of the batch decoder (include/flacb200.h "batch decode").
//
//   [H2D blob if host] -> meta -> sync(count) -> scan -> sync(write) -> size+scan -> frames -> chain -> scan -> post
// Two host synchronisations per batch (candidate count, decoded PCM size) because device buffers are sized from them.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <functional>
#include <vector>

#include "../../include/flacb200.h"
#include "fb_common.cuh"
#include "dec_common.cuh"

struct flacb200_ctx;
int fb_ctx_device(flacb200_ctx*);
cudaStream_t fb_ctx_stream(flacb200_ctx*);
void** fb_ctx_dec_slot(flacb200_ctx*, void (*)(void*));
int fb_ctx_fail(flacb200_ctx*, int, const char*, cudaError_t);
void fb_ctx_add_launches(flacb200_ctx*, uint64_t);

namespace fb {
void launch_dec_meta(const uint8_t*, const uint64_t*, const uint64_t*, int, int, const DecStreamMeta&, DecStreamMeta*, cudaStream_t);
void launch_dec_sync(const uint8_t*, const uint64_t*, const uint64_t*, const DecStreamMeta*, const DecSegment*, int, int, uint32_t*, const uint32_t*, DecCand*, cudaStream_t);
void launch_dec_scan(const uint32_t*, int, uint32_t*, uint64_t*, uint64_t*, cudaStream_t);
void launch_dec_cand_size(const DecCand*, int, uint32_t*, cudaStream_t);
void launch_dec_frames(const uint8_t*, const uint64_t*, const uint64_t*, DecCand*, int, const uint64_t*, int32_t*, cudaStream_t);
void launch_dec_chain(DecCand*, const uint32_t*, const DecStreamMeta*, const uint64_t*, int, int, DecStreamResult*, uint32_t*, unsigned int*, cudaStream_t);
void launch_dec_crc(const uint8_t*, const uint64_t*, DecCand*, int, cudaStream_t);
void launch_dec_assign(DecStreamResult*, const uint64_t*, int, cudaStream_t);
void launch_dec_post(const uint8_t*, const uint64_t*, DecCand*, int, const uint64_t*, const int32_t*, DecStreamResult*, void*, int, cudaStream_t);
void launch_dec_cand_first(const uint32_t*, const uint32_t*, int, int, const uint64_t*, uint32_t*, cudaStream_t);
void launch_dec_mirror(const void*, void*, size_t, cudaStream_t);
}
using namespace fb;

namespace {

struct Buf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = n + n / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct DecState {
    Buf blob, soff, slen, meta, segs, segcount, segbase, cands, sizes, slotoff, samples, candfirst, res, ss32, pcmoff, pcm, total, firstseg;
    // readbacks go through mapped pinned memory written by tiny kernels: a D2H memcpy would queue on the copy engine behind
    // the bulk PCM transfer of the previous chunk (flacb200_decode_batch_host)
    uint64_t* m_total = nullptr;            // [0] scan total
    unsigned char* m_res = nullptr; size_t m_res_cap = 0;
    std::vector<DecSegment> h_segs;
    std::vector<uint32_t> h_first_seg;      // first segment of each stream (+1 sentinel)
    std::vector<uint64_t> c_off, c_len;     // layout the segment table on the device was built for (a repeated layout skips the rebuild + upload)
    int n_streams = 0, n_cands = 0;
    uint64_t total_elems = 0;
    uint32_t out_bytes = 0;
    bool have = false;
    cudaEvent_t ev[6] = {nullptr};      // [5]: between the frame decode and the CRC kernel
    std::vector<DecStreamResult> h_res;
    DecState* alt = nullptr;                // second state for flacb200_decode_batch_host's double buffering
    // pipelined host path
    Buf hblob;                              // device copy of the caller's blob
    cudaStream_t h2d = nullptr, d2h = nullptr;
    std::vector<cudaEvent_t> ev_h2d;
    cudaEvent_t ev_pcm_free = nullptr, ev_post = nullptr;
    bool pcm_busy = false;
};

void dec_free(void* p) {
    DecState* d = (DecState*)p;
    Buf* bufs[] = {&d->blob, &d->soff, &d->slen, &d->meta, &d->segs, &d->segcount, &d->segbase, &d->cands, &d->sizes, &d->slotoff, &d->samples,
                   &d->candfirst, &d->res, &d->ss32, &d->pcmoff, &d->pcm, &d->total, &d->firstseg};
    if (d->m_total) cudaFreeHost(d->m_total);
    if (d->m_res) cudaFreeHost(d->m_res);
    for (Buf* b : bufs) b->release();
    d->hblob.release();
    for (auto& e : d->ev) if (e) cudaEventDestroy(e);
    for (auto& e : d->ev_h2d) if (e) cudaEventDestroy(e);
    if (d->ev_pcm_free) cudaEventDestroy(d->ev_pcm_free);
    if (d->ev_post) cudaEventDestroy(d->ev_post);
    if (d->h2d) cudaStreamDestroy(d->h2d);
    if (d->d2h) cudaStreamDestroy(d->d2h);
    if (d->alt) dec_free(d->alt);
    delete d;
}

DecState* state(flacb200_ctx* ctx, int which = 0) {
    void** slot = fb_ctx_dec_slot(ctx, dec_free);
    if (!*slot) { DecState* d = new DecState(); for (auto& e : d->ev) cudaEventCreate(&e); *slot = d; }
    DecState* d = (DecState*)*slot;
    if (which == 0) return d;
    if (!d->alt) { d->alt = new DecState(); for (auto& e : d->alt->ev) cudaEventCreate(&e); }     // second lane of the pipelined host path
    return d->alt;
}

}  // namespace

#define CKD(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fb_ctx_fail(ctx, FLACB200_ERR_CUDA, #call, e_); } while (0)

static int decode_core(flacb200_ctx* ctx, DecState* d, cudaStream_t st, const uint8_t* d_blob, int ns, const uint64_t* stream_off,
                       const uint64_t* stream_len, uint32_t out_container_bytes, const flacb200_dec_raw_params* raw,
                       const std::function<int()>* after_uploads = nullptr);

extern "C" int flacb200_decode_batch(flacb200_ctx* ctx, const uint8_t* blob, int blob_is_device, uint64_t blob_bytes,
                                     uint32_t n_streams, const uint64_t* stream_off, const uint64_t* stream_len,
                                     uint32_t out_container_bytes, const flacb200_dec_raw_params* raw) {
    if (!ctx) return FLACB200_ERR_NO_DEVICE;
    if ((!blob && blob_bytes) || (n_streams && (!stream_off || !stream_len))) return fb_ctx_fail(ctx, FLACB200_ERR_ARG, "null argument", cudaSuccess);
    if (out_container_bytes != 0 && out_container_bytes != 2 && out_container_bytes != 4) return fb_ctx_fail(ctx, FLACB200_ERR_ARG, "out_container_bytes must be 0, 2 or 4", cudaSuccess);
    cudaSetDevice(fb_ctx_device(ctx));
    DecState* d = state(ctx);
    cudaStream_t st = fb_ctx_stream(ctx);
    d->have = false;
    const int ns = (int)n_streams;
    for (int s = 0; s < ns; s++) {
        if (stream_off[s] + stream_len[s] > blob_bytes) return fb_ctx_fail(ctx, FLACB200_ERR_ARG, "stream exceeds blob", cudaSuccess);
        if (stream_len[s] >= 0xFFFFFFF0ull) return fb_ctx_fail(ctx, FLACB200_ERR_UNSUPPORTED, "streams of 4 GiB or more are not supported", cudaSuccess);
    }
    const uint8_t* d_blob = blob;
    if (!blob_is_device && ns) {
        CKD(d->blob.reserve(blob_bytes + 64));
        CKD(cudaMemcpyAsync(d->blob.p, blob, blob_bytes, cudaMemcpyHostToDevice, st));
        CKD(cudaMemsetAsync((uint8_t*)d->blob.p + blob_bytes, 0, 16, st));
        d_blob = (const uint8_t*)d->blob.p;
    }
    return decode_core(ctx, d, st, d_blob, ns, stream_off, stream_len, out_container_bytes, raw);
}

// One decode pass over streams whose bytes are already in HBM (d_blob + stream_off[s]).
static int decode_core(flacb200_ctx* ctx, DecState* d, cudaStream_t st, const uint8_t* d_blob, int ns, const uint64_t* stream_off,
                       const uint64_t* stream_len, uint32_t out_container_bytes, const flacb200_dec_raw_params* raw,
                       const std::function<int()>* after_uploads) {
    d->have = false;
    // segments: every stream is cut into 4096-byte pieces scanned by one warp each.  The table (75 entries per 300 KB stream: 3.7 MB
    // for the 4096-stream batch) took 1 ms of host time to build and upload per call; a batch with the layout of the previous one
    // finds it on the device.
    const bool same_layout = ns > 0 && d->c_off.size() == (size_t)ns && memcmp(d->c_off.data(), stream_off, 8 * (size_t)ns) == 0 &&
                             memcmp(d->c_len.data(), stream_len, 8 * (size_t)ns) == 0;
    if (!same_layout) {
        d->c_off.clear(); d->c_len.clear();
        d->h_segs.clear(); d->h_first_seg.assign(ns + 1, 0);
        for (int s = 0; s < ns; s++) {
            d->h_first_seg[s] = (uint32_t)d->h_segs.size();
            for (uint64_t o = 0; o < stream_len[s]; o += kDecSegBytes) {
                DecSegment g; g.stream = (uint32_t)s; g.start = (uint32_t)o;
                g.bytes = (uint32_t)((stream_len[s] - o < kDecSegBytes) ? (stream_len[s] - o) : kDecSegBytes);
                d->h_segs.push_back(g);
            }
        }
        d->h_first_seg[ns] = (uint32_t)d->h_segs.size();
    }
    const int nsegs = (int)d->h_segs.size();
    d->n_streams = ns; d->n_cands = 0; d->total_elems = 0;
    if (ns == 0) { d->have = true; d->out_bytes = out_container_bytes ? out_container_bytes : 2; return 0; }

    CKD(d->soff.reserve(8 * (size_t)(ns + 1))); CKD(d->slen.reserve(8 * (size_t)(ns + 1)));
    CKD(d->meta.reserve(sizeof(DecStreamMeta) * (size_t)ns));
    CKD(d->segs.reserve(sizeof(DecSegment) * (size_t)(nsegs + 1)));
    CKD(d->segcount.reserve(4 * (size_t)(nsegs + 1))); CKD(d->segbase.reserve(4 * (size_t)(nsegs + 2)));
    CKD(d->candfirst.reserve(4 * (size_t)(ns + 2)));
    CKD(d->res.reserve(sizeof(DecStreamResult) * (size_t)ns)); CKD(d->ss32.reserve(4 * (size_t)(ns + 1)));
    CKD(d->pcmoff.reserve(8 * (size_t)(ns + 1))); CKD(d->total.reserve(64));
    if (!d->m_total) CKD(cudaHostAlloc((void**)&d->m_total, 64, cudaHostAllocMapped));
    if (d->m_res_cap < sizeof(DecStreamResult) * (size_t)ns) {
        if (d->m_res) cudaFreeHost(d->m_res);
        d->m_res = nullptr; d->m_res_cap = 0;
        const size_t want = sizeof(DecStreamResult) * (size_t)ns * 2 + 256;
        CKD(cudaHostAlloc((void**)&d->m_res, want, cudaHostAllocMapped));
        d->m_res_cap = want;
    }
    CKD(d->firstseg.reserve(4 * (size_t)(ns + 2)));
    if (!same_layout) {
        CKD(cudaMemcpyAsync(d->firstseg.p, d->h_first_seg.data(), 4 * (size_t)(ns + 1), cudaMemcpyHostToDevice, st));
        CKD(cudaMemcpyAsync(d->soff.p, stream_off, 8 * (size_t)ns, cudaMemcpyHostToDevice, st));
        CKD(cudaMemcpyAsync(d->slen.p, stream_len, 8 * (size_t)ns, cudaMemcpyHostToDevice, st));
        if (nsegs) CKD(cudaMemcpyAsync(d->segs.p, d->h_segs.data(), sizeof(DecSegment) * (size_t)nsegs, cudaMemcpyHostToDevice, st));
        d->c_off.assign(stream_off, stream_off + ns); d->c_len.assign(stream_len, stream_len + ns);
    }
    // the last H2D copies of this pass are in the queue: the pipelined host path now submits the next chunk's bytes
    if (after_uploads) { const int rcu = (*after_uploads)(); if (rcu) return rcu; }

    DecStreamMeta rawp; memset(&rawp, 0, sizeof rawp);
    if (raw) {
        rawp.sample_rate = raw->sample_rate; rawp.channels = raw->channels; rawp.bps = raw->bits_per_sample;
        rawp.min_blocksize = rawp.max_blocksize = raw->fixed_blocksize;
        rawp.have_last = (raw->flags & 2u) ? 1u : 0u; rawp.last_blocksize = raw->last_blocksize; rawp.next_sample = raw->next_sample;
    }
    CKD(cudaEventRecord(d->ev[0], st));
    launch_dec_meta(d_blob, (const uint64_t*)d->soff.p, (const uint64_t*)d->slen.p, ns, raw ? 1 : 0, rawp, (DecStreamMeta*)d->meta.p, st);
    uint64_t h_total = 0;
    if (nsegs) {
        launch_dec_sync(d_blob, (const uint64_t*)d->soff.p, (const uint64_t*)d->slen.p, (const DecStreamMeta*)d->meta.p, (const DecSegment*)d->segs.p, nsegs, 0,
                        (uint32_t*)d->segcount.p, nullptr, nullptr, st);
        launch_dec_scan((const uint32_t*)d->segcount.p, nsegs, (uint32_t*)d->segbase.p, nullptr, (uint64_t*)d->total.p, st);
        launch_dec_mirror(d->total.p, d->m_total, 8, st);
        CKD(cudaStreamSynchronize(st));                                  // sync #1: number of candidates
        h_total = d->m_total[0];
    }
    const int nc = (int)h_total;
    d->n_cands = nc;
    CKD(d->cands.reserve(sizeof(DecCand) * (size_t)(nc + 1)));
    CKD(d->sizes.reserve(4 * (size_t)(nc + 1))); CKD(d->slotoff.reserve(8 * (size_t)(nc + 1)));
    if (nsegs && nc)
        launch_dec_sync(d_blob, (const uint64_t*)d->soff.p, (const uint64_t*)d->slen.p, (const DecStreamMeta*)d->meta.p, (const DecSegment*)d->segs.p, nsegs, 1,
                        (uint32_t*)d->segcount.p, (const uint32_t*)d->segbase.p, (DecCand*)d->cands.p, st);
    // first candidate of each stream = scan value at the stream's first segment (+ sentinel = nc), computed on the device
    launch_dec_cand_first((const uint32_t*)d->segbase.p, (const uint32_t*)d->firstseg.p, ns, nsegs, (const uint64_t*)d->total.p, (uint32_t*)d->candfirst.p, st);
    CKD(cudaEventRecord(d->ev[1], st));
    uint64_t h_slots = 0;
    if (nc) {
        launch_dec_cand_size((const DecCand*)d->cands.p, nc, (uint32_t*)d->sizes.p, st);
        launch_dec_scan((const uint32_t*)d->sizes.p, nc, nullptr, (uint64_t*)d->slotoff.p, (uint64_t*)d->total.p, st);
        launch_dec_mirror(d->total.p, d->m_total, 8, st);
        CKD(cudaStreamSynchronize(st));                                  // sync #2: scratch size for candidate samples
        h_slots = d->m_total[0];
        CKD(d->samples.reserve(4 * (size_t)(h_slots + 16)));
        launch_dec_frames(d_blob, (const uint64_t*)d->soff.p, (const uint64_t*)d->slen.p, (DecCand*)d->cands.p, nc, (const uint64_t*)d->slotoff.p, (int32_t*)d->samples.p, st);
    }
    CKD(cudaEventRecord(d->ev[5], st));
    if (nc) launch_dec_crc(d_blob, (const uint64_t*)d->soff.p, (DecCand*)d->cands.p, nc, st);
    CKD(cudaEventRecord(d->ev[2], st));
    // total (8 bytes) is followed by the "some stream has silence to fill" flag
    CKD(cudaMemsetAsync((uint8_t*)d->total.p + 8, 0, 8, st));
    const int eof = (raw && (raw->flags & 1u)) ? 0 : 1;         // raw bit 0: more input may follow (streaming callers)
    launch_dec_chain((DecCand*)d->cands.p, (const uint32_t*)d->candfirst.p, (const DecStreamMeta*)d->meta.p, (const uint64_t*)d->slen.p, ns, eof,
                     (DecStreamResult*)d->res.p, (uint32_t*)d->ss32.p, (unsigned int*)((uint8_t*)d->total.p + 8), st);
    launch_dec_scan((const uint32_t*)d->ss32.p, ns, nullptr, (uint64_t*)d->pcmoff.p, (uint64_t*)d->total.p, st);
    launch_dec_assign((DecStreamResult*)d->res.p, (const uint64_t*)d->pcmoff.p, ns, st);
    d->h_res.resize(ns);
    uint64_t h_elems = 0;
    launch_dec_mirror(d->total.p, d->m_total, 16, st);
    launch_dec_mirror(d->res.p, d->m_res, sizeof(DecStreamResult) * (size_t)ns, st);
    CKD(cudaStreamSynchronize(st));                                      // sync #3: PCM size
    h_elems = d->m_total[0];
    const bool any_gap = d->m_total[1] != 0;
    memcpy(d->h_res.data(), d->m_res, sizeof(DecStreamResult) * (size_t)ns);
    uint32_t ob = out_container_bytes;
    if (ob == 0) { ob = 2; for (int s = 0; s < ns; s++) if (d->h_res[s].bps > 16) ob = 4; }
    for (int s = 0; s < ns; s++) if (ob == 2 && d->h_res[s].bps > 16 && d->h_res[s].n_frames) return fb_ctx_fail(ctx, FLACB200_ERR_ARG, "int16 output requested for a >16-bit stream", cudaSuccess);
    d->out_bytes = ob; d->total_elems = h_elems;
    CKD(d->pcm.reserve((size_t)h_elems * ob + 64));
    if (any_gap) CKD(cudaMemsetAsync(d->pcm.p, 0, (size_t)h_elems * ob, st));     // missing frames stand as silence (rare: corrupted input)
    CKD(cudaEventRecord(d->ev[3], st));
    launch_dec_post(d_blob, (const uint64_t*)d->soff.p, (DecCand*)d->cands.p, nc, (const uint64_t*)d->slotoff.p, (const int32_t*)d->samples.p,
                    (DecStreamResult*)d->res.p, d->pcm.p, (int)ob, st);
    CKD(cudaEventRecord(d->ev[4], st));
    CKD(cudaGetLastError());
    fb_ctx_add_launches(ctx, 12);
    d->have = true;
    return 0;
}

extern "C" int flacb200_decode_result(flacb200_ctx* ctx, flacb200_dec_result* res) {
    if (!ctx || !res) return FLACB200_ERR_ARG;
    DecState* d = state(ctx);
    if (!d->have) return fb_ctx_fail(ctx, FLACB200_ERR_ARG, "no decode batch", cudaSuccess);
    cudaSetDevice(fb_ctx_device(ctx));
    cudaStream_t st = fb_ctx_stream(ctx);
    if (d->n_streams) {
        CKD(cudaMemcpyAsync(d->h_res.data(), d->res.p, sizeof(DecStreamResult) * (size_t)d->n_streams, cudaMemcpyDeviceToHost, st));   // post may have flagged CRC errors
    }
    CKD(cudaStreamSynchronize(st));
    uint32_t nf = 0;
    for (int s = 0; s < d->n_streams; s++) nf += d->h_res[s].n_frames;
    res->total_elems = d->total_elems; res->n_streams = (uint32_t)d->n_streams; res->n_frames = nf; res->out_container_bytes = d->out_bytes;
    res->n_candidates = (uint32_t)d->n_cands; res->d_pcm = d->pcm.p;
    return 0;
}

extern "C" int flacb200_decode_fetch(flacb200_ctx* ctx, void* pcm, size_t pcm_cap, flacb200_dec_stream_info* streams, uint32_t* frame_samples, uint32_t frame_cap) {
    flacb200_dec_result r;
    int rc = flacb200_decode_result(ctx, &r);
    if (rc) return rc;
    DecState* d = state(ctx);
    cudaStream_t st = fb_ctx_stream(ctx);
    if (pcm) {
        const size_t bytes = (size_t)r.total_elems * r.out_container_bytes;
        if (pcm_cap < bytes) return fb_ctx_fail(ctx, FLACB200_ERR_ARG, "pcm buffer too small", cudaSuccess);
        if (bytes) CKD(cudaMemcpyAsync(pcm, d->pcm.p, bytes, cudaMemcpyDeviceToHost, st));
    }
    std::vector<DecCand> hc;
    if (frame_samples && d->n_cands) {
        hc.resize(d->n_cands);
        CKD(cudaMemcpyAsync(hc.data(), d->cands.p, sizeof(DecCand) * (size_t)d->n_cands, cudaMemcpyDeviceToHost, st));
    }
    CKD(cudaStreamSynchronize(st));
    if (streams) {
        for (int s = 0; s < d->n_streams; s++) {
            const DecStreamResult& h = d->h_res[s];
            flacb200_dec_stream_info& o = streams[s];
            o.total_samples = h.total_samples; o.pcm_off = h.pcm_off; o.consumed = h.consumed; o.n_frames = h.n_frames; o.status = h.status;
            o.sample_rate = h.sample_rate; o.channels = h.channels; o.bits_per_sample = h.bps; o.max_blocksize = h.max_blocksize;
            o.n_events = h.n_events; o.gap_samples = h.gap_samples; o.next_sample = h.next_sample; o.last_blocksize = h.last_blocksize; o.have_last = h.have_last;
            for (int k = 0; k < kDecMaxEvents; k++) { o.ev_frame[k] = h.ev_frame[k]; o.ev_status[k] = h.ev_status[k]; }
        }
    }
    if (frame_samples) { uint32_t k = 0; for (auto& c : hc) if (c.valid && k < frame_cap) frame_samples[k++] = c.blocksize; }
    return 0;
}

// sample offset (within its stream's PCM, silence included) of every delivered frame, in the order of flacb200_decode_fetch's frame_samples
extern "C" int flacb200_decode_fetch_frame_offsets(flacb200_ctx* ctx, uint64_t* frame_sample_off, uint32_t frame_cap) {
    if (!ctx || !frame_sample_off) return FLACB200_ERR_ARG;
    DecState* d = state(ctx);
    if (!d->have) return fb_ctx_fail(ctx, FLACB200_ERR_ARG, "no decode batch", cudaSuccess);
    cudaSetDevice(fb_ctx_device(ctx));
    cudaStream_t st = fb_ctx_stream(ctx);
    std::vector<DecCand> hc(d->n_cands);
    if (d->n_cands) CKD(cudaMemcpyAsync(hc.data(), d->cands.p, sizeof(DecCand) * (size_t)d->n_cands, cudaMemcpyDeviceToHost, st));
    CKD(cudaStreamSynchronize(st));
    uint32_t k = 0;
    for (auto& c : hc) if (c.valid && k < frame_cap) frame_sample_off[k++] = c.sample_off;
    return 0;
}

// Host -> host decode in one call: the streams are cut into chunks; chunk c+1's bytes travel to the GPU while chunk c
// decodes and chunk c-1's PCM travels back (two decode states alternate, so a chunk's PCM buffer is not reused before
// its copy has left).  PCM lands in `pcm` in stream order, streams[s].pcm_off indexes it.
extern "C" int flacb200_decode_batch_host(flacb200_ctx* ctx, const uint8_t* blob, uint64_t blob_bytes, uint32_t n_streams,
                                          const uint64_t* stream_off, const uint64_t* stream_len, uint32_t out_container_bytes,
                                          const flacb200_dec_raw_params* raw, void* pcm, size_t pcm_cap, uint64_t* total_elems,
                                          flacb200_dec_stream_info* streams) {
    if (!ctx) return FLACB200_ERR_NO_DEVICE;
    if ((!blob && blob_bytes) || !pcm || (n_streams && (!stream_off || !stream_len))) return fb_ctx_fail(ctx, FLACB200_ERR_ARG, "null argument", cudaSuccess);
    if (out_container_bytes != 2 && out_container_bytes != 4) return fb_ctx_fail(ctx, FLACB200_ERR_ARG, "out_container_bytes must be 2 or 4", cudaSuccess);
    cudaSetDevice(fb_ctx_device(ctx));
    const int ns = (int)n_streams;
    if (total_elems) *total_elems = 0;
    for (int s = 0; s < ns; s++) {
        if (stream_off[s] + stream_len[s] > blob_bytes) return fb_ctx_fail(ctx, FLACB200_ERR_ARG, "stream exceeds blob", cudaSuccess);
        if (stream_len[s] >= 0xFFFFFFF0ull) return fb_ctx_fail(ctx, FLACB200_ERR_UNSUPPORTED, "streams of 4 GiB or more are not supported", cudaSuccess);
    }
    if (ns == 0) return 0;
    DecState* D[2] = {state(ctx, 0), state(ctx, 1)};
    DecState* d0 = D[0];
    cudaStream_t st = fb_ctx_stream(ctx);
    if (!d0->h2d) { CKD(cudaStreamCreateWithFlags(&d0->h2d, cudaStreamNonBlocking)); CKD(cudaStreamCreateWithFlags(&d0->d2h, cudaStreamNonBlocking)); }
    for (DecState* d : D) {
        if (!d->ev_pcm_free) { CKD(cudaEventCreateWithFlags(&d->ev_pcm_free, cudaEventDisableTiming)); CKD(cudaEventCreateWithFlags(&d->ev_post, cudaEventDisableTiming)); }
        d->pcm_busy = false; d->have = false;
    }
    // chunks of whole streams with about equal byte counts
    // Chunks must still fill the GPU (the frame kernel decodes one frame per thread and a launch cannot finish faster than
    // one frame's serial decode, about 2 ms): one chunk per ~200 MB of FLAC (measured on 1.26 GB: 1/2/4/6/8 chunks take
    // 72/68/57.5/55/56 ms).  All table uploads of a chunk are submitted before
    // the next chunk's bytes and every readback goes through mapped memory, so the bulk copies never delay the small ones.
    int nchunks = (int)(blob_bytes / (200ull << 20));
    if (const char* ev = getenv("FLACB200_DEC_CHUNKS")) { const int v = atoi(ev); if (v > 0) nchunks = v; }
    if (nchunks < 1) nchunks = 1;
    if (nchunks > 12) nchunks = 12;
    if (nchunks > ns) nchunks = ns;
    std::vector<int> cs(nchunks + 1, 0);
    {
        uint64_t tot = 0; for (int s = 0; s < ns; s++) tot += stream_len[s];
        uint64_t acc = 0; int c = 1;
        for (int s = 0; s < ns && c < nchunks; s++) { acc += stream_len[s]; if (acc * nchunks >= tot * c && s + 1 >= c) cs[c++] = s + 1; }
        for (; c <= nchunks; c++) cs[c] = ns;
    }
    while ((int)d0->ev_h2d.size() < nchunks) { cudaEvent_t e; CKD(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); d0->ev_h2d.push_back(e); }
    CKD(d0->hblob.reserve(blob_bytes + 64));
    uint8_t* d_blob = (uint8_t*)d0->hblob.p;
    CKD(cudaMemsetAsync(d_blob + blob_bytes, 0, 16, d0->h2d));
    bool monotonic = true;
    for (int s = 1; s < ns; s++) if (stream_off[s] < stream_off[s - 1] + stream_len[s - 1]) { monotonic = false; break; }
    // blob chunk c+1 is enqueued right before chunk c is decoded: the copy engine runs in submission order, so chunk c's
    // small table uploads (issued inside decode_core) only wait for transfers they need anyway
    auto enqueue_blob = [&](int c) -> int {
        const int s0 = cs[c], s1 = cs[c + 1];
        if (s1 > s0) {
            if (monotonic) {
                const uint64_t b0 = stream_off[s0], b1 = stream_off[s1 - 1] + stream_len[s1 - 1];
                CKD(cudaMemcpyAsync(d_blob + b0, blob + b0, b1 - b0, cudaMemcpyHostToDevice, d0->h2d));
            } else {
                for (int s = s0; s < s1; s++) CKD(cudaMemcpyAsync(d_blob + stream_off[s], blob + stream_off[s], stream_len[s], cudaMemcpyHostToDevice, d0->h2d));
            }
        }
        CKD(cudaEventRecord(d0->ev_h2d[c], d0->h2d));
        return 0;
    };
    { const int rc0 = enqueue_blob(0); if (rc0) return rc0; }
    uint64_t elem_base = 0;
    int status = 0;
    for (int c = 0; c < nchunks; c++) {
        const int s0 = cs[c], s1 = cs[c + 1];
        if (s1 <= s0) continue;
        DecState* d = D[c & 1];
        CKD(cudaStreamWaitEvent(st, d0->ev_h2d[c], 0));
        const std::function<int()> next_blob = [&]() -> int { return (c + 1 < nchunks) ? enqueue_blob(c + 1) : 0; };
        if (d->pcm_busy) { CKD(cudaStreamWaitEvent(st, d->ev_pcm_free, 0)); d->pcm_busy = false; }     // its previous PCM must have left
        const int rc = decode_core(ctx, d, st, d_blob, s1 - s0, stream_off + s0, stream_len + s0, out_container_bytes, raw, &next_blob);
        if (rc) { status = rc; break; }
        CKD(cudaEventRecord(d->ev_post, st));
        // per-stream results of the chunk (post may flag CRC errors: read them after it)
        launch_dec_mirror(d->res.p, d->m_res, sizeof(DecStreamResult) * (size_t)(s1 - s0), st);
        const size_t bytes = (size_t)d->total_elems * out_container_bytes;
        if ((elem_base + d->total_elems) * out_container_bytes > pcm_cap) { status = fb_ctx_fail(ctx, FLACB200_ERR_ARG, "pcm buffer too small", cudaSuccess); break; }
        CKD(cudaStreamWaitEvent(d0->d2h, d->ev_post, 0));
        if (bytes) CKD(cudaMemcpyAsync((uint8_t*)pcm + (size_t)elem_base * out_container_bytes, d->pcm.p, bytes, cudaMemcpyDeviceToHost, d0->d2h));
        CKD(cudaEventRecord(d->ev_pcm_free, d0->d2h));
        d->pcm_busy = true;
        CKD(cudaStreamSynchronize(st));
        memcpy(d->h_res.data(), d->m_res, sizeof(DecStreamResult) * (size_t)(s1 - s0));
        if (streams) {
            for (int s = s0; s < s1; s++) {
                const DecStreamResult& h = d->h_res[s - s0];
                flacb200_dec_stream_info& o = streams[s];
                o.total_samples = h.total_samples; o.pcm_off = h.pcm_off + elem_base; o.consumed = h.consumed; o.n_frames = h.n_frames; o.status = h.status;
                o.sample_rate = h.sample_rate; o.channels = h.channels; o.bits_per_sample = h.bps; o.max_blocksize = h.max_blocksize;
                o.n_events = h.n_events; o.gap_samples = h.gap_samples; o.next_sample = h.next_sample; o.last_blocksize = h.last_blocksize; o.have_last = h.have_last;
                for (int k = 0; k < kDecMaxEvents; k++) { o.ev_frame[k] = h.ev_frame[k]; o.ev_status[k] = h.ev_status[k]; }
            }
        }
        elem_base += d->total_elems;
    }
    CKD(cudaStreamSynchronize(d0->d2h));
    CKD(cudaStreamSynchronize(d0->h2d));
    for (DecState* d : D) d->pcm_busy = false;
    if (status) return status;
    if (total_elems) *total_elems = elem_base;
    return 0;
}

extern "C" int flacb200_decode_kernel_times(flacb200_ctx* ctx, float* ms) {
    if (!ctx || !ms) return FLACB200_ERR_ARG;
    DecState* d = state(ctx);
    if (!d->have || !d->n_streams) return FLACB200_ERR_ARG;
    cudaSetDevice(fb_ctx_device(ctx));
    CKD(cudaStreamSynchronize(fb_ctx_stream(ctx)));
    for (int i = 0; i < 4; i++) CKD(cudaEventElapsedTime(&ms[i], d->ev[i], d->ev[i + 1]));
    CKD(cudaEventElapsedTime(&ms[4], d->ev[5], d->ev[2]));       // the CRC-16 share of ms[1]
    ms[5] = 0.0f;
    return 0;
}
