// flac_api_enc.cu -- drop-in FLAC__stream_encoder_* layer (include/flacb200_flac_api.h) on top of the batch engine.
//
// Host logic only: settings, the libFLAC framing rules (a frame is produced once blocksize+1
// samples are buffered; finish() flushes the remainder as a short frame), stream prologue, STREAMINFO
// rewrite through seek/tell, callback dispatch in order.  All per-frame arithmetic happens in the CUDA
// kernels via flacb200_encode_batch; one process_interleaved() call encodes every complete frame it holds
// in ONE batch.  MD5 of a handle-fed stream is accumulated incrementally on the host as the samples arrive
// (the batch ABI hashes on the GPU; see DESIGN.md "MD5 placement").
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <vector>

#include "../../include/flacb200.h"
#include "../../include/flacb200_flac_api.h"
#include "fb_common.cuh"

extern "C" {
// libFLAC 1.4.3 string tables (values read from the reference binary; pyflac/encoder.py:42,54)
const char *const FLAC__StreamEncoderStateString[] = {
    "FLAC__STREAM_ENCODER_OK", "FLAC__STREAM_ENCODER_UNINITIALIZED", "FLAC__STREAM_ENCODER_OGG_ERROR",
    "FLAC__STREAM_ENCODER_VERIFY_DECODER_ERROR", "FLAC__STREAM_ENCODER_VERIFY_MISMATCH_IN_AUDIO_DATA",
    "FLAC__STREAM_ENCODER_CLIENT_ERROR", "FLAC__STREAM_ENCODER_IO_ERROR", "FLAC__STREAM_ENCODER_FRAMING_ERROR",
    "FLAC__STREAM_ENCODER_MEMORY_ALLOCATION_ERROR"};
const char *const FLAC__StreamEncoderInitStatusString[] = {
    "FLAC__STREAM_ENCODER_INIT_STATUS_OK", "FLAC__STREAM_ENCODER_INIT_STATUS_ENCODER_ERROR",
    "FLAC__STREAM_ENCODER_INIT_STATUS_UNSUPPORTED_CONTAINER", "FLAC__STREAM_ENCODER_INIT_STATUS_INVALID_CALLBACKS",
    "FLAC__STREAM_ENCODER_INIT_STATUS_INVALID_NUMBER_OF_CHANNELS", "FLAC__STREAM_ENCODER_INIT_STATUS_INVALID_BITS_PER_SAMPLE",
    "FLAC__STREAM_ENCODER_INIT_STATUS_INVALID_SAMPLE_RATE", "FLAC__STREAM_ENCODER_INIT_STATUS_INVALID_BLOCK_SIZE",
    "FLAC__STREAM_ENCODER_INIT_STATUS_INVALID_MAX_LPC_ORDER", "FLAC__STREAM_ENCODER_INIT_STATUS_INVALID_QLP_COEFF_PRECISION",
    "FLAC__STREAM_ENCODER_INIT_STATUS_BLOCK_SIZE_TOO_SMALL_FOR_LPC_ORDER", "FLAC__STREAM_ENCODER_INIT_STATUS_NOT_STREAMABLE",
    "FLAC__STREAM_ENCODER_INIT_STATUS_INVALID_METADATA", "FLAC__STREAM_ENCODER_INIT_STATUS_ALREADY_INITIALIZED"};
const char *FLAC__VENDOR_STRING = "reference libFLAC 1.4.3 20230623";   // byte-identical files need the oracle's vendor string
}

namespace {

enum { ST_OK = 0, ST_UNINITIALIZED = 1, ST_VERIFY_DECODER_ERROR = 3, ST_VERIFY_MISMATCH = 4, ST_CLIENT_ERROR = 5, ST_IO_ERROR = 6, ST_FRAMING_ERROR = 7, ST_MEMORY_ALLOCATION_ERROR = 8 };
enum { INIT_OK = 0, INIT_ENCODER_ERROR = 1, INIT_UNSUPPORTED_CONTAINER = 2, INIT_INVALID_CALLBACKS = 3, INIT_ALREADY_INITIALIZED = 13 };

// RFC 1321, incremental (host side of the handle API only)
struct Md5 {
    uint32_t h[4]; uint64_t len; uint8_t buf[64]; uint32_t fill;
    void init() { h[0] = 0x67452301u; h[1] = 0xefcdab89u; h[2] = 0x98badcfeu; h[3] = 0x10325476u; len = 0; fill = 0; }
    static uint32_t rol(uint32_t v, int s) { return (v << s) | (v >> (32 - s)); }
    void block(const uint8_t* p) {
        static const uint32_t K[64] = {
            0xd76aa478,0xe8c7b756,0x242070db,0xc1bdceee,0xf57c0faf,0x4787c62a,0xa8304613,0xfd469501,0x698098d8,0x8b44f7af,0xffff5bb1,0x895cd7be,
            0x6b901122,0xfd987193,0xa679438e,0x49b40821,0xf61e2562,0xc040b340,0x265e5a51,0xe9b6c7aa,0xd62f105d,0x02441453,0xd8a1e681,0xe7d3fbc8,
            0x21e1cde6,0xc33707d6,0xf4d50d87,0x455a14ed,0xa9e3e905,0xfcefa3f8,0x676f02d9,0x8d2a4c8a,0xfffa3942,0x8771f681,0x6d9d6122,0xfde5380c,
            0xa4beea44,0x4bdecfa9,0xf6bb4b60,0xbebfbc70,0x289b7ec6,0xeaa127fa,0xd4ef3085,0x04881d05,0xd9d4d039,0xe6db99e5,0x1fa27cf8,0xc4ac5665,
            0xf4292244,0x432aff97,0xab9423a7,0xfc93a039,0x655b59c3,0x8f0ccc92,0xffeff47d,0x85845dd1,0x6fa87e4f,0xfe2ce6e0,0xa3014314,0x4e0811a1,
            0xf7537e82,0xbd3af235,0x2ad7d2bb,0xeb86d391};
        static const int S[64] = {7,12,17,22,7,12,17,22,7,12,17,22,7,12,17,22,5,9,14,20,5,9,14,20,5,9,14,20,5,9,14,20,
                                  4,11,16,23,4,11,16,23,4,11,16,23,4,11,16,23,6,10,15,21,6,10,15,21,6,10,15,21,6,10,15,21};
        uint32_t w[16], a = h[0], b = h[1], c = h[2], d = h[3];
        for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] | (uint32_t)p[4 * i + 1] << 8 | (uint32_t)p[4 * i + 2] << 16 | (uint32_t)p[4 * i + 3] << 24;
        for (int i = 0; i < 64; i++) {
            uint32_t f; int g;
            if (i < 16) { f = (b & c) | (~b & d); g = i; }
            else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
            else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
            else { f = c ^ (b | ~d); g = (7 * i) & 15; }
            const uint32_t t = a + f + K[i] + w[g];
            a = d; d = c; c = b; b = b + rol(t, S[i]);
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d;
    }
    void update(const uint8_t* p, size_t n) {
        len += n;
        while (n) {
            size_t k = 64 - fill; if (k > n) k = n;
            memcpy(buf + fill, p, k); fill += (uint32_t)k; p += k; n -= k;
            if (fill == 64) { block(buf); fill = 0; }
        }
    }
    void final(uint8_t out[16]) {
        const uint64_t bits = len * 8; const uint8_t pad = 0x80, z = 0; uint8_t lb[8];
        update(&pad, 1);
        while (fill != 56) update(&z, 1);
        for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (8 * i));
        update(lb, 8);
        for (int i = 0; i < 16; i++) out[i] = (uint8_t)(h[i >> 2] >> (8 * (i & 3)));
    }
};

// ------------------------------------------------------------------------------------------------ dispatcher
// pyFLAC runs one FLAC__StreamEncoder per stream and, for many streams, one thread per encoder
// (/root/reference/pyflac/encoder.py:293-330; libFLAC handles are independent).  A GPU wants the opposite:
// many streams per launch.  The dispatcher sits between the two: every process_interleaved() / finish() that has whole
// frames to encode submits a job and blocks; the first submitter becomes the leader, waits a short gather window for the
// other handles that are in flight, encodes ALL compatible jobs (same settings) in ONE flacb200_encode_batch call and hands
// every job its own frames.  Each caller then fires its write callbacks on its own thread, outside any lock -- a callback
// may use another encoder.  One dispatcher (and one engine context) per CUDA device; handles pick a device from
// FLACB200_DEVICE (one index) or round-robin over FLACB200_DEVICES (comma separated list).
struct EncJob {
    // in
    flacb200_enc_config cfg; const int32_t* pcm; uint64_t n; uint32_t ffn; uint8_t prev_ca; bool loose, verify;
    // out
    std::vector<uint8_t> bytes; std::vector<uint32_t> flen, fsmp; uint8_t last_ca = 0;
    int rc = 0;                 // 0 ok, 1 engine error, 2 verify decoder error, 3 verify mismatch
    uint64_t v_sample = 0; uint32_t v_channel = 0; int32_t v_expected = 0, v_got = 0;
    bool done = false;
};

struct Dispatcher {
    std::mutex mu; std::condition_variable cv;
    std::deque<EncJob*> q;
    bool running = false;
    int device = 0; flacb200_ctx* ctx = nullptr; int ctx_rc = -1;
    std::atomic<int> active{0};                 // initialised encoder handles on this device
    // leader-only scratch (pinned)
    void* h_pcm = nullptr; size_t h_pcm_cap = 0;
    std::vector<uint8_t> arena; std::vector<uint64_t> foff; std::vector<uint32_t> flen, fsmp, fstream; std::vector<uint8_t> ca;
    uint64_t batches = 0, jobs = 0;
    int gather_us = 400;                        // adaptive: halves whenever a window closes with a single job, back to 400 when a batch had company

    flacb200_ctx* context() {
        if (ctx_rc == -1) ctx_rc = flacb200_create(&ctx, device);
        return ctx_rc == 0 ? ctx : nullptr;
    }
    static bool same_key(const flacb200_enc_config& a, const flacb200_enc_config& b) { return memcmp(&a, &b, sizeof a) == 0; }

    void run(std::vector<EncJob*>& B);
    void submit(EncJob* job) {
        std::unique_lock<std::mutex> lk(mu);
        q.push_back(job);
        cv.notify_all();
        while (!job->done) {
            if (running) { cv.wait(lk); continue; }
            running = true;                                   // this thread leads one batch
            const int others = active.load();
            if (others > 1 && gather_us > 0 && !getenv("FLACB200_NO_GATHER")) {
                // gather window: the other handles of this device are probably about to submit too (while a batch runs,
                // later submissions queue up and form the next batch by themselves; the window only helps the first one)
                const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(gather_us);
                while ((int)q.size() < others && cv.wait_until(lk, deadline) != std::cv_status::timeout) {}
            }
            std::vector<EncJob*> B;
            const flacb200_enc_config key = q.front()->cfg;
            for (auto it = q.begin(); it != q.end();) { if (same_key((*it)->cfg, key) && (*it)->loose == q.front()->loose) { B.push_back(*it); it = q.erase(it); } else ++it; }
            lk.unlock();
            run(B);
            lk.lock();
            for (EncJob* j : B) j->done = true;
            batches++; jobs += B.size();
            gather_us = (B.size() > 1 || !q.empty()) ? 400 : gather_us / 2;
            running = false;
            cv.notify_all();
        }
    }
};

void Dispatcher::run(std::vector<EncJob*>& B) {
    auto fail_all = [&](int rc) { for (EncJob* j : B) j->rc = rc; };
    flacb200_ctx* c = context();
    if (!c) { fail_all(1); return; }
    cudaSetDevice(device);
    flacb200_enc_config cfg = B[0]->cfg;
    const uint32_t ch = cfg.channels;
    const bool narrow = cfg.bits_per_sample <= 16;            // int16 container: half the H2D bytes, and 16-bit stereo takes the TMA-staged kernels
    cfg.container_bytes = narrow ? 2u : 4u;
    const size_t ns = B.size();
    std::vector<uint64_t> off(ns), cnt(ns); std::vector<uint32_t> ffn(ns); std::vector<uint8_t> prev(ns);
    uint64_t elems = 0;
    for (size_t s = 0; s < ns; s++) { off[s] = elems; cnt[s] = B[s]->n; ffn[s] = B[s]->ffn; prev[s] = B[s]->prev_ca; elems += B[s]->n * ch; }
    const size_t bytes = (size_t)elems * cfg.container_bytes;
    if (bytes > h_pcm_cap) {
        if (h_pcm) cudaFreeHost(h_pcm);
        h_pcm = nullptr; h_pcm_cap = 0;
        const size_t want = bytes + bytes / 4 + 4096;
        if (cudaHostAlloc(&h_pcm, want, cudaHostAllocDefault) != cudaSuccess) { fail_all(1); return; }
        h_pcm_cap = want;
    }
    for (size_t s = 0; s < ns; s++) {
        const size_t n = (size_t)B[s]->n * ch;
        if (narrow) { int16_t* d = (int16_t*)h_pcm + off[s]; const int32_t* x = B[s]->pcm; for (size_t i = 0; i < n; i++) d[i] = (int16_t)x[i]; }
        else memcpy((int32_t*)h_pcm + off[s], B[s]->pcm, n * 4);
    }
    if (B[0]->loose && flacb200_encode_set_prev_assignment(c, prev.data(), (uint32_t)ns) != 0) { fail_all(1); return; }
    if (flacb200_encode_batch(c, &cfg, h_pcm, 0, elems, (uint32_t)ns, off.data(), cnt.data(), ffn.data()) != 0) { fail_all(1); return; }
    flacb200_enc_result r;
    if (flacb200_encode_result(c, &r) != 0) { fail_all(1); return; }
    arena.resize(r.total_bytes ? r.total_bytes : 1); foff.resize(r.n_frames); flen.resize(r.n_frames); fsmp.resize(r.n_frames); fstream.resize(r.n_frames);
    if (flacb200_encode_fetch(c, arena.data(), arena.size(), foff.data(), flen.data(), fsmp.data(), fstream.data(), nullptr) != 0) { fail_all(1); return; }
    if (B[0]->loose && r.n_frames) { ca.resize(r.n_frames); if (flacb200_encode_fetch_assignments(c, ca.data(), ca.size()) != 0) { fail_all(1); return; } }
    // frames are ordered by stream: hand every job its own
    uint32_t f = 0;
    for (size_t s = 0; s < ns; s++) {
        EncJob* j = B[s];
        j->flen.clear(); j->fsmp.clear(); j->bytes.clear();
        const uint32_t f0 = f;
        while (f < r.n_frames && fstream[f] == s) f++;
        if (f > f0) {
            const uint64_t b0 = foff[f0], b1 = foff[f - 1] + flen[f - 1];
            j->bytes.assign(arena.begin() + (long)b0, arena.begin() + (long)b1);
            j->flen.assign(flen.begin() + f0, flen.begin() + f); j->fsmp.assign(fsmp.begin() + f0, fsmp.begin() + f);
            if (B[0]->loose) j->last_ca = ca[f - 1];
        }
    }
    // up: verify_write_callback_ / FLAC__stream_encoder_set_verify (stream_encoder.h:781-801): every frame is decoded again (here:
    // all verifying jobs of the batch in one GPU decode, headerless mode) and compared with the input before it is delivered
    std::vector<EncJob*> V;
    for (EncJob* j : B) if (j->verify && !j->flen.empty()) V.push_back(j);
    if (!V.empty()) {
        std::vector<uint8_t> blob; std::vector<uint64_t> boff(V.size()), blen(V.size());
        for (size_t v = 0; v < V.size(); v++) { boff[v] = blob.size(); blen[v] = V[v]->bytes.size(); blob.insert(blob.end(), V[v]->bytes.begin(), V[v]->bytes.end()); }
        blob.resize(blob.size() + 16, 0);
        flacb200_dec_raw_params rp; memset(&rp, 0, sizeof rp); rp.sample_rate = cfg.sample_rate; rp.channels = ch; rp.bits_per_sample = cfg.bits_per_sample; rp.fixed_blocksize = cfg.blocksize;
        std::vector<flacb200_dec_stream_info> di(V.size());
        uint64_t tot = 0; for (EncJob* j : V) tot += j->n * ch;
        std::vector<int32_t> back((size_t)tot + 1);
        if (flacb200_decode_batch(c, blob.data(), 0, blob.size() - 16, (uint32_t)V.size(), boff.data(), blen.data(), 4, &rp) != 0 ||
            flacb200_decode_fetch(c, back.data(), back.size() * 4, di.data(), nullptr, 0) != 0) { for (EncJob* j : V) j->rc = 2; return; }
        for (size_t v = 0; v < V.size(); v++) {
            EncJob* j = V[v];
            if (di[v].status != 0) { j->rc = 2; continue; }
            const int32_t* got = back.data() + di[v].pcm_off;
            bool same = di[v].total_samples == j->n;
            const size_t n = (size_t)j->n * ch;
            for (size_t i = 0; same && i < n; i++)
                if (got[i] != j->pcm[i]) { same = false; j->v_sample = i / ch; j->v_channel = (uint32_t)(i % ch); j->v_expected = j->pcm[i]; j->v_got = got[i]; }
            if (!same) j->rc = 3;
        }
    }
}

std::mutex g_disp_mu;
std::vector<Dispatcher*> g_disp;               // index = CUDA device
std::atomic<unsigned> g_rr{0};

// device for a new handle: FLACB200_DEVICE=<i>, or round robin over FLACB200_DEVICES=<i,j,...>; default 0
int pick_device() {
    if (const char* l = getenv("FLACB200_DEVICES")) {
        std::vector<int> d; const char* p = l;
        while (*p) { char* e; const long v = strtol(p, &e, 10); if (e == p) break; d.push_back((int)v); p = (*e == ',') ? e + 1 : e; }
        if (!d.empty()) return d[g_rr.fetch_add(1) % d.size()];
    }
    if (const char* l = getenv("FLACB200_DEVICE")) return atoi(l);
    return 0;
}
Dispatcher* dispatcher(int device) {
    std::lock_guard<std::mutex> lk(g_disp_mu);
    if (device < 0) device = 0;
    if ((size_t)device >= g_disp.size()) g_disp.resize((size_t)device + 1, nullptr);
    if (!g_disp[device]) { g_disp[device] = new Dispatcher(); g_disp[device]->device = device; }
    return g_disp[device];
}

struct EncImpl {
    // settings (libFLAC defaults: stream_encoder.h set_defaults_: 2 ch, 16 bit, 44.1 kHz, level 5)
    FLAC__bool verify = 0, streamable_subset = 1, limit_min_bitrate = 0;
    uint32_t channels = 2, bps = 16, sample_rate = 44100, level = 5, blocksize = 0;
    uint64_t total_samples_estimate = 0;
    // fine-grained settings (stream_encoder.h:  set_do_mid_side_stereo ... set_apodization): set_compression_level writes the level's
    // presets into them, the individual setters change them afterwards, exactly as in libFLAC
    FLAC__bool ms = 1, loose = 0, prec_search = 0, exhaustive = 0;
    uint32_t max_lpc = 8, qlp_prec = 0, min_po = 0, max_po = 5, rice_dist = 0, apod_parts = 1;
    uint32_t qlp_resolved = 0;             // what get_qlp_coeff_precision reports while the encoder is initialised
    float apod_p = 0.5f;
    bool custom_tuning = false;            // a setting this build cannot honour (exhaustive searches, other window families): init fails loudly
    int state = ST_UNINITIALIZED;
    // session
    FLAC__StreamEncoderWriteCallback write_cb = nullptr; FLAC__StreamEncoderSeekCallback seek_cb = nullptr;
    FLAC__StreamEncoderTellCallback tell_cb = nullptr; FLAC__StreamEncoderMetadataCallback meta_cb = nullptr;
    FLAC__StreamEncoderProgressCallback progress_cb = nullptr;
    void* client = nullptr;
    FILE* file = nullptr; bool own_file = false;
    uint32_t N = 0;                        // resolved blocksize
    std::vector<int32_t> pending;          // interleaved samples not yet framed
    Dispatcher* disp = nullptr; bool counted = false;      // this handle's device queue; counted in disp->active while initialised
    uint32_t frame_number = 0;
    uint8_t last_ca = 0;                   // loose mid/side: channel assignment of the last frame delivered
    uint64_t vstat_sample = 0; uint32_t vstat_channel = 0; int32_t vstat_expected = 0, vstat_got = 0;   // verify mismatch report
    uint64_t samples_written = 0, bytes_written = 0;
    uint32_t min_fs = 0, max_fs = 0, frames_written = 0;
    bool being_deleted = false;            // FLAC__stream_encoder_delete on an encoder that was not finished: finish() without flush or callbacks
    uint64_t streaminfo_offset = 0;        // byte position of the STREAMINFO block header as told by the tell callback
    Md5 md5;
    std::vector<uint8_t> arena; std::vector<uint64_t> foff; std::vector<uint32_t> flen, fsmp;
};

struct Handle { FLAC__StreamEncoder pub; EncImpl impl; };
inline EncImpl* I(const FLAC__StreamEncoder* e) { return e ? (EncImpl*)e->private_ : nullptr; }

// level presets (stream_encoder.h:845-853)
static const struct { int ms, loose; uint32_t lpc, po, parts; } kLv[9] = {{0,0,0,3,1},{1,1,0,3,1},{1,0,0,3,1},{0,0,6,4,1},{1,1,8,4,1},{1,0,8,5,1},{1,0,8,6,2},{1,0,12,6,2},{1,0,12,6,3}};
void apply_level(EncImpl* m, uint32_t v) {
    m->level = v > 8 ? 8 : v;
    m->ms = kLv[m->level].ms; m->loose = kLv[m->level].loose; m->max_lpc = kLv[m->level].lpc; m->max_po = kLv[m->level].po;
    m->apod_parts = kLv[m->level].parts; m->apod_p = 0.5f;
    m->qlp_prec = 0; m->qlp_resolved = 0; m->min_po = 0; m->rice_dist = 0; m->prec_search = 0; m->exhaustive = 0; m->custom_tuning = false;
}
void reset_settings(EncImpl* m) {
    m->verify = 0; m->streamable_subset = 1; m->limit_min_bitrate = 0;
    m->channels = 2; m->bps = 16; m->sample_rate = 44100; m->blocksize = 0;
    m->total_samples_estimate = 0;
    apply_level(m, 5);
}
void fill_tuning(const EncImpl* m, flacb200_enc_config* cfg) {
    cfg->tune = 1; cfg->do_mid_side = m->ms ? 1u : 0u; cfg->loose_mid_side = m->loose ? 1u : 0u; cfg->max_lpc_order = m->max_lpc;
    cfg->qlp_coeff_precision = m->qlp_prec; cfg->max_residual_partition_order = m->max_po; cfg->apod_parts = m->apod_parts; cfg->apod_p = m->apod_p;
}

// ref: format.h:546-557 -- the 34-byte STREAMINFO body
void put_streaminfo(uint8_t* p, const EncImpl* m, uint32_t min_fs, uint32_t max_fs, uint64_t total, const uint8_t md5[16]) {
    p[0] = (uint8_t)(m->N >> 8); p[1] = (uint8_t)m->N; p[2] = p[0]; p[3] = p[1];
    p[4] = (uint8_t)(min_fs >> 16); p[5] = (uint8_t)(min_fs >> 8); p[6] = (uint8_t)min_fs;
    p[7] = (uint8_t)(max_fs >> 16); p[8] = (uint8_t)(max_fs >> 8); p[9] = (uint8_t)max_fs;
    const uint64_t v = ((uint64_t)m->sample_rate << 44) | ((uint64_t)(m->channels - 1) << 41) | ((uint64_t)(m->bps - 1) << 36) | (total & 0xFFFFFFFFFull);
    for (int i = 0; i < 8; i++) p[10 + i] = (uint8_t)(v >> (56 - 8 * i));
    memcpy(p + 18, md5, 16);
}

// deliver bytes: user callback (stream mode) or stdio (file mode); bookkeeping as libFLAC's write_frame_
bool deliver(FLAC__StreamEncoder* e, const uint8_t* buf, size_t bytes, uint32_t samples, uint32_t frame) {
    EncImpl* m = I(e);
    if (m->file) {
        if (fwrite(buf, 1, bytes, m->file) != bytes) { m->state = ST_IO_ERROR; return false; }
        if (samples > 0 && m->progress_cb)
            m->progress_cb(e, m->bytes_written + bytes, m->samples_written + samples, m->frames_written + 1, 0, m->client);
    } else {
        // up: write_frame_ -- tell (if given) before every write; remember where the STREAMINFO block went by
        if (m->tell_cb) {
            FLAC__uint64 posn = 0;
            const int ts = m->tell_cb(e, &posn, m->client);
            if (ts == 1) { m->state = ST_CLIENT_ERROR; return false; }           // TELL_STATUS_ERROR
            if (ts == 0 && samples == 0 && bytes > 0 && (buf[0] & 0x7f) == 0 && m->streaminfo_offset == 0) m->streaminfo_offset = posn;
        }
        if (m->write_cb(e, buf, bytes, samples, frame, m->client) != 0) { m->state = ST_CLIENT_ERROR; return false; }
    }
    m->bytes_written += bytes;
    if (samples > 0) {
        m->samples_written += samples; m->frames_written++;
        if (m->min_fs == 0 || bytes < m->min_fs) m->min_fs = (uint32_t)bytes;
        if (bytes > m->max_fs) m->max_fs = (uint32_t)bytes;
    }
    return true;
}

// encode `n` pending inter-channel samples (whole frames, or everything at finish): one job for the device's dispatcher,
// which batches it with whatever the other handles have submitted; then MD5 and callbacks on this thread, no lock held
bool encode_pending(FLAC__StreamEncoder* e, uint64_t n) {
    EncImpl* m = I(e);
    if (n == 0) return true;
    if (!m->disp) { m->state = ST_MEMORY_ALLOCATION_ERROR; return false; }
    EncJob job;
    memset(&job.cfg, 0, sizeof job.cfg);
    job.cfg.sample_rate = m->sample_rate; job.cfg.channels = m->channels; job.cfg.bits_per_sample = m->bps;
    job.cfg.compression_level = m->level; job.cfg.blocksize = m->N; job.cfg.container_bytes = 4;
    job.cfg.write_prologue = 0; job.cfg.do_md5 = 0; job.cfg.streamable_subset = 0; job.cfg.debug_trace = 0;
    job.cfg.limit_min_bitrate = m->limit_min_bitrate ? 1u : 0u;
    fill_tuning(m, &job.cfg);
    job.pcm = m->pending.data(); job.n = n; job.ffn = m->frame_number; job.prev_ca = m->last_ca;
    job.loose = m->loose && m->ms && m->channels == 2;
    job.verify = m->verify != 0;
    m->disp->submit(&job);
    if (job.rc == 1) { m->state = ST_FRAMING_ERROR; return false; }
    // MD5 over (bps+7)/8 little-endian bytes per sample, interleaved (up: md5.c FLAC__MD5Accumulate)
    {
        const uint32_t bytes = (m->bps + 7) / 8;
        const size_t vals = (size_t)n * m->channels;
        std::vector<uint8_t> tmp(vals * bytes);
        size_t k = 0;
        if (bytes == 2) for (size_t i = 0; i < vals; i++) { const uint32_t v = (uint32_t)m->pending[i]; tmp[k++] = (uint8_t)v; tmp[k++] = (uint8_t)(v >> 8); }
        else for (size_t i = 0; i < vals; i++) { const uint32_t v = (uint32_t)m->pending[i]; for (uint32_t b = 0; b < bytes; b++) tmp[k++] = (uint8_t)(v >> (8 * b)); }
        m->md5.update(tmp.data(), tmp.size());
    }
    if (job.rc == 2) { m->state = ST_VERIFY_DECODER_ERROR; return false; }
    if (job.rc == 3) {
        m->vstat_sample = m->samples_written + job.v_sample; m->vstat_channel = job.v_channel; m->vstat_expected = job.v_expected; m->vstat_got = job.v_got;
        m->state = ST_VERIFY_MISMATCH; return false;
    }
    if (job.loose && !job.flen.empty()) m->last_ca = job.last_ca;
    size_t at = 0;
    for (size_t f = 0; f < job.flen.size(); f++) {
        if (!deliver(e, job.bytes.data() + at, job.flen[f], job.fsmp[f], m->frame_number)) return false;
        at += job.flen[f];
        m->frame_number++;
    }
    m->pending.erase(m->pending.begin(), m->pending.begin() + (size_t)n * m->channels);
    return true;
}

int init_common(FLAC__StreamEncoder* e) {
    EncImpl* m = I(e);
    flacb200_enc_config cfg{};
    cfg.sample_rate = m->sample_rate; cfg.channels = m->channels; cfg.bits_per_sample = m->bps; cfg.compression_level = m->level;
    cfg.blocksize = m->blocksize; cfg.container_bytes = 4; cfg.streamable_subset = (uint32_t)m->streamable_subset;
    fill_tuning(m, &cfg);
    const int st = flacb200_enc_validate(&cfg);
    if (st != 0) return st;
    m->N = m->blocksize ? m->blocksize : (m->max_lpc == 0 ? 1152u : 4096u);
    // up: init_stream_internal_ -- mid/side is a two-channel matter, loose mid/side a mid/side matter: the getters say so from here on
    if (m->channels != 2) { m->ms = 0; m->loose = 0; } else if (!m->ms) m->loose = 0;
    // up: init_stream_internal_ resolves a precision of 0 ("let the encoder choose") from the sample size and the blocksize, and
    // FLAC__stream_encoder_get_qlp_coeff_precision reports that value from then on (the kernels resolve it the same way: engine.cu)
    m->qlp_resolved = m->qlp_prec ? m->qlp_prec
                    : m->bps < 16 ? (2 + m->bps / 2 < 5 ? 5u : 2 + m->bps / 2)
                    : m->bps == 16 ? (m->N <= 192 ? 7u : m->N <= 384 ? 8u : m->N <= 576 ? 9u : m->N <= 1152 ? 10u : m->N <= 2304 ? 11u : m->N <= 4608 ? 12u : 13u)
                    : (m->N <= 384 ? 13u : m->N <= 1152 ? 14u : 15u);
    // limits of this build fail loudly here instead of producing a different stream (DESIGN.md "limits"): searches libFLAC's presets
    // never run (exhaustive model / precision search, a minimum partition order), window families other than tukey, orders above 12
    if (m->custom_tuning || m->exhaustive || m->prec_search || (m->min_po != 0 && m->max_po != 0) || m->max_lpc > 12 || m->max_po > 6 || m->qlp_prec > 15 ||
        m->apod_parts < 1 || m->apod_parts > 3) { m->state = ST_FRAMING_ERROR; return INIT_ENCODER_ERROR; }
    {
        Dispatcher* d = dispatcher(pick_device());
        std::lock_guard<std::mutex> lk(d->mu);
        if (!d->context()) { m->state = ST_MEMORY_ALLOCATION_ERROR; return INIT_ENCODER_ERROR; }   // no CUDA device: no CPU fallback
        m->disp = d;
        if (!m->counted) { d->active.fetch_add(1); m->counted = true; }
    }
    m->pending.clear(); m->frame_number = 0; m->last_ca = 0; m->samples_written = 0; m->bytes_written = 0; m->min_fs = m->max_fs = 0; m->frames_written = 0; m->streaminfo_offset = 0;
    m->md5.init();
    m->state = ST_OK;
    // stream prologue: "fLaC", STREAMINFO (frame sizes / MD5 zero, total = estimate), VORBIS_COMMENT
    uint8_t b[64];
    memcpy(b, "fLaC", 4);
    if (!deliver(e, b, 4, 0, 0)) return INIT_ENCODER_ERROR;
    uint8_t zero[16] = {0};
    b[0] = 0x00; b[1] = 0; b[2] = 0; b[3] = 34;
    put_streaminfo(b + 4, m, 0, 0, m->total_samples_estimate, zero);
    if (!deliver(e, b, 38, 0, 0)) return INIT_ENCODER_ERROR;
    const uint32_t vl = (uint32_t)strlen(FLAC__VENDOR_STRING), len = 4 + vl + 4;
    b[0] = 0x84; b[1] = (uint8_t)(len >> 16); b[2] = (uint8_t)(len >> 8); b[3] = (uint8_t)len;
    b[4] = (uint8_t)vl; b[5] = (uint8_t)(vl >> 8); b[6] = (uint8_t)(vl >> 16); b[7] = (uint8_t)(vl >> 24);
    memcpy(b + 8, FLAC__VENDOR_STRING, vl);
    memset(b + 8 + vl, 0, 4);
    if (!deliver(e, b, 4 + len, 0, 0)) return INIT_ENCODER_ERROR;
    // up: init_stream_internal_ asks for the position once more after the metadata (audio_offset); observed on the reference binary
    if (!m->file && m->tell_cb) { FLAC__uint64 posn = 0; if (m->tell_cb(e, &posn, m->client) == 1) { m->state = ST_CLIENT_ERROR; return INIT_ENCODER_ERROR; } }
    return INIT_OK;
}

}  // namespace

extern "C" {

// batches / jobs the encoder dispatcher of `device` has run so far (tests: cross-object batching really happens)
int flacb200_dispatch_stats(int device, uint64_t* batches, uint64_t* jobs) {
    Dispatcher* d = dispatcher(device);
    std::lock_guard<std::mutex> lk(d->mu);
    if (batches) *batches = d->batches;
    if (jobs) *jobs = d->jobs;
    return 0;
}

FLAC__StreamEncoder* FLAC__stream_encoder_new(void) {
    Handle* h = new Handle();
    h->pub.protected_ = nullptr; h->pub.private_ = &h->impl;
    return &h->pub;
}
void FLAC__stream_encoder_delete(FLAC__StreamEncoder* e) {
    if (!e) return;
    // up: FLAC__stream_encoder_delete (pinned on the binary, tools/host_logic_check.py): an encoder that is deleted without finish()
    // is torn down without a word -- no last frame, no STREAMINFO rewrite, no callback of any kind (a Python object may be collected
    // long after its callbacks stopped making sense)
    I(e)->being_deleted = true;
    if (I(e)->state != ST_UNINITIALIZED) FLAC__stream_encoder_finish(e);
    delete reinterpret_cast<Handle*>(e);
}

#define SETTER(name, type, field) FLAC__bool FLAC__stream_encoder_set_##name(FLAC__StreamEncoder* e, type v) { \
    EncImpl* m = I(e); if (!m || m->state != ST_UNINITIALIZED) return 0; m->field = v; return 1; }
SETTER(verify, FLAC__bool, verify)
SETTER(channels, uint32_t, channels)
SETTER(bits_per_sample, uint32_t, bps)
SETTER(sample_rate, uint32_t, sample_rate)
SETTER(blocksize, uint32_t, blocksize)
SETTER(streamable_subset, FLAC__bool, streamable_subset)
SETTER(limit_min_bitrate, FLAC__bool, limit_min_bitrate)
#undef SETTER
// STREAMINFO holds total_samples in 36 bits; libFLAC clamps the estimate on the way in (ref: format.h:536 FLAC__STREAM_METADATA_STREAMINFO_TOTAL_SAMPLES_LEN)
FLAC__bool FLAC__stream_encoder_set_total_samples_estimate(FLAC__StreamEncoder* e, FLAC__uint64 v) {
    EncImpl* m = I(e); if (!m || m->state != ST_UNINITIALIZED) return 0;
    const FLAC__uint64 lim = (1ull << 36) - 1;
    m->total_samples_estimate = v < lim ? v : lim; return 1;
}
FLAC__bool FLAC__stream_encoder_set_compression_level(FLAC__StreamEncoder* e, uint32_t v) {
    EncImpl* m = I(e); if (!m || m->state != ST_UNINITIALIZED) return 0;
    apply_level(m, v); return 1;
}
// tuning knobs outside pyFLAC's surface: remember that the presets were left, init then refuses (fails loudly)
#define SETTER(name, type, stmt) FLAC__bool FLAC__stream_encoder_set_##name(FLAC__StreamEncoder* e, type v) { \
    EncImpl* m = I(e); if (!m || m->state != ST_UNINITIALIZED) return 0; stmt; return 1; }
SETTER(do_mid_side_stereo, FLAC__bool, m->ms = v)
SETTER(loose_mid_side_stereo, FLAC__bool, m->loose = v)
SETTER(max_lpc_order, uint32_t, m->max_lpc = v)
SETTER(qlp_coeff_precision, uint32_t, m->qlp_prec = v)
SETTER(do_qlp_coeff_prec_search, FLAC__bool, m->prec_search = v)
SETTER(do_exhaustive_model_search, FLAC__bool, m->exhaustive = v)
SETTER(min_residual_partition_order, uint32_t, m->min_po = v)
SETTER(max_residual_partition_order, uint32_t, m->max_po = v)
SETTER(rice_parameter_search_dist, uint32_t, (void)v)                    // accepted and dropped, as libFLAC 1.4.3 does (the getter keeps reading 0)
#undef SETTER
// up: FLAC__stream_encoder_set_apodization.  One window of the tukey family is what this build runs: "tukey(P)" and
// "subdivide_tukey(N[/P])" (the two every compression level uses); a list, or any other family, makes init fail loudly.
FLAC__bool FLAC__stream_encoder_set_apodization(FLAC__StreamEncoder* e, const char* spec) {
    EncImpl* m = I(e);
    if (!m || m->state != ST_UNINITIALIZED || !spec) return 0;
    m->custom_tuning = false;
    const size_t n = strlen(spec);
    if (strchr(spec, ';') && strchr(spec, ';')[1] != 0) { m->custom_tuning = true; return 1; }
    if (n > 7 && strncmp(spec, "tukey(", 6) == 0) {
        const float p = (float)strtod(spec + 6, nullptr);
        m->apod_parts = 1; m->apod_p = (p >= 0.0f && p <= 1.0f) ? p : 0.5f;        // out of range: libFLAC falls back to its default tukey(0.5)
    } else if (n > 16 && strncmp(spec, "subdivide_tukey(", 16) == 0) {
        const int parts = (int)strtod(spec + 16, nullptr);
        if (parts > 1) {
            const char* sl = strchr(spec, '/');
            float p = sl ? (float)strtod(sl + 1, nullptr) : 0.5f;
            if (p > 1.0f) p = 1.0f; else if (p < 0.0f) p = 0.0f;
            m->apod_parts = (uint32_t)parts; m->apod_p = p;
        } else { m->apod_parts = 1; m->apod_p = 0.5f; }
    } else m->custom_tuning = true;
    return 1;
}

int FLAC__stream_encoder_get_state(const FLAC__StreamEncoder* e) { return I(e) ? I(e)->state : ST_UNINITIALIZED; }
const char* FLAC__stream_encoder_get_resolved_state_string(const FLAC__StreamEncoder* e) { return FLAC__StreamEncoderStateString[FLAC__stream_encoder_get_state(e)]; }
void FLAC__stream_encoder_get_verify_decoder_error_stats(const FLAC__StreamEncoder* e, FLAC__uint64* a, uint32_t* f, uint32_t* c, uint32_t* s, FLAC__int32* x, FLAC__int32* g) {
    const EncImpl* m = I(e);
    const uint32_t N = m && m->N ? m->N : 1u;
    if (a) *a = m ? m->vstat_sample : 0; if (f) *f = m ? (uint32_t)(m->vstat_sample / N) : 0; if (c) *c = m ? m->vstat_channel : 0;
    if (s) *s = m ? (uint32_t)(m->vstat_sample % N) : 0; if (x) *x = m ? m->vstat_expected : 0; if (g) *g = m ? m->vstat_got : 0;
}
FLAC__bool FLAC__stream_encoder_get_verify(const FLAC__StreamEncoder* e) { return I(e)->verify; }
FLAC__bool FLAC__stream_encoder_get_streamable_subset(const FLAC__StreamEncoder* e) { return I(e)->streamable_subset; }
uint32_t FLAC__stream_encoder_get_channels(const FLAC__StreamEncoder* e) { return I(e)->channels; }
uint32_t FLAC__stream_encoder_get_bits_per_sample(const FLAC__StreamEncoder* e) { return I(e)->bps; }
uint32_t FLAC__stream_encoder_get_sample_rate(const FLAC__StreamEncoder* e) { return I(e)->sample_rate; }
uint32_t FLAC__stream_encoder_get_blocksize(const FLAC__StreamEncoder* e) { return I(e)->state == ST_UNINITIALIZED ? I(e)->blocksize : I(e)->N; }
FLAC__bool FLAC__stream_encoder_get_do_mid_side_stereo(const FLAC__StreamEncoder* e) { return I(e)->ms; }
FLAC__bool FLAC__stream_encoder_get_loose_mid_side_stereo(const FLAC__StreamEncoder* e) { return I(e)->loose; }
uint32_t FLAC__stream_encoder_get_max_lpc_order(const FLAC__StreamEncoder* e) { return I(e)->max_lpc; }
uint32_t FLAC__stream_encoder_get_qlp_coeff_precision(const FLAC__StreamEncoder* e) { const EncImpl* m = I(e); return m->state != ST_UNINITIALIZED && m->qlp_resolved ? m->qlp_resolved : m->qlp_prec; }
FLAC__bool FLAC__stream_encoder_get_do_qlp_coeff_prec_search(const FLAC__StreamEncoder* e) { return I(e)->prec_search; }
FLAC__bool FLAC__stream_encoder_get_do_escape_coding(const FLAC__StreamEncoder*) { return 0; }
FLAC__bool FLAC__stream_encoder_get_do_exhaustive_model_search(const FLAC__StreamEncoder* e) { return I(e)->exhaustive; }
uint32_t FLAC__stream_encoder_get_min_residual_partition_order(const FLAC__StreamEncoder* e) { return I(e)->min_po; }
uint32_t FLAC__stream_encoder_get_max_residual_partition_order(const FLAC__StreamEncoder* e) { return I(e)->max_po; }
uint32_t FLAC__stream_encoder_get_rice_parameter_search_dist(const FLAC__StreamEncoder* e) { return I(e)->rice_dist; }
FLAC__uint64 FLAC__stream_encoder_get_total_samples_estimate(const FLAC__StreamEncoder* e) { return I(e)->total_samples_estimate; }
FLAC__bool FLAC__stream_encoder_get_limit_min_bitrate(const FLAC__StreamEncoder* e) { return I(e)->limit_min_bitrate; }

int FLAC__stream_encoder_init_stream(FLAC__StreamEncoder* e, FLAC__StreamEncoderWriteCallback w, FLAC__StreamEncoderSeekCallback s,
                                     FLAC__StreamEncoderTellCallback t, FLAC__StreamEncoderMetadataCallback mcb, void* client) {
    EncImpl* m = I(e);
    if (m->state != ST_UNINITIALIZED) return INIT_ALREADY_INITIALIZED;
    if (!w || (s && !t)) return INIT_INVALID_CALLBACKS;          // up: init_stream_internal_ (tests/test_encoder.py:202-207)
    m->write_cb = w; m->seek_cb = s; m->tell_cb = t; m->meta_cb = mcb; m->progress_cb = nullptr; m->client = client; m->file = nullptr; m->own_file = false;
    return init_common(e);
}
int FLAC__stream_encoder_init_FILE(FLAC__StreamEncoder* e, FILE* f, FLAC__StreamEncoderProgressCallback p, void* client) {
    EncImpl* m = I(e);
    if (m->state != ST_UNINITIALIZED) return INIT_ALREADY_INITIALIZED;
    if (!f) { m->state = ST_IO_ERROR; return INIT_ENCODER_ERROR; }
    m->write_cb = nullptr; m->seek_cb = nullptr; m->tell_cb = nullptr; m->meta_cb = nullptr; m->progress_cb = p; m->client = client; m->file = f; m->own_file = true;
    const int rc = init_common(e);
    if (rc != INIT_OK && m->file) { fclose(m->file); m->file = nullptr; }
    return rc;
}
int FLAC__stream_encoder_init_file(FLAC__StreamEncoder* e, const char* filename, FLAC__StreamEncoderProgressCallback p, void* client) {
    EncImpl* m = I(e);
    if (m->state != ST_UNINITIALIZED) return INIT_ALREADY_INITIALIZED;
    FILE* f = filename ? fopen(filename, "w+b") : stdout;
    if (!f) { m->state = ST_IO_ERROR; return INIT_ENCODER_ERROR; }
    return FLAC__stream_encoder_init_FILE(e, f, p, client);
}
int FLAC__stream_encoder_init_ogg_stream(FLAC__StreamEncoder*, FLAC__StreamEncoderReadCallback, FLAC__StreamEncoderWriteCallback, FLAC__StreamEncoderSeekCallback,
                                         FLAC__StreamEncoderTellCallback, FLAC__StreamEncoderMetadataCallback, void*) { return INIT_UNSUPPORTED_CONTAINER; }
int FLAC__stream_encoder_init_ogg_FILE(FLAC__StreamEncoder*, FILE*, FLAC__StreamEncoderProgressCallback, void*) { return INIT_UNSUPPORTED_CONTAINER; }
int FLAC__stream_encoder_init_ogg_file(FLAC__StreamEncoder*, const char*, FLAC__StreamEncoderProgressCallback, void*) { return INIT_UNSUPPORTED_CONTAINER; }

FLAC__bool FLAC__stream_encoder_process_interleaved(FLAC__StreamEncoder* e, const FLAC__int32 buffer[], uint32_t samples) {
    EncImpl* m = I(e);
    if (!m || m->state != ST_OK) return 0;
    m->pending.insert(m->pending.end(), buffer, buffer + (size_t)samples * m->channels);
    const uint64_t have = m->pending.size() / m->channels;
    // a frame needs blocksize+1 buffered samples (the over-read sample stays pending): SURVEY A.2
    if (have > m->N) {
        const uint64_t frames = (have - 1) / m->N;
        if (!encode_pending(e, frames * m->N)) return 0;
    }
    return 1;
}
FLAC__bool FLAC__stream_encoder_process(FLAC__StreamEncoder* e, const FLAC__int32* const buffer[], uint32_t samples) {
    EncImpl* m = I(e);
    if (!m || m->state != ST_OK) return 0;
    std::vector<int32_t> tmp((size_t)samples * m->channels);
    for (uint32_t i = 0; i < samples; i++) for (uint32_t c = 0; c < m->channels; c++) tmp[(size_t)i * m->channels + c] = buffer[c][i];
    return FLAC__stream_encoder_process_interleaved(e, tmp.data(), samples);
}

FLAC__bool FLAC__stream_encoder_finish(FLAC__StreamEncoder* e) {
    EncImpl* m = I(e);
    if (!m || m->state == ST_UNINITIALIZED) return 1;
    // up: FLAC__stream_encoder_finish (pinned on the binary, tools/host_logic_check.py and tests/test_gpu_zz_host_logic.py): only what goes
    // wrong INSIDE this call makes it fail and leaves the error state standing -- the last frame, the STREAMINFO rewrite.  An
    // encoder that was already in an error state (a process() call failed) is simply reset: true, UNINITIALIZED.
    bool ok = true;
    if (m->state == ST_OK && !m->being_deleted) {
        const uint64_t have = m->pending.size() / m->channels;
        if (have > 0) ok = encode_pending(e, have);           // remainder -> final (possibly short) frame
    }
    uint8_t digest[16];
    m->md5.final(digest);
    if (m->state == ST_OK && !m->being_deleted) {
        // up: update_metadata_ -- rewrite STREAMINFO in place: MD5 @26 (16 B), total samples @21 (5 B), frame sizes @12 (6 B)
        uint8_t si[34];
        const uint32_t min_fs = m->min_fs ? m->min_fs : 0xFFFFFFu;      // no frame at all: libFLAC's minimum is still where it started (2^24 - 1)
        put_streaminfo(si, m, min_fs, m->max_fs, m->samples_written, digest);
        // offsets are relative to the STREAMINFO block header (4 when the stream starts at byte 0): +22, +17, +8
        const uint64_t so = m->file ? 4 : (m->streaminfo_offset ? m->streaminfo_offset : 4);
        struct { uint64_t off; const uint8_t* p; size_t n; } patch[3] = {{so + 22, si + 18, 16}, {so + 17, si + 13, 5}, {so + 8, si + 4, 6}};
        if (m->file) {
            for (auto& q : patch) { if (fseek(m->file, (long)q.off, SEEK_SET) != 0 || fwrite(q.p, 1, q.n, m->file) != q.n) { ok = false; m->state = ST_IO_ERROR; break; } }
            fseek(m->file, 0, SEEK_END);
        } else if (m->seek_cb) {
            for (auto& q : patch) {
                if (m->seek_cb(e, q.off, m->client) != 0) { ok = false; m->state = ST_CLIENT_ERROR; break; }
                if (m->write_cb(e, q.p, q.n, 0, 0, m->client) != 0) { ok = false; m->state = ST_CLIENT_ERROR; break; }
            }
        }
        if (m->meta_cb) {
            FLAC__StreamMetadata md; memset(&md, 0, sizeof md);
            md.type = 0; md.is_last = 0; md.length = 34;
            md.data.stream_info.min_blocksize = md.data.stream_info.max_blocksize = m->N;
            md.data.stream_info.min_framesize = min_fs; md.data.stream_info.max_framesize = m->max_fs;
            md.data.stream_info.sample_rate = m->sample_rate; md.data.stream_info.channels = m->channels;
            md.data.stream_info.bits_per_sample = m->bps; md.data.stream_info.total_samples = m->samples_written;
            memcpy(md.data.stream_info.md5sum, digest, 16);
            m->meta_cb(e, &md, m->client);
        }
    }
    if (m->file && m->own_file) { if (m->file != stdout) fclose(m->file); }
    m->file = nullptr;
    m->pending.clear();
    if (m->counted && m->disp) { m->disp->active.fetch_sub(1); m->counted = false; }
    reset_settings(m);                                         // stream_encoder.h:225-227: back to defaults
    if (ok) m->state = ST_UNINITIALIZED;
    return ok ? 1 : 0;
}

}  // extern "C"
