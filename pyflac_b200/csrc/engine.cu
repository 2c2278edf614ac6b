// engine.cu -- host side of the batch encoder: settings resolution, frame slicing, window tables,
// device buffers, kernel sequencing.  C ABI in include/flacb200.h.
//
// Pipeline per batch (all on one CUDA stream, MD5 forked onto the output set's side stream):
//   [H2D pcm if host] -> autoc -> analyze -> pack -> scan_kernel -> compact_kernel -> finalize_kernel
//                     \-> md5_kernel ------------------------------------------------> md5_patch_kernel
// Host -> host: flacb200_encode_batch_host (one synchronous call, chunks of streams pipelined over H2D / kernels / D2H) and
// flacb200_encode_host_submit / _collect (the same with up to three batches in flight).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sched.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/flacb200.h"
#include "fb_common.cuh"
#include "fb_math.cuh"
#include "md5_host.h"
#include "md5_mb.h"

namespace fb {
int launch_analyze(const void*, const FrameDesc*, const float*, const EncParams&, int, SubframePlan*, uint8_t*,
                   SignalDebug*, EncStats*, size_t, void*, cudaStream_t, cudaEvent_t*);
size_t analyze_smem_bytes(const EncParams&);
size_t analyze_work_stride(const EncParams&);
void analyze_layout(EncParams&);
void launch_pack(const void*, const FrameDesc*, const EncParams&, int, const SubframePlan*, const uint8_t*, uint8_t*,
                 uint32_t, uint32_t*, cudaStream_t);
size_t pack_smem_bytes(const EncParams&, uint32_t);
bool fused_eligible(const EncParams&, uint32_t, int);
void launch_autoc_unshifted(const void*, const FrameDesc*, const float*, const EncParams&, int, void*, cudaStream_t);
void launch_fused(const void*, const FrameDesc*, const EncParams&, int, const void*, size_t, SubframePlan*, uint8_t*, EncStats*, uint8_t*, uint32_t,
                  uint32_t*, cudaStream_t, cudaEvent_t);
void launch_md5(const void*, uint32_t, const uint64_t*, const uint64_t*, int, uint32_t, uint32_t, uint8_t*, cudaStream_t);
void launch_md5_gated(const void*, uint32_t, const uint64_t*, const uint64_t*, int, uint32_t, uint32_t, uint8_t*, const Md5Gate&, cudaStream_t);
void launch_layout(const uint32_t*, const FrameDesc*, int, uint32_t, uint64_t, uint64_t*, uint64_t*, cudaStream_t);
void launch_compact(const uint8_t*, uint32_t, const uint32_t*, const uint64_t*, uint8_t*, int, cudaStream_t);
void launch_finalize(const uint32_t*, const uint64_t*, const uint32_t*, const uint32_t*, const uint64_t*, const uint8_t*,
                     int, const EncParams&, uint32_t, uint8_t*, StreamInfoOut*, cudaStream_t);
void launch_md5_patch(const uint8_t*, const uint32_t*, int, uint32_t, uint8_t*, StreamInfoOut*, cudaStream_t);
}  // namespace fb

using namespace fb;

// growable device buffer
struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Host threads of one ctx for the MD5 of host-resident PCM: created once, parked on a condition variable between calls
// (a fresh std::thread per call and worker cost 0.3-0.5 ms of the 12 ms host path and more on an oversubscribed host).
struct HostPool {
    std::vector<std::thread> threads;
    std::mutex mu; std::condition_variable cv_go, cv_done;
    std::function<void()> job; uint64_t gen = 0; int want = 0, running = 0; bool stop = false;
    void start(int n, std::function<void()> fn) {
        std::unique_lock<std::mutex> lk(mu);
        while ((int)threads.size() < n) threads.emplace_back([this, idx = (int)threads.size()] { loop(idx); });
        job = std::move(fn); want = n; running = n; gen++;
        cv_go.notify_all();
    }
    void wait() { std::unique_lock<std::mutex> lk(mu); cv_done.wait(lk, [this] { return running == 0; }); }
    void loop(int idx) {
        uint64_t seen = 0;
        for (;;) {
            std::function<void()> fn;
            { std::unique_lock<std::mutex> lk(mu); cv_go.wait(lk, [&] { return stop || (gen != seen && idx < want); }); if (stop) return; seen = gen; fn = job; }
            fn();
            { std::unique_lock<std::mutex> lk(mu); if (--running == 0) cv_done.notify_all(); }
        }
    }
    ~HostPool() { { std::unique_lock<std::mutex> lk(mu); stop = true; cv_go.notify_all(); } for (auto& t : threads) t.join(); }
};

// Hardware threads this process may use, shared fairly with the other ranks of the node: the affinity mask (a cgroup cpuset
// shows up there; std::thread::hardware_concurrency() reports the whole machine) divided by LOCAL_WORLD_SIZE (torchrun).
static unsigned host_thread_budget() {
    unsigned n = 0;
    cpu_set_t set; CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof set, &set) == 0) n = (unsigned)CPU_COUNT(&set);
    if (n == 0) n = std::thread::hardware_concurrency();
    if (n == 0) n = 8;
    unsigned ranks = 1;
    if (const char* ev = getenv("FLACB200_LOCAL_RANKS")) { const int v = atoi(ev); if (v > 0) ranks = (unsigned)v; }
    else if (const char* ev2 = getenv("LOCAL_WORLD_SIZE")) { const int v = atoi(ev2); if (v > 0) ranks = (unsigned)v; }
    n /= ranks;
    return n < 1 ? 1 : n;
}

struct flacb200_ctx {
    int device = 0;
    // stream = the caller's stream (flacb200_set_stream; default: own_stream): the engine only records / waits on it.  All encode
    // work runs on enc_stream, which waits for the caller's stream as it was at the call (whatever produced the PCM), and the MD5
    // chain of a batch waits for exactly that too -- NOT for the encode kernels of earlier batches, which on one shared stream
    // stood between the call and its chain and left every run with an MD5 tail behind its last batch.  flacb200_join / result / fetch
    // make the caller's stream (or the host) wait for the engine.
    cudaStream_t stream = nullptr, own_stream = nullptr, md5_stream = nullptr, enc_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_user = nullptr;
    bool profiling = false;
    cudaEvent_t ev_k[10] = {nullptr};      // analysis start, analysis end, pack end, layout end, compact end, finalize end, md5 start, md5 end, OR/AND end, autocorrelation end
    std::string err;
    uint64_t launches = 0;
    int max_smem_optin = 0;

    // last batch
    EncParams P{};
    flacb200_enc_config cfg{};
    int n_frames = 0, n_streams = 0;
    uint32_t scratch_stride = 0;
    bool have_batch = false, debug = false;
    bool use_fused = false;                // this batch runs enc_fused.cu (16-bit stereo, one kernel per frame from PCM to bytes)
    std::vector<FrameDesc> h_frames;
    std::vector<uint32_t> h_stream_first, h_stream_nframes;
    std::vector<uint64_t> h_stream_off, h_stream_samples;
    std::vector<uint32_t> h_first_frame;
    std::vector<uint8_t> prev_ca;              // loose mid/side: channel assignment before the batch, per stream (one-shot)
    bool prev_ca_pending = false;
    std::vector<float> h_windows;
    std::map<uint32_t, uint32_t> window_off;   // blocksize -> float offset in h_windows
    float window_p = -1.0f;

    // host<->device pipeline of flacb200_encode_batch_host: copy streams, per-chunk events, host MD5 workers
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    static constexpr int kMaxChunks = 16;
    cudaEvent_t ev_h2d[kMaxChunks] = {nullptr}, ev_done[kMaxChunks] = {nullptr};
    void* dec = nullptr;                   // decode engine state (dec_engine.cu)
    void (*dec_free)(void*) = nullptr;
    uint64_t* h_totals = nullptr;          // pinned: per-chunk arena byte counts
    EncStats* h_stats = nullptr;           // pinned: the batch's counters (libm-log guard)
    DevBuf d_totals;
    uint64_t e2e_last_bytes = 0;
    double e2e_ms[10] = {0};              // last host call: host md5 done, enqueue, kernels drained, d2h done, md5 join, total, GPU md5 done, streams hashed on the GPU, host md5 threads, chunks
    HostPool pool;
    cudaEvent_t ev_md5 = nullptr, ev_d2h = nullptr;      // blocking-sync events: the calling thread sleeps instead of spinning on a core the MD5 workers need
    uint8_t* h_digests = nullptr; size_t h_digests_cap = 0;   // pinned: digests of the streams hashed on the GPU
    std::atomic<uint64_t> gpu_md5_done_us{0};
    std::chrono::steady_clock::time_point t_call;
    // libm-log guard (DESIGN.md "log guard"): decisions inside the band are logged per output set, repeated on the host with the
    // libm the reference links against, and -- should the host decide otherwise -- the batch is encoded again with overrides
    static constexpr uint32_t kGuardCap = 1024;
    double guard_rel = 1e-12; uint32_t guard_flip = 0;
    std::vector<LogGuardOverride> h_ovr; DevBuf d_guard_ovr;
    uint64_t guard_info[4] = {0, 0, 0, 0};    // last batch: decisions inside the band, confirmed by the host, overridden, not checked (log full)
    const void* last_d_pcm = nullptr; bool last_pcm_staged = false; bool in_rerun = false;
    uint64_t batch_seq = 0, settled_seq = 0;   // the guard of a batch is settled once, however often its result is asked for
    struct HostJob;                       // one in-flight flacb200_encode_host_submit (defined below)
    static constexpr int kJobs = 3;
    HostJob* jobs[kJobs] = {nullptr};
    int next_job = 0;
    int md5_gpu_chunks = -1;              // chunks (from the front of the batch) whose streams the GPU hashes; -1: not decided yet
    uint64_t md5_split_key = 0;           // batch shape the split was tuned for

    DevBuf d_pcm, d_frames, d_windows, d_plans, d_ca, d_scratch, d_work, d_flags;
    DevBuf d_sfirst, d_snframes, d_soff, d_ssamples, d_debug;

    // Output buffers exist three times and rotate per batch: the MD5 of a batch (a serial chain per stream, longer
    // than analysis + packing for long streams) and the STREAMINFO finalisation that needs it run on the set's own
    // side stream, so the next batch's kernels do not wait for them; a set is reused only after its finalize is done.
    struct OutSet {
        DevBuf flen, foff, arena, sinfo, md5, total, stats, guard_log;
        cudaStream_t side = nullptr;
        cudaEvent_t ev_main = nullptr, ev_free = nullptr;
        bool busy = false;
    };
    static constexpr int kSets = 5;
    OutSet sets[kSets];
    int cur = 0;
    OutSet& set() { return sets[cur]; }
};

static int wait_all_sets(flacb200_ctx* ctx, cudaStream_t st);
void fb_ctx_free_jobs(flacb200_ctx* ctx);
bool fb_ctx_jobs_in_flight(flacb200_ctx* ctx);
static int fail(flacb200_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess) {
    char buf[512];
    if (e != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    else snprintf(buf, sizeof buf, "%s", what);
    if (c) c->err = buf;
    return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, FLACB200_ERR_CUDA, #call, e_); } while (0)

// ---------------------------------------------------------------- settings ----
// ref: pyflac/include/FLAC/stream_encoder.h:845-853 -- the compression-level table
static const struct { int ms, loose; uint32_t max_lpc, max_po; int parts; } kLevels[9] = {
    {0, 0, 0, 3, 1}, {1, 1, 0, 3, 1}, {1, 0, 0, 3, 1}, {0, 0, 6, 4, 1}, {1, 1, 8, 4, 1},
    {1, 0, 8, 5, 1}, {1, 0, 8, 6, 2}, {1, 0, 12, 6, 2}, {1, 0, 12, 6, 3}};

extern "C" int flacb200_enc_validate(const flacb200_enc_config* c) {
    // order of checks follows FLAC__stream_encoder_init_stream (verified against the reference binary);
    // values: pyflac/builder/encoder.py:65-80
    if (c->channels == 0 || c->channels > 8) return 4;
    if (c->bits_per_sample < 4 || c->bits_per_sample > 32) return 5;
    if (c->sample_rate > 1048575u) return 6;
    const uint32_t lvl = c->compression_level > 8 ? 8 : c->compression_level;
    const uint32_t maxlpc = c->tune ? c->max_lpc_order : kLevels[lvl].max_lpc;
    const uint32_t maxpo = c->tune ? c->max_residual_partition_order : kLevels[lvl].max_po;
    uint32_t bs = c->blocksize ? c->blocksize : (maxlpc == 0 ? 1152u : 4096u);
    if (bs < 16 || bs > 65535) return 7;
    if (maxlpc > 32) return 8;                                           // INVALID_MAX_LPC_ORDER
    if (bs < maxlpc) return 10;
    if (c->tune && c->qlp_coeff_precision != 0 && (c->qlp_coeff_precision < 5 || c->qlp_coeff_precision > 15)) return 9;   // INVALID_QLP_COEFF_PRECISION
    if (c->streamable_subset) {
        const uint32_t b = c->bits_per_sample;
        if (!(b == 8 || b == 12 || b == 16 || b == 20 || b == 24 || b == 32)) return 11;
        if (c->sample_rate <= 48000 && (bs > 4608 || maxlpc > 12)) return 11;
        if (bs > 16384) return 11;
        // a subset frame header must be able to carry the rate: above 16 bits only multiples of 10 Hz can (format.h:  FLAC__format_sample_rate_is_subset)
        if (c->sample_rate >= (1u << 16) && c->sample_rate % 10u != 0u) return 11;
        if (maxpo > 8) return 11;                                        // FLAC__SUBSET_MAX_RICE_PARTITION_ORDER
    }
    return 0;
}

static int resolve_params(flacb200_ctx* ctx, const flacb200_enc_config& c, EncParams& P) {
    const int st = flacb200_enc_validate(&c);
    if (st != 0) { char b[64]; snprintf(b, sizeof b, "init status %d", st); return fail(ctx, FLACB200_ERR_CONFIG, b); }
    const uint32_t lvl = c.compression_level > 8 ? 8 : c.compression_level;
    memset(&P, 0, sizeof P);
    P.channels = c.channels; P.bps = c.bits_per_sample; P.sample_rate = c.sample_rate;
    // the level's presets, or the caller's fine-grained settings (stream_encoder.h: set_do_mid_side_stereo ... set_apodization)
    bool ms = kLevels[lvl].ms != 0, loose = kLevels[lvl].loose != 0;
    P.max_lpc_order = kLevels[lvl].max_lpc; P.max_part_order = kLevels[lvl].max_po; P.apod_parts = (uint32_t)kLevels[lvl].parts;
    uint32_t qlp = 0;
    if (c.tune) {
        ms = c.do_mid_side != 0; loose = c.loose_mid_side != 0;
        P.max_lpc_order = c.max_lpc_order; P.max_part_order = c.max_residual_partition_order > 15 ? 15 : c.max_residual_partition_order;
        P.apod_parts = c.apod_parts; qlp = c.qlp_coeff_precision;
        if (P.max_lpc_order > (uint32_t)kMaxOrder || P.max_part_order > (uint32_t)kMaxPartOrder || P.apod_parts < 1 || P.apod_parts > 3 ||
            qlp > 15 || !(c.apod_p >= 0.0f && c.apod_p <= 1.0f))
            return fail(ctx, FLACB200_ERR_UNSUPPORTED, "fine-grained settings outside this build's range (max_lpc_order <= 12, max_residual_partition_order <= 6, tukey / subdivide_tukey(2..3), qlp_coeff_precision <= 15)");
    }
    P.blocksize = c.blocksize ? c.blocksize : (P.max_lpc_order == 0 ? 1152u : 4096u);
    P.do_mid_side = (ms && c.channels == 2) ? 1u : 0u;
    P.rice_limit = c.bits_per_sample > 16 ? 31u : 15u;
    P.container_bytes = c.container_bytes;
    P.n_signals = c.channels + (P.do_mid_side ? 2u : 0u);
    if (qlp) P.qlp_precision = qlp;
    else if (c.bits_per_sample < 16) { uint32_t p = 2 + c.bits_per_sample / 2; P.qlp_precision = p < 5 ? 5 : p; }
    else if (c.bits_per_sample == 16) {
        const uint32_t b = P.blocksize;
        P.qlp_precision = b <= 192 ? 7 : b <= 384 ? 8 : b <= 576 ? 9 : b <= 1152 ? 10 : b <= 2304 ? 11 : b <= 4608 ? 12 : 13;
    } else {
        const uint32_t b = P.blocksize;
        P.qlp_precision = b <= 384 ? 13 : b <= 1152 ? 14 : 15;
    }
    // up: FLAC__stream_encoder_init_*: loose_mid_side_stereo_frames = (uint32_t)(sample_rate * 0.4 / blocksize + 0.5), at least 1
    if (loose && P.do_mid_side) {
        P.loose_frames = (uint32_t)((double)c.sample_rate * 0.4 / (double)P.blocksize + 0.5);
        if (P.loose_frames == 0) P.loose_frames = 1;
    }
    P.limit_min_bitrate = c.limit_min_bitrate ? 1u : 0u;
    // limits of this build (DESIGN.md "limits")
    if (c.container_bytes != 2 && c.container_bytes != 4) return fail(ctx, FLACB200_ERR_ARG, "container_bytes must be 2 or 4");
    if (c.container_bytes == 2 && c.bits_per_sample > 16) return fail(ctx, FLACB200_ERR_ARG, "int16 container needs bits_per_sample <= 16");
    P.smem_stride = ((P.blocksize + 3) / 4) * 4;
    analyze_layout(P);
    return 0;
}

// up: window.c FLAC__window_tukey -- generated on the host with the same libm cosf the reference
// binary imports (cosf@GLIBC), never on the device
static void window_tukey(float* w, int32_t L, float p) {
    for (int32_t n = 0; n < L; n++) w[n] = 1.0f;
    if (p <= 0.0f) return;
    if (p >= 1.0f) { const int32_t N = L - 1; for (int32_t n = 0; n < L; n++) w[n] = (float)(0.5f - 0.5f * cosf(2.0f * (float)M_PI * n / N)); return; }
    const int32_t Np = (int32_t)(p / 2.0f * L) - 1;
    if (Np > 0) {
        for (int32_t n = 0; n <= Np; n++) {
            w[n] = (float)(0.5f - 0.5f * cosf((float)(M_PI * n / Np)));
            w[L - Np - 1 + n] = (float)(0.5f - 0.5f * cosf((float)(M_PI * (n + Np) / Np)));
        }
    }
}

static uint32_t max_frame_bytes(const EncParams& P) {
    // header <= 16, per channel: 8-bit subframe header + unary wasted + N*(bps+1) + Rice slack (N/2 + params), CRC-16
    const uint64_t per_ch_bits = 8 + 32 + (uint64_t)P.blocksize * (P.bps + 1) + P.blocksize / 2 + 64 * 5 + 64;
    uint64_t bytes = 16 + (uint64_t)P.channels * ((per_ch_bits + 7) / 8 + 1) + 2 + 16;
    return (uint32_t)((bytes + 15) / 16 * 16);
}

// ---------------------------------------------------------------- ctx ----
extern "C" int flacb200_create(flacb200_ctx** out, int device) {
    if (!out) return FLACB200_ERR_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return FLACB200_ERR_NO_DEVICE;
    if (device < 0 || device >= n) return FLACB200_ERR_ARG;
    flacb200_ctx* ctx = new flacb200_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return FLACB200_ERR_NO_DEVICE; }
    cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->md5_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->enc_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->ev_user, cudaEventDisableTiming);
    for (auto& S : ctx->sets) {
        cudaStreamCreateWithFlags(&S.side, cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&S.ev_main, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&S.ev_free, cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    for (auto& e : ctx->ev_k) cudaEventCreate(&e);
    cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking);
    for (auto& e : ctx->ev_h2d) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (auto& e : ctx->ev_done) cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync);
    cudaEventCreateWithFlags(&ctx->ev_md5, cudaEventDisableTiming | cudaEventBlockingSync);
    cudaEventCreateWithFlags(&ctx->ev_d2h, cudaEventDisableTiming | cudaEventBlockingSync);
    cudaHostAlloc((void**)&ctx->h_totals, sizeof(uint64_t) * flacb200_ctx::kMaxChunks, cudaHostAllocDefault);
    cudaHostAlloc((void**)&ctx->h_stats, sizeof(EncStats), cudaHostAllocDefault);
    ctx->stream = ctx->own_stream;
    if (const char* ev = getenv("FLACB200_LOG_GUARD_REL")) { const double v = atof(ev); if (v > 0.0) ctx->guard_rel = v; }      // tests widen the band
    if (const char* ev = getenv("FLACB200_LOG_GUARD_FLIP")) ctx->guard_flip = (uint32_t)atoi(ev);                             // tests: wrong device decisions
    *out = ctx;
    return FLACB200_OK;
}

extern "C" void flacb200_destroy(flacb200_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->enc_stream);
    cudaStreamSynchronize(ctx->stream);
    DevBuf* bufs[] = {&ctx->d_guard_ovr, &ctx->d_flags, &ctx->d_pcm, &ctx->d_frames, &ctx->d_windows, &ctx->d_plans, &ctx->d_ca, &ctx->d_scratch, &ctx->d_work, &ctx->d_sfirst, &ctx->d_snframes,
                      &ctx->d_soff, &ctx->d_ssamples, &ctx->d_debug};
    for (DevBuf* b : bufs) b->release();
    for (auto& S : ctx->sets) {
        if (S.side) cudaStreamSynchronize(S.side);
        DevBuf* sb[] = {&S.flen, &S.foff, &S.arena, &S.sinfo, &S.md5, &S.total, &S.stats, &S.guard_log};
        for (DevBuf* b : sb) b->release();
        if (S.ev_main) cudaEventDestroy(S.ev_main);
        if (S.ev_free) cudaEventDestroy(S.ev_free);
        if (S.side) cudaStreamDestroy(S.side);
    }
    cudaEventDestroy(ctx->ev_fork); cudaEventDestroy(ctx->ev_join); cudaEventDestroy(ctx->ev_user);
    cudaStreamSynchronize(ctx->enc_stream); cudaStreamDestroy(ctx->enc_stream);
    for (auto& e : ctx->ev_k) cudaEventDestroy(e);
    for (auto& e : ctx->ev_h2d) cudaEventDestroy(e);
    for (auto& e : ctx->ev_done) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->h2d_stream); cudaStreamDestroy(ctx->d2h_stream);
    fb_ctx_free_jobs(ctx);
    if (ctx->dec && ctx->dec_free) ctx->dec_free(ctx->dec);
    if (ctx->h_totals) cudaFreeHost(ctx->h_totals);
    if (ctx->h_stats) cudaFreeHost(ctx->h_stats);
    if (ctx->h_digests) cudaFreeHost(ctx->h_digests);
    if (ctx->ev_md5) cudaEventDestroy(ctx->ev_md5);
    if (ctx->ev_d2h) cudaEventDestroy(ctx->ev_d2h);
    ctx->d_totals.release();
    cudaStreamDestroy(ctx->own_stream); cudaStreamDestroy(ctx->md5_stream);
    delete ctx;
}

extern "C" const char* flacb200_last_error(const flacb200_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context (no CUDA device?)"; }
extern "C" int flacb200_set_stream(flacb200_ctx* ctx, void* s) { if (!ctx) return FLACB200_ERR_ARG; ctx->stream = s ? (cudaStream_t)s : ctx->own_stream; return 0; }
extern "C" int flacb200_sync(flacb200_ctx* ctx) {
    if (!ctx) return FLACB200_ERR_ARG;
    cudaSetDevice(ctx->device);
    CK(cudaStreamSynchronize(ctx->enc_stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto& S : ctx->sets) if (S.busy) CK(cudaStreamSynchronize(S.side));
    return 0;
}
// accessors for the decode engine (dec_engine.cu keeps its own state behind ctx->dec)
int fb_ctx_device(flacb200_ctx* c) { return c->device; }
cudaStream_t fb_ctx_stream(flacb200_ctx* c) { return c->stream; }
void** fb_ctx_dec_slot(flacb200_ctx* c, void (*freer)(void*)) { c->dec_free = freer; return &c->dec; }
int fb_ctx_fail(flacb200_ctx* c, int code, const char* what, cudaError_t e) { return fail(c, code, what, e); }
void fb_ctx_add_launches(flacb200_ctx* c, uint64_t n) { c->launches += n; }

// Make the ctx stream wait for every side stream (MD5 + finalize of in-flight batches): after this, an event recorded
// on the ctx stream covers all work issued so far.
extern "C" int flacb200_join(flacb200_ctx* ctx) {
    if (!ctx) return FLACB200_ERR_ARG;
    cudaSetDevice(ctx->device);
    CK(cudaEventRecord(ctx->ev_join, ctx->enc_stream));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    return wait_all_sets(ctx, ctx->stream);
}

extern "C" int flacb200_host_path_times(flacb200_ctx* ctx, double* ms) { if (!ctx || !ms) return FLACB200_ERR_ARG; for (int i = 0; i < 6; i++) ms[i] = ctx->e2e_ms[i]; return 0; }
extern "C" int flacb200_host_path_info(flacb200_ctx* ctx, double* v, int n) { if (!ctx || !v || n < 0) return FLACB200_ERR_ARG; for (int i = 0; i < n && i < 10; i++) v[i] = ctx->e2e_ms[i]; return 0; }

extern "C" int flacb200_set_profiling(flacb200_ctx* ctx, int on) { if (!ctx) return FLACB200_ERR_ARG; ctx->profiling = on != 0; return 0; }
// ms[0..8] = analysis (3 kernels), pack, layout(scan), compact, finalize (incl. waiting for MD5), md5 (side stream),
// then the analysis split: OR/AND, autocorrelation, decisions -- of the last batch
extern "C" int flacb200_kernel_times(flacb200_ctx* ctx, float* ms) {
    if (!ctx || !ms || !ctx->profiling || !ctx->have_batch) return FLACB200_ERR_ARG;
    cudaSetDevice(ctx->device);
    CK(cudaStreamSynchronize(ctx->enc_stream));
    for (auto& S : ctx->sets) CK(cudaStreamSynchronize(S.side));
    for (int i = 0; i < 5; i++) CK(cudaEventElapsedTime(&ms[i], ctx->ev_k[i], ctx->ev_k[i + 1]));
    ms[5] = 0.0f;
    if (ctx->cfg.do_md5) CK(cudaEventElapsedTime(&ms[5], ctx->ev_k[6], ctx->ev_k[7]));
    // the three kernels of the analysis: ms[6] OR/AND, ms[7] autocorrelation (0 when the level has no LPC), ms[8] decisions
    CK(cudaEventElapsedTime(&ms[6], ctx->ev_k[0], ctx->ev_k[8]));
    CK(cudaEventElapsedTime(&ms[7], ctx->ev_k[8], ctx->ev_k[9]));
    CK(cudaEventElapsedTime(&ms[8], ctx->ev_k[9], ctx->ev_k[1]));
    ms[9] = ctx->use_fused ? 1.0f : 0.0f;        // 1: the TMA-staged kernels of enc_fused.cu ran: ms[7] autocorrelation, ms[8] analysis, ms[1] pack; ms[6] (OR/AND pass) is zero
    return 0;
}
extern "C" int flacb200_log_guard_info(flacb200_ctx* ctx, uint64_t* v) { if (!ctx || !v) return FLACB200_ERR_ARG; for (int i = 0; i < 4; i++) v[i] = ctx->guard_info[i]; return 0; }
extern "C" int flacb200_set_log_guard(flacb200_ctx* ctx, double rel, int flip) { if (!ctx) return FLACB200_ERR_ARG; ctx->guard_rel = rel > 0.0 ? rel : 1e-12; ctx->guard_flip = flip ? 1u : 0u; return 0; }
extern "C" uint64_t flacb200_launch_count(const flacb200_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------- batch encode ----
static bool same_layout(const flacb200_ctx* c, const flacb200_enc_config& cfg, uint32_t ns, const uint64_t* off,
                        const uint64_t* smp, const uint32_t* ffn) {
    if (!c->have_batch || memcmp(&c->cfg, &cfg, sizeof cfg) != 0 || (uint32_t)c->n_streams != ns) return false;
    if (memcmp(c->h_stream_off.data(), off, ns * sizeof(uint64_t)) != 0) return false;
    if (memcmp(c->h_stream_samples.data(), smp, ns * sizeof(uint64_t)) != 0) return false;
    for (uint32_t s = 0; s < ns; s++) if (c->h_first_frame[s] != (ffn ? ffn[s] : 0u)) return false;
    return true;
}

static int plan_batch(flacb200_ctx* ctx, const flacb200_enc_config& cfg, uint32_t ns, const uint64_t* off,
                      const uint64_t* smp, const uint32_t* ffn) {
    EncParams P;
    int rc = resolve_params(ctx, cfg, P);
    if (rc) return rc;
    const float ap = cfg.tune ? cfg.apod_p : 0.5f;
    const float wp = P.apod_parts == 1 ? ap : ap / (float)P.apod_parts;       // up: set_apodization: tukey(p); subdivide_tukey: p/parts in float
    if (wp != ctx->window_p) { ctx->h_windows.clear(); ctx->window_off.clear(); ctx->window_p = wp; }
    ctx->h_frames.clear();
    ctx->h_stream_first.assign(ns, 0); ctx->h_stream_nframes.assign(ns, 0);
    ctx->h_stream_off.assign(off, off + ns); ctx->h_stream_samples.assign(smp, smp + ns);
    ctx->h_first_frame.assign(ns, 0);
    bool windows_grew = false;
    for (uint32_t s = 0; s < ns; s++) {
        ctx->h_stream_first[s] = (uint32_t)ctx->h_frames.size();
        ctx->h_first_frame[s] = ffn ? ffn[s] : 0u;
        uint32_t fn = ctx->h_first_frame[s];
        for (uint64_t done = 0; done < smp[s];) {
            const uint32_t N = (smp[s] - done >= P.blocksize) ? P.blocksize : (uint32_t)(smp[s] - done);
            FrameDesc fd;
            fd.pcm_off = off[s] + done * P.channels;
            fd.blocksize = N; fd.frame_number = fn; fd.stream = s; fd.window_off = 0; fd.lead = 0; fd.pad = 0;
            if (P.loose_frames) {
                const uint32_t phase = fn % P.loose_frames, k = fn - ctx->h_first_frame[s];     // k = index within this batch
                if (phase != 0) fd.lead = (k >= phase) ? phase : (kLeadForced | (ctx->prev_ca.size() > s ? (ctx->prev_ca[s] & 3u) : 0u));
            }
            fn++;
            if (P.max_lpc_order > 0) {
                auto it = ctx->window_off.find(N);
                if (it == ctx->window_off.end()) {
                    const uint32_t o = (uint32_t)ctx->h_windows.size();
                    ctx->h_windows.resize(o + N);
                    window_tukey(ctx->h_windows.data() + o, (int32_t)N, wp);
                    it = ctx->window_off.emplace(N, o).first;
                    windows_grew = true;
                }
                fd.window_off = it->second;
            }
            ctx->h_frames.push_back(fd);
            done += N;
        }
        ctx->h_stream_nframes[s] = (uint32_t)ctx->h_frames.size() - ctx->h_stream_first[s];
    }
    ctx->prev_ca.clear(); ctx->prev_ca_pending = false;       // one-shot: the next batch starts fresh unless set again
    ctx->P = P; ctx->cfg = cfg; ctx->n_streams = (int)ns; ctx->n_frames = (int)ctx->h_frames.size();
    ctx->scratch_stride = max_frame_bytes(P);
    ctx->debug = cfg.debug_trace != 0;

    ctx->use_fused = !ctx->debug && !getenv("FLACB200_NO_FUSED") && fused_eligible(P, ctx->scratch_stride, ctx->max_smem_optin);
    const size_t sa = analyze_smem_bytes(P), sp = pack_smem_bytes(P, ctx->scratch_stride);
    if (!ctx->use_fused && ((int)sa > ctx->max_smem_optin || (int)sp > ctx->max_smem_optin))
        return fail(ctx, FLACB200_ERR_UNSUPPORTED, "blocksize x channels exceeds the shared-memory frame tile of this build");

    const int nf = ctx->n_frames;
    cudaStream_t st = ctx->enc_stream;
    CK(ctx->d_frames.reserve(sizeof(FrameDesc) * (size_t)(nf ? nf : 1)));
    CK(ctx->d_plans.reserve(sizeof(SubframePlan) * (size_t)(nf ? nf : 1) * P.n_signals));
    CK(ctx->d_work.reserve(analyze_work_stride(P) * (size_t)(nf ? nf : 1)));
    CK(ctx->d_ca.reserve((size_t)nf + 16));
    CK(ctx->d_scratch.reserve((size_t)nf * ctx->scratch_stride + 64));
    for (auto& S : ctx->sets) CK(S.flen.reserve(sizeof(uint32_t) * (size_t)(nf + 1)));
    for (auto& S : ctx->sets) CK(S.foff.reserve(sizeof(uint64_t) * (size_t)(nf + 1)));
    for (auto& S : ctx->sets) CK(S.arena.reserve((size_t)nf * ctx->scratch_stride + (size_t)ns * kStreamPrologueBytes + 64 + 256 * (flacb200_ctx::kMaxChunks + 1)));
    for (auto& S : ctx->sets) CK(S.total.reserve(64));
    for (auto& S : ctx->sets) CK(S.stats.reserve(sizeof(EncStats)));
    for (auto& S : ctx->sets) CK(S.guard_log.reserve(sizeof(LogGuardEntry) * flacb200_ctx::kGuardCap));
    CK(ctx->d_sfirst.reserve(sizeof(uint32_t) * (ns + 1)));
    CK(ctx->d_snframes.reserve(sizeof(uint32_t) * (ns + 1)));
    CK(ctx->d_soff.reserve(sizeof(uint64_t) * (ns + 1)));
    CK(ctx->d_ssamples.reserve(sizeof(uint64_t) * (ns + 1)));
    for (auto& S : ctx->sets) CK(S.md5.reserve(16 * (size_t)(ns + 1)));
    for (auto& S : ctx->sets) CK(S.sinfo.reserve(sizeof(StreamInfoOut) * (size_t)(ns + 1)));
    if (ctx->debug) CK(ctx->d_debug.reserve(sizeof(SignalDebug) * (size_t)(nf ? nf : 1) * P.n_signals));
    if (nf) CK(cudaMemcpyAsync(ctx->d_frames.p, ctx->h_frames.data(), sizeof(FrameDesc) * nf, cudaMemcpyHostToDevice, st));
    if (ns) {
        CK(cudaMemcpyAsync(ctx->d_sfirst.p, ctx->h_stream_first.data(), sizeof(uint32_t) * ns, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->d_snframes.p, ctx->h_stream_nframes.data(), sizeof(uint32_t) * ns, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->d_soff.p, ctx->h_stream_off.data(), sizeof(uint64_t) * ns, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->d_ssamples.p, ctx->h_stream_samples.data(), sizeof(uint64_t) * ns, cudaMemcpyHostToDevice, st));
    }
    if (windows_grew || (ctx->d_windows.cap < ctx->h_windows.size() * sizeof(float))) {
        CK(ctx->d_windows.reserve(ctx->h_windows.size() * sizeof(float) + 64));
        windows_grew = true;
    }
    if (windows_grew && !ctx->h_windows.empty())
        CK(cudaMemcpyAsync(ctx->d_windows.p, ctx->h_windows.data(), ctx->h_windows.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    // the host vectors above are pageable: make sure the copies are done before they can change
    CK(cudaStreamSynchronize(st));
    ctx->have_batch = true;
    return 0;
}

// make sure no side stream still owns any output set (host path and teardown use this)
static int wait_all_sets(flacb200_ctx* ctx, cudaStream_t st) {
    for (auto& S : ctx->sets) if (S.busy) { CK(cudaStreamWaitEvent(st, S.ev_free, 0)); }
    return 0;
}

// the batch's settings with the guard fields of output set S filled in
static EncParams params_for_set(flacb200_ctx* ctx, flacb200_ctx::OutSet& S) {
    EncParams P = ctx->P;
    P.guard_rel = ctx->guard_rel; P.guard_flip = ctx->in_rerun ? 0u : ctx->guard_flip; P.guard_cap = flacb200_ctx::kGuardCap;
    P.guard_log = (LogGuardEntry*)S.guard_log.p;
    P.guard_n_ovr = (uint32_t)ctx->h_ovr.size(); P.guard_ovr = (const LogGuardOverride*)ctx->d_guard_ovr.p;
    return P;
}

// Repeat the logged decisions with the host's libm log (the one the reference binary imports).  Returns the number of decisions
// the host makes differently; those land in ctx->h_ovr / d_guard_ovr for a second pass over the batch.
static int guard_settle(flacb200_ctx* ctx, flacb200_ctx::OutSet& S, uint64_t seen, size_t* n_new) {
    *n_new = 0;
    ctx->guard_info[0] = seen; ctx->guard_info[1] = ctx->guard_info[2] = 0; ctx->guard_info[3] = 0;
    if (!seen) return 0;
    const size_t n = (size_t)std::min<uint64_t>(seen, flacb200_ctx::kGuardCap);
    ctx->guard_info[3] = seen - n;
    std::vector<LogGuardEntry> log(n);
    CK(cudaMemcpy(log.data(), S.guard_log.p, n * sizeof(LogGuardEntry), cudaMemcpyDeviceToHost));
    std::vector<LogGuardOverride> ovr;
    for (const LogGuardEntry& e : log) {
        // up: lpc.c FLAC__lpc_compute_best_order / stream_encoder.c evaluate_lpc_subframe_ -- the same arithmetic as the kernels, host libm
        const double escale = 0.5 / (double)e.N;
        double best = 4294967295.0; int guess = 1;
        for (uint32_t o = 1; o <= e.max_order && o <= (uint32_t)kMaxOrder; o++) {
            bool ul;
            const double bits = expected_bits_per_sample(e.lperr[o - 1], escale, &ul) * (double)(e.N - o) + (double)(o * e.overhead);
            if (bits < best) { best = bits; guess = (int)o; }
        }
        bool ul2;
        const double rbps = expected_bits_per_sample(e.lperr[guess - 1], 0.5 / (double)(e.N - (uint32_t)guess), &ul2);
        const int skip = rbps >= (double)e.sbps ? 1 : 0;
        if (guess != e.guess || skip != e.skip) ovr.push_back(LogGuardOverride{e.stream, e.frame_number, e.signal, e.step, guess, skip});
        else ctx->guard_info[1]++;
    }
    ctx->guard_info[2] = ovr.size();
    *n_new = ovr.size();
    if (!ovr.empty()) {
        ctx->h_ovr = ovr;
        CK(ctx->d_guard_ovr.reserve(ovr.size() * sizeof(LogGuardOverride)));
        CK(cudaMemcpy(ctx->d_guard_ovr.p, ovr.data(), ovr.size() * sizeof(LogGuardOverride), cudaMemcpyHostToDevice));
    }
    return 0;
}

// pcm_staged: the PCM was copied into ctx->d_pcm on the encode stream (host input): the MD5 chain waits for that copy; otherwise it
// waits for the caller's stream as it was at the call (ev_user) and for nothing else
static int run_batch(flacb200_ctx* ctx, const void* d_pcm, bool pcm_staged) {
    ctx->last_d_pcm = d_pcm; ctx->last_pcm_staged = pcm_staged; ctx->batch_seq++;
    const int nf = ctx->n_frames, ns = ctx->n_streams;
    cudaStream_t st = ctx->enc_stream;
    ctx->cur = (ctx->cur + 1) % flacb200_ctx::kSets;
    flacb200_ctx::OutSet& S = ctx->set();
    const EncParams P = params_for_set(ctx, S);
    if (S.busy) { CK(cudaStreamWaitEvent(st, S.ev_free, 0)); S.busy = false; }   // its previous finalize must be done before reuse
    CK(cudaMemsetAsync(S.stats.p, 0, sizeof(EncStats), st));
    CK(cudaMemsetAsync(S.total.p, 0, 8, st));
    if (nf == 0) return 0;
    const bool md5 = ctx->cfg.do_md5 != 0;
    const bool prof = ctx->profiling;
    if (md5) {
        if (pcm_staged) { CK(cudaEventRecord(ctx->ev_fork, st)); CK(cudaStreamWaitEvent(S.side, ctx->ev_fork, 0)); }
        else CK(cudaStreamWaitEvent(S.side, ctx->ev_user, 0));
        if (prof) CK(cudaEventRecord(ctx->ev_k[6], S.side));
        launch_md5(d_pcm, P.container_bytes, (const uint64_t*)ctx->d_soff.p, (const uint64_t*)ctx->d_ssamples.p, ns, P.channels, P.bps,
                   (uint8_t*)S.md5.p, S.side);
        if (prof) CK(cudaEventRecord(ctx->ev_k[7], S.side));
        ctx->launches++;
    }
    if (prof) CK(cudaEventRecord(ctx->ev_k[0], st));
    int n_an;
    if (ctx->use_fused) {
        // one kernel from PCM to frame bytes (enc_fused.cu); the split events collapse onto its end
        // enc_fused.cu path: autocorrelation (un-shifted, no OR/AND pass needed) -> TMA-staged analysis -> TMA-staged pack
        if (prof) CK(cudaEventRecord(ctx->ev_k[8], st));
        launch_autoc_unshifted(d_pcm, (const FrameDesc*)ctx->d_frames.p, (const float*)ctx->d_windows.p, P, nf, ctx->d_work.p, st);
        if (prof) CK(cudaEventRecord(ctx->ev_k[9], st));
        launch_fused(d_pcm, (const FrameDesc*)ctx->d_frames.p, P, nf, (const uint8_t*)ctx->d_work.p + 64, analyze_work_stride(P), (SubframePlan*)ctx->d_plans.p,
                     (uint8_t*)ctx->d_ca.p, (EncStats*)S.stats.p, (uint8_t*)ctx->d_scratch.p, ctx->scratch_stride, (uint32_t*)S.flen.p, st,
                     prof ? ctx->ev_k[1] : nullptr);
        n_an = (P.max_lpc_order > 0 ? 1 : 0) + 1;
    } else {
        n_an = launch_analyze(d_pcm, (const FrameDesc*)ctx->d_frames.p, (const float*)ctx->d_windows.p, P, nf, (SubframePlan*)ctx->d_plans.p,
                              (uint8_t*)ctx->d_ca.p, ctx->debug ? (SignalDebug*)ctx->d_debug.p : nullptr, (EncStats*)S.stats.p,
                              analyze_smem_bytes(P), ctx->d_work.p, st, prof ? &ctx->ev_k[8] : nullptr);
        if (prof) CK(cudaEventRecord(ctx->ev_k[1], st));
        launch_pack(d_pcm, (const FrameDesc*)ctx->d_frames.p, P, nf, (const SubframePlan*)ctx->d_plans.p, (const uint8_t*)ctx->d_ca.p,
                    (uint8_t*)ctx->d_scratch.p, ctx->scratch_stride, (uint32_t*)S.flen.p, st);
    }
    if (prof) CK(cudaEventRecord(ctx->ev_k[2], st));
    const uint32_t pro = ctx->cfg.write_prologue ? (uint32_t)kStreamPrologueBytes : 0u;
    launch_layout((const uint32_t*)S.flen.p, (const FrameDesc*)ctx->d_frames.p, nf, pro, 0ull, (uint64_t*)S.foff.p, (uint64_t*)S.total.p, st);
    if (prof) CK(cudaEventRecord(ctx->ev_k[3], st));
    launch_compact((const uint8_t*)ctx->d_scratch.p, ctx->scratch_stride, (const uint32_t*)S.flen.p, (const uint64_t*)S.foff.p, (uint8_t*)S.arena.p, nf, st);
    if (prof) CK(cudaEventRecord(ctx->ev_k[4], st));
    // finalize (stream prologues, per-stream info; MD5 fields zero) closes the batch on the main stream: from here on the
    // frames and the index are final.  The MD5 chain -- serial per stream, longer than everything else for long streams --
    // keeps running on the set's side stream and patches its 16 bytes per stream into place when it is done; the main
    // stream moves on to the next batch (flacb200_encode_result_frames / flacb200_encode_fetch_md5 expose the two moments).
    launch_finalize((const uint32_t*)S.flen.p, (const uint64_t*)S.foff.p, (const uint32_t*)ctx->d_sfirst.p, (const uint32_t*)ctx->d_snframes.p,
                    (const uint64_t*)ctx->d_ssamples.p, nullptr, ns, P, pro ? 1u : 0u, (uint8_t*)S.arena.p, (StreamInfoOut*)S.sinfo.p, st);
    if (prof) CK(cudaEventRecord(ctx->ev_k[5], st));
    if (md5) {
        CK(cudaEventRecord(S.ev_main, st)); CK(cudaStreamWaitEvent(S.side, S.ev_main, 0));
        launch_md5_patch((const uint8_t*)S.md5.p, (const uint32_t*)ctx->d_snframes.p, ns, pro ? 1u : 0u, (uint8_t*)S.arena.p, (StreamInfoOut*)S.sinfo.p, S.side);
        CK(cudaEventRecord(S.ev_free, S.side)); S.busy = true;
        ctx->launches++;
    }
    ctx->launches += 4 + n_an;        // analysis kernels, pack, scan, compact, finalize
    CK(cudaGetLastError());
    return 0;
}

extern "C" int flacb200_encode_batch(flacb200_ctx* ctx, const flacb200_enc_config* cfg, const void* pcm, int pcm_is_device,
                                     uint64_t pcm_elems, uint32_t n_streams, const uint64_t* stream_off,
                                     const uint64_t* stream_samples, const uint32_t* first_frame_number) {
    if (!ctx) return FLACB200_ERR_NO_DEVICE;
    if (!cfg || (!pcm && pcm_elems) || (n_streams && (!stream_off || !stream_samples))) return fail(ctx, FLACB200_ERR_ARG, "null argument");
    cudaSetDevice(ctx->device);
    if (!ctx->in_rerun) ctx->h_ovr.clear();
    for (uint32_t s = 0; s < n_streams; s++)
        if (stream_off[s] + stream_samples[s] * cfg->channels > pcm_elems) return fail(ctx, FLACB200_ERR_ARG, "stream exceeds pcm buffer");
    if (ctx->prev_ca_pending || !same_layout(ctx, *cfg, n_streams, stream_off, stream_samples, first_frame_number)) {
        ctx->have_batch = false;
        CK(cudaStreamSynchronize(ctx->enc_stream));
        for (auto& S : ctx->sets) if (S.busy) { CK(cudaStreamSynchronize(S.side)); S.busy = false; }
        int rc = plan_batch(ctx, *cfg, n_streams, stream_off, stream_samples, first_frame_number);
        if (rc) return rc;
    }
    // the encode stream picks up behind whatever the caller's stream holds right now (the producer of the PCM)
    CK(cudaEventRecord(ctx->ev_user, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->enc_stream, ctx->ev_user, 0));
    const void* d_pcm = pcm;
    if (!pcm_is_device) {
        const size_t bytes = (size_t)pcm_elems * cfg->container_bytes;
        // the MD5 of an earlier host-PCM batch may still be reading the staging buffer on its side stream
        { int wrc = wait_all_sets(ctx, ctx->enc_stream); if (wrc) return wrc; }
        if (bytes + 64 > ctx->d_pcm.cap) for (auto& S : ctx->sets) if (S.busy) { CK(cudaStreamSynchronize(S.side)); S.busy = false; }   // about to free it
        if (bytes + 64 > ctx->d_pcm.cap) CK(cudaStreamSynchronize(ctx->enc_stream));
        CK(ctx->d_pcm.reserve(bytes + 64));
        CK(cudaMemcpyAsync(ctx->d_pcm.p, pcm, bytes, cudaMemcpyHostToDevice, ctx->enc_stream));
        d_pcm = ctx->d_pcm.p;
    }
    return run_batch(ctx, d_pcm, !pcm_is_device);
}

extern "C" int flacb200_encode_set_prev_assignment(flacb200_ctx* ctx, const uint8_t* prev, uint32_t n_streams) {
    if (!ctx) return FLACB200_ERR_NO_DEVICE;
    ctx->prev_ca.clear();
    if (prev) ctx->prev_ca.assign(prev, prev + n_streams);
    ctx->prev_ca_pending = true;           // forces a re-plan: the cached layout may carry other forced decisions
    return 0;
}

extern "C" int flacb200_encode_fetch_assignments(flacb200_ctx* ctx, uint8_t* frame_ca, size_t cap) {
    if (!ctx || !frame_ca) return FLACB200_ERR_ARG;
    if (!ctx->have_batch) return fail(ctx, FLACB200_ERR_ARG, "no batch");
    if (cap < (size_t)ctx->n_frames) return fail(ctx, FLACB200_ERR_ARG, "assignment buffer too small");
    cudaSetDevice(ctx->device);
    if (ctx->n_frames) {
        CK(cudaMemcpyAsync(frame_ca, ctx->d_ca.p, (size_t)ctx->n_frames, cudaMemcpyDeviceToHost, ctx->enc_stream));
        CK(cudaStreamSynchronize(ctx->enc_stream));
    }
    return 0;
}

static int encode_result_impl(flacb200_ctx* ctx, flacb200_enc_result* res, bool wait_md5);
extern "C" int flacb200_encode_result(flacb200_ctx* ctx, flacb200_enc_result* res) { return encode_result_impl(ctx, res, true); }
extern "C" int flacb200_encode_result_frames(flacb200_ctx* ctx, flacb200_enc_result* res) { return encode_result_impl(ctx, res, false); }
extern "C" int flacb200_encode_fetch_md5(flacb200_ctx* ctx, uint8_t* digests, size_t cap) {
    if (!ctx || !digests) return FLACB200_ERR_ARG;
    if (!ctx->have_batch) return fail(ctx, FLACB200_ERR_ARG, "no batch");
    if (cap < (size_t)ctx->n_streams * 16) return fail(ctx, FLACB200_ERR_ARG, "digest buffer too small");
    cudaSetDevice(ctx->device);
    if (!ctx->cfg.do_md5) { memset(digests, 0, (size_t)ctx->n_streams * 16); return 0; }
    cudaStream_t side = ctx->set().side;
    if (ctx->n_streams && ctx->n_frames) CK(cudaMemcpyAsync(digests, ctx->set().md5.p, (size_t)ctx->n_streams * 16, cudaMemcpyDeviceToHost, side));
    else memset(digests, 0, (size_t)ctx->n_streams * 16);
    CK(cudaStreamSynchronize(side));
    return 0;
}
// digests of an EARLIER batch of the same layout (back = 1: the one before the last): rounds of a pipeline collect them one round
// late, when the chain has long finished, instead of waiting for the last batch's
extern "C" int flacb200_encode_fetch_md5_back(flacb200_ctx* ctx, int back, uint8_t* digests, size_t cap) {
    if (!ctx || !digests || back < 0 || back >= flacb200_ctx::kSets) return FLACB200_ERR_ARG;
    if (!ctx->have_batch) return fail(ctx, FLACB200_ERR_ARG, "no batch");
    if (cap < (size_t)ctx->n_streams * 16) return fail(ctx, FLACB200_ERR_ARG, "digest buffer too small");
    cudaSetDevice(ctx->device);
    if (!ctx->cfg.do_md5 || !ctx->n_streams || !ctx->n_frames) { memset(digests, 0, (size_t)ctx->n_streams * 16); return 0; }
    flacb200_ctx::OutSet& S = ctx->sets[(ctx->cur - back + 2 * flacb200_ctx::kSets) % flacb200_ctx::kSets];
    CK(cudaMemcpyAsync(digests, S.md5.p, (size_t)ctx->n_streams * 16, cudaMemcpyDeviceToHost, S.side));
    CK(cudaStreamSynchronize(S.side));
    return 0;
}
static int encode_result_impl(flacb200_ctx* ctx, flacb200_enc_result* res, bool wait_md5) {
    if (!ctx || !res) return FLACB200_ERR_ARG;
    if (!ctx->have_batch) return fail(ctx, FLACB200_ERR_ARG, "no batch");
    cudaSetDevice(ctx->device);
    uint64_t total = 0; EncStats stt{};
    CK(cudaMemcpyAsync(&total, ctx->set().total.p, 8, cudaMemcpyDeviceToHost, ctx->enc_stream));
    CK(cudaMemcpyAsync(&stt, ctx->set().stats.p, sizeof stt, cudaMemcpyDeviceToHost, ctx->enc_stream));
    CK(cudaStreamSynchronize(ctx->enc_stream));
    if (ctx->n_frames && !ctx->in_rerun && ctx->settled_seq != ctx->batch_seq) {
        ctx->settled_seq = ctx->batch_seq;
        // decisions inside the libm-log guard band: the host repeats them; if it decides otherwise the batch is encoded once more
        // with the host's decisions (the caller's PCM is still there: it must stay unchanged until the results are taken)
        size_t n_new = 0;
        int grc = guard_settle(ctx, ctx->set(), stt.log_ambiguous, &n_new);
        if (grc) return grc;
        if (n_new) {
            ctx->in_rerun = true;
            int rrc = run_batch(ctx, ctx->last_d_pcm, ctx->last_pcm_staged);
            if (!rrc) rrc = encode_result_impl(ctx, res, wait_md5);
            ctx->in_rerun = false; ctx->h_ovr.clear(); ctx->settled_seq = ctx->batch_seq;
            if (rrc) return rrc;
            res->log_guard_hits = ctx->guard_info[3];
            return 0;
        }
    }
    if (wait_md5 && ctx->set().busy) CK(cudaStreamSynchronize(ctx->set().side));     // the MD5 of this batch lives on the set's side stream
    res->total_bytes = total; res->n_frames = (uint32_t)ctx->n_frames; res->n_streams = (uint32_t)ctx->n_streams;
    res->log_guard_hits = ctx->in_rerun ? 0 : ctx->guard_info[3];       // decisions inside the band the host could not check (log full); 0 = settled
    res->d_arena = (const uint8_t*)ctx->set().arena.p; res->d_frame_off = (const uint64_t*)ctx->set().foff.p; res->d_frame_len = (const uint32_t*)ctx->set().flen.p;
    return 0;
}

extern "C" int flacb200_encode_fetch(flacb200_ctx* ctx, uint8_t* arena, size_t arena_cap, uint64_t* frame_off, uint32_t* frame_len,
                                     uint32_t* frame_samples, uint32_t* frame_stream, flacb200_stream_info* streams) {
    flacb200_enc_result r;
    int rc = flacb200_encode_result(ctx, &r);
    if (rc) return rc;
    cudaStream_t st = ctx->enc_stream;
    if (arena) {
        if (arena_cap < r.total_bytes) return fail(ctx, FLACB200_ERR_ARG, "arena too small");
        if (r.total_bytes) CK(cudaMemcpyAsync(arena, ctx->set().arena.p, r.total_bytes, cudaMemcpyDeviceToHost, st));
    }
    const int nf = ctx->n_frames;
    if (frame_off && nf) CK(cudaMemcpyAsync(frame_off, ctx->set().foff.p, sizeof(uint64_t) * nf, cudaMemcpyDeviceToHost, st));
    if (frame_len && nf) CK(cudaMemcpyAsync(frame_len, ctx->set().flen.p, sizeof(uint32_t) * nf, cudaMemcpyDeviceToHost, st));
    static_assert(sizeof(flacb200_stream_info) == sizeof(StreamInfoOut), "stream info layout");
    if (streams && ctx->n_streams) CK(cudaMemcpyAsync(streams, ctx->set().sinfo.p, sizeof(StreamInfoOut) * ctx->n_streams, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int f = 0; f < nf; f++) {
        if (frame_samples) frame_samples[f] = ctx->h_frames[f].blocksize;
        if (frame_stream) frame_stream[f] = ctx->h_frames[f].stream;
    }
    return 0;
}

extern "C" int flacb200_encode_fetch_trace(flacb200_ctx* ctx, void* plans, size_t plans_bytes, uint8_t* frame_ca, void* debug, size_t debug_bytes) {
    if (!ctx || !ctx->have_batch) return FLACB200_ERR_ARG;
    cudaSetDevice(ctx->device);
    const size_t np = sizeof(SubframePlan) * (size_t)ctx->n_frames * ctx->P.n_signals;
    const size_t nd = sizeof(SignalDebug) * (size_t)ctx->n_frames * ctx->P.n_signals;
    CK(cudaStreamSynchronize(ctx->enc_stream));
    if (plans) { if (plans_bytes < np) return fail(ctx, FLACB200_ERR_ARG, "plans buffer too small"); CK(cudaMemcpy(plans, ctx->d_plans.p, np, cudaMemcpyDeviceToHost)); }
    if (frame_ca) CK(cudaMemcpy(frame_ca, ctx->d_ca.p, (size_t)ctx->n_frames, cudaMemcpyDeviceToHost));
    if (debug) {
        if (!ctx->debug) return fail(ctx, FLACB200_ERR_ARG, "batch was not run with debug_trace");
        if (debug_bytes < nd) return fail(ctx, FLACB200_ERR_ARG, "debug buffer too small");
        CK(cudaMemcpy(debug, ctx->d_debug.p, nd, cudaMemcpyDeviceToHost));
    }
    return 0;
}

// Host -> host path (what a pyFLAC-style caller has: PCM in host memory, packed bytes wanted in host memory).
// Streams are cut into up to 12 chunks of whole streams; chunk c+1's H2D copy, chunk c's kernels and chunk c-1's
// D2H copy run concurrently on three CUDA streams, and the MD5 digests (serial chain per stream) are computed
// by host threads straight from the caller's PCM while the GPU encodes (md5_host.h; DESIGN.md "MD5 placement").
// The host arena is contiguous: stream images back to back in stream order; frame_off / streams[].byte_off index it.
extern "C" int flacb200_encode_batch_host(flacb200_ctx* ctx, const flacb200_enc_config* cfg, const void* pcm_host, uint64_t pcm_elems,
                                          uint32_t n_streams, const uint64_t* stream_off, const uint64_t* stream_samples,
                                          uint8_t* arena, size_t arena_cap, uint64_t* total_bytes, uint64_t* frame_off,
                                          uint32_t* frame_len, flacb200_stream_info* streams) {
    if (!ctx) return FLACB200_ERR_NO_DEVICE;
    if (!cfg || !pcm_host || !arena || (n_streams && (!stream_off || !stream_samples))) return fail(ctx, FLACB200_ERR_ARG, "null argument");
    cudaSetDevice(ctx->device);
    if (!ctx->in_rerun) ctx->h_ovr.clear();
    if (!ctx->in_rerun && fb_ctx_jobs_in_flight(ctx)) return fail(ctx, FLACB200_ERR_ARG, "collect the submitted batches before a synchronous host call");
    for (uint32_t s = 0; s < n_streams; s++)
        if (stream_off[s] + stream_samples[s] * cfg->channels > pcm_elems) return fail(ctx, FLACB200_ERR_ARG, "stream exceeds pcm buffer");
    if (ctx->prev_ca_pending || !same_layout(ctx, *cfg, n_streams, stream_off, stream_samples, nullptr)) {
        ctx->have_batch = false;
        CK(cudaStreamSynchronize(ctx->enc_stream));
        for (auto& S : ctx->sets) if (S.busy) { CK(cudaStreamSynchronize(S.side)); S.busy = false; }
        int rc = plan_batch(ctx, *cfg, n_streams, stream_off, stream_samples, nullptr);
        if (rc) return rc;
    }
    const EncParams P = params_for_set(ctx, ctx->set());
    const int nf = ctx->n_frames, ns = ctx->n_streams;
    const uint32_t cont = cfg->container_bytes;
    if (total_bytes) *total_bytes = 0;
    if (nf == 0) return 0;
    cudaStream_t st = ctx->enc_stream;
    CK(ctx->d_pcm.reserve((size_t)pcm_elems * cont + 64));
    CK(ctx->d_totals.reserve(sizeof(uint64_t) * flacb200_ctx::kMaxChunks));

    // ---- chunk boundaries (whole streams, roughly equal sample counts) ----
    int nchunks = 12;      // measured on B200 + PCIe 5: 4 -> 13.7 ms, 8 -> 12.7, 12 -> 12.3 per 491 MB batch (the H2D copy itself is ~11.2)
    if (const char* ev = getenv("FLACB200_CHUNKS")) { const int v = atoi(ev); if (v > 0 && v <= flacb200_ctx::kMaxChunks) nchunks = v; }
    if (nchunks > ns) nchunks = ns;
    std::vector<int> cs(nchunks + 1, 0);
    {
        uint64_t tot = 0; for (int s = 0; s < ns; s++) tot += stream_samples[s];
        uint64_t acc = 0; int c = 1;
        for (int s = 0; s < ns && c < nchunks; s++) {
            acc += stream_samples[s];
            if (acc * nchunks >= tot * c && s + 1 >= c) { cs[c++] = s + 1; }
        }
        for (; c <= nchunks; c++) cs[c] = ns;
        cs[nchunks] = ns;
    }

    const auto t_start = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
    std::atomic<uint64_t> md5_done_us{0};
    // ---- MD5 (one serial chain per stream): host threads hash the caller's buffer while the GPU encodes; when the host cannot
    // finish by the time the transfer does (few cores per rank), the streams of the first chunks -- the ones that reach HBM
    // first -- are hashed by md5_kernel on a side stream instead, and the split is re-balanced from the measured finish times ----
    const bool want_md5 = cfg->do_md5 != 0;
    std::vector<uint8_t> digests((size_t)ns * 16, 0);
    std::atomic<int> next_stream{0};
    bool pool_running = false;
    int gpu_chunks = 0, g_streams = 0;
    unsigned nt = 0;
    std::atomic<uint64_t>& gpu_md5_done_us = ctx->gpu_md5_done_us;  // stamped by a host function on the MD5 stream
    ctx->t_call = t_start;
    const uint32_t bytes_per = (P.bps + 7) / 8, chn = P.channels;
    const bool raw_bytes = (bytes_per == cont);                    // the container bytes are the hashed bytes
    const int G = fb::md5_mb16_available() ? 16 : 8;               // streams per SIMD pass (AVX-512 / AVX2)
    // Hashing competes with the H2D DMA for host memory bandwidth (measured: 8+ threads finish the MD5s in 6 ms but
    // stretch the 491 MB copy from 11 to 13.5 ms), so use just enough threads to finish when the transfer does:
    // threads = (bytes / calibrated per-thread rate) / (bytes / ~42 GB/s PCIe), within this rank's share of the host.
    static const double gbps_per_thread = [] {
        std::vector<uint8_t> buf(16u << 16, 0x5a);
        const uint8_t* d[16]; size_t l[16]; uint8_t dig[16][16];
        for (int i = 0; i < 16; i++) { d[i] = buf.data() + ((size_t)i << 16); l[i] = 1u << 16; }
        fb::md5_group16(d, l, 16, dig);
        const auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < 4; r++) fb::md5_group16(d, l, 16, dig);
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return sec > 0 ? 4.0 * buf.size() / sec / 1e9 : 4.0;
    }();
    const double total_mb = (double)pcm_elems * cont / 1e6, t_h2d_ms = total_mb / 44.0;
    if (want_md5) {
        const unsigned budget = host_thread_budget();
        const unsigned max_workers = budget > 2 ? budget - 1 : budget;      // the calling thread sleeps in blocking waits most of the time
        unsigned want = (unsigned)(42.0 / gbps_per_thread + 0.5);
        if (want < 1) want = 1;
        if (want > max_workers) want = max_workers;
        if (const char* ev = getenv("FLACB200_MD5_THREADS")) { const int v = atoi(ev); if (v > 0) want = (unsigned)v; }
        nt = std::min<unsigned>(std::min<unsigned>(want, 64u), (unsigned)((ns + G - 1) / G));
        const bool gpu_fast = raw_bytes || (bytes_per == 3 && cont == 4);   // md5_kernel's staged paths
        uint64_t max_stream_bytes = 0;
        for (int s = 0; s < ns; s++) max_stream_bytes = std::max<uint64_t>(max_stream_bytes, stream_samples[s] * chn * bytes_per);
        const double host_all_ms = total_mb / ((double)nt * gbps_per_thread);   // MB / (GB/s) = ms
        const double gpu_chain_ms = (double)max_stream_bytes * 8.3e-6;      // measured: 15.9 ms per 1.92 MB stream, whatever the stream count
        const uint64_t key = (uint64_t)ns * 1000003ull ^ (uint64_t)pcm_elems * 31ull ^ ((uint64_t)nt << 48) ^ ((uint64_t)P.bps << 56) ^ (uint64_t)nchunks;
        if (ctx->md5_split_key != key || ctx->md5_gpu_chunks < 0) {
            int best = 0; double best_t = host_all_ms;
            if (gpu_fast && host_all_ms > t_h2d_ms * 1.1)
                for (int g = 1; g <= nchunks; g++) {
                    const double fr = (double)cs[g] / (double)ns;
                    const double t = std::max(fr * t_h2d_ms + gpu_chain_ms, (1.0 - fr) * host_all_ms);
                    if (t < best_t - 0.2) { best_t = t; best = g; }
                }
            ctx->md5_gpu_chunks = best; ctx->md5_split_key = key;
        }
        gpu_chunks = gpu_fast ? std::min(ctx->md5_gpu_chunks, nchunks) : 0;
        if (const char* ev = getenv("FLACB200_MD5_GPU_CHUNKS")) { const int v = atoi(ev); if (v >= 0 && gpu_fast) gpu_chunks = std::min(v, nchunks); }
        g_streams = cs[gpu_chunks];
        next_stream.store(g_streams);
        if (g_streams > 0 && (size_t)g_streams * 16 > ctx->h_digests_cap) {
            if (ctx->h_digests) cudaFreeHost(ctx->h_digests);
            ctx->h_digests = nullptr; ctx->h_digests_cap = 0;
            CK(cudaHostAlloc((void**)&ctx->h_digests, (size_t)ns * 16 + 64, cudaHostAllocDefault));
            ctx->h_digests_cap = (size_t)ns * 16 + 64;
        }
        gpu_md5_done_us.store(0);
        if (g_streams < ns) {
            ctx->pool.start((int)nt, [&, bytes_per, chn, raw_bytes, G]() {
                for (;;) {
                    const int g = next_stream.fetch_add(G);
                    if (g >= ns) {
                        const uint64_t now = (uint64_t)(since() * 1000.0); uint64_t prev = md5_done_us.load();
                        while (prev < now && !md5_done_us.compare_exchange_weak(prev, now)) {}
                        break;
                    }
                    const int n = std::min(G, ns - g);
                    if (raw_bytes) {
                        const uint8_t* d[16]; size_t l[16]; uint8_t dig[16][16];
                        for (int i = 0; i < n; i++) { d[i] = (const uint8_t*)pcm_host + stream_off[g + i] * cont; l[i] = (size_t)stream_samples[g + i] * chn * cont; }
                        fb::md5_group16(d, l, n, dig);
                        for (int i = 0; i < n; i++) memcpy(&digests[(size_t)(g + i) * 16], dig[i], 16);
                    } else {
                        for (int i = 0; i < n; i++) {
                            fb::Md5 m; m.init();
                            m.update_samples((const uint8_t*)pcm_host + stream_off[g + i] * cont, (size_t)stream_samples[g + i] * chn, cont, bytes_per);
                            m.final(&digests[(size_t)(g + i) * 16]);
                        }
                    }
                }
            });
            pool_running = true;
        }
    }
    auto join_workers = [&]() { if (pool_running) { ctx->pool.wait(); pool_running = false; } };

    // ---- enqueue: H2D per chunk, then kernels per chunk ----
    auto bail = [&](int code) { join_workers(); return code; };
#define CKJ(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return bail(fail(ctx, FLACB200_ERR_CUDA, #call, e_)); } while (0)
    { int wrc = wait_all_sets(ctx, st); if (wrc) return bail(wrc); }
    CKJ(cudaMemsetAsync(ctx->set().stats.p, 0, sizeof(EncStats), st));
    bool monotonic = true;
    for (int s = 1; s < ns; s++) if (stream_off[s] < stream_off[s - 1] + stream_samples[s - 1] * P.channels) { monotonic = false; break; }
    if (g_streams > 0) {
        // ONE md5_kernel launch for all GPU-hashed streams, ahead of the data: its warps wait for the arrival flag of their chunk
        // (set on the copy stream behind each chunk), so the chains start as their bytes land and run side by side
        CKJ(ctx->d_flags.reserve(sizeof(uint32_t) * flacb200_ctx::kMaxChunks));
        CKJ(cudaMemsetAsync(ctx->d_flags.p, 0, sizeof(uint32_t) * flacb200_ctx::kMaxChunks, ctx->h2d_stream));
        CKJ(cudaEventRecord(ctx->ev_fork, ctx->h2d_stream));
        CKJ(cudaStreamWaitEvent(ctx->md5_stream, ctx->ev_fork, 0));
        Md5Gate gate; gate.flags = (const uint32_t*)ctx->d_flags.p; gate.nchunks = gpu_chunks;
        for (int c = 0; c <= nchunks && c < 17; c++) gate.cs[c] = cs[c];
        launch_md5_gated(ctx->d_pcm.p, cont, (const uint64_t*)ctx->d_soff.p, (const uint64_t*)ctx->d_ssamples.p, g_streams, P.channels, P.bps,
                         (uint8_t*)ctx->set().md5.p, gate, ctx->md5_stream);
        ctx->launches++;
    }
    for (int c = 0; c < nchunks; c++) {
        const int s0 = cs[c], s1 = cs[c + 1];
        if (s1 > s0) {
            if (monotonic) {
                const uint64_t e0 = stream_off[s0], e1 = stream_off[s1 - 1] + stream_samples[s1 - 1] * P.channels;
                CKJ(cudaMemcpyAsync((uint8_t*)ctx->d_pcm.p + e0 * cont, (const uint8_t*)pcm_host + e0 * cont, (e1 - e0) * cont, cudaMemcpyHostToDevice, ctx->h2d_stream));
            } else {
                for (int s = s0; s < s1; s++)
                    CKJ(cudaMemcpyAsync((uint8_t*)ctx->d_pcm.p + stream_off[s] * cont, (const uint8_t*)pcm_host + stream_off[s] * cont,
                                        stream_samples[s] * P.channels * cont, cudaMemcpyHostToDevice, ctx->h2d_stream));
            }
        }
        if (c < gpu_chunks) CKJ(cudaMemsetAsync((uint32_t*)ctx->d_flags.p + c, 0xff, sizeof(uint32_t), ctx->h2d_stream));
        CKJ(cudaEventRecord(ctx->ev_h2d[c], ctx->h2d_stream));
    }
    const uint32_t pro = cfg->write_prologue ? (uint32_t)kStreamPrologueBytes : 0u;
    std::vector<uint64_t> dev_base(nchunks + 1, 0);
    for (int c = 0; c < nchunks; c++) {
        const int s0 = cs[c], s1 = cs[c + 1];
        const int f0 = s0 < ns ? (int)ctx->h_stream_first[s0] : nf, f1 = s1 < ns ? (int)ctx->h_stream_first[s1] : nf;
        const int cnf = f1 - f0, cns = s1 - s0;
        dev_base[c + 1] = dev_base[c] + (((uint64_t)cnf * ctx->scratch_stride + (uint64_t)cns * kStreamPrologueBytes + 255) / 256) * 256;
        CKJ(cudaStreamWaitEvent(st, ctx->ev_h2d[c], 0));
        if (cnf > 0) {
            int n_an = 0;
            if (ctx->use_fused) {
                launch_autoc_unshifted(ctx->d_pcm.p, (const FrameDesc*)ctx->d_frames.p + f0, (const float*)ctx->d_windows.p, P, cnf,
                                       (uint8_t*)ctx->d_work.p + (size_t)f0 * analyze_work_stride(P), st);
                launch_fused(ctx->d_pcm.p, (const FrameDesc*)ctx->d_frames.p + f0, P, cnf, (const uint8_t*)ctx->d_work.p + (size_t)f0 * analyze_work_stride(P) + 64,
                             analyze_work_stride(P), (SubframePlan*)ctx->d_plans.p + (size_t)f0 * P.n_signals, (uint8_t*)ctx->d_ca.p + f0, (EncStats*)ctx->set().stats.p,
                             (uint8_t*)ctx->d_scratch.p + (size_t)f0 * ctx->scratch_stride, ctx->scratch_stride, (uint32_t*)ctx->set().flen.p + f0, st, nullptr);
                n_an = (P.max_lpc_order > 0 ? 1 : 0) + 1;
            } else {
            n_an = launch_analyze(ctx->d_pcm.p, (const FrameDesc*)ctx->d_frames.p + f0, (const float*)ctx->d_windows.p, P, cnf,
                                            (SubframePlan*)ctx->d_plans.p + (size_t)f0 * P.n_signals, (uint8_t*)ctx->d_ca.p + f0, nullptr, (EncStats*)ctx->set().stats.p,
                                            analyze_smem_bytes(P), (uint8_t*)ctx->d_work.p + (size_t)f0 * analyze_work_stride(P), st, nullptr);
            launch_pack(ctx->d_pcm.p, (const FrameDesc*)ctx->d_frames.p + f0, P, cnf, (const SubframePlan*)ctx->d_plans.p + (size_t)f0 * P.n_signals,
                        (const uint8_t*)ctx->d_ca.p + f0, (uint8_t*)ctx->d_scratch.p + (size_t)f0 * ctx->scratch_stride, ctx->scratch_stride,
                        (uint32_t*)ctx->set().flen.p + f0, st);
            }
            launch_layout((const uint32_t*)ctx->set().flen.p + f0, (const FrameDesc*)ctx->d_frames.p + f0, cnf, pro, dev_base[c],
                          (uint64_t*)ctx->set().foff.p + f0, (uint64_t*)ctx->d_totals.p + c, st);
            launch_compact((const uint8_t*)ctx->d_scratch.p + (size_t)f0 * ctx->scratch_stride, ctx->scratch_stride, (const uint32_t*)ctx->set().flen.p + f0,
                           (const uint64_t*)ctx->set().foff.p + f0, (uint8_t*)ctx->set().arena.p, cnf, st);
            launch_finalize((const uint32_t*)ctx->set().flen.p, (const uint64_t*)ctx->set().foff.p, (const uint32_t*)ctx->d_sfirst.p + s0,
                            (const uint32_t*)ctx->d_snframes.p + s0, (const uint64_t*)ctx->d_ssamples.p + s0, nullptr, cns, P, pro ? 1u : 0u,
                            (uint8_t*)ctx->set().arena.p, (StreamInfoOut*)ctx->set().sinfo.p + s0, st);
            ctx->launches += 4 + n_an;
        } else {
            CKJ(cudaMemsetAsync((uint64_t*)ctx->d_totals.p + c, 0, 8, st));
        }
        CKJ(cudaMemcpyAsync(ctx->h_totals + c, (uint64_t*)ctx->d_totals.p + c, 8, cudaMemcpyDeviceToHost, st));
        CKJ(cudaEventRecord(ctx->ev_done[c], st));
    }
    if (dev_base[nchunks] > ctx->set().arena.cap) return bail(fail(ctx, FLACB200_ERR_CUDA, "device arena too small"));
    if (g_streams > 0) {
        CKJ(cudaMemcpyAsync(ctx->h_digests, ctx->set().md5.p, (size_t)g_streams * 16, cudaMemcpyDeviceToHost, ctx->md5_stream));
        CKJ(cudaLaunchHostFunc(ctx->md5_stream, [](void* p) {
            flacb200_ctx* c = (flacb200_ctx*)p;
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - c->t_call).count();
            c->gpu_md5_done_us.store((uint64_t)(ms * 1000.0));
        }, ctx));
        CKJ(cudaEventRecord(ctx->ev_md5, ctx->md5_stream));
    }

    ctx->e2e_ms[1] = since();
    // ---- drain: as each chunk finishes, copy exactly its bytes to the next free spot of the host arena ----
    std::vector<uint64_t> host_base(nchunks + 1, 0);
    for (int c = 0; c < nchunks; c++) {
        CKJ(cudaEventSynchronize(ctx->ev_done[c]));
        const uint64_t bytes = ctx->h_totals[c];
        if (host_base[c] + bytes > arena_cap) return bail(fail(ctx, FLACB200_ERR_ARG, "arena too small"));
        if (bytes) CKJ(cudaMemcpyAsync(arena + host_base[c], (const uint8_t*)ctx->set().arena.p + dev_base[c], bytes, cudaMemcpyDeviceToHost, ctx->d2h_stream));
        host_base[c + 1] = host_base[c] + bytes;
    }
    ctx->e2e_ms[2] = since();
    std::vector<uint64_t> tmp_off;
    uint64_t* foff_host = frame_off;
    if (!foff_host) { tmp_off.resize(nf); foff_host = tmp_off.data(); }
    CKJ(cudaMemcpyAsync(foff_host, ctx->set().foff.p, sizeof(uint64_t) * nf, cudaMemcpyDeviceToHost, ctx->d2h_stream));
    if (frame_len) CKJ(cudaMemcpyAsync(frame_len, ctx->set().flen.p, sizeof(uint32_t) * nf, cudaMemcpyDeviceToHost, ctx->d2h_stream));
    std::vector<flacb200_stream_info> tmp_info;
    flacb200_stream_info* info_host = streams;
    if (!info_host) { tmp_info.resize(ns); info_host = tmp_info.data(); }
    CKJ(cudaMemcpyAsync(info_host, ctx->set().sinfo.p, sizeof(StreamInfoOut) * ns, cudaMemcpyDeviceToHost, ctx->d2h_stream));
    CKJ(cudaMemcpyAsync(ctx->h_stats, ctx->set().stats.p, sizeof(EncStats), cudaMemcpyDeviceToHost, ctx->d2h_stream));
    CKJ(cudaEventRecord(ctx->ev_d2h, ctx->d2h_stream));
    CKJ(cudaEventSynchronize(ctx->ev_d2h));
    CKJ(cudaGetLastError());
    ctx->e2e_ms[3] = since();
    join_workers();
    if (g_streams > 0) {
        CKJ(cudaEventSynchronize(ctx->ev_md5));
        memcpy(digests.data(), ctx->h_digests, (size_t)g_streams * 16);
    }
#undef CKJ
    if (!ctx->in_rerun) {
        // libm-log guard: decisions inside the band are repeated on the host; a different host decision encodes the batch again
        size_t n_new = 0;
        int grc = guard_settle(ctx, ctx->set(), ctx->h_stats->log_ambiguous, &n_new);
        if (grc) return grc;
        if (n_new) {
            ctx->in_rerun = true;
            const int rrc = flacb200_encode_batch_host(ctx, cfg, pcm_host, pcm_elems, n_streams, stream_off, stream_samples, arena, arena_cap,
                                                       total_bytes, frame_off, frame_len, streams);
            ctx->in_rerun = false; ctx->h_ovr.clear();
            return rrc;
        }
    }
    ctx->e2e_ms[4] = since();
    ctx->e2e_ms[0] = (double)md5_done_us.load() / 1000.0;       // when the last MD5 worker ran out of streams
    ctx->e2e_ms[6] = (double)gpu_md5_done_us.load() / 1000.0;   // when the GPU's digests had reached the host
    ctx->e2e_ms[7] = (double)g_streams; ctx->e2e_ms[8] = (double)nt; ctx->e2e_ms[9] = (double)nchunks;
    if (want_md5 && !getenv("FLACB200_MD5_GPU_CHUNKS")) {
        // re-balance for the next call of this shape: move one chunk across when the measured finish times say it pays
        const double t_host = ctx->e2e_ms[0], t_gpu = ctx->e2e_ms[6], t_rest = ctx->e2e_ms[3];
        const double now_max = std::max(std::max(t_host, t_gpu), t_rest);
        const bool gpu_fast = raw_bytes || (bytes_per == 3 && cont == 4);
        const double step = t_h2d_ms / nchunks;
        if (gpu_fast && gpu_chunks < nchunks && g_streams < ns && t_host >= now_max - 1e-9) {
            const double rem = (double)(nchunks - gpu_chunks);
            const double new_host = t_host * (rem - 1.0) / rem;
            uint64_t msb = 0; for (int s = 0; s < ns; s++) msb = std::max<uint64_t>(msb, stream_samples[s] * chn * bytes_per);
            const double new_gpu = (gpu_chunks > 0 ? t_gpu : (double)msb * 8.3e-6) + step;
            if (std::max(std::max(new_host, new_gpu), t_rest) < now_max - 0.3) ctx->md5_gpu_chunks = gpu_chunks + 1;
        } else if (gpu_chunks > 0 && t_gpu >= now_max - 1e-9) {
            const double rem = (double)(nchunks - gpu_chunks);
            const double new_host = rem > 0 ? t_host * (rem + 1.0) / rem : total_mb / ((double)nt * gbps_per_thread) / nchunks;
            const double new_gpu = gpu_chunks > 1 ? t_gpu - step : 0.0;
            if (std::max(std::max(new_host, new_gpu), t_rest) < now_max - 0.3) ctx->md5_gpu_chunks = gpu_chunks - 1;
        }
    }
    // device-arena offsets -> host-arena offsets; MD5 digests into STREAMINFO
    for (int c = 0; c < nchunks; c++) {
        const int s0 = cs[c], s1 = cs[c + 1];
        const int f0 = s0 < ns ? (int)ctx->h_stream_first[s0] : nf, f1 = s1 < ns ? (int)ctx->h_stream_first[s1] : nf;
        for (int f = f0; f < f1; f++) foff_host[f] = foff_host[f] - dev_base[c] + host_base[c];
        for (int s = s0; s < s1; s++) {
            flacb200_stream_info& si = info_host[s];
            if (si.n_frames) si.byte_off = si.byte_off - dev_base[c] + host_base[c];
            if (want_md5) {
                memcpy(si.md5, &digests[(size_t)s * 16], 16);
                if (pro && si.n_frames) memcpy(arena + si.byte_off + 26, si.md5, 16);
            }
        }
    }
    ctx->e2e_last_bytes = host_base[nchunks];
    ctx->e2e_ms[5] = since();
    if (total_bytes) *total_bytes = host_base[nchunks];
    return 0;
}

// ---------------------------------------------------------------- pipelined host -> host encode ----
// flacb200_encode_host_submit / _collect: the same work as flacb200_encode_batch_host, but the call returns once everything is
// enqueued and up to kJobs batches are in flight, so a caller with a queue of batches (a library of files) keeps the PCIe link
// busy in both directions all the time and nothing waits for a serial MD5 chain: the digests come from md5_kernel, chunk by chunk
// as the bytes land in HBM, on a side stream that overlaps the following batches (the host threads -- and the second read of the PCM
// from host memory, which is what limits many ranks on one host -- are not needed).  Each job owns a PCM staging buffer, an
// output set, its events and a drain thread that sleeps in blocking waits and issues each chunk's D2H copy when its size is known.
struct flacb200_ctx::HostJob {
    bool active = false;
    int rc = 0; std::string err;
    DevBuf d_pcm, d_totals, d_flags;
    cudaEvent_t ev_h2d[kMaxChunks] = {nullptr}, ev_done[kMaxChunks] = {nullptr}, ev_md5 = nullptr, ev_d2h = nullptr, ev_flags = nullptr;
    uint64_t* h_totals = nullptr; uint8_t* h_digests = nullptr; size_t h_digests_cap = 0;
    EncStats* h_stats = nullptr; uint64_t pcm_elems = 0;
    int set_idx = 0, nchunks = 0, nf = 0, ns = 0;
    bool want_md5 = false, gpu_md5 = false; uint32_t pro = 0;
    std::vector<int> cs; std::vector<uint64_t> dev_base;
    std::vector<uint32_t> stream_first;
    const void* pcm_host = nullptr; uint32_t cont = 0, chn = 0, bytes_per = 0;
    std::vector<uint64_t> s_off, s_smp;
    uint8_t* arena = nullptr; size_t arena_cap = 0; uint64_t* frame_off = nullptr; uint32_t* frame_len = nullptr; flacb200_stream_info* streams = nullptr;
    uint64_t total = 0;
    std::thread drain;
    void init() {
        for (auto& e : ev_h2d) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        for (auto& e : ev_done) cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync);
        cudaEventCreateWithFlags(&ev_md5, cudaEventDisableTiming | cudaEventBlockingSync);
        cudaEventCreateWithFlags(&ev_d2h, cudaEventDisableTiming | cudaEventBlockingSync);
        cudaEventCreateWithFlags(&ev_flags, cudaEventDisableTiming);
        cudaHostAlloc((void**)&h_totals, sizeof(uint64_t) * kMaxChunks, cudaHostAllocDefault);
        cudaHostAlloc((void**)&h_stats, sizeof(EncStats), cudaHostAllocDefault);
    }
    void destroy() {
        if (drain.joinable()) drain.join();
        for (auto& e : ev_h2d) if (e) cudaEventDestroy(e);
        for (auto& e : ev_done) if (e) cudaEventDestroy(e);
        if (ev_md5) cudaEventDestroy(ev_md5);
        if (ev_d2h) cudaEventDestroy(ev_d2h);
        if (ev_flags) cudaEventDestroy(ev_flags);
        if (h_totals) cudaFreeHost(h_totals);
        if (h_stats) cudaFreeHost(h_stats);
        if (h_digests) cudaFreeHost(h_digests);
        d_pcm.release(); d_totals.release(); d_flags.release();
    }
};

void fb_ctx_free_jobs(flacb200_ctx* ctx) {
    for (auto& j : ctx->jobs) if (j) { j->destroy(); delete j; j = nullptr; }
}

bool fb_ctx_jobs_in_flight(flacb200_ctx* ctx) { for (auto& j : ctx->jobs) if (j && j->active) return true; return false; }

static void chunk_bounds(int ns, const uint64_t* stream_samples, int nchunks, std::vector<int>& cs) {
    cs.assign(nchunks + 1, 0);
    uint64_t tot = 0; for (int s = 0; s < ns; s++) tot += stream_samples[s];
    uint64_t acc = 0; int c = 1;
    for (int s = 0; s < ns && c < nchunks; s++) {
        acc += stream_samples[s];
        if (acc * nchunks >= tot * c && s + 1 >= c) { cs[c++] = s + 1; }
    }
    for (; c <= nchunks; c++) cs[c] = ns;
    cs[nchunks] = ns;
}

// the drain thread of one job: D2H of every chunk as soon as its byte count is known, then the index, then the digests
static void drain_job(flacb200_ctx* ctx, flacb200_ctx::HostJob* J) {
    cudaSetDevice(ctx->device);
    flacb200_ctx::OutSet& S = ctx->sets[J->set_idx];
    auto failj = [&](const char* what, cudaError_t e) { J->rc = FLACB200_ERR_CUDA; J->err = std::string(what) + ": " + cudaGetErrorString(e); };
#define CKD(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { failj(#call, e_); return; } } while (0)
    const int nchunks = J->nchunks, nf = J->nf, ns = J->ns;
    std::vector<uint64_t> host_base(nchunks + 1, 0);
    for (int c = 0; c < nchunks; c++) {
        CKD(cudaEventSynchronize(J->ev_done[c]));
        const uint64_t bytes = J->h_totals[c];
        if (host_base[c] + bytes > J->arena_cap) { J->rc = FLACB200_ERR_ARG; J->err = "arena too small"; return; }
        if (bytes) CKD(cudaMemcpyAsync(J->arena + host_base[c], (const uint8_t*)S.arena.p + J->dev_base[c], bytes, cudaMemcpyDeviceToHost, ctx->d2h_stream));
        host_base[c + 1] = host_base[c] + bytes;
    }
    std::vector<uint64_t> tmp_off; std::vector<flacb200_stream_info> tmp_info;
    uint64_t* foff_host = J->frame_off;
    if (!foff_host) { tmp_off.resize(nf); foff_host = tmp_off.data(); }
    CKD(cudaMemcpyAsync(foff_host, S.foff.p, sizeof(uint64_t) * nf, cudaMemcpyDeviceToHost, ctx->d2h_stream));
    if (J->frame_len) CKD(cudaMemcpyAsync(J->frame_len, S.flen.p, sizeof(uint32_t) * nf, cudaMemcpyDeviceToHost, ctx->d2h_stream));
    flacb200_stream_info* info_host = J->streams;
    if (!info_host) { tmp_info.resize(ns); info_host = tmp_info.data(); }
    CKD(cudaMemcpyAsync(info_host, S.sinfo.p, sizeof(StreamInfoOut) * ns, cudaMemcpyDeviceToHost, ctx->d2h_stream));
    CKD(cudaMemcpyAsync(J->h_stats, S.stats.p, sizeof(EncStats), cudaMemcpyDeviceToHost, ctx->d2h_stream));
    CKD(cudaEventRecord(J->ev_d2h, ctx->d2h_stream));
    CKD(cudaEventSynchronize(J->ev_d2h));
    std::vector<uint8_t> host_dig;
    const uint8_t* dig = nullptr;
    if (J->want_md5) {
        if (J->gpu_md5) { CKD(cudaEventSynchronize(J->ev_md5)); dig = J->h_digests; }
        else {
            // sample sizes md5_kernel has no staged path for: hashed here, sixteen streams per SIMD pass where the bytes allow
            host_dig.assign((size_t)ns * 16, 0);
            for (int s = 0; s < ns; s++) {
                fb::Md5 m; m.init();
                m.update_samples((const uint8_t*)J->pcm_host + J->s_off[s] * J->cont, (size_t)J->s_smp[s] * J->chn, J->cont, J->bytes_per);
                m.final(&host_dig[(size_t)s * 16]);
            }
            dig = host_dig.data();
        }
    }
#undef CKD
    for (int c = 0; c < nchunks; c++) {
        const int s0 = J->cs[c], s1 = J->cs[c + 1];
        const int f0 = s0 < ns ? (int)J->stream_first[s0] : nf, f1 = s1 < ns ? (int)J->stream_first[s1] : nf;
        for (int f = f0; f < f1; f++) foff_host[f] = foff_host[f] - J->dev_base[c] + host_base[c];
        for (int s = s0; s < s1; s++) {
            flacb200_stream_info& si = info_host[s];
            if (si.n_frames) si.byte_off = si.byte_off - J->dev_base[c] + host_base[c];
            if (dig) {
                memcpy(si.md5, dig + (size_t)s * 16, 16);
                if (J->pro && si.n_frames) memcpy(J->arena + si.byte_off + 26, si.md5, 16);
            }
        }
    }
    J->total = host_base[nchunks];
}

extern "C" int flacb200_encode_host_submit(flacb200_ctx* ctx, const flacb200_enc_config* cfg, const void* pcm_host, uint64_t pcm_elems,
                                           uint32_t n_streams, const uint64_t* stream_off, const uint64_t* stream_samples,
                                           uint8_t* arena, size_t arena_cap, uint64_t* frame_off, uint32_t* frame_len,
                                           flacb200_stream_info* streams, int* ticket) {
    if (!ctx) return FLACB200_ERR_NO_DEVICE;
    if (!cfg || !pcm_host || !arena || !ticket || (n_streams && (!stream_off || !stream_samples))) return fail(ctx, FLACB200_ERR_ARG, "null argument");
    cudaSetDevice(ctx->device);
    for (uint32_t s = 0; s < n_streams; s++)
        if (stream_off[s] + stream_samples[s] * cfg->channels > pcm_elems) return fail(ctx, FLACB200_ERR_ARG, "stream exceeds pcm buffer");
    const int slot = ctx->next_job;
    if (!ctx->jobs[slot]) { ctx->jobs[slot] = new flacb200_ctx::HostJob(); ctx->jobs[slot]->init(); }
    flacb200_ctx::HostJob* J = ctx->jobs[slot];
    if (J->active) return fail(ctx, FLACB200_ERR_ARG, "too many batches in flight: collect the oldest ticket first");
    if (ctx->prev_ca_pending || !same_layout(ctx, *cfg, n_streams, stream_off, stream_samples, nullptr)) {
        // the frame table on the device is shared by the batches in flight: a new layout waits for them
        for (auto& j : ctx->jobs) if (j && j->active) return fail(ctx, FLACB200_ERR_ARG, "collect the batches in flight before submitting a different layout");
        ctx->have_batch = false;
        CK(cudaStreamSynchronize(ctx->enc_stream));
        for (auto& S : ctx->sets) if (S.busy) { CK(cudaStreamSynchronize(S.side)); S.busy = false; }
        int rc = plan_batch(ctx, *cfg, n_streams, stream_off, stream_samples, nullptr);
        if (rc) return rc;
    }
    ctx->h_ovr.clear();
    const int nf = ctx->n_frames, ns = ctx->n_streams;
    const uint32_t cont = cfg->container_bytes;
    cudaStream_t st = ctx->enc_stream;
    J->rc = 0; J->err.clear(); J->total = 0; J->pcm_elems = pcm_elems; J->h_stats->log_ambiguous = 0;
    const EncParams P = params_for_set(ctx, ctx->sets[(ctx->cur + 1) % flacb200_ctx::kSets]);
    J->nf = nf; J->ns = ns; J->arena = arena; J->arena_cap = arena_cap; J->frame_off = frame_off; J->frame_len = frame_len; J->streams = streams;
    J->pcm_host = pcm_host; J->cont = cont; J->chn = P.channels; J->bytes_per = (P.bps + 7) / 8;
    J->want_md5 = cfg->do_md5 != 0;
    J->gpu_md5 = J->want_md5 && (J->bytes_per == cont || (J->bytes_per == 3 && cont == 4));
    J->pro = cfg->write_prologue ? (uint32_t)kStreamPrologueBytes : 0u;
    J->nchunks = 0;
    *ticket = slot;
    if (nf == 0) { J->active = true; ctx->next_job = (slot + 1) % flacb200_ctx::kJobs; return 0; }
    CK(J->d_pcm.reserve((size_t)pcm_elems * cont + 64));
    CK(J->d_totals.reserve(sizeof(uint64_t) * flacb200_ctx::kMaxChunks));
    if (J->gpu_md5 && (size_t)ns * 16 > J->h_digests_cap) {
        if (J->h_digests) cudaFreeHost(J->h_digests);
        J->h_digests = nullptr; J->h_digests_cap = 0;
        CK(cudaHostAlloc((void**)&J->h_digests, (size_t)ns * 16 + 64, cudaHostAllocDefault));
        J->h_digests_cap = (size_t)ns * 16 + 64;
    }
    if (!J->gpu_md5 && J->want_md5) { J->s_off.assign(stream_off, stream_off + ns); J->s_smp.assign(stream_samples, stream_samples + ns); }
    int nchunks = 12;
    if (const char* ev = getenv("FLACB200_CHUNKS")) { const int v = atoi(ev); if (v > 0 && v <= flacb200_ctx::kMaxChunks) nchunks = v; }
    if (nchunks > ns) nchunks = ns;
    J->nchunks = nchunks;
    chunk_bounds(ns, stream_samples, nchunks, J->cs);
    const std::vector<int>& cs = J->cs;
    J->stream_first = ctx->h_stream_first;
    // the output set: rotate like the device-resident path; its previous user (MD5 patch on the side stream) must be done
    ctx->cur = (ctx->cur + 1) % flacb200_ctx::kSets;
    J->set_idx = ctx->cur;
    flacb200_ctx::OutSet& S = ctx->sets[J->set_idx];
    if (S.busy) { CK(cudaStreamWaitEvent(st, S.ev_free, 0)); S.busy = false; }
    CK(cudaMemsetAsync(S.stats.p, 0, sizeof(EncStats), st));
    bool monotonic = true;
    for (int s = 1; s < ns; s++) if (stream_off[s] < stream_off[s - 1] + stream_samples[s - 1] * P.channels) { monotonic = false; break; }
    if (J->gpu_md5) {
        // ONE md5_kernel launch per batch, ahead of the data (see Md5Gate): every chain starts when its chunk has landed
        CK(J->d_flags.reserve(sizeof(uint32_t) * flacb200_ctx::kMaxChunks));
        CK(cudaMemsetAsync(J->d_flags.p, 0, sizeof(uint32_t) * flacb200_ctx::kMaxChunks, ctx->h2d_stream));
        CK(cudaEventRecord(J->ev_flags, ctx->h2d_stream));
        CK(cudaStreamWaitEvent(S.side, J->ev_flags, 0));
        Md5Gate gate; gate.flags = (const uint32_t*)J->d_flags.p; gate.nchunks = nchunks;
        for (int c = 0; c <= nchunks && c < 17; c++) gate.cs[c] = cs[c];
        launch_md5_gated(J->d_pcm.p, cont, (const uint64_t*)ctx->d_soff.p, (const uint64_t*)ctx->d_ssamples.p, ns, P.channels, P.bps, (uint8_t*)S.md5.p, gate, S.side);
        ctx->launches++;
    }
    for (int c = 0; c < nchunks; c++) {
        const int s0 = cs[c], s1 = cs[c + 1];
        if (s1 > s0) {
            if (monotonic) {
                const uint64_t e0 = stream_off[s0], e1 = stream_off[s1 - 1] + stream_samples[s1 - 1] * P.channels;
                CK(cudaMemcpyAsync((uint8_t*)J->d_pcm.p + e0 * cont, (const uint8_t*)pcm_host + e0 * cont, (e1 - e0) * cont, cudaMemcpyHostToDevice, ctx->h2d_stream));
            } else {
                for (int s = s0; s < s1; s++)
                    CK(cudaMemcpyAsync((uint8_t*)J->d_pcm.p + stream_off[s] * cont, (const uint8_t*)pcm_host + stream_off[s] * cont,
                                       stream_samples[s] * P.channels * cont, cudaMemcpyHostToDevice, ctx->h2d_stream));
            }
        }
        if (J->gpu_md5) CK(cudaMemsetAsync((uint32_t*)J->d_flags.p + c, 0xff, sizeof(uint32_t), ctx->h2d_stream));
        CK(cudaEventRecord(J->ev_h2d[c], ctx->h2d_stream));
    }
    J->dev_base.assign(nchunks + 1, 0);
    for (int c = 0; c < nchunks; c++) {
        const int s0 = cs[c], s1 = cs[c + 1];
        const int f0 = s0 < ns ? (int)ctx->h_stream_first[s0] : nf, f1 = s1 < ns ? (int)ctx->h_stream_first[s1] : nf;
        const int cnf = f1 - f0, cns = s1 - s0;
        J->dev_base[c + 1] = J->dev_base[c] + (((uint64_t)cnf * ctx->scratch_stride + (uint64_t)cns * kStreamPrologueBytes + 255) / 256) * 256;
        CK(cudaStreamWaitEvent(st, J->ev_h2d[c], 0));
        if (cnf > 0) {
            int n_an = 0;
            if (ctx->use_fused) {
                launch_autoc_unshifted(J->d_pcm.p, (const FrameDesc*)ctx->d_frames.p + f0, (const float*)ctx->d_windows.p, P, cnf,
                                       (uint8_t*)ctx->d_work.p + (size_t)f0 * analyze_work_stride(P), st);
                launch_fused(J->d_pcm.p, (const FrameDesc*)ctx->d_frames.p + f0, P, cnf, (const uint8_t*)ctx->d_work.p + (size_t)f0 * analyze_work_stride(P) + 64,
                             analyze_work_stride(P), (SubframePlan*)ctx->d_plans.p + (size_t)f0 * P.n_signals, (uint8_t*)ctx->d_ca.p + f0, (EncStats*)S.stats.p,
                             (uint8_t*)ctx->d_scratch.p + (size_t)f0 * ctx->scratch_stride, ctx->scratch_stride, (uint32_t*)S.flen.p + f0, st, nullptr);
                n_an = (P.max_lpc_order > 0 ? 1 : 0) + 1;
            } else {
                n_an = launch_analyze(J->d_pcm.p, (const FrameDesc*)ctx->d_frames.p + f0, (const float*)ctx->d_windows.p, P, cnf,
                                      (SubframePlan*)ctx->d_plans.p + (size_t)f0 * P.n_signals, (uint8_t*)ctx->d_ca.p + f0, nullptr, (EncStats*)S.stats.p,
                                      analyze_smem_bytes(P), (uint8_t*)ctx->d_work.p + (size_t)f0 * analyze_work_stride(P), st, nullptr);
                launch_pack(J->d_pcm.p, (const FrameDesc*)ctx->d_frames.p + f0, P, cnf, (const SubframePlan*)ctx->d_plans.p + (size_t)f0 * P.n_signals,
                            (const uint8_t*)ctx->d_ca.p + f0, (uint8_t*)ctx->d_scratch.p + (size_t)f0 * ctx->scratch_stride, ctx->scratch_stride,
                            (uint32_t*)S.flen.p + f0, st);
            }
            launch_layout((const uint32_t*)S.flen.p + f0, (const FrameDesc*)ctx->d_frames.p + f0, cnf, J->pro, J->dev_base[c],
                          (uint64_t*)S.foff.p + f0, (uint64_t*)J->d_totals.p + c, st);
            launch_compact((const uint8_t*)ctx->d_scratch.p + (size_t)f0 * ctx->scratch_stride, ctx->scratch_stride, (const uint32_t*)S.flen.p + f0,
                           (const uint64_t*)S.foff.p + f0, (uint8_t*)S.arena.p, cnf, st);
            launch_finalize((const uint32_t*)S.flen.p, (const uint64_t*)S.foff.p, (const uint32_t*)ctx->d_sfirst.p + s0,
                            (const uint32_t*)ctx->d_snframes.p + s0, (const uint64_t*)ctx->d_ssamples.p + s0, nullptr, cns, P, J->pro ? 1u : 0u,
                            (uint8_t*)S.arena.p, (StreamInfoOut*)S.sinfo.p + s0, st);
            ctx->launches += 4 + n_an;
        } else {
            CK(cudaMemsetAsync((uint64_t*)J->d_totals.p + c, 0, 8, st));
        }
        CK(cudaMemcpyAsync(J->h_totals + c, (uint64_t*)J->d_totals.p + c, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(J->ev_done[c], st));
    }
    if (J->dev_base[nchunks] > S.arena.cap) return fail(ctx, FLACB200_ERR_CUDA, "device arena too small");
    if (J->gpu_md5) {
        CK(cudaMemcpyAsync(J->h_digests, S.md5.p, (size_t)ns * 16, cudaMemcpyDeviceToHost, S.side));
        CK(cudaEventRecord(J->ev_md5, S.side));
        CK(cudaEventRecord(S.ev_free, S.side)); S.busy = true;          // the set is free again when its digests have left
    }
    CK(cudaGetLastError());
    J->active = true;
    ctx->next_job = (slot + 1) % flacb200_ctx::kJobs;
    J->drain = std::thread(drain_job, ctx, J);
    return 0;
}

extern "C" int flacb200_encode_host_collect(flacb200_ctx* ctx, int ticket, uint64_t* total_bytes) {
    if (!ctx) return FLACB200_ERR_NO_DEVICE;
    if (ticket < 0 || ticket >= flacb200_ctx::kJobs || !ctx->jobs[ticket] || !ctx->jobs[ticket]->active) return fail(ctx, FLACB200_ERR_ARG, "no such batch in flight");
    flacb200_ctx::HostJob* J = ctx->jobs[ticket];
    if (J->drain.joinable()) J->drain.join();
    J->active = false;
    if (total_bytes) *total_bytes = J->total;
    if (J->rc) return fail(ctx, J->rc, J->err.c_str());
    ctx->e2e_last_bytes = J->total;
    if (J->nf) {
        // libm-log guard (see flacb200_encode_batch_host): settled here; the rare second pass runs as one synchronous call once the
        // other batches in flight have drained (their results stay where they are until they are collected)
        size_t n_new = 0;
        cudaSetDevice(ctx->device);
        int grc = guard_settle(ctx, ctx->sets[J->set_idx], J->h_stats->log_ambiguous, &n_new);
        if (grc) return grc;
        if (n_new) {
            for (auto& o : ctx->jobs) if (o && o->drain.joinable()) o->drain.join();
            ctx->in_rerun = true;
            uint64_t tb = 0;
            const int rrc = flacb200_encode_batch_host(ctx, &ctx->cfg, J->pcm_host, J->pcm_elems, (uint32_t)ctx->n_streams, ctx->h_stream_off.data(),
                                                       ctx->h_stream_samples.data(), J->arena, J->arena_cap, &tb, J->frame_off, J->frame_len, J->streams);
            ctx->in_rerun = false; ctx->h_ovr.clear();
            if (total_bytes) *total_bytes = tb;
            return rrc;
        }
    }
    return 0;
}
