/*
 * flacb200.h -- C ABI of libflacb200.so, the B200-native FLAC encode/decode engine.
 *
 * Two layers, both plain C (pointers + sizes, no torch / C++ types):
 *
 *  (1) BATCH ABI (additive; the only way to keep a GPU busy): many independent streams per call,
 *      PCM either resident in HBM (e.g. a torch tensor's data_ptr) or in host memory.
 *      This is what bench.py times and what the per-handle layer is built on.
 *
 *  (2) DROP-IN ABI: the subset of libFLAC's stream encoder/decoder C API that pyFLAC's cffi
 *      modules bind, with identical names, signatures, status values and callback contracts:
 *        encoder  -- /root/reference/pyflac/builder/encoder.py:251-256 (callbacks), :266-322 (functions),
 *                    called from pyflac/encoder.py:77,115,132,141,145-231,319,401
 *        decoder  -- /root/reference/pyflac/builder/decoder.py:368-375 (callbacks), :387-475 (functions),
 *                    called from pyflac/decoder.py:85,99,108,170,196,271,294,372,388
 *      Repointing pyFLAC's build_args.py:49-51 at this library is the whole integration
 *      (INTEGRATION.md).  Declared in flacb200_flac_api.h.
 *
 * Every compute entry point fails loudly (non-zero status + flacb200_last_error) when no CUDA
 * device is usable; there is no CPU fallback.
 */
#ifndef FLACB200_H
#define FLACB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct flacb200_ctx flacb200_ctx;

enum {
    FLACB200_OK = 0,
    FLACB200_ERR_NO_DEVICE = 1,      /* CUDA runtime reports no usable device */
    FLACB200_ERR_CUDA = 2,           /* a CUDA call failed; see flacb200_last_error */
    FLACB200_ERR_CONFIG = 3,         /* settings libFLAC itself would reject (status in flacb200_last_error) */
    FLACB200_ERR_UNSUPPORTED = 4,    /* valid FLAC settings outside this build's range (see DESIGN.md "limits") */
    FLACB200_ERR_ARG = 5,
    FLACB200_ERR_BITSTREAM = 6       /* decode: malformed input */
};

/* Encoder settings == what pyFLAC can set (encoder.py:293-316): everything else follows
 * libFLAC's compression-level table (stream_encoder.h:845-853). */
typedef struct {
    uint32_t sample_rate;
    uint32_t channels;
    uint32_t bits_per_sample;     /* 4..32 */
    uint32_t compression_level;   /* 0..8 */
    uint32_t blocksize;           /* 0 = libFLAC default (1152 for levels 0-2, else 4096) */
    uint32_t container_bytes;     /* PCM element size in memory: 2 (int16, bps<=16) or 4 (int32) */
    uint32_t write_prologue;      /* 1: every stream's bytes in the arena start with fLaC+STREAMINFO+VORBIS_COMMENT (== FileEncoder output) */
    uint32_t do_md5;              /* 1: STREAMINFO carries the MD5 of the PCM (libFLAC default) */
    uint32_t streamable_subset;   /* validation only (stream_encoder.h:1005-1017) */
    uint32_t debug_trace;         /* 1: keep per-signal analysis traces (tests) */
    uint32_t limit_min_bitrate;   /* FLAC__stream_encoder_set_limit_min_bitrate (stream_encoder.h:1105-1115) */
    /* Fine-grained settings (builder/encoder.py:274-284; libFLAC's set_do_mid_side_stereo ... set_apodization).  tune != 0: the seven
     * fields below replace what the compression level stands for (all of them: fill in the level's own values where nothing changes);
     * tune == 0: they are ignored.  Supported range of this build: max_lpc_order <= 12, max_residual_partition_order <= 6, one
     * apodization of the tukey family -- apod_parts 1 = tukey(apod_p), 2 or 3 = subdivide_tukey(parts/apod_p) -- anything else is
     * FLACB200_ERR_UNSUPPORTED. */
    uint32_t tune;
    uint32_t do_mid_side, loose_mid_side;
    uint32_t max_lpc_order;       /* 0 = fixed predictors only */
    uint32_t qlp_coeff_precision; /* 0 = libFLAC's automatic choice, else 5..15 */
    uint32_t max_residual_partition_order;
    uint32_t apod_parts;
    float    apod_p;
} flacb200_enc_config;

typedef struct {
    uint64_t total_samples;
    uint64_t byte_off;            /* offset of the stream's first byte in the arena */
    uint64_t byte_len;
    uint32_t min_framesize, max_framesize;
    uint32_t n_frames, pad;
    uint8_t  md5[16];
} flacb200_stream_info;

typedef struct {
    uint64_t total_bytes;         /* bytes used in the arena */
    uint32_t n_frames;
    uint32_t n_streams;
    uint64_t log_guard_hits;      /* decisions inside the libm-log guard band (0 in every test) */
    /* device pointers (owned by the ctx, valid until the next call on it) */
    const uint8_t  *d_arena;
    const uint64_t *d_frame_off;
    const uint32_t *d_frame_len;
} flacb200_enc_result;

int  flacb200_create(flacb200_ctx **out, int device);
void flacb200_destroy(flacb200_ctx *ctx);
const char *flacb200_last_error(const flacb200_ctx *ctx);
/* The caller's CUDA stream (cudaStream_t as void*, e.g. torch's current stream; NULL = a ctx-owned one).  Decode work runs on it.
 * Encode work runs on the engine's own streams, ordered behind this stream as it stands when flacb200_encode_batch is called (so
 * PCM produced by earlier work on it is complete) -- and behind nothing else: the MD5 chain of a batch does not queue behind the
 * encode kernels of earlier batches. */
int  flacb200_set_stream(flacb200_ctx *ctx, void *cuda_stream);
int  flacb200_sync(flacb200_ctx *ctx);
/* Batches overlap: encode kernels on the engine's encode stream, the MD5 + STREAMINFO patch of each batch on a side stream (five
 * rotating output sets).  flacb200_join makes the caller's stream wait for all of it, so that an event recorded there afterwards
 * covers every batch issued so far; result / fetch / sync wait on the host.  The PCM of a batch must stay unchanged until then. */
int  flacb200_join(flacb200_ctx *ctx);

/* libFLAC's init-time validation for these settings: returns the FLAC__StreamEncoderInitStatus
 * value (0 = OK), pyflac/builder/encoder.py:65-80. */
int  flacb200_enc_validate(const flacb200_enc_config *cfg);

/* Encode n_streams independent streams.  Stream s = stream_samples[s] inter-channel samples starting
 * at ELEMENT offset stream_off[s] of `pcm` (interleaved [sample][channel], container_bytes each).
 * first_frame_number may be NULL (all zero).  Asynchronous on the ctx stream when pcm_is_device;
 * results are fetched with flacb200_encode_result / flacb200_encode_fetch. */
int  flacb200_encode_batch(flacb200_ctx *ctx, const flacb200_enc_config *cfg,
                           const void *pcm, int pcm_is_device, uint64_t pcm_elems,
                           uint32_t n_streams, const uint64_t *stream_off, const uint64_t *stream_samples,
                           const uint32_t *first_frame_number);
/* Loose mid/side (compression levels 1 and 4 on stereo, stream_encoder.h:881-909) decides between independent and
 * mid/side coding once every round(0.4 s) of frames and the frames in between follow that decision.  A batch whose
 * streams continue earlier batches (first_frame_number not a multiple of that period) needs each stream's channel
 * assignment of the frame before the batch: prev_assignment[s] in 0..3 (format.h:388-393 order).  Applies to the
 * next flacb200_encode_batch call only; NULL / not called = every stream starts at a decision frame or with 0. */
int  flacb200_encode_set_prev_assignment(flacb200_ctx *ctx, const uint8_t *prev_assignment, uint32_t n_streams);
/* Channel assignment chosen for every frame of the last batch (n_frames bytes). */
int  flacb200_encode_fetch_assignments(flacb200_ctx *ctx, uint8_t *frame_ca, size_t cap);
/* Synchronises, then reports sizes and device pointers. */
int  flacb200_encode_result(flacb200_ctx *ctx, flacb200_enc_result *res);
/* The MD5 of a stream is one serial chain over all of its samples and ends long after the frames of a batch are final
 * (about 20 ms for 10 s of stereo audio, whatever the batch size).  flacb200_encode_result_frames returns as soon as the
 * frames, the index and the stream prologues are final; the 16 MD5 bytes of every STREAMINFO (arena offset byte_off + 26)
 * are still zero then -- exactly what libFLAC writes before its seek-back at finish() (stream_encoder.h:1744-1770).
 * flacb200_encode_fetch_md5 waits for the chain, returns the digests (16 bytes per stream) and by then the arena and the
 * per-stream info hold them too. */
int  flacb200_encode_result_frames(flacb200_ctx *ctx, flacb200_enc_result *res);
int  flacb200_encode_fetch_md5(flacb200_ctx *ctx, uint8_t *digests, size_t cap);
/* The digests of an earlier batch of the same layout (back = 1: the batch before the last, up to 4): a pipeline of rounds
 * collects them one round late, when that chain has long finished, instead of waiting for the last batch's. */
int  flacb200_encode_fetch_md5_back(flacb200_ctx *ctx, int back, uint8_t *digests, size_t cap);
/* Copy results to host memory (any pointer may be NULL). arena_cap in bytes. */
int  flacb200_encode_fetch(flacb200_ctx *ctx, uint8_t *arena, size_t arena_cap,
                           uint64_t *frame_off, uint32_t *frame_len, uint32_t *frame_samples,
                           uint32_t *frame_stream, flacb200_stream_info *streams);
/* Debug: copy the analysis traces of the last batch (debug_trace=1). Layout = fb::SignalDebug / SubframePlan. */
int  flacb200_encode_fetch_trace(flacb200_ctx *ctx, void *plans, size_t plans_bytes, uint8_t *frame_ca,
                                 void *debug, size_t debug_bytes);
/* One-call host->host convenience used for end-to-end timing: H2D, encode, D2H of the arena + index. */
int  flacb200_encode_batch_host(flacb200_ctx *ctx, const flacb200_enc_config *cfg,
                                const void *pcm_host, uint64_t pcm_elems,
                                uint32_t n_streams, const uint64_t *stream_off, const uint64_t *stream_samples,
                                uint8_t *arena, size_t arena_cap, uint64_t *total_bytes,
                                uint64_t *frame_off, uint32_t *frame_len, flacb200_stream_info *streams);
/* The same work without waiting: submit returns once the batch is enqueued (H2D chunks, kernels, MD5 on the GPU's side stream)
 * and hands out a ticket; up to 3 batches may be in flight, so the PCIe link stays busy in both directions and no call waits
 * for a serial MD5 chain.  collect blocks until that batch's images, index and STREAMINFO digests are in the host buffers given
 * to submit (they and pcm_host must stay valid until then) and returns the byte count.  Tickets are collected in submission
 * order; batches in flight share one stream layout (a different layout needs the earlier ones collected first). */
int  flacb200_encode_host_submit(flacb200_ctx *ctx, const flacb200_enc_config *cfg,
                                 const void *pcm_host, uint64_t pcm_elems,
                                 uint32_t n_streams, const uint64_t *stream_off, const uint64_t *stream_samples,
                                 uint8_t *arena, size_t arena_cap, uint64_t *frame_off, uint32_t *frame_len,
                                 flacb200_stream_info *streams, int *ticket);
int  flacb200_encode_host_collect(flacb200_ctx *ctx, int ticket, uint64_t *total_bytes);
/* ------------------------------------------------------------------ batch decode ----
 * n_streams independent FLAC byte strings (stream s = stream_len[s] bytes at byte offset stream_off[s] of `blob`).
 * Output: interleaved PCM [sample][channel] per stream, back to back in stream order, in `out_container_bytes`
 * sized elements (2 = int16 for <=16-bit streams, 4 = int32; 0 = choose 2 when every stream is <=16 bit).
 * raw != NULL switches to headerless input: every "stream" is a run of frames without fLaC/metadata and raw
 * supplies what STREAMINFO would (used by the drop-in stream decoder to decode as bytes arrive). */
typedef struct {
    uint64_t total_samples;       /* inter-channel samples decoded */
    uint64_t pcm_off;             /* ELEMENT offset of the stream's first sample in the output */
    uint64_t consumed;            /* bytes covered by metadata + successfully chained frames */
    uint32_t n_frames;
    int32_t  status;              /* 0 ok, else the FIRST problem met: 2 not FLAC, 3 bad metadata, 4 bad frame, 5 incomplete frame, 6 lost sync,
                                     7 CRC-16 mismatch, 8 unsupported, 9 reserved values.  Decoding goes on behind a damaged frame the way
                                     libFLAC 1.4.3's does (stream_decoder.h:1440-1460): the frame is dropped, the search resumes, and when a
                                     later frame's number shows that frames are missing they stand as silence in the PCM. */
    uint32_t sample_rate, channels, bits_per_sample, max_blocksize;
    uint32_t n_events;            /* error-callback events libFLAC would have fired; the first 16 are logged below */
    uint32_t gap_samples;         /* samples of silence inserted for missing frames */
    uint32_t ev_frame[16];        /* good frames delivered before event k */
    uint8_t  ev_status[16];       /* FLAC__StreamDecoderErrorStatus of event k (0 LOST_SYNC, 1 BAD_HEADER, 2 FRAME_CRC_MISMATCH, 3 UNPARSEABLE_STREAM) */
    uint64_t next_sample;         /* stream position behind the last delivered frame, by the frame headers: hand it to the next headerless batch */
    uint32_t last_blocksize, have_last;
} flacb200_dec_stream_info;

/* flags bit 0: more input may follow (streaming): errors behind the last good frame are not reported yet;
 * bit 1: frames of this stream were delivered by an earlier batch: next_sample / last_blocksize (from that batch's
 * flacb200_dec_stream_info) let the decoder see frames missing across the batch boundary.  fixed_blocksize: STREAMINFO's
 * blocksize when min == max, else 0. */
typedef struct { uint32_t sample_rate, channels, bits_per_sample, flags; uint64_t next_sample; uint32_t last_blocksize, fixed_blocksize; } flacb200_dec_raw_params;

typedef struct {
    uint64_t total_elems;         /* PCM elements (samples x channels) produced over all streams */
    uint32_t n_streams, n_frames;
    uint32_t out_container_bytes;
    uint32_t n_candidates;        /* frame-start candidates examined (>= n_frames) */
    const void *d_pcm;            /* device pointer, valid until the next call on the ctx */
} flacb200_dec_result;

int  flacb200_decode_batch(flacb200_ctx *ctx, const uint8_t *blob, int blob_is_device, uint64_t blob_bytes,
                           uint32_t n_streams, const uint64_t *stream_off, const uint64_t *stream_len,
                           uint32_t out_container_bytes, const flacb200_dec_raw_params *raw);
int  flacb200_decode_result(flacb200_ctx *ctx, flacb200_dec_result *res);
/* Copy PCM and per-stream info to host (either may be NULL). pcm_cap in bytes.  frame_samples (optional,
 * cap entries) receives the blocksize of every decoded frame in stream order. */
int  flacb200_decode_fetch(flacb200_ctx *ctx, void *pcm, size_t pcm_cap, flacb200_dec_stream_info *streams,
                           uint32_t *frame_samples, uint32_t frame_cap);
/* Sample offset (within its stream's PCM, inserted silence included) of every delivered frame, in the order of
 * flacb200_decode_fetch's frame_samples. */
int  flacb200_decode_fetch_frame_offsets(flacb200_ctx *ctx, uint64_t *frame_sample_off, uint32_t frame_cap);
/* One-call host -> host decode used for end-to-end work: chunks of streams are pipelined (H2D of the next chunk,
 * kernels of the current one, D2H of the previous one run concurrently).  out_container_bytes must be 2 or 4.
 * PCM lands in `pcm` in stream order; streams[s].pcm_off (elements) indexes it; *total_elems = elements written. */
int  flacb200_decode_batch_host(flacb200_ctx *ctx, const uint8_t *blob, uint64_t blob_bytes, uint32_t n_streams,
                                const uint64_t *stream_off, const uint64_t *stream_len, uint32_t out_container_bytes,
                                const flacb200_dec_raw_params *raw, void *pcm, size_t pcm_cap, uint64_t *total_elems,
                                flacb200_dec_stream_info *streams);
/* ms[0..5] = metadata+sync scan, candidate decode + CRC-16, chain+layout, post (undo stereo / interleave), the CRC-16 kernel's share of ms[1], 0 */
int  flacb200_decode_kernel_times(flacb200_ctx *ctx, float *ms);

/* Per-kernel device times of the last batch, measured with CUDA events on the launching streams:
 * ms[0..9] = analysis (all three kernels), pack, scan, compact, finalize(+MD5 join), md5 (side stream), then the
 * analysis split into its kernels: frame_bits (OR/AND), autoc (autocorrelation), analyze (decisions); ms[9] = 1 when
 * the batch ran the fused path (16-bit stereo, csrc/enc_fused.cu): ms[7] is its autocorrelation kernel, ms[8] its
 * persistent worker kernel (TMA-staged frame tile -> frame bytes), ms[1] and ms[6] are zero. */
int  flacb200_set_profiling(flacb200_ctx *ctx, int on);
int  flacb200_kernel_times(flacb200_ctx *ctx, float *ms);
/* Wall-clock breakdown (ms since entry) of the last flacb200_encode_batch_host call:
 * ms[1] work enqueued, ms[2] all chunks' kernels finished, ms[3] D2H finished, ms[4] host MD5 joined, ms[5] return. */
int  flacb200_host_path_times(flacb200_ctx *ctx, double *ms);
/* The same plus how the MD5 work of that call was placed: v[6] ms when the digests hashed on the GPU had reached the host
 * (0: none were), v[7] streams hashed on the GPU, v[8] host MD5 threads, v[9] chunks; n = entries of v to fill (<= 10).
 * Host threads hash the caller's buffer while the GPU encodes; their number is this rank's share of the CPUs the process may
 * run on (sched_getaffinity / LOCAL_WORLD_SIZE).  When they cannot finish by the time the transfer does, the streams of the
 * first chunks are hashed by md5_kernel as their bytes land in HBM, and the split follows the measured finish times. */
int  flacb200_host_path_info(flacb200_ctx *ctx, double *v, int n);
/* Tuning knobs of the two host -> host calls, read from the environment at call time (defaults are measured on a
 * B200 / PCIe 5 host and normally right):
 *   FLACB200_CHUNKS       pieces the PCM of flacb200_encode_batch_host is cut into for the H2D / kernel / D2H pipeline (default 12)
 *   FLACB200_MD5_THREADS  host threads hashing the caller's PCM meanwhile (default: calibrated so they finish with the H2D copy,
 *                         at most this rank's share of the host: affinity mask / FLACB200_LOCAL_RANKS or LOCAL_WORLD_SIZE)
 *   FLACB200_MD5_WARPS    chain warps per md5_kernel CTA, 1..4 (default 4: one per SM sub-partition, the fewest SMs shared with the encode kernels)
 *   FLACB200_MD5_GPU_CHUNKS  leading chunks whose streams md5_kernel hashes instead of the host (default: balanced automatically)
 *   FLACB200_DEC_CHUNKS   groups of streams flacb200_decode_batch_host pipelines (default: one per 200 MB of FLAC, at most 12) */
/* Drop-in layer (flacb200_flac_api.h): concurrent FLAC__stream_encoder_process_interleaved() calls of different handles are
 * coalesced into shared GPU batches by a per-device dispatcher; this reports how many batches / jobs it has run.  Handles
 * choose their device from FLACB200_DEVICE=<index> or round-robin over FLACB200_DEVICES=<i,j,...> (default 0). */
int  flacb200_dispatch_stats(int device, uint64_t *batches, uint64_t *jobs);
/* libm-log guard.  libFLAC's LPC order guess and its "don't even try" test compare costs computed with libm log(); CUDA's log is within
 * an ulp of glibc's but not bit-identical.  Every such decision whose runner-up lies within 1e-12 relative of the winner (thousands of ulps)
 * is logged by the kernels and repeated on the host with the libm the reference links against; when the host decides otherwise the
 * batch is encoded once more with the host's decisions before any result is handed out.  v[0] decisions inside the band in the last
 * batch, v[1] confirmed by the host, v[2] overridden (second pass ran), v[3] not checked (more than 1024 in one batch; also reported as
 * flacb200_enc_result.log_guard_hits).  flacb200_set_log_guard is a test hook: band width, and flip != 0 makes the kernels take the
 * runner-up inside the band, i.e. decide wrongly on purpose. */
int  flacb200_log_guard_info(flacb200_ctx *ctx, uint64_t *v);
int  flacb200_set_log_guard(flacb200_ctx *ctx, double rel, int flip);
/* Kernel launches issued by this ctx so far (bench.py's gpu_launches). */
uint64_t flacb200_launch_count(const flacb200_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
