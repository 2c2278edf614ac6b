/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked or loaded by the product path).
 *
 * Thin C driver around the reference's own CPU implementation: the prebuilt libFLAC 1.4.3
 * that pyFLAC binds (/root/reference/pyflac/libraries/linux-x86_64/libFLAC-12.1.0.so,
 * build_args.py:49-51).  oracle/Makefile copies that binary into the git-ignored
 * oracle/_ref/ and builds this file into oracle/_ref/libflacref.so next to it, so both
 * travel to the GPU box.  Used by:
 *   - tests/        : libFLAC bytes / PCM as the ground truth the restatement (flac_oracle.c)
 *                     and the CUDA path are compared against;
 *   - bench.py      : the `cpu_baseline` leg and `--impl reference` (pthreads, one
 *                     FLAC__StreamEncoder / Decoder per thread, C callbacks).
 *
 * The entry points mirror what pyFLAC itself calls (pyflac/encoder.py:77,115,132,153-231,319;
 * pyflac/decoder.py:85,99,170,196).  The prototypes below are re-declared by hand from the
 * public libFLAC C API (pyflac/include/FLAC/stream_encoder.h, stream_decoder.h) -- no
 * reference header is copied into this repository.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>
#include <time.h>

typedef struct FLAC__StreamEncoder FLAC__StreamEncoder;
typedef struct FLAC__StreamDecoder FLAC__StreamDecoder;
typedef int FLAC__bool;

/* stream_encoder.h: callback signatures (:520-620) and the calls pyFLAC makes */
typedef int (*enc_write_cb)(const FLAC__StreamEncoder *, const uint8_t *, size_t, uint32_t, uint32_t, void *);
typedef int (*enc_seek_cb)(const FLAC__StreamEncoder *, uint64_t, void *);
typedef int (*enc_tell_cb)(const FLAC__StreamEncoder *, uint64_t *, void *);
typedef void (*enc_meta_cb)(const FLAC__StreamEncoder *, const void *, void *);

extern FLAC__StreamEncoder *FLAC__stream_encoder_new(void);
extern void FLAC__stream_encoder_delete(FLAC__StreamEncoder *);
extern FLAC__bool FLAC__stream_encoder_set_verify(FLAC__StreamEncoder *, FLAC__bool);
extern FLAC__bool FLAC__stream_encoder_set_streamable_subset(FLAC__StreamEncoder *, FLAC__bool);
extern FLAC__bool FLAC__stream_encoder_set_channels(FLAC__StreamEncoder *, uint32_t);
extern FLAC__bool FLAC__stream_encoder_set_bits_per_sample(FLAC__StreamEncoder *, uint32_t);
extern FLAC__bool FLAC__stream_encoder_set_sample_rate(FLAC__StreamEncoder *, uint32_t);
extern FLAC__bool FLAC__stream_encoder_set_compression_level(FLAC__StreamEncoder *, uint32_t);
extern FLAC__bool FLAC__stream_encoder_disable_instruction_set(FLAC__StreamEncoder *, uint32_t);   /* libFLAC's test hook: bit mask of SIMD levels to turn off */
extern FLAC__bool FLAC__stream_encoder_set_blocksize(FLAC__StreamEncoder *, uint32_t);
extern FLAC__bool FLAC__stream_encoder_set_limit_min_bitrate(FLAC__StreamEncoder *, FLAC__bool);
extern FLAC__bool FLAC__stream_encoder_set_do_md5(FLAC__StreamEncoder *, FLAC__bool);
extern FLAC__bool FLAC__stream_encoder_set_max_lpc_order(FLAC__StreamEncoder *, uint32_t);
extern FLAC__bool FLAC__stream_encoder_set_qlp_coeff_precision(FLAC__StreamEncoder *, uint32_t);
extern FLAC__bool FLAC__stream_encoder_set_do_exhaustive_model_search(FLAC__StreamEncoder *, FLAC__bool);
extern FLAC__bool FLAC__stream_encoder_set_min_residual_partition_order(FLAC__StreamEncoder *, uint32_t);
extern FLAC__bool FLAC__stream_encoder_set_max_residual_partition_order(FLAC__StreamEncoder *, uint32_t);
extern int FLAC__stream_encoder_init_stream(FLAC__StreamEncoder *, enc_write_cb, enc_seek_cb, enc_tell_cb, enc_meta_cb, void *);
extern FLAC__bool FLAC__stream_encoder_process_interleaved(FLAC__StreamEncoder *, const int32_t *, uint32_t);
extern FLAC__bool FLAC__stream_encoder_finish(FLAC__StreamEncoder *);
extern int FLAC__stream_encoder_get_state(const FLAC__StreamEncoder *);

/* stream_decoder.h */
typedef int (*dec_read_cb)(const FLAC__StreamDecoder *, uint8_t *, size_t *, void *);
typedef int (*dec_write_cb)(const FLAC__StreamDecoder *, const void *frame, const int32_t *const buffer[], void *);
typedef void (*dec_error_cb)(const FLAC__StreamDecoder *, int, void *);
extern FLAC__StreamDecoder *FLAC__stream_decoder_new(void);
extern void FLAC__stream_decoder_delete(FLAC__StreamDecoder *);
extern FLAC__bool FLAC__stream_decoder_set_md5_checking(FLAC__StreamDecoder *, FLAC__bool);
extern int FLAC__stream_decoder_init_stream(FLAC__StreamDecoder *, dec_read_cb, void *, void *, void *, void *, dec_write_cb, void *, dec_error_cb, void *);
extern FLAC__bool FLAC__stream_decoder_process_until_end_of_stream(FLAC__StreamDecoder *);
extern FLAC__bool FLAC__stream_decoder_finish(FLAC__StreamDecoder *);
extern uint32_t FLAC__stream_decoder_get_channels(const FLAC__StreamDecoder *);
extern uint32_t FLAC__stream_decoder_get_bits_per_sample(const FLAC__StreamDecoder *);
extern uint32_t FLAC__stream_decoder_get_sample_rate(const FLAC__StreamDecoder *);
extern uint32_t FLAC__stream_decoder_get_blocksize(const FLAC__StreamDecoder *);
extern const char *FLAC__VENDOR_STRING;

/* ------------------------------------------------------------------ encode ---- */

typedef struct {
    uint8_t *buf;
    size_t cap, len, pos;
    /* frame index: one entry per write callback with samples > 0 */
    uint64_t *frame_off;
    uint32_t *frame_len;
    uint32_t *frame_samples;
    uint32_t nframes, frames_cap;
    int overflow;
} memfile;

static int mf_write(const FLAC__StreamEncoder *e, const uint8_t *b, size_t n, uint32_t samples, uint32_t frame, void *cd)
{
    memfile *m = (memfile *)cd;
    (void)e; (void)frame;
    if (m->pos + n > m->cap) { m->overflow = 1; return 1; /* FATAL_ERROR */ }
    memcpy(m->buf + m->pos, b, n);
    if (samples > 0 && m->frame_off && m->nframes < m->frames_cap) {
        m->frame_off[m->nframes] = m->pos;
        m->frame_len[m->nframes] = (uint32_t)n;
        m->frame_samples[m->nframes] = samples;
        m->nframes++;
    }
    m->pos += n;
    if (m->pos > m->len) m->len = m->pos;
    return 0;
}
static int mf_seek(const FLAC__StreamEncoder *e, uint64_t off, void *cd)
{
    memfile *m = (memfile *)cd; (void)e;
    if (off > m->cap) return 1;
    m->pos = (size_t)off;
    return 0;
}
static int mf_tell(const FLAC__StreamEncoder *e, uint64_t *off, void *cd)
{
    memfile *m = (memfile *)cd; (void)e;
    *off = m->pos;
    return 0;
}

typedef struct {
    uint32_t sample_rate, channels, bps, level, blocksize;
    int seekable;         /* 1: seek/tell callbacks given => STREAMINFO rewritten (== FileEncoder bytes) */
    int limit_min_bitrate;
    int streamable_subset;
    int do_md5;           /* libFLAC's undocumented test hook; 1 = default behaviour */
} ref_enc_cfg;

const char *ref_vendor_string(void) { return FLAC__VENDOR_STRING; }

/* Encode one stream (interleaved int32 samples as pyFLAC passes them, encoder.py:112-115).
 * `chunk` > 0 feeds the encoder in pieces of that many samples (exercises the over-read framing).
 * Returns total bytes written or -1. */
long ref_encode_stream(const ref_enc_cfg *cfg, const int32_t *pcm, uint64_t nsamples, uint32_t chunk,
                       uint8_t *out, size_t out_cap,
                       uint64_t *frame_off, uint32_t *frame_len, uint32_t *frame_samples,
                       uint32_t frames_cap, uint32_t *nframes)
{
    memfile m;
    FLAC__StreamEncoder *e = FLAC__stream_encoder_new();
    int ok = 1;
    if (!e) return -1;
    memset(&m, 0, sizeof m);
    m.buf = out; m.cap = out_cap;
    m.frame_off = frame_off; m.frame_len = frame_len; m.frame_samples = frame_samples; m.frames_cap = frames_cap;
    FLAC__stream_encoder_set_channels(e, cfg->channels);
    FLAC__stream_encoder_set_bits_per_sample(e, cfg->bps);
    FLAC__stream_encoder_set_sample_rate(e, cfg->sample_rate);
    FLAC__stream_encoder_set_compression_level(e, cfg->level);
    { const char *ev = getenv("REF_DISABLE_SIMD"); if (ev) FLAC__stream_encoder_disable_instruction_set(e, (uint32_t)strtoul(ev, 0, 0)); }
    /* tuning away from the presets (stream_encoder.h:993-1115): only used to make DECODE fixtures whose subframes the
     * presets never produce (orders above 12, partition orders above 6, exhaustive order search) */
    { const char *ev;
      if ((ev = getenv("REF_MAX_LPC_ORDER"))) FLAC__stream_encoder_set_max_lpc_order(e, (uint32_t)atoi(ev));
      if ((ev = getenv("REF_QLP_PRECISION"))) FLAC__stream_encoder_set_qlp_coeff_precision(e, (uint32_t)atoi(ev));
      if ((ev = getenv("REF_EXHAUSTIVE"))) FLAC__stream_encoder_set_do_exhaustive_model_search(e, atoi(ev));
      if ((ev = getenv("REF_MIN_PART_ORDER"))) FLAC__stream_encoder_set_min_residual_partition_order(e, (uint32_t)atoi(ev));
      if ((ev = getenv("REF_MAX_PART_ORDER"))) FLAC__stream_encoder_set_max_residual_partition_order(e, (uint32_t)atoi(ev)); }
    FLAC__stream_encoder_set_blocksize(e, cfg->blocksize);
    FLAC__stream_encoder_set_streamable_subset(e, cfg->streamable_subset);
    FLAC__stream_encoder_set_limit_min_bitrate(e, cfg->limit_min_bitrate);
    if (!cfg->do_md5) FLAC__stream_encoder_set_do_md5(e, 0);
    if (FLAC__stream_encoder_init_stream(e, mf_write, cfg->seekable ? mf_seek : 0, cfg->seekable ? mf_tell : 0, 0, &m) != 0) {
        FLAC__stream_encoder_delete(e);
        return -2;
    }
    if (chunk == 0) chunk = 0x7fffffffu;
    for (uint64_t done = 0; ok && done < nsamples;) {
        uint64_t n = nsamples - done;
        if (n > chunk) n = chunk;
        ok = FLAC__stream_encoder_process_interleaved(e, pcm + done * cfg->channels, (uint32_t)n);
        done += n;
    }
    if (!FLAC__stream_encoder_finish(e)) ok = 0;
    FLAC__stream_encoder_delete(e);
    if (nframes) *nframes = m.nframes;
    if (!ok || m.overflow) return -3;
    return (long)m.len;
}

/* ------------------------------------------------------------------ decode ---- */

typedef struct {
    const uint8_t *in; size_t in_len, in_pos;
    int32_t *out; uint64_t out_cap /* in interleaved samples (frames) */, out_frames;
    uint32_t channels, bps, sample_rate;
    int errors, overflow;
} decctx;

/* FLAC__Frame header begins with: blocksize, sample_rate, channels, channel_assignment, bits_per_sample
 * (format.h:418-440), all 32-bit; that prefix is all the harness reads. */
typedef struct { uint32_t blocksize, sample_rate, channels, channel_assignment, bits_per_sample; } frame_hdr_prefix;

static int dc_read(const FLAC__StreamDecoder *d, uint8_t *buf, size_t *bytes, void *cd)
{
    decctx *c = (decctx *)cd; (void)d;
    size_t n = c->in_len - c->in_pos;
    if (n == 0) { *bytes = 0; return 1; /* END_OF_STREAM */ }
    if (n > *bytes) n = *bytes;
    memcpy(buf, c->in + c->in_pos, n);
    c->in_pos += n; *bytes = n;
    return 0;
}
static int dc_write(const FLAC__StreamDecoder *d, const void *frame, const int32_t *const buffer[], void *cd)
{
    decctx *c = (decctx *)cd; (void)d;
    const frame_hdr_prefix *h = (const frame_hdr_prefix *)frame;
    c->channels = h->channels; c->bps = h->bits_per_sample; c->sample_rate = h->sample_rate;
    if (c->out) {
        if (c->out_frames + h->blocksize > c->out_cap) { c->overflow = 1; return 1; }
        for (uint32_t ch = 0; ch < h->channels; ch++) {
            const int32_t *src = buffer[ch];
            int32_t *dst = c->out + c->out_frames * h->channels + ch;
            for (uint32_t i = 0; i < h->blocksize; i++) dst[(size_t)i * h->channels] = src[i];
        }
    }
    c->out_frames += h->blocksize;
    return 0;
}
static void dc_error(const FLAC__StreamDecoder *d, int status, void *cd)
{
    decctx *c = (decctx *)cd; (void)d; (void)status;
    c->errors++;
}

/* Decode a whole .flac byte string into interleaved int32. Returns number of inter-channel
 * samples decoded, or negative on error.  info[0..3] = channels, bps, sample_rate, error count. */
long ref_decode_stream(const uint8_t *in, size_t in_len, int32_t *out, uint64_t out_cap, uint32_t *info)
{
    decctx c;
    FLAC__StreamDecoder *d = FLAC__stream_decoder_new();
    if (!d) return -1;
    memset(&c, 0, sizeof c);
    c.in = in; c.in_len = in_len; c.out = out; c.out_cap = out_cap;
    FLAC__stream_decoder_set_md5_checking(d, 1);
    if (FLAC__stream_decoder_init_stream(d, dc_read, 0, 0, 0, 0, dc_write, 0, dc_error, &c) != 0) {
        FLAC__stream_decoder_delete(d);
        return -2;
    }
    int ok = FLAC__stream_decoder_process_until_end_of_stream(d);
    int md5ok = FLAC__stream_decoder_finish(d);
    FLAC__stream_decoder_delete(d);
    if (info) { info[0] = c.channels; info[1] = c.bps; info[2] = c.sample_rate; info[3] = (uint32_t)c.errors + (md5ok ? 0u : 1000u); }
    if (!ok || c.overflow) return -3;
    return (long)c.out_frames;
}

/* --------------------------------------------------- multi-threaded timing ---- */

typedef struct {
    const ref_enc_cfg *cfg;
    const int32_t *pcm;        /* [n_streams][nsamples][channels] int32 */
    const int16_t *pcm16;      /* alternative container: int16, widened per stream inside the timed region like encoder.py:112 */
    uint64_t nsamples;
    uint32_t n_streams, n_threads, tid;
    uint64_t bytes_out;
    uint8_t *out_all; size_t out_stride; uint64_t *out_lens;   /* optional: keep every stream's bytes for comparison */
    int fail;
} enc_job;

static void *enc_worker(void *arg)
{
    enc_job *j = (enc_job *)arg;
    size_t cap = (size_t)j->nsamples * j->cfg->channels * 5 + 65536;
    uint8_t *out = (uint8_t *)malloc(cap);
    int32_t *wide = j->pcm16 ? (int32_t *)malloc((size_t)j->nsamples * j->cfg->channels * 4) : 0;
    for (uint32_t s = j->tid; s < j->n_streams; s += j->n_threads) {
        const int32_t *src;
        size_t n = (size_t)j->nsamples * j->cfg->channels;
        if (j->pcm16) {
            const int16_t *p = j->pcm16 + (size_t)s * n;
            for (size_t i = 0; i < n; i++) wide[i] = p[i];
            src = wide;
        } else src = j->pcm + (size_t)s * n;
        long r;
        if (j->out_all) {
            r = ref_encode_stream(j->cfg, src, j->nsamples, 0, j->out_all + (size_t)s * j->out_stride, j->out_stride, 0, 0, 0, 0, 0);
            if (r >= 0) j->out_lens[s] = (uint64_t)r;
        } else r = ref_encode_stream(j->cfg, src, j->nsamples, 0, out, cap, 0, 0, 0, 0, 0);
        if (r < 0) j->fail = 1; else j->bytes_out += (uint64_t)r;
    }
    free(out); free(wide);
    return 0;
}

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

/* Encode n_streams equal-length streams on n_threads pthreads (streams dealt round-robin).
 * Exactly one of pcm32 / pcm16 is non-NULL. Returns elapsed seconds (<0 on failure). */
double ref_encode_mt(const ref_enc_cfg *cfg, const int32_t *pcm32, const int16_t *pcm16, uint64_t nsamples,
                     uint32_t n_streams, uint32_t n_threads, uint64_t *bytes_out,
                     uint8_t *out_all, size_t out_stride, uint64_t *out_lens)
{
    pthread_t *th = (pthread_t *)calloc(n_threads, sizeof *th);
    enc_job *jobs = (enc_job *)calloc(n_threads, sizeof *jobs);
    double t0 = now_s();
    for (uint32_t t = 0; t < n_threads; t++) {
        jobs[t].cfg = cfg; jobs[t].pcm = pcm32; jobs[t].pcm16 = pcm16; jobs[t].nsamples = nsamples;
        jobs[t].n_streams = n_streams; jobs[t].n_threads = n_threads; jobs[t].tid = t;
        jobs[t].out_all = out_all; jobs[t].out_stride = out_stride; jobs[t].out_lens = out_lens;
        pthread_create(&th[t], 0, enc_worker, &jobs[t]);
    }
    uint64_t total = 0; int fail = 0;
    for (uint32_t t = 0; t < n_threads; t++) { pthread_join(th[t], 0); total += jobs[t].bytes_out; fail |= jobs[t].fail; }
    double dt = now_s() - t0;
    if (bytes_out) *bytes_out = total;
    free(th); free(jobs);
    return fail ? -1.0 : dt;
}

typedef struct {
    const uint8_t *blob; const uint64_t *off; const uint64_t *len;
    uint32_t n_streams, n_threads, tid;
    uint64_t max_frames; uint64_t frames_out; int fail;
} dec_job;

static void *dec_worker(void *arg)
{
    dec_job *j = (dec_job *)arg;
    for (uint32_t s = j->tid; s < j->n_streams; s += j->n_threads) {
        uint32_t info[4];
        long r = ref_decode_stream(j->blob + j->off[s], (size_t)j->len[s], 0, 0, info);
        if (r < 0 || info[3]) j->fail = 1; else j->frames_out += (uint64_t)r;
    }
    return 0;
}

/* Decode n_streams .flac byte strings (blob + offsets) on n_threads; PCM is produced by libFLAC
 * into its own buffers and dropped (the write callback only counts) -- the decode work itself is timed. */
double ref_decode_mt(const uint8_t *blob, const uint64_t *off, const uint64_t *len, uint32_t n_streams,
                     uint32_t n_threads, uint64_t *frames_out)
{
    pthread_t *th = (pthread_t *)calloc(n_threads, sizeof *th);
    dec_job *jobs = (dec_job *)calloc(n_threads, sizeof *jobs);
    double t0 = now_s();
    for (uint32_t t = 0; t < n_threads; t++) {
        jobs[t].blob = blob; jobs[t].off = off; jobs[t].len = len;
        jobs[t].n_streams = n_streams; jobs[t].n_threads = n_threads; jobs[t].tid = t;
        pthread_create(&th[t], 0, dec_worker, &jobs[t]);
    }
    uint64_t total = 0; int fail = 0;
    for (uint32_t t = 0; t < n_threads; t++) { pthread_join(th[t], 0); total += jobs[t].frames_out; fail |= jobs[t].fail; }
    double dt = now_s() - t0;
    if (frames_out) *frames_out = total;
    free(th); free(jobs);
    return fail ? -1.0 : dt;
}
