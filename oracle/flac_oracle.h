/*
 * flac_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("port") of the libFLAC 1.4.3 encode / decode hot path that pyFLAC drives
 * (pyflac/encoder.py:115,132 ; pyflac/decoder.py:196,294,388).  The arithmetic lives in the
 * third-party dependency xiph/flac 1.4.3 (src/libFLAC/), whose source is NOT in /root/reference
 * (only headers + a stripped binary are vendored: pyflac/include/FLAC, pyflac/libraries).  This file
 * therefore restates the published algorithm (RFC 9639 + upstream stream_encoder.c / lpc.c /
 * fixed.c / window.c / stream_encoder_framing.c / bitwriter.c semantics, summarised and verified
 * in SURVEY.md Appendix A) and is PINNED byte-for-byte against the reference binary itself
 * (oracle/_ref, tests/test_oracle_vs_ref.py) and against committed golden vectors generated from
 * that binary (tests/golden/, tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 */
#ifndef FLAC_ORACLE_H
#define FLAC_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FO_MAX_CHANNELS 8
#define FO_MAX_LPC_ORDER 32
#define FO_MAX_PART_ORDER 8
#define FO_MAX_APOD_STEPS 16

typedef struct {
    uint32_t sample_rate, channels, bps, level, blocksize; /* blocksize 0 => libFLAC default */
    int32_t seekable;          /* 1 => STREAMINFO finalised (md5, total samples, frame sizes) like FileEncoder */
    int32_t limit_min_bitrate;
    int32_t streamable_subset;
} fo_enc_cfg;

/* Decisions for one coded subframe (what ends up in the bitstream). */
typedef struct {
    int32_t type;              /* 0 CONSTANT, 1 VERBATIM, 2 FIXED, 3 LPC */
    int32_t order;
    int32_t wasted;
    int32_t sbps;              /* subframe bits per sample (after wasted bits, +1 for side) */
    int32_t precision;         /* qlp coefficient precision (LPC) */
    int32_t shift;             /* quantization level (LPC) */
    int32_t qlp[FO_MAX_LPC_ORDER];
    int32_t partition_order;
    int32_t rice2;             /* 1 => 5-bit parameters */
    uint32_t rice[1u << FO_MAX_PART_ORDER];
    uint32_t bits_est;         /* libFLAC's estimate used for the decisions (not the real size) */
} fo_subframe;

/* Per-signal analysis trace (debug aid for the CUDA path): every intermediate that decides something. */
typedef struct {
    int32_t wasted, sbps;
    uint64_t fixed_err[5];
    int32_t fixed_order;
    uint32_t fixed_bits;                       /* candidate estimate, 0 if not evaluated */
    int32_t is_constant;
    int32_t n_apod;                            /* LPC candidates tried */
    double autoc[FO_MAX_APOD_STEPS][FO_MAX_LPC_ORDER + 1];
    double lpc_err[FO_MAX_APOD_STEPS][FO_MAX_LPC_ORDER];
    int32_t lpc_order[FO_MAX_APOD_STEPS];      /* guessed order, 0 = step skipped */
    uint32_t lpc_bits[FO_MAX_APOD_STEPS];      /* candidate estimate, 0 = rejected */
    fo_subframe best;
} fo_signal_trace;

typedef struct {
    uint32_t blocksize, frame_number;
    int32_t channel_assignment;                /* 0 independent, 1 left/side, 2 right/side, 3 mid/side */
    int32_t n_signals;                         /* channels, +2 (mid, side) if mid/side analysed */
    fo_signal_trace sig[FO_MAX_CHANNELS + 2];  /* [0..ch-1] channels, [ch] mid, [ch+1] side */
} fo_frame_trace;

/* Encode a complete stream. pcm = interleaved int32 [nsamples][channels] (what pyFLAC passes to
 * FLAC__stream_encoder_process_interleaved).  Returns bytes written (>0) or a negative error:
 * -1 bad config, -2 output too small, -3 configuration outside the restated range.
 * frame_off/len (optional, frames_cap entries) receive the byte span of every audio frame.
 * traces (optional, traces_cap entries) receive per-frame analysis traces. */
long fo_encode_stream(const fo_enc_cfg *cfg, const int32_t *pcm, uint64_t nsamples,
                      uint8_t *out, size_t out_cap,
                      uint64_t *frame_off, uint32_t *frame_len, uint32_t frames_cap, uint32_t *nframes,
                      fo_frame_trace *traces, uint32_t traces_cap);

/* Validate settings exactly like FLAC__stream_encoder_init_stream does; returns the
 * FLAC__StreamEncoderInitStatus value (0 == OK). */
int fo_encoder_init_status(const fo_enc_cfg *cfg, int has_write_cb, int has_seek_cb, int has_tell_cb);

/* Decode a complete .flac byte string to interleaved int32. Returns inter-channel sample count or
 * negative error (-1 not FLAC, -2 out too small, -3 bitstream error, -4 CRC mismatch, -5 MD5 mismatch).
 * info[0..3] = channels, bps, sample_rate, blocksize of first frame. out may be NULL (count only). */
long fo_decode_stream(const uint8_t *in, size_t in_len, int32_t *out, uint64_t out_cap, uint32_t *info);

/* Helpers exposed for unit tests */
void fo_window_tukey(float *w, int32_t L, float p);
void fo_md5(const uint8_t *data, size_t len, uint8_t digest[16]);
uint8_t fo_crc8(const uint8_t *data, size_t len);
uint16_t fo_crc16(const uint8_t *data, size_t len);

#ifdef __cplusplus
}
#endif
#endif
