/*
 * flacb200_flac_api.h -- the DROP-IN layer of libflacb200.so: the libFLAC stream encoder / decoder entry
 * points that pyFLAC's cffi modules bind, re-implemented on top of the batch engine (flacb200.h).
 *
 * Every prototype below replaces the one pyFLAC declares in
 *     /root/reference/pyflac/builder/encoder.py:266-322   (encoder functions, callbacks :251-256)
 *     /root/reference/pyflac/builder/decoder.py:387-475   (decoder functions, callbacks :368-375)
 * and that pyflac/encoder.py / pyflac/decoder.py call (file:line next to each group).  Names, argument
 * meaning, status values (builder/encoder.py:51-106, builder/decoder.py:49-136), callback contracts and
 * the exported string tables are libFLAC 1.4.3's (pyflac/include/FLAC/stream_encoder.h, stream_decoder.h).
 * Types are spelled with plain C types so that this header stands alone.
 *
 * Behavioural contract kept from libFLAC (verified against the reference binary):
 *   - init_stream fires the write callback 3 times ("fLaC", STREAMINFO, VORBIS_COMMENT) with samples=0;
 *   - a frame is emitted only once blocksize+1 samples are buffered; one write callback == one frame;
 *   - finish() codes the remainder as a short frame, rewrites STREAMINFO through seek/tell when given
 *     (MD5 @26, total samples @21, frame sizes @12), fires the metadata callback and resets the handle;
 *   - setters fail once initialised; callbacks run synchronously on the calling thread.
 * What differs: the arithmetic runs on the GPU (no CPU fallback -- init returns
 * FLAC__STREAM_ENCODER_INIT_STATUS_ENCODER_ERROR when no CUDA device is usable), and the fine-grained
 * tuning setters pyFLAC never calls are accepted only at their compression-level values.
 */
#ifndef FLACB200_FLAC_API_H
#define FLACB200_FLAC_API_H
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int FLAC__bool;
typedef uint8_t FLAC__byte;
typedef int32_t FLAC__int32;
typedef uint64_t FLAC__uint64;

/* opaque handles (layout-compatible with builder/encoder.py:115-118, builder/decoder.py:140-145) */
typedef struct { void *protected_; void *private_; } FLAC__StreamEncoder;
typedef struct { void *protected_; void *private_; } FLAC__StreamDecoder;

/* builder/encoder.py:129-137 + :234-248 -- the encoder only ever delivers STREAMINFO (metadata callback at finish) */
typedef struct {
    uint32_t min_blocksize, max_blocksize;
    uint32_t min_framesize, max_framesize;
    uint32_t sample_rate;
    uint32_t channels;
    uint32_t bits_per_sample;
    FLAC__uint64 total_samples;
    FLAC__byte md5sum[16];
} FLAC__StreamMetadata_StreamInfo;
/* builder/decoder.py:256-349 -- the other block types, as the decoder's metadata callback hands them over once
 * FLAC__stream_decoder_set_metadata_respond*() asked for them (the encoder only ever delivers STREAMINFO).  Pointers inside a
 * block are valid for the duration of the callback. */
typedef struct { int dummy; } FLAC__StreamMetadata_Padding;
typedef struct { FLAC__byte id[4]; FLAC__byte *data; } FLAC__StreamMetadata_Application;
typedef struct { FLAC__uint64 sample_number, stream_offset; uint32_t frame_samples; } FLAC__StreamMetadata_SeekPoint;
typedef struct { uint32_t num_points; FLAC__StreamMetadata_SeekPoint *points; } FLAC__StreamMetadata_SeekTable;
typedef struct { uint32_t length; FLAC__byte *entry; } FLAC__StreamMetadata_VorbisComment_Entry;
typedef struct {
    FLAC__StreamMetadata_VorbisComment_Entry vendor_string;
    uint32_t num_comments;
    FLAC__StreamMetadata_VorbisComment_Entry *comments;
} FLAC__StreamMetadata_VorbisComment;
typedef struct { FLAC__uint64 offset; FLAC__byte number; } FLAC__StreamMetadata_CueSheet_Index;
typedef struct {
    FLAC__uint64 offset;
    FLAC__byte number;
    char isrc[13];
    uint32_t type : 1;
    uint32_t pre_emphasis : 1;
    FLAC__byte num_indices;
    FLAC__StreamMetadata_CueSheet_Index *indices;
} FLAC__StreamMetadata_CueSheet_Track;
typedef struct {
    char media_catalog_number[129];
    FLAC__uint64 lead_in;
    FLAC__bool is_cd;
    uint32_t num_tracks;
    FLAC__StreamMetadata_CueSheet_Track *tracks;
} FLAC__StreamMetadata_CueSheet;
typedef struct {
    int type;                 /* FLAC__StreamMetadata_Picture_Type */
    char *mime_type;
    FLAC__byte *description;
    uint32_t width, height, depth, colors;
    uint32_t data_length;
    FLAC__byte *data;
} FLAC__StreamMetadata_Picture;
typedef struct { FLAC__byte *data; } FLAC__StreamMetadata_Unknown;
typedef struct {
    int type;                 /* FLAC__MetadataType: 0 STREAMINFO, 1 PADDING, 2 APPLICATION, 3 SEEKTABLE, 4 VORBIS_COMMENT, 5 CUESHEET, 6 PICTURE */
    FLAC__bool is_last;
    uint32_t length;
    union {
        FLAC__StreamMetadata_StreamInfo stream_info;
        FLAC__StreamMetadata_Padding padding;
        FLAC__StreamMetadata_Application application;
        FLAC__StreamMetadata_SeekTable seek_table;
        FLAC__StreamMetadata_VorbisComment vorbis_comment;
        FLAC__StreamMetadata_CueSheet cue_sheet;
        FLAC__StreamMetadata_Picture picture;
        FLAC__StreamMetadata_Unknown unknown;
        uint64_t pad_[24];
    } data;
} FLAC__StreamMetadata;

/* builder/decoder.py:146-231 -- decode write-callback payload.  pyFLAC reads header.{blocksize,sample_rate,
 * channels,bits_per_sample} (decoder.py:500-524); the subframe array is filled with type/wasted_bits only. */
typedef struct {
    uint32_t blocksize, sample_rate, channels;
    int channel_assignment;
    uint32_t bits_per_sample;
    int number_type;
    union { uint32_t frame_number; FLAC__uint64 sample_number; } number;
    uint8_t crc;
} FLAC__FrameHeader;
typedef struct { int type; uint64_t data_[54]; uint32_t wasted_bits; } FLAC__Subframe;   /* sizeof == 448, as libFLAC 1.4.3 on LP64 */
typedef struct { uint16_t crc; } FLAC__FrameFooter;
typedef struct { FLAC__FrameHeader header; FLAC__Subframe subframes[8]; FLAC__FrameFooter footer; } FLAC__Frame;

/* ---- encoder callbacks: builder/encoder.py:251-256 ---- */
typedef int (*FLAC__StreamEncoderReadCallback)(const FLAC__StreamEncoder *, FLAC__byte buffer[], size_t *bytes, void *client_data);
typedef int (*FLAC__StreamEncoderWriteCallback)(const FLAC__StreamEncoder *, const FLAC__byte buffer[], size_t bytes, uint32_t samples, uint32_t current_frame, void *client_data);
typedef int (*FLAC__StreamEncoderSeekCallback)(const FLAC__StreamEncoder *, FLAC__uint64 absolute_byte_offset, void *client_data);
typedef int (*FLAC__StreamEncoderTellCallback)(const FLAC__StreamEncoder *, FLAC__uint64 *absolute_byte_offset, void *client_data);
typedef void (*FLAC__StreamEncoderMetadataCallback)(const FLAC__StreamEncoder *, const FLAC__StreamMetadata *metadata, void *client_data);
typedef void (*FLAC__StreamEncoderProgressCallback)(const FLAC__StreamEncoder *, FLAC__uint64 bytes_written, FLAC__uint64 samples_written, uint32_t frames_written, uint32_t total_frames_estimate, void *client_data);

extern const char *const FLAC__StreamEncoderStateString[];        /* pyflac/encoder.py:42 */
extern const char *const FLAC__StreamEncoderInitStatusString[];   /* pyflac/encoder.py:54 */
extern const char *FLAC__VENDOR_STRING;

/* pyflac/encoder.py:77 */
FLAC__StreamEncoder *FLAC__stream_encoder_new(void);
void FLAC__stream_encoder_delete(FLAC__StreamEncoder *encoder);
/* pyflac/encoder.py:145-231 (setters return false once initialised, stream_encoder.h:218-223) */
FLAC__bool FLAC__stream_encoder_set_verify(FLAC__StreamEncoder *encoder, FLAC__bool value);
FLAC__bool FLAC__stream_encoder_set_channels(FLAC__StreamEncoder *encoder, uint32_t value);
FLAC__bool FLAC__stream_encoder_set_bits_per_sample(FLAC__StreamEncoder *encoder, uint32_t value);
FLAC__bool FLAC__stream_encoder_set_sample_rate(FLAC__StreamEncoder *encoder, uint32_t value);
FLAC__bool FLAC__stream_encoder_set_compression_level(FLAC__StreamEncoder *encoder, uint32_t value);
FLAC__bool FLAC__stream_encoder_set_blocksize(FLAC__StreamEncoder *encoder, uint32_t value);
FLAC__bool FLAC__stream_encoder_set_streamable_subset(FLAC__StreamEncoder *encoder, FLAC__bool value);
FLAC__bool FLAC__stream_encoder_set_limit_min_bitrate(FLAC__StreamEncoder *encoder, FLAC__bool value);
FLAC__bool FLAC__stream_encoder_set_total_samples_estimate(FLAC__StreamEncoder *encoder, FLAC__uint64 value);
/* in pyFLAC's cdef (builder/encoder.py:274-284), never called by pyFLAC.  Applied like libFLAC's (after the compression level's presets);
 * bit-exact within this build's range: max_lpc_order <= 12, max_residual_partition_order <= 6, qlp_coeff_precision 0 / 5..15, mid/side
 * and loose mid/side, apodization "tukey(P)" or "subdivide_tukey(N[/P])" with N <= 3.  Outside it (exhaustive model / precision
 * search, a minimum partition order, other window families or lists) init returns ENCODER_ERROR instead of encoding something else. */
FLAC__bool FLAC__stream_encoder_set_do_mid_side_stereo(FLAC__StreamEncoder *encoder, FLAC__bool value);
FLAC__bool FLAC__stream_encoder_set_loose_mid_side_stereo(FLAC__StreamEncoder *encoder, FLAC__bool value);
FLAC__bool FLAC__stream_encoder_set_apodization(FLAC__StreamEncoder *encoder, const char *specification);
FLAC__bool FLAC__stream_encoder_set_max_lpc_order(FLAC__StreamEncoder *encoder, uint32_t value);
FLAC__bool FLAC__stream_encoder_set_qlp_coeff_precision(FLAC__StreamEncoder *encoder, uint32_t value);
FLAC__bool FLAC__stream_encoder_set_do_qlp_coeff_prec_search(FLAC__StreamEncoder *encoder, FLAC__bool value);
FLAC__bool FLAC__stream_encoder_set_do_exhaustive_model_search(FLAC__StreamEncoder *encoder, FLAC__bool value);
FLAC__bool FLAC__stream_encoder_set_min_residual_partition_order(FLAC__StreamEncoder *encoder, uint32_t value);
FLAC__bool FLAC__stream_encoder_set_max_residual_partition_order(FLAC__StreamEncoder *encoder, uint32_t value);
FLAC__bool FLAC__stream_encoder_set_rice_parameter_search_dist(FLAC__StreamEncoder *encoder, uint32_t value);
/* pyflac/encoder.py:141, :153-231 */
int FLAC__stream_encoder_get_state(const FLAC__StreamEncoder *encoder);
const char *FLAC__stream_encoder_get_resolved_state_string(const FLAC__StreamEncoder *encoder);
void FLAC__stream_encoder_get_verify_decoder_error_stats(const FLAC__StreamEncoder *encoder, FLAC__uint64 *absolute_sample, uint32_t *frame_number, uint32_t *channel, uint32_t *sample, FLAC__int32 *expected, FLAC__int32 *got);
FLAC__bool FLAC__stream_encoder_get_verify(const FLAC__StreamEncoder *encoder);
FLAC__bool FLAC__stream_encoder_get_streamable_subset(const FLAC__StreamEncoder *encoder);
uint32_t FLAC__stream_encoder_get_channels(const FLAC__StreamEncoder *encoder);
uint32_t FLAC__stream_encoder_get_bits_per_sample(const FLAC__StreamEncoder *encoder);
uint32_t FLAC__stream_encoder_get_sample_rate(const FLAC__StreamEncoder *encoder);
uint32_t FLAC__stream_encoder_get_blocksize(const FLAC__StreamEncoder *encoder);
FLAC__bool FLAC__stream_encoder_get_do_mid_side_stereo(const FLAC__StreamEncoder *encoder);
FLAC__bool FLAC__stream_encoder_get_loose_mid_side_stereo(const FLAC__StreamEncoder *encoder);
uint32_t FLAC__stream_encoder_get_max_lpc_order(const FLAC__StreamEncoder *encoder);
uint32_t FLAC__stream_encoder_get_qlp_coeff_precision(const FLAC__StreamEncoder *encoder);
FLAC__bool FLAC__stream_encoder_get_do_qlp_coeff_prec_search(const FLAC__StreamEncoder *encoder);
FLAC__bool FLAC__stream_encoder_get_do_escape_coding(const FLAC__StreamEncoder *encoder);
FLAC__bool FLAC__stream_encoder_get_do_exhaustive_model_search(const FLAC__StreamEncoder *encoder);
uint32_t FLAC__stream_encoder_get_min_residual_partition_order(const FLAC__StreamEncoder *encoder);
uint32_t FLAC__stream_encoder_get_max_residual_partition_order(const FLAC__StreamEncoder *encoder);
uint32_t FLAC__stream_encoder_get_rice_parameter_search_dist(const FLAC__StreamEncoder *encoder);
FLAC__uint64 FLAC__stream_encoder_get_total_samples_estimate(const FLAC__StreamEncoder *encoder);
FLAC__bool FLAC__stream_encoder_get_limit_min_bitrate(const FLAC__StreamEncoder *encoder);
/* pyflac/encoder.py:319 (init_stream), :401 (init_file), :115 (process_interleaved), :132 (finish) */
int FLAC__stream_encoder_init_stream(FLAC__StreamEncoder *encoder, FLAC__StreamEncoderWriteCallback write_callback, FLAC__StreamEncoderSeekCallback seek_callback, FLAC__StreamEncoderTellCallback tell_callback, FLAC__StreamEncoderMetadataCallback metadata_callback, void *client_data);
int FLAC__stream_encoder_init_FILE(FLAC__StreamEncoder *encoder, FILE *file, FLAC__StreamEncoderProgressCallback progress_callback, void *client_data);
int FLAC__stream_encoder_init_file(FLAC__StreamEncoder *encoder, const char *filename, FLAC__StreamEncoderProgressCallback progress_callback, void *client_data);
/* Ogg is compiled out of the reference binary too (--with-ogg=no): these return UNSUPPORTED_CONTAINER */
int FLAC__stream_encoder_init_ogg_stream(FLAC__StreamEncoder *encoder, FLAC__StreamEncoderReadCallback read_callback, FLAC__StreamEncoderWriteCallback write_callback, FLAC__StreamEncoderSeekCallback seek_callback, FLAC__StreamEncoderTellCallback tell_callback, FLAC__StreamEncoderMetadataCallback metadata_callback, void *client_data);
int FLAC__stream_encoder_init_ogg_FILE(FLAC__StreamEncoder *encoder, FILE *file, FLAC__StreamEncoderProgressCallback progress_callback, void *client_data);
int FLAC__stream_encoder_init_ogg_file(FLAC__StreamEncoder *encoder, const char *filename, FLAC__StreamEncoderProgressCallback progress_callback, void *client_data);
FLAC__bool FLAC__stream_encoder_finish(FLAC__StreamEncoder *encoder);
FLAC__bool FLAC__stream_encoder_process(FLAC__StreamEncoder *encoder, const FLAC__int32 *const buffer[], uint32_t samples);
FLAC__bool FLAC__stream_encoder_process_interleaved(FLAC__StreamEncoder *encoder, const FLAC__int32 buffer[], uint32_t samples);

/* ---- decoder callbacks: builder/decoder.py:368-375 ---- */
typedef int (*FLAC__StreamDecoderReadCallback)(const FLAC__StreamDecoder *, FLAC__byte buffer[], size_t *bytes, void *client_data);
typedef int (*FLAC__StreamDecoderSeekCallback)(const FLAC__StreamDecoder *, FLAC__uint64 absolute_byte_offset, void *client_data);
typedef int (*FLAC__StreamDecoderTellCallback)(const FLAC__StreamDecoder *, FLAC__uint64 *absolute_byte_offset, void *client_data);
typedef int (*FLAC__StreamDecoderLengthCallback)(const FLAC__StreamDecoder *, FLAC__uint64 *stream_length, void *client_data);
typedef FLAC__bool (*FLAC__StreamDecoderEofCallback)(const FLAC__StreamDecoder *, void *client_data);
typedef int (*FLAC__StreamDecoderWriteCallback)(const FLAC__StreamDecoder *, const FLAC__Frame *frame, const FLAC__int32 *const buffer[], void *client_data);
typedef void (*FLAC__StreamDecoderMetadataCallback)(const FLAC__StreamDecoder *, const FLAC__StreamMetadata *metadata, void *client_data);
typedef void (*FLAC__StreamDecoderErrorCallback)(const FLAC__StreamDecoder *, int status, void *client_data);

extern const char *const FLAC__StreamDecoderStateString[];        /* pyflac/decoder.py:46 */
extern const char *const FLAC__StreamDecoderInitStatusString[];   /* pyflac/decoder.py:60 */
extern const char *const FLAC__StreamDecoderErrorStatusString[];  /* pyflac/decoder.py:546 */

/* pyflac/decoder.py:85 */
FLAC__StreamDecoder *FLAC__stream_decoder_new(void);
void FLAC__stream_decoder_delete(FLAC__StreamDecoder *decoder);
FLAC__bool FLAC__stream_decoder_set_md5_checking(FLAC__StreamDecoder *decoder, FLAC__bool value);
/* builder/decoder.py:392-397 -- which metadata blocks reach the metadata callback (default: STREAMINFO only), as libFLAC's filter */
FLAC__bool FLAC__stream_decoder_set_metadata_respond(FLAC__StreamDecoder *decoder, int type);
FLAC__bool FLAC__stream_decoder_set_metadata_respond_application(FLAC__StreamDecoder *decoder, const FLAC__byte id[4]);
FLAC__bool FLAC__stream_decoder_set_metadata_respond_all(FLAC__StreamDecoder *decoder);
FLAC__bool FLAC__stream_decoder_set_metadata_ignore(FLAC__StreamDecoder *decoder, int type);
FLAC__bool FLAC__stream_decoder_set_metadata_ignore_application(FLAC__StreamDecoder *decoder, const FLAC__byte id[4]);
FLAC__bool FLAC__stream_decoder_set_metadata_ignore_all(FLAC__StreamDecoder *decoder);
/* pyflac/decoder.py:108 */
int FLAC__stream_decoder_get_state(const FLAC__StreamDecoder *decoder);
const char *FLAC__stream_decoder_get_resolved_state_string(const FLAC__StreamDecoder *decoder);
FLAC__bool FLAC__stream_decoder_get_md5_checking(const FLAC__StreamDecoder *decoder);
FLAC__uint64 FLAC__stream_decoder_get_total_samples(const FLAC__StreamDecoder *decoder);
uint32_t FLAC__stream_decoder_get_channels(const FLAC__StreamDecoder *decoder);
int FLAC__stream_decoder_get_channel_assignment(const FLAC__StreamDecoder *decoder);
uint32_t FLAC__stream_decoder_get_bits_per_sample(const FLAC__StreamDecoder *decoder);
uint32_t FLAC__stream_decoder_get_sample_rate(const FLAC__StreamDecoder *decoder);
uint32_t FLAC__stream_decoder_get_blocksize(const FLAC__StreamDecoder *decoder);
FLAC__bool FLAC__stream_decoder_get_decode_position(const FLAC__StreamDecoder *decoder, FLAC__uint64 *position);
/* pyflac/decoder.py:170,372 (init_stream), :271 (init_file) */
int FLAC__stream_decoder_init_stream(FLAC__StreamDecoder *decoder, FLAC__StreamDecoderReadCallback read_callback, FLAC__StreamDecoderSeekCallback seek_callback, FLAC__StreamDecoderTellCallback tell_callback, FLAC__StreamDecoderLengthCallback length_callback, FLAC__StreamDecoderEofCallback eof_callback, FLAC__StreamDecoderWriteCallback write_callback, FLAC__StreamDecoderMetadataCallback metadata_callback, FLAC__StreamDecoderErrorCallback error_callback, void *client_data);
int FLAC__stream_decoder_init_ogg_stream(FLAC__StreamDecoder *decoder, FLAC__StreamDecoderReadCallback read_callback, FLAC__StreamDecoderSeekCallback seek_callback, FLAC__StreamDecoderTellCallback tell_callback, FLAC__StreamDecoderLengthCallback length_callback, FLAC__StreamDecoderEofCallback eof_callback, FLAC__StreamDecoderWriteCallback write_callback, FLAC__StreamDecoderMetadataCallback metadata_callback, FLAC__StreamDecoderErrorCallback error_callback, void *client_data);
int FLAC__stream_decoder_init_FILE(FLAC__StreamDecoder *decoder, FILE *file, FLAC__StreamDecoderWriteCallback write_callback, FLAC__StreamDecoderMetadataCallback metadata_callback, FLAC__StreamDecoderErrorCallback error_callback, void *client_data);
int FLAC__stream_decoder_init_ogg_FILE(FLAC__StreamDecoder *decoder, FILE *file, FLAC__StreamDecoderWriteCallback write_callback, FLAC__StreamDecoderMetadataCallback metadata_callback, FLAC__StreamDecoderErrorCallback error_callback, void *client_data);
int FLAC__stream_decoder_init_file(FLAC__StreamDecoder *decoder, const char *filename, FLAC__StreamDecoderWriteCallback write_callback, FLAC__StreamDecoderMetadataCallback metadata_callback, FLAC__StreamDecoderErrorCallback error_callback, void *client_data);
int FLAC__stream_decoder_init_ogg_file(FLAC__StreamDecoder *decoder, const char *filename, FLAC__StreamDecoderWriteCallback write_callback, FLAC__StreamDecoderMetadataCallback metadata_callback, FLAC__StreamDecoderErrorCallback error_callback, void *client_data);
/* pyflac/decoder.py:99 (finish), :196,:294 (process_until_end_of_stream), :388 (process_single) */
FLAC__bool FLAC__stream_decoder_finish(FLAC__StreamDecoder *decoder);
FLAC__bool FLAC__stream_decoder_flush(FLAC__StreamDecoder *decoder);
FLAC__bool FLAC__stream_decoder_reset(FLAC__StreamDecoder *decoder);
FLAC__bool FLAC__stream_decoder_process_single(FLAC__StreamDecoder *decoder);
FLAC__bool FLAC__stream_decoder_process_until_end_of_metadata(FLAC__StreamDecoder *decoder);
FLAC__bool FLAC__stream_decoder_process_until_end_of_stream(FLAC__StreamDecoder *decoder);
FLAC__bool FLAC__stream_decoder_skip_single_frame(FLAC__StreamDecoder *decoder);
FLAC__bool FLAC__stream_decoder_seek_absolute(FLAC__StreamDecoder *decoder, FLAC__uint64 sample);

#ifdef __cplusplus
}
#endif
#endif
