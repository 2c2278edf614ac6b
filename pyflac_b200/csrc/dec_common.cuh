// dec_common.cuh -- structures shared by the decode kernels and the host decode engine.
#pragma once
#include <stdint.h>

namespace fb {

constexpr uint32_t kDecSegBytes = 4096;       // bytes scanned per warp for frame-start candidates

enum DecStatus : int {
    kDecOk = 0,
    kDecPending = 1,
    kDecNotFlac = 2,         // no "fLaC" marker
    kDecBadMetadata = 3,
    kDecBadFrame = 4,        // reserved bits / impossible field values inside a frame
    kDecIncomplete = 5,      // frame runs past the end of the buffer
    kDecLostSync = 6,        // no frame starts where the previous one ended
    kDecCrcMismatch = 7,     // CRC-16 of a chained frame does not match
    kDecUnsupported = 8,     // valid FLAC outside this build's range
    kDecUnparseable = 9      // reserved field values inside a frame (libFLAC: UNPARSEABLE_STREAM)
};

constexpr int kDecMaxEvents = 16;     // error events logged per stream and batch (more are counted, not logged)

struct DecStreamMeta {
    uint32_t first_frame;    // byte offset of the first audio frame within the stream
    uint32_t sample_rate, channels, bps;
    uint32_t min_blocksize, max_blocksize;
    uint64_t total_samples;
    int32_t  status;
    uint8_t  md5[16];
    uint32_t have_last;      // headerless continuation: a frame was delivered before this batch ...
    uint32_t last_blocksize; // ... with this blocksize ...
    uint64_t next_sample;    // ... and the stream's next sample number behind it (gap detection across batches)
};

struct DecSegment { uint32_t stream; uint32_t start; uint32_t bytes; };

struct DecCand {
    uint32_t stream;
    uint32_t pos;            // byte offset of the sync code within the stream
    uint32_t blocksize;
    uint32_t hdr_bytes;      // header length including CRC-8
    uint32_t sample_rate;
    uint8_t  channels, ca, bps, variable;
    uint64_t number;         // frame or sample number
    int32_t  status;
    uint32_t end_pos;        // byte offset just past the frame's CRC-16; for a candidate that failed to decode: where the parse stopped
    uint32_t valid;          // 1 when the stream's frame chain passes through this candidate
    uint32_t sample_slot;
    uint64_t sample_off;     // inter-channel sample index of the frame within its stream
};

struct DecStreamResult {
    uint64_t total_samples;  // inter-channel samples decoded (incl. the silence that stands in for missing frames)
    uint64_t pcm_off;        // element offset of the stream's PCM in the output
    uint64_t consumed;       // bytes of the stream covered by metadata + everything up to the end of the last good frame
    uint32_t n_frames;       // frames delivered (good ones; silence is derived from their sample offsets)
    int32_t  status;         // first error met (decoding goes on behind it, as libFLAC's does)
    uint32_t sample_rate, channels, bps, max_blocksize;
    // errors in the order libFLAC's error callback would see them: ev_status = FLAC__StreamDecoderErrorStatus,
    // ev_frame = good frames delivered before the event
    uint32_t n_events;
    uint32_t gap_samples;    // samples of silence inserted for missing frames
    uint32_t ev_frame[kDecMaxEvents];
    uint8_t  ev_status[kDecMaxEvents];
    uint64_t next_sample;    // sample number behind the last delivered frame (by the frame headers), for the next batch of the stream
    uint32_t last_blocksize, have_last;
};

}  // namespace fb
