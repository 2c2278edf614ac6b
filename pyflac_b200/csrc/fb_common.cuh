// fb_common.cuh -- data structures shared by the host engine and the sm_100a kernels.
//
// Vocabulary follows the reference's domain (pyflac/include/FLAC/format.h): streams, frames,
// subframes, partitions, Rice parameters.  One *frame* = one blocksize-long slice of one stream,
// all channels; one *signal* = one candidate subframe input (a channel, or mid / side).
#pragma once
#include <stdint.h>

namespace fb {

constexpr int kMaxChannels   = 8;
constexpr int kMaxSignals    = kMaxChannels;       // stereo mid/side analysis uses 4 signals (L, R, M, S)
constexpr int kMaxOrder      = 12;                 // compression levels 0..8 never exceed order 12 (stream_encoder.h:845-853)
constexpr int kMaxLags       = kMaxOrder + 1;
constexpr int kMaxPartOrder  = 6;                  // level table tops out at partition order 6
constexpr int kMaxParts      = 1 << kMaxPartOrder;
constexpr int kMaxApodSteps  = 9;                  // subdivide_tukey(3): 1 + 2 + 6 analyses
constexpr int kStreamPrologueBytes = 4 + 38 + 44;  // "fLaC" + STREAMINFO block + VORBIS_COMMENT(vendor)  

enum SubframeType : uint8_t { kConstant = 0, kVerbatim = 1, kFixed = 2, kLpc = 3 };

// Settings resolved from (compression_level, blocksize, bps, channels) exactly as
// FLAC__stream_encoder_init_* resolves them.  Uniform over one batch.
struct EncParams {
    uint32_t channels, bps, sample_rate;
    uint32_t blocksize;          // nominal blocksize (last frame of a stream may be shorter)
    uint32_t do_mid_side;        // analyse L,R,M,S and pick 1 of 4 channel assignments
    uint32_t max_lpc_order;      // 0 => fixed predictors only
    uint32_t qlp_precision;
    uint32_t max_part_order;     // level's limit; per frame it is min(this, ctz(N))
    uint32_t apod_parts;         // 1 = tukey(0.5); 2,3 = subdivide_tukey(parts)
    uint32_t rice_limit;         // 15 (bps<=16) or 31
    uint32_t container_bytes;    // 2 = int16 samples in HBM, 4 = int32
    uint32_t n_signals;          // channels (+2 when do_mid_side)
    uint32_t smem_stride;        // int32 words reserved per signal in shared memory (>= max blocksize, multiple of 4)
    uint32_t apod_steps;         // analysis kernel: LPC candidates per signal = steps of libFLAC's apodization walk (1, 3 or 9)
    uint32_t reserved0;
    uint32_t an_stride;          // analysis kernel: int32 words staged per signal (32 padded rows)
    uint32_t loose_frames;       // loose mid/side (levels 1, 4 on stereo): a full L/R/M/S decision every this many frames; 0 = off
    uint32_t limit_min_bitrate;  // up: process_subframes_ -- never emit a frame made of constant subframes only
    // libm-log guard (DESIGN.md "log guard"): decisions whose runner-up lies within guard_rel (relative) are logged for the host,
    // which repeats them with the libm log the reference links against and sends back overrides when it decides otherwise
    uint32_t guard_flip;         // test hook: take the runner-up inside the band (a deliberately wrong device decision)
    uint32_t guard_cap, guard_n_ovr, guard_pad;
    double   guard_rel;
    struct LogGuardEntry* guard_log;
    const struct LogGuardOverride* guard_ovr;
};

struct LogGuardEntry {
    uint32_t stream, frame_number, signal, step;
    uint32_t N, sbps, max_order, overhead;
    int32_t  guess, skip;        // what the device decided: LPC order, and whether the candidate was dropped as hopeless
    uint32_t kind, pad;          // bit 0: the order choice was inside the band, bit 1: the "don't even try" test was
    double   lperr[kMaxOrder];
};
struct LogGuardOverride { uint32_t stream, frame_number, signal, step; int32_t guess, skip; };

struct FrameDesc {
    uint64_t pcm_off;            // element offset (samples*channels) of the frame's first sample in the PCM buffer
    uint32_t blocksize;
    uint32_t frame_number;
    uint32_t window_off;         // float offset of this blocksize's window in the window table
    uint32_t stream;
    // loose mid/side: 0 = this frame makes the decision (full analysis); k > 0 = follow the channel assignment of the
    // frame k places earlier in the batch; kLeadForced | ca = follow a decision made before this batch
    uint32_t lead;
    uint32_t pad;
};
constexpr uint32_t kLeadForced = 0x80000000u;

// What the analysis kernel decides for one signal and the pack kernel turns into bits.
struct alignas(16) SubframePlan {
    uint8_t  type, order, wasted, sbps;
    uint8_t  precision, part_order, rice2, pad0;
    int32_t  shift;
    uint32_t bits_est;           // libFLAC's estimate (decision metric, not the coded size)
    int32_t  qlp[kMaxOrder];
    uint8_t  rice[kMaxParts];
};
static_assert(sizeof(SubframePlan) == 128, "plan layout");

// Optional per-signal trace (parity debugging against oracle/flac_oracle.c's fo_signal_trace).
struct SignalDebug {
    uint64_t fixed_err[5];
    int32_t  fixed_order;
    uint32_t fixed_bits;
    int32_t  is_constant;
    int32_t  n_apod;
    double   autoc[kMaxApodSteps][kMaxLags + 1];
    double   lpc_err[kMaxApodSteps][kMaxOrder];
    int32_t  lpc_order[kMaxApodSteps];
    uint32_t lpc_bits[kMaxApodSteps];
};

struct EncStats {
    unsigned long long log_ambiguous;   // LPC order / skip decisions that fell inside the libm-log guard band
    unsigned long long frames;
};

// md5_kernel launched ahead of its input (host -> host path): flags[c] != 0 once chunk c = streams cs[c] .. cs[c+1]-1 is in HBM.
struct Md5Gate {
    const uint32_t* flags;       // nullptr: the PCM is already there
    int nchunks;
    int cs[17];
};

// Per-stream results of the finalize step.
struct StreamInfoOut {
    uint64_t total_samples;
    uint64_t byte_off;           // offset of the stream's "fLaC" marker in the output arena
    uint64_t byte_len;           // prologue + all frames
    uint32_t min_framesize, max_framesize;
    uint32_t n_frames, pad;
    uint8_t  md5[16];
};

}  // namespace fb
