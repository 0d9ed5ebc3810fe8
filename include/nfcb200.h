/*
 * nfcb200.h -- C ABI of the B200-native NFC IQ demodulation path (libnfcb200.so).
 *
 * This is the drop-in boundary: the reference has no FFI on this path, its seam is link-time -- whoever provides
 * liblab-radio provides lab::NfcDecoder (lab-radio/src/main/include/lab/nfc/NfcDecoder.h:33-122).  The shim in
 * nfc_laboratory_b200/shim/ implements that class on top of the entry points below; every entry point
 * names the reference interface it replaces.  Plain pointers and sizes only, no torch / CUDA types.
 *
 * All functions return 0 on success or a negative nfcb200_status; nfcb200_last_error() gives the text.  A handle is
 * single-caller like the reference decoder (RadioDecoderTask.cpp:92-151); CUDA streams and events are internal.
 * There is no CPU fallback: without a CUDA device nfcb200_create() fails with NFCB200_ERR_NO_DEVICE.
 */
#ifndef NFCB200_H
#define NFCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nfcb200_handle nfcb200_handle;

typedef enum nfcb200_status
{
   NFCB200_OK = 0,
   NFCB200_ERR_NO_DEVICE = -1,   /* no CUDA device / driver: the product has no CPU path        */
   NFCB200_ERR_INVALID = -2,     /* bad argument (reference: RadioDecoderTask InvalidConfig = -2) */
   NFCB200_ERR_CUDA = -3,        /* CUDA runtime error, see nfcb200_last_error                    */
   NFCB200_ERR_CAPACITY = -4,    /* output or internal pool too small; call again with more room  */
   NFCB200_ERR_UNSUPPORTED = -5  /* sample rate / signal type outside what the kernels implement  */
} nfcb200_status;

/* sample formats.  1 and 2 are hw::SignalType values (hw-dev SignalType.h:27-36); 3 and 4 are the WAV ingest formats
 * of hw::RecordDevice (RecordDevice.cpp:281-311: int16 / 32768.f) decoded on the device */
typedef enum nfcb200_sigtype
{
   NFCB200_SIG_IQ_F32 = 1,   /* SIGNAL_TYPE_RADIO_IQ: interleaved float32 I,Q (RadioDeviceTask.cpp:547-655 fused in) */
   NFCB200_SIG_MAG_F32 = 2,  /* SIGNAL_TYPE_RADIO_SAMPLES: float32 magnitude (NfcDecoder::nextFrames input)          */
   NFCB200_SIG_MAG_S16 = 3,  /* mono int16 PCM                                                                       */
   NFCB200_SIG_IQ_S16 = 4    /* interleaved int16 I,Q                                                                */
} nfcb200_sigtype;

enum { NFCB200_TECH_A = 0, NFCB200_TECH_B = 1, NFCB200_TECH_F = 2, NFCB200_TECH_V = 3 };

/* POD mirror of lab::RawFrame (lab-data RawFrame.cpp:26-39); `stream` is the index of the capture in the batch */
typedef struct nfcb200_frame
{
   uint32_t stream;
   uint32_t tech_type;    /* FrameTech  0x0100 any, 0x0101 A, 0x0102 B, 0x0103 F, 0x0104 V        */
   uint32_t frame_type;   /* FrameType  0x0100 carrier off, 0x0101 carrier on, 0x0102 poll, 0x0103 listen */
   uint32_t frame_flags;  /* FrameFlags: ShortFrame 1, Encrypted 2, Truncated 8, ParityError 0x10, CrcError 0x20, SyncError 0x40 */
   uint32_t frame_phase;  /* FramePhase: 0x0101 carrier, 0x0102 selection, 0x0103 application      */
   uint32_t frame_rate;   /* symbols per second                                                    */
   uint32_t length;       /* payload bytes                                                         */
   uint32_t reserved;
   uint64_t sample_start;
   uint64_t sample_end;
   uint64_t sample_rate;
   double time_start;     /* double(sample_start) / double(sample_rate)                            */
   double time_end;
   double date_time;      /* streamTime + time_start                                               */
   uint8_t data[512];     /* `length` payload bytes, zero padded to the next 64-byte boundary; the rest is not written */
} nfcb200_frame;

/* decoder configuration: the NfcDecoder setters (NfcDecoder.h:47-117) / RadioDecoderTask JSON keys
 * (RadioDecoderTask.cpp:207-366).  nfcb200_config_default() fills the reference defaults. */
typedef struct nfcb200_config
{
   int device;                    /* CUDA device ordinal                                                          */
   uint32_t enabled;              /* bit t enables tech t: setEnableNfcA/B/F/V                                    */
   float power_level_threshold;   /* setPowerLevelThreshold, default 0.01                                         */
   float correlation_threshold[4];/* setCorrelationThresholdNfcX: 0.75 0.50 0.50 0.50                             */
   float modulation_min[4];       /* setModulationThresholdNfcX min: 0.90 0.10 0.10 0.90                          */
   float modulation_max[4];       /* setModulationThresholdNfcX max: 1.00 0.90 0.90 1.00                          */
   uint32_t stream_time;          /* setStreamTime                                                                */
   uint32_t use_tma;              /* 1: stage screening tiles with cp.async.bulk (default); 0: plain loads (debug) */
   uint32_t max_rounds;           /* bound on speculation rounds (0 = default)                                    */
   uint32_t segments_per_lane;    /* 0 = choose from the batch size; >= 1 forces the lane grouping                */
   uint32_t exact;                /* 1: one warp lane per stream carries the reference's float state across the whole
                                     capture (running sums included): bit-exact on float input, slower.  0 (default): lanes
                                     cold-start their running sums -- exact on 16-bit input, sums within 2e-6 on float input */
   uint32_t reserved[3];
} nfcb200_config;

/* counters and device timings of the last nfcb200_decode_batch call */
typedef struct nfcb200_stats
{
   uint64_t samples;          /* n_streams * n_samples                          */
   uint64_t blocks;           /* screening blocks                               */
   uint64_t active_blocks;    /* blocks handed to lanes                         */
   uint64_t segments;         /* segments found by the screen                   */
   uint64_t lanes;            /* lanes (groups of consecutive segments)         */
   uint64_t live_lanes;       /* segments not swallowed by a predecessor        */
   uint64_t lane_runs;        /* lane executions over all rounds                */
   uint64_t lane_samples;     /* samples stepped by lanes over all rounds       */
   uint64_t rounds;           /* speculation rounds                             */
   uint64_t frames;           /* frames returned                                */
   uint64_t kernel_launches;  /* kernels of this library launched by the call   */
   float ms_h2d;              /* host -> device copy of the samples (0 when the input is device resident) */
   float ms_screen;           /* K1 screening kernel                            */
   float ms_segment;          /* segment construction + front pass              */
   float ms_lanes;            /* all lane + chain kernels                       */
   float ms_gather;           /* frame gather incl. device -> host copy         */
   float ms_total;            /* whole call, device events                      */
   float ms_wall;             /* whole call, host clock                         */
   float ms_front;            /* front pass (part of ms_segment .. ms_lanes: own event pair) */
   float straggler_lanes;     /* thread lanes that held the launch and were decoded again by a warp lane (a count)  */
   uint64_t feature_samples;  /* samples the front pass wrote to the feature pool */
} nfcb200_stats;

void nfcb200_config_default(nfcb200_config *cfg);

/* replaces `new NfcDecoder()` + setters (RadioDecoderTask.cpp:67, 224-341; test-sdr main.cpp:149-156) */
int nfcb200_create(const nfcb200_config *cfg, nfcb200_handle **out);

/* replaces the NfcDecoder destructor / cleanup() (NfcDecoder.cpp:365-369) */
void nfcb200_destroy(nfcb200_handle *h);

/* re-configure; takes effect at the next decode (NfcDecoder::initialize, NfcDecoder.cpp:295-360) */
int nfcb200_configure(nfcb200_handle *h, const nfcb200_config *cfg);

/*
 * Batch decode: n_streams independent captures of n_samples each, laid out [n_streams][n_samples] in `sigtype`
 * format.  Replaces one NfcDecoder instance per stream fed by nextFrames() until exhausted (NfcDecoder.cpp:374-467),
 * preceded by the IQ -> magnitude step of RadioDeviceTask (RadioDeviceTask.cpp:547-655) when sigtype is an IQ format.
 *   samples_on_device != 0: `samples` is a device pointer on the handle's device (no copy)
 *   samples_on_device == 0: `samples` is host memory (pinned memory copies asynchronously)
 * Frames are written to out[0 .. min(*n_out, cap)) ordered by (stream, decode order); *n_out is the number of frames
 * decoded (if it exceeds cap the call returns NFCB200_ERR_CAPACITY after filling cap frames).  Carrier on/off frames
 * are included (filter on frame_type like test-sdr main.cpp:171-174 if not wanted).
 */
int nfcb200_decode_batch(nfcb200_handle *h, const void *samples, int samples_on_device, int sigtype, uint32_t n_streams, uint64_t n_samples,
                         uint32_t sample_rate, nfcb200_frame *out, uint64_t cap, uint64_t *n_out);

/*
 * Streaming decode of ONE capture in arbitrary chunks: replaces NfcDecoder::nextFrames(SignalBuffer) called per
 * buffer by RadioDecoderTask::signalDecode (RadioDecoderTask.cpp:377-401) and test-sdr (main.cpp:163-176).  Frames
 * whose decode is complete are returned; state is carried across calls.  n == 0 flushes: the pending tail is decoded
 * as end of stream and, like nextFrames({}) (NfcDecoder.cpp:449-463), one carrier frame at the current clock is added.
 */
int nfcb200_stream_push(nfcb200_handle *h, const void *samples, int sigtype, uint64_t n, uint32_t sample_rate, nfcb200_frame *out, uint64_t cap,
                        uint64_t *n_out);

/*
 * Frames of the stream that did not fit the buffer of an earlier nfcb200_stream_push (which then returned
 * NFCB200_ERR_CAPACITY after delivering `cap` frames): delivers up to cap of them, *n_left = how many remain.  Nothing is
 * lost on overflow; the stream state has advanced regardless.
 */
int nfcb200_stream_pending(nfcb200_handle *h, nfcb200_frame *out, uint64_t cap, uint64_t *n_out, uint64_t *n_left);

/*
 * Per-sample value tap on the device (debug / test): ONE lane of the decoder over a host capture, one row of 8 floats per
 * sample from `first` on -- [x, w, deviation, average, channel 4, channel 5, lock state, 0], the channels of the
 * reference's signal debugger (NfcTech.h:32-37, NfcDecoder::setEnableDebug), NaN where a channel was not written.
 * first = 0: the exact stream start; otherwise a cold start with `warm` warm-up samples.  rows must hold (n - first) * 8
 * floats.  Slow by design (one GPU thread); the product kernels compile the taps away.
 */
int nfcb200_debug_trace(const nfcb200_config *cfg, const void *samples, int sigtype, uint64_t n, uint32_t sample_rate, uint32_t first,
                        uint32_t warm, float *rows);

/* forget the streaming state (NfcDecoder::initialize on a sample-rate change, NfcDecoder.cpp:383-388) */
int nfcb200_stream_reset(nfcb200_handle *h);

int nfcb200_get_stats(nfcb200_handle *h, nfcb200_stats *stats);

/* debug tap: per-block screening flags of the last batch, [n_streams][n_blocks] bytes (bit0 trigger, bit1 active) */
int nfcb200_get_block_flags(nfcb200_handle *h, uint8_t *out, uint64_t cap, uint64_t *n_blocks_per_stream);

/*
 * Time shards of ONE long capture (BASELINE.json configs[4]): a shard that does not start at the capture's first sample
 * continues its predecessor's decoder.  After a single-stream nfcb200_decode_batch, nfcb200_carry_before returns the carry
 * of the decoder in front of the first lane that begins at or after `sample` (*lane_begin: that lane's begin, an idle point
 * of the capture; ~0 when there is none).  nfcb200_set_carry hands such a carry to the NEXT single-stream decode of a handle
 * (one shot; clock_shift is subtracted from the absolute sample times it holds: the next window counts from its own first
 * sample).  The blob is opaque, nfcb200_carry_size() bytes: protocol state (FSD / FWT / SFGT from RATS / ATS / ATTRIB, the
 * Encrypted flag, lastCommand -- NfcA.cpp:1592-1790, NfcB.cpp:1153-1258), carrier flags, the carrier edge time.
 * A window decoded from an injected carry is a CONTINUATION, not a decoder start: the reference's start-of-stream
 * behaviour (two carrier-off frames at samples 0 and 1, detectors held off for 1 024 samples) does not apply, its first
 * 2 048 samples are warm-up only (start the window that far, or further, in front of the idle point the carry belongs
 * to: nfc_laboratory_b200/dist.py decode_long_capture_carry uses 6 144).
 */
int nfcb200_carry_size(void);
int nfcb200_default_carry(nfcb200_handle *h, void *blob, uint64_t cap); /* what a cold-started lane assumes: power-on state, carrier on */
int nfcb200_set_carry(nfcb200_handle *h, const void *blob, uint64_t size, uint32_t clock_shift);
int nfcb200_carry_before(nfcb200_handle *h, uint64_t sample, void *blob, uint64_t cap, uint64_t *size, uint64_t *lane_begin);

/*
 * Multi-GPU frame gather (no counterpart in the reference, which has no second device): the frames of the
 * last nfcb200_decode_batch call as they sit in DEVICE memory -- ordered by (stream, time), 128-byte records (stream,
 * header fields, the first 80 payload bytes) plus 128-byte extension chunks for longer payloads -- so that a rank can hand
 * them to NCCL without a host round trip; nfcb200_emit_records converts gathered records (host memory) into ABI frames
 * on the receiving rank, adding stream_offset to the stream index.  The pointers stay valid until the next decode.
 */
int nfcb200_device_frames(nfcb200_handle *h, const void **records, uint64_t *n_records, const void **ext, uint64_t *n_ext_chunks);
int nfcb200_emit_records(nfcb200_handle *h, const void *records, uint64_t n_records, const void *ext, uint64_t n_ext_chunks, uint32_t stream_offset,
                         uint32_t sample_rate, nfcb200_frame *out, uint64_t cap, uint64_t *n_out);

/*
 * Wire format of the multi-GPU frame gather (no reference counterpart: the reference is one decoder per process).
 * Writes [u64 count][count x 80-byte frame headers, stream index raised by stream_offset][payload bytes back to back]
 * to `out`; *n_bytes is the size needed (call with out == NULL to size the buffer).  Host-only, no CUDA call.
 */
int nfcb200_pack_frames(const nfcb200_frame *frames, uint64_t n, uint32_t stream_offset, uint8_t *out, uint64_t cap, uint64_t *n_bytes);

const char *nfcb200_last_error(void);

/* library / build identification, e.g. "nfcb200 0.1 sm_100a" */
const char *nfcb200_version(void);

#ifdef __cplusplus
}
#endif

#endif
