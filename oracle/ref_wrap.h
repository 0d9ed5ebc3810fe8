/*
 * oracle/ref_wrap.h -- C ABI over the UNMODIFIED reference decoder (test infrastructure only).
 *
 * The library built from this header (oracle/_ref/libnfcref.so) links the reference's own
 * lab::NfcDecoder sources where they lie under /root/reference (see oracle/Makefile); nothing in the
 * product path may include, link or call it.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker / CPU baseline.
 */
#ifndef NFC_ORACLE_REF_WRAP_H
#define NFC_ORACLE_REF_WRAP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* POD mirror of lab::RawFrame (reference: lab-data RawFrame.cpp:26-39) */
typedef struct nfcref_frame
{
   uint32_t tech_type;    /* FrameTech  0x0100..0x0104 */
   uint32_t frame_type;   /* FrameType  0x0100..0x0103 */
   uint32_t frame_flags;  /* FrameFlags bit-or */
   uint32_t frame_phase;  /* FramePhase */
   uint32_t frame_rate;   /* symbols / s */
   uint32_t length;       /* payload bytes */
   uint64_t sample_start;
   uint64_t sample_end;
   uint64_t sample_rate;
   double time_start;
   double time_end;
   double date_time;
   uint8_t data[512];
} nfcref_frame;

typedef struct nfcref_decoder nfcref_decoder;

/* create a reference decoder with all four techs enabled (mirrors test-sdr main.cpp:149-156) */
nfcref_decoder *nfcref_create(void);

void nfcref_destroy(nfcref_decoder *d);

/* tech enable mask: bit0 A, bit1 B, bit2 F, bit3 V */
void nfcref_set_enabled(nfcref_decoder *d, unsigned mask);

/* NaN leaves a value unchanged (reference semantics) */
void nfcref_set_thresholds(nfcref_decoder *d, int tech, float corr, float mod_min, float mod_max);

void nfcref_set_power_threshold(nfcref_decoder *d, float value);

void nfcref_set_stream_time(nfcref_decoder *d, long t);

/*
 * Feed n magnitude samples (float32, SIGNAL_TYPE_RADIO_SAMPLES) in buffers of `chunk` samples, exactly like
 * test-sdr does with chunk = 65536.  Decoded frames (ALL frame types, carrier frames included) are appended
 * to out[0..cap); returns the number of frames produced (may exceed cap; only cap are stored).
 */
long nfcref_push(nfcref_decoder *d, const float *mag, uint64_t n, uint32_t sample_rate, uint32_t chunk,
                 nfcref_frame *out, long cap);

/* flush: nextFrames({}) -> one carrier frame (NfcDecoder.cpp:449-463) */
long nfcref_flush(nfcref_decoder *d, nfcref_frame *out, long cap);

/* one-shot convenience: create, push everything, destroy */
long nfcref_decode(const float *mag, uint64_t n, uint32_t sample_rate, uint32_t chunk, unsigned enabled_mask,
                   nfcref_frame *out, long cap);

/* reference IQ->magnitude, scalar path (RadioDeviceTask.cpp:627-637): sqrtf(I*I+Q*Q) */
void nfcref_iq_magnitude(const float *iq, uint64_t n, float *mag);

/*
 * Time the reference on `threads` host threads: stream s (s < n_streams) is mag + s*n, one NfcDecoder per thread,
 * streams distributed round-robin.  If iq != NULL the IQ->magnitude step is done (and timed) first from
 * iq + 2*s*n.  Returns elapsed seconds (wall clock over the whole batch); *frames_out = total frames.
 */
double nfcref_time_batch(const float *mag, const float *iq, uint64_t n, uint32_t n_streams, uint32_t sample_rate,
                         uint32_t chunk, int threads, long *frames_out);

/*
 * One 64-bit hash per reference frame (every field RawFrame::operator== compares plus the payload): hashes[s * cap + i]
 * for frame i of stream s, counts[s] = its frame count.  The full-size differential of bench.py compares these with the
 * hashes of the GPU's frames (nfc_laboratory_b200/dist.py frame_hashes).  Returns elapsed seconds.
 */
double nfcref_hash_batch(const float *mag, const float *iq, uint64_t n, uint32_t n_streams, uint32_t sample_rate, uint32_t chunk, int threads,
                         uint64_t *hashes, uint32_t cap, uint32_t *counts);

#ifdef __cplusplus
}
#endif

#endif
