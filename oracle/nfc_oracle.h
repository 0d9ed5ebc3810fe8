/*
 * oracle/nfc_oracle.h -- C API of the plain-C restatement of the reference decoder (oracle/nfc_oracle.c).
 * TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; never the product.
 */
#ifndef NFC_ORACLE_H
#define NFC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* same layout as oracle/ref_wrap.h nfcref_frame (POD mirror of lab::RawFrame) */
typedef struct nfcoracle_frame
{
   uint32_t tech_type, frame_type, frame_flags, frame_phase, frame_rate, length;
   uint64_t sample_start, sample_end, sample_rate;
   double time_start, time_end, date_time;
   uint8_t data[512];
} nfcoracle_frame;

typedef struct nfcoracle_decoder nfcoracle_decoder;

nfcoracle_decoder *nfcoracle_create(void);
void nfcoracle_destroy(nfcoracle_decoder *d);
void nfcoracle_set_enabled(nfcoracle_decoder *d, unsigned mask); /* bit0 A, bit1 B, bit2 F, bit3 V */

/* feed n float magnitude samples; frames of ALL types are appended to out[0..cap); returns the number produced */
long nfcoracle_push(nfcoracle_decoder *d, const float *mag, uint64_t n, uint32_t sample_rate, nfcoracle_frame *out, long cap);
long nfcoracle_decode(const float *mag, uint64_t n, uint32_t sample_rate, unsigned enabled_mask, nfcoracle_frame *out, long cap);
void nfcoracle_iq_magnitude(const float *iq, uint64_t n, float *mag);

#ifdef __cplusplus
}
#endif

#endif
