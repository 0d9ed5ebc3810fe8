/*
 * oracle/ref_tap.cpp -- TEST INFRASTRUCTURE: per-sample oracle.  The reference decoder with its own signal debugger
 * enabled (NfcDecoder::setEnableDebug), recording into memory through oracle/tap/hw/RecordDevice.h instead of a 16-bit
 * WAV.  One row of 10 floats per sample clock (NfcTech.h:32-37, NfcTech.cpp:96-101):
 *    0 samplingValue x   1 filteredValue w   2 meanDeviation   3 signalAverage
 *    4..6 what the function that ran LAST on that sample stored (e.g. NfcA.cpp:259-264: filterIntegrate / period2,
 *         correlatedSD, searchValueThreshold; the sync markers 0.75 / 0.50 overwrite channel 5 on single samples)
 * Row r belongs to sample r; the row of the last sample is never flushed (NfcSignalDebug::block writes a row when the
 * clock advances, NfcTech.h:83-98).
 */
#include <cstdint>
#include <cstring>
#include <algorithm>

#include <lab/nfc/NfcDecoder.h>
#include <hw/SignalBuffer.h>
#include <hw/SignalType.h>

#include "tap/hw/RecordDevice.h"
#undef RecordDevice

extern "C" {

/* decode `n` magnitude samples with the techs of `enabled_mask` (bit0 A, bit1 B, bit2 F, bit3 V) and copy up to
 * cap_rows rows of 10 floats to rows_out; returns the number of rows the debugger produced */
long nfcref_tap_decode(const float *mag, uint64_t n, uint32_t sample_rate, uint32_t chunk, unsigned enabled_mask, float *rows_out, long cap_rows)
{
   hw::tapRows().clear();

   {
      lab::NfcDecoder decoder;
      decoder.setEnableNfcA((enabled_mask & 1) != 0);
      decoder.setEnableNfcB((enabled_mask & 2) != 0);
      decoder.setEnableNfcF((enabled_mask & 4) != 0);
      decoder.setEnableNfcV((enabled_mask & 8) != 0);
      decoder.setEnableDebug(true);

      if (chunk == 0)
         chunk = 65536;

      for (uint64_t pos = 0; pos < n; pos += chunk)
      {
         uint32_t len = (uint32_t) std::min<uint64_t>(chunk, n - pos);
         hw::SignalBuffer samples(len, 1, 1, sample_rate, 0, 0, hw::SignalType::SIGNAL_TYPE_RADIO_SAMPLES, 0);
         samples.put(mag + pos, len).flip();
         decoder.nextFrames(samples);
      }
   }

   const long rows = (long) (hw::tapRows().size() / 10);
   const long keep = std::min(rows, cap_rows);
   if (rows_out && keep > 0)
      std::memcpy(rows_out, hw::tapRows().data(), (size_t) keep * 10 * sizeof(float));
   hw::tapRows().clear();
   hw::tapRows().shrink_to_fit();
   return rows;
}

}
