/*
 * oracle/ref_wrap.cpp -- thin C wrapper over the unmodified reference lab::NfcDecoder.
 * TEST INFRASTRUCTURE ONLY (see ref_wrap.h).  Compiled together with the reference's own sources by
 * oracle/Makefile into oracle/_ref/libnfcref.so; no reference source is copied into this repository.
 *
 * Mirrors the call sequence of the reference regression tool (src/nfc-test/test-sdr/src/main/cpp/main.cpp:141-180):
 * SignalBuffer(65536 * ch, ch, 1, rate, 0, 0, SIGNAL_TYPE_RADIO_SAMPLES) -> decoder.nextFrames(samples).
 */
#include <cmath>
#include <cstring>
#include <chrono>
#include <thread>
#include <vector>
#include <list>
#include <atomic>

#include <hw/SignalType.h>
#include <hw/SignalBuffer.h>
#include <lab/data/RawFrame.h>
#include <lab/nfc/NfcDecoder.h>

#include "ref_wrap.h"

struct nfcref_decoder
{
   lab::NfcDecoder decoder;
};

static void convert(const lab::RawFrame &f, nfcref_frame *o)
{
   std::memset(o, 0, sizeof(*o));
   o->tech_type = f.techType();
   o->frame_type = f.frameType();
   o->frame_flags = f.frameFlags();
   o->frame_phase = f.framePhase();
   o->frame_rate = f.frameRate();
   o->sample_start = f.sampleStart();
   o->sample_end = f.sampleEnd();
   o->sample_rate = f.sampleRate();
   o->time_start = f.timeStart();
   o->time_end = f.timeEnd();
   o->date_time = f.dateTime();
   unsigned int n = f.limit();
   if (n > sizeof(o->data))
      n = sizeof(o->data);
   o->length = n;
   for (unsigned int i = 0; i < n; i++)
      o->data[i] = f[i];
}

static long feed(lab::NfcDecoder &decoder, const float *mag, uint64_t n, uint32_t rate, uint32_t chunk, nfcref_frame *out, long cap, long produced)
{
   if (chunk == 0)
      chunk = 65536;

   for (uint64_t pos = 0; pos < n; pos += chunk)
   {
      uint32_t len = (uint32_t) std::min<uint64_t>(chunk, n - pos);

      hw::SignalBuffer samples(len, 1, 1, rate, 0, 0, hw::SignalType::SIGNAL_TYPE_RADIO_SAMPLES, 0);

      samples.put(mag + pos, len).flip();

      for (const lab::RawFrame &frame: decoder.nextFrames(samples))
      {
         if (out && produced < cap)
            convert(frame, out + produced);

         produced++;
      }
   }

   return produced;
}

extern "C" {

nfcref_decoder *nfcref_create(void)
{
   auto *d = new nfcref_decoder();
   d->decoder.setEnableNfcA(true);
   d->decoder.setEnableNfcB(true);
   d->decoder.setEnableNfcF(true);
   d->decoder.setEnableNfcV(true);
   return d;
}

void nfcref_destroy(nfcref_decoder *d)
{
   delete d;
}

void nfcref_set_enabled(nfcref_decoder *d, unsigned mask)
{
   d->decoder.setEnableNfcA(mask & 1);
   d->decoder.setEnableNfcB(mask & 2);
   d->decoder.setEnableNfcF(mask & 4);
   d->decoder.setEnableNfcV(mask & 8);
}

void nfcref_set_thresholds(nfcref_decoder *d, int tech, float corr, float mod_min, float mod_max)
{
   switch (tech)
   {
      case 0:
         d->decoder.setCorrelationThresholdNfcA(corr);
         d->decoder.setModulationThresholdNfcA(mod_min, mod_max);
         break;
      case 1:
         d->decoder.setCorrelationThresholdNfcB(corr);
         d->decoder.setModulationThresholdNfcB(mod_min, mod_max);
         break;
      case 2:
         d->decoder.setCorrelationThresholdNfcF(corr);
         d->decoder.setModulationThresholdNfcF(mod_min, mod_max);
         break;
      case 3:
         d->decoder.setCorrelationThresholdNfcV(corr);
         d->decoder.setModulationThresholdNfcV(mod_min, mod_max);
         break;
   }
}

void nfcref_set_power_threshold(nfcref_decoder *d, float value)
{
   d->decoder.setPowerLevelThreshold(value);
}

void nfcref_set_stream_time(nfcref_decoder *d, long t)
{
   d->decoder.setStreamTime(t);
}

long nfcref_push(nfcref_decoder *d, const float *mag, uint64_t n, uint32_t sample_rate, uint32_t chunk, nfcref_frame *out, long cap)
{
   return feed(d->decoder, mag, n, sample_rate, chunk, out, cap, 0);
}

long nfcref_flush(nfcref_decoder *d, nfcref_frame *out, long cap)
{
   long produced = 0;

   for (const lab::RawFrame &frame: d->decoder.nextFrames({}))
   {
      if (out && produced < cap)
         convert(frame, out + produced);

      produced++;
   }

   return produced;
}

long nfcref_decode(const float *mag, uint64_t n, uint32_t sample_rate, uint32_t chunk, unsigned enabled_mask, nfcref_frame *out, long cap)
{
   nfcref_decoder *d = nfcref_create();
   nfcref_set_enabled(d, enabled_mask);
   long produced = nfcref_push(d, mag, n, sample_rate, chunk, out, cap);
   nfcref_destroy(d);
   return produced;
}

void nfcref_iq_magnitude(const float *iq, uint64_t n, float *mag)
{
   for (uint64_t i = 0; i < n; i++)
   {
      float I = iq[2 * i + 0];
      float Q = iq[2 * i + 1];
      mag[i] = sqrtf(I * I + Q * Q);
   }
}

double nfcref_time_batch(const float *mag, const float *iq, uint64_t n, uint32_t n_streams, uint32_t sample_rate, uint32_t chunk, int threads, long *frames_out)
{
   if (threads < 1)
      threads = 1;

   std::atomic<long> total {0};
   std::atomic<uint32_t> next {0};

   auto start = std::chrono::steady_clock::now();

   std::vector<std::thread> pool;

   for (int t = 0; t < threads; t++)
   {
      pool.emplace_back([&]() {
         std::vector<float> scratch;

         for (;;)
         {
            uint32_t s = next.fetch_add(1);

            if (s >= n_streams)
               break;

            const float *src = mag ? mag + (uint64_t) s * n : nullptr;

            if (iq)
            {
               scratch.resize(n);
               nfcref_iq_magnitude(iq + 2 * (uint64_t) s * n, n, scratch.data());
               src = scratch.data();
            }

            lab::NfcDecoder decoder;
            decoder.setEnableNfcA(true);
            decoder.setEnableNfcB(true);
            decoder.setEnableNfcF(true);
            decoder.setEnableNfcV(true);

            total += feed(decoder, src, n, sample_rate, chunk, nullptr, 0, 0);
         }
      });
   }

   for (auto &th: pool)
      th.join();

   auto stop = std::chrono::steady_clock::now();

   if (frames_out)
      *frames_out = total.load();

   return std::chrono::duration<double>(stop - start).count();
}


/*
 * Full-size differential (bench.py cpu_baseline leg, tests): one 64-bit hash per reference frame over every field
 * RawFrame::operator== compares (lab-data RawFrame.cpp:82-98) plus the payload, hashes[s * cap + i] for frame i of stream s,
 * counts[s] = frames of stream s (may exceed cap: the excess is not stored).  nfc_laboratory_b200/dist.py frame_hashes()
 * computes the same value from the GPU's frames.  Returns the seconds spent (IQ -> magnitude included when iq != NULL).
 */
static uint64_t hash_frame(const lab::RawFrame &f)
{
   const uint64_t fields[8] = {f.techType(), f.frameType(), f.frameFlags(), f.framePhase(), f.frameRate(), f.limit() > 512 ? 512u : f.limit(),
                               f.sampleStart(), f.sampleEnd()};
   uint64_t h = 0;
   for (int k = 0; k < 8; k++)
      h = h * 0x100000001B3ull + fields[k] + (uint64_t) (k + 1);
   unsigned int n = f.limit() > 512 ? 512 : f.limit();
   for (unsigned int j = 0; j < n; j++)
      h += ((uint64_t) (unsigned char) f[j] + 1ull) * ((uint64_t) j * 0x9E3779B97F4A7C15ull + 1ull);
   return h;
}

double nfcref_hash_batch(const float *mag, const float *iq, uint64_t n, uint32_t n_streams, uint32_t sample_rate, uint32_t chunk, int threads,
                         uint64_t *hashes, uint32_t cap, uint32_t *counts)
{
   if (threads < 1)
      threads = 1;
   if (chunk == 0)
      chunk = 65536;

   std::atomic<uint32_t> next {0};
   auto start = std::chrono::steady_clock::now();
   std::vector<std::thread> pool;

   for (int t = 0; t < threads; t++)
   {
      pool.emplace_back([&]() {
         std::vector<float> scratch;

         for (;;)
         {
            uint32_t s = next.fetch_add(1);
            if (s >= n_streams)
               break;

            const float *src = mag ? mag + (uint64_t) s * n : nullptr;
            if (iq)
            {
               scratch.resize(n);
               nfcref_iq_magnitude(iq + 2 * (uint64_t) s * n, n, scratch.data());
               src = scratch.data();
            }

            lab::NfcDecoder decoder;
            decoder.setEnableNfcA(true);
            decoder.setEnableNfcB(true);
            decoder.setEnableNfcF(true);
            decoder.setEnableNfcV(true);

            uint32_t produced = 0;
            for (uint64_t pos = 0; pos < n; pos += chunk)
            {
               uint32_t len = (uint32_t) std::min<uint64_t>(chunk, n - pos);
               hw::SignalBuffer samples(len, 1, 1, sample_rate, 0, 0, hw::SignalType::SIGNAL_TYPE_RADIO_SAMPLES, 0);
               samples.put(src + pos, len).flip();
               for (const lab::RawFrame &frame: decoder.nextFrames(samples))
               {
                  if (produced < cap)
                     hashes[(uint64_t) s * cap + produced] = hash_frame(frame);
                  produced++;
               }
            }
            counts[s] = produced;
         }
      });
   }

   for (auto &th: pool)
      th.join();

   return std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
}

}
