/*
 * NfcDecoderB200.cpp -- drop-in implementation of the reference class lab::NfcDecoder on top of libnfcb200.so.
 *
 * The reference has no plugin / FFI seam on this path: whoever provides liblab-radio provides the decoder
 * (lab-radio/src/main/include/lab/nfc/NfcDecoder.h:33-122, pimpl std::shared_ptr<Impl>).  This file is compiled INSTEAD
 * of the reference's lab-radio/src/main/cpp/{NfcDecoder,NfcTech}.cpp and tech/Nfc{A,B,F,V}.cpp, against the reference's
 * own UNMODIFIED headers, and forwards every call to the C ABI of include/nfcb200.h.  Everything above it -- test-sdr,
 * RadioDecoderTask, nfc-rx, the Qt app -- compiles and links unchanged.
 *
 * Behaviour mirrored from NfcDecoder.cpp: setters only store (they take effect at the next initialize(), i.e. at the
 * next sample-rate change or explicit initialize() call, NfcDecoder.cpp:383-388); nextFrames() of an invalid buffer
 * flushes and returns one carrier frame (NfcDecoder.cpp:449-463); SIGNAL_TYPE_RADIO_IQ buffers (stride 2) are accepted
 * as an extension (the reference would spin forever on them) and take the fused IQ path.
 */
#include <cmath>
#include <list>
#include <memory>
#include <stdexcept>
#include <vector>

#include <hw/SignalType.h>
#include <hw/SignalBuffer.h>

#include <lab/data/RawFrame.h>
#include <lab/nfc/NfcDecoder.h>

#include <nfcb200.h>

namespace lab {

struct NfcDecoder::Impl
{
   nfcb200_config cfg {};
   nfcb200_handle *handle = nullptr;
   bool debugEnabled = false;
   long sampleRate = 0;
   bool dirty = true; // configuration changed since the last initialize; applied by the next nextFrames / initialize
   std::vector<nfcb200_frame> frames;
   int lastStatus = 0;       // status of the last library call (0 = ok): lab::NfcDecoder has no error channel and its callers
   std::string lastMessage;  // (RadioDecoderTask.cpp:377-401) no try / catch -- failures are reported here, never thrown

   Impl()
   {
      nfcb200_config_default(&cfg);
      frames.resize(4096);
   }

   ~Impl()
   {
      if (handle)
         nfcb200_destroy(handle);
   }

   bool note(int rc)
   {
      lastStatus = rc;
      lastMessage = rc ? nfcb200_last_error() : "";
      return rc == 0;
   }

   bool ensure()
   {
      if (!handle)
      {
         if (!note(nfcb200_create(&cfg, &handle)))
         {
            handle = nullptr;
            return false;
         }
         dirty = false;
      }
      return true;
   }

   void initialize()
   {
      if (!ensure())
         return;
      note(nfcb200_configure(handle, &cfg));
      nfcb200_stream_reset(handle);
      dirty = false;
   }

   static RawFrame convert(const nfcb200_frame &f)
   {
      RawFrame frame(f.tech_type, f.frame_type);
      frame.setFramePhase(f.frame_phase);
      frame.setFrameFlags(f.frame_flags);
      frame.setFrameRate(f.frame_rate);
      frame.setSampleStart(f.sample_start);
      frame.setSampleEnd(f.sample_end);
      frame.setSampleRate(f.sample_rate);
      frame.setTimeStart(f.time_start);
      frame.setTimeEnd(f.time_end);
      frame.setDateTime(f.date_time);
      frame.put(f.data, f.length).flip();
      return frame;
   }

   std::list<RawFrame> nextFrames(hw::SignalBuffer &samples)
   {
      std::list<RawFrame> result;

      if (!ensure())
         return result; // no device: an empty list, the reason is in status()

      if (dirty)
      {
         // a setter ran since the last initialize: the thresholds reach the device before the next samples
         note(nfcb200_configure(handle, &cfg));
         dirty = false;
      }

      uint64_t count = 0;
      int rc;

      if (samples.isValid())
      {
         // re-configure decoder parameters on sample rate changes (NfcDecoder.cpp:383-388)
         if (sampleRate != (long) samples.sampleRate())
         {
            sampleRate = samples.sampleRate();
            initialize();
         }

         int sigtype;

         if (samples.type() == hw::SignalType::SIGNAL_TYPE_RADIO_SAMPLES)
            sigtype = NFCB200_SIG_MAG_F32;
         else if (samples.type() == hw::SignalType::SIGNAL_TYPE_RADIO_IQ)
            sigtype = NFCB200_SIG_IQ_F32;
         else
            return result; // not a radio buffer: nothing to decode (the reference would not consume it either)

         const unsigned int stride = sigtype == NFCB200_SIG_IQ_F32 ? 2 : 1;
         const uint64_t n = samples.remaining() / stride;

         if (n == 0)
            return result;

         const float *data = samples.data() + samples.position();

         rc = nfcb200_stream_push(handle, data, sigtype, n, (uint32_t) sampleRate, frames.data(), frames.size(), &count);
      }
      else
      {
         rc = nfcb200_stream_push(handle, nullptr, NFCB200_SIG_MAG_F32, 0, (uint32_t) sampleRate, frames.data(), frames.size(), &count);
      }

      for (uint64_t i = 0; i < count; i++)
         result.push_back(convert(frames[i]));

      // more frames than the buffer holds: drain the rest (nothing is dropped, nothing is thrown)
      while (rc == NFCB200_ERR_CAPACITY)
      {
         uint64_t left = 0;
         if (nfcb200_stream_pending(handle, frames.data(), frames.size(), &count, &left) != 0)
            break;
         for (uint64_t i = 0; i < count; i++)
            result.push_back(convert(frames[i]));
         if (left == 0)
            rc = 0;
      }

      note(rc);

      return result;
   }

   void setEnabled(unsigned bit, bool on)
   {
      if (on)
         cfg.enabled |= bit;
      else
         cfg.enabled &= ~bit;
      dirty = true;
   }

   void setThreshold(int tech, float corr, float mn, float mx)
   {
      // NaN leaves a value unchanged (NfcA.cpp:2027-2045)
      if (!std::isnan(corr))
         cfg.correlation_threshold[tech] = corr;
      if (!std::isnan(mn))
         cfg.modulation_min[tech] = mn;
      if (!std::isnan(mx))
         cfg.modulation_max[tech] = mx;
      dirty = true;
   }
};

NfcDecoder::NfcDecoder() : impl(std::make_shared<Impl>())
{
}

void NfcDecoder::initialize()
{
   impl->initialize();
}

void NfcDecoder::cleanup()
{
}

std::list<RawFrame> NfcDecoder::nextFrames(hw::SignalBuffer samples)
{
   return impl->nextFrames(samples);
}

bool NfcDecoder::isDebugEnabled() const { return impl->debugEnabled; }
void NfcDecoder::setEnableDebug(bool enabled) { impl->debugEnabled = enabled; }

bool NfcDecoder::isNfcAEnabled() const { return impl->cfg.enabled & 1; }
void NfcDecoder::setEnableNfcA(bool enabled) { impl->setEnabled(1, enabled); }
bool NfcDecoder::isNfcBEnabled() const { return impl->cfg.enabled & 2; }
void NfcDecoder::setEnableNfcB(bool enabled) { impl->setEnabled(2, enabled); }
bool NfcDecoder::isNfcFEnabled() const { return impl->cfg.enabled & 4; }
void NfcDecoder::setEnableNfcF(bool enabled) { impl->setEnabled(4, enabled); }
bool NfcDecoder::isNfcVEnabled() const { return impl->cfg.enabled & 8; }
void NfcDecoder::setEnableNfcV(bool enabled) { impl->setEnabled(8, enabled); }

long NfcDecoder::sampleRate() const { return impl->sampleRate; }
void NfcDecoder::setSampleRate(long sampleRate) { impl->sampleRate = sampleRate; }

long NfcDecoder::streamTime() const { return impl->cfg.stream_time; }
void NfcDecoder::setStreamTime(long referenceTime) { impl->cfg.stream_time = (uint32_t) referenceTime; impl->dirty = true; }

float NfcDecoder::powerLevelThreshold() const { return impl->cfg.power_level_threshold; }
void NfcDecoder::setPowerLevelThreshold(float value) { impl->cfg.power_level_threshold = value; impl->dirty = true; }

float NfcDecoder::modulationThresholdNfcAMin() const { return impl->cfg.modulation_min[0]; }
float NfcDecoder::modulationThresholdNfcAMax() const { return impl->cfg.modulation_max[0]; }
void NfcDecoder::setModulationThresholdNfcA(float min, float max) { impl->setThreshold(0, NAN, min, max); }
float NfcDecoder::modulationThresholdNfcBMin() const { return impl->cfg.modulation_min[1]; }
float NfcDecoder::modulationThresholdNfcBMax() const { return impl->cfg.modulation_max[1]; }
void NfcDecoder::setModulationThresholdNfcB(float min, float max) { impl->setThreshold(1, NAN, min, max); }
float NfcDecoder::modulationThresholdNfcFMin() const { return impl->cfg.modulation_min[2]; }
float NfcDecoder::modulationThresholdNfcFMax() const { return impl->cfg.modulation_max[2]; }
void NfcDecoder::setModulationThresholdNfcF(float min, float max) { impl->setThreshold(2, NAN, min, max); }
float NfcDecoder::modulationThresholdNfcVMin() const { return impl->cfg.modulation_min[3]; }
float NfcDecoder::modulationThresholdNfcVMax() const { return impl->cfg.modulation_max[3]; }
void NfcDecoder::setModulationThresholdNfcV(float min, float max) { impl->setThreshold(3, NAN, min, max); }

float NfcDecoder::correlationThresholdNfcA() const { return impl->cfg.correlation_threshold[0]; }
void NfcDecoder::setCorrelationThresholdNfcA(float value) { impl->setThreshold(0, value, NAN, NAN); }
float NfcDecoder::correlationThresholdNfcB() const { return impl->cfg.correlation_threshold[1]; }
void NfcDecoder::setCorrelationThresholdNfcB(float value) { impl->setThreshold(1, value, NAN, NAN); }
float NfcDecoder::correlationThresholdNfcF() const { return impl->cfg.correlation_threshold[2]; }
void NfcDecoder::setCorrelationThresholdNfcF(float value) { impl->setThreshold(2, value, NAN, NAN); }
float NfcDecoder::correlationThresholdNfcV() const { return impl->cfg.correlation_threshold[3]; }
void NfcDecoder::setCorrelationThresholdNfcV(float value) { impl->setThreshold(3, value, NAN, NAN); }

}
