/*
 * oracle/tap/hw/RecordDevice.h -- TEST INFRASTRUCTURE.  Put FIRST on the include path of the reference's six lab-radio
 * translation units only (oracle/Makefile, target tap): lab-radio/src/main/cpp/NfcTech.h:30 includes <hw/RecordDevice.h>
 * for its signal debugger (NfcSignalDebug, NfcTech.h:47-126), which writes ten float channels per sample to a WAV file
 * (16-bit PCM: useless for a 1e-5 comparison).  This header gives the debugger a recorder with the same interface that
 * keeps the raw float rows in memory instead; no reference file is edited, the other translation units keep the real
 * hw::RecordDevice (different class name, no ODR clash).
 */
#ifndef NFCB200_ORACLE_TAP_RECORDDEVICE_H
#define NFCB200_ORACLE_TAP_RECORDDEVICE_H

#include <string>
#include <vector>

#include <hw/SignalDevice.h>

namespace hw {

// rows of `channels` floats, one row per sample clock (filled by RecordDeviceTap::write, read by oracle/ref_tap.cpp)
inline std::vector<float> &tapRows()
{
   static std::vector<float> rows;
   return rows;
}

class RecordDeviceTap : public SignalDevice
{
   public:

      explicit RecordDeviceTap(const std::string &name) : name(name) {}

      bool open(Mode) override { opened = true; return true; }

      void close() override { opened = false; }

      using Device::get;

      using Device::set;

      rt::Variant get(int id, int) const override
      {
         switch (id)
         {
            case PARAM_DEVICE_NAME: return name;
            case PARAM_CHANNEL_COUNT: return channels;
            case PARAM_SAMPLE_RATE: return sampleRate;
            default: return false;
         }
      }

      bool set(int id, const rt::Variant &value, int) override
      {
         if (auto v = std::get_if<unsigned int>(&value))
         {
            if (id == PARAM_CHANNEL_COUNT) channels = *v;
            if (id == PARAM_SAMPLE_RATE) sampleRate = *v;
            return true;
         }
         if (auto v = std::get_if<int>(&value))
         {
            if (id == PARAM_CHANNEL_COUNT) channels = (unsigned int) *v;
            if (id == PARAM_SAMPLE_RATE) sampleRate = (unsigned int) *v;
            return true;
         }
         return false;
      }

      bool isOpen() const override { return opened; }

      bool isEof() const override { return false; }

      bool isReady() const override { return opened; }

      long read(SignalBuffer &) override { return 0; }

      long write(const SignalBuffer &buffer) override
      {
         SignalBuffer copy = buffer; // the copy shares the storage and has its own cursor (rt/Buffer.h)
         const unsigned int count = copy.remaining();
         std::vector<float> &rows = tapRows();
         const size_t at = rows.size();
         rows.resize(at + count);
         if (count)
            copy.get(rows.data() + at, count);
         return (long) buffer.limit();
      }

   private:

      std::string name;
      unsigned int channels = 0;
      unsigned int sampleRate = 0;
      bool opened = false;
};

}

#define RecordDevice RecordDeviceTap

#endif
