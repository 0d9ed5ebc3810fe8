/*
 * tests/native/host_sim.cpp -- TEST-ONLY host build of the device lane machine (csrc/nfc_core.h).
 *
 * The product has no CPU path: nfc_core.h is device code and the shipped library (libnfcb200.so) only runs it inside
 * CUDA kernels.  This file compiles the same header with g++ so that the lane logic can be exercised in the GPU-less
 * CI container (`pytest -m "not gpu"`) against the reference oracle; it is never linked into the product library and
 * nothing in nfc_laboratory_b200/ imports it.
 *
 * Build (tests/conftest.py does this): g++ -O2 -msse2 -mfpmath=sse -ffp-contract=off -shared -fPIC
 */
#include <cstring>
#include <cstdlib>
#include <vector>

#include "../../nfc_laboratory_b200/csrc/nfc_core.h"

using namespace nfcb200;

extern "C" {

struct sim_frame
{
   uint32_t tech, type, flags, phase, rate, start, end, len;
   uint8_t data[512];
};

struct sim_result
{
   uint32_t stop;     // absolute index of the first sample NOT consumed
   uint32_t dormant;  // lane state at stop
   uint32_t locked;
   uint32_t reserved;
};

}

struct Sink
{
   sim_frame *out;
   long cap;
   long count;

   void frame(const FrameOut &f, const u8 *payload)
   {
      if (count < cap)
      {
         sim_frame &o = out[count];
         o.tech = f.tech;
         o.type = f.type;
         o.flags = f.flags;
         o.phase = f.phase;
         o.rate = f.rate;
         o.start = f.start;
         o.end = f.end;
         o.len = f.len;
         std::memset(o.data, 0, sizeof(o.data));
         std::memcpy(o.data, payload, f.len);
      }
      count++;
   }
};

extern "C" {

int hostsim_carry_size(void)
{
   return (int) sizeof(Carry);
}

int hostsim_params(uint32_t sampleRate, uint32_t enabled, Params *P)
{
   std::memset(P, 0, sizeof(*P));
   params_defaults(P);
   P->enabled = enabled;
   params_init(P, sampleRate);
   return P->valid;
}

int hostsim_params_size(void)
{
   return (int) sizeof(Params);
}

/*
 * Run one lane over mag[first .. n) (mag indexed by absolute sample).  carry_in == NULL: power-on carry.
 * The lane stops at n, or -- when own_end > 0 -- at the first sample >= own_end where it is dormant.
 */
long hostsim_run(const float *mag, uint64_t n, uint32_t sampleRate, uint32_t enabled, uint32_t first, uint32_t warm, uint32_t own_end,
                 const void *carry_in, void *carry_out, sim_frame *out, long cap, sim_result *res)
{
   Params P;
   if (!hostsim_params(sampleRate, enabled, &P))
      return -1;

   std::vector<float> scratch(NFCB200_SCRATCH_FLOATS, 0.0f);
   std::vector<u8> sb(512, 0);

   Carry carry;
   if (carry_in)
      std::memcpy(&carry, carry_in, sizeof(Carry));
   else
      carry_init(carry, P);

   Lane L;
   lane_begin(L, P, carry, first, warm);

   Sink sink {out, cap, 0};
   Machine<1, Sink> M(P, L, scratch.data(), sb.data(), sink);

   uint64_t pos = first;

   for (; pos < n; pos++)
   {
      if (own_end && pos >= own_end && M.dormant())
         break;

      M.step(mag[pos]);
   }

   if (carry_out)
      std::memcpy(carry_out, &L.c, sizeof(Carry));

   if (res)
   {
      res->stop = (uint32_t) pos;
      res->dormant = M.dormant();
      res->locked = L.lock;
      res->reserved = 0;
   }

   return sink.count;
}

}
