/*
 * nfc_trace.cu -- the per-sample value tap ON THE DEVICE (debug / test entry point of libnfcb200.so).
 *
 * The reference's signal debugger (NfcTech.h:47-126, NfcDecoder::setEnableDebug) records, for every sample, the front-end
 * signals and the locked decoder's integrator / correlation values.  This unit compiles the lane machine (nfc_core.h)
 * with its NFC_TRACE taps switched on -- the product kernels in nfcb200.cu compile them away -- and runs ONE lane over a
 * capture on the GPU, one row of 8 floats per sample: [x, w, deviation, average, ch4, ch5, lock state, 0] (channels as in
 * NfcTech.cpp:98-101, NfcA.cpp:259-264, 845-851), NaN where a channel was not written.  tests/test_gpu_tap.py compares
 * the rows with oracle/ref_tap.cpp (the reference's own debugger, recorded losslessly).
 */
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>

#define NFCB200_TRACE_VALUES 1
__device__ float *g_nfcb200_trace_row = nullptr;
__host__ __device__ __forceinline__ void nfcb200_trace_value(int channel, float value)
{
#if defined(__CUDA_ARCH__)
   if (g_nfcb200_trace_row && channel >= 0 && channel < 6)
      g_nfcb200_trace_row[channel] = value;
#else
   (void) channel;
   (void) value;
#endif
}

#include "nfc_chain.h"
#include "../../include/nfcb200.h"

using namespace nfcb200;

namespace {

struct NullSink
{
   __device__ void frame(const FrameOut &, const u8 *) {}
};

__device__ __forceinline__ float trace_sample(const void *samples, int sigtype, uint64_t idx)
{
   switch (sigtype)
   {
      case NFCB200_SIG_IQ_F32:
      {
         float2 v = ((const float2 *) samples)[idx];
         return sqrtf(v.x * v.x + v.y * v.y);
      }
      case NFCB200_SIG_MAG_F32:
         return ((const float *) samples)[idx];
      case NFCB200_SIG_MAG_S16:
         return (float) ((const short *) samples)[idx] / 32768.0f;
      default:
      {
         short2 v = ((const short2 *) samples)[idx];
         float I = (float) v.x / 32768.0f, Q = (float) v.y / 32768.0f;
         return sqrtf(I * I + Q * Q);
      }
   }
}

__global__ void trace_kernel(const void *samples, int sigtype, uint64_t n, uint32_t first, uint32_t warm, float *rows, float *scratch, u8 *sbuf,
                             Lane *lane, const __grid_constant__ Params dP)
{
   if (threadIdx.x != 0 || blockIdx.x != 0)
      return;

   Lane &L = *lane;
   Carry carry;
   if (first)
      carry_speculate(carry, dP);
   else
   {
      carry_init(carry, dP);
      carry_canon(carry);
   }
   lane_begin(L, dP, carry, first, warm);

   NullSink sink;
   Machine<1, NullSink, 0, false> M(dP, L, L.fe, scratch, sbuf, sink);
   M.reload_front();

   const float nan = __int_as_float(0x7fc00000);
   for (uint64_t pos = first; pos < n; pos++)
   {
      float *row = rows + 8 * (pos - first);
      for (int c = 0; c < 8; c++)
         row[c] = nan;
      g_nfcb200_trace_row = row;
      M.step(trace_sample(samples, sigtype, pos));
      row[6] = (float) L.fe.lock;
      row[7] = 0;
   }
   g_nfcb200_trace_row = nullptr;
}

}

extern "C" int nfcb200_debug_trace(const nfcb200_config *cfg, const void *samples, int sigtype, uint64_t n, uint32_t sample_rate, uint32_t first,
                                   uint32_t warm, float *rows)
{
   if (!cfg || !samples || !rows || n == 0 || first >= n || sigtype < NFCB200_SIG_IQ_F32 || sigtype > NFCB200_SIG_IQ_S16)
      return NFCB200_ERR_INVALID;

   Params P;
   memset(&P, 0, sizeof(P));
   params_defaults(&P);
   P.enabled = cfg->enabled & 0xF;
   P.streamTime = cfg->stream_time;
   P.power = cfg->power_level_threshold;
   for (int t = 0; t < 4; t++)
   {
      P.thr[t].corr = cfg->correlation_threshold[t];
      P.thr[t].modMin = cfg->modulation_min[t];
      P.thr[t].modMax = cfg->modulation_max[t];
   }
   params_init(&P, sample_rate);
   if (!P.valid)
      return NFCB200_ERR_UNSUPPORTED;

   if (cudaSetDevice(cfg->device) != cudaSuccess)
      return NFCB200_ERR_NO_DEVICE;

   const size_t bs = sigtype == NFCB200_SIG_IQ_F32 ? 8 : sigtype == NFCB200_SIG_MAG_S16 ? 2 : 4;
   void *dS = nullptr;
   float *dRows = nullptr, *dScratch = nullptr;
   u8 *dSb = nullptr;
   Lane *dLane = nullptr;
   const size_t rowBytes = (size_t) (n - first) * 8 * sizeof(float);
   int rc = 0;

   if (cudaMalloc(&dS, n * bs) != cudaSuccess || cudaMalloc(&dRows, rowBytes) != cudaSuccess ||
       cudaMalloc(&dScratch, NFCB200_SCRATCH_FLOATS * sizeof(float)) != cudaSuccess || cudaMalloc(&dSb, 512) != cudaSuccess ||
       cudaMalloc(&dLane, sizeof(Lane)) != cudaSuccess)
      rc = NFCB200_ERR_CUDA;

   if (!rc)
   {
      cudaMemcpy(dS, samples, n * bs, cudaMemcpyHostToDevice);
      cudaMemset(dScratch, 0, NFCB200_SCRATCH_FLOATS * sizeof(float));
      cudaMemset(dSb, 0, 512);
      trace_kernel<<<1, 32>>>(dS, sigtype, n, first, warm, dRows, dScratch, dSb, dLane, P);
      if (cudaDeviceSynchronize() != cudaSuccess)
         rc = NFCB200_ERR_CUDA;
      else
         cudaMemcpy(rows, dRows, rowBytes, cudaMemcpyDeviceToHost);
   }

   cudaFree(dS);
   cudaFree(dRows);
   cudaFree(dScratch);
   cudaFree(dSb);
   cudaFree(dLane);
   return rc;
}
