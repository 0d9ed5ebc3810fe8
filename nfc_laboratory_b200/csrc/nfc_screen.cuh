/*
 * nfc_screen.cuh -- K1: the dense, HBM-bound pass over every sample of every capture stream.
 *
 * Fuses, per tile of one stream, what the reference does one sample at a time on the CPU for every idle sample
 * (85 % of its run time):
 *   IQ -> magnitude                 sqrtf(I*I + Q*Q)                     RadioDeviceTask.cpp:627-637
 *   DC-removal IIR edge signal      w = x - x[-1] + 0.9 w[-1]            NfcTech.cpp:56-62      (NFC-B detector input)
 *   half-symbol moving sums         C[t] = sum x[t-p2+1 .. t]            NfcA.cpp:246-250 (A x3 rates, F x2 share them)
 *   sliding correlators             S0 - S1 = C[t] - 2 C[t-(p1-p2)] + C[t-1]   NfcA.cpp:253-255, NfcF.cpp:247-249
 *   NFC-V pulse correlator          C[t-(p1-p2)] - C[t]                  NfcV.cpp:274
 * and tests them against the detectors' trigger thresholds with a safety margin.  The moving sums are differences of
 * a mean-removed block prefix sum (warp-shuffle scan), the IIR is an affine warp scan.  The kernel writes 5 bytes per
 * 256-sample block (trigger flag + block sum) and never materialises magnitude or features in HBM: algorithmic traffic
 * is one read of the input (8 B / sample for float2 IQ).
 *
 * Input tiles are staged into shared memory by the TMA engine (cp.async.bulk, 1-D) under an mbarrier, double
 * buffered, one persistent CTA per SM slot.
 *
 * The screen is CONSERVATIVE, not exact: any sample on which a reference detector could leave its idle state lies in
 * a flagged block (margins: thresholds x 0.9 against the block envelope, +-blocks dilation in nfc_chain.h).  The exact
 * decisions are taken by the lanes (nfc_decode.cuh) on the flagged regions only.
 */
#ifndef NFCB200_SCREEN_CUH
#define NFCB200_SCREEN_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "nfc_chain.h"

namespace nfcb200 {

// sample formats accepted at the boundary (hw/SignalType.h:27-36 for 1 and 2; 3 and 4 are WAV ingest, RecordDevice.cpp:281-311)
enum { SIG_IQ_F32 = 1, SIG_MAG_F32 = 2, SIG_MAG_S16 = 3, SIG_IQ_S16 = 4 };

__host__ __device__ inline uint32_t sig_bytes(int sigtype)
{
   return sigtype == SIG_IQ_F32 ? 8 : sigtype == SIG_MAG_F32 ? 4 : sigtype == SIG_MAG_S16 ? 2 : 4;
}

// exact magnitude of one sample (reference operation order, no FMA: the unit is compiled with -fmad=false)
__device__ __forceinline__ float sample_from_raw(const void *tile, int sigtype, uint32_t i)
{
   switch (sigtype)
   {
      case SIG_IQ_F32:
      {
         float2 v = ((const float2 *) tile)[i];
         return sqrtf(v.x * v.x + v.y * v.y);
      }
      case SIG_MAG_F32:
         return ((const float *) tile)[i];
      case SIG_MAG_S16:
         return (float) ((const short *) tile)[i] / 32768.0f;
      default:
      {
         short2 v = ((const short2 *) tile)[i];
         float I = (float) v.x / 32768.0f, Q = (float) v.y / 32768.0f;
         return sqrtf(I * I + Q * Q);
      }
   }
}

// magnitude for SCREENING only: approximate reciprocal square root (2 ulp) instead of the IEEE sqrt sequence; the exact
// decoder lanes recompute the magnitude with sample_from_raw / load_sample
__device__ __forceinline__ float screen_mag(const void *tile, int sigtype, uint32_t i)
{
   switch (sigtype)
   {
      case SIG_IQ_F32:
      {
         float2 v = ((const float2 *) tile)[i];
         float p = v.x * v.x + v.y * v.y;
         return p * rsqrtf(fmaxf(p, 1e-30f));
      }
      case SIG_MAG_F32:
         return ((const float *) tile)[i];
      case SIG_MAG_S16:
         return (float) ((const short *) tile)[i] * (1.0f / 32768.0f);
      default:
      {
         short2 v = ((const short2 *) tile)[i];
         float I = (float) v.x * (1.0f / 32768.0f), Q = (float) v.y * (1.0f / 32768.0f);
         float p = I * I + Q * Q;
         return p * rsqrtf(fmaxf(p, 1e-30f));
      }
   }
}

#define SCR_THREADS 256
#define SCR_PER_THREAD 17
#define SCR_SPAN (SCR_THREADS * SCR_PER_THREAD)   /* 4352 samples staged per tile            */
#define SCR_HALO 512                               /* history before the tile's own samples   */
#define SCR_TILE (SCR_SPAN - SCR_HALO)             /* 3840 = 15 blocks of 256 own samples     */
#define SCR_TILE_BLOCKS (SCR_TILE / NFCB200_BLOCK)

struct ScreenConfig
{
   const void *samples;    // [n_streams][n_samples] of sigtype
   uint64_t n_samples;     // per stream
   uint32_t n_streams;
   int sigtype;
   uint32_t n_blocks;      // blocks per stream
   uint32_t tiles_per_stream;
   uint8_t *flags;         // [n_streams][n_blocks]
   float *bsum;            // [n_streams][n_blocks] block sums of x
   // correlator geometry (samples): A/F rates 106, 212, 424 and NFC-V
   uint32_t p1[3], p2[3];
   uint32_t vp1, vp2;
   float thrA[3];          // |C[t] - C[t-q]| > thrA[r] * envelope flags rate r (see the derivation in the kernel)
   float thrV;             // same for the NFC-V pulse correlator
   float kB;               // |w| > kB * envelope flags an NFC-B edge
   int use_tma;
};

// ---- mbarrier / TMA bulk copy (PTX; SASS: SYNCS.*, UBLKCP) ---------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
   return (uint32_t) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
   asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "NFCB200_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra NFCB200_DONE_%=;\n"
      "bra NFCB200_WAIT_%=;\n"
      "NFCB200_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                "r"(smem_u32(bar))
                : "memory");
}

// ---- the kernel --------------------------------------------------------------------------------------------------------

struct ScreenSmem
{
   // raw staging, two stages, 16-byte aligned; sized for the widest format (float2)
   unsigned char raw[2][SCR_SPAN * 8];
   float P[SCR_SPAN + 1];   // mean-removed inclusive prefix sum, P[0] = 0
   float warpAgg[8];        // cross-warp scan scratch (prefix)
   float warpA[8], warpB[8];// cross-warp scan scratch (affine)
   float lastX[8];          // last magnitude of every warp (x[n-1] of the next warp's first sample)
   float envB[SCR_TILE_BLOCKS]; // per-block envelope reference
   uint32_t blockHit[SCR_TILE_BLOCKS];
   uint64_t bar[2];
};

struct ScreenTaps
{
   int p20, q0, p21, q1, p22, q2, pv, qv;
   float t0, t1, t2, tv, tb;
};

/*
 * Trigger tests of one thread's 17 consecutive own samples, branch-free.
 * Evaluation density: 424k window every sample, 212k at chunk offsets 0, 2, .., 16, 106k at 0, 4, .., 16, NFC-V at 0, 8,
 * 16 (the gap to the next chunk's offset 0 is one sample) -- the thresholds were lowered on the host by the change
 * |C[t] - C[t - q]| can undergo between evaluations (2 xmax per sample).  Chunk-relative phases keep every warp uniform.
 * The chunk spans at most two screening blocks: their envelope-scaled thresholds are formed once and selected per sample.
 */
__device__ __forceinline__ void screen_tests(ScreenSmem &s, const ScreenTaps &tp, int first, float exclPref, float wCarry, const float (&loc)[SCR_PER_THREAD],
                                             const float (&wl)[SCR_PER_THREAD])
{
   const int blkA = (first - SCR_HALO) >> 8;
   const int blkB = (first + SCR_PER_THREAD - 1 - SCR_HALO) >> 8;
   const int split = SCR_HALO + (blkB << 8) - first; // samples i < split belong to blkA (split >= 17 when blkA == blkB)
   const float envA = s.envB[blkA], envB = s.envB[blkB];

   const float bA = tp.tb * envA, bB = tp.tb * envB;
   const float a2A = tp.t2 * envA, a2B = tp.t2 * envB;
   const float a1A = tp.t1 * envA, a1B = tp.t1 * envB;
   const float a0A = tp.t0 * envA, a0B = tp.t0 * envB;
   const float avA = tp.tv * envA, avB = tp.tv * envB;

   const float *P = s.P + first + 1; // P[i] = inclusive prefix at the chunk's i-th sample

   bool hitA = false, hitB = false;
   float decay = 1.0f; // 0.9^(i + 1): what is left of the IIR state that entered the chunk

#pragma unroll
   for (int i = 0; i < SCR_PER_THREAD; i++)
   {
      const bool inA = i < split;
      const float Pt = exclPref + loc[i];

      decay *= 0.9f;
      bool hit = fabsf(wl[i] + decay * wCarry) > (inA ? bA : bB);

      hit |= fabsf((Pt - P[i - tp.p22]) - (P[i - tp.q2] - P[i - tp.q2 - tp.p22])) > (inA ? a2A : a2B);

      if ((i & 1) == 0)
         hit |= fabsf((Pt - P[i - tp.p21]) - (P[i - tp.q1] - P[i - tp.q1 - tp.p21])) > (inA ? a1A : a1B);

      if ((i & 3) == 0)
         hit |= fabsf((Pt - P[i - tp.p20]) - (P[i - tp.q0] - P[i - tp.q0 - tp.p20])) > (inA ? a0A : a0B);

      if ((i & 7) == 0)
         hit |= fabsf((Pt - P[i - tp.pv]) - (P[i - tp.qv] - P[i - tp.qv - tp.pv])) > (inA ? avA : avB);

      hitA |= hit && inA;
      hitB |= hit && !inA;
   }

   if (hitA)
      s.blockHit[blkA] = 1; // benign race: all writers store 1
   if (hitB)
      s.blockHit[blkB] = 1;
}

// work item -> (stream, tile); staged range in samples [lo, hi) clipped to the stream, `base` = index of smem slot 0
struct TileGeom
{
   uint32_t stream, tile;
   int64_t base;  // sample index of staged slot 0 (may be negative for tile 0)
   int64_t lo, hi; // valid samples [lo, hi)
};

__device__ __forceinline__ TileGeom tile_geom(const ScreenConfig &c, uint64_t item)
{
   TileGeom g;
   g.stream = (uint32_t) (item / c.tiles_per_stream);
   g.tile = (uint32_t) (item % c.tiles_per_stream);
   g.base = (int64_t) g.tile * SCR_TILE - SCR_HALO;
   g.lo = g.base < 0 ? 0 : g.base;
   g.hi = g.base + SCR_SPAN;
   if (g.hi > (int64_t) c.n_samples)
      g.hi = (int64_t) c.n_samples;
   return g;
}

__device__ __forceinline__ void tile_issue(const ScreenConfig &c, ScreenSmem &s, int stage, uint64_t item)
{
   // one elected thread arms the barrier and launches the bulk copy of the valid part of the tile
   TileGeom g = tile_geom(c, item);
   uint32_t bs = sig_bytes(c.sigtype);
   const unsigned char *src = (const unsigned char *) c.samples + ((uint64_t) g.stream * c.n_samples + (uint64_t) g.lo) * bs;
   uint32_t bytes = (uint32_t) (g.hi - g.lo) * bs;
   uint32_t dstoff = (uint32_t) (g.lo - g.base) * bs;
   // cp.async.bulk needs 16-byte aligned addresses and sizes: the head is aligned by construction (tile and halo are
   // multiples of 8 samples, stream pitch checked on the host); a ragged tail is finished with plain loads below
   uint32_t bulk = bytes & ~15u;
   mbar_expect_tx(&s.bar[stage], bulk);
   if (bulk)
      tma_load_1d(s.raw[stage] + dstoff, src, bulk, &s.bar[stage]);
}

__global__ void __launch_bounds__(SCR_THREADS, 2) screen_kernel(ScreenConfig c, uint64_t n_items)
{
   extern __shared__ __align__(128) unsigned char smem_raw[];
   ScreenSmem &s = *reinterpret_cast<ScreenSmem *>(smem_raw);

   const int tid = threadIdx.x;
   const int lane = tid & 31;
   const int warp = tid >> 5;
   const uint32_t bs = sig_bytes(c.sigtype);

   if (c.use_tma && tid == 0)
   {
      mbar_init(&s.bar[0], 1);
      mbar_init(&s.bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   __syncthreads();

   uint64_t item = blockIdx.x;
   uint32_t phase[2] = {0, 0};
   int stage = 0;

   if (c.use_tma && tid == 0 && item < n_items)
      tile_issue(c, s, 0, item);

   for (; item < n_items; item += gridDim.x, stage ^= 1)
   {
      TileGeom g = tile_geom(c, item);

      if (c.use_tma)
      {
         // prefetch the next tile into the other stage, then wait for this one
         uint64_t next = item + gridDim.x;
         if (tid == 0 && next < n_items)
            tile_issue(c, s, stage ^ 1, next);

         mbar_wait(&s.bar[stage], phase[stage]);
         phase[stage] ^= 1;

         // ragged tail (< 16 bytes) of the last tile of a stream
         uint32_t bytes = (uint32_t) (g.hi - g.lo) * bs;
         uint32_t bulk = bytes & ~15u;
         if (tid < (int) (bytes - bulk))
         {
            const unsigned char *src = (const unsigned char *) c.samples + ((uint64_t) g.stream * c.n_samples + (uint64_t) g.lo) * bs;
            s.raw[stage][(uint32_t) (g.lo - g.base) * bs + bulk + tid] = src[bulk + tid];
         }
      }
      else
      {
         // plain coalesced 16-byte loads (debug knob; same staging layout)
         uint32_t bytes = (uint32_t) (g.hi - g.lo) * bs;
         const unsigned char *src = (const unsigned char *) c.samples + ((uint64_t) g.stream * c.n_samples + (uint64_t) g.lo) * bs;
         unsigned char *dst = s.raw[stage] + (uint32_t) (g.lo - g.base) * bs;
         uint32_t vec = bytes >> 4;
         for (uint32_t i = tid; i < vec; i += SCR_THREADS)
            ((uint4 *) dst)[i] = __ldg(((const uint4 *) src) + i);
         for (uint32_t i = (vec << 4) + tid; i < bytes; i += SCR_THREADS)
            dst[i] = src[i];
      }

      if (tid < SCR_TILE_BLOCKS)
         s.blockHit[tid] = 0;

      __syncthreads();

      // ---- per-thread chunk: magnitude, local prefix, local IIR ------------------------------------------------------
      const int validLo = (int) (g.lo - g.base); // slots below hold no data (stream start): replicate the first sample
      const int validHi = (int) (g.hi - g.base); // slots at / above hold no data (stream end): replicate the last sample
      const void *raw = s.raw[stage];
      const bool whole = validLo == 0 && validHi == SCR_SPAN;

      float xs[SCR_PER_THREAD];
      const int first = tid * SCR_PER_THREAD;

      // reference level for the mean-removed prefix: the first valid sample of the staged span
      const float mu = screen_mag(raw, c.sigtype, (uint32_t) validLo);

      if (whole)
      {
#pragma unroll
         for (int i = 0; i < SCR_PER_THREAD; i++)
            xs[i] = screen_mag(raw, c.sigtype, (uint32_t) (first + i));
      }
      else
      {
#pragma unroll
         for (int i = 0; i < SCR_PER_THREAD; i++)
         {
            int slot = first + i;
            slot = slot < validLo ? validLo : (slot >= validHi ? validHi - 1 : slot);
            xs[i] = screen_mag(raw, c.sigtype, (uint32_t) slot);
         }
      }

      // additive scan of (x - mu): thread-local inclusive prefix, then warp shuffle scan of the chunk totals
      float loc[SCR_PER_THREAD];
      float run = 0;
#pragma unroll
      for (int i = 0; i < SCR_PER_THREAD; i++)
      {
         run += xs[i] - mu;
         loc[i] = run;
      }

      float incl = run;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1)
      {
         float o = __shfl_up_sync(0xffffffffu, incl, d);
         if (lane >= d)
            incl += o;
      }
      if (lane == 31)
      {
         s.warpAgg[warp] = incl;
         s.lastX[warp] = xs[SCR_PER_THREAD - 1];
      }

      // affine scan of w[n] = 0.9 w[n-1] + (x[n] - x[n-1]); thread-local with zero carry first.  x[n-1] of the first
      // sample of a chunk comes from the previous thread (previous warp through shared memory)
      float prevx = __shfl_up_sync(0xffffffffu, xs[SCR_PER_THREAD - 1], 1);
      __syncthreads();
      if (lane == 0)
         prevx = warp ? s.lastX[warp - 1] : xs[0];

      float wl[SCR_PER_THREAD];
      float A = 1.0f, B;
      {
         float w = 0, px = prevx;
#pragma unroll
         for (int i = 0; i < SCR_PER_THREAD; i++)
         {
            w = w * 0.9f + (xs[i] - px);
            px = xs[i];
            wl[i] = w;
            A *= 0.9f;
         }
         B = w;
      }

      // warp-level inclusive scan of the affine maps (A, B): compose(earlier, later) = (Ae * Al, Al * Be + Bl)
      float sA = A, sB = B;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1)
      {
         float oA = __shfl_up_sync(0xffffffffu, sA, d);
         float oB = __shfl_up_sync(0xffffffffu, sB, d);
         if (lane >= d)
         {
            sB = sA * oB + sB;
            sA = sA * oA;
         }
      }
      if (lane == 31)
      {
         s.warpA[warp] = sA;
         s.warpB[warp] = sB;
      }
      __syncthreads();

      // cross-warp carries (8 warps: serial, tiny)
      float prefBase = 0, wCarryWarp = 0;
      for (int v = 0; v < warp; v++)
      {
         prefBase += s.warpAgg[v];
         wCarryWarp = s.warpA[v] * wCarryWarp + s.warpB[v];
      }

      const float exclPref = prefBase + (incl - run);
      const float eA = __shfl_up_sync(0xffffffffu, sA, 1);
      const float eB = __shfl_up_sync(0xffffffffu, sB, 1);
      const float wCarry = lane ? (eA * wCarryWarp + eB) : wCarryWarp; // IIR state entering this thread's chunk

#pragma unroll
      for (int i = 0; i < SCR_PER_THREAD; i++)
         s.P[first + i + 1] = exclPref + loc[i];
      if (tid == 0)
         s.P[0] = 0;

      if (tid < SCR_TILE_BLOCKS)
         s.blockHit[tid] = 0;

      __syncthreads();

      // envelope reference per block: min(mean of this block, mean of the previous block) -- in idle both equal the
      // reference's envelope EMA to within the noise; during a pause the smaller one only makes the test stricter
      if (tid < SCR_TILE_BLOCKS)
      {
         int bslot = SCR_HALO + (tid << 8);
         float meanCur = (s.P[bslot + NFCB200_BLOCK] - s.P[bslot]) * (1.0f / NFCB200_BLOCK) + mu;
         float meanPrev = (s.P[bslot] - s.P[bslot - NFCB200_BLOCK]) * (1.0f / NFCB200_BLOCK) + mu;
         float env = fminf(meanCur, meanPrev);
         s.envB[tid] = env < 0 ? 0 : env;
      }

      __syncthreads();

      // ---- correlators and trigger tests on the tile's own samples ---------------------------------------------------
      // With C[t] = P[t] - P[t - p2] (half-symbol moving sum) the reference's correlator is
      //    S0 - S1 = (C[t] - C[t - q]) - (C[t - q] - C[t - 1]) = 2 (C[t] - C[t - q]) - (x[t] - x[t - p2])
      // so  |S0 - S1| <= 2 |C[t] - C[t - q]| + xmax, and a detector needing |S0 - S1| / p2 > T env cannot trigger while
      //    |C[t] - C[t - q]| <= thr env,   thr = min(0.9 T p2, T p2 - 1.25) / 2          (xmax <= 1.25 env)
      // One difference of two moving sums (3 shared-memory taps) per rate and sample.  The long windows change slowly
      // (by at most 2 xmax per sample), so the 212k correlator is evaluated on every 2nd sample, the 106k one on every
      // 4th and the NFC-V one on every 8th, with the thresholds lowered by the possible change in between.
      {
         ScreenTaps tp;
         tp.p20 = (int) c.p2[0];
         tp.q0 = (int) (c.p1[0] - c.p2[0]);
         tp.p21 = (int) c.p2[1];
         tp.q1 = (int) (c.p1[1] - c.p2[1]);
         tp.p22 = (int) c.p2[2];
         tp.q2 = (int) (c.p1[2] - c.p2[2]);
         tp.pv = (int) c.vp2;
         tp.qv = (int) (c.vp1 - c.vp2);
         tp.t0 = c.thrA[0];
         tp.t1 = c.thrA[1];
         tp.t2 = c.thrA[2];
         tp.tv = c.thrV;
         tp.tb = c.kB;

         const int ownEnd = (int) ((int64_t) c.n_samples - g.base); // first slot past the stream
         const int lastSlot = first + SCR_PER_THREAD - 1;

         if (first >= SCR_HALO && lastSlot < ownEnd)
         {
            // whole chunk inside the tile's own samples: branch-free tests
            screen_tests(s, tp, first, exclPref, wCarry, loc, wl);
         }
         else if (lastSlot >= SCR_HALO && first < ownEnd)
         {
            // chunk straddling the halo boundary or the end of the stream: same tests, checked per sample, no decimation
            float a = 1.0f;
#pragma unroll
            for (int i = 0; i < SCR_PER_THREAD; i++)
            {
               a *= 0.9f;
               const int slot = first + i;
               if (slot < SCR_HALO || slot >= ownEnd)
                  continue;
               const int blk = (slot - SCR_HALO) >> 8;
               const float env = s.envB[blk];
               const int t = slot + 1;
               const float Pt = s.P[t];
               bool hit = fabsf(wl[i] + a * wCarry) > tp.tb * env;
               hit |= fabsf((Pt - s.P[t - tp.p20]) - (s.P[t - tp.q0] - s.P[t - tp.q0 - tp.p20])) > tp.t0 * env;
               hit |= fabsf((Pt - s.P[t - tp.p21]) - (s.P[t - tp.q1] - s.P[t - tp.q1 - tp.p21])) > tp.t1 * env;
               hit |= fabsf((Pt - s.P[t - tp.p22]) - (s.P[t - tp.q2] - s.P[t - tp.q2 - tp.p22])) > tp.t2 * env;
               hit |= fabsf((Pt - s.P[t - tp.pv]) - (s.P[t - tp.qv] - s.P[t - tp.qv - tp.pv])) > tp.tv * env;
               if (hit)
                  s.blockHit[blk] = 1;
            }
         }
      }

      __syncthreads();

      if (tid < SCR_TILE_BLOCKS)
      {
         uint32_t b = g.tile * SCR_TILE_BLOCKS + tid;
         if (b < c.n_blocks)
         {
            int bslot = SCR_HALO + (tid << 8);
            c.flags[(uint64_t) g.stream * c.n_blocks + b] = s.blockHit[tid] ? SCR_TRIGGER : 0;
            // block sum of x over the samples that exist (the replicated tail contributes nothing real: the last block
            // of a stream is always active through the trailing margin, so its sum is only used for the envelope)
            c.bsum[(uint64_t) g.stream * c.n_blocks + b] = (s.P[bslot + NFCB200_BLOCK] - s.P[bslot]) + mu * NFCB200_BLOCK;
         }
      }

      __syncthreads(); // the staging buffer of this stage is free again before the next issue targets it
   }
}

}

#endif
