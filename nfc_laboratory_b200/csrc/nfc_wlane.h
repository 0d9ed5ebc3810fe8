/*
 * nfc_wlane.h -- the warp lane: one WARP decodes one lane (a run of consecutive capture segments of one stream) from
 * tiles in shared memory (host + device; tests/native/host_sim.cpp runs the same code with a one-lane "warp").
 *
 * Round 1 ran the per-sample machine of nfc_core.h one THREAD per segment with 26 kB of delay lines per thread in global
 * memory: every tap was a DRAM round trip (1.6 % of the HBM roofline).  Here the work is split by what is sequential in
 * the reference's arithmetic and what is not:
 *
 *   front pass   (front_pass, one THREAD per segment, registers only)   the four float recurrences of nextSample
 *                (NfcTech.cpp:39-68: gated envelope, DC-removal IIR, deviation and mean EMAs) are sequential by
 *                construction -- every step rounds -- but touch no history: they run thread-per-segment at full SIMT
 *                width and leave w / envelope / deviation / mean per sample in a FEATURE pool in HBM.
 *   warp lane    (WLane, one WARP per lane, all history in SHARED memory)  reads the features 32 samples at a time
 *                (coalesced), and advances the decoder:
 *                  - search mode with every detector idle (85 % of the reference's CPU time): the six
 *                    running sums (NfcA.cpp:246-250, NfcF.cpp:240-244, NfcV.cpp:258-270) advance as six sequential float
 *                    chains on six threads -- the reference's own recurrence, add then subtract, same rounding -- and
 *                    the 32 threads evaluate every detector's trigger condition for 32 samples at once (ff_search);
 *                    only a sample on which some detector would leave its idle state goes to
 *                  - the exact per-sample machine (nfc_core.h Machine::step) on one thread, rings in shared memory.
 *   exact carry  a lane that skips the idle stretch between two segments keeps its running sums EXACT: when no add
 *                in between can round and the detectors run throughout (decided from the screening
 *                pass's block means, 16-bit mono input) the sum moves by the exact change of its window sum; on float
 *                input the lane walks the stretch with the sums alone (WALK), the reference's own add / subtract.  A stream decoded by ONE lane therefore carries the reference's
 *                float state across the whole capture: no cold start, no speculation, bit-exact on float input.
 *
 * Nothing here has a counterpart in the reference (one CPU thread per stream, one sample at a time).
 */
#ifndef NFCB200_WLANE_H
#define NFCB200_WLANE_H

#include "nfc_chain.h"

namespace nfcb200 {

struct Feat
{
   float w, env, dev, avg; // filteredValue, signalEnvelope, meanDeviation, signalAverage after the sample (NfcTech.cpp:53-68)
};

// one segment with its feature range [first, end)
struct SegRec
{
   u32 stream;
   u32 first;  // first sample of the front pass (its warm-up starts here)
   u32 begin;  // own region [begin, end), block aligned
   u32 end;
   unsigned long long featOff; // index of the feature of sample `first` in the pool
   float tEnv, tAvg, tDev, tF1; // front-end state after sample end - 1: a lane that runs past the range continues from it
   u32 tPulse;
   u32 band;   // the warm-up was the long one (carrier average exact)
   u32 hasFeat; // the front pass of this segment ran: pool[featOff ..] and the t* fields are valid.  Without it a warp lane
                // runs the front-end recurrences itself (fill_front): the straggler pass of throughput mode
};

#define NFCB200_PREROLL 400u /* samples of sums + rings before the detectors of a skipped-to segment open (> longest period) */

// gated envelope, DC-removal IIR, deviation and mean EMAs of NfcDecoderStatus::nextSample (NfcTech.cpp:39-68) over
// [first, end) from a zero state: expression for expression what Machine::front() computes (nfc_core.h)
template <class LOADX, class STORE>
NFC_HD void front_pass(const Params &P, u32 first, u32 end, LOADX loadx, STORE store, SegRec &S)
{
   float env = 0, avg = 0, dev = 0, f1 = 0;
   u32 pulse = 0;
   const u32 etu = (u32) P.etu, hold = (u32) (P.etu * 10);

   // the samples are fetched eight at a time, ahead of the recurrences that consume them (a thread walks its own stream:
   // every load is a cache miss, and the recurrences cannot start before their sample arrives)
   for (u32 pos0 = first; pos0 < end; pos0 += 8)
   {
      float xs[8];
      const u32 cnt = end - pos0 < 8 ? end - pos0 : 8;
      for (u32 i = 0; i < 8; i++)
         xs[i] = i < cnt ? loadx(pos0 + i) : 0.0f;

      for (u32 i = 0; i < 8; i++)
      {
         if (i >= cnt)
            break;

         const u32 pos = pos0 + i;
         const float x = xs[i];

         ++pulse;

         const float adiff = fabsf(x - env);
         const bool open = gate_open(adiff, env);

         if (open || pulse > hold)
         {
            pulse = 0;
            env = env * P.envW0 + x * P.envW1;
         }
         else if (pos - first < etu)
         {
            env = x;
         }

         float n0 = x + f1 * P.iirA;
         float w = n0 - f1;
         f1 = n0;

         dev = dev * P.mdevW0 + fabsf(w) * P.mdevW1;
         avg = avg * P.meanW0 + x * P.meanW1;

         store(pos - first, w, env, dev, avg);
      }
   }

   S.tEnv = env;
   S.tAvg = avg;
   S.tDev = dev;
   S.tF1 = f1;
   S.tPulse = pulse;
}

// a "warp" of one thread: the host build runs every strided loop to completion and every collective is the identity
struct HostWarp
{
   static NFC_HD u32 lane() { return 0; }
   static NFC_HD u32 width() { return 1; }
   static NFC_HD void sync() {}
   static NFC_HD u32 min_u32(u32 v) { return v; }
   static NFC_HD u32 max_u32(u32 v) { return v; }
   static NFC_HD u32 add_u32(u32 v) { return v; }
   static NFC_HD float add_f32(float v) { return v; }
   static NFC_HD u32 or_u32(u32 v) { return v; }
   static NFC_HD unsigned long long clock() { return 0; }

   // the detectors' running sums over the cnt samples after the current step: the reference's own recurrence
   template <class WL>
   static NFC_HD void sum_chains(WL &wl, u32 cnt)
   {
      wl.sum_chains_seq(cnt);
   }

   // an idle stretch [pos, target) with the sums alone
   template <class WL>
   static NFC_HD void walk(WL &wl, u32 pos, u32 target)
   {
      wl.walk_generic(pos, target);
   }
};

// per-warp scratch next to the rings (shared memory on the device)
struct WShared
{
   float lin[6][40]; // running sums of the current fast-forward span, entry a = after a samples (a = 1 .. 32)
   float cavg[32];   // carrier average of the chunk's samples
   u32 act;          // control: what the warp does next (WLANE_*)
   u32 pos, n, mode, si, j, stepped, blockActive;
   float scalEnv, scalAvg, scalDev, scalF1; // front-end state of a lane past its feature range (fill_front)
   u32 scalPulse;
   u32 scalAge;      // samples since a cold start of that state, saturating (the envelope follows x over the first etu)
   u32 nextB, nextSeg, nextCls, nextFrom; // cached look-ahead of the current inactive run: next active block, its segment, gap class
   u32 jumpCls, jumpTa, jumpGs, jumpT, jumpG, jumpB, jumpSeg;
   float delta[6];
   unsigned long long cyc[8];   // cycles per phase (WPH_*), samples per phase: development counters, summed into the handle
   unsigned long long cnt[8];
   u32 barPhase;                // device walk: parity bits of the staging barriers
   unsigned long long bar[4];   // device walk: mbarriers of the staging tiles
};

enum { WLANE_DONE = 0, WLANE_CHUNK = 1, WLANE_JUMP = 2, WLANE_WALK = 3 };
enum { WMODE_FEAT = 0, WMODE_SCAL = 1 };
enum { WPH_CONTROL = 0, WPH_FILL = 1, WPH_SEARCH = 2, WPH_MACHINE = 3, WPH_WALK = 4, WPH_JUMP = 5, WPH_SCALAR = 6, WPH_LOCKED = 7 };

// the six running-sum chains: NFC-A 106 / 212 / 424, NFC-F 212 / 424, NFC-V
struct SumChain
{
   u32 corr, p1, p2, sdd, fi, tech;
};

NFC_HD SumChain sum_chain(const Params &P, u32 d)
{
   SumChain c;
   const RateParams &b = d < 3 ? P.A[d] : (d < 5 ? P.F[d - 2] : P.V);
   c.corr = b.corr;
   c.p1 = b.p1;
   c.p2 = b.p2;
   c.sdd = b.sdd;
   c.fi = d < 3 ? d : (d < 5 ? d + 2 : 7); // index of the Mod inside Carry (nfc_core.h FI())
   c.tech = d < 3 ? EN_A : (d < 5 ? EN_F : EN_V);
   return c;
}

// index of the highest set bit (v != 0)
NFC_HD u32 high_bit(u32 v)
{
#if defined(__CUDA_ARCH__)
   return 31u - (u32) __clz((int) v);
#else
   u32 h = 31;
   while (!((v >> h) & 1u))
      h--;
   return h;
#endif
}

// (ph + m) mod p for ph < p without a division (m is at most a few periods)
NFC_HD u32 wrap_add(u32 ph, u32 m, u32 p)
{
   ph += m;
   while (ph >= p)
      ph -= p;
   return ph;
}

#define NFCB200_LEAD (NFCB200_HALO_SHORT) /* a lane re-enters the features this many samples before the next active block */

/*
 * SRC (per lane):  float x(u32 pos)            magnitude of sample pos (reference operation order)
 *                  Feat feat(u64 index)        feature pool entry
 *                  bool active(u32 pos)        the block of pos is active
 *                  float bmean(u32 b)          mean sample of block b (screening pass)
 *                  const SegRec &seg(u32 i), u32 nseg()      the segment table (all streams, ordered by stream, time)
 *                  bool exact_int()            samples are small integers over a power of two (16-bit mono): adds never round
 */
template <class W, class SINK, class SRC>
struct WLane
{
   typedef Machine<1, SINK, 0, false> Mach;

   const Params &P;
   Lane &L;
   Front &F;
   float *rg;
   WShared &sh;
   const SRC &src;
   Mach M;
   bool noff; // self-check: every sample goes through the per-sample machine

   NFC_HD WLane(const Params &p, Lane &l, float *r, u8 *s, SINK &k, WShared &w, const SRC &sr)
      : P(p), L(l), F(l.fe), rg(r), sh(w), src(sr), M(p, l, l.fe, r, s, k), noff(false)
   {
   }

   NFC_HD u32 slot(u32 k, u32 delay) const
   {
      return (k + F.kbase - delay) & (NFCB200_RING - 1);
   }

   // ------------------------------------------------------------------------------------------------------------------
   // ring fill: features of the chunk [pos, pos + n) into the sample rings; chunk sample i gets local step k0 + 1 + i
   // ------------------------------------------------------------------------------------------------------------------
   NFC_HD void fill_feat(u32 pos, u32 n, u32 k0, const SegRec &S)
   {
      u32 closedCnt = 0;
      const bool quietBlock = sh.blockActive == 0;

      for (u32 i = W::lane(); i < n; i += W::width())
      {
         const u32 s = slot(k0 + 1 + i, 0);
         const float x = src.x(pos + i);
         const Feat f = src.feat(S.featOff + (pos + i - S.first));
         // gate state for the retirement heuristic (Front::closed): the envelope before the sample is the previous feature
         if (quietBlock)
         {
            const float envPrev = i ? src.feat(S.featOff + (pos + i - 1 - S.first)).env : rg[NFCB200_OFF_M + slot(k0, 0)];
            if (!gate_open(fabsf(x - envPrev), envPrev))
               closedCnt++;
         }
         rg[NFCB200_OFF_X + s] = x;
         rg[NFCB200_OFF_W + s] = f.w;
         rg[NFCB200_OFF_D + s] = f.dev;
         rg[NFCB200_OFF_M + s] = f.env;
         sh.cavg[i] = f.avg;
      }

      if (quietBlock)
         closedCnt = W::add_u32(closedCnt);

      if (W::lane() == 0)
      {
         // the lane may only retire / skip ahead after two chunks of an inactive block with the envelope gate open throughout
         if (!quietBlock || closedCnt)
            F.closed = 64;
         else
            F.closed = F.closed >= 32 ? F.closed - 32 : 0;
      }
   }

   // the same for a chunk WITHOUT features: the recurrences of nextSample (NfcTech.cpp:39-68) from the lane's own state
   NFC_HD void fill_front(u32 pos, u32 n, u32 k0)
   {
      fill_x(pos, n, k0);
      W::sync();

      if (W::lane() == 0)
      {
         float env = sh.scalEnv, avg = sh.scalAvg, dev = sh.scalDev, f1 = sh.scalF1;
         u32 pulse = sh.scalPulse, closed = F.closed, age = sh.scalAge;
         const u32 hold = (u32) (P.etu * 10), etu = (u32) P.etu;

         for (u32 i = 0; i < n; i++)
         {
            const u32 s = slot(k0 + 1 + i, 0);
            const float x = rg[NFCB200_OFF_X + s];

            ++pulse;

            const float adiff = fabsf(x - env);
            const bool open = gate_open(adiff, env);

            if (open)
               closed = closed ? closed - 1 : 0;
            else if (closed < 4096)
               closed++;

            if (open || pulse > hold)
            {
               pulse = 0;
               env = env * P.envW0 + x * P.envW1;
            }
            else if (age < etu)
            {
               env = x; // cold start (front_pass, Machine::front): the envelope follows the signal over the first etu
            }
            if (age < 0xFFFFu)
               age++;

            float n0 = x + f1 * P.iirA;
            float w = n0 - f1;
            f1 = n0;

            dev = dev * P.mdevW0 + fabsf(w) * P.mdevW1;
            avg = avg * P.meanW0 + x * P.meanW1;

            rg[NFCB200_OFF_W + s] = w;
            rg[NFCB200_OFF_D + s] = dev;
            rg[NFCB200_OFF_M + s] = env;
            sh.cavg[i] = avg;
         }

         sh.scalEnv = env;
         sh.scalAvg = avg;
         sh.scalDev = dev;
         sh.scalF1 = f1;
         sh.scalPulse = pulse;
         sh.scalAge = age;
         F.closed = closed;
      }
   }

   NFC_HD void fill_x(u32 pos, u32 n, u32 k0)
   {
      for (u32 i = W::lane(); i < n; i += W::width())
         rg[NFCB200_OFF_X + slot(k0 + 1 + i, 0)] = src.x(pos + i);
   }

   // ------------------------------------------------------------------------------------------------------------------
   // running sums of `cnt` samples following local step F.k: results in sh.lin[d][1 .. cnt]
   // ------------------------------------------------------------------------------------------------------------------
   // the reference's recurrence, one chain per thread (all of them on the one thread of the host build)
   NFC_HD void sum_chains_seq(u32 cnt)
   {
      for (u32 d = W::lane(); d < 6; d += W::width())
      {
         const SumChain c = sum_chain(P, d);
         if (!(P.enabled & c.tech))
            continue;
         float s = F.fi[c.fi];
         u32 s0 = slot(F.k + 1, c.sdd), s1 = slot(F.k + 1, c.sdd + c.p2);
         float *lin = sh.lin[d];
         lin[0] = s;
         for (u32 a = 1; a <= cnt; a++)
         {
            s += rg[NFCB200_OFF_X + s0];
            s -= rg[NFCB200_OFF_X + s1];
            lin[a] = s;
            s0 = (s0 + 1) & (NFCB200_RING - 1);
            s1 = (s1 + 1) & (NFCB200_RING - 1);
         }
      }
   }

   // value of chain d's correlation ring as seen by the sample `a` steps after F.k, `back` samples before it
   NFC_HD float corr_at(u32 d, const SumChain &c, u32 phase0, u32 a, u32 back) const
   {
      if (back < a)
         return sh.lin[d][a - back];
      // older than this span: still in the ring (the span's own values are committed afterwards)
      return rg[c.corr + wrap_add(phase0, a + c.p1 - back, c.p1)];
   }

   NFC_HD u32 phase_of(u32 d) const
   {
      return d < 3 ? F.cA[d] : (d < 5 ? F.cF[d - 3] : F.cV1);
   }

   // clocks and ring phases after m more samples
   NFC_HD void advance(u32 m)
   {
      F.k += m;
      F.clk += m;
      F.pulseFilter += m;
      for (int r = 0; r < 3; r++)
         F.cA[r] = wrap_add(F.cA[r], m % P.A[r].p1, P.A[r].p1);
      for (int r = 0; r < 2; r++)
         F.cF[r] = wrap_add(F.cF[r], m % P.F[r + 1].p1, P.F[r + 1].p1);
      F.cV1 = wrap_add(F.cV1, m % P.V.p1, P.V.p1);
      F.cV0 = wrap_add(F.cV0, m % P.V.p0, P.V.p0);
   }

   // commit m (<= 32) samples of a fast-forward span: rings, sums, clocks, ring phases
   NFC_HD void commit(u32 m, bool sums)
   {
      if (sums)
      {
         for (u32 d = 0; d < 6; d++)
         {
            const SumChain c = sum_chain(P, d);
            if (!(P.enabled & c.tech))
               continue;
            const u32 ph0 = phase_of(d);
      #pragma unroll 1
      for (u32 a = 1 + W::lane(); a <= m; a += W::width())
               rg[c.corr + wrap_add(ph0, a, c.p1)] = sh.lin[d][a];
         }
      }

      W::sync();

      if (W::lane() == 0)
      {
         if (sums)
            for (u32 d = 0; d < 6; d++)
            {
               const SumChain c = sum_chain(P, d);
               if (P.enabled & c.tech)
                  F.fi[c.fi] = sh.lin[d][m];
            }
         F.k += m;
         F.clk += m;
         F.pulseFilter += m;
         for (int r = 0; r < 3; r++)
            F.cA[r] = wrap_add(F.cA[r], m, P.A[r].p1);
         for (int r = 0; r < 2; r++)
            F.cF[r] = wrap_add(F.cF[r], m, P.F[r + 1].p1);
         F.cV1 = wrap_add(F.cV1, m, P.V.p1);
         F.cV0 = wrap_add(F.cV0, m, P.V.p0);
      }

      W::sync();
   }

   /*
    * Search mode, every detector idle (F.lock == LOCK_NONE, F.busy == 0): advance over the samples of the chunk from
    * index j on which nothing but the running sums moves.  Returns the number of samples consumed; the sample after them
    * (if inside the chunk) needs the scalar machine: a carrier event, a carrier edge, or a detector leaving its idle state
    * (the conditions are the idle fast paths of A_detect / B_detect / F_detect / V_detect, nfc_core.h).
    */
   NFC_HD u32 ff_search(u32 j, u32 n, u32 carrierOn, u32 carrierOff)
   {
      const u32 k = F.k, gate = F.gate, gateSum = F.gateSum, warm = F.warm;
      const u32 rem = n - j;

      // ---- regime of every sample (0 off, 1 sums only, 2 detectors) and the events the machine must see ---------------
      u32 regime0 = 0;
      {
         const float env = rg[NFCB200_OFF_M + slot(k + 1, 0)];
         regime0 = (k < gateSum || env < P.power) ? 0 : (k < gate ? 1 : 2);
      }

      u32 cut = rem; // samples [0, cut) of the span are event-free and share regime0
      u32 below = 0; // bit a-1: |w| < low threshold (edge peak reset)

#pragma unroll 1
      for (u32 a = 1 + W::lane(); a <= rem; a += W::width())
      {
         const u32 ka = k + a;
         const u32 s = slot(ka, 0);
         const float env = rg[NFCB200_OFF_M + s];
         const float rect = fabsf(rg[NFCB200_OFF_W + s]);
         const float avg = sh.cavg[j + a - 1];
         const u32 regime = (ka - 1 < gateSum || env < P.power) ? 0 : (ka - 1 < gate ? 1 : 2);
         bool ev = regime != regime0;
         ev |= rect > P.highThr && ka > F.edgeHold;
         if (ka > warm)
            ev |= (avg > P.highThr && !carrierOn) || (!(avg > P.highThr) && avg < P.lowThr && !carrierOff);
         if (ev && a - 1 < cut)
            cut = a - 1;
         if (rect < P.lowThr && ka > F.edgeHold)
            below |= 1u << (a - 1);
      }

      cut = W::min_u32(cut);
      below = W::or_u32(below);

      if (cut == 0)
         return 0;

      u32 m = cut;

      if (regime0 != 0)
      {
         W::sum_chains(*this, cut);
         W::sync();
      }

      if (regime0 == 2)
      {
         // ---- trigger conditions, one sample per thread --------------------------------------------------------------
         u32 first = cut + 1; // first a whose sample must go to the machine

   #pragma unroll 1
      for (u32 a = 1 + W::lane(); a <= cut; a += W::width())
         {
            const u32 ka = k + a;
            const float env = rg[NFCB200_OFF_M + slot(ka, 0)];
            bool trig = false;

            if (P.enabled & EN_A)
            {
               const float thr = env * P.thr[TECH_A].corr;
               for (u32 d = 0; d < 3; d++)
               {
                  const SumChain c = sum_chain(P, d);
                  const float c1 = sh.lin[d][a];
                  const float c2 = corr_at(d, c, F.cA[d], a, c.p1 - c.p2);
                  const float c3 = corr_at(d, c, F.cA[d], a, 1);
                  const float s0 = c1 - c2, s1 = c2 - c3;
                  trig |= !((s0 - s1) > -0.5f * thr * (float) c.p2);
               }
            }

            if (P.enabled & EN_B)
            {
               const float thr = env * P.thr[TECH_B].modMin;
               for (u32 r = 0; r < 2; r++)
               {
                  const float edge = rg[NFCB200_OFF_W + slot(ka, P.B[r].sdd)];
                  trig |= edge < -thr;
               }
            }

            if (P.enabled & EN_F)
            {
               const float thr = env * P.thr[TECH_F].corr;
               for (u32 d = 3; d < 5; d++)
               {
                  const SumChain c = sum_chain(P, d);
                  const float c1 = sh.lin[d][a];
                  const float c2 = corr_at(d, c, F.cF[d - 3], a, c.p1 - c.p2);
                  const float c3 = corr_at(d, c, F.cF[d - 3], a, 1);
                  const float q0 = c1 - c2, q1 = c2 - c3;
                  trig |= !(fabsf(q0 - q1) < 0.5f * thr * (float) c.p2);
               }
            }

            if (P.enabled & EN_V)
            {
               const SumChain c = sum_chain(P, 5);
               const float c1 = sh.lin[5][a];
               const float c2 = corr_at(5, c, F.cV1, a, c.p1 - c.p2);
               const float s0 = (c2 - c1) / (float) c.p2;
               trig |= s0 > env * P.thr[TECH_V].corr;
            }

            if (trig && a < first)
               first = a;
         }

         first = W::min_u32(first);
         m = first - 1;
      }

      if (m == 0)
         return 0;

      // edge peak: no sample of the span exceeds the high threshold; any sample below the low one clears the peak
      if (W::lane() == 0)
      {
         const u32 mask = m >= 32 ? 0xffffffffu : ((1u << m) - 1u);
         if (below & mask)
            F.edgePeak = 0;
         F.env = rg[NFCB200_OFF_M + slot(k + m, 0)];
         F.avg = sh.cavg[j + m - 1];
      }

      commit(m, regime0 != 0);
      return m;
   }


   NFC_HD static u32 float_bits(float v)
   {
      union
      {
         float f;
         u32 u;
      } c;
      c.f = v;
      return c.u;
   }

   // carrier-edge tracker (NfcTech.cpp:77-92, Machine::edge_track) over the m samples after local step k, in closed form:
   // the peak restarts at every sample below the low threshold; inside a stretch without such a sample the tracker records
   // strictly increasing values above the high threshold, so its last record is the FIRST occurrence of the stretch's maximum
   NFC_HD void edge_span(u32 k, u32 clk, u32 m)
   {
      u32 hi = 0, lo = 0;

#pragma unroll 1
      for (u32 a = 1 + W::lane(); a <= m; a += W::width())
      {
         const float rect = fabsf(rg[NFCB200_OFF_W + slot(k + a, 0)]);
         if (k + a <= F.edgeHold)
            continue;
         if (rect > P.highThr)
            hi |= 1u << (a - 1);
         else if (rect < P.lowThr)
            lo |= 1u << (a - 1);
      }

      hi = W::or_u32(hi);
      lo = W::or_u32(lo);

      if (!(hi | lo))
         return;

      if (!hi)
      {
         if (W::lane() == 0)
            F.edgePeak = 0;
         return;
      }

      // the stretch that holds the last sample above the high threshold
      const u32 h = high_bit(hi);
      const u32 loBelow = lo & ((1u << h) - 1u);
      const u32 segStart = loBelow ? high_bit(loBelow) + 1 : 0; // first bit of that stretch
      const u32 segMask = hi & ~((1u << segStart) - 1u);
      const bool firstSeg = loBelow == 0;
      const bool resetAfter = (lo >> h) > 1u; // bit h of lo is clear (the sample is above the high threshold)

      u32 best = 0;
#pragma unroll 1
      for (u32 a = 1 + W::lane(); a <= m; a += W::width())
         if ((segMask >> (a - 1)) & 1u)
         {
            const u32 v = float_bits(fabsf(rg[NFCB200_OFF_W + slot(k + a, 0)]));
            best = v > best ? v : best;
         }
      best = W::max_u32(best);

      u32 firstAt = 0xFFFFFFFFu;
#pragma unroll 1
      for (u32 a = 1 + W::lane(); a <= m; a += W::width())
         if (((segMask >> (a - 1)) & 1u) && float_bits(fabsf(rg[NFCB200_OFF_W + slot(k + a, 0)])) == best && a < firstAt)
            firstAt = a;
      firstAt = W::min_u32(firstAt);

      if (W::lane() == 0)
      {
         union
         {
            float f;
            u32 u;
         } c;
         c.u = best;
         float peak = F.edgePeak;
         if (!firstSeg || c.f > peak)
         {
            peak = c.f;
            F.edgeTime = clk + firstAt;
         }
         F.edgePeak = resetAfter ? 0.0f : peak;
      }
   }

   // ring writes, running sum and clocks of m samples of a LOCKED NFC-A decoder (rate r): only its own sum moved
   NFC_HD void commit_locked_A(u32 r, u32 m)
   {
      const RateParams &b = P.A[r];
      const u32 ph0 = F.cA[r];
#pragma unroll 1
      for (u32 a = 1 + W::lane(); a <= m; a += W::width())
         rg[b.corr + wrap_add(ph0, a, b.p1)] = sh.lin[r][a];

      W::sync();

      if (W::lane() == 0)
      {
         F.fi[r] = sh.lin[r][m];
         F.env = rg[NFCB200_OFF_M + slot(F.k + m, 0)];
         advance(m);
      }

      W::sync();
   }

   // half-symbol running sum of a locked NFC-A decoder over cnt samples (one thread: the recurrence is sequential); the
   // addends come from the x ring (poll frames) or from the w^2 * 10 integration ring (106 kbps listen frames)
   NFC_HD void chain_locked_A(u32 r, u32 cnt, u32 ringOff)
   {
      if (W::lane() == 0)
      {
         const RateParams &b = P.A[r];
         float s = F.fi[r];
         u32 s0 = slot(F.k + 1, b.sdd), s1 = slot(F.k + 1, b.sdd + b.p2);
         float *lin = sh.lin[r];
         lin[0] = s;
         for (u32 a = 1; a <= cnt; a++)
         {
            s += rg[ringOff + s0];
            s -= rg[ringOff + s1];
            lin[a] = s;
            s0 = (s0 + 1) & (NFCB200_RING - 1);
            s1 = (s1 + 1) & (NFCB200_RING - 1);
         }
      }
   }

   // w^2 * 10 of the span's samples into the integration ring (NfcA.cpp:963, 1118)
   NFC_HD void integrate_w2(u32 sdd, u32 cnt)
   {
#pragma unroll 1
      for (u32 a = 1 + W::lane(); a <= cnt; a += W::width())
      {
         const u32 s = slot(F.k + a, sdd);
         const float w = rg[NFCB200_OFF_W + s];
         rg[NFCB200_OFF_I + s] = w * w * 10;
      }
   }

   /*
    * Locked NFC-A symbol decoders that share one shape -- decodePollFrameSymbolAsk (kind 0, NfcA.cpp:812-934) and
    * decodeListenFrameSymbolAsk (kind 1, NfcA.cpp:1095-1214): integrate every sample; inside [searchStartTime,
    * searchEndTime] keep the FIRST maximum of the correlation (above the threshold for poll frames), latch the values at
    * searchSyncTime, classify at searchEndTime.  One call consumes the chunk's samples up to the end of the window; the
    * window itself is evaluated by the 32 threads at once, the symbol and frame logic behind it by Machine's own code.
    */
   NFC_HD u32 ff_symbol_A(u32 j, u32 n, int kind)
   {
      const u32 r = F.lockRate;
      const RateParams &b = P.A[r];
      Mod &m = L.c.mA[r];
      const u32 k = F.k, clk = F.clk;
      const u32 end = m.searchEndTime, start = m.searchStartTime, syncT = m.searchSyncTime;

      if (end <= clk)
         return 0;

      u32 cnt = n - j;
      if (cnt > end - clk)
         cnt = end - clk;

      if (kind == 1)
      {
         integrate_w2(b.sdd, cnt);
         W::sync();
      }

      chain_locked_A(r, cnt, kind == 0 ? NFCB200_OFF_X : NFCB200_OFF_I);
      W::sync();

      const float thr = m.searchValueThreshold, peak0 = m.correlatedPeakValue;
      const SumChain c = sum_chain(P, r);
      u32 best = 0, latchA = 0xFFFFFFFFu;
      float mySd = 0, myS0 = 0, myS1 = 0;
      bool myCand = false;

#pragma unroll 1
      for (u32 a = 1 + W::lane(); a <= cnt; a += W::width())
      {
         const u32 ca = clk + a;
         if (ca < start)
            continue;
         const float c1 = sh.lin[r][a];
         const float c2 = corr_at(r, c, F.cA[r], a, c.p1 - c.p2);
         const float c3 = corr_at(r, c, F.cA[r], a, 1);
         const float s0 = c1 - c2, s1 = c2 - c3;
         const float sd = kind == 0 ? fabsf(s0 - s1) / (float) b.p2 : fabsf(s0 - s1);
         const bool cand = kind == 0 ? sd > thr : true;
         if (cand)
         {
            const u32 v = float_bits(sd);
            best = v > best ? v : best;
         }
         if (ca == syncT)
         {
            m.searchCorrDValue = sd;
            m.searchCorr0Value = s0;
            m.searchCorr1Value = s1;
         }
         // one sample per thread on the device; the host's single thread keeps its last in-window sample only for the
         // second pass below, which recomputes instead
         mySd = sd;
         myS0 = s0;
         myS1 = s1;
         myCand = cand;
      }

      best = W::max_u32(best);
      (void) latchA;
      (void) myS0;
      (void) myS1;
      (void) mySd;
      (void) myCand;

      {
         union
         {
            float f;
            u32 u;
         } cv;
         cv.u = best;
         if (best != 0 && cv.f > peak0)
         {
            // first sample that attains the maximum
            u32 firstAt = 0xFFFFFFFFu;
      #pragma unroll 1
      for (u32 a = 1 + W::lane(); a <= cnt; a += W::width())
            {
               const u32 ca = clk + a;
               if (ca < start)
                  continue;
               const float c1 = sh.lin[r][a];
               const float c2 = corr_at(r, c, F.cA[r], a, c.p1 - c.p2);
               const float c3 = corr_at(r, c, F.cA[r], a, 1);
               const float s0 = c1 - c2, s1 = c2 - c3;
               const float sd = kind == 0 ? fabsf(s0 - s1) / (float) b.p2 : fabsf(s0 - s1);
               if (float_bits(sd) == best && (kind == 0 ? sd > thr : true) && a < firstAt)
                  firstAt = a;
            }
            firstAt = W::min_u32(firstAt);
            if (W::lane() == 0)
            {
               m.correlatedPeakValue = cv.f;
               m.correlatedPeakTime = clk + firstAt;
            }
         }
      }

      W::sync();

      edge_span(k, clk, cnt);
      commit_locked_A(r, cnt);

      if (clk + cnt == end)
      {
         if (W::lane() == 0)
         {
            if (kind == 0)
            {
               const int pattern = M.A_poll_symbol_tail();
               if (pattern > (int) Mach::A_No)
                  M.A_poll_after(pattern);
            }
            else
            {
               const int pattern = M.A_listen_symbol_ask_tail();
               if (pattern > (int) Mach::A_No)
                  M.A_listen_ask_after(pattern);
            }
            if (F.lock == LOCK_NONE)
               M.refresh_busy();
         }
         W::sync();
      }

      return cnt;
   }

   /*
    * decodeListenFrameStartAsk (NfcA.cpp:939-1090) while nothing happens: the integrator runs on every sample; before
    * guardEnd nothing else does, after it the search acts only on a correlation beyond the threshold, a deep modulation,
    * the time-out or a window end.  Consumes the samples before the first such sample.
    */
   NFC_HD u32 ff_listen_start_A(u32 j, u32 n)
   {
      const u32 r = F.lockRate;
      const RateParams &b = P.A[r];
      Mod &m = L.c.mA[r];
      const FrameSt &fs = L.c.t[TECH_A].fs;
      const u32 k = F.k, clk = F.clk;
      const u32 cnt = n - j;

      integrate_w2(b.sdd, cnt);
      W::sync();
      chain_locked_A(r, cnt, NFCB200_OFF_I);
      W::sync();

      const SumChain c = sum_chain(P, r);
      const float thr = m.searchValueThreshold, peak = m.correlatedPeakValue;
      const bool second = m.symbolStartTime != 0;
      u32 first = cnt + 1;

#pragma unroll 1
      for (u32 a = 1 + W::lane(); a <= cnt; a += W::width())
      {
         const u32 ca = clk + a;
         if (ca < fs.guardEnd)
            continue;
         bool ev = ca == fs.guardEnd || ca > fs.waitingEnd || ca == m.searchEndTime;
         {
            const u32 s = slot(k + a, 0);
            const float x = rg[NFCB200_OFF_X + s], env = rg[NFCB200_OFF_M + s];
            const float clamped = x < 0.0f ? 0.0f : (env < x ? env : x);
            ev |= (env - clamped) / env > P.thr[TECH_A].modMin;
         }
         const float s0 = sh.lin[r][a] - corr_at(r, c, F.cA[r], a, c.p1 - c.p2);
         ev |= second ? (s0 < -thr && s0 < peak) : (s0 > thr && s0 > peak);
         if (ev && a < first)
            first = a;
      }

      first = W::min_u32(first);
      const u32 mcount = first - 1;

      if (mcount == 0)
         return 0;

      edge_span(k, clk, mcount);
      commit_locked_A(r, mcount);
      return mcount;
   }

   // an idle stretch [pos, target), sums alone: no features exist there and nothing else can move (the correlation
   // rings are not maintained: the lane re-enters the features NFCB200_LEAD samples before the next active block and
   // refills them before its detectors open)
   NFC_HDN void walk_generic(u32 pos, u32 target)
   {
      while (pos < target)
      {
         u32 n = target - pos;
         if (n > 32)
            n = 32;
         fill_x(pos, n, F.k);
         W::sync();
         W::sum_chains(*this, n);
         W::sync();
         if (W::lane() == 0)
         {
            for (u32 d = 0; d < 6; d++)
            {
               const SumChain c = sum_chain(P, d);
               if (P.enabled & c.tech)
                  F.fi[c.fi] = sh.lin[d][n];
            }
            advance(n);
         }
         W::sync();
         pos += n;
      }
   }

   // ------------------------------------------------------------------------------------------------------------------
   // exact carry of the running sums over an idle stretch (ta = last sample consumed, gs = first sample of the pre-roll)
   // ------------------------------------------------------------------------------------------------------------------
   // 1: the detectors run on every sample between ta and gs AND every add is exact (16-bit mono input: the sums are small
   //    integers over 2^15) -> the sums move by the change of their window sums; 2: detectors off throughout (envelope below
   //    the power threshold: sums frozen); 3: detectors run, adds may round (float input: even on an idle carrier the sum
   //    crosses into the next binade often enough) -> the stretch is walked with the sums alone; 0: cannot tell, keep stepping.
   // The level test uses the block means of the screening pass: an inactive block holds no level shift (nfc_decode.cuh
   // segment_flags_kernel), so its mean is the envelope to within the noise.
   NFC_HDN int gap_class(u32 ta, u32 gs) const
   {
      const u32 lookback = P.V.sdd + P.V.p2 + 8;
      const u32 b0 = (ta > lookback ? ta - lookback : 0) / NFCB200_BLOCK, b1 = gs / NFCB200_BLOCK;
      float mn = 3.0e38f, mx = -3.0e38f;

      for (u32 b = b0; b <= b1; b++)
      {
         const float m = src.bmean(b);
         mn = m < mn ? m : mn;
         mx = m > mx ? m : mx;
      }

      if (mx < 0.5f * P.power)
         return 2;

      if (!(mn > 2.0f * P.power))
         return 0;

      return src.exact_int() ? 1 : 3;
   }

   // change of the six window sums between ta and gs - 1 (exact: gap_class == 1)
   NFC_HDN void gap_delta(u32 ta, u32 gs)
   {
      for (u32 d = 0; d < 6; d++)
      {
         const SumChain c = sum_chain(P, d);
         float acc = 0;
         if (P.enabled & c.tech)
            for (u32 i = W::lane(); i < c.p2; i += W::width())
               acc += src.x(gs - 1 - c.sdd - i) - src.x(ta - c.sdd - i);
         acc = W::add_f32(acc);
         if (W::lane() == 0)
            sh.delta[d] = acc;
      }
      W::sync();
   }

   // ------------------------------------------------------------------------------------------------------------------
   // control (one thread): what comes next at sh.pos
   // ------------------------------------------------------------------------------------------------------------------
   NFC_HD bool seg_of_lane(u32 i, const LaneRec &R) const
   {
      if (i >= src.nseg())
         return false;
      const SegRec &S = src.seg(i);
      return S.stream == R.stream && S.begin < R.end;
   }

   NFC_HD void enter_scalar(const SegRec &S)
   {
      // the lane ran past its feature range without settling: it continues with the per-sample front end from the state
      // the front pass left at the end of the range
      sh.scalEnv = S.tEnv;
      sh.scalAvg = S.tAvg;
      sh.scalDev = S.tDev;
      sh.scalF1 = S.tF1;
      sh.scalPulse = S.tPulse;
      sh.scalAge = 0xFFFFu;
      sh.mode = WMODE_SCAL;
   }

   // a range without features (SegRec::hasFeat == 0) entered at `at`: the front end starts from zero at the range's first
   // sample like the front pass would, and is brought up to `at` here (one thread, no ring traffic)
   NFC_HD void enter_scalar_cold(const SegRec &S, u32 at)
   {
      SegRec tmp = S;
      if (at > S.first)
      {
         front_pass(P, S.first, at, [&](u32 p) { return src.x(p); }, [](u32, float, float, float, float) {}, tmp);
         enter_scalar(tmp);
         const u32 age = at - S.first;
         sh.scalAge = age < 0xFFFFu ? age : 0xFFFFu;
      }
      else
      {
         sh.scalEnv = sh.scalAvg = sh.scalDev = sh.scalF1 = 0.0f;
         sh.scalPulse = 0;
         sh.scalAge = 0;
         sh.mode = WMODE_SCAL;
      }
   }

   NFC_HD void plain_chunk(u32 pos, u32 limit)
   {
      u32 n = 32 - (pos & 31);
      if (n > limit - pos)
         n = limit - pos;
      sh.n = n;
   }

   NFC_HDN void control(const LaneRec &R, u32 nsamples)
   {
      const u32 pos = sh.pos;

      sh.act = WLANE_CHUNK;

      if (pos >= nsamples)
      {
         sh.act = WLANE_DONE;
         return;
      }

      const bool inFeat = sh.mode == WMODE_FEAT && pos < src.seg(sh.si).end;
      const bool active = src.active(pos);
      sh.blockActive = active ? 1u : 0u;

      if (active)
         sh.nextB = 0;

      // outside the active blocks a settled lane retires, or skips ahead to NFCB200_LEAD samples before the next active block
      if ((pos & 31) == 0 && !active && F.lock == LOCK_NONE && F.busy == 0 && F.closed < 16 && M.dormant())
      {
         // look ahead once per inactive run: the next active block B, the segment that holds it, the class of the stretch
         if (sh.nextB == 0 || pos >= sh.nextB || sh.nextFrom > pos)
         {
            u32 B = 0, ni = sh.si;
            bool found = false;

            if (inFeat)
            {
               const u32 segEnd = src.seg(sh.si).end;
               for (u32 p = (pos | (NFCB200_BLOCK - 1)) + 1; p < segEnd && !found; p += NFCB200_BLOCK)
                  if (src.active(p))
                  {
                     B = p;
                     found = true;
                  }
            }
            else
            {
               if (pos >= R.end)
               {
                  sh.act = WLANE_DONE;
                  return;
               }

               while (seg_of_lane(ni, R) && src.seg(ni).end <= pos)
                  ni++;

               if (seg_of_lane(ni, R))
               {
                  const SegRec &N = src.seg(ni);
                  if (N.begin > pos)
                  {
                     B = N.begin;
                     found = true;
                  }
                  else if (N.hasFeat && pos >= N.first + NFCB200_HALO_SHORT)
                  {
                     // inside that segment's feature range: its features are exact once the front pass has converged (the
                     // same contraction a lane start relies on), and they continue the lane's own rings without a seam
                     sh.mode = WMODE_FEAT;
                     sh.si = ni;
                     plain_chunk(pos, N.end);
                     return;
                  }
               }
            }

            sh.nextB = found ? B : 0xFFFFFFFFu;
            sh.nextSeg = ni;
            sh.nextFrom = pos;
            sh.nextCls = 0;
            if (found && B >= NFCB200_LEAD + 512 && B - NFCB200_LEAD >= pos + 512)
               sh.nextCls = (u32) gap_class(pos - 1, B - 512 - NFCB200_PREROLL);
         }
         else if (!inFeat && pos >= R.end)
         {
            sh.act = WLANE_DONE;
            return;
         }

         const u32 B = sh.nextB;

         if (B != 0xFFFFFFFFu && sh.nextCls != 0 && B - NFCB200_LEAD >= pos + 512)
         {
            const u32 T = B - NFCB200_LEAD, g = B - 512, gs = g - NFCB200_PREROLL;
            sh.act = sh.nextCls == 3 ? WLANE_WALK : WLANE_JUMP;
            sh.jumpCls = sh.nextCls;
            sh.jumpTa = pos - 1;
            sh.jumpGs = gs;
            sh.jumpT = T;
            sh.jumpG = g;
            sh.jumpB = B;
            sh.jumpSeg = sh.nextSeg;
            sh.nextB = 0;
            return;
         }
      }

      if (inFeat)
      {
         plain_chunk(pos, src.seg(sh.si).end);
         return;
      }

      if (sh.mode == WMODE_FEAT)
         enter_scalar(src.seg(sh.si));

      plain_chunk(pos, nsamples);
   }

   // ------------------------------------------------------------------------------------------------------------------
   // one lane run
   // ------------------------------------------------------------------------------------------------------------------
   NFC_HD void run(const LaneRec &R, u32 seg0, u32 nsamples)
   {
      if (W::lane() == 0)
      {
         lane_begin(L, P, R.in, R.first, R.begin - R.first);
         M.reload_front();
         sh.pos = R.first;
         sh.mode = WMODE_FEAT;
         sh.si = seg0;
         if (!src.seg(seg0).hasFeat)
            enter_scalar_cold(src.seg(seg0), R.first);
         sh.stepped = 0;
         sh.nextB = 0;
      }
      W::sync();

      for (;;)
      {
         unsigned long long t0 = W::clock();
         if (W::lane() == 0)
            control(R, nsamples);
         W::sync();
         tick(WPH_CONTROL, t0, 1);

         const u32 act = sh.act;

         if (act == WLANE_DONE)
            break;

         if (act == WLANE_JUMP || act == WLANE_WALK)
         {
            const u32 T = sh.jumpT, from = sh.pos;

            if (act == WLANE_WALK)
               W::walk(*this, from, T);
            else if (sh.jumpCls == 1)
               gap_delta(sh.jumpTa, sh.jumpGs);

            W::sync();
            tick(act == WLANE_WALK ? WPH_WALK : WPH_JUMP, t0, T - from);

            if (W::lane() == 0)
            {
               if (act == WLANE_WALK)
               {
                  // the sums ran through: they go on with the features' first sample
                  F.gateSum = F.k;
               }
               else
               {
                  if (sh.jumpCls == 1)
                     for (u32 d = 0; d < 6; d++)
                     {
                        const SumChain c = sum_chain(P, d);
                        if (P.enabled & c.tech)
                           F.fi[c.fi] += sh.delta[d];
                     }
                  F.clk = T - 1;
                  F.gateSum = F.k + (sh.jumpGs - T);
               }
               F.gate = F.k + (sh.jumpG - T);
               F.warm = F.k + (sh.jumpB - T);
               F.edgeHold = F.k + 64; // the features it enters may start with the front pass's cold-start transient
               F.closed = 64;
               sh.si = sh.jumpSeg;
               sh.mode = WMODE_FEAT;
               if (!src.seg(sh.jumpSeg).hasFeat)
                  enter_scalar_cold(src.seg(sh.jumpSeg), T);
               sh.pos = T;
               sh.stepped += T - from;
            }
            W::sync();
            continue;
         }

         const u32 pos = sh.pos, n = sh.n, mode = sh.mode;
         const u32 k0 = F.k;

         if (mode == WMODE_FEAT)
         {
            fill_feat(pos, n, k0, src.seg(sh.si));
            W::sync();
            tick(WPH_FILL, t0, n);
            chunk_feat(n);
         }
         else
         {
            // past the feature range: the front-end recurrences of the chunk run here, on one thread, from the lane's own
            // state; everything behind them is the same chunk path
            fill_front(pos, n, k0);
            W::sync();
            tick(WPH_SCALAR, t0, n);
            chunk_feat(n);
         }

         if (W::lane() == 0)
         {
            sh.pos = pos + n;
            sh.stepped += n;
         }
         W::sync();
      }
   }

   // a locked decoder with a fast path is about to run (the machine hands over after every sample it had to take)
   NFC_HD bool fast_locked() const
   {
      if (F.lock != LOCK_A)
         return false;
      const FrameSt &fs = L.c.t[TECH_A].fs;
      return fs.frameType == FT_Poll || (fs.frameType == FT_Listen && F.lockRate == 0);
   }

   // development counters: cycles since t0 go to phase ph, t0 restarts
   NFC_HD void tick(u32 ph, unsigned long long &t0, u32 samples)
   {
      const unsigned long long t1 = W::clock();
      if (W::lane() == 0)
      {
         sh.cyc[ph] += t1 - t0;
         sh.cnt[ph] += samples;
      }
      t0 = t1;
   }

   // the samples of one filled chunk
   NFC_HD void chunk_feat(u32 n)
   {
      u32 j = 0;
      unsigned long long t0 = W::clock();

      while (j < n)
      {
         const bool idle = F.lock == LOCK_NONE && F.busy == 0 && !noff;
         const u32 carrierOn = L.c.carrierOn, carrierOff = L.c.carrierOff;
         bool lockedA = false;
         int aKind = 0;
         if (!noff && F.lock == LOCK_A)
         {
            const FrameSt &fs = L.c.t[TECH_A].fs;
            if (fs.frameType == FT_Poll)
            {
               lockedA = true;
               aKind = 0;
            }
            else if (fs.frameType == FT_Listen && F.lockRate == 0)
            {
               lockedA = true;
               aKind = fs.frameStart ? 1 : 2;
            }
         }
         W::sync();

         if (idle)
         {
            const u32 m = ff_search(j, n, carrierOn, carrierOff);
            j += m;
            tick(WPH_SEARCH, t0, m);
            if (j >= n)
               break;
         }
         else if (lockedA)
         {
            // locked NFC-A decoders with a fast path; they return 0 on the samples that need the machine
            u32 m = 0;
            if (aKind == 2)
               m = ff_listen_start_A(j, n);
            else
               m = ff_symbol_A(j, n, aKind);
            j += m;
            tick(WPH_LOCKED, t0, m);
            if (j >= n)
               break;
            if (m)
               continue;
         }

         // the machine, until it is idle again (or the chunk ends); after a fast-forward span at least one sample
         if (W::lane() == 0)
         {
            M.featMode = true;
            u32 i = j;
            do
            {
               M.featAvg = sh.cavg[i];
               M.step(0.0f);
               i++;
            } while (i < n && (noff || !((F.lock == LOCK_NONE && F.busy == 0) || fast_locked())));
            sh.j = i;
         }
         W::sync();
         const u32 jn = sh.j;
         tick(WPH_MACHINE, t0, jn - j);
         j = jn;
         W::sync();
      }
   }
};

}

#endif
