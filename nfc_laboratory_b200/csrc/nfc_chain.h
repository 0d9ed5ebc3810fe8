/*
 * nfc_chain.h -- segment construction and the speculative carry chain (host + device).
 *
 * The reference decoder is one sequential state machine per capture stream (NfcDecoder.cpp:393-442).  Here a stream is
 * cut into SEGMENTS around the blocks the dense screening pass flagged; every segment is decoded by an independent lane
 * that cold-starts HALO samples early (front end and rings re-converge bit-exactly) from a SPECULATED
 * carry.  chain_walk() then walks the lanes of one stream in time order, composes the true carry, and marks the lanes
 * whose speculation was wrong; those are re-run.  At the fixed point every lane ran from exactly the carry its
 * predecessor left, i.e. the lanes together performed the reference's sequential computation.
 *
 * Nothing here has a counterpart in the reference (which has no parallelism on this path).
 */
#ifndef NFCB200_CHAIN_H
#define NFCB200_CHAIN_H

#include "nfc_core.h"

namespace nfcb200 {

#define NFCB200_BLOCK 256       /* samples per screening block                                   */
#define NFCB200_HALO 4096       /* lane warm-up, samples                                          */
#define NFCB200_PRE_BLOCKS 2    /* active margin before a flagged block                           */
#define NFCB200_POST_BLOCKS 4   /* active margin after a flagged block                            */
#define NFCB200_GAP_BLOCKS 32   /* regions closer than this form one segment (>= 2 * HALO / BLOCK) */
#define NFCB200_START_BLOCKS 8  /* the stream start is always a segment                           */

// screening flag bits (one byte per block)
enum { SCR_TRIGGER = 1, SCR_ACTIVE = 2 };

struct LaneRec
{
   u32 stream;
   u32 begin;      // own region [begin, end), samples, block aligned (end clipped to the stream length)
   u32 end;
   u32 first;      // first sample fed to the lane (begin - HALO, or 0)
   u32 stop;       // first sample NOT consumed by the last run
   u32 lockedMask; // techs locked during the last run
   u32 gen;        // number of runs so far (frames carry the generation that produced them)
   u32 dirty;      // must (re-)run
   u32 dead;       // swallowed by its predecessor
   u32 nframes;    // frames emitted by the last run
   u32 pad[2];
   Carry in;       // carry the last run started from
   Carry out;      // canonical carry the last run retired with
};

// dilate the raw trigger flags into active blocks, in place: bit SCR_ACTIVE
NFC_HD void blocks_activate(u8 *flags, u32 nb)
{
   // forward reach (POST) and backward reach (PRE) of every trigger
   int reach = 0;
   for (u32 b = 0; b < nb; b++)
   {
      if (flags[b] & SCR_TRIGGER)
         reach = NFCB200_POST_BLOCKS + 1;
      if (reach > 0 || b < NFCB200_START_BLOCKS)
         flags[b] |= SCR_ACTIVE;
      if (reach > 0)
         reach--;
   }
   reach = 0;
   for (u32 b = nb; b-- > 0;)
   {
      if (flags[b] & SCR_TRIGGER)
         reach = NFCB200_PRE_BLOCKS + 1;
      if (reach > 0)
         flags[b] |= SCR_ACTIVE;
      if (reach > 0)
         reach--;
   }
}

// emit the segments of one stream; returns the number of segments (only `cap` are stored)
NFC_HD u32 blocks_segments(const u8 *flags, u32 nb, u32 nsamples, u32 stream, LaneRec *out, u32 cap)
{
   u32 count = 0;
   u32 b = 0;

   while (b < nb)
   {
      if (!(flags[b] & SCR_ACTIVE))
      {
         b++;
         continue;
      }

      u32 last = b;
      u32 e = b;

      while (e < nb && ((flags[e] & SCR_ACTIVE) || e - last < NFCB200_GAP_BLOCKS))
      {
         if (flags[e] & SCR_ACTIVE)
            last = e;
         e++;
      }

      if (count < cap)
      {
         LaneRec &l = out[count];
         l.stream = stream;
         l.begin = b * NFCB200_BLOCK;
         l.end = (last + 1) * NFCB200_BLOCK;
         if (l.end > nsamples)
            l.end = nsamples;
         l.first = l.begin > NFCB200_HALO ? l.begin - NFCB200_HALO : 0;
         l.stop = 0;
         l.lockedMask = 0;
         l.gen = 0;
         l.dirty = 1;
         l.dead = 0;
         l.nframes = 0;
      }

      count++;
      b = last + 1;
   }

   return count;
}

// speculated carry of a lane that does not start at sample 0: power-on state with the carrier already detected
NFC_HD void carry_speculate(Carry &c, const Params &P)
{
   carry_init(c, P);
   carry_canon(c);
   c.carrierOn = 1;
}

/*
 * Walk the lanes [0, n) of ONE stream (time ordered).  Returns the number of lanes left dirty.
 *   - a lane whose predecessor was still busy less than HALO samples before its own region is swallowed: the
 *     predecessor's region is extended over it (and the predecessor re-runs if it had already retired earlier)
 *   - otherwise the lane's last run is valid iff, on every carry group it can observe, it started from the carry
 *     composed so far; groups it cannot observe (protocol state of techs it never locked) pass through
 *   - the carry after a lane is PREDICTED for lanes that must re-run: groups the last run left unchanged are assumed
 *     to pass the corrected value through (pure heuristic -- validity is only ever established by the equality test)
 */
NFC_HD u32 chain_walk(LaneRec *lanes, u32 n, const Params &P)
{
   Carry cur;
   carry_init(cur, P);
   carry_canon(cur);

   u32 ndirty = 0;
   int prev = -1;

   for (u32 j = 0; j < n; j++)
   {
      LaneRec &L = lanes[j];

      if (L.dead)
         continue;

      if (prev >= 0)
      {
         LaneRec &Q = lanes[prev];

         // Q's reach: where its last run stopped, or -- once its region was extended -- at least the new region end
         u32 reach = Q.stop > Q.end ? Q.stop : Q.end;

         if (Q.gen > 0 && (!Q.dirty || Q.stop < Q.end) && reach + NFCB200_HALO > L.begin)
         {
            L.dead = 1;
            L.dirty = 0;
            if (Q.end < L.end)
               Q.end = L.end;
            if (Q.stop < Q.end && !Q.dirty)
            {
               Q.dirty = 1;
               ndirty++;
            }
            continue;
         }
      }

      const bool ran = L.gen > 0;
      u32 touched = 0x10F | ((L.lockedMask & 0xF) << 4); // groups 0..3 and 8 always, 4 + t when tech t was locked
      bool ok = ran;

      if (ran)
      {
         for (int g = 0; g < NFCB200_GROUPS; g++)
            if ((touched >> g) & 1)
               if (!group_equal(L.in, cur, g))
               {
                  ok = false;
#ifdef NFCB200_CHAIN_DEBUG
                  printf("   lane %u [%u,%u) gen %u: group %d differs\n", j, L.begin, L.end, L.gen, g);
                  u32 *pa, *pb, wa, wb;
                  carry_group(L.in, g, pa, wa);
                  carry_group(cur, g, pb, wb);
                  for (u32 i = 0; i < wa; i++)
                     if (pa[i] != pb[i])
                        printf("        word %u: assumed %08x true %08x\n", i, pa[i], pb[i]);
#endif
               }
      }

      // carry after this lane
      Carry next = cur;

      if (ran)
      {
         // per-word prediction: a word the last run left unchanged is assumed to pass the corrected value through, any
         // other word keeps the value the run produced.  The NFC-F pulse counters (searchPulseWidth) accumulate across
         // lanes (NfcF.cpp:307 increments, the recover path :260-271 does not clear), so they are predicted by delta.
         for (int g = 0; g < NFCB200_GROUPS; g++)
         {
            if (!((touched >> g) & 1))
               continue;

            u32 *pn, *po, *pi, wn, wo, wi;
            carry_group(next, g, pn, wn);
            carry_group(L.out, g, po, wo);
            carry_group(L.in, g, pi, wi);

            for (u32 w = 0; w < wn; w++)
            {
               if (po[w] == pi[w])
                  continue;

               const u32 pulseWord = 4; // offset of Mod::searchPulseWidth
               if (g == 2 && (w % (sizeof(Mod) / 4)) == pulseWord)
                  pn[w] = pn[w] + (po[w] - pi[w]);
               else
                  pn[w] = po[w];
            }
         }
      }

      if (!ok)
      {
         if (ran)
            L.in = cur; // a lane that never ran keeps the carry it was created with (power-on or speculated)
         L.dirty = 1;
      }

      if (L.dirty)
         ndirty++;

      cur = next;
      prev = (int) j;
   }

   return ndirty;
}

}

#endif
