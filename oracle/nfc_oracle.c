/*
 * oracle/nfc_oracle.c -- plain-C restatement of the reference decoder path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this file; the product
 * (nfc_laboratory_b200/libnfcb200.so) never links or executes it and has no CPU path of its own.
 *
 * What it restates (all file:line citations are relative to /root/reference/src/nfc-lib/lib-lab/lab-radio/src/main):
 *   front end                NfcDecoderStatus::nextSample            cpp/NfcTech.cpp:28-105
 *   search / dispatch loop   NfcDecoder::Impl::nextFrames            cpp/NfcDecoder.cpp:374-467
 *   carrier on / off         NfcDecoder::Impl::detectCarrier         cpp/NfcDecoder.cpp:472-523
 *   parameters               NfcDecoder::Impl::initialize + NfcX::initialize   cpp/NfcDecoder.cpp:295-360, tech/Nfc{A,B,F,V}.cpp
 *   NFC-A / B / F / V        detectModulation, decodePollFrame*, decodeListenFrame*, process*, checkCrc
 *                            cpp/tech/NfcA.cpp:217-2005, NfcB.cpp:238-1283, NfcF.cpp:206-1226, NfcV.cpp:236-1205
 *   CRC                      Crc::ccitt16                            lib-lab/lab-data/src/main/cpp/Crc.cpp:96-112
 * Every function below carries the reference lines it follows.  The reference runs one blocking loop per symbol
 * decoder; this restatement is a single sequential per-sample transition function (one decoder instance per stream,
 * no parallelism, no screening, no speculation), with the reference's rings and arithmetic order kept literally.
 *
 * PINNED: tests/test_golden_oracle.py checks this restatement against all 19 golden vectors of the reference
 * (wav/test_*.wav / .json, 300 frames) and against the compiled reference (oracle/_ref/libnfcref.so) frame for frame,
 * carrier frames included.
 *
 * Known deviations (documented, not observable on the goldens): frame bytes beyond the frame length read as zero (the
 * reference reads recycled pool memory, rt/Buffer.h:656-668); the 512-byte stream buffer is bounds-checked (the
 * reference writes past it when maxFrameSize > 512, NfcA.cpp:1716).
 *
 * Build: gcc -std=c11 -O2 -msse2 -mfpmath=sse -ffp-contract=off -fPIC -shared nfc_oracle.c -o libnfcoracle.so -lm
 * (no FMA contraction: the reference's x86 build has none, CMakeLists.txt:36-40)
 */
#include <math.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "nfc_oracle.h"

typedef uint32_t u32;
typedef uint8_t u8;

enum { TECH_A = 0, TECH_B = 1, TECH_F = 2, TECH_V = 3 };
enum { EN_A = 1, EN_B = 2, EN_F = 4, EN_V = 8 };
enum { FT_CarrierOff = 0x0100, FT_CarrierOn = 0x0101, FT_Poll = 0x0102, FT_Listen = 0x0103 };
enum { TT_Any = 0x0100, TT_A = 0x0101, TT_B = 0x0102, TT_F = 0x0103, TT_V = 0x0104 };
enum { PH_Carrier = 0x0101, PH_Selection = 0x0102, PH_Application = 0x0103 };
enum { FL_Short = 0x01, FL_Encrypted = 0x02, FL_Truncated = 0x08, FL_Parity = 0x10, FL_Crc = 0x20, FL_Sync = 0x40 };
enum { LOCK_NONE = 0, LOCK_A = 1, LOCK_B = 2, LOCK_F = 3, LOCK_V = 4 };

#define NFCB200_RING 1024
#define NFCB200_OFF_X 0
#define NFCB200_OFF_W 1024
#define NFCB200_OFF_D 2048
#define NFCB200_OFF_M 3072
#define NFCB200_OFF_I 4096
#define NFCB200_OFF_CA 5120
#define NFCB200_CA_LEN 1024
#define NFCB200_OFF_CF (5120 + 3 * 1024)
#define NFCB200_CF_LEN 1024
#define NFCB200_OFF_CV (5120 + 5 * 1024)
#define NFCB200_SCRATCH_FLOATS (5120 + 6 * 1024)

/* NfcBitrateParams (NfcTech.h:168-194), the fields this path uses */
typedef struct RateParams
{
   u32 p0, p1, p2, p4, p8, sdd, sps, pre1, c1, c0, corr;
} RateParams;

typedef struct TechThresholds
{
   float corr, modMin, modMax;
} TechThresholds;

typedef struct Params
{
   u32 sampleRate, enabled, streamTime;
   int etu;
   double stu;
   float iirA, envW0, envW1, mdevW0, mdevW1, meanW0, meanW1, power, lowThr, highThr;
   RateParams A[3], B[3], F[3], V;
   TechThresholds thr[4];
   u32 A_sfgt, A_fgt, A_fwt, A_rgt, A_fwtAtqa, fwtActivation;
   u32 B_sfgt, B_fgt, B_fwt, B_rgt, B_tr0min, B_fwtAtqb, B_tr1Min, B_tr1Max, B_s1Min, B_s1Max, B_s2Min, B_s2Max, B_eofComp;
   u32 F_sfgt, F_fgt, F_fwt, F_rgt;
   u32 V_sfgt, V_fgt, V_fwt, V_rgt, V_s1Min, V_s1Max, V_s2Min, V_s2Max, V_len2, V_len8;
} Params;

static const int NFC_FDS_TABLE_[16] = {16, 24, 32, 40, 48, 64, 96, 128, 256, 512, 1024, 2048, 4096, 0, 0, 0}; /* Nfc.h:45 */
static const int NFC_XGT_TABLE_[16] = {4096, 8192, 16384, 32768, 65536, 131072, 262144, 524288, 1048576, 2097152, 4194304, 8388608, 16777216, 33554432, 67108864, 134217728}; /* Nfc.h:48-52 */

/* NfcModulationStatus (NfcTech.h:221-259) without the two rings */
typedef struct Mod
{
   u32 searchModeState, searchStartTime, searchEndTime, searchSyncTime, searchPulseWidth;
   float searchValueThreshold, searchPhaseThreshold, searchLastPhase, searchLastValue, searchSyncValue, searchCorrDValue, searchCorr0Value, searchCorr1Value;
   u32 symbolStartTime, symbolEndTime, symbolRiseTime;
   float filterIntegrate, phaseIntegrate, correlatedPeakValue, detectorPeakValue;
   u32 correlatedPeakTime, detectorPeakTime;
} Mod;

typedef struct Sym { u32 pattern, value, start, end, edge, length; } Sym;                     /* NfcSymbolStatus  NfcTech.h:264-273 */
typedef struct Bits { u32 previous, pattern, bits, skip, data, flags, parity, bytes; } Bits;  /* NfcStreamStatus  NfcTech.h:278-289 */
typedef struct FrameSt                                                                         /* NfcFrameStatus   NfcTech.h:294-315 */
{
   u32 lastCommand, frameType, symbolRate, frameStart, frameEnd, guardEnd, waitingEnd, frameGuardTime, frameWaitingTime, startUpGuardTime, requestGuardTime;
} FrameSt;
typedef struct Proto { u32 maxFrameSize, frameGuardTime, frameWaitingTime, startUpGuardTime, requestGuardTime; } Proto; /* NfcA.cpp:74-91 */
typedef struct TechSt { FrameSt fs; Proto ps; u32 chained; } TechSt;

typedef struct State
{
   Mod mA[3], mB[2], mF[2], mV;
   TechSt t[4];
   u32 carrierOn, carrierOff;
} State;

typedef struct Front                                                                           /* NfcDecoderStatus NfcTech.h:317-393 */
{
   u32 clk, k, pulseFilter;
   float env, avg, dev, f1, edgePeak;
   u32 edgeTime;
   u32 cA[3], cF[2], cV1, cV0;
} Front;

typedef struct Lane
{
   Front fe;
   State c;
   Sym sym;
   Bits st;
   u32 lock, lockRate, pulseBits, warm, gate;
} Lane;

struct nfcoracle_decoder
{
   Params P;
   Lane L;
   float rg[NFCB200_SCRATCH_FLOATS];
   u8 sb[512];
   nfcoracle_frame *out;
   long cap, nframes;
   int initialised;
   u32 enabledCfg;
};

typedef struct nfcoracle_decoder Dec;

#define P (d->P)
#define L (d->L)
#define rg (d->rg)
#define sb (d->sb)
#define RG(off, i) rg[(off) + (i)]
#define SMP(off, delay) RG(off, (L.fe.k - (delay)) & (NFCB200_RING - 1))

/* Crc::ccitt16, lab-data Crc.cpp:96-112 (table driven there, bitwise here) */
static unsigned short crc_ccitt16(const u8 *data, u32 from, u32 to, unsigned short init, bool refin)
{
   unsigned short crc = init;

   if (to == from)
      return (unsigned short) ~init;

   for (u32 i = from; i < to; i++)
   {
      if (refin)
      {
         crc ^= data[i];
         for (int k = 0; k < 8; k++)
            crc = (crc & 1) ? (unsigned short) ((crc >> 1) ^ 0x8408) : (unsigned short) (crc >> 1);
      }
      else
      {
         crc ^= (unsigned short) (data[i] << 8);
         for (int k = 0; k < 8; k++)
            crc = (crc & 0x8000) ? (unsigned short) ((crc << 1) ^ 0x1021) : (unsigned short) (crc << 1);
      }
   }

   return crc;
}

/* NfcA::Impl::checkParity, NfcA.cpp:1994-2005 */
static bool odd_parity_ok(u32 value, u32 parity)
{
   for (u32 i = 0; i < 8; i++)
      if (value & (1u << i))
         parity ^= 1;
   return parity != 0;
}

/* ---- parameter derivation: NfcDecoder.cpp:295-360 and NfcX::initialize ------------------------------------------- */
static void rate_fill(RateParams *r, double stu, int shiftBase, int rate, u32 sdd, u32 corrOff)
{
   r->p0 = (u32) (int) round(stu * (double) ((shiftBase) >> rate));
   r->p1 = (u32) (int) round(stu * (double) ((shiftBase / 2) >> rate));
   r->p2 = (u32) (int) round(stu * (double) ((shiftBase / 4) >> rate));
   r->p4 = (u32) (int) round(stu * (double) ((shiftBase / 8) >> rate));
   r->p8 = (u32) (int) round(stu * (double) ((shiftBase / 16) >> rate));
   r->sdd = sdd;
   r->pre1 = 0;
   r->c1 = r->p1 ? (NFCB200_RING - sdd) % r->p1 : 0;
   r->c0 = r->p0 ? (NFCB200_RING - sdd) % r->p0 : 0;
   r->corr = corrOff;
}

static void params_init(Params *p, u32 sampleRate)
{
   const float NFC_FC = 13.56E6f; /* Nfc.h:36, a float constant */
   p->sampleRate = sampleRate;
   p->stu = (double) sampleRate / (double) NFC_FC;   /* NfcDecoder.cpp:307 */
   p->etu = (int) (p->stu * 128);                    /* :310 */
   p->iirA = (float) 0.9;                            /* :313 */
   p->envW0 = (float) (1 - 5E5 / sampleRate);        /* :316-317 */
   p->envW1 = (float) (1 - p->envW0);
   p->mdevW0 = (float) (1 - 2E5 / sampleRate);       /* :320-321 */
   p->mdevW1 = (float) (1 - p->mdevW0);
   p->meanW0 = (float) (1 - 5E4 / sampleRate);       /* :324-325 */
   p->meanW1 = (float) (1 - p->meanW0);
   p->lowThr = p->power / 1.25f;                     /* :328-329 */
   p->highThr = p->power * 1.25f;
   double stu = p->stu;
   for (int rate = 0; rate < 3; rate++)
   {
      u32 sddA = rate > 0 ? p->A[rate - 1].sdd + p->A[rate - 1].p1 : 0; /* NfcA.cpp:167 */
      rate_fill(&p->A[rate], stu, 256, rate, sddA, NFCB200_OFF_CA + rate * NFCB200_CA_LEN);
      p->A[rate].sps = (u32) (int) roundf(NFC_FC / (float) (128 >> rate));
      u32 sddB = rate > 0 ? p->B[rate - 1].sdd + p->B[rate - 1].p1 : 0;
      rate_fill(&p->B[rate], stu, 256, rate, sddB, 0);
      p->B[rate].sps = (u32) (int) roundf(NFC_FC / (128 >> rate));
      rate_fill(&p->F[rate], stu, 256, rate, 0, rate ? NFCB200_OFF_CF + (rate - 1) * NFCB200_CF_LEN : 0);
      p->F[rate].sps = (u32) (int) roundf(NFC_FC / (float) (128 >> rate));
      p->F[rate].pre1 = (u32) (int) round(stu * (128 >> rate) * 48);  /* NfcF.cpp:156 */
   }
   rate_fill(&p->V, stu, 512, 0, 0, NFCB200_OFF_CV);                  /* NfcV.cpp:154-173 */
   p->V.sdd = p->V.p0;
   if (!p->V.p1 || !p->V.p0 || p->V.sdd >= NFCB200_RING)
      return;
   p->V.c1 = (NFCB200_RING - p->V.sdd) % p->V.p1;
   p->V.c0 = (NFCB200_RING - p->V.sdd) % p->V.p0;
   p->V.sps = (u32) (int) roundf(NFC_FC / 256);
   p->A_sfgt = (u32) (int) (stu * (256 * 16));
   p->A_fgt = (u32) (int) (stu * 1024);
   p->A_fwt = (u32) (int) (stu * (256 * 16 * 16));
   p->A_rgt = (u32) (int) (stu * 7000);
   p->A_fwtAtqa = (u32) (stu * (128 * 18));
   p->fwtActivation = (u32) (int) (stu * 71680);
   p->B_sfgt = (u32) (int) (stu * (256 * 16));
   p->B_fgt = (u32) (int) (stu * 1024);
   p->B_fwt = (u32) (int) (stu * (256 * 16 * 16));
   p->B_rgt = (u32) (int) (stu * 7000);
   p->B_tr0min = (u32) (stu * 1024);
   p->B_fwtAtqb = (u32) (stu * 7680);
   p->B_tr1Min = (u32) (int) (stu * 1024);
   p->B_tr1Max = (u32) (int) (stu * 3200);
   p->B_s1Min = (u32) (int) (stu * 1272);
   p->B_s1Max = (u32) (int) (stu * 1416);
   p->B_s2Min = (u32) (int) (stu * 248);
   p->B_s2Max = (u32) (int) (stu * 392);
   p->B_eofComp = (u32) (int) (stu * 352);            /* NfcB.cpp:622 */
   p->F_sfgt = (u32) (int) (stu * 4096);
   p->F_fgt = (u32) (int) (stu * 1024);
   p->F_fwt = (u32) (int) (stu * (256 * 16 * 16));
   p->F_rgt = (u32) (int) (stu * 7000);
   p->V_sfgt = (u32) (int) (stu * 4096);
   p->V_fgt = (u32) (int) (stu * 1024);
   p->V_fwt = (u32) (int) (stu * (256 * 16 * 16));
   p->V_rgt = (u32) (int) (stu * 7000);
   p->V_s1Min = (u32) (int) (stu * (768 - 32));
   p->V_s1Max = (u32) (int) (stu * (768 + 32));
   p->V_s2Min = (u32) (int) (stu * (256 - 32));
   p->V_s2Max = (u32) (int) (stu * (256 + 32));
   p->V_len2 = (u32) (int) round(4 * stu * 256);      /* NfcV.cpp:224 */
   p->V_len8 = (u32) (int) round(256 * stu * 256);
}

static void zero_mod(Mod *m)
{
   m->searchModeState = 0;
   m->searchStartTime = 0;
   m->searchEndTime = 0;
   m->searchSyncTime = 0;
   m->searchPulseWidth = 0;
   m->searchValueThreshold = 0;
   m->searchPhaseThreshold = 0;
   m->searchLastPhase = 0;
   m->searchLastValue = 0;
   m->searchSyncValue = 0;
   m->searchCorrDValue = 0;
   m->searchCorr0Value = 0;
   m->searchCorr1Value = 0;
   m->symbolStartTime = 0;
   m->symbolEndTime = 0;
   m->symbolRiseTime = 0;
   m->filterIntegrate = 0;
   m->phaseIntegrate = 0;
   m->correlatedPeakValue = 0;
   m->detectorPeakValue = 0;
   m->correlatedPeakTime = 0;
   m->detectorPeakTime = 0;
}

static void zero_ring(Dec *d, u32 off, u32 len)
{
   for (u32 i = 0; i < len; i++)
      RG(off, i) = 0;
}

static void clear_bits(Dec *d)
{
   L.st.previous = L.st.pattern = L.st.bits = L.st.skip = L.st.data = L.st.flags = L.st.parity = L.st.bytes = 0;
}

static void clear_sym(Dec *d)
{
   L.sym.pattern = L.sym.value = L.sym.start = L.sym.end = L.sym.edge = L.sym.length = 0;
}

static void put_byte(Dec *d, u32 value)
{
   // the reference writes streamStatus.buffer[512] unchecked (maxFrameSize may reach 4096, NfcA.cpp:1716); cap here
   if (L.st.bytes < 512)
      sb[L.st.bytes] = (u8) value;
   L.st.bytes++;
}

// the "clear modulation status for receiving card response" block shared by all techs (NfcA.cpp:491-511,
// NfcB.cpp:515-535, NfcF.cpp:483-503, NfcV.cpp:509-529): field-wise clear + memset of both rings
static void clear_for_listen(Dec *d, Mod *m, u32 corrOff, u32 corrLen)
{
   m->symbolStartTime = 0;
   m->symbolEndTime = 0;
   m->filterIntegrate = 0;
   m->phaseIntegrate = 0;
   m->searchModeState = 0;
   m->searchSyncTime = 0;
   m->searchStartTime = 0;
   m->searchEndTime = 0;
   m->searchPulseWidth = 0;
   m->searchLastValue = 0;
   m->searchLastPhase = 0;
   m->searchValueThreshold = 0;
   m->searchPhaseThreshold = 0;
   m->correlatedPeakValue = 0;
   zero_ring(d, NFCB200_OFF_I, NFCB200_RING);
   if (corrLen)
      zero_ring(d, corrOff, corrLen);
}

static void emit(Dec *d, u32 tech, u32 type, u32 flags, u32 phase, u32 rate, u32 start, u32 end, const u8 *payload, u32 len)
{
   if (d->nframes < d->cap)
   {
      nfcoracle_frame *o = &d->out[d->nframes];
      memset(o, 0, sizeof(*o));
      o->tech_type = tech;
      o->frame_type = type;
      o->frame_flags = flags;
      o->frame_phase = phase;
      o->frame_rate = rate;
      o->sample_start = start;
      o->sample_end = end;
      o->sample_rate = P.sampleRate;
      o->time_start = (double) start / (double) P.sampleRate;
      o->time_end = (double) end / (double) P.sampleRate;
      o->date_time = (double) P.streamTime + o->time_start;
      o->length = len > 512 ? 512 : len;
      memcpy(o->data, payload, o->length);
   }
   d->nframes++;
}

// ------------------------------------------------------------------------------------------------------------------
// front end: NfcDecoderStatus::nextSample, NfcTech.cpp:28-105
// ------------------------------------------------------------------------------------------------------------------
static void front(Dec *d, float x)
{
   Front *f = &(L.fe);

   ++f->clk;
   ++f->k;
   ++f->pulseFilter;

   // correlation ring phases (signalIndex % period), kept incrementally
   for (int r = 0; r < 3; r++)
      if (++f->cA[r] == P.A[r].p1)
         f->cA[r] = 0;
   for (int r = 0; r < 2; r++)
      if (++f->cF[r] == P.F[r + 1].p1)
         f->cF[r] = 0;
   if (++f->cV1 == P.V.p1)
      f->cV1 = 0;
   if (++f->cV0 == P.V.p0)
      f->cV0 = 0;

   float diff = fabsf(x - f->env) / f->env; // NfcTech.cpp:39 (inf / NaN at env == 0 compare false, as there)

   if (diff < 0.05f || f->pulseFilter > (u32) (P.etu * 10))
   {
      f->pulseFilter = 0;
      f->env = f->env * P.envW0 + x * P.envW1;
   }
   else if (f->k - 1 < (u32) P.etu) // reference: signalClock < elementaryTimeUnit (k - 1 == clock on a fresh stream)
   {
      f->env = x;
   }

   float n0 = x + f->f1 * P.iirA; // :56
   float w = n0 - f->f1;          // :59
   f->f1 = n0;                    // :62

   f->dev = f->dev * P.mdevW0 + fabsf(w) * P.mdevW1; // :65
   f->avg = f->avg * P.meanW0 + x * P.meanW1;        // :68

   float clamped = x < 0.0f ? 0.0f : (f->env < x ? f->env : x); // std::clamp(x, 0, env), :74

   SMP(NFCB200_OFF_X, 0) = x;
   SMP(NFCB200_OFF_W, 0) = w;
   SMP(NFCB200_OFF_D, 0) = f->dev;
   SMP(NFCB200_OFF_M, 0) = (f->env - clamped) / f->env;

   float rect = fabsf(w); // :77-92

   if (rect > P.highThr)
   {
      if (rect > f->edgePeak)
      {
         f->edgePeak = rect;
         f->edgeTime = f->clk;
      }
   }
   else if (rect < P.lowThr)
   {
      f->edgePeak = 0;
   }
}

// NfcDecoder::Impl::detectCarrier, NfcDecoder.cpp:472-523
static void detect_carrier(Dec *d)
{
   Front *f = &(L.fe);

   if (f->avg > P.highThr)
   {
      if (!L.c.carrierOn)
      {
         L.c.carrierOn = f->edgeTime ? f->edgeTime : f->clk;
         emit(d, TT_Any, FT_CarrierOn, 0, PH_Carrier, 0, L.c.carrierOn, L.c.carrierOn, sb, 0);
         L.c.carrierOff = 0;
         f->edgeTime = 0;
      }
   }
   else if (f->avg < P.lowThr)
   {
      if (!L.c.carrierOff)
      {
         L.c.carrierOff = f->edgeTime ? f->edgeTime : f->clk;
         emit(d, TT_Any, FT_CarrierOff, 0, PH_Carrier, 0, L.c.carrierOff, L.c.carrierOff, sb, 0);
         L.c.carrierOn = 0;
         f->edgeTime = 0;
      }
   }
}

// S0 / S1 of the half-symbol correlator over a ring of period p1 (NfcA.cpp:241-255): C is written at slot c
static void corr_points(u32 c, u32 p1, u32 p2, u32 *fp2, u32 *fp3)
{
   *fp2 = c + p2;
   if (*fp2 >= p1)
      *fp2 -= p1; /* (signalIndex + p2) % p1, valid because p2 < p1 */
   *fp3 = c ? c - 1 : p1 - 1; /* (signalIndex + p1 - 1) % p1 */
}

// ------------------------------------------------------------------------------------------------------------------
// NFC-A
// ------------------------------------------------------------------------------------------------------------------

// NfcA::Impl::resetModulation, NfcA.cpp:1451-1475
static void A_reset(Dec *d)
{
   for (int r = 0; r < 3; r++)
   {
      zero_mod(&L.c.mA[r]);
      zero_ring(d, P.A[r].corr, P.A[r].p1);
   }
   // integrationData of all rates: one shared ring, all-zero whenever a listen phase starts
   zero_ring(d, NFCB200_OFF_I, NFCB200_RING);
   clear_bits(d);
   clear_sym(d);
   L.c.t[TECH_A].fs.frameType = 0;
   L.c.t[TECH_A].fs.frameStart = 0;
   L.c.t[TECH_A].fs.frameEnd = 0;
   L.lock = LOCK_NONE;
}

// NfcA::Impl::detectModulation, NfcA.cpp:217-411 (the clock / envelope gates are applied by the caller)
static bool A_detect(Dec *d)
{
   const float env = L.fe.env;
   const u32 clk = L.fe.clk;
   const float minimumCorrelationValue = env * P.thr[TECH_A].corr;
   const float minDeep = P.thr[TECH_A].modMin;

   for (int rate = 0; rate < 3; rate++)
   {
      const RateParams *b = &(P.A[rate]);
      Mod *m = &(L.c.mA[rate]);

      u32 fp1 = L.fe.cA[rate], fp2, fp3;
      corr_points(fp1, b->p1, b->p2, &fp2, &fp3);

      // :246-250
      m->filterIntegrate += SMP(NFCB200_OFF_X, b->sdd);
      m->filterIntegrate -= SMP(NFCB200_OFF_X, b->sdd + b->p2);
      RG(b->corr, fp1) = m->filterIntegrate;

      // :253-255
      float s0 = m->filterIntegrate - RG(b->corr, fp2);
      float s1 = RG(b->corr, fp2) - RG(b->corr, fp3);
      float sd = (s0 - s1) / (float) b->p2;

      // :268-279 recover status from previous partial search
      if (m->correlatedPeakTime && clk > m->correlatedPeakTime + b->p1)
      {
         m->symbolStartTime = 0;
         m->symbolEndTime = 0;
         m->searchStartTime = 0;
         m->searchEndTime = 0;
         m->searchSyncTime = 0;
         m->detectorPeakTime = 0;
         m->detectorPeakValue = 0;
         m->correlatedPeakTime = 0;
         m->correlatedPeakValue = 0;
      }

      if (clk < m->searchStartTime) // :282
         continue;

      if (!m->symbolStartTime) // :285-306
      {
         float deep = SMP(NFCB200_OFF_M, b->sdd + b->p8);

         if (sd < -minimumCorrelationValue)
         {
            if (sd < m->correlatedPeakValue)
            {
               m->correlatedPeakValue = sd;
               m->correlatedPeakTime = clk;
               m->searchEndTime = clk + b->p4;
            }

            if (deep > m->detectorPeakValue)
            {
               m->detectorPeakValue = deep;
               m->detectorPeakTime = clk;
            }
         }
      }
      else // :307-318
      {
         if (sd > minimumCorrelationValue)
         {
            if (sd > m->correlatedPeakValue)
            {
               m->correlatedPeakValue = sd;
               m->correlatedPeakTime = clk;
            }
         }
      }

      if (clk != m->searchEndTime) // :321
         continue;

      if (!m->symbolStartTime) // :324-350
      {
         if (m->detectorPeakValue < minDeep)
         {
            m->symbolStartTime = 0;
            m->symbolEndTime = 0;
            m->searchSyncTime = 0;
            m->searchStartTime = 0;
            m->searchEndTime = 0;
            m->searchPulseWidth = 0;
            m->correlatedPeakTime = 0;
            m->correlatedPeakValue = 0;
            m->detectorPeakTime = 0;
            m->detectorPeakValue = 0;
            continue;
         }

         m->searchSyncTime = m->correlatedPeakTime + b->p2;
         m->searchStartTime = m->searchSyncTime - b->p8;
         m->searchEndTime = m->searchSyncTime + b->p8;
         m->symbolStartTime = m->correlatedPeakTime - b->p2;
         m->correlatedPeakTime = 0;
         m->correlatedPeakValue = 0;
         continue;
      }

      // :353-378
      m->symbolEndTime = m->correlatedPeakTime;
      m->searchPulseWidth = m->symbolEndTime - m->symbolStartTime;

      u32 minimumPulseWidth = b->p1 - b->p4; // compared unsigned in the reference (unsigned < int)
      u32 maximumPulseWidth = b->p1 + b->p4;

      if (m->correlatedPeakTime == 0 || m->detectorPeakValue < minDeep || m->searchPulseWidth < minimumPulseWidth || m->searchPulseWidth > maximumPulseWidth)
      {
         m->symbolStartTime = 0;
         m->symbolEndTime = 0;
         m->searchSyncTime = 0;
         m->searchStartTime = 0;
         m->searchEndTime = 0;
         m->searchPulseWidth = 0;
         m->correlatedPeakTime = 0;
         m->correlatedPeakValue = 0;
         m->detectorPeakTime = 0;
         m->detectorPeakValue = 0;
         continue;
      }

      // :381-407 lock
      m->searchSyncTime = m->symbolEndTime + b->p1;
      m->searchStartTime = m->searchSyncTime - b->p8;
      m->searchEndTime = m->searchSyncTime + b->p8;
      m->searchValueThreshold = m->correlatedPeakValue / 2;
      m->searchCorr0Value = 0;
      m->searchCorr1Value = 0;
      m->correlatedPeakTime = 0;
      m->correlatedPeakValue = 0;

      FrameSt *fs = &(L.c.t[TECH_A].fs);
      fs->frameType = FT_Poll;
      fs->symbolRate = b->sps;
      fs->frameStart = m->symbolStartTime - b->sdd;
      fs->frameEnd = 0;

      L.sym.value = 0;
      L.sym.start = m->symbolStartTime - b->sdd;
      L.sym.end = m->symbolEndTime - b->sdd;
      L.sym.length = L.sym.end - L.sym.start;
      L.sym.pattern = 4; // PatternZ

      L.lock = LOCK_A;
      L.lockRate = rate;
      return true;
   }

   return false;
}

enum { A_Invalid = 0, A_No = 1, A_X = 2, A_Y = 3, A_Z = 4, A_D = 5, A_E = 6, A_F = 7, A_M = 8, A_N = 9, A_S = 10, A_O = 11 };

// NfcA::Impl::checkCrc, NfcA.cpp:1978-1989
static bool A_crc_ok(Dec *d, u32 size)
{
   if (size < 2)
      return true;
   unsigned short crc = crc_ccitt16(sb, 0, size - 2, 0x6363, true);
   unsigned short res = (unsigned short) ((sb[size - 2] & 0xff) | ((sb[size - 1] & 0xff) << 8));
   return res == crc;
}

// frame byte access: the reference reads RawFrame storage unchecked (Buffer.h:656-668); bytes beyond the frame
// length are recycled pool memory there, zero here
static u32 fb(Dec *d, u32 i, u32 len)
{
   return i < len && i < 512 ? sb[i] : 0;
}

static void A_default_protocol(Dec *d, Proto *ps)
{
   ps->maxFrameSize = 256;
   ps->startUpGuardTime = P.A_sfgt;
   ps->frameGuardTime = P.A_fgt;
   ps->frameWaitingTime = P.A_fwt;
   ps->requestGuardTime = P.A_rgt;
}

// NfcA::Impl::process and the processXXX chain, NfcA.cpp:1480-1973
static void A_process(Dec *d, u32 type, u32 len, u32 *flags_, u32 *phase_)
{
   u32 flags = *flags_, phase = *phase_;
   TechSt *t = &(L.c.t[TECH_A]);
   FrameSt *fs = &(t->fs);
   Proto *ps = &(t->ps);
   const bool poll = type == FT_Poll;
   const u32 b0 = fb(d, 0, len);

   if (poll)
   {
      fs->startUpGuardTime = ps->startUpGuardTime;
      fs->frameWaitingTime = ps->frameWaitingTime;
      fs->frameGuardTime = ps->frameGuardTime;
      fs->requestGuardTime = ps->requestGuardTime;
   }
   else
   {
      fs->frameGuardTime = ps->frameGuardTime;
   }

   bool done = false;

   // processREQA :1592-1631
   if (poll)
   {
      if ((b0 == 0x26 || b0 == 0x52) && len == 1)
      {
         phase = PH_Selection;
         fs->lastCommand = b0;
         A_default_protocol(d, ps);
         fs->frameGuardTime = P.A_fgt;
         fs->frameWaitingTime = P.A_fwtAtqa;
         t->chained = 0;
         done = true;
      }
   }
   else if (fs->lastCommand == 0x26 || fs->lastCommand == 0x52)
   {
      phase = PH_Selection;
      done = true;
   }

   // processHLTA :1636-1665
   if (!done && poll && b0 == 0x50 && len == 4 && !(flags & FL_Crc))
   {
      phase = PH_Selection;
      flags |= !A_crc_ok(d, len) ? FL_Crc : 0;
      fs->lastCommand = b0;
      A_default_protocol(d, ps);
      t->chained = 0;
      A_reset(d);
      done = true;
   }

   if (!done)
   {
      if (!(t->chained & FL_Encrypted))
      {
         do
         {
            // processSELn :1670-1699
            if (poll)
            {
               if (b0 == 0x93 || b0 == 0x95 || b0 == 0x97)
               {
                  phase = PH_Selection;
                  fs->lastCommand = b0;
                  fs->frameGuardTime = P.A_fgt;
                  fs->frameWaitingTime = P.A_fwtAtqa;
                  break;
               }
            }
            else if (fs->lastCommand == 0x93 || fs->lastCommand == 0x95 || fs->lastCommand == 0x97)
            {
               phase = PH_Selection;
               break;
            }

            // processRATS :1704-1790
            if (poll)
            {
               if (b0 == 0xE0)
               {
                  int fsdi = (fb(d, 1, len) >> 4) & 0x0F;
                  fs->lastCommand = b0;
                  ps->maxFrameSize = (u32) NFC_FDS_TABLE_[fsdi];
                  fs->frameWaitingTime = P.fwtActivation;
                  phase = PH_Selection;
                  flags |= !A_crc_ok(d, len) ? FL_Crc : 0;
                  break;
               }
            }
            else if (fs->lastCommand == 0xE0)
            {
               u32 offset = 0;
               u32 tl = fb(d, offset++, len);

               if (tl > 0)
               {
                  u32 t0 = fb(d, offset++, len);

                  if (t0 & 0x10)
                     offset++;

                  if (t0 & 0x20)
                  {
                     u32 tb = fb(d, offset++, len);
                     u32 sfgi = tb & 0x0f;
                     u32 fwi = (tb >> 4) & 0x0f;
                     if (sfgi == 15)
                        sfgi = 0;
                     if (fwi == 15)
                        fwi = 4;
                     ps->startUpGuardTime = (u32) (int) (P.stu * NFC_XGT_TABLE_[sfgi]);
                     ps->frameWaitingTime = (u32) (int) (P.stu * NFC_XGT_TABLE_[fwi]);
                  }
                  else
                  {
                     ps->startUpGuardTime = P.A_sfgt;
                     ps->frameWaitingTime = P.A_fwt;
                  }
               }

               phase = PH_Selection;
               flags |= !A_crc_ok(d, len) ? FL_Crc : 0;
               break;
            }

            // processPPSr :1795-1822
            if (poll)
            {
               if ((b0 & 0xF0) == 0xD0)
               {
                  fs->lastCommand = b0 & 0xF0;
                  phase = PH_Selection;
                  flags |= !A_crc_ok(d, len) ? FL_Crc : 0;
                  break;
               }
            }
            else if (fs->lastCommand == 0xD0)
            {
               phase = PH_Selection;
               flags |= !A_crc_ok(d, len) ? FL_Crc : 0;
               break;
            }

            // processAUTH :1827-1868
            if (poll)
            {
               if (b0 == 0x60 || b0 == 0x61)
               {
                  fs->lastCommand = b0;
                  phase = PH_Application;
                  flags |= !A_crc_ok(d, len) ? FL_Crc : 0;
                  break;
               }
            }
            else if (fs->lastCommand == 0x60 || fs->lastCommand == 0x61)
            {
               t->chained = FL_Encrypted;
               phase = PH_Application;
               break;
            }

            // processIBlock :1873-1900
            if (poll)
            {
               if ((b0 & 0xE2) == 0x02 && len > 4)
               {
                  fs->lastCommand = b0 & 0xE2;
                  phase = PH_Application;
                  flags |= !A_crc_ok(d, len) ? FL_Crc : 0;
                  break;
               }
            }
            else if (fs->lastCommand == 0x02)
            {
               phase = PH_Application;
               flags |= !A_crc_ok(d, len) ? FL_Crc : 0;
               break;
            }

            // processRBlock :1905-1932
            if (poll)
            {
               if ((b0 & 0xE6) == 0xA2 && len == 3)
               {
                  fs->lastCommand = b0 & 0xE6;
                  phase = PH_Application;
                  flags |= !A_crc_ok(d, len) ? FL_Crc : 0;
                  break;
               }
            }
            else if (fs->lastCommand == 0xA2)
            {
               phase = PH_Application;
               flags |= !A_crc_ok(d, len) ? FL_Crc : 0;
               break;
            }

            // processSBlock :1937-1964
            if (poll)
            {
               if ((b0 & 0xC7) == 0xC0 && len == 4)
               {
                  fs->lastCommand = b0 & 0xC7;
                  phase = PH_Application;
                  flags |= !A_crc_ok(d, len) ? FL_Crc : 0;
                  break;
               }
            }
            else if (fs->lastCommand == 0xC0)
            {
               phase = PH_Application;
               flags |= !A_crc_ok(d, len) ? FL_Crc : 0;
               break;
            }

            // processOther :1969-1973
            phase = PH_Application;
            flags |= !A_crc_ok(d, len) ? FL_Crc : 0;
         }
         while (false);
      }
      else
      {
         flags &= ~(u32) FL_Parity; // :1536
         phase = PH_Application;
      }
   }

   flags |= t->chained; // :1545

   if (poll) // :1548-1577
   {
      if (L.lock == LOCK_A)
      {
         u32 sdd = P.A[L.lockRate].sdd;
         fs->guardEnd = fs->frameEnd + fs->frameGuardTime + sdd;
         fs->waitingEnd = fs->frameEnd + fs->frameWaitingTime + sdd;
         fs->frameType = FT_Listen;
      }
   }
   else
   {
      if (L.lock == LOCK_A)
         fs->guardEnd = fs->frameEnd + fs->frameGuardTime + P.A[L.lockRate].sdd;
      fs->frameType = 0;
      fs->lastCommand = 0;
   }

   fs->frameStart = 0;
   fs->frameEnd = 0;
   *flags_ = flags; *phase_ = phase;
}

// one sample of decodePollFrameSymbolAsk, NfcA.cpp:812-934.  Returns a pattern or A_Invalid (no symbol yet).
static int A_poll_symbol(Dec *d)
{
   const RateParams *b = &(P.A[L.lockRate]);
   Mod *m = &(L.c.mA[L.lockRate]);
   const u32 clk = L.fe.clk;

   u32 fp1 = L.fe.cA[L.lockRate], fp2, fp3;
   corr_points(fp1, b->p1, b->p2, &fp2, &fp3);

   m->filterIntegrate += SMP(NFCB200_OFF_X, b->sdd);
   m->filterIntegrate -= SMP(NFCB200_OFF_X, b->sdd + b->p2);
   RG(b->corr, fp1) = m->filterIntegrate;

   float s0 = m->filterIntegrate - RG(b->corr, fp2);
   float s1 = RG(b->corr, fp2) - RG(b->corr, fp3);
   float sd = fabsf(s0 - s1) / (float) b->p2;

   if (clk < m->searchStartTime)
      return A_Invalid;

   if (sd > m->correlatedPeakValue && sd > m->searchValueThreshold) // :858
   {
      m->correlatedPeakValue = sd;
      m->correlatedPeakTime = clk;
   }

   if (clk == m->searchSyncTime) // :865
   {
      m->searchCorrDValue = sd;
      m->searchCorr0Value = s0;
      m->searchCorr1Value = s1;
   }

   if (clk != m->searchEndTime)
      return A_Invalid;

   if (m->searchCorrDValue < m->searchValueThreshold) // :877 Pattern-Y
   {
      m->symbolStartTime = m->symbolEndTime;
      m->symbolEndTime = m->searchSyncTime;
      m->symbolRiseTime = m->symbolStartTime;
      L.sym.value = 1;
      L.sym.pattern = A_Y;
   }
   else if (m->searchCorr0Value > m->searchCorr1Value) // :890 Pattern-Z
   {
      m->symbolStartTime = m->symbolEndTime;
      m->symbolEndTime = m->correlatedPeakTime;
      m->symbolRiseTime = m->correlatedPeakTime - b->p2;
      L.sym.value = 0;
      L.sym.pattern = A_Z;
   }
   else // Pattern-X
   {
      m->symbolStartTime = m->symbolEndTime;
      m->symbolEndTime = m->correlatedPeakTime;
      m->symbolRiseTime = m->correlatedPeakTime;
      L.sym.value = 1;
      L.sym.pattern = A_X;
   }

   m->searchSyncTime = m->symbolEndTime + b->p1; // :916-923
   m->searchStartTime = m->searchSyncTime - b->p8;
   m->searchEndTime = m->searchSyncTime + b->p8;
   m->searchCorrDValue = 0;
   m->searchCorr0Value = 0;
   m->searchCorr1Value = 0;
   m->correlatedPeakTime = 0;
   m->correlatedPeakValue = 0;

   L.sym.start = m->symbolStartTime - b->sdd;
   L.sym.end = m->symbolEndTime - b->sdd;
   L.sym.edge = m->symbolRiseTime - b->sdd;
   L.sym.length = L.sym.end - L.sym.start;

   return (int) L.sym.pattern;
}

// one sample of decodePollFrame, NfcA.cpp:432-563
static void A_poll_step(Dec *d)
{
   int pattern = A_poll_symbol(d);

   if (pattern <= A_No)
      return;

   TechSt *t = &(L.c.t[TECH_A]);
   Bits *st = &(L.st);
   bool frameEnd = false, truncateError = false;

   st->pattern = (u32) pattern;

   if (st->pattern == A_Y && (st->previous == A_Y || st->previous == A_Z))
      frameEnd = true;
   else if (st->bytes == t->ps.maxFrameSize)
      truncateError = true;

   if (frameEnd || truncateError)
   {
      if (st->bytes > 0 || st->bits == 7)
      {
         if (st->bits >= 7)
            put_byte(d, st->data);

         u32 flags = 0, phase = 0;

         if (st->flags & FL_Parity)
            flags |= FL_Parity;
         if (truncateError)
            flags |= FL_Truncated;
         if (st->bytes == 1 && st->bits == 7)
            flags |= FL_Short;

         u32 len = st->bytes, rate = t->fs.symbolRate, start = t->fs.frameStart, end = t->fs.frameEnd;

         A_process(d, FT_Poll, len, &flags, &phase);

         emit(d, TT_A, FT_Poll, flags, phase, rate, start, end, sb, len);

         clear_bits(d);

         if (L.lock == LOCK_A) // :491-511
            clear_for_listen(d, &L.c.mA[L.lockRate], P.A[L.lockRate].corr, P.A[L.lockRate].p1);

         return;
      }

      A_reset(d);
      return;
   }

   if (L.sym.edge) // :525
      t->fs.frameEnd = L.sym.edge;

   if (st->previous)
   {
      u32 value = (st->previous == A_X);

      if (st->bits < 8)
      {
         st->data = st->data | (value << st->bits++);
      }
      else if (st->bytes < t->ps.maxFrameSize)
      {
         put_byte(d, st->data);
         st->flags |= !odd_parity_ok(st->data, value) ? FL_Parity : 0;
         st->data = st->bits = 0;
      }
      else
      {
         A_reset(d);
         return;
      }
   }

   st->previous = st->pattern;
}

// NfcA::Impl::resetFrameSearch, NfcA.cpp:1426-1446
static void A_reset_frame_search(Dec *d)
{
   if (L.lock == LOCK_A)
   {
      Mod *m = &(L.c.mA[L.lockRate]);
      m->symbolStartTime = 0;
      m->symbolEndTime = 0;
      m->symbolRiseTime = 0;
      m->searchSyncTime = 0;
      m->searchStartTime = 0;
      m->searchEndTime = 0;
      m->searchPulseWidth = 0;
      m->correlatedPeakTime = 0;
      m->correlatedPeakValue = 0;
      m->detectorPeakTime = 0;
      m->detectorPeakValue = 0;
   }
   L.c.t[TECH_A].fs.frameStart = 0;
}

// common part of the ASK listen integrator (w^2 * 10 over half a symbol), NfcA.cpp:955-973 / 1110-1130
static void A_listen_ask_integrate(Dec *d, const RateParams *b, Mod *m, float *s0_, float *s1_)
{
   float s0, s1;
   u32 fp1 = L.fe.cA[L.lockRate], fp2, fp3;
   corr_points(fp1, b->p1, b->p2, &fp2, &fp3);

   float data = SMP(NFCB200_OFF_W, b->sdd);
   float v = data * data * 10;

   SMP(NFCB200_OFF_I, b->sdd) = v;

   m->filterIntegrate += v;
   m->filterIntegrate -= SMP(NFCB200_OFF_I, b->sdd + b->p2);

   RG(b->corr, fp1) = m->filterIntegrate;

   s0 = m->filterIntegrate - RG(b->corr, fp2);
   s1 = RG(b->corr, fp2) - RG(b->corr, fp3);
   *s0_ = s0; *s1_ = s1;
}

// one sample of decodeListenFrameStartAsk, NfcA.cpp:939-1090
static int A_listen_start_ask(Dec *d)
{
   const RateParams *b = &(P.A[L.lockRate]);
   Mod *m = &(L.c.mA[L.lockRate]);
   FrameSt *fs = &(L.c.t[TECH_A].fs);
   const u32 clk = L.fe.clk;

   float s0, s1;
   A_listen_ask_integrate(d, b, m, &s0, &s1);

   float deep = SMP(NFCB200_OFF_M, 0); // futureIndex

   if (clk < fs->guardEnd)
      return A_Invalid;

   if (clk == fs->guardEnd)
      m->searchValueThreshold = SMP(NFCB200_OFF_D, b->sdd) * (float) b->p8;

   if (clk > fs->waitingEnd)
      return A_No;

   if (deep > P.thr[TECH_A].modMin)
      return A_No;

   if (!m->symbolStartTime)
   {
      if (s0 > m->searchValueThreshold && s0 > m->correlatedPeakValue)
      {
         m->correlatedPeakValue = s0;
         m->correlatedPeakTime = clk;
         m->searchEndTime = clk + b->p4;
      }
   }
   else
   {
      if (s0 < -m->searchValueThreshold && s0 < m->correlatedPeakValue)
      {
         m->correlatedPeakValue = s0;
         m->correlatedPeakTime = clk;
      }
   }

   if (clk != m->searchEndTime)
      return A_Invalid;

   if (!m->symbolStartTime) // :1032-1040
   {
      m->searchSyncTime = m->correlatedPeakTime + b->p2;
      m->searchEndTime = m->searchEndTime + b->p2;
      m->symbolStartTime = m->correlatedPeakTime - b->p2;
      m->correlatedPeakTime = 0;
      m->correlatedPeakValue = 0;
      return A_Invalid;
   }

   m->symbolEndTime = m->correlatedPeakTime;
   m->searchPulseWidth = m->symbolEndTime - m->symbolStartTime;

   u32 minimumPulseWidth = b->p1 - b->p8;
   u32 maximumPulseWidth = b->p1 + b->p8;

   if (m->correlatedPeakTime == 0 || m->searchPulseWidth < minimumPulseWidth || m->searchPulseWidth > maximumPulseWidth)
   {
      m->symbolStartTime = 0;
      m->symbolEndTime = 0;
      m->searchSyncTime = 0;
      m->searchStartTime = 0;
      m->searchEndTime = 0;
      m->searchPulseWidth = 0;
      m->correlatedPeakTime = 0;
      m->correlatedPeakValue = 0;
      m->detectorPeakTime = 0;
      m->detectorPeakValue = 0;
      return A_Invalid;
   }

   m->searchSyncTime = m->symbolEndTime + b->p1;
   m->searchStartTime = m->searchSyncTime - b->p8;
   m->searchEndTime = m->searchSyncTime + b->p8;
   m->searchValueThreshold = fabsf(m->correlatedPeakValue * 0.25f);
   m->searchCorr0Value = 0;
   m->searchCorr1Value = 0;
   m->correlatedPeakTime = 0;
   m->correlatedPeakValue = 0;

   L.sym.value = 1;
   L.sym.start = m->symbolStartTime - b->sdd;
   L.sym.end = m->symbolEndTime - b->sdd;
   L.sym.length = L.sym.end - L.sym.start;
   L.sym.pattern = A_D;

   return A_D;
}

// one sample of decodeListenFrameSymbolAsk, NfcA.cpp:1095-1214
static int A_listen_symbol_ask(Dec *d)
{
   const RateParams *b = &(P.A[L.lockRate]);
   Mod *m = &(L.c.mA[L.lockRate]);
   const u32 clk = L.fe.clk;

   float s0, s1;
   A_listen_ask_integrate(d, b, m, &s0, &s1);
   float sd = fabsf(s0 - s1);

   if (clk < m->searchStartTime)
      return A_Invalid;

   if (sd > m->correlatedPeakValue)
   {
      m->correlatedPeakValue = sd;
      m->correlatedPeakTime = clk;
   }

   if (clk == m->searchSyncTime)
   {
      m->searchCorrDValue = sd;
      m->searchCorr0Value = s0;
      m->searchCorr1Value = s1;
   }

   if (clk != m->searchEndTime)
      return A_Invalid;

   if (m->searchCorrDValue > m->searchValueThreshold)
   {
      m->symbolStartTime = m->symbolEndTime;
      m->symbolEndTime = m->correlatedPeakTime;
      m->searchValueThreshold = m->correlatedPeakValue * 0.25f;

      if (m->searchCorr0Value > m->searchCorr1Value)
      {
         m->symbolRiseTime = m->searchSyncTime;
         L.sym.value = 0;
         L.sym.pattern = A_E;
      }
      else
      {
         m->symbolRiseTime = m->searchSyncTime - b->p2;
         L.sym.value = 1;
         L.sym.pattern = A_D;
      }
   }
   else
   {
      m->symbolStartTime = m->symbolEndTime;
      m->symbolEndTime = m->searchSyncTime;
      m->symbolRiseTime = 0;
      L.sym.pattern = A_F;
   }

   m->searchSyncTime = m->symbolEndTime + b->p1;
   m->searchStartTime = m->searchSyncTime - b->p8;
   m->searchEndTime = m->searchSyncTime + b->p8;
   m->correlatedPeakTime = 0;
   m->correlatedPeakValue = 0;

   L.sym.start = m->symbolStartTime - b->sdd;
   L.sym.end = m->symbolEndTime - b->sdd;
   L.sym.edge = m->symbolRiseTime - b->sdd;
   L.sym.length = L.sym.end - L.sym.start;

   return (int) L.sym.pattern;
}

// one sample of decodeListenFrameStartBpsk, NfcA.cpp:1220-1329
static int A_listen_start_bpsk(Dec *d)
{
   const RateParams *b = &(P.A[L.lockRate]);
   Mod *m = &(L.c.mA[L.lockRate]);
   FrameSt *fs = &(L.c.t[TECH_A].fs);
   const u32 clk = L.fe.clk;

   float data = SMP(NFCB200_OFF_W, b->sdd);
   float delay1 = SMP(NFCB200_OFF_W, b->sdd + b->p1);
   float deep = SMP(NFCB200_OFF_M, 0);

   float v = data * delay1 * 10;
   SMP(NFCB200_OFF_I, b->sdd) = v;

   if (clk < fs->guardEnd)
      return A_Invalid;

   if (clk == fs->guardEnd)
      m->searchValueThreshold = SMP(NFCB200_OFF_D, b->sdd);

   if (clk > fs->waitingEnd)
      return A_No;

   if (deep > P.thr[TECH_A].modMin)
      return A_No;

   m->phaseIntegrate += v;
   m->phaseIntegrate -= SMP(NFCB200_OFF_I, b->sdd + b->p4);

   if (m->phaseIntegrate > m->searchValueThreshold) // :1277
   {
      if (!m->symbolStartTime)
         m->symbolStartTime = clk;

      m->searchEndTime = clk + b->p2;
   }

   if (!m->symbolEndTime && (m->phaseIntegrate < 0 || clk == m->searchEndTime)) // :1286
   {
      int preambleSyncLength = (int) (clk - m->symbolStartTime);

      if (preambleSyncLength < P.etu * 3 || preambleSyncLength > P.etu * 4)
      {
         m->symbolStartTime = 0;
         m->symbolEndTime = 0;
         m->searchEndTime = 0;
         return A_Invalid;
      }

      m->symbolEndTime = m->searchEndTime + b->p2;
   }

   if (clk != m->searchEndTime)
      return A_Invalid;

   m->searchSyncTime = m->symbolEndTime + b->p2; // :1311-1316
   m->searchLastPhase = m->phaseIntegrate;
   m->searchPhaseThreshold = fabsf(m->phaseIntegrate * 0.25f);
   m->detectorPeakTime = 0;

   L.sym.value = 0;
   L.sym.start = m->symbolStartTime - b->p1 - b->sdd;
   L.sym.end = m->symbolEndTime - b->p1 - b->sdd;
   L.sym.length = L.sym.end - L.sym.start;
   L.sym.pattern = A_S;

   return A_S;
}

// shared BPSK symbol step of NFC-A (NfcA.cpp:1334-1421) and NFC-B (NfcB.cpp:954-1040): identical apart from the
// pattern codes.  Returns 0 none, 1 end-of-frame (PatternO), 2 symbol
static int bpsk_symbol(Dec *d, const RateParams *b, Mod *m, bool *toggled_)
{
   bool toggled = *toggled_;
   const u32 clk = L.fe.clk;

   float data = SMP(NFCB200_OFF_W, b->sdd);
   float delay1 = SMP(NFCB200_OFF_W, b->sdd + b->p1);

   float v = data * delay1 * 10;
   SMP(NFCB200_OFF_I, b->sdd) = v;

   m->phaseIntegrate += v;
   m->phaseIntegrate -= SMP(NFCB200_OFF_I, b->sdd + b->p4);

   if (!m->detectorPeakTime)
   {
      if ((m->phaseIntegrate > 0 && m->searchLastPhase < 0) || (m->phaseIntegrate < 0 && m->searchLastPhase > 0))
      {
         m->detectorPeakTime = clk;
         m->searchSyncTime = clk + b->p2;
         m->searchLastPhase = m->phaseIntegrate;
      }
   }

   if (clk != m->searchSyncTime)
      { *toggled_ = toggled; return 0; }

   if (fabsf(m->phaseIntegrate) < fabsf(m->searchPhaseThreshold))
      { *toggled_ = toggled; return 1; }

   m->symbolStartTime = m->symbolEndTime;
   m->symbolEndTime = m->searchSyncTime + b->p2;
   m->searchSyncTime = m->searchSyncTime + b->p1;
   m->searchLastPhase = m->phaseIntegrate;
   m->detectorPeakTime = 0;

   toggled = false;

   if (m->phaseIntegrate < -m->searchPhaseThreshold)
      toggled = true;
   else
      m->searchPhaseThreshold = m->phaseIntegrate * 0.25f;

   { *toggled_ = toggled; return 2; }
   *toggled_ = toggled;
}

static void A_emit_listen(Dec *d, u32 flags)
{
   TechSt *t = &(L.c.t[TECH_A]);
   u32 phase = 0;
   u32 len = L.st.bytes, rate = P.A[L.lockRate].sps, start = t->fs.frameStart, end = t->fs.frameEnd;
   A_process(d, FT_Listen, len, &flags, &phase);
   emit(d, TT_A, FT_Listen, flags, phase, rate, start, end, sb, len);
   A_reset(d);
}

// one sample of decodeListenFrame, NfcA.cpp:568-807
static void A_listen_step(Dec *d)
{
   TechSt *t = &(L.c.t[TECH_A]);
   FrameSt *fs = &(t->fs);
   Bits *st = &(L.st);
   bool frameEnd = false, truncateError = false;

   if (L.lockRate == 0) // 106k ASK / Manchester
   {
      if (!fs->frameStart)
      {
         int pattern = A_listen_start_ask(d);

         if (pattern == A_D)
            fs->frameStart = L.sym.start;
         else if (pattern == A_No)
            A_reset(d);

         return;
      }

      int pattern = A_listen_symbol_ask(d);

      if (pattern <= A_No)
         return;

      if (pattern == A_F)
         frameEnd = true;
      else if (st->bytes == t->ps.maxFrameSize)
         truncateError = true;

      if (frameEnd || truncateError)
      {
         if (st->bytes > 0 || st->bits == 4)
         {
            if (st->bits == 4)
               put_byte(d, st->data);

            u32 flags = 0;
            if (st->flags & FL_Parity)
               flags |= FL_Parity;
            if (truncateError)
               flags |= FL_Truncated;
            if (st->bytes == 1 && st->bits == 4)
               flags |= FL_Short;

            A_emit_listen(d, flags);
            return;
         }

         A_reset_frame_search(d); // :653
         return;
      }

      if (L.sym.edge)
         fs->frameEnd = L.sym.edge;

      if (st->bits < 8)
      {
         st->data |= (L.sym.value << st->bits++);
      }
      else if (st->bytes < t->ps.maxFrameSize)
      {
         put_byte(d, st->data);
         st->flags |= !odd_parity_ok(st->data, L.sym.value) ? FL_Parity : 0;
         st->data = st->bits = 0;
      }
      else
      {
         A_reset(d);
      }

      return;
   }

   // 212k / 424k BPSK
   if (!fs->frameStart)
   {
      int pattern = A_listen_start_bpsk(d);

      if (pattern == A_S)
         fs->frameStart = L.sym.start;
      else if (pattern == A_No)
         A_reset(d);

      return;
   }

   const RateParams *b = &(P.A[L.lockRate]);
   Mod *m = &(L.c.mA[L.lockRate]);
   bool toggled = false;
   int r = bpsk_symbol(d, b, m, &toggled);

   if (r == 0)
      return;

   int pattern;

   if (r == 1)
   {
      pattern = A_O;
   }
   else
   {
      if (toggled)
      {
         L.sym.value = !L.sym.value;
         L.sym.pattern = (L.sym.pattern == A_M) ? A_N : A_M;
      }

      L.sym.start = m->symbolStartTime - b->p1 - b->sdd;
      L.sym.end = m->symbolEndTime - b->p1 - b->sdd;
      L.sym.length = L.sym.end - L.sym.start;
      pattern = (int) L.sym.pattern;

      if (pattern <= A_No) // `while ((pattern = ...) > NoPattern)`
         return;
   }

   if (pattern == A_O)
      frameEnd = true;
   else if (st->bytes == t->ps.maxFrameSize)
      truncateError = true;

   if (frameEnd || truncateError)
   {
      if (st->bits == 9)
      {
         put_byte(d, st->data);
         st->flags |= odd_parity_ok(st->data, st->parity) ? FL_Parity : 0; // last byte: inverted parity, :734
      }

      if (st->bytes > 0)
      {
         fs->frameEnd = L.sym.end;

         u32 flags = 0;
         if (st->flags & FL_Parity)
            flags |= FL_Parity;
         if (truncateError)
            flags |= FL_Truncated;

         A_emit_listen(d, flags);
         return;
      }

      A_reset(d);
      return;
   }

   if (st->bits < 8)
   {
      st->data |= (L.sym.value << st->bits);
   }
   else if (st->bits < 9)
   {
      st->parity = L.sym.value;
   }
   else
   {
      put_byte(d, st->data);
      st->flags |= !odd_parity_ok(st->data, st->parity) ? FL_Parity : 0;
      st->data = L.sym.value;
      st->bits = 0;
   }

   st->bits++;
}

// ------------------------------------------------------------------------------------------------------------------
// NFC-B
// ------------------------------------------------------------------------------------------------------------------
enum { B_Invalid = 0, B_No = 1, B_L = 2, B_H = 3, B_S = 4, B_M = 5, B_N = 6, B_O = 7 };

// NfcB::Impl::resetModulation, NfcB.cpp:1045-1069
static void B_reset(Dec *d)
{
   zero_mod(&L.c.mB[0]);
   zero_mod(&L.c.mB[1]);
   zero_ring(d, NFCB200_OFF_I, NFCB200_RING);
   clear_bits(d);
   clear_sym(d);
   L.c.t[TECH_B].fs.frameType = 0;
   L.c.t[TECH_B].fs.frameStart = 0;
   L.c.t[TECH_B].fs.frameEnd = 0;
   L.lock = LOCK_NONE;
}

static void B_clear_search(Mod *m, bool sync)
{
   m->symbolStartTime = 0;
   m->symbolEndTime = 0;
   m->searchStartTime = 0;
   m->searchEndTime = 0;
   if (sync)
      m->searchSyncTime = 0;
   m->detectorPeakTime = 0;
   m->detectorPeakValue = 0;
}

// NfcB::Impl::detectModulation, NfcB.cpp:238-432
static bool B_detect(Dec *d)
{
   const u32 clk = L.fe.clk;
   const float env = L.fe.env;

   for (int rate = 0; rate <= 1; rate++)
   {
      const RateParams *b = &(P.B[rate]);
      Mod *m = &(L.c.mB[rate]);

      float edge = SMP(NFCB200_OFF_W, b->sdd);
      float deep = SMP(NFCB200_OFF_M, b->sdd);

      // :265-274
      if (deep > P.thr[TECH_B].modMax || (m->detectorPeakTime && clk > m->detectorPeakTime + b->p1))
         B_clear_search(m, true);

      if (!m->symbolStartTime) // :277-305
      {
         m->searchValueThreshold = env * P.thr[TECH_B].modMin;

         if (edge < -m->searchValueThreshold && edge < m->detectorPeakValue)
         {
            m->detectorPeakValue = edge;
            m->detectorPeakTime = clk;
            m->searchEndTime = clk + b->p4;
         }

         if (clk != m->searchEndTime)
            continue;

         m->symbolStartTime = m->detectorPeakTime - b->p8;
         m->searchStartTime = m->symbolStartTime + (10 * b->p1) - b->p2;
         m->searchEndTime = m->symbolStartTime + (11 * b->p1) + b->p2;
         m->searchValueThreshold = fabsf(m->detectorPeakValue * 0.5f);
         m->detectorPeakValue = 0;
         m->detectorPeakTime = 0;
         continue;
      }

      if (!m->symbolEndTime) // :308-361
      {
         if (clk < m->searchStartTime)
         {
            if (edge > m->searchValueThreshold)
            {
               m->symbolStartTime = 0;
               m->symbolEndTime = 0;
               m->searchStartTime = 0;
               m->searchEndTime = 0;
               m->detectorPeakTime = 0;
               m->detectorPeakValue = 0;
            }
            continue;
         }

         if (edge > m->searchValueThreshold && edge > m->detectorPeakValue)
         {
            m->detectorPeakValue = edge;
            m->detectorPeakTime = clk;
            m->searchEndTime = clk + b->p4;
         }

         if (clk != m->searchEndTime)
            continue;

         if (!m->detectorPeakTime)
         {
            m->symbolStartTime = 0;
            m->symbolEndTime = 0;
            m->searchStartTime = 0;
            m->searchEndTime = 0;
            m->detectorPeakValue = 0;
            continue;
         }

         m->symbolEndTime = m->detectorPeakTime;
         m->searchStartTime = m->detectorPeakTime + (2 * b->p1) - b->p2;
         m->searchEndTime = m->detectorPeakTime + (3 * b->p1) + b->p2;
         m->searchValueThreshold = fabsf(m->detectorPeakValue) / 2;
         m->detectorPeakValue = 0;
         m->detectorPeakTime = 0;
         continue;
      }

      if (clk < m->searchStartTime) // :364-377
      {
         if (edge < -m->searchValueThreshold)
         {
            m->symbolStartTime = 0;
            m->symbolEndTime = 0;
            m->searchStartTime = 0;
            m->searchEndTime = 0;
            m->detectorPeakTime = 0;
            m->detectorPeakValue = 0;
         }
         continue;
      }

      if (edge < -m->searchValueThreshold && m->detectorPeakValue > edge) // :380
      {
         m->detectorPeakValue = edge;
         m->detectorPeakTime = clk;
         m->searchEndTime = clk + b->p4;
      }

      if (clk != m->searchEndTime)
         continue;

      if (!m->detectorPeakTime) // :392-402 (note: `break`, the 212k detector is skipped for this sample)
      {
         m->symbolStartTime = 0;
         m->symbolEndTime = 0;
         m->searchStartTime = 0;
         m->searchEndTime = 0;
         m->detectorPeakTime = 0;
         m->detectorPeakValue = 0;
         break;
      }

      m->symbolEndTime = m->detectorPeakTime; // :408-428
      m->searchSyncTime = m->symbolEndTime + b->p2;
      m->searchStartTime = 0;
      m->searchEndTime = 0;
      m->searchValueThreshold = fabsf(m->detectorPeakValue * 0.5f);
      m->detectorPeakTime = 0;
      m->detectorPeakValue = 0;

      FrameSt *fs = &(L.c.t[TECH_B].fs);
      fs->frameType = FT_Poll;
      fs->symbolRate = b->sps;
      fs->frameStart = m->symbolStartTime - b->sdd;
      fs->frameEnd = 0;

      L.lock = LOCK_B;
      L.lockRate = rate;
      return true;
   }

   return false;
}

// NfcB::Impl::checkCrc, NfcB.cpp:1272-1283
static bool B_crc_ok(Dec *d, u32 size)
{
   if (size < 3)
      return false;
   unsigned short crc = (unsigned short) ~crc_ccitt16(sb, 0, size - 2, 0xFFFF, true);
   unsigned short res = (unsigned short) ((sb[size - 2] & 0xff) | ((sb[size - 1] & 0xff) << 8));
   return res == crc;
}

// NfcB::Impl::process, NfcB.cpp:1074-1267
static void B_process(Dec *d, u32 type, u32 len, u32 *flags_, u32 *phase_)
{
   u32 flags = *flags_, phase = *phase_;
   TechSt *t = &(L.c.t[TECH_B]);
   FrameSt *fs = &(t->fs);
   Proto *ps = &(t->ps);
   const bool poll = type == FT_Poll;
   const u32 b0 = fb(d, 0, len);

   if (poll)
   {
      fs->startUpGuardTime = ps->startUpGuardTime;
      fs->frameWaitingTime = ps->frameWaitingTime;
      fs->frameGuardTime = ps->frameGuardTime;
      fs->requestGuardTime = ps->requestGuardTime;
   }
   else
   {
      fs->frameGuardTime = ps->frameGuardTime;
   }

   do
   {
      // processREQB :1153-1206
      if (poll)
      {
         if (b0 == 0x05 && len == 5)
         {
            fs->lastCommand = b0;
            ps->maxFrameSize = 256;
            ps->startUpGuardTime = P.B_sfgt;
            ps->frameGuardTime = P.B_fgt;
            ps->frameWaitingTime = P.B_fwt;
            ps->requestGuardTime = P.B_rgt;
            fs->frameGuardTime = P.B_tr0min;
            fs->frameWaitingTime = P.B_fwtAtqb;
            t->chained = 0;
            phase = PH_Selection;
            flags |= !B_crc_ok(d, len) ? FL_Crc : 0;
            break;
         }
      }
      else if (fs->lastCommand == 0x05)
      {
         int fdsi = (fb(d, 10, len) >> 4) & 0x0f;
         int fwi = (fb(d, 11, len) >> 4) & 0x0f;
         ps->maxFrameSize = (u32) NFC_FDS_TABLE_[fdsi];
         ps->frameWaitingTime = (u32) (int) (P.stu * NFC_XGT_TABLE_[fwi]);
         phase = PH_Selection;
         flags |= !B_crc_ok(d, len) ? FL_Crc : 0;
         break;
      }

      // processATTRIB :1212-1258
      if (poll)
      {
         if (b0 == 0x1d && len > 10)
         {
            fs->lastCommand = b0;
            u32 param1 = fb(d, 5, len), param2 = fb(d, 6, len);
            u32 tr0i = (param1 >> 6) & 0x3;
            u32 fdsi = param2 & 0xf;
            ps->maxFrameSize = (u32) NFC_FDS_TABLE_[fdsi];
            if (!tr0i)
               ps->frameGuardTime = P.B_fgt;
            else
               ps->frameGuardTime = (u32) (int) (P.stu * (tr0i == 1 ? 48 * 16 : tr0i == 2 ? 16 * 16 : 0));
            fs->frameWaitingTime = P.fwtActivation;
            t->chained = 0;
            phase = PH_Selection;
            flags |= !B_crc_ok(d, len) ? FL_Crc : 0;
            break;
         }
      }
      else if (fs->lastCommand == 0x1d)
      {
         phase = PH_Selection;
         break;
      }

      // processOther :1263-1267
      phase = PH_Application;
      flags |= !B_crc_ok(d, len) ? FL_Crc : 0;
   }
   while (false);

   flags |= t->chained;

   if (poll)
   {
      if (L.lock == LOCK_B)
      {
         u32 sdd = P.B[L.lockRate].sdd;
         fs->guardEnd = fs->frameEnd + fs->frameGuardTime + sdd;
         fs->waitingEnd = fs->frameEnd + fs->frameWaitingTime + sdd;
         fs->frameType = FT_Listen;
      }
   }
   else
   {
      if (L.lock == LOCK_B)
         fs->guardEnd = fs->frameEnd + fs->frameGuardTime + P.B[L.lockRate].sdd;
      fs->frameType = 0;
      fs->lastCommand = 0;
   }

   fs->frameStart = 0;
   fs->frameEnd = 0;
   *flags_ = flags; *phase_ = phase;
}

// one sample of decodePollFrameSymbolAsk, NfcB.cpp:684-762
static int B_poll_symbol(Dec *d)
{
   const RateParams *b = &(P.B[L.lockRate]);
   Mod *m = &(L.c.mB[L.lockRate]);
   const u32 clk = L.fe.clk;

   float edge = SMP(NFCB200_OFF_W, b->sdd);
   float deep = SMP(NFCB200_OFF_M, b->sdd);

   if (clk > m->searchStartTime && clk < m->searchEndTime)
   {
      edge = fabsf(edge);

      if (edge > m->searchValueThreshold && m->detectorPeakValue < edge)
      {
         m->detectorPeakValue = edge;
         m->searchSyncTime = clk + b->p2;
      }
   }

   if (clk != m->searchSyncTime)
      return B_Invalid;

   m->symbolStartTime = m->symbolEndTime;
   m->symbolEndTime = m->searchSyncTime + b->p2;
   m->searchStartTime = m->searchSyncTime + b->p4;
   m->searchEndTime = m->searchStartTime + b->p2;
   m->searchSyncTime = m->searchSyncTime + b->p1;
   m->detectorPeakValue = 0;

   if (deep > P.thr[TECH_B].modMin)
   {
      L.sym.value = 0;
      L.sym.pattern = B_L;
   }
   else
   {
      L.sym.value = 1;
      L.sym.pattern = B_H;
   }

   L.sym.start = m->symbolStartTime - b->sdd;
   L.sym.end = m->symbolEndTime - b->sdd;
   L.sym.length = L.sym.end - L.sym.start;

   return (int) L.sym.pattern;
}

// one sample of decodePollFrame, NfcB.cpp:453-567
static void B_poll_step(Dec *d)
{
   int pattern = B_poll_symbol(d);

   if (pattern <= B_No)
      return;

   TechSt *t = &(L.c.t[TECH_B]);
   Bits *st = &(L.st);
   bool frameEnd = false, truncateError = false, streamError = false;

   if (st->bits == 9 && !st->data && pattern == B_L)
      frameEnd = true;
   else if (st->bits == 9 && pattern == B_L)
      streamError = true;
   else if (st->bits == 0 && pattern == B_H && st->skip == 6)
      streamError = true;
   else if (st->bytes == t->ps.maxFrameSize)
      truncateError = true;
   else if ((st->bits == 0 && pattern == B_H) && ++st->skip)
      return;

   if (frameEnd || streamError || truncateError)
   {
      if (st->bytes > 2)
      {
         t->fs.frameEnd = L.sym.end;

         u32 flags = 0, phase = 0;
         if (truncateError || streamError)
            flags |= FL_Truncated;

         u32 len = st->bytes, rate = P.B[L.lockRate].sps, start = t->fs.frameStart, end = t->fs.frameEnd;

         B_process(d, FT_Poll, len, &flags, &phase);
         emit(d, TT_B, FT_Poll, flags, phase, rate, start, end, sb, len);

         clear_bits(d);

         if (L.lock == LOCK_B)
            clear_for_listen(d, &L.c.mB[L.lockRate], 0, 0);

         return;
      }

      B_reset(d);
      return;
   }

   if (st->bits < 9)
   {
      if (st->bits > 0)
         st->data |= (L.sym.value << (st->bits - 1));
      st->bits++;
   }
   else
   {
      put_byte(d, st->data);
      st->data = 0;
      st->bits = 0;
      st->skip = 0;
   }
}

// one sample of decodeListenFrameStartBpsk, NfcB.cpp:767-949
static int B_listen_start(Dec *d)
{
   const RateParams *b = &(P.B[L.lockRate]);
   Mod *m = &(L.c.mB[L.lockRate]);
   FrameSt *fs = &(L.c.t[TECH_B].fs);
   const u32 clk = L.fe.clk;

   float data = SMP(NFCB200_OFF_W, b->sdd);
   float delay1 = SMP(NFCB200_OFF_W, b->sdd + b->p1);
   float deep = SMP(NFCB200_OFF_M, 0);

   float v = data * delay1 * 10;
   SMP(NFCB200_OFF_I, b->sdd) = v;

   m->phaseIntegrate += v; // integrates always (:793-794)
   m->phaseIntegrate -= SMP(NFCB200_OFF_I, b->sdd + b->p4);

   if (clk < fs->guardEnd)
      return B_Invalid;

   if (clk == fs->guardEnd)
      m->searchValueThreshold = SMP(NFCB200_OFF_D, b->sdd);

   if (clk > fs->waitingEnd)
      return B_No;

   if (deep > P.thr[TECH_B].modMax)
      return B_No;

   if (clk < m->searchStartTime)
      return B_Invalid;

   if (m->phaseIntegrate > m->searchValueThreshold)
   {
      if (!m->symbolStartTime)
         m->symbolStartTime = clk;

      m->searchEndTime = clk + b->p2;
   }

   if (clk != m->searchEndTime && m->phaseIntegrate > 0)
      return B_Invalid;

   u32 length, lo, hi;

   switch (m->searchModeState)
   {
      case 0: // LISTEN_MODE_TR1
         length = clk - m->symbolStartTime;
         lo = P.B_tr1Min;
         hi = P.B_tr1Max;
         break;
      case 1: // LISTEN_MODE_SOS_S1
         length = clk - m->symbolEndTime;
         lo = P.B_s1Min;
         hi = P.B_s1Max;
         break;
      case 2: // LISTEN_MODE_SOS_S2
         length = clk - m->symbolEndTime;
         lo = P.B_s2Min;
         hi = P.B_s2Max;
         break;
      default:
         return B_Invalid; // the reference's switch has no default: falls out of the switch, loops
   }

   if (length < lo || length > hi) // int vs unsigned in the reference: compared unsigned
   {
      m->searchModeState = 0;
      m->searchStartTime = 0;
      m->searchEndTime = 0;
      m->symbolStartTime = 0;
      m->symbolEndTime = 0;
      return B_Invalid;
   }

   m->symbolEndTime = clk;

   if (m->searchModeState < 2)
   {
      m->searchModeState++;
      m->searchStartTime = clk + b->p1 + b->p4;
      m->searchEndTime = 0;
      return B_Invalid;
   }

   m->searchSyncTime = clk + b->p2; // :927-943
   m->searchLastPhase = m->phaseIntegrate;
   m->searchPhaseThreshold = fabsf(m->detectorPeakValue * 0.25f);
   m->searchStartTime = 0;
   m->searchEndTime = 0;
   m->detectorPeakValue = 0;

   L.sym.value = 1;
   L.sym.start = m->symbolStartTime - b->p1 - b->sdd;
   L.sym.end = m->symbolEndTime - b->p1 - b->sdd;
   L.sym.length = L.sym.end - L.sym.start;
   L.sym.pattern = B_S;

   return B_S;
}

// one sample of decodeListenFrame, NfcB.cpp:572-679
static void B_listen_step(Dec *d)
{
   TechSt *t = &(L.c.t[TECH_B]);
   FrameSt *fs = &(t->fs);
   Bits *st = &(L.st);

   if (!fs->frameStart)
   {
      int pattern = B_listen_start(d);

      if (pattern == B_S)
         fs->frameStart = L.sym.start;
      else if (pattern == B_No)
         B_reset(d);

      return;
   }

   const RateParams *b = &(P.B[L.lockRate]);
   Mod *m = &(L.c.mB[L.lockRate]);
   bool toggled = false;
   int r = bpsk_symbol(d, b, m, &toggled);

   if (r == 0)
      return;

   int pattern;

   if (r == 1)
   {
      pattern = B_O;
   }
   else
   {
      if (toggled)
      {
         L.sym.value = !L.sym.value;
         L.sym.pattern = (L.sym.pattern == B_M) ? B_N : B_M;
      }

      L.sym.start = m->symbolStartTime - b->p1 - b->sdd;
      L.sym.end = m->symbolEndTime - b->p1 - b->sdd;
      L.sym.length = L.sym.end - L.sym.start;
      pattern = (int) L.sym.pattern;

      if (pattern <= B_No)
         return;
   }

   bool frameEnd = false, truncateError = false, streamError = false;

   if (st->bits == 9 && !st->data && pattern == B_M)
      frameEnd = true;
   else if ((st->bits == 0 && pattern == B_N) || (st->bits == 9 && pattern == B_M))
      streamError = true;
   else if (st->bytes == t->ps.maxFrameSize)
      truncateError = true;

   if (frameEnd || streamError || truncateError)
   {
      if (st->bytes > 0)
      {
         fs->frameEnd = L.sym.end + P.B_eofComp;

         u32 flags = 0, phase = 0;
         if (truncateError || streamError)
            flags |= FL_Truncated;

         u32 len = st->bytes, rate = b->sps, start = fs->frameStart, end = fs->frameEnd;

         B_process(d, FT_Listen, len, &flags, &phase);
         emit(d, TT_B, FT_Listen, flags, phase, rate, start, end, sb, len);
      }

      B_reset(d);
      return;
   }

   if (st->bits < 9)
   {
      if (st->bits > 0)
         st->data |= (L.sym.value << (st->bits - 1));
      st->bits++;
   }
   else
   {
      put_byte(d, st->data);
      st->data = 0;
      st->bits = 0;
   }
}

// ------------------------------------------------------------------------------------------------------------------
// NFC-F
// ------------------------------------------------------------------------------------------------------------------
enum { F_Invalid = 0, F_No = 1, F_L = 2, F_H = 3, F_S = 4, F_E = 5 };

// NfcF::Impl::resetModulation, NfcF.cpp:1047-1071
static void F_reset(Dec *d)
{
   for (int r = 0; r < 2; r++)
   {
      zero_mod(&L.c.mF[r]);
      zero_ring(d, P.F[r + 1].corr, P.F[r + 1].p1);
   }
   clear_bits(d);
   clear_sym(d);
   L.c.t[TECH_F].fs.frameType = 0;
   L.c.t[TECH_F].fs.frameStart = 0;
   L.c.t[TECH_F].fs.frameEnd = 0;
   L.lock = LOCK_NONE;
}

static void F_restart_search(Mod *m)
{
   m->symbolStartTime = 0;
   m->symbolEndTime = 0;
   m->searchSyncTime = 0;
   m->searchSyncValue = 0;
   m->searchStartTime = 0;
   m->searchEndTime = 0;
   m->searchPulseWidth = 0;
   m->searchValueThreshold = 0;
   m->correlatedPeakValue = 0;
   m->correlatedPeakTime = 0;
}

// the preamble tracker shared by detectModulation (NfcF.cpp:273-404) and decodeListenFrameStartAsk (:810-932).
// `ge` selects the listen variant's `>=` threshold test (:814).  Returns true when the preamble->sync transition
// was accepted (symbol timings are left in m).
static bool F_track_preamble(Dec *d, const RateParams *b, Mod *m, float s0, float sd, float minimumCorrelationValue, bool ge)
{
   const u32 clk = L.fe.clk;

   if (clk < m->searchStartTime)
      return false;

   if (ge ? (sd >= minimumCorrelationValue) : (sd > minimumCorrelationValue))
   {
      if (sd > m->correlatedPeakValue)
      {
         m->correlatedPeakValue = sd;
         m->correlatedPeakTime = clk;

         if (!m->searchSyncTime)
         {
            m->searchSyncValue = sd;
            m->searchCorr0Value = s0;
            m->searchEndTime = clk + b->p8;
         }
      }
   }

   if (clk == m->searchSyncTime)
   {
      m->searchSyncValue = sd;
      m->searchLastValue = s0;
   }

   if (clk != m->searchEndTime)
      return false;

   if (m->searchPulseWidth++ < 94)
   {
      if (m->correlatedPeakTime == 0 || m->searchSyncValue < m->searchValueThreshold)
      {
         F_restart_search(m);
         return false;
      }
   }

   if (m->searchSyncValue > m->searchValueThreshold)
   {
      if (!m->symbolStartTime)
         m->symbolStartTime = m->correlatedPeakTime - b->p2;

      m->symbolEndTime = m->correlatedPeakTime;
      m->searchSyncTime = m->symbolEndTime + b->p2;
      m->searchStartTime = m->searchSyncTime - b->p8;
      m->searchEndTime = m->searchSyncTime + b->p8;
      m->searchValueThreshold = m->correlatedPeakValue / 2;
      m->searchLastPhase = m->searchLastValue;
      m->correlatedPeakTime = 0;
      m->correlatedPeakValue = 0;
      return false;
   }

   if ((m->searchLastPhase < 0 && m->searchCorr0Value < 0) || (m->searchLastPhase > 0 && m->searchCorr0Value > 0))
      m->symbolStartTime -= b->p2;

   int preambleLength = (int) (m->symbolEndTime - m->symbolStartTime);
   int preambleMinLength = (int) (b->pre1 - b->p4);
   int preambleMaxLength = (int) (b->pre1 + b->p4);

   if (preambleLength < preambleMinLength || preambleLength > preambleMaxLength)
   {
      F_restart_search(m);
      return false;
   }

   m->searchModeState = m->searchLastPhase > 0 ? 0 : 1; // OBSERVED : REVERSED
   m->searchSyncTime = m->searchSyncTime + b->p2;
   m->searchStartTime = m->searchSyncTime - b->p4;
   m->searchEndTime = m->searchSyncTime + b->p4;
   m->correlatedPeakTime = 0;
   m->correlatedPeakValue = 0;
   return true;
}

static void F_correlate(Dec *d, const RateParams *b, Mod *m, u32 fp1, float *s0_, float *s1_, float *sd_)
{
   float s0, s1, sd;
   u32 fp2, fp3;
   corr_points(fp1, b->p1, b->p2, &fp2, &fp3);
   RG(b->corr, fp1) = m->filterIntegrate;
   s0 = m->filterIntegrate - RG(b->corr, fp2);
   s1 = RG(b->corr, fp2) - RG(b->corr, fp3);
   sd = fabsf(s0 - s1) / (float) b->p2;
   *s0_ = s0; *s1_ = s1; *sd_ = sd;
}

// NfcF::Impl::detectModulation, NfcF.cpp:206-408
static bool F_detect(Dec *d)
{
   const u32 clk = L.fe.clk;
   const float minimumCorrelationValue = L.fe.env * P.thr[TECH_F].corr;

   for (int rate = 1; rate <= 2; rate++)
   {
      const RateParams *b = &(P.F[rate]);
      Mod *m = &(L.c.mF[rate - 1]);

      float deep = SMP(NFCB200_OFF_M, b->sdd);

      m->filterIntegrate += SMP(NFCB200_OFF_X, b->sdd);
      m->filterIntegrate -= SMP(NFCB200_OFF_X, b->sdd + b->p2);

      float s0, s1, sd;
      F_correlate(d, b, m, L.fe.cF[rate - 1], &s0, &s1, &sd);

      // :260-271
      if (deep > P.thr[TECH_F].modMax || (m->correlatedPeakTime && clk > m->correlatedPeakTime + b->p1))
      {
         m->symbolStartTime = 0;
         m->symbolEndTime = 0;
         m->searchStartTime = 0;
         m->searchEndTime = 0;
         m->searchSyncTime = 0;
         m->detectorPeakTime = 0;
         m->detectorPeakValue = 0;
         m->correlatedPeakTime = 0;
         m->correlatedPeakValue = 0;
      }

      if (!F_track_preamble(d, b, m, s0, sd, minimumCorrelationValue, false))
         continue;

      L.sym.start = m->symbolStartTime; // :390-402
      L.sym.end = m->symbolEndTime;
      L.sym.length = L.sym.end - L.sym.start;
      L.sym.pattern = F_S;

      FrameSt *fs = &(L.c.t[TECH_F].fs);
      fs->frameType = FT_Poll;
      fs->symbolRate = b->sps;
      fs->frameStart = L.sym.start;
      fs->frameEnd = 0;

      L.lock = LOCK_F;
      L.lockRate = rate;
      return true;
   }

   return false;
}

// NfcF::Impl::checkCrc, NfcF.cpp:1215-1226 (payload only: the two sync bytes were stripped)
static bool F_crc_ok(Dec *d, const u8 *pl, u32 size)
{
   if (size < 2)
      return false;
   unsigned short crc = crc_ccitt16(pl, 0, size - 2, 0x0000, false);
   unsigned short res = (unsigned short) (((pl[size - 2] & 0xff) << 8) | (pl[size - 1] & 0xff));
   return res == crc;
}

// NfcF::Impl::process, NfcF.cpp:1076-1210.  d = payload after the sync bytes
static void F_process(Dec *d, u32 type, const u8 *pl, u32 len, u32 *flags_, u32 *phase_)
{
   u32 flags = *flags_, phase = *phase_;
   TechSt *t = &(L.c.t[TECH_F]);
   FrameSt *fs = &(t->fs);
   Proto *ps = &(t->ps);
   const bool poll = type == FT_Poll;

   if (poll)
   {
      fs->startUpGuardTime = ps->startUpGuardTime;
      fs->frameWaitingTime = ps->frameWaitingTime;
      fs->frameGuardTime = ps->frameGuardTime;
      fs->requestGuardTime = ps->requestGuardTime;
   }
   else
   {
      fs->frameGuardTime = ps->frameGuardTime;
   }

   bool done = false;

   // processREQC :1152-1201
   if (poll)
   {
      u32 b1 = 1 < len ? pl[1] : 0;

      if (b1 == 0x00)
      {
         fs->lastCommand = b1;
         int tsn = (int) (5 < len ? pl[5] : 0);
         ps->maxFrameSize = 256;
         ps->startUpGuardTime = P.F_sfgt;
         ps->frameGuardTime = P.F_fgt;
         ps->frameWaitingTime = P.F_fwt;
         ps->requestGuardTime = P.F_rgt;
         fs->frameGuardTime = (u32) (P.stu * 1024);
         fs->frameWaitingTime = (u32) (P.stu * (512 * 64 + (tsn + 1) * (256 * 64)));
         t->chained = 0;
         phase = PH_Selection;
         flags |= !F_crc_ok(d, pl, len) ? FL_Crc : 0;
         done = true;
      }
   }
   else if (fs->lastCommand == 0x00)
   {
      phase = PH_Selection;
      flags |= !F_crc_ok(d, pl, len) ? FL_Crc : 0;
      done = true;
   }

   if (!done) // processOther
   {
      phase = PH_Application;
      flags |= !F_crc_ok(d, pl, len) ? FL_Crc : 0;
   }

   flags |= t->chained;

   if (poll)
   {
      if (L.lock == LOCK_F)
      {
         u32 sdd = P.F[L.lockRate].sdd;
         fs->guardEnd = fs->frameEnd + fs->frameGuardTime + sdd;
         fs->waitingEnd = fs->frameEnd + fs->frameWaitingTime + sdd;
         fs->frameType = FT_Listen;
      }
   }
   else
   {
      if (L.lock == LOCK_F)
         fs->guardEnd = fs->frameEnd + fs->frameGuardTime + P.F[L.lockRate].sdd;
      fs->frameType = 0;
      fs->lastCommand = 0;
   }

   fs->frameStart = 0;
   fs->frameEnd = 0;
   *flags_ = flags; *phase_ = phase;
}

// one sample of decodePollFrameSymbolAsk / decodeListenFrameSymbolAsk, NfcF.cpp:641-744, 941-1042 (identical bodies)
static int F_symbol(Dec *d)
{
   const RateParams *b = &(P.F[L.lockRate]);
   Mod *m = &(L.c.mF[L.lockRate - 1]);
   const u32 clk = L.fe.clk;

   m->filterIntegrate += SMP(NFCB200_OFF_X, b->sdd);
   m->filterIntegrate -= SMP(NFCB200_OFF_X, b->sdd + b->p2);

   float s0, s1, sd;
   F_correlate(d, b, m, L.fe.cF[L.lockRate - 1], &s0, &s1, &sd);

   if (clk < m->searchStartTime)
      return F_Invalid;

   if (sd > m->searchValueThreshold && sd > m->correlatedPeakValue)
   {
      m->correlatedPeakValue = sd;
      m->correlatedPeakTime = clk;
   }

   if (clk == m->searchSyncTime)
   {
      m->searchCorr0Value = s0;
      m->searchCorr1Value = s1;
   }

   if (clk != m->searchEndTime)
      return F_Invalid;

   if (!m->correlatedPeakTime)
      return F_E;

   m->symbolStartTime = m->symbolEndTime;
   m->symbolEndTime = m->correlatedPeakTime;
   m->searchSyncTime = m->symbolEndTime + b->p1;
   m->searchStartTime = m->searchSyncTime - b->p4;
   m->searchEndTime = m->searchSyncTime + b->p4;
   m->searchValueThreshold = m->correlatedPeakValue / 2;
   m->correlatedPeakTime = 0;
   m->correlatedPeakValue = 0;

   L.sym.start = m->symbolStartTime - b->sdd;
   L.sym.end = m->symbolEndTime - b->sdd;
   L.sym.length = L.sym.end - L.sym.start;

   if ((m->searchModeState == 0 && m->searchCorr0Value > m->searchCorr1Value) || (m->searchModeState == 1 && m->searchCorr0Value < m->searchCorr1Value))
   {
      L.sym.value = 0;
      L.sym.pattern = F_L;
   }
   else
   {
      L.sym.value = 1;
      L.sym.pattern = F_H;
   }

   return (int) L.sym.pattern;
}

// frame assembly shared by decodePollFrame (NfcF.cpp:428-529) and decodeListenFrame (:534-636)
static void F_frame_step(Dec *d, int pattern, u32 type)
{
   TechSt *t = &(L.c.t[TECH_F]);
   Bits *st = &(L.st);
   bool frameEnd = false, truncateError = false;

   if (pattern == F_E)
      frameEnd = true;
   else if (st->bytes == t->ps.maxFrameSize)
      truncateError = true;

   if (frameEnd || truncateError)
   {
      if (st->bytes > 2)
      {
         t->fs.frameEnd = L.sym.end;

         u32 flags = 0, phase = 0;
         if (truncateError)
            flags |= FL_Truncated;
         if (sb[0] != 0xB2 || sb[1] != 0x4D)
            flags |= FL_Sync;

         u32 total = st->bytes > 512 ? 512 : st->bytes;
         u32 len = total - 2, rate = P.F[L.lockRate].sps, start = t->fs.frameStart, end = t->fs.frameEnd;

         F_process(d, type, sb + 2, len, &flags, &phase);
         emit(d, TT_F, type, flags, phase, rate, start, end, sb + 2, len);

         if (type == FT_Poll)
         {
            clear_bits(d);

            if (L.lock == LOCK_F)
               clear_for_listen(d, &L.c.mF[L.lockRate - 1], P.F[L.lockRate].corr, P.F[L.lockRate].p1);

            return;
         }
      }

      F_reset(d);
      return;
   }

   st->data = (st->data << 1) | L.sym.value;

   if (++st->bits == 8)
   {
      put_byte(d, st->data);
      st->data = 0;
      st->bits = 0;
   }
}

// one sample of decodeListenFrameStartAsk, NfcF.cpp:749-936
static int F_listen_start(Dec *d)
{
   const RateParams *b = &(P.F[L.lockRate]);
   Mod *m = &(L.c.mF[L.lockRate - 1]);
   FrameSt *fs = &(L.c.t[TECH_F].fs);
   const u32 clk = L.fe.clk;

   m->filterIntegrate += SMP(NFCB200_OFF_X, b->sdd);
   m->filterIntegrate -= SMP(NFCB200_OFF_X, b->sdd + b->p2);

   if (clk < (fs->guardEnd - b->p1))
      return F_Invalid;

   float s0, s1, sd;
   F_correlate(d, b, m, L.fe.cF[L.lockRate - 1], &s0, &s1, &sd);

   if (clk < fs->guardEnd)
      return F_Invalid;

   if (clk == fs->guardEnd)
      m->searchValueThreshold = SMP(NFCB200_OFF_D, b->sdd) * 10;

   if (clk > fs->waitingEnd)
      return F_No;

   // note: the threshold passed for the `>=` test is searchValueThreshold itself (:814)
   if (!F_track_preamble(d, b, m, s0, sd, m->searchValueThreshold, true))
      return F_Invalid;

   L.sym.start = m->symbolStartTime - b->sdd;
   L.sym.end = m->symbolEndTime - b->sdd;
   L.sym.length = L.sym.end - L.sym.start;
   L.sym.pattern = F_S;
   return F_S;
}

static void F_poll_step(Dec *d)
{
   int pattern = F_symbol(d);
   if (pattern > F_No)
      F_frame_step(d, pattern, FT_Poll);
}

static void F_listen_step(Dec *d)
{
   FrameSt *fs = &(L.c.t[TECH_F].fs);

   if (!fs->frameStart)
   {
      int pattern = F_listen_start(d);

      if (pattern == F_S)
         fs->frameStart = L.sym.start;
      else if (pattern == F_No)
         F_reset(d);

      return;
   }

   int pattern = F_symbol(d);
   if (pattern > F_No)
      F_frame_step(d, pattern, FT_Listen);
}

// ------------------------------------------------------------------------------------------------------------------
// NFC-V
// ------------------------------------------------------------------------------------------------------------------
enum { V_Invalid = 0, V_No = 1, V_0 = 2, V_1 = 3, V_2 = 4, V_8 = 5, V_S = 6, V_E = 7 };

// NfcV::Impl::resetModulation, NfcV.cpp:1079-1103
static void V_reset(Dec *d)
{
   clear_bits(d);
   clear_sym(d);
   zero_mod(&L.c.mV);
   zero_ring(d, P.V.corr, P.V.p0 > P.V.p1 ? P.V.p0 : P.V.p1);
   zero_ring(d, NFCB200_OFF_I, NFCB200_RING);
   L.c.t[TECH_V].fs.frameType = 0;
   L.c.t[TECH_V].fs.frameStart = 0;
   L.c.t[TECH_V].fs.frameEnd = 0;
   L.pulseBits = 0;
   L.lock = LOCK_NONE;
}

static void V_clear_search(Mod *m)
{
   m->symbolStartTime = 0;
   m->symbolEndTime = 0;
   m->searchStartTime = 0;
   m->searchEndTime = 0;
   m->correlatedPeakTime = 0;
   m->correlatedPeakValue = 0;
   m->detectorPeakTime = 0;
   m->detectorPeakValue = 0;
}

// half-symbol pulse correlator of detectModulation / decodePollFrameSymbolPpm (NfcV.cpp:258-274, 688-703)
static float V_pulse_corr(Dec *d, Mod *m, float *signalData_)
{
   float signalData = 0;
   const RateParams *b = &(P.V);
   u32 fp1 = L.fe.cV1;
   u32 fp2 = fp1 + b->p2;
   if (fp2 >= b->p1)
      fp2 -= b->p1;

   signalData = SMP(NFCB200_OFF_X, b->sdd);

   m->filterIntegrate += signalData;
   m->filterIntegrate -= SMP(NFCB200_OFF_X, b->sdd + b->p2);

   RG(b->corr, fp1) = m->filterIntegrate;

   { *signalData_ = signalData; return (RG(b->corr, fp2) - m->filterIntegrate) / (float) b->p2; }
   *signalData_ = signalData;
}

// NfcV::Impl::detectModulation, NfcV.cpp:236-435
static bool V_detect(Dec *d)
{
   const RateParams *b = &(P.V);
   Mod *m = &(L.c.mV);
   const u32 clk = L.fe.clk;
   const float minimumCorrelationValue = L.fe.env * P.thr[TECH_V].corr;

   float signalData;
   float s0 = V_pulse_corr(d, m, &signalData);
   float deep = SMP(NFCB200_OFF_M, b->sdd + b->p8);

   if (m->correlatedPeakTime && clk > m->correlatedPeakTime + b->p0) // :287-298
   {
      m->symbolStartTime = 0;
      m->symbolEndTime = 0;
      m->searchStartTime = 0;
      m->searchEndTime = 0;
      m->searchSyncTime = 0;
      m->detectorPeakTime = 0;
      m->detectorPeakValue = 0;
      m->correlatedPeakTime = 0;
      m->correlatedPeakValue = 0;
   }

   if (clk < m->searchStartTime)
      return false;

   if (s0 > minimumCorrelationValue)
   {
      if (s0 > m->correlatedPeakValue)
      {
         m->correlatedPeakValue = s0;
         m->correlatedPeakTime = clk;
         m->searchEndTime = clk + b->p4;
      }

      if (deep > m->detectorPeakValue)
      {
         m->detectorPeakValue = deep;
         m->detectorPeakTime = clk;
      }
   }

   if (clk != m->searchEndTime)
      return false;

   if (signalData < minimumCorrelationValue || m->correlatedPeakTime == 0 || m->detectorPeakValue < P.thr[TECH_V].modMin)
   {
      V_clear_search(m);
      return false;
   }

   if (!m->symbolStartTime) // :345-359
   {
      m->symbolStartTime = m->correlatedPeakTime - b->p2;
      m->searchStartTime = m->symbolStartTime + (2 * b->p1);
      m->searchEndTime = m->symbolStartTime + (4 * b->p1);
      m->correlatedPeakTime = 0;
      m->correlatedPeakValue = 0;
      m->detectorPeakTime = 0;
      m->detectorPeakValue = 0;
      return false;
   }

   FrameSt *fs = &(L.c.t[TECH_V].fs);

   if (m->correlatedPeakTime > (m->symbolStartTime + 3 * b->p1 - b->p8) && m->correlatedPeakTime < (m->symbolStartTime + 3 * b->p1 + b->p8))
   {
      m->symbolEndTime = m->correlatedPeakTime + b->p1;
      m->searchSyncTime = m->symbolEndTime;
      m->searchStartTime = m->searchSyncTime;
      m->searchEndTime = m->searchSyncTime + P.V_len2;
      fs->symbolRate = b->sps / 2;
      L.pulseBits = 2;
   }
   else if (m->correlatedPeakTime > (m->symbolStartTime + 4 * b->p1 - b->p8) && m->correlatedPeakTime < (m->symbolStartTime + 4 * b->p1 + b->p8))
   {
      m->symbolEndTime = m->correlatedPeakTime;
      m->searchSyncTime = m->symbolEndTime;
      m->searchStartTime = m->searchSyncTime;
      m->searchEndTime = m->searchSyncTime + P.V_len8;
      fs->symbolRate = b->sps / 32;
      L.pulseBits = 8;
   }
   else
   {
      V_clear_search(m);
      return false;
   }

   fs->frameType = FT_Poll;
   fs->frameStart = m->symbolStartTime - b->sdd;
   fs->frameEnd = 0;

   m->correlatedPeakTime = 0;
   m->correlatedPeakValue = 0;
   m->searchValueThreshold = minimumCorrelationValue;

   L.lock = LOCK_V;
   L.lockRate = 0;
   return true;
}

// NfcV::Impl::checkCrc, NfcV.cpp:1194-1205
static bool V_crc_ok(Dec *d, u32 size)
{
   if (size < 3)
      return false;
   unsigned short crc = (unsigned short) ~crc_ccitt16(sb, 0, size - 2, 0xFFFF, true);
   unsigned short res = (unsigned short) ((sb[size - 2] & 0xff) | ((sb[size - 1] & 0xff) << 8));
   return res == crc;
}

// NfcV::Impl::process, NfcV.cpp:1108-1189
static void V_process(Dec *d, u32 type, u32 len, u32 *flags_, u32 *phase_)
{
   u32 flags = *flags_, phase = *phase_;
   TechSt *t = &(L.c.t[TECH_V]);
   FrameSt *fs = &(t->fs);
   const bool poll = type == FT_Poll;

   if (poll)
   {
      fs->frameGuardTime = t->ps.frameGuardTime;
      fs->frameWaitingTime = t->ps.frameWaitingTime;
   }
   else
   {
      fs->frameGuardTime = t->ps.frameGuardTime;
   }

   phase = PH_Application;
   flags |= !V_crc_ok(d, len) ? FL_Crc : 0;
   flags |= t->chained;

   if (poll)
   {
      if (L.lock == LOCK_V)
      {
         fs->guardEnd = fs->frameEnd + fs->frameGuardTime - P.V.sdd; // minus: NfcV.cpp:1147-1150
         fs->waitingEnd = fs->frameEnd + fs->frameWaitingTime - P.V.sdd;
         fs->frameType = FT_Listen;
      }
   }
   else
   {
      if (L.lock == LOCK_V)
         fs->guardEnd = fs->frameEnd + fs->frameGuardTime + P.V.sdd;
      fs->frameType = 0;
      fs->lastCommand = 0;
   }

   fs->frameStart = 0;
   fs->frameEnd = 0;
   *flags_ = flags; *phase_ = phase;
}

// one sample of decodePollFrameSymbolPpm, NfcV.cpp:672-795
static int V_poll_symbol(Dec *d)
{
   const RateParams *b = &(P.V);
   Mod *m = &(L.c.mV);
   const u32 clk = L.fe.clk;

   float signalData;
   float s0 = V_pulse_corr(d, m, &signalData);

   if (clk < m->searchStartTime)
      return V_Invalid;

   if (s0 > m->searchValueThreshold)
   {
      if (s0 > m->correlatedPeakValue)
      {
         m->correlatedPeakValue = s0;
         m->correlatedPeakTime = clk;
         m->searchEndTime = clk + b->p4;
      }
   }

   if (clk != m->searchEndTime)
      return V_Invalid;

   // EOF: pulse in the first half of the second slot (:734-751)
   if (m->correlatedPeakTime > (m->searchStartTime + 1 * b->p1 + b->p4) && m->correlatedPeakTime < (m->searchStartTime + 2 * b->p1 - b->p4))
   {
      m->symbolEndTime = m->correlatedPeakTime + b->p2;
      L.sym.value = 0;
      L.sym.start = m->symbolStartTime - b->sdd;
      L.sym.end = m->symbolEndTime - b->sdd;
      L.sym.length = L.sym.end - L.sym.start;
      L.sym.pattern = V_S;
      return V_S;
   }

   L.sym.value = 0;
   L.sym.start = m->symbolStartTime - b->sdd;
   L.sym.end = m->symbolEndTime - b->sdd;
   L.sym.length = L.sym.end - L.sym.start;
   L.sym.pattern = V_E;

   const u32 periods = 1u << L.pulseBits;
   const u32 length = L.pulseBits == 2 ? P.V_len2 : P.V_len8;

   // slot search (:761-789).  slot->end = round((i + 1) * stu * 256), NfcV.cpp:228-232
   for (u32 i = 0; i < periods; i++)
   {
      u32 slotEnd = (u32) (int) round((double) (i + 1) * P.stu * 256);

      if (m->correlatedPeakTime > (m->searchStartTime + slotEnd - b->p4) && m->correlatedPeakTime < (m->searchStartTime + slotEnd + b->p4))
      {
         m->symbolStartTime = m->correlatedPeakTime - slotEnd;
         m->symbolEndTime = m->symbolStartTime + length;
         m->searchSyncTime = m->symbolEndTime;
         m->searchStartTime = m->searchSyncTime;
         m->searchEndTime = m->searchSyncTime + length;
         m->correlatedPeakTime = 0;
         m->correlatedPeakValue = 0;

         L.sym.value = i;
         L.sym.start = m->symbolStartTime - b->sdd;
         L.sym.end = m->symbolEndTime - b->sdd;
         L.sym.length = L.sym.end - L.sym.start;
         L.sym.pattern = L.pulseBits == 2 ? V_2 : V_8;
         return (int) L.sym.pattern;
      }
   }

   return V_E;
}

// frame assembly of decodePollFrame (NfcV.cpp:450-556) / decodeListenFrame (:561-667)
static void V_frame_step(Dec *d, int pattern, u32 type)
{
   TechSt *t = &(L.c.t[TECH_V]);
   Bits *st = &(L.st);
   bool frameEnd = false, truncateError = false, streamError = false;

   if (pattern == V_S)
      frameEnd = true;
   else if (pattern == V_E)
      streamError = true;
   else if (st->bytes == t->ps.maxFrameSize)
      truncateError = true;

   if (frameEnd || streamError || truncateError)
   {
      if (st->bytes > 0)
      {
         if (st->bits == 8)
            put_byte(d, st->data);

         t->fs.frameEnd = L.sym.end;

         u32 flags = 0, phase = 0;
         if (truncateError || streamError)
            flags |= FL_Truncated;

         u32 len = st->bytes, rate = t->fs.symbolRate, start = t->fs.frameStart, end = t->fs.frameEnd;

         V_process(d, type, len, &flags, &phase);
         emit(d, TT_V, type, flags, phase, rate, start, end, sb, len);

         if (type == FT_Poll)
         {
            clear_bits(d);

            if (L.lock == LOCK_V)
               clear_for_listen(d, &L.c.mV, P.V.corr, P.V.p0 > P.V.p1 ? P.V.p0 : P.V.p1);

            return;
         }
      }

      V_reset(d);
      return;
   }

   if (st->bits == 8)
   {
      put_byte(d, st->data);
      st->data = 0;
      st->bits = 0;
   }

   st->data |= (L.sym.value << st->bits);
   st->bits += (type == FT_Poll) ? L.pulseBits : 1;
}

// full-symbol w^2 * 10 correlator of the NFC-V listen decoders (NfcV.cpp:817-835, 1000-1018)
static float V_listen_corr(Dec *d, Mod *m)
{
   const RateParams *b = &(P.V);
   u32 fp1 = L.fe.cV0;
   u32 fp2 = fp1 + b->p1;
   if (fp2 >= b->p0)
      fp2 -= b->p0;

   float data = SMP(NFCB200_OFF_W, b->sdd);
   float v = data * data * 10;
   SMP(NFCB200_OFF_I, b->sdd) = v;

   m->filterIntegrate += v;
   m->filterIntegrate -= SMP(NFCB200_OFF_I, b->sdd + b->p1);

   RG(b->corr, fp1) = m->filterIntegrate;

   return RG(b->corr, fp2) - m->filterIntegrate;
}

// one sample of decodeListenFrameStartAsk, NfcV.cpp:800-980
static int V_listen_start(Dec *d)
{
   const RateParams *b = &(P.V);
   Mod *m = &(L.c.mV);
   FrameSt *fs = &(L.c.t[TECH_V].fs);
   const u32 clk = L.fe.clk;

   float s0 = V_listen_corr(d, m);
   float deep = SMP(NFCB200_OFF_M, 0);

   if (clk < fs->guardEnd)
      return V_Invalid;

   if (clk == fs->guardEnd)
      m->searchValueThreshold = SMP(NFCB200_OFF_D, b->sdd);

   if (clk > fs->waitingEnd)
      return V_No;

   if (deep > P.thr[TECH_V].modMax)
      return V_No;

   if (clk < m->searchStartTime)
      return V_Invalid;

   if (s0 < -m->searchValueThreshold && s0 < m->correlatedPeakValue)
   {
      m->correlatedPeakValue = s0;
      m->correlatedPeakTime = clk;
      m->searchEndTime = clk + b->p8;
   }

   if (s0 > m->searchValueThreshold && s0 > m->correlatedPeakValue)
   {
      m->correlatedPeakValue = s0;
      m->correlatedPeakTime = clk;
      m->searchEndTime = clk + b->p8;
   }

   if (clk != m->searchEndTime)
      return V_Invalid;

   if (m->searchModeState == 0) // LISTEN_MODE_PREAMBLE1
   {
      if (!m->symbolStartTime)
      {
         m->symbolStartTime = m->correlatedPeakTime - b->p1;
         m->searchStartTime = m->correlatedPeakTime + b->p0;
         m->searchEndTime = m->searchStartTime + b->p1;
         m->correlatedPeakValue = 0;
         m->correlatedPeakTime = 0;
         return V_Invalid;
      }

      m->symbolEndTime = m->correlatedPeakTime;

      u32 preambleS1Length = m->symbolEndTime - m->symbolStartTime - b->p1; // int vs unsigned: compared unsigned

      if (m->correlatedPeakTime == 0 || preambleS1Length < P.V_s1Min || preambleS1Length > P.V_s1Max)
      {
         m->searchModeState = 0;
         m->searchStartTime = 0;
         m->searchEndTime = 0;
         m->symbolStartTime = 0;
         m->symbolEndTime = 0;
         return V_Invalid;
      }

      m->searchModeState = 1;
      m->searchStartTime = m->correlatedPeakTime + b->p1 - b->p2;
      m->searchEndTime = m->searchStartTime + b->p1;
      m->correlatedPeakValue = 0;
      m->correlatedPeakTime = 0;
      return V_Invalid;
   }

   if (m->searchModeState == 1) // LISTEN_MODE_PREAMBLE2
   {
      u32 preambleS2Length = m->correlatedPeakTime - m->symbolEndTime;

      if (m->correlatedPeakTime == 0 || preambleS2Length < P.V_s2Min || preambleS2Length > P.V_s2Max)
      {
         m->searchModeState = 0;
         m->searchStartTime = 0;
         m->searchEndTime = 0;
         m->symbolStartTime = 0;
         m->symbolEndTime = 0;
         return V_Invalid;
      }

      m->symbolEndTime = m->correlatedPeakTime;
      m->searchSyncTime = m->symbolEndTime + b->p0;
      m->searchStartTime = m->searchSyncTime - b->p4;
      m->searchEndTime = m->searchSyncTime + b->p4;
      m->searchValueThreshold = (float) (m->correlatedPeakValue * 0.25);
      m->searchCorr0Value = 0;
      m->searchCorr1Value = 0;
      m->correlatedPeakTime = 0;
      m->correlatedPeakValue = 0;

      L.sym.value = 0;
      L.sym.start = m->symbolStartTime - b->sdd;
      L.sym.end = m->symbolEndTime - b->sdd;
      L.sym.length = L.sym.end - L.sym.start;
      L.sym.pattern = V_S;
      return V_S;
   }

   return V_Invalid;
}

// one sample of decodeListenFrameSymbolAsk, NfcV.cpp:985-1074
static int V_listen_symbol(Dec *d)
{
   const RateParams *b = &(P.V);
   Mod *m = &(L.c.mV);
   const u32 clk = L.fe.clk;

   float s0 = V_listen_corr(d, m);
   float sd = fabsf(s0);

   if (clk < m->searchStartTime)
      return V_Invalid;

   if (sd > m->searchValueThreshold && sd > m->correlatedPeakValue)
   {
      m->searchCorr0Value = s0;
      m->searchCorr1Value = -s0;
      m->correlatedPeakValue = sd;
      m->symbolEndTime = clk;
   }

   if (clk != m->searchEndTime)
      return V_Invalid;

   if (m->correlatedPeakValue < m->searchValueThreshold)
      return V_S;

   m->symbolStartTime = m->symbolEndTime;
   m->symbolEndTime = m->symbolStartTime + b->p0;
   m->searchSyncTime = m->symbolEndTime;
   m->searchStartTime = m->searchSyncTime - b->p4;
   m->searchEndTime = m->searchSyncTime + b->p4;
   m->searchValueThreshold = (float) (m->correlatedPeakValue * 0.25);
   m->correlatedPeakTime = 0;
   m->correlatedPeakValue = 0;

   L.sym.value = m->searchCorr0Value > m->searchCorr1Value ? 0 : 1;
   L.sym.start = m->symbolStartTime - b->sdd;
   L.sym.end = m->symbolEndTime - b->sdd;
   L.sym.length = L.sym.end - L.sym.start;
   L.sym.pattern = L.sym.value ? V_1 : V_0;

   return (int) L.sym.pattern;
}

static void V_poll_step(Dec *d)
{
   int pattern = V_poll_symbol(d);
   if (pattern > V_No)
      V_frame_step(d, pattern, FT_Poll);
}

static void V_listen_step(Dec *d)
{
   FrameSt *fs = &(L.c.t[TECH_V].fs);

   if (!fs->frameStart)
   {
      int pattern = V_listen_start(d);

      if (pattern == V_S)
         fs->frameStart = L.sym.start;
      else if (pattern == V_No)
         V_reset(d);

      return;
   }

   int pattern = V_listen_symbol(d);
   if (pattern > V_No)
      V_frame_step(d, pattern, FT_Listen);
}

// ------------------------------------------------------------------------------------------------------------------
// dispatch: NfcDecoder::Impl::nextFrames inner loops, NfcDecoder.cpp:393-442, one sample per call
// ------------------------------------------------------------------------------------------------------------------
static void step(Dec *d, float x)
{
   front(d, x);

   if (L.lock == LOCK_NONE)
   {
      if (L.fe.k > L.warm)
         detect_carrier(d);

      // `signalClock < BUFFER_SIZE` and `signalEnvelope < powerLevelThreshold` gates of every detectModulation
      if (L.fe.k - 1 < L.gate || L.fe.env < P.power)
         return;

      if ((P.enabled & EN_A) && A_detect(d))
         return;
      if ((P.enabled & EN_B) && B_detect(d))
         return;
      if ((P.enabled & EN_F) && F_detect(d))
         return;
      if ((P.enabled & EN_V) && V_detect(d))
         return;

      return;
   }

   u32 frameType;

   switch (L.lock)
   {
      case LOCK_A:
         frameType = L.c.t[TECH_A].fs.frameType;
         if (frameType == FT_Poll)
            A_poll_step(d);
         else if (frameType == FT_Listen)
            A_listen_step(d);
         break;
      case LOCK_B:
         frameType = L.c.t[TECH_B].fs.frameType;
         if (frameType == FT_Poll)
            B_poll_step(d);
         else if (frameType == FT_Listen)
            B_listen_step(d);
         break;
      case LOCK_F:
         frameType = L.c.t[TECH_F].fs.frameType;
         if (frameType == FT_Poll)
            F_poll_step(d);
         else if (frameType == FT_Listen)
            F_listen_step(d);
         break;
      default:
         frameType = L.c.t[TECH_V].fs.frameType;
         if (frameType == FT_Poll)
            V_poll_step(d);
         else if (frameType == FT_Listen)
            V_listen_step(d);
         break;
   }
}


/* ---- decoder life cycle: NfcDecoder::Impl::initialize / NfcX::initialize ------------------------------------------- */
static void decoder_initialise(Dec *d, u32 sampleRate)
{
   u32 enabled = d->enabledCfg;
   u32 streamTime = P.streamTime;
   float power = P.power;
   TechThresholds thr[4];
   memcpy(thr, P.thr, sizeof(thr));
   memset(&P, 0, sizeof(P));
   P.enabled = enabled;
   P.streamTime = streamTime;
   P.power = power;
   memcpy(P.thr, thr, sizeof(thr));
   params_init(&P, sampleRate);

   memset(&L, 0, sizeof(L));
   memset(rg, 0, sizeof(rg));
   memset(sb, 0, sizeof(sb));

   const u32 def[4][4] = {
      {P.A_sfgt, P.A_fgt, P.A_fwt, P.A_rgt}, {P.B_sfgt, P.B_fgt, P.B_fwt, P.B_rgt}, {P.F_sfgt, P.F_fgt, P.F_fwt, P.F_rgt}, {P.V_sfgt, P.V_fgt, P.V_fwt, P.V_rgt}};

   for (int t = 0; t < 4; t++) /* NfcA.cpp:195-205 and twins */
   {
      L.c.t[t].ps.maxFrameSize = 256;
      L.c.t[t].ps.startUpGuardTime = L.c.t[t].fs.startUpGuardTime = def[t][0];
      L.c.t[t].ps.frameGuardTime = L.c.t[t].fs.frameGuardTime = def[t][1];
      L.c.t[t].ps.frameWaitingTime = L.c.t[t].fs.frameWaitingTime = def[t][2];
      L.c.t[t].ps.requestGuardTime = L.c.t[t].fs.requestGuardTime = def[t][3];
   }

   L.fe.clk = (u32) -1; /* signalClock = -1, NfcTech.h:338 */
   for (int r = 0; r < 3; r++)
      L.fe.cA[r] = P.A[r].c1 ? P.A[r].c1 - 1 : P.A[r].p1 - 1;
   for (int r = 0; r < 2; r++)
      L.fe.cF[r] = P.F[r + 1].c1 ? P.F[r + 1].c1 - 1 : P.F[r + 1].p1 - 1;
   L.fe.cV1 = P.V.c1 ? P.V.c1 - 1 : P.V.p1 - 1;
   L.fe.cV0 = P.V.c0 ? P.V.c0 - 1 : P.V.p0 - 1;
   L.gate = NFCB200_RING;
   d->initialised = 1;
}

#undef P
#undef L
#undef rg
#undef sb

/* ---- public API (oracle/nfc_oracle.h) ----------------------------------------------------------------------------- */
nfcoracle_decoder *nfcoracle_create(void)
{
   Dec *d = (Dec *) calloc(1, sizeof(Dec));
   d->enabledCfg = EN_A | EN_B | EN_F | EN_V;
   d->P.power = 0.01f;                                                  /* NfcTech.h:347 */
   d->P.thr[TECH_A] = (TechThresholds) {0.75f, 0.90f, 1.00f};            /* NfcA.cpp:94-100 */
   d->P.thr[TECH_B] = (TechThresholds) {0.50f, 0.10f, 0.90f};            /* NfcB.cpp:103-109 */
   d->P.thr[TECH_F] = (TechThresholds) {0.50f, 0.10f, 0.90f};            /* NfcF.cpp:88-94 */
   d->P.thr[TECH_V] = (TechThresholds) {0.50f, 0.90f, 1.00f};            /* NfcV.cpp:101-107 */
   return d;
}

void nfcoracle_destroy(nfcoracle_decoder *d)
{
   free(d);
}

void nfcoracle_set_enabled(nfcoracle_decoder *d, unsigned mask)
{
   d->enabledCfg = mask & 0xF;
   d->P.enabled = mask & 0xF;
}

long nfcoracle_push(nfcoracle_decoder *d, const float *mag, uint64_t n, uint32_t sample_rate, nfcoracle_frame *out, long cap)
{
   if (!d->initialised || d->P.sampleRate != sample_rate) /* NfcDecoder.cpp:383-388 */
      decoder_initialise(d, sample_rate);

   d->out = out;
   d->cap = out ? cap : 0;
   d->nframes = 0;

   for (uint64_t i = 0; i < n; i++)
      step(d, mag[i]);

   return d->nframes;
}

long nfcoracle_decode(const float *mag, uint64_t n, uint32_t sample_rate, unsigned enabled_mask, nfcoracle_frame *out, long cap)
{
   nfcoracle_decoder *d = nfcoracle_create();
   nfcoracle_set_enabled(d, enabled_mask);
   long produced = nfcoracle_push(d, mag, n, sample_rate, out, cap);
   nfcoracle_destroy(d);
   return produced;
}

/* IQ -> magnitude, scalar path of RadioDeviceTask::processQueue (lab-tasks RadioDeviceTask.cpp:627-637) */
void nfcoracle_iq_magnitude(const float *iq, uint64_t n, float *mag)
{
   for (uint64_t i = 0; i < n; i++)
   {
      float I = iq[2 * i + 0];
      float Q = iq[2 * i + 1];
      mag[i] = sqrtf(I * I + Q * Q);
   }
}
