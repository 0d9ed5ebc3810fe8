/*
 * nfc_params.h -- decoder parameter block (host-computed once per sample rate, read-only on the device).
 *
 * Restates the parameter derivation of the reference:
 *   NfcDecoder::Impl::initialize      lab-radio/src/main/cpp/NfcDecoder.cpp:295-360
 *   NfcA::Impl::initialize            lab-radio/src/main/cpp/tech/NfcA.cpp:115-212
 *   NfcB::Impl::initialize            lab-radio/src/main/cpp/tech/NfcB.cpp:124-233
 *   NfcF::Impl::initialize            lab-radio/src/main/cpp/tech/NfcF.cpp:106-204
 *   NfcV::Impl::initialize            lab-radio/src/main/cpp/tech/NfcV.cpp:119-234
 *   protocol constants                lab-radio/src/main/include/lab/nfc/Nfc.h:27-169
 */
#ifndef NFCB200_PARAMS_H
#define NFCB200_PARAMS_H

#include <stdint.h>
#include <math.h>

namespace nfcb200 {

typedef uint32_t u32;
typedef uint8_t u8;

enum { TECH_A = 0, TECH_B = 1, TECH_F = 2, TECH_V = 3 };

enum { EN_A = 1, EN_B = 2, EN_F = 4, EN_V = 8 };

// frame constants (lab-data RawFrame.h:29-86)
enum { FT_CarrierOff = 0x0100, FT_CarrierOn = 0x0101, FT_Poll = 0x0102, FT_Listen = 0x0103 };
enum { TT_Any = 0x0100, TT_A = 0x0101, TT_B = 0x0102, TT_F = 0x0103, TT_V = 0x0104 };
enum { PH_Carrier = 0x0101, PH_Selection = 0x0102, PH_Application = 0x0103 };
enum { FL_Short = 0x01, FL_Encrypted = 0x02, FL_Truncated = 0x08, FL_Parity = 0x10, FL_Crc = 0x20, FL_Sync = 0x40 };

// sample ring length of the reference (NfcTech.h:40)
#define NFCB200_RING 1024

// per-lane scratch layout, in floats
#define NFCB200_OFF_X 0
#define NFCB200_OFF_W 1024
#define NFCB200_OFF_D 2048
#define NFCB200_OFF_M 3072
#define NFCB200_OFF_I 4096
#define NFCB200_OFF_CA 5120          /* 3 x 192 : NFC-A correlation rings (period1 <= 192)  */
#define NFCB200_CA_LEN 192
#define NFCB200_OFF_CF (5120 + 576)  /* 2 x 96  : NFC-F correlation rings (period1 <= 96)   */
#define NFCB200_CF_LEN 96
#define NFCB200_OFF_CV (5120 + 768)  /* 1 x 768 : NFC-V correlation ring  (period0 <= 768)  */
#define NFCB200_CV_LEN 768
#define NFCB200_SCRATCH_FLOATS (5120 + 768 + 768)

struct RateParams
{
   u32 p0, p1, p2, p4, p8; // symbol periods in samples (NfcTech.h:174-179)
   u32 sdd;                // symbolDelayDetect
   u32 sps;                // symbolsPerSecond
   u32 pre1;               // preamble1Samples (NFC-F)
   u32 c1;                 // (BUFFER_SIZE - sdd) % p1 : correlation ring phase at local step 0
   u32 c0;                 // (BUFFER_SIZE - sdd) % p0
   u32 corr;               // scratch offset of the correlation ring
};

struct TechThresholds
{
   float corr;   // correlationThreshold
   float modMin; // minimumModulationDeep
   float modMax; // maximumModulationDeep
};

struct Params
{
   u32 sampleRate;
   u32 enabled;
   u32 streamTime;
   int etu;           // elementaryTimeUnit (int, truncated)
   double stu;        // sampleTimeUnit
   float iirA;
   float envW0, envW1;
   float mdevW0, mdevW1;
   float meanW0, meanW1;
   float power, lowThr, highThr;

   RateParams A[3];
   RateParams B[3];
   RateParams F[3]; // index by rate type, [0] unused
   RateParams V;

   TechThresholds thr[4];

   // default protocol timings (samples)
   u32 A_sfgt, A_fgt, A_fwt, A_rgt, A_fwtAtqa, fwtActivation;
   u32 B_sfgt, B_fgt, B_fwt, B_rgt, B_tr0min, B_fwtAtqb;
   u32 B_tr1Min, B_tr1Max, B_s1Min, B_s1Max, B_s2Min, B_s2Max, B_eofComp;
   u32 F_sfgt, F_fgt, F_fwt, F_rgt;
   u32 V_sfgt, V_fgt, V_fwt, V_rgt, V_s1Min, V_s1Max, V_s2Min, V_s2Max;
   u32 V_len2, V_len8; // pulse->length for 1-of-4 / 1-of-256

   int valid; // 0 when the sample rate cannot be represented by the scratch layout
};

// Nfc.h:45-52 tables (functions, so that device code can index them)
#if defined(__CUDACC__)
#define NFC_TAB_HD __host__ __device__ inline
#else
#define NFC_TAB_HD inline
#endif

// NFC_FDS_TABLE: FSDI -> frame size
NFC_TAB_HD int nfc_fds_table(int i)
{
   switch (i & 15)
   {
      case 0: return 16;
      case 1: return 24;
      case 2: return 32;
      case 3: return 40;
      case 4: return 48;
      case 5: return 64;
      case 6: return 96;
      case 7: return 128;
      case 8: return 256;
      case 9: return 512;
      case 10: return 1024;
      case 11: return 2048;
      case 12: return 4096;
      default: return 0;
   }
}

// NFC_SFGT_TABLE / NFC_FWT_TABLE: 256 * 16 * 2^i in 1/fc units
NFC_TAB_HD int nfc_xgt_table(int i)
{
   return 4096 << (i & 15);
}

static inline void rate_fill(RateParams *r, double stu, int shiftBase, int rate, u32 sdd, u32 corrOff)
{
   // NfcA.cpp:157-176 (same block in B and F); NFC-V uses shiftBase 512 (NfcV.cpp:157-161)
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

static inline void params_init(Params *P, u32 sampleRate)
{
   const float NFC_FC = 13.56E6f; // Nfc.h:36 (float constant!)

   P->sampleRate = sampleRate;
   P->valid = 0;

   if (!sampleRate)
      return;

   // NfcDecoder.cpp:307-329
   P->stu = (double) sampleRate / (double) NFC_FC;
   P->etu = (int) (P->stu * 128);
   P->iirA = (float) 0.9;
   P->envW0 = (float) (1 - 5E5 / sampleRate);
   P->envW1 = (float) (1 - P->envW0);
   P->mdevW0 = (float) (1 - 2E5 / sampleRate);
   P->mdevW1 = (float) (1 - P->mdevW0);
   P->meanW0 = (float) (1 - 5E4 / sampleRate);
   P->meanW1 = (float) (1 - P->meanW0);
   P->lowThr = P->power / 1.25f;
   P->highThr = P->power * 1.25f;

   double stu = P->stu;

   // NFC-A / NFC-B: 106, 212, 424 with cascading detection delay (NfcA.cpp:141-176, NfcB.cpp:150-185)
   for (int rate = 0; rate < 3; rate++)
   {
      u32 sddA = rate > 0 ? P->A[rate - 1].sdd + P->A[rate - 1].p1 : 0;
      rate_fill(&P->A[rate], stu, 256, rate, sddA, NFCB200_OFF_CA + rate * NFCB200_CA_LEN);
      P->A[rate].sps = (u32) (int) roundf(NFC_FC / (float) (128 >> rate)); // NfcA.cpp:157

      u32 sddB = rate > 0 ? P->B[rate - 1].sdd + P->B[rate - 1].p1 : 0;
      rate_fill(&P->B[rate], stu, 256, rate, sddB, 0);
      P->B[rate].sps = (u32) (int) roundf(NFC_FC / (128 >> rate)); // NfcB.cpp:166

      // NFC-F: 212 and 424 only, no detection delay (NfcF.cpp:132-168)
      rate_fill(&P->F[rate], stu, 256, rate, 0, rate ? NFCB200_OFF_CF + (rate - 1) * NFCB200_CF_LEN : 0);
      P->F[rate].sps = (u32) (int) roundf(NFC_FC / (float) (128 >> rate));
      P->F[rate].pre1 = (u32) (int) round(stu * (128 >> rate) * 48);
   }

   // NFC-V: single rate, delay = period0 (NfcV.cpp:154-173)
   rate_fill(&P->V, stu, 512, 0, 0, NFCB200_OFF_CV);
   P->V.sdd = P->V.p0;
   if (!P->V.p1 || !P->V.p0 || P->V.sdd >= NFCB200_RING)
      return; // sample rate far outside the range the reference supports
   P->V.c1 = (NFCB200_RING - P->V.sdd) % P->V.p1;
   P->V.c0 = (NFCB200_RING - P->V.sdd) % P->V.p0;
   P->V.sps = (u32) (int) roundf(NFC_FC / 256);

   // protocol defaults (Nfc.h) scaled with int() truncation (NfcA.cpp:195-199 etc.)
   P->A_sfgt = (u32) (int) (stu * (256 * 16));
   P->A_fgt = (u32) (int) (stu * 1024);
   P->A_fwt = (u32) (int) (stu * (256 * 16 * 16));
   P->A_rgt = (u32) (int) (stu * 7000);
   P->A_fwtAtqa = (u32) (stu * (128 * 18));
   P->fwtActivation = (u32) (int) (stu * 71680);

   P->B_sfgt = (u32) (int) (stu * (256 * 16));
   P->B_fgt = (u32) (int) (stu * 1024);
   P->B_fwt = (u32) (int) (stu * (256 * 16 * 16));
   P->B_rgt = (u32) (int) (stu * 7000);
   P->B_tr0min = (u32) (stu * 1024);
   P->B_fwtAtqb = (u32) (stu * 7680);
   P->B_tr1Min = (u32) (int) (stu * 1024);
   P->B_tr1Max = (u32) (int) (stu * 3200);
   P->B_s1Min = (u32) (int) (stu * 1272);
   P->B_s1Max = (u32) (int) (stu * 1416);
   P->B_s2Min = (u32) (int) (stu * 248);
   P->B_s2Max = (u32) (int) (stu * 392);
   P->B_eofComp = (u32) (int) (stu * 352); // NfcB.cpp:622

   P->F_sfgt = (u32) (int) (stu * 4096);
   P->F_fgt = (u32) (int) (stu * 1024);
   P->F_fwt = (u32) (int) (stu * (256 * 16 * 16));
   P->F_rgt = (u32) (int) (stu * 7000);

   P->V_sfgt = (u32) (int) (stu * 4096);
   P->V_fgt = (u32) (int) (stu * 1024);
   P->V_fwt = (u32) (int) (stu * (256 * 16 * 16));
   P->V_rgt = (u32) (int) (stu * 7000);
   P->V_s1Min = (u32) (int) (stu * (768 - 32));
   P->V_s1Max = (u32) (int) (stu * (768 + 32));
   P->V_s2Min = (u32) (int) (stu * (256 - 32));
   P->V_s2Max = (u32) (int) (stu * (256 + 32));
   P->V_len2 = (u32) (int) round(4 * stu * 256);   // NfcV.cpp:224
   P->V_len8 = (u32) (int) round(256 * stu * 256);

   // the scratch layout bounds the supported sample rates (reference bound is period0 <= 1024)
   P->valid = P->A[0].p1 <= NFCB200_CA_LEN && P->F[1].p1 <= NFCB200_CF_LEN && P->V.p0 <= NFCB200_CV_LEN &&
              P->V.sdd + P->V.p0 < NFCB200_RING && P->A[2].p8 >= 1 && P->F[2].p8 >= 1;
}

static inline void params_defaults(Params *P)
{
   P->enabled = EN_A | EN_B | EN_F | EN_V;
   P->streamTime = 0;
   P->power = 0.01f;                                   // NfcTech.h:347
   P->thr[TECH_A] = TechThresholds {0.75f, 0.90f, 1.00f}; // NfcA.cpp:94-100
   P->thr[TECH_B] = TechThresholds {0.50f, 0.10f, 0.90f}; // NfcB.cpp:103-109
   P->thr[TECH_F] = TechThresholds {0.50f, 0.10f, 0.90f}; // NfcF.cpp:88-94
   P->thr[TECH_V] = TechThresholds {0.50f, 0.90f, 1.00f}; // NfcV.cpp:101-107
}

}

#endif
