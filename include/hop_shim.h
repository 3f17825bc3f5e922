/*
 * hop_shim.h -- C++ glue between the HM-15 / HEVC-HOP encoder classes and the libhopgpu C ABI.
 *
 * Header-only; compiled INSIDE the reference tree (it uses TComPattern, TComMv, TComPicYuv, TComRdCost,
 * TComDataCU exactly as TEncSearch does).  It keeps the call signatures of
 *   TEncSearch::xPatternSearch   (TLibEncoder/TEncSearch.h:471-486)
 *   TEncSearch::xPatternSearchGT (TLibEncoder/TEncSearch.h:488-517)
 * so that their bodies become one call each (INTEGRATION.md shows the patch).  Error convention of the
 * reference: no error returns -- a failing GPU call prints to stderr and exits (no CPU fallback).
 */
#ifndef HOP_SHIM_H
#define HOP_SHIM_H

#include <cstdio>
#include <cstdlib>
#include <ctime>
#include "hop_gpu.h"
#include "hop_border.h"

namespace hopshim {

/* HOP_STATS=1: wall time spent inside the library per entry point, printed when the encoder exits */
struct Stats {
  double   sec[6];
  unsigned long long calls[6];
  double   shape_sec[2][17][17];               // [K1|K2][cols/4][rows/4]
  unsigned long long shape_calls[2][17][17];
  bool     on;
  Stats() : on(getenv("HOP_STATS") != NULL)
  {
    for (int i = 0; i < 6; i++) { sec[i] = 0; calls[i] = 0; }
    for (int k = 0; k < 2; k++) for (int a = 0; a < 17; a++) for (int b = 0; b < 17; b++) { shape_sec[k][a][b] = 0; shape_calls[k][a][b] = 0; }
  }
  ~Stats()
  {
    if (!on) return;
    static const char* name[6] = {"xPatternSearch", "xPatternSearchGT", "refUpdate", "refReset+create", "prefetch", "hostBorder"};
    for (int i = 0; i < 6; i++)
      fprintf(stderr, "hopshim: %-18s %9llu calls %9.3f s (%.1f us/call)\n", name[i], calls[i], sec[i],
              calls[i] ? 1e6 * sec[i] / calls[i] : 0.0);
    for (int k = 0; k < 2; k++) for (int a = 0; a < 17; a++) for (int b = 0; b < 17; b++)
      if (shape_calls[k][a][b])
        fprintf(stderr, "hopshim:   %-16s %2dx%-2d %8llu calls %8.3f s (%.1f us/call)\n", name[k], 4 * a, 4 * b,
                shape_calls[k][a][b], shape_sec[k][a][b], 1e6 * shape_sec[k][a][b] / shape_calls[k][a][b]);
  }
};
inline Stats& stats() { static Stats s; return s; }
struct Timer {
  int k, cols, rows; timespec t0;
  explicit Timer(int kind, int c = 0, int r = 0) : k(kind), cols(c), rows(r) { if (stats().on) clock_gettime(CLOCK_MONOTONIC, &t0); }
  ~Timer()
  {
    if (!stats().on) return;
    timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1);
    const double dt = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    stats().sec[k] += dt;
    stats().calls[k]++;
    if (k < 2 && cols > 0 && cols <= 64 && rows > 0 && rows <= 64) { stats().shape_sec[k][cols / 4][rows / 4] += dt; stats().shape_calls[k][cols / 4][rows / 4]++; }
  }
};

struct State {
  HopCtx*    ctx;
  const Pel* origin;       // host address of sample (0,0) of the mirrored SS reference plane
  const Pel* buf_lo;       // first / one-past-last host sample of that plane incl. margins
  const Pel* buf_hi;
  int        stride;
  bool       recording;    // inside prefetchBegin()/prefetchEnd(): searches are enqueued speculatively, not awaited
  State() : ctx(NULL), origin(NULL), buf_lo(NULL), buf_hi(NULL), stride(0), recording(false) {}
};

inline State& state() { static State s; return s; }

inline void check(int status, const char* what)
{
  if (status != HOP_OK) {
    fprintf(stderr, "libhopgpu: %s failed (%d): %s\n", what, status, hop_last_error());
    exit(EXIT_FAILURE);
  }
}

/* HOP_FUSED=0 keeps the three stages as separate calls (fractional refinement on the host) */
inline bool fused() { static int f = -1; if (f < 0) { const char* e = getenv("HOP_FUSED"); f = (e && e[0] == '0') ? 0 : 1; } return f != 0; }

inline HopCtx* ctx()
{
  State& s = state();
  if (!s.ctx) {
    const char* dev = getenv("HOP_DEVICE");
    Timer tm(3);
    check(hop_ctx_create(dev ? atoi(dev) : 0, &s.ctx), "hop_ctx_create");
  }
  return s.ctx;
}

/* TEncTop::create (TEncTop.cpp:89-110): one context per encoder instance = per GPU (HOP_DEVICE selects it). */
inline void create() { ctx(); }

/* A front end that runs several encodes in one process (integration/hop_batch_main.cpp) keeps the context -- its
 * stream, pinned result slots and scratch buffers -- across TEncTop::destroy / create; only the mirror is detached. */
inline bool& keepContext() { static bool keep = false; return keep; }

/* TEncTop::destroy (TEncTop.cpp:204-213): release the context; HOP_STATS=1 prints what the single-PU path did. */
inline void destroy(bool force = false)
{
  State& s = state();
  if (!s.ctx) return;
  if (keepContext() && !force) { s.origin = NULL; s.buf_lo = NULL; s.buf_hi = NULL; s.recording = false; return; }
  if (stats().on) {
    HopCtxStats cs;
    if (hop_ctx_stats(s.ctx, &cs) == HOP_OK) {
      const double gpu_s = stats().sec[0] + stats().sec[1] + stats().sec[4];
      fprintf(stderr, "hopshim: single-PU calls %llu  cache hits %llu  misses %llu  prefetched %llu  dropped %llu\n",
              (unsigned long long)cs.single_calls, (unsigned long long)cs.cache_hits, (unsigned long long)cs.cache_misses,
              (unsigned long long)cs.prefetched, (unsigned long long)cs.prefetch_dropped);
      fprintf(stderr, "hopshim: HOP candidates scored for the encoder %llu in %.3f s of search calls = %.3g candidates/s in-encoder\n",
              (unsigned long long)cs.candidates, gpu_s, gpu_s > 0 ? (double)cs.candidates / gpu_s : 0.0);
    }
  }
  hop_ctx_destroy(s.ctx);
  s = State();
}

/* The SS reference is not (re)wired for this slice: nothing may be searched on a stale mirror. */
inline void refInvalidate() { State& s = state(); s.origin = NULL; s.buf_lo = NULL; s.buf_hi = NULL; }

/* Speculation window (TEncCu::xCompressCU, before the inter modes of a CU are tried, TEncCu.cpp:456-633): the
 * caller runs the reference's own xCheckRDCostInter for the partition modes it is going to try; inside the window
 * xMotionSearchSS does not wait -- it enqueues the search of the mode's FIRST PU (whose inputs depend on nothing
 * inside the CU) and reports "no valid vector", which makes predInterSearch return at once (TEncSearch.cpp:3964).
 * The real pass afterwards builds the same requests and finds the results waiting. */
inline bool prefetchEnabled() { static int f = -1; if (f < 0) { const char* e = getenv("HOP_PREFETCH"); f = (e && e[0] == '0') ? 0 : 1; } return f != 0 && fused(); }
inline bool prefetchAmp()     { static int f = -1; if (f < 0) { const char* e = getenv("HOP_PREFETCH_AMP"); f = (e && e[0] == '0') ? 0 : 1; } return f != 0; }
inline void prefetchBegin() { state().recording = true; }
inline void prefetchEnd()   { state().recording = false; }

/* After TComSlice::setRefPicList wired the SS reference of an ISS or PSS slice (TEncGOP.cpp:794,
 * TComSlice.cpp:241-254, 366-377, 496-502): the plane was reset to NOT_VALID and border-extended on the host;
 * mirror it as it is. */
inline void refReset(TComPicYuv* pic)
{
  HopCtx* c = ctx();
  Timer tm(3);
  State& s = state();
  const int m = pic->getLumaMargin(), w = pic->getWidth(), h = pic->getHeight();
  check(hop_ref_create(c, w, h, m), "hop_ref_create");
  const size_t samples = (size_t)pic->getStride() * (h + 2 * m);
  check(hop_ref_upload(s.ctx, pic->getBufY(), samples), "hop_ref_upload");
  s.origin = pic->getLumaAddr();
  s.buf_lo = pic->getBufY();
  s.buf_hi = pic->getBufY() + samples;
  s.stride = pic->getStride();
}

/* Tail of TEncCu::xCopyYuv2SSRef after copyToPicYuv put a CU's reconstruction into the SS reference
 * (TEncCu.cpp:1694-1696): re-extend the picture border and bring the device mirror up to date.  For the mirrored
 * plane the host border is re-extended incrementally (hop_border.h: same samples as the reference's full pass,
 * which costs megabytes per CU on large pictures); HOP_HOST_BORDER=full keeps the reference's own call. */
inline bool hostBorderIncremental() { static int f = -1; if (f < 0) { const char* e = getenv("HOP_HOST_BORDER"); f = (e && e[0] == 'f') ? 0 : 1; } return f != 0; }

inline void refCommit(TComPicYuv* pic, int x, int y, int w, int h)
{
  State& s = state();
  const bool mine = pic->getLumaAddr() == s.origin;
  if (mine && hostBorderIncremental()) {
    Timer tb(5);
    const int m = pic->getLumaMargin(), mc = pic->getChromaMargin();
    hop_extend_patch_border(pic->getLumaAddr(), pic->getStride(), pic->getWidth(), pic->getHeight(), m, m, x, y, w, h);
    hop_extend_patch_border(pic->getCbAddr(), pic->getCStride(), pic->getWidth() >> 1, pic->getHeight() >> 1, mc, mc, x >> 1, y >> 1, w >> 1, h >> 1);
    hop_extend_patch_border(pic->getCrAddr(), pic->getCStride(), pic->getWidth() >> 1, pic->getHeight() >> 1, mc, mc, x >> 1, y >> 1, w >> 1, h >> 1);
    pic->setBorderExtension(true);
  } else {
    Timer tb(5);
    pic->setBorderExtension(false);
    pic->extendPicBorder();
  }
  if (!mine) return;     // not the mirrored plane
  Timer tm(2);
  check(hop_ref_update(s.ctx, x, y, w, h, pic->getLumaAddr() + (size_t)y * pic->getStride() + x, pic->getStride()),
        "hop_ref_update");
}

inline bool owns(const Pel* p) { const State& s = state(); return s.ctx && p >= s.buf_lo && p < s.buf_hi; }

inline HopCostState costState(TComRdCost* rd)
{
  HopCostState c;
  c.lambda_cost = rd->hopGetCost();
  c.cost_scale  = rd->hopGetCostScale();
  c.pred.hor    = rd->hopGetPredictor().getHor();
  c.pred.ver    = rd->hopGetPredictor().getVer();
  return c;
}

/* Body of TEncSearch::xPatternSearch for an SS reference (TEncSearch.cpp:6262-6371). */
inline void xPatternSearch(TComPattern* pcPatternKey, Pel* piRefY, Int iRefStride, TComMv* pcMvSrchRngLT,
                           TComMv* pcMvSrchRngRB, TComMv& rcMv, UInt& ruiSAD, Int riOffsetX, Int riOffsetY,
                           TComMv* ssBestCand, Bool isSSE, Bool useFastEnc, Int bitDepth, TComRdCost* rd)
{
  Timer tm(0, pcPatternKey->getROIYWidth(), pcPatternKey->getROIYHeight());
  State& s = state();
  HopSearchJob j;
  j.org_off = 0;
  j.ref_off = piRefY - s.origin;
  j.org_stride = pcPatternKey->getPatternLStride();
  j.ref_stride = iRefStride;
  j.cols = pcPatternKey->getROIYWidth();
  j.rows = pcPatternKey->getROIYHeight();
  j.rng_left = pcMvSrchRngLT->getHor(); j.rng_top = pcMvSrchRngLT->getVer();
  j.rng_right = pcMvSrchRngRB->getHor(); j.rng_bottom = pcMvSrchRngRB->getVer();
  j.offset_x = riOffsetX; j.offset_y = riOffsetY;
  j.is_ss = isSSE ? 1 : 0;
  j.fast_enc = useFastEnc ? 1 : 0;
  j.bit_depth = bitDepth;
  j.cost = costState(rd);
  HopSearchResult r;
  const size_t org_samples = (size_t)(j.rows - 1) * j.org_stride + j.cols;
  check(hop_pattern_search_batch(s.ctx, 1, &j, pcPatternKey->getROIY(), org_samples, NULL, 0, &r), "hop_pattern_search_batch");
  if (!r.found) { ruiSAD = MAX_UINT; return; }          // :6356-6360, rcMv / ssBestCand untouched
  rcMv.set(r.mv.hor, r.mv.ver);                          // :6363
  ssBestCand[0].set(r.mv.hor, r.mv.ver);                 // :6344 (IT_SS_NUMBER_OF_BEST_CAND 1)
  ruiSAD = r.sad;                                        // :6365
}

/* Body of TEncSearch::xPatternSearchGT, diamond branch (TEncSearch.cpp:4686-4790, 5093-5467). */
inline void xPatternSearchGT(TComDataCU* pcCU, TComPattern* pcPatternKey, Pel* piRefY, Int iRefStride,
                             TComMv* pcMvInt, TComMv* rcMvHalf, TComMv* rcMvQter, TComMv* rcGT0, TComMv* rcGT1,
                             TComMv* rcGT2, TComMv* rcGT3, Bool& gtFlag, UInt& ruiCost, TComMv* bestSSCand,
                             Bool useHADME, Int bitDepth, TComRdCost* rd)
{
  Timer tm(1, pcPatternKey->getROIYWidth(), pcPatternKey->getROIYHeight());
  State& s = state();
  HopGtJob j;
  j.org_off = 0;
  j.ref_off = piRefY - s.origin;
  j.org_stride = pcPatternKey->getPatternLStride();
  j.ref_stride = iRefStride;
  j.cols = pcPatternKey->getROIYWidth();
  j.rows = pcPatternKey->getROIYHeight();
  j.ss_cand.hor = bestSSCand[0].getHor(); j.ss_cand.ver = bestSSCand[0].getVer();
  AMVPInfo* amvp = pcCU->getCUMvField(REF_PIC_LIST_0)->getAMVPInfo();     // :5100-5104
  j.num_pred = amvp->iN;
  for (int i = 0; i < HOP_MAX_PRED; i++) {
    j.amvp[i].hor = i < amvp->iN ? amvp->m_acMvCand[i].getHor() : 0;
    j.amvp[i].ver = i < amvp->iN ? amvp->m_acMvCand[i].getVer() : 0;
  }
  j.threshold = ruiCost;
  j.use_had = useHADME ? 1 : 0;
  j.bit_depth = bitDepth;
  j.cost = costState(rd);
  HopGtResult r;
  const size_t org_samples = (size_t)(j.rows - 1) * j.org_stride + j.cols;
  check(hop_pattern_search_gt_batch(s.ctx, 1, &j, pcPatternKey->getROIY(), org_samples, NULL, 0, &r), "hop_pattern_search_gt_batch");
  rcGT0->set(r.gt[0].hor, r.gt[0].ver); rcGT1->set(r.gt[1].hor, r.gt[1].ver);
  rcGT2->set(r.gt[2].hor, r.gt[2].ver); rcGT3->set(r.gt[3].hor, r.gt[3].ver);
  gtFlag = r.gt_flag != 0;
  if (r.gt_flag) {                                       // :5441-5457
    ruiCost = r.cost;
    pcMvInt->set(r.mv_int.hor, r.mv_int.ver);
    rcMvHalf->set(0, 0);
    rcMvQter->set(0, 0);
  }
}

/* The GPU part of TEncSearch::xMotionEstimation for an SS reference in ONE call (TEncSearch.cpp:4572-4642):
 * xPatternSearch -> validity check -> xPatternSearchFracDIF -> xPatternSearchGT.  Returns false when the
 * integer search found nothing valid or the zero vector (the caller then sets bNotValCU, :4603-4611);
 * outputs are then exactly what the reference leaves at that point. */
inline bool xMotionSearchSS(TComDataCU* pcCU, TComPattern* pcPatternKey, Pel* piRefY, Int iRefStride,
                            TComMv* pcMvSrchRngLT, TComMv* pcMvSrchRngRB, Int riOffsetX, Int riOffsetY,
                            Bool useFastEnc, Bool useHADME, Int bitDepth, TComRdCost* rd, Bool bUseGT,
                            TComMv& rcMv, UInt& ruiCost, TComMv& rcMvHalf, TComMv& rcMvQter,
                            TComMv& rcGT0, TComMv& rcGT1, TComMv& rcGT2, TComMv& rcGT3, Bool& gtFlag, TComMv* ssBestCand)
{
  State& s = state();
  HopMotionJob mj;
  HopSearchJob& j = mj.search;
  j.org_off = 0;
  j.ref_off = piRefY - s.origin;
  j.org_stride = pcPatternKey->getPatternLStride();
  j.ref_stride = iRefStride;
  j.cols = pcPatternKey->getROIYWidth();
  j.rows = pcPatternKey->getROIYHeight();
  j.rng_left = pcMvSrchRngLT->getHor(); j.rng_top = pcMvSrchRngLT->getVer();
  j.rng_right = pcMvSrchRngRB->getHor(); j.rng_bottom = pcMvSrchRngRB->getVer();
  j.offset_x = riOffsetX; j.offset_y = riOffsetY;
  j.is_ss = 1;
  j.fast_enc = useFastEnc ? 1 : 0;
  j.bit_depth = bitDepth;
  j.cost = costState(rd);                                   /* cost scale 2 at this point (:4560) */
  mj.use_had = useHADME ? 1 : 0;
  mj.use_gt = bUseGT ? 1 : 0;
  AMVPInfo* amvp = pcCU->getCUMvField(REF_PIC_LIST_0)->getAMVPInfo();     /* :5100-5104 */
  mj.num_pred = amvp->iN;
  for (int i = 0; i < HOP_MAX_PRED; i++) {
    mj.amvp[i].hor = i < amvp->iN ? amvp->m_acMvCand[i].getHor() : 0;
    mj.amvp[i].ver = i < amvp->iN ? amvp->m_acMvCand[i].getVer() : 0;
  }
  HopMotionResult r;
  const size_t org_samples = (size_t)(j.rows - 1) * j.org_stride + j.cols;
  gtFlag = false;
  rcGT0.set(0, 0); rcGT1.set(0, 0); rcGT2.set(0, 0); rcGT3.set(0, 0);
  if (s.recording) {                                                          /* speculation window: enqueue, do not wait */
    Timer tp(4);
    check(hop_motion_search_prefetch(s.ctx, 1, &mj, pcPatternKey->getROIY(), org_samples), "hop_motion_search_prefetch");
    ruiCost = MAX_UINT;
    return false;
  }
  Timer tm(1, pcPatternKey->getROIYWidth(), pcPatternKey->getROIYHeight());
  check(hop_motion_search_batch(s.ctx, 1, &mj, pcPatternKey->getROIY(), org_samples, NULL, 0, &r), "hop_motion_search_batch");
  if (!r.search.found) { ruiCost = MAX_UINT; return false; }               /* :6356-6360 */
  rcMv.set(r.search.mv.hor, r.search.mv.ver);                              /* :6363 */
  ssBestCand[0].set(r.search.mv.hor, r.search.mv.ver);
  ruiCost = r.search.sad;                                                   /* :6365 */
  if (!r.refined) return false;                                             /* zero vector, :4604 */
  rcMvHalf.set(r.frac.half.hor, r.frac.half.ver);                          /* :6564-6610 */
  rcMvQter.set(r.frac.qter.hor, r.frac.qter.ver);
  ruiCost = r.frac.cost;
  if (bUseGT && r.gt.gt_flag) {                                             /* :5441-5457 */
    gtFlag = true;
    rcGT0.set(r.gt.gt[0].hor, r.gt.gt[0].ver); rcGT1.set(r.gt.gt[1].hor, r.gt.gt[1].ver);
    rcGT2.set(r.gt.gt[2].hor, r.gt.gt[2].ver); rcGT3.set(r.gt.gt[3].hor, r.gt.gt[3].ver);
    ruiCost = r.gt.cost;
    rcMv.set(r.gt.mv_int.hor, r.gt.mv_int.ver);
    rcMvHalf.set(0, 0);
    rcMvQter.set(0, 0);
  }
  return true;
}

}  // namespace hopshim
#endif
