/*
 * ref_harness.cpp -- C entry points into the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).
 *
 * Linked against the reference's own objects (compiled by oracle/Makefile straight from
 * /root/reference/source with the reference's flags) into oracle/_ref/libhopref.so.  It subclasses
 * TEncSearch so that the protected members xPatternSearch / xPatternSearchGT
 * (TLibEncoder/TEncSearch.h:471-517) can be driven with synthetic inputs, and exposes the DistFunc
 * table of TComRdCost and TComPicYuv::extendPicBorder.  Nothing of the reference is re-implemented
 * here: every result below is computed by reference code.
 *
 * Used (a) to pin oracle/hop_oracle.c and to generate tests/golden/, (b) as the "reference" CPU
 * baseline of bench.py.  The job/result structs are those of include/hop_gpu.h.
 */
// K6 drives xPredInterLumaBlk / xPredInterChromaBlk, which take a TComDataCU* and a TComPicYuv* only to turn
// (CU address, z-order index) into a sample address.  The harness points a TComPicYuv at the caller's buffer by
// setting its (private) geometry members directly -- access specifiers are opened for THIS translation unit only;
// the reference's sources and objects are untouched and the class layouts do not depend on access specifiers.
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <fstream>
#include <sstream>
#include <list>
#include <map>
#include <string>
#include <vector>
#define private public
#define protected public
#include "TLibEncoder/TEncSearch.h"
#include "TLibEncoder/TEncCfg.h"
#include "TLibCommon/TComRdCost.h"
#include "TLibCommon/TComDataCU.h"
#include "TLibCommon/TComPattern.h"
#include "TLibCommon/TComPicYuv.h"
#include "TLibCommon/TComRom.h"

#include "TLibCommon/TComYuv.h"
#undef private
#undef protected

#include "../include/hop_gpu.h"
#include <string.h>

namespace {

class RefSearch : public TEncSearch
{
public:
  TEncCfg     cfg;
  TComRdCost  rd;
  TComDataCU  cu;      // only getCUMvField(REF_PIC_LIST_0)->getAMVPInfo() is read (TEncSearch.cpp:5100)
  TComPattern pattern;

  RefSearch()
  {
    m_pcEncCfg = &cfg;
    m_pcRdCost = &rd;
    rd.init();                  // TComRdCost.cpp:177-233, fills m_afpDistortFunc
    initTempBuff();             // TComPrediction.cpp:81, allocates m_filteredBlock*
    m_cDistParam.bApplyWeight = false;   // what setWpScalingDistParam leaves for SS slices
    predYuv = NULL;
  }

  void setCost(const HopCostState& cs)
  {
    // getMotionCost(1,0) copies m_uiLambdaMotionSAD to m_uiCost; setLambda derives it from a double.
    // m_uiLambdaMotionSAD is private, so drive it through the public pair with an exact inverse:
    // floor(65536*sqrt(l)) must equal lambda_cost -> search the double that does.
    double s = ((double)cs.lambda_cost + 0.5) / 65536.0;
    rd.setLambda(s * s);
    rd.getMotionCost(true, 0);
    TComMv pred(cs.pred.hor, cs.pred.ver);
    rd.setPredictor(pred);
    rd.setCostScale(cs.cost_scale);
  }

  void patternSearch(const HopSearchJob* j, const int16_t* org, const int16_t* ref, HopSearchResult* out)
  {
    g_bitDepthY = j->bit_depth;
    cfg.setUseFastEnc(j->fast_enc != 0);
    setCost(j->cost);
    pattern.initPattern((Pel*)org + j->org_off, NULL, NULL, j->cols, j->rows, j->org_stride, 0, 0);
    TComMv lt(j->rng_left, j->rng_top), rb(j->rng_right, j->rng_bottom);
    TComMv mv(0x7fff, 0x7fff);
    TComMv ssBest[1]; ssBest[0].set(0x7fff, 0x7fff);
    UInt sad = 0;
    xPatternSearch(&pattern, (Pel*)ref + j->ref_off, j->ref_stride, &lt, &rb, mv, sad,
                   j->offset_x, j->offset_y, ssBest, j->is_ss != 0);
    memset(out, 0, sizeof(*out));
    out->sad = sad;
    if (sad == MAX_UINT && mv.getHor() == 0x7fff) { out->found = 0; out->cost = MAX_UINT; return; }
    out->found = 1;
    out->mv.hor = mv.getHor(); out->mv.ver = mv.getVer();
    out->cost = sad + rd.getCost(mv.getHor(), mv.getVer());
    // ssBestCand[0] must equal rcMv (IT_SS_NUMBER_OF_BEST_CAND 1); flag a mismatch loudly
    if (ssBest[0].getHor() != mv.getHor() || ssBest[0].getVer() != mv.getVer()) out->found = -1;
  }

  void patternSearchGT(const HopGtJob* j, const int16_t* org, const int16_t* ref, HopGtResult* out)
  {
    g_bitDepthY = j->bit_depth;
    cfg.setUseHADME(j->use_had != 0);
    setCost(j->cost);
    pattern.initPattern((Pel*)org + j->org_off, NULL, NULL, j->cols, j->rows, j->org_stride, 0, 0);
    AMVPInfo* amvp = cu.getCUMvField(REF_PIC_LIST_0)->getAMVPInfo();
    amvp->iN = j->num_pred;
    for (int i = 0; i < HOP_MAX_PRED; i++) amvp->m_acMvCand[i].set(j->amvp[i].hor, j->amvp[i].ver);
    TComMv ssBest[1]; ssBest[0].set(j->ss_cand.hor, j->ss_cand.ver);
    // pcMvInt / half / quarter only feed the prologue's scratch interpolation in the diamond branch
    TComMv mvInt(j->ss_cand.hor, j->ss_cand.ver), mvHalf(0, 0), mvQter(0, 0);
    TComMv gt0, gt1, gt2, gt3;
    Bool gtFlag = false;
    UInt cost = j->threshold;
    xPatternSearchGT(&cu, &pattern, (Pel*)ref + j->ref_off, j->ref_stride, &mvInt, &mvHalf, &mvQter,
                     &gt0, &gt1, &gt2, &gt3, gtFlag, cost, false, ssBest);
    memset(out, 0, sizeof(*out));
    out->gt_flag = gtFlag ? 1 : 0;
    out->gt[0].hor = gt0.getHor(); out->gt[0].ver = gt0.getVer();
    out->gt[1].hor = gt1.getHor(); out->gt[1].ver = gt1.getVer();
    out->gt[2].hor = gt2.getHor(); out->gt[2].ver = gt2.getVer();
    out->gt[3].hor = gt3.getHor(); out->gt[3].ver = gt3.getVer();
    out->cost = cost;
    if (gtFlag) { out->mv_int.hor = mvInt.getHor(); out->mv_int.ver = mvInt.getVer(); }
    out->best_index = -2;      // not observable from outside the reference
    out->n_candidates = 0;
  }

  void fracSearch(const HopFracJob* j, const int16_t* org, const int16_t* ref, HopFracResult* out)
  {
    g_bitDepthY = j->bit_depth;
    cfg.setUseHADME(j->use_had != 0);
    HopCostState cs = j->cost;
    cs.cost_scale = 1;                         // xMotionEstimation: setCostScale(1) before the call (:4615)
    setCost(cs);
    pattern.initPattern((Pel*)org + j->org_off, NULL, NULL, j->cols, j->rows, j->org_stride, 0, 0);
    TComMv mvInt(j->mv_int.hor, j->mv_int.ver), mvHalf, mvQter;
    UInt cost = 0;
    xPatternSearchFracDIF(&cu, &pattern, (Pel*)ref + j->ref_off, j->ref_stride, &mvInt, mvHalf, mvQter, cost, false);
    memset(out, 0, sizeof(*out));
    out->half.hor = mvHalf.getHor(); out->half.ver = mvHalf.getVer();
    out->qter.hor = mvQter.getHor(); out->qter.ver = mvQter.getVer();
    out->cost = cost;
  }

  // debugging aid for tests: copy one of the 16 interpolated planes left behind by the last frac search
  void fracPlane(int ver, int hor, int16_t* dst, int cols, int rows)
  {
    Pel* p = m_filteredBlock[ver][hor].getLumaAddr();
    Int st = m_filteredBlock[ver][hor].getStride();
    for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) dst[r * cols + c] = p[r * st + c];
  }

  // protected members of TComPrediction (TComPrediction.h:106,110), re-exported
  void calcParam(Int* x, Int* y, Double* h, Int w, Int hh) { calcParamProjective(x, y, h, w, hh); }
  void warp(Pel* r, Pel* aux, Double* h, Int W, Int H, Int stride, Int nss) { ProjectiveTransform(r, aux, h, W, H, stride, nss); }

  // K6: the reference's own xPredInterLumaBlk / xPredInterChromaBlk on the caller's plane (TComPrediction.cpp:639-720,
  // 1235-1347), then -- in the order of xGetInterPredictionError (TEncSearch.cpp:2951-2977) and xGetTemplateCost
  // (:4390-4477) -- the reference's own setDistParam / DistFunc, isValidPattern, getDistPart and calcRdCost.
  TComYuv* predYuv;
  UInt     zero_offsets[256];
  void predict(const HopPredJob* j, const int16_t* org, const int16_t* refbuf, int16_t* dst, HopPredResult* out)
  {
    const bool chroma = j->comp != 0;
    g_bitDepthY = j->bit_depth; g_bitDepthC = j->bit_depth;
    if (!predYuv) { predYuv = new TComYuv; predYuv->create(64, 64); memset(zero_offsets, 0, sizeof(zero_offsets)); }
    TComPicYuv pic;                                   // a view of the caller's plane: origin = the PU's own position
    pic.m_cuOffsetY = (Int*)zero_offsets; pic.m_cuOffsetC = (Int*)zero_offsets;
    pic.m_buOffsetY = (Int*)zero_offsets; pic.m_buOffsetC = (Int*)zero_offsets;
    pic.m_piPicOrgY = (Pel*)refbuf + j->ref_off; pic.m_piPicOrgU = (Pel*)refbuf + j->ref_off; pic.m_piPicOrgV = (Pel*)refbuf + j->ref_off;
    pic.m_iLumaMarginX = 0; pic.m_iChromaMarginX = 0;
    pic.m_iPicWidth = chroma ? 2 * j->ref_stride : j->ref_stride;      // getStride() / getCStride() == ref_stride
    cu.m_uiCUAddr = 0; cu.m_uiAbsIdxInLCU = 0;
    TComMv mv(j->mv.hor, j->mv.ver), g0(j->gt[0].hor, j->gt[0].ver), g1(j->gt[1].hor, j->gt[1].ver),
           g2(j->gt[2].hor, j->gt[2].ver), g3(j->gt[3].hor, j->gt[3].ver), zero(0, 0);
    memset(out, 0, sizeof(*out));
    out->valid = 1;
    const int bw = chroma ? j->cols >> 1 : j->cols, bh = chroma ? j->rows >> 1 : j->rows;
    if (j->template_cost) {
      TComMv probe(j->mv_probe.hor, j->mv_probe.ver);
      if (j->is_ss && !rd.isValidPattern(pic.getLumaAddr(0, 0), pic.getStride(), probe, j->cols, j->rows)) {   // :4420-4436
        out->valid = 0; out->cost = MAX_INT;
        pic.m_cuOffsetY = pic.m_cuOffsetC = pic.m_buOffsetY = pic.m_buOffsetC = NULL; pic.m_piPicOrgY = pic.m_piPicOrgU = pic.m_piPicOrgV = NULL;
        return;
      }
      xPredInterLumaBlk(&cu, &pic, 0, &mv, j->cols, j->rows, predYuv, false, false, &zero, &zero, &zero, &zero);   // :4451-4455
      double s = ((double)j->lambda_sad + 0.5) / 65536.0;      // m_uiLambdaMotionSAD is private: set it through setLambda
      rd.setLambda(s * s);
      UInt c = rd.getDistPart(g_bitDepthY, predYuv->getLumaAddr(0), predYuv->getStride(), (Pel*)org + j->org_off, j->org_stride,
                              j->cols, j->rows, TEXT_LUMA, DF_SAD);                                                 // :4473
      out->dist = c;
      out->cost = (UInt)rd.calcRdCost(j->mvp_bits, c, false, DF_SAD);                                               // :4474
    } else {
      if (!chroma) xPredInterLumaBlk(&cu, &pic, 0, &mv, j->cols, j->rows, predYuv, false, j->gt_flag != 0, &g0, &g1, &g2, &g3);
      else         xPredInterChromaBlk(&cu, &pic, 0, &mv, j->cols, j->rows, predYuv, false, j->gt_flag != 0, &g0, &g1, &g2, &g3);
      if (j->dist_func) {
        DistParam dp;
        dp.bApplyWeight = false;
        Pel* p = chroma ? predYuv->getCbAddr(0) : predYuv->getLumaAddr(0);
        rd.setDistParam(dp, j->bit_depth, (Pel*)org + j->org_off, j->org_stride, p, chroma ? predYuv->getCStride() : predYuv->getStride(),
                        bw, bh, j->dist_func == HOP_DF_HADS);                                                      // :2971-2975
        out->dist = dp.DistFunc(&dp);
        out->cost = out->dist;
      }
    }
    if (j->dst_off >= 0 && dst) {
      Pel* p = chroma ? predYuv->getCbAddr(0) : predYuv->getLumaAddr(0);
      const int st = chroma ? predYuv->getCStride() : predYuv->getStride();
      for (int y = 0; y < bh; y++) for (int x = 0; x < bw; x++) dst[j->dst_off + y * j->dst_stride + x] = p[y * st + x];
    }
    // the view owns nothing: detach before the destructor of TComPicYuv runs
    pic.m_cuOffsetY = pic.m_cuOffsetC = pic.m_buOffsetY = pic.m_buOffsetC = NULL; pic.m_piPicOrgY = pic.m_piPicOrgU = pic.m_piPicOrgV = NULL;
  }

  // K7: the reference's own predIntraLumaAng + calcHAD for the 35 modes (TEncSearch.cpp:2451-2464).  m_piYuvExt gets the
  // caller's reference samples in the layout initAdiPattern leaves (border row / column of the unfiltered buffer, then of
  // the filtered one, TComPattern.cpp:583-607); the interior of the two buffers is never read by the predictors.
  void intraPrescreen(const HopIntraJob* j, const int16_t* org, const int32_t* refs, uint32_t* out)
  {
    g_bitDepthY = j->bit_depth;
    const int n = j->size, sw = 2 * n + 1;
    const int32_t* r = refs + j->refs_off;
    for (int f = 0; f < 2; f++) {
      Int* buf = m_piYuvExt + f * sw * sw;
      for (int k = 0; k < sw; k++) { buf[k] = r[(2 * f) * sw + k]; buf[k * sw] = r[(2 * f + 1) * sw + k]; }
      buf[0] = r[(2 * f) * sw];
    }
    if (!predYuv) { predYuv = new TComYuv; predYuv->create(64, 64); memset(zero_offsets, 0, sizeof(zero_offsets)); }
    Pel* pred = predYuv->getLumaAddr(0);
    const UInt stride = predYuv->getStride();
    for (UInt mode = 0; mode < 35; mode++) {
      predIntraLumaAng(&pattern, mode, pred, stride, n, n, j->above_avail != 0, j->left_avail != 0);
      out[mode] = rd.calcHAD(g_bitDepthY, (Pel*)org + j->org_off, j->org_stride, pred, stride, n, n);
    }
  }

  uint32_t dist(const HopDistJob* j, const int16_t* org, const int16_t* cur)
  {
    DistParam dp;
    pattern.initPattern((Pel*)org + j->org_off, NULL, NULL, j->cols, j->rows, j->org_stride, 0, 0);
    if (j->func == HOP_DF_HADS)
      rd.setDistParam(&pattern, (Pel*)cur + j->cur_off, j->cur_stride, 1, dp, true);   // TComRdCost.cpp:332
    else
      rd.setDistParam(&pattern, (Pel*)cur + j->cur_off, j->cur_stride, dp);            // TComRdCost.cpp:298
    dp.iSubShift = (j->func == HOP_DF_HADS) ? 0 : j->sub_shift;
    dp.bitDepth = j->bit_depth;
    dp.bApplyWeight = false;
    dp.uiComp = 0;
    return dp.DistFunc(&dp);
  }
};

RefSearch* g_ref = NULL;
RefSearch* ref()
{
  if (!g_ref) {
    g_uiMaxCUWidth = 64; g_uiMaxCUHeight = 64; g_uiMaxCUDepth = 4; g_uiAddCUDepth = 1;
    g_bitDepthY = 8; g_bitDepthC = 8;
    initROM();
    g_ref = new RefSearch();
  }
  return g_ref;
}

}  // namespace

extern "C" {

void ref_pattern_search(const HopSearchJob* job, const int16_t* org, const int16_t* refbuf, HopSearchResult* out)
{ ref()->patternSearch(job, org, refbuf, out); }

void ref_pattern_search_gt(const HopGtJob* job, const int16_t* org, const int16_t* refbuf, HopGtResult* out)
{ ref()->patternSearchGT(job, org, refbuf, out); }

uint32_t ref_dist(const HopDistJob* job, const int16_t* org, const int16_t* cur)
{ return ref()->dist(job, org, cur); }

void ref_pattern_search_batch(int n, const HopSearchJob* jobs, const int16_t* org, const int16_t* refbuf, HopSearchResult* out)
{ for (int i = 0; i < n; i++) ref()->patternSearch(&jobs[i], org, refbuf, &out[i]); }

void ref_pattern_search_gt_batch(int n, const HopGtJob* jobs, const int16_t* org, const int16_t* refbuf, HopGtResult* out)
{ for (int i = 0; i < n; i++) ref()->patternSearchGT(&jobs[i], org, refbuf, &out[i]); }

void ref_frac_search_batch(int n, const HopFracJob* jobs, const int16_t* org, const int16_t* refbuf, HopFracResult* out)
{ for (int i = 0; i < n; i++) ref()->fracSearch(&jobs[i], org, refbuf, &out[i]); }

void ref_predict_batch(int n, const HopPredJob* jobs, const int16_t* org, const int16_t* refbuf, int16_t* dst, HopPredResult* out)
{ for (int i = 0; i < n; i++) ref()->predict(&jobs[i], org, refbuf, dst, &out[i]); }

void ref_intra_prescreen_batch(int n, const HopIntraJob* jobs, const int16_t* org, const int32_t* refs, uint32_t* out)
{ for (int i = 0; i < n; i++) ref()->intraPrescreen(&jobs[i], org, refs, out + (size_t)i * HOP_INTRA_MODES); }

void ref_frac_plane(int ver, int hor, int16_t* dst, int cols, int rows) { ref()->fracPlane(ver, hor, dst, cols, rows); }

/* TComRdCost::xGetComponentBits is private; getBitsGT (public, TComRdCost.h:204) exposes it. */
uint32_t ref_bits_gt(int x0, int y0, int x1, int y1, int x2, int y2)
{ return ref()->rd.getBitsGT(x0, y0, x1, y1, x2, y2, 0, 0); }

uint32_t ref_get_cost_xy(const HopCostState* cs, int x, int y)
{ ref()->setCost(*cs); return ref()->rd.getCost(x, y); }

/* TComPrediction::calcParamProjective / ProjectiveTransform are public members of TComPrediction */
void ref_calc_param_projective(const int32_t x[4], const int32_t y[4], double h[9], int width, int height)
{ ref()->calcParam((Int*)x, (Int*)y, h, width, height); }

void ref_projective_transform(const int16_t* refsrch, int16_t* aux, const double h[9], int W, int H, int stride, int nss_window)
{ ref()->warp((Pel*)refsrch, aux, (Double*)h, W, H, stride, nss_window); }

/* TComPicYuv::extendPicBorder on a luma plane: `plane` is (pic_h+2m) x (pic_w+2m) incl. margins (in/out) */
int ref_extend_border(int16_t* plane, int pic_w, int pic_h)
{
  ref();
  TComPicYuv pic;
  pic.create(pic_w, pic_h, 64, 64, 4);
  int m = pic.getLumaMargin();
  int stride = pic.getStride();
  size_t n = (size_t)stride * (pic_h + 2 * m);
  memcpy(pic.getBufY(), plane, n * sizeof(int16_t));
  pic.setBorderExtension(false);
  pic.extendPicBorder();
  memcpy(plane, pic.getBufY(), n * sizeof(int16_t));
  pic.destroy();
  return m;
}

}  // extern "C"
