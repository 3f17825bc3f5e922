/*
 * hop_oracle.h -- CPU restatement of the HEVC-HOP hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this; the product (libhopgpu) never links, imports or executes anything under oracle/.
 *
 * Parity pin: the reference ships no tests or golden vectors.  This restatement is
 * pinned against the reference ITSELF: oracle/_ref/libhopref.so is the unmodified reference compiled
 * from /root/reference/source plus ref_harness.cpp, and tests/test_oracle_cpu.py +
 * tests/golden/ (made by tests/golden/make_golden.py from libhopref.so) hold the outputs of the real
 * xPatternSearch / xPatternSearchGT / DistFunc on seeded inputs.
 *
 * Job/result structs are the public ones of include/hop_gpu.h so results compare field by field.
 */
#ifndef HOP_ORACLE_H
#define HOP_ORACLE_H
#include "../include/hop_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* TComRdCost::xGetComponentBits, TComRdCost.cpp:270-284 */
uint32_t orc_component_bits(int32_t val);
/* TComRdCost::getBits / getCost(x,y) / getCost(b), TComRdCost.h:185-202 */
uint32_t orc_get_bits(const HopCostState* cs, int32_t x, int32_t y);
uint32_t orc_get_cost_xy(const HopCostState* cs, int32_t x, int32_t y);
uint32_t orc_get_cost_bits(const HopCostState* cs, uint32_t bits);
/* TComRdCost::getBitsGT (IT_GT_CODING 0, IT_GT_AFFINE 1, W_GT 1), TComRdCost.h:204-216 */
uint32_t orc_get_bits_gt(int32_t x0, int32_t y0, int32_t x1, int32_t y1, int32_t x2, int32_t y2);

/* SAD family, TComRdCost.cpp:513-1010 (dispatch by width as setDistParam does, :298-329) */
uint32_t orc_sad(const int16_t* org, int org_stride, const int16_t* cur, int cur_stride,
                 int cols, int rows, int sub_shift, int bit_depth);
/* xGetHADs, TComRdCost.cpp:1641-1708 with xCalcHADs2x2/4x4/8x8 :1366-1575 */
uint32_t orc_hads(const int16_t* org, int org_stride, const int16_t* cur, int cur_stride,
                  int cols, int rows, int bit_depth);
uint32_t orc_dist(const HopDistJob* job, const int16_t* org, const int16_t* cur);

/* calcParamProjective, TComPrediction.cpp:807-832 */
void orc_calc_param_projective(const int32_t x[4], const int32_t y[4], double h[9], int width, int height);
/* ProjectiveTransform (IT_GT_GRID_SIZE 2, bilinear), TComPrediction.cpp:904-1030.
 * ref points at piRefSrch (window + (cols/2, rows/2)); W,H are the 2x grid sizes. */
void orc_projective_transform(const int16_t* ref, int16_t* aux, const double h[9],
                              int W, int H, int stride, int nss_window);
/* m_filteredBlock[0][0] after xExtDIFUpSamplingH: filterCopy first+last == clamp to [0,2^bd-1]
 * (TEncSearch.cpp:7818-7837, TComInterpolationFilter.cpp:113-154). dst is w x h, stride w. */
void orc_stage_window(const int16_t* src, int src_stride, int16_t* dst, int w, int h, int bit_depth);

/* K1: xPatternSearch, TEncSearch.cpp:6262-6371 */
void orc_pattern_search(const HopSearchJob* job, const int16_t* org, const int16_t* ref,
                        HopSearchResult* out);
/* K2: xPatternSearchGT diamond branch, TEncSearch.cpp:4686-4790 + 5093-5467 */
void orc_pattern_search_gt(const HopGtJob* job, const int16_t* org, const int16_t* ref,
                           HopGtResult* out);

/* batch helpers (loop over jobs; used by the timing legs of bench.py) */
void orc_pattern_search_batch(int n, const HopSearchJob* jobs, const int16_t* org, const int16_t* ref,
                              HopSearchResult* out);
void orc_pattern_search_gt_batch(int n, const HopGtJob* jobs, const int16_t* org, const int16_t* ref,
                                 HopGtResult* out);

/* K5: xPatternSearchFracDIF, TEncSearch.cpp:6564-6610 (+ xPatternRefinement :709-761, xExtDIFUpSamplingH/Q
 * :7818-8011, TComInterpolationFilter.cpp:92-254).  orc_interp_block is the plane the reference would read
 * for the quarter-pel displacement (qx,qy) from the integer position `src`. */
void orc_interp_block(const int16_t* src, int stride, int qx, int qy, int cols, int rows, int bit_depth, int16_t* dst);
void orc_frac_search(const HopFracJob* job, const int16_t* org, const int16_t* ref, HopFracResult* out);
void orc_frac_search_batch(int n, const HopFracJob* jobs, const int16_t* org, const int16_t* ref, HopFracResult* out);
/* xMotionEstimation's GPU part in one go: K1 -> frac -> GT (TEncSearch.cpp:4572-4642) */
void orc_motion_search_batch(int n, const HopMotionJob* jobs, const int16_t* org, const int16_t* ref, HopMotionResult* out);

/* Exhaustive sweep (reference mode IT_GT_SEARCH 1 + IT_GT_GRID_SIZE 1, TEncSearch.cpp:4989-5091): the job's
 * ss_cand is pcMvInt; amvp is ignored.  The slice [cand_begin, cand_end) counts AFFINE candidates in loop
 * order (7200 for N = 2). */
uint64_t orc_gt_sweep_key(const HopGtJob* job, const int16_t* org, const int16_t* ref,
                          int cand_begin, int cand_end, uint32_t* n_scored);
void orc_gt_sweep_finalize(const HopGtJob* job, uint64_t key, HopGtResult* out);
void orc_gt_sweep_batch(int n, const HopGtJob* jobs, const int16_t* org, const int16_t* ref, HopGtResult* out);
void orc_gt_sweep_keys_batch(int n, const HopGtJob* jobs, const int16_t* org, const int16_t* ref,
                             int cand_begin, int cand_end, uint64_t* keys);

/* K6: xPredInterLumaBlk / xPredInterChromaBlk incl. the GT branches (TComPrediction.cpp:639-805, 1235-1420), and
 * xGetInterPredictionError / xGetTemplateCost on top (TEncSearch.cpp:2951-2977, 4390-4477). */
void orc_predict(const HopPredJob* job, const int16_t* org, const int16_t* ref, int16_t* dst, HopPredResult* out);
void orc_predict_batch(int n, const HopPredJob* jobs, const int16_t* org, const int16_t* ref, int16_t* dst, HopPredResult* out);

/* K7: the 35 intra predictions of predIntraLumaAng (TComPrediction.cpp:129-170, 192-348, 1468-1546) and calcHAD
 * (TComRdCost.cpp:391-425) for one PU -- the pre-screen loop of estIntraPredQT (TEncSearch.cpp:2451-2464). */
void orc_intra_predict(const int32_t* above, const int32_t* left, int n, int mode, int bit_depth, int above_avail, int left_avail, int16_t* dst);
void orc_intra_prescreen(const HopIntraJob* job, const int16_t* org, const int32_t* refs, uint32_t* out);
void orc_intra_prescreen_batch(int n, const HopIntraJob* jobs, const int16_t* org, const int32_t* refs, uint32_t* out);

/* K4: TComPicYuv::extendPicBorder luma part, TComPicYuv.cpp:236-274. plane points at sample (0,0). */
void orc_extend_border(int16_t* origin, int stride, int pic_w, int pic_h, int margin);

#ifdef __cplusplus
}
#endif
#endif
