/*
 * hop_gpu.h -- C ABI of libhopgpu, the B200 (sm_100a) implementation of the HEVC-HOP encoder hot path.
 *
 * The library replaces, inside an otherwise unchanged HM-15/HEVC-HOP encoder, the bodies of
 *
 *   TEncSearch::xPatternSearch      source/Lib/TLibEncoder/TEncSearch.cpp:6262-6371   (K1, SS full search)
 *   TEncSearch::xPatternSearchGT    source/Lib/TLibEncoder/TEncSearch.cpp:4686-5467   (K2, HOP/GT diamond search)
 *   TComRdCost::m_afpDistortFunc[]  source/Lib/TLibCommon/TComRdCost.cpp:177-221      (K3, SAD / HADs table)
 *   TEncCu::xCopyYuv2SSRef          source/Lib/TLibEncoder/TEncCu.cpp:1677-1715       (K4, SS reference update)
 *   + TComPicYuv::extendPicBorder   source/Lib/TLibCommon/TComPicYuv.cpp:236-274
 *
 * Conventions
 *   - plain C, plain pointers and sizes; no C++/torch types cross this line.
 *   - `Pel` of the reference is int16_t (TypeDef.h:300); NOT_VALID samples are -1 (CommonDef.h:126).
 *   - every function returns HOP_OK (0) or a negative HopStatus; there is no CPU fallback: when no
 *     CUDA device / kernel image is usable the call fails with HOP_ERR_CUDA (hop_last_error() has text).
 *   - the reference has no error returns (Void + exit()); the host shim (hop_shim.h) turns a non-zero
 *     status into fprintf(stderr)+exit(EXIT_FAILURE), as the reference does for its own fatal errors.
 *   - host entry points copy in/out and never retain host pointers; `_dev` entry points take device
 *     pointers plus a CUstream/cudaStream_t passed as void* (NULL = the context's own stream).
 *   - a context is used by one host thread at a time (the reference encoder is single threaded).
 */
#ifndef HOP_GPU_H
#define HOP_GPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HOP_ABI_VERSION 2

typedef enum HopStatus {
  HOP_OK            = 0,
  HOP_ERR_ARG       = -1,  /* bad argument (NULL, size out of range, unsupported block shape) */
  HOP_ERR_CUDA      = -2,  /* CUDA runtime / launch failure, or no sm_100 device               */
  HOP_ERR_NOMEM     = -3,
  HOP_ERR_STATE     = -4   /* call order violated (e.g. search before hop_ref_reset)            */
} HopStatus;

#define HOP_MAX_UINT   0xFFFFFFFFu   /* MAX_UINT, CommonDef.h */
#define HOP_NOT_VALID  (-1)          /* NOT_VALID, CommonDef.h:126 */
#define HOP_MAX_PU     64            /* largest PU edge (CTU 64) */
#define HOP_MAX_PRED   3             /* AMVP_MAX_NUM_CANDS_MEM, CommonDef.h */

/* TComMv (TComMv.h): two Shorts. */
typedef struct HopMv { int16_t hor, ver; } HopMv;

/* Motion-cost state of TComRdCost read by getCost()/getBits() (TComRdCost.h:128-136, 185-202). */
typedef struct HopCostState {
  uint32_t lambda_cost;   /* m_uiCost, i.e. m_uiLambdaMotionSAD after getMotionCost(1,0)  */
  int32_t  cost_scale;    /* m_iCostScale: 2 inside xPatternSearch, 0 inside xPatternSearchGT */
  HopMv    pred;          /* m_mvPredictor (quarter-pel)                                   */
} HopCostState;

/* ---------------------------------------------------------------------------------------------
 * K1 -- one xPatternSearch call (TEncSearch.cpp:6262-6371).
 * `org`/`ref` are sample offsets into the two buffers handed to the batch call; ref_off addresses
 * piRefY, i.e. the PU's own position in the SS reference plane.
 * ------------------------------------------------------------------------------------------- */
typedef struct HopSearchJob {
  int64_t  org_off;        /* pcPatternKey->getROIY()                          */
  int64_t  ref_off;        /* piRefY                                           */
  int32_t  org_stride;     /* pcPatternKey->getPatternLStride()                */
  int32_t  ref_stride;     /* iRefStride                                       */
  int32_t  cols, rows;     /* getROIYWidth/Height: 4..64, see hop_shape_supported */
  int32_t  rng_left, rng_top, rng_right, rng_bottom;   /* pcMvSrchRngLT / RB (integer pel) */
  int32_t  offset_x, offset_y;                         /* riOffsetX / riOffsetY            */
  int32_t  is_ss;          /* isSSE: apply the causal gate + isValidPattern    */
  int32_t  fast_enc;       /* m_pcEncCfg->getUseFastEnc(): rows>8 => iSubShift=1 (:6303-6309) */
  int32_t  bit_depth;      /* g_bitDepthY                                      */
  HopCostState cost;
} HopSearchJob;

typedef struct HopSearchResult {
  int32_t  found;          /* isValid (:6356); 0 => sad == HOP_MAX_UINT, mv untouched by the shim */
  HopMv    mv;             /* rcMv == ssBestCand[0] (IT_SS_NUMBER_OF_BEST_CAND 1)                 */
  uint32_t sad;            /* ruiSAD = best - getCost(best)                                        */
  uint32_t cost;           /* uiSadBest (SAD + motion cost) -- extra, for tests                    */
} HopSearchResult;

/* ---------------------------------------------------------------------------------------------
 * K2 -- one xPatternSearchGT call, diamond branch (TEncSearch.cpp:4686-4790, 5093-5467) with the
 * compile-time configuration the reference ships (TypeDef.h:207-240): IT_GT_AFFINE 1,
 * IT_GT_SEARCH 2, IT_GT_GRID_SIZE 2, IT_MAX_NSS_Iteration 6, IT_Independent_Iterations 1,
 * bilinear warp in IEEE binary64 (TComPrediction.cpp:807-832, 904-1030), W_GT 1, IT_GT_CODING 0.
 * ------------------------------------------------------------------------------------------- */
typedef struct HopGtJob {
  int64_t  org_off;
  int64_t  ref_off;
  int32_t  org_stride;
  int32_t  ref_stride;
  int32_t  cols, rows;
  HopMv    ss_cand;               /* bestSSCand[0], integer pel (:5116-5123)                       */
  int32_t  num_pred;              /* AMVPInfo::iN (:5101)                                          */
  HopMv    amvp[HOP_MAX_PRED];    /* AMVPInfo::m_acMvCand, raw quarter-pel (:5104, 5144-5153)      */
  uint32_t threshold;             /* ruiCost on entry = cost to beat (:4769)                       */
  int32_t  use_had;               /* m_pcEncCfg->getUseHADME() (:4772)                             */
  int32_t  bit_depth;             /* g_bitDepthY                                                   */
  HopCostState cost;              /* cost_scale is 0 here (:4619)                                  */
} HopGtJob;

typedef struct HopGtResult {
  int32_t  gt_flag;        /* gtFlag (:5441/:5461)                                              */
  HopMv    gt[4];          /* rcGT0..rcGT3 (:5448-5451), zero when !gt_flag                      */
  uint32_t cost;           /* ruiCost on exit (unchanged threshold when !gt_flag)                */
  HopMv    mv_int;         /* pcMvInt on exit (:5455); valid only when gt_flag                   */
  int32_t  best_index;     /* extra, for tests: (start*8 + pass)*64 + candidate, or -1           */
  uint32_t n_candidates;   /* extra: HOP candidates warped + scored by this call                */
} HopGtResult;

/* ---------------------------------------------------------------------------------------------
 * K5 -- one xPatternSearchFracDIF call (TEncSearch.cpp:6564-6610): half- then quarter-pel refinement of
 * the integer vector with the 8-tap DCT-IF planes (xExtDIFUpSamplingH/Q :7818-8011,
 * TComInterpolationFilter.cpp:92-254) and 2 x 9 Hadamard/SAD costs (xPatternRefinement :709-761).
 * The cost it returns is the threshold xPatternSearchGT has to beat.
 * ------------------------------------------------------------------------------------------- */
typedef struct HopFracJob {
  int64_t  org_off;
  int64_t  ref_off;        /* piRefY: the PU's own position in the reference plane                  */
  int32_t  org_stride;
  int32_t  ref_stride;
  int32_t  cols, rows;
  HopMv    mv_int;         /* *pcMvInt, integer pel                                                  */
  int32_t  use_had;        /* m_pcEncCfg->getUseHADME()                                              */
  int32_t  bit_depth;
  HopCostState cost;       /* lambda_cost and pred; cost_scale is set by the stage (1 half, 0 quarter) */
} HopFracJob;

typedef struct HopFracResult {
  HopMv    half;           /* rcMvHalf  in {-1,0,1}^2                                                */
  HopMv    qter;           /* rcMvQter  in {-1,0,1}^2                                                */
  uint32_t cost;           /* ruiCost after the quarter-pel stage                                    */
  uint32_t cost_half;      /* extra, for tests: best cost of the half-pel stage                      */
} HopFracResult;

/* ---------------------------------------------------------------------------------------------
 * The whole SS motion search of one PU in ONE call: xPatternSearch -> (valid and non-zero vector?) ->
 * xPatternSearchFracDIF -> xPatternSearchGT, i.e. the GPU part of TEncSearch::xMotionEstimation
 * (TEncSearch.cpp:4572-4642) without a host round trip between the stages.
 * ------------------------------------------------------------------------------------------- */
typedef struct HopMotionJob {
  HopSearchJob search;     /* cost.cost_scale must be 2 (:4560)                                      */
  int32_t  use_had;        /* getUseHADME(): frac refinement and GT distortion                        */
  int32_t  use_gt;         /* bUseGT (:4627)                                                          */
  int32_t  num_pred;       /* AMVPInfo::iN                                                            */
  HopMv    amvp[HOP_MAX_PRED];
} HopMotionJob;

typedef struct HopMotionResult {
  HopSearchResult search;  /* found == 0 or mv == (0,0): the later stages did not run (:4603-4611)    */
  int32_t  refined;        /* 1 when the frac (and, if use_gt, the GT) stage ran                      */
  HopFracResult   frac;
  HopGtResult     gt;      /* threshold = frac.cost, ss_cand = search.mv                              */
} HopMotionResult;

/* ---------------------------------------------------------------------------------------------
 * K3 -- one DistFunc call (TComRdCost.cpp:513-1010, 1366-1708).
 * ------------------------------------------------------------------------------------------- */
typedef enum HopDistFunc {
  HOP_DF_SAD  = 8,    /* DF_SAD  (TypeDef.h DFunc): xGetSAD{,4,8,12,16,16N,24,32,48,64} by width */
  HOP_DF_HADS = 22    /* DF_HADS: xGetHADs (8x8 / 4x4 / 2x2 tiles)                                */
} HopDistFunc;

typedef struct HopDistJob {
  int64_t  org_off, cur_off;
  int32_t  org_stride, cur_stride;
  int32_t  cols, rows;
  int32_t  func;           /* HopDistFunc */
  int32_t  sub_shift;      /* DistParam::iSubShift (SAD only) */
  int32_t  bit_depth;
} HopDistJob;

/* ---------------------------------------------------------------------------------------------
 * K6 -- motion-compensated prediction of one PU from one reference (uni-prediction, bi = false) and what the
 * encoder computes from it:
 *   TComPrediction::xPredInterLumaBlk / xPredInterChromaBlk  (TLibCommon/TComPrediction.cpp:639-720, 1235-1347)
 *     - no GT: 8-tap / 4-tap DCT-IF interpolation at the vector's fraction (TComInterpolationFilter.cpp:92-420)
 *     - GT   : the 2W x 2H region around the block is fetched (interpolated when the vector is fractional) and
 *              warped by xPredGTLuma / xPredGTChroma (:723-805, 1351-1420) with calcParamProjective[C] (:807-859)
 *              and ProjectiveTransform (:904-1030); chroma uses the half GT vectors in binary64
 *   dist_func != 0: the distortion of the prediction against the original block, setDistParam(.., bHadamard) +
 *              DistFunc as TEncSearch::xGetInterPredictionError (TLibEncoder/TEncSearch.cpp:2951-2977)
 *   template != 0 : TEncSearch::xGetTemplateCost (:4390-4477): isValidPattern gate on the unclipped candidate
 *              (TComRdCost.cpp:430-442), luma prediction at the clipped candidate, SAD, calcRdCost(bits, sad, DF_SAD)
 * ------------------------------------------------------------------------------------------- */
typedef struct HopPredJob {
  int64_t  ref_off;        /* the PU's own position in the reference plane of `comp` (luma, Cb or Cr plane)  */
  int64_t  org_off;        /* original block of that component (dist / template), ignored otherwise           */
  int64_t  dst_off;        /* where the prediction goes in the dst buffer; < 0: not stored                     */
  int32_t  ref_stride, org_stride, dst_stride;
  int32_t  cols, rows;     /* LUMA size of the PU; a chroma job predicts cols/2 x rows/2                       */
  int32_t  comp;           /* 0 = luma, 1 = chroma (either chroma plane)                                        */
  HopMv    mv;             /* quarter-pel luma vector, already clipped (TComDataCU::clipMv is host code)        */
  int32_t  gt_flag;        /* bUseGT; all-zero GT vectors take the plain branch as in the reference             */
  HopMv    gt[4];          /* mGT0..mGT3                                                                        */
  int32_t  bit_depth;      /* g_bitDepthY / g_bitDepthC                                                         */
  int32_t  dist_func;      /* 0 = none, HOP_DF_SAD, HOP_DF_HADS                                                 */
  int32_t  template_cost;  /* 1 = xGetTemplateCost semantics (luma, no GT)                                      */
  int32_t  is_ss;          /* template: the reference is the SS reference => isValidPattern gate applies        */
  HopMv    mv_probe;       /* template: cMvCand before clipMv                                                   */
  uint32_t mvp_bits;       /* template: m_auiMVPIdxCost[iMVPIdx][iMVPNum]                                       */
  uint32_t lambda_sad;     /* template: m_uiLambdaMotionSAD                                                     */
} HopPredJob;

typedef struct HopPredResult {
  int32_t  valid;          /* template: 0 when the gate rejected the candidate (cost = MAX_INT then); else 1    */
  uint32_t dist;           /* DistFunc result (0 when dist_func == 0 and template_cost == 0)                    */
  uint32_t cost;           /* template: the returned uiCost; otherwise == dist                                  */
} HopPredResult;

/* ---------------------------------------------------------------------------------------------
 * K7 -- intra mode pre-screen of one PU: the 35 predictions of TComPrediction::predIntraLumaAng
 * (TLibCommon/TComPrediction.cpp:316-348: xPredIntraPlanar :1468-1510, xPredIntraAng :192-314 with the DC value
 * :129-170, the edge filters of the pure horizontal / vertical modes and xDCPredFiltering :1524-1546 for blocks up
 * to 16x16) and their Hadamard cost TComRdCost::calcHAD (TLibCommon/TComRdCost.cpp:391-425) against the original
 * block -- the loop TEncSearch::estIntraPredQT runs per PU (TLibEncoder/TEncSearch.cpp:2451-2464).  The mode bits
 * (CABAC state) and the candidate list stay on the host.
 * Reference samples: what TComPattern::initAdiPattern left in m_piYuvExt, border row / column only, int32:
 *   refs[refs_off + 0*(2N+1) + k] = unfiltered above  (k = 0 is the corner sample, k = 1..2N the row above)
 *   refs[refs_off + 1*(2N+1) + k] = unfiltered left   (k = 0 the corner again, k = 1..2N the column to the left)
 *   refs[refs_off + 2*(2N+1) + k], [.. + 3*(2N+1) + k] = the same two from the filtered buffer
 * (TComPattern::getPredictorPtr :583-607 picks filtered / unfiltered per mode and size.)
 * ------------------------------------------------------------------------------------------- */
#define HOP_INTRA_MODES 35
typedef struct HopIntraJob {
  int64_t  org_off;        /* original block, size x size                                                  */
  int64_t  refs_off;       /* into the int32 refs buffer, 4 * (2 * size + 1) entries                        */
  int32_t  org_stride;
  int32_t  size;           /* N = 4, 8, 16, 32 or 64                                                        */
  int32_t  above_avail;    /* bAbove / bLeft of initAdiPattern: only the DC mode reads them                  */
  int32_t  left_avail;
  int32_t  bit_depth;
  int32_t  reserved;
} HopIntraJob;

/* ---------------------------------------------------------------------------------------------
 * Context: one per encoder instance = per GPU.  Owns a stream, scratch buffers and (optionally) the
 * device mirror of the SS reference luma plane (TEncTop::m_cSSRef / TComPicYuv, margin 80).
 * ------------------------------------------------------------------------------------------- */
typedef struct HopCtx HopCtx;

int         hop_abi_version(void);
const char* hop_last_error(void);                 /* thread-local text of the last failure */
int         hop_device_count(void);               /* number of CUDA devices, <0 on error    */

int  hop_ctx_create(int device, HopCtx** out);
void hop_ctx_destroy(HopCtx* ctx);
int  hop_ctx_sync(HopCtx* ctx);
void* hop_ctx_stream(HopCtx* ctx);                /* the context's cudaStream_t */

/* 1 if (cols,rows) is a PU shape the kernels accept: both in {4,8,12,16,24,32,48,64}, not 4x4. */
int  hop_shape_supported(int cols, int rows);

/* SS reference mirror (K4).  pic_w/pic_h luma size, margin as TComPicYuv (80), stride = pic_w+2*margin. */
int  hop_ref_create(HopCtx* ctx, int pic_w, int pic_h, int margin);
int  hop_ref_reset(HopCtx* ctx, int value);       /* setPicPel(NOT_VALID), TComSlice.cpp:253 */
/* upload a whole host plane incl. margins ((pic_h+2m) x stride samples), e.g. right after the reset */
int  hop_ref_upload(HopCtx* ctx, const int16_t* plane, size_t plane_samples);
/* copy a w x h block of reconstruction to (x,y) and re-extend the borders (TEncCu.cpp:1694-1696) */
int  hop_ref_update(HopCtx* ctx, int x, int y, int w, int h, const int16_t* src, int src_stride);
/* read back the whole plane incl. margins ((pic_h+2m) x stride samples) -- tests / debugging */
int  hop_ref_download(HopCtx* ctx, int16_t* dst, size_t dst_samples);
int  hop_ref_stride(HopCtx* ctx);
/* device pointer of sample (0,0) of the mirror, for the _dev entry points */
const int16_t* hop_ref_origin_dev(HopCtx* ctx);

/* ---- host entry points (copy in, run, copy out, synchronous) -------------------------------
 * `org`/`ref` are HOST buffers of org_samples / ref_samples int16 samples; job offsets index them.
 * ref == NULL means "the context's SS reference mirror" (offsets are then relative to its sample
 * (0,0) and may be negative into the margin). */
int hop_pattern_search_batch(HopCtx* ctx, int n, const HopSearchJob* jobs,
                             const int16_t* org, size_t org_samples,
                             const int16_t* ref, size_t ref_samples,
                             HopSearchResult* out);
int hop_pattern_search_gt_batch(HopCtx* ctx, int n, const HopGtJob* jobs,
                                const int16_t* org, size_t org_samples,
                                const int16_t* ref, size_t ref_samples,
                                HopGtResult* out);
/* Asynchronous form of hop_pattern_search_gt_batch for callers that stream many batches: the call only
 * enqueues (input copies on a copy stream, kernel and result copy on the context stream) and returns; up
 * to HOP_ASYNC_SLOTS batches are in flight, so the input copy of batch k+1 overlaps the kernel of batch k.
 * jobs/org/ref/out must stay valid (and should be page-locked) until hop_ctx_sync() returns. */
#define HOP_ASYNC_SLOTS 4
int hop_pattern_search_gt_batch_async(HopCtx* ctx, int n, const HopGtJob* jobs,
                                      const int16_t* org, size_t org_samples,
                                      const int16_t* ref, size_t ref_samples,
                                      HopGtResult* out);
int hop_dist_batch(HopCtx* ctx, int n, const HopDistJob* jobs,
                   const int16_t* org, size_t org_samples,
                   const int16_t* cur, size_t cur_samples,
                   uint32_t* out);
int hop_frac_search_batch(HopCtx* ctx, int n, const HopFracJob* jobs,
                          const int16_t* org, size_t org_samples,
                          const int16_t* ref, size_t ref_samples,
                          HopFracResult* out);
/* K6.  `ref` is the plane (of the component) the jobs' ref_off index; `dst` (dst_samples int16) receives the
 * predictions of the jobs with dst_off >= 0 and may be NULL when no job stores one.  ref == NULL: luma jobs read the
 * context's SS mirror. */
int hop_predict_batch(HopCtx* ctx, int n, const HopPredJob* jobs,
                      const int16_t* org, size_t org_samples,
                      const int16_t* ref, size_t ref_samples,
                      int16_t* dst, size_t dst_samples, HopPredResult* out);
/* K7.  out receives HOP_INTRA_MODES values per job (mode 0 = planar, 1 = DC, 2..34 angular): uiSad of :2458. */
int hop_intra_prescreen_batch(HopCtx* ctx, int n, const HopIntraJob* jobs,
                              const int16_t* org, size_t org_samples,
                              const int32_t* refs, size_t refs_count, uint32_t* out);
int hop_motion_search_batch(HopCtx* ctx, int n, const HopMotionJob* jobs,
                            const int16_t* org, size_t org_samples,
                            const int16_t* ref, size_t ref_samples,
                            HopMotionResult* out);
/* Speculative form for the encoder (the first PU of every partition mode of a CU searches the same
 * SS-reference state, TEncCu.cpp:456-633): enqueue the fused motion search of each job against the SS mirror on a
 * side stream and return immediately.  A later hop_motion_search_batch(n = 1, ref = NULL) whose request is
 * IDENTICAL (every job field, every sample of the original block, SS mirror not modified in between) returns that
 * launch's result instead of launching again; any other request takes the normal path, so correctness never
 * depends on what was predicted.  Up to HOP_PREFETCH_SLOTS requests are kept; older unclaimed ones are dropped. */
#define HOP_PREFETCH_SLOTS 15
int hop_motion_search_prefetch(HopCtx* ctx, int n, const HopMotionJob* jobs,
                               const int16_t* org, size_t org_samples);

/* ---- device entry points (everything already resident in HBM, asynchronous on `stream`) ---- */
/* cols / rows / nx_max / ny_max: when every job of the batch has this PU shape (8-bit content) and a search window
 * of at most nx_max x ny_max positions, the per-width throughput kernel runs with exactly the shared memory that
 * needs; pass zeros for mixed batches.  Batched searches of one context are issued on one stream at a time. */
int hop_pattern_search_batch_dev(HopCtx* ctx, int n, const HopSearchJob* d_jobs,
                                 const int16_t* d_org, const int16_t* d_ref,
                                 HopSearchResult* d_out, int cols, int rows, int nx_max, int ny_max, void* stream);
/* max_cols/max_rows: upper bound of the PU shapes in d_jobs (sizes the shared-memory window and CTA).
 * ref_samples: addressable int16 samples behind d_ref -- start vectors are raw AMVP vectors whose 2W x 2H window may
 * leave the buffer (the reference then reads whatever lies beyond its plane); the kernels keep every read inside
 * [d_ref, d_ref + ref_samples).  Ignored (mirror geometry used) when d_ref == hop_ref_origin_dev(ctx). */
int hop_pattern_search_gt_batch_dev(HopCtx* ctx, int n, const HopGtJob* d_jobs,
                                    const int16_t* d_org, const int16_t* d_ref, size_t ref_samples,
                                    HopGtResult* d_out, int max_cols, int max_rows, void* stream);
int hop_dist_batch_dev(HopCtx* ctx, int n, const HopDistJob* d_jobs,
                       const int16_t* d_org, const int16_t* d_cur,
                       uint32_t* d_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Exhaustive HOP parameter sweep -- the reference's compile-time search mode IT_GT_SEARCH 1
 * ("square unit search", N = 2: 25^4 corner sets around the initial rectangle) as it builds with
 * IT_GT_GRID_SIZE 1 (TEncSearch.cpp:4989-5091, TypeDef.h:216,228).  Job = HopGtJob with ss_cand = pcMvInt
 * (half = quarter = 0; amvp ignored).  The HOP_SWEEP_CANDS affine corner sets are indexed in loop order;
 * a slice [cand_begin, cand_end) yields per PU the partial argmin key (cost << 32 | flat loop index,
 * all-ones when nothing scored) -- the words a multi-GPU run min-reduces (ncclMin) before finalize.
 * ------------------------------------------------------------------------------------------- */
#define HOP_SWEEP_CANDS 7200
int hop_gt_sweep_keys_dev(HopCtx* ctx, int n, const HopGtJob* d_jobs, const int16_t* d_org, const int16_t* d_ref,
                          size_t ref_samples, int max_cols, int max_rows, int cand_begin, int cand_end,
                          uint64_t* d_keys, uint32_t* d_counts /* may be NULL */, void* stream);
int hop_gt_sweep_finalize_dev(HopCtx* ctx, int n, const HopGtJob* d_jobs, const uint64_t* d_keys,
                              const uint32_t* d_counts /* may be NULL */, HopGtResult* d_out, void* stream);
/* Sharded sweep without a collective call (one process per GPU, all GPUs of one NVLink / NVSwitch node): every rank
 * keeps one merge word per PU in its own HBM and exports it (CUDA IPC); hop_gt_sweep_sharded_dev scores this rank's
 * slice of the candidate range and its kernel min-reduces the partial keys into EVERY rank's merge words through
 * peer-mapped memory (64-bit atomicMin over NVLink, issued as each PU completes, overlapped with the CTAs still
 * computing); a second kernel waits for all ranks' arrival counters and finalises.  Every rank calls with the same
 * n / jobs / inputs and obtains the same results as hop_gt_sweep_batch.  Set-up: create -> exchange the handles of
 * all ranks by any means (e.g. torch.distributed.all_gather_object) -> connect. */
typedef struct HopSweepHandle { unsigned char bytes[64]; } HopSweepHandle;
int hop_sweep_exchange_create(HopCtx* ctx, int max_pus, HopSweepHandle* mine);
int hop_sweep_exchange_connect(HopCtx* ctx, int world, int rank, const HopSweepHandle* all /* [world] */);
int hop_gt_sweep_sharded_dev(HopCtx* ctx, int n, const HopGtJob* d_jobs, const int16_t* d_org, const int16_t* d_ref,
                             size_t ref_samples, int max_cols, int max_rows, HopGtResult* d_out, void* stream);
/* host convenience: whole candidate range on this GPU */
int hop_gt_sweep_batch(HopCtx* ctx, int n, const HopGtJob* jobs, const int16_t* org, size_t org_samples,
                       const int16_t* ref, size_t ref_samples, HopGtResult* out);

/* Number of kernel launches issued through this context so far (bench.py's gpu_launches). */
uint64_t hop_ctx_launch_count(HopCtx* ctx);

/* Counters of the single-PU (in-encoder) path since the context was created. */
typedef struct HopCtxStats {
  uint64_t single_calls;      /* n = 1 calls against the SS mirror (search, GT or fused motion search)           */
  uint64_t cache_hits;        /* ... of the fused kind answered by a speculative search                            */
  uint64_t cache_misses;      /* ... that launched their own kernels                                               */
  uint64_t prefetched;        /* speculative searches enqueued by hop_motion_search_prefetch                       */
  uint64_t prefetch_dropped;  /* ... whose result was never asked for and whose slot was reused                    */
  uint64_t candidates;        /* HOP candidates warped + scored for the single-PU calls that returned (n_candidates) */
} HopCtxStats;
int hop_ctx_stats(HopCtx* ctx, HopCtxStats* out);

/* ALU peak probes for the roofline: run a dependent-free instruction loop on every
 * SM and return achieved Gop/s (lane-operations) for the named pipe.
 * what: 0 = int32 IADD3, 1 = VABSDIFF4-accumulate (4 byte-SADs per lane-op), 2 = fp64 DADD,
 *       3 = fp64 DMUL, 4 = fp64 DFMA, 5 = int32 LOP3/SHF mix. */
int hop_probe_alu(HopCtx* ctx, int what, double* gops_out, double* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* HOP_GPU_H */
