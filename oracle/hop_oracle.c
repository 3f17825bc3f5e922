/*
 * hop_oracle.c -- CPU restatement of the HEVC-HOP hot path.  TEST INFRASTRUCTURE ONLY (see hop_oracle.h).
 *
 * Each function cites the reference lines (relative to /root/reference/source/Lib) it restates.
 * Build: gcc -O2 -ffp-contract=off, x86-64 baseline (SSE2 doubles, no FMA) -- the warp is IEEE
 * binary64 in the reference (build/linux/common/makefile.base:50,66: -O3, no -march) and must not be
 * contracted.  Parity pin: tests/test_oracle_cpu.py (+ test_frac_motion.py, test_sweep.py, test_predict.py) and tests/golden/ compare this file with the
 * compiled reference itself (oracle/_ref/libhopref.so).
 */
#include "hop_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define DIST_SHIFT(bit_depth) ((bit_depth) - 8)   /* DISTORTION_PRECISION_ADJUSTMENT, TypeDef.h:162-167 */

/* ------------------------------------------------------------------------------------------------
 * bit costs -- TLibCommon/TComRdCost.cpp:270-284, TComRdCost.h:185-216
 * ---------------------------------------------------------------------------------------------- */
uint32_t orc_component_bits(int32_t val)
{
  uint32_t length = 1;
  uint32_t temp = (val <= 0) ? (uint32_t)((-val << 1) + 1) : (uint32_t)(val << 1);
  while (temp != 1) { temp >>= 1; length += 2; }
  return length;
}

uint32_t orc_get_bits(const HopCostState* cs, int32_t x, int32_t y)
{
  /* FIX203 branch, TComRdCost.h:196-199: (x << m_iCostScale) - m_mvPredictor */
  return orc_component_bits((x << cs->cost_scale) - cs->pred.hor) +
         orc_component_bits((y << cs->cost_scale) - cs->pred.ver);
}

uint32_t orc_get_cost_xy(const HopCostState* cs, int32_t x, int32_t y)
{
  return (cs->lambda_cost * orc_get_bits(cs, x, y)) >> 16;   /* UInt arithmetic, TComRdCost.h:187 */
}

uint32_t orc_get_cost_bits(const HopCostState* cs, uint32_t bits)
{
  return (cs->lambda_cost * bits) >> 16;                      /* TComRdCost.h:193 */
}

uint32_t orc_get_bits_gt(int32_t x0, int32_t y0, int32_t x1, int32_t y1, int32_t x2, int32_t y2)
{
  /* IT_GT_CODING 0, IT_GT_AFFINE 1 (GT3 not coded), W_GT 1 -- TComRdCost.h:204-216 */
  return 1u * (orc_component_bits(x0) + orc_component_bits(y0) + orc_component_bits(x1) +
               orc_component_bits(y1) + orc_component_bits(x2) + orc_component_bits(y2));
}

/* ------------------------------------------------------------------------------------------------
 * SAD -- TComRdCost.cpp:513-1010.  setDistParam (:298-329) picks xGetSAD4/8/16/32/64 through
 * g_aucConvertToBit (powers of two), xGetSAD12/24/48 by explicit test, else the generic xGetSAD,
 * which ignores iSubShift (:513-539).
 * ---------------------------------------------------------------------------------------------- */
static int sad_width_has_subshift(int cols)
{
  return cols == 4 || cols == 8 || cols == 16 || cols == 32 || cols == 64 ||
         cols == 12 || cols == 24 || cols == 48;
}

uint32_t orc_sad(const int16_t* org, int org_stride, const int16_t* cur, int cur_stride,
                 int cols, int rows, int sub_shift, int bit_depth)
{
  uint32_t sum = 0;
  if (!sad_width_has_subshift(cols)) sub_shift = 0;
  int step = 1 << sub_shift;
  for (int r = rows; r != 0; r -= step) {
    for (int n = 0; n < cols; n++) sum += (uint32_t)abs((int)org[n] - (int)cur[n]);
    org += org_stride * step;
    cur += cur_stride * step;
  }
  sum <<= sub_shift;
  return sum >> DIST_SHIFT(bit_depth);
}

/* ------------------------------------------------------------------------------------------------
 * Hadamard -- TComRdCost.cpp:1366-1575 (tiles), :1641-1708 (xGetHADs).  The reference's butterfly
 * networks are full 2-D Walsh-Hadamard transforms; sum(|coeff|) does not depend on coefficient order
 * or sign conventions, only on the per-tile rounding: 8x8 (sad+2)>>2, 4x4 (satd+1)>>1, 2x2 none.
 * ---------------------------------------------------------------------------------------------- */
static uint32_t had_tile(const int16_t* org, const int16_t* cur, int so, int sc, int n)
{
  int d[64];
  for (int y = 0; y < n; y++)
    for (int x = 0; x < n; x++) d[y * n + x] = (int)org[y * so + x] - (int)cur[y * sc + x];
  for (int y = 0; y < n; y++)                 /* rows */
    for (int len = 1; len < n; len <<= 1)
      for (int i = 0; i < n; i += len << 1)
        for (int j = i; j < i + len; j++) {
          int a = d[y * n + j], b = d[y * n + j + len];
          d[y * n + j] = a + b; d[y * n + j + len] = a - b;
        }
  for (int x = 0; x < n; x++)                 /* columns */
    for (int len = 1; len < n; len <<= 1)
      for (int i = 0; i < n; i += len << 1)
        for (int j = i; j < i + len; j++) {
          int a = d[j * n + x], b = d[(j + len) * n + x];
          d[j * n + x] = a + b; d[(j + len) * n + x] = a - b;
        }
  int satd = 0;
  for (int k = 0; k < n * n; k++) satd += abs(d[k]);
  if (n == 8) return (uint32_t)((satd + 2) >> 2);
  if (n == 4) return (uint32_t)((satd + 1) >> 1);
  return (uint32_t)satd;
}

uint32_t orc_hads(const int16_t* org, int org_stride, const int16_t* cur, int cur_stride,
                  int cols, int rows, int bit_depth)
{
  uint32_t sum = 0;
  int n;
  if ((rows % 8 == 0) && (cols % 8 == 0)) n = 8;
  else if ((rows % 4 == 0) && (cols % 4 == 0)) n = 4;
  else if ((rows % 2 == 0) && (cols % 2 == 0)) n = 2;
  else return HOP_MAX_UINT;  /* assert(false) in the reference */
  for (int y = 0; y < rows; y += n)
    for (int x = 0; x < cols; x += n)
      sum += had_tile(org + y * org_stride + x, cur + y * cur_stride + x, org_stride, cur_stride, n);
  return sum >> DIST_SHIFT(bit_depth);
}

uint32_t orc_dist(const HopDistJob* job, const int16_t* org, const int16_t* cur)
{
  const int16_t* o = org + job->org_off;
  const int16_t* c = cur + job->cur_off;
  if (job->func == HOP_DF_HADS)
    return orc_hads(o, job->org_stride, c, job->cur_stride, job->cols, job->rows, job->bit_depth);
  return orc_sad(o, job->org_stride, c, job->cur_stride, job->cols, job->rows, job->sub_shift,
                 job->bit_depth);
}

/* ------------------------------------------------------------------------------------------------
 * K1 -- TLibEncoder/TEncSearch.cpp:6262-6371 (+ isValidPattern TComRdCost.cpp:444-458)
 * ---------------------------------------------------------------------------------------------- */
void orc_pattern_search(const HopSearchJob* job, const int16_t* org_buf, const int16_t* ref_buf,
                        HopSearchResult* out)
{
  const int16_t* org = org_buf + job->org_off;
  const int16_t* ref_y = ref_buf + job->ref_off;
  const int stride = job->ref_stride;
  uint32_t sad_best = HOP_MAX_UINT;
  int best_x = 0, best_y = 0;
  int is_valid = 0;
  int sub_shift = 0;
  if (job->fast_enc && job->rows > 8) sub_shift = 1;          /* :6303-6309 */

  ref_y += job->rng_top * stride;                              /* :6311 */
  for (int y = job->rng_top; y <= job->rng_bottom; y++) {
    for (int x = job->rng_left; x <= job->rng_right; x++) {
      const int16_t* srch = ref_y + x;
      uint32_t sad = orc_sad(org, job->org_stride, srch, stride, job->cols, job->rows, sub_shift,
                             job->bit_depth);                   /* :6323, before the gates */
      if (job->is_ss) {
        if ((x >= job->offset_x) && (y > job->offset_y)) continue;            /* :6328 */
        const int16_t* lb = srch + (job->rows + 4) * stride;                  /* isValidPattern */
        const int16_t* rb = lb + (job->cols + 4);
        if (!((*lb != HOP_NOT_VALID) && (*rb != HOP_NOT_VALID))) continue;    /* :6330 */
      }
      is_valid = 1;
      sad += orc_get_cost_xy(&job->cost, x, y);                 /* :6336 */
      if (sad < sad_best) { sad_best = sad; best_x = x; best_y = y; }         /* :6338-6349 */
    }
    ref_y += stride;
  }
  memset(out, 0, sizeof(*out));
  if (!is_valid) {                                              /* :6356-6360 */
    out->found = 0; out->sad = HOP_MAX_UINT; out->cost = HOP_MAX_UINT;
    return;
  }
  out->found = 1;
  out->mv.hor = (int16_t)best_x; out->mv.ver = (int16_t)best_y;
  out->cost = sad_best;
  out->sad = sad_best - orc_get_cost_xy(&job->cost, best_x, best_y);          /* :6365 */
}

/* ------------------------------------------------------------------------------------------------
 * warp -- TLibCommon/TComPrediction.cpp:807-832, 904-1030
 * ---------------------------------------------------------------------------------------------- */
void orc_calc_param_projective(const int32_t x[4], const int32_t y[4], double h[9], int width, int height)
{
  double H, W, dx[4], dy[4];
  W = (double)width - 1.0;
  H = (double)height - 1.0;
  dx[1] = (double)x[1] - x[2];
  dx[2] = (double)x[3] - x[2];
  dx[3] = (double)x[0] - x[1] + x[2] - x[3];
  dy[1] = (double)y[1] - y[2];
  dy[2] = (double)y[3] - y[2];
  dy[3] = (double)y[0] - y[1] + y[2] - y[3];
  h[2] = ((dx[3] * dy[2] - dx[2] * dy[3]) / (dx[1] * dy[2] - dx[2] * dy[1])) / W;
  h[5] = ((dx[1] * dy[3] - dx[3] * dy[1]) / (dx[1] * dy[2] - dx[2] * dy[1])) / H;
  h[0] = (double)(x[1] - x[0]) / W + h[2] * x[1];
  h[3] = (double)(x[3] - x[0]) / H + h[5] * x[3];
  h[6] = (double)x[0];
  h[1] = (double)(y[1] - y[0]) / W + h[2] * y[1];
  h[4] = (double)(y[3] - y[0]) / H + h[5] * y[3];
  h[7] = (double)y[0];
  h[8] = 1.0;
}

void orc_projective_transform(const int16_t* ref, int16_t* aux_out, const double h[9],
                              int W, int H, int stride, int nss_window)
{
  /* IT_GT_GRID_SIZE == 2, IT_GT_Interpolation_Filter == 0 */
  const int G = 2;
  int off_x = W / 2 - (W / G / 2);
  int off_y = H / 2 - (H / G / 2);
  for (int y = off_y; y < off_y + H / G; y++) {
    for (int x = off_x; x < off_x + W / G; x++) {
      double Fx = (h[0] * x + h[3] * y + h[6]) / (h[2] * x + h[5] * y + h[8]);
      double Fy = (h[1] * x + h[4] * y + h[7]) / (h[2] * x + h[5] * y + h[8]);
      int Y = (int)Fy - off_y;
      int X = (int)Fx - off_x;
      double q = (Fy - off_y - (double)Y);
      double p = (Fx - off_x - (double)X);
      if (Y < -nss_window / G) Y = -nss_window / G;
      if (X < -nss_window / G) X = -nss_window / G;
      if (Y > nss_window / G + H / G - 1) Y = nss_window / G + H / G - 1;
      if (X > nss_window / G + W / G - 1) X = nss_window / G + W / G - 1;
      if (Y + 1 > nss_window / G + H / G - 1) Y = nss_window / G + H / G - 2;
      if (X + 1 > nss_window / G + W / G - 1) X = nss_window / G + W / G - 2;
      const int16_t* pa = ref + Y * stride;
      double aux = (1.0 - q) * ((1.0 - p) * (double)(pa[X]) + p * (double)(pa[X + 1]));
      pa = ref + (Y + 1) * stride;
      aux += q * ((1.0 - p) * (double)(pa[X]) + p * (double)(pa[X + 1]));
      if (aux > 255) aux = 255;     /* hard-coded 8-bit clip, also for 10-bit input (:969-972) */
      if (aux < 0) aux = 0;
      aux_out[x - off_x] = (int16_t)(aux + 0.5);
    }
    aux_out += W / G;
  }
}

void orc_stage_window(const int16_t* src, int src_stride, int16_t* dst, int w, int h, int bit_depth)
{
  /* filterCopy(isFirst) then filterCopy(isLast): TComInterpolationFilter.cpp:113-154.
   * Short arithmetic kept literally; equals clamp(src, 0, 2^bd-1) for src in [-1, 2^bd-1]. */
  int shift = 14 - bit_depth;                         /* IF_INTERNAL_PREC - bitDepth */
  int16_t offset = (int16_t)(1 << 13);                /* IF_INTERNAL_OFFS */
  int16_t offset_last = (int16_t)(offset + (shift ? (1 << (shift - 1)) : 0));
  int16_t max_val = (int16_t)((1 << bit_depth) - 1);
  for (int r = 0; r < h; r++) {
    for (int c = 0; c < w; c++) {
      int16_t v = (int16_t)(src[c] << shift);
      v = (int16_t)(v - offset);
      int16_t t = (int16_t)((v + offset_last) >> shift);
      if (t < 0) t = 0;
      if (t > max_val) t = max_val;
      dst[c] = t;
    }
    src += src_stride;
    dst += w;
  }
}

/* ------------------------------------------------------------------------------------------------
 * K2 -- TLibEncoder/TEncSearch.cpp:4686-4790 (prologue), 5093-5467 (diamond branch)
 * ---------------------------------------------------------------------------------------------- */
void orc_pattern_search_gt(const HopGtJob* job, const int16_t* org_buf, const int16_t* ref_buf,
                           HopGtResult* out)
{
  const int cols = job->cols, rows = job->rows;
  const int G = 2;                                   /* IT_GT_GRID_SIZE */
  const int16_t* org = org_buf + job->org_off;
  const int16_t* ref_y = ref_buf + job->ref_off;
  const int ref_stride = job->ref_stride;
  const int win_w = cols * 2, win_h = rows * 2;
  int16_t* window = (int16_t*)malloc(sizeof(int16_t) * win_w * win_h);   /* m_filteredBlock[0][0] */
  int16_t* aux = (int16_t*)malloc(sizeof(int16_t) * rows * cols);        /* piAux :4749 */
  double proj[9];
  int best_cx[4] = {0, 0, 0, 0}, best_cy[4] = {0, 0, 0, 0};              /* :4776 */
  int cur_cx[4], cur_cy[4];
  int best_nx[4], best_ny[4], cur_nx[4], cur_ny[4];
  const int max_iter = 6;                                                 /* IT_MAX_NSS_Iteration */
  int nss_window = ((rows < cols) ? rows : cols) >> 1;                    /* :4756 */
  nss_window *= G;                                                        /* :4758 */
  int last_step = nss_window >> max_iter;                                 /* :4763 */
  if (last_step == 0) last_step = 1;
  uint32_t dist, dist_best = job->threshold;                              /* :4769 */
  int best_ss_x = 0, best_ss_y = 0;                                       /* :5096-5097 */
  int best_index = -1;
  uint32_t n_cand = 0;

  int num_pred = job->num_pred;
  for (int b = 0; b < 1 + num_pred; b++) {                                /* :5106-5110 */
    int16_t Hor, Ver;
    int i_offset;
    if (b < 1) {
      if (job->ss_cand.hor == 0 && job->ss_cand.ver == 0) continue;       /* :5116 */
      Ver = job->ss_cand.ver; Hor = job->ss_cand.hor;
      i_offset = job->ss_cand.hor - cols / 2 + (job->ss_cand.ver - rows / 2) * ref_stride;   /* :5120 */
      Ver = (int16_t)(Ver << 2); Hor = (int16_t)(Hor << 2);               /* :5122-5123 */
    } else {
      HopMv e = job->amvp[b - 1];
      if (e.hor == 0 && e.ver == 0) continue;                             /* :5144 */
      Ver = e.ver; Hor = e.hor;
      Ver = (int16_t)(Ver >> 2); Hor = (int16_t)(Hor >> 2);               /* :5148-5149 */
      i_offset = Hor - cols / 2 + (Ver - rows / 2) * ref_stride;          /* :5150 */
      Ver = (int16_t)(Ver << 2); Hor = (int16_t)(Hor << 2);               /* :5152-5153 */
    }
    /* :5161-5165 -> m_filteredBlock[0][0] = clamped copy of the 2W x 2H window */
    orc_stage_window(ref_y + i_offset, ref_stride, window, win_w, win_h, job->bit_depth);
    const int16_t* ref_srch = window + (cols / 2) + (rows / 2) * win_w;   /* :5177 */

    int iter = 1;
    int pass = 0;
    for (int j0 = nss_window; (j0 > 1) && (iter <= max_iter); j0 /= 2, pass++) {   /* :5181 */
      iter++;
      if (j0 == nss_window) {                                             /* :5183-5203 */
        cur_nx[0] = best_nx[0] = 0;            cur_ny[0] = best_ny[0] = 0;
        cur_nx[1] = best_nx[1] = cols * G - 1; cur_ny[1] = best_ny[1] = 0;
        cur_nx[2] = best_nx[2] = cols * G - 1; cur_ny[2] = best_ny[2] = rows * G - 1;
        cur_nx[3] = best_nx[3] = 0;            cur_ny[3] = best_ny[3] = rows * G - 1;
      } else {                                                            /* :5204-5214 */
        for (int k = 0; k < 4; k++) { cur_nx[k] = best_nx[k]; cur_ny[k] = best_ny[k]; }
      }
      const int s = j0 / 2;
      int cand = 0;   /* index among the affine candidates of this pass, in loop order */
      /* 8 nested loops, each (s, 0, -s), outermost y0 .. innermost x3; per corner only the diamond
       * points (x==0 || y==0) are visited (:5217-5287) */
      for (int y0 = s; y0 >= -s; y0 -= s) { cur_cy[0] = cur_ny[0] + y0;
      for (int x0 = s; x0 >= -s; x0 -= s) { if (!(y0 == 0 || x0 == 0)) continue; cur_cx[0] = cur_nx[0] + x0;
      for (int y1 = s; y1 >= -s; y1 -= s) { cur_cy[1] = cur_ny[1] + y1;
      for (int x1 = s; x1 >= -s; x1 -= s) { if (!(y1 == 0 || x1 == 0)) continue; cur_cx[1] = cur_nx[1] + x1;
      for (int y2 = s; y2 >= -s; y2 -= s) { cur_cy[2] = cur_ny[2] + y2;
      for (int x2 = s; x2 >= -s; x2 -= s) { if (!(y2 == 0 || x2 == 0)) continue; cur_cx[2] = cur_nx[2] + x2;
      for (int y3 = s; y3 >= -s; y3 -= s) { cur_cy[3] = cur_ny[3] + y3;
      for (int x3 = s; x3 >= -s; x3 -= s) { if (!(y3 == 0 || x3 == 0)) continue; cur_cx[3] = cur_nx[3] + x3;
        if (x0 == x1 && x0 == x2 && x0 == x3 && y0 == y1 && y0 == y2 && y0 == y3) continue;  /* :5289 */
        orc_calc_param_projective(cur_cx, cur_cy, proj, cols * G, rows * G);                 /* :5316 */
        if (!(proj[2] == 0.0 && proj[5] == 0.0)) continue;                                   /* :5323 */
        orc_projective_transform(ref_srch, aux, proj, cols * G, rows * G, win_w, nss_window); /* :5336 */
        if (job->use_had) dist = orc_hads(org, job->org_stride, aux, cols, cols, rows, job->bit_depth);
        else dist = orc_sad(org, job->org_stride, aux, cols, cols, rows, 0, job->bit_depth);  /* :5344 */
        dist += orc_get_cost_xy(&job->cost, Hor, Ver);                                       /* :5345 */
        dist += orc_get_cost_bits(&job->cost, orc_get_bits_gt(                               /* :5346-5358 */
            cur_cx[0] / last_step, cur_cy[0] / last_step,
            (cur_cx[1] - cols * G + 1) / last_step, cur_cy[1] / last_step,
            (cur_cx[2] - cols * G + 1) / last_step, (cur_cy[2] - rows * G + 1) / last_step));
        n_cand++;
        if (dist < dist_best) {                                                              /* :5361-5384 */
          dist_best = dist;
          for (int k = 0; k < 4; k++) {
            best_cx[k] = cur_cx[k]; best_cy[k] = cur_cy[k];
            best_nx[k] = cur_cx[k]; best_ny[k] = cur_cy[k];
          }
          best_ss_x = Hor; best_ss_y = Ver;
          best_index = (b * 8 + pass) * 64 + cand;
        }
        cand++;
      }}}}}}}}
    }
  }
  free(window); free(aux);

  memset(out, 0, sizeof(*out));
  out->n_candidates = n_cand;
  out->best_index = -1;
  out->cost = job->threshold;
  int any = 0;
  for (int k = 0; k < 4; k++) any |= (best_cx[k] != 0) | (best_cy[k] != 0);
  if (any) {                                                                                 /* :5436-5459 */
    out->gt_flag = 1;
    out->gt[0].hor = (int16_t)(best_cx[0] / last_step);                 out->gt[0].ver = (int16_t)(best_cy[0] / last_step);
    out->gt[1].hor = (int16_t)((best_cx[1] - cols * G + 1) / last_step); out->gt[1].ver = (int16_t)(best_cy[1] / last_step);
    out->gt[2].hor = (int16_t)((best_cx[2] - cols * G + 1) / last_step); out->gt[2].ver = (int16_t)((best_cy[2] - rows * G + 1) / last_step);
    out->gt[3].hor = (int16_t)(best_cx[3] / last_step);                 out->gt[3].ver = (int16_t)((best_cy[3] - rows * G + 1) / last_step);
    out->cost = dist_best;
    out->mv_int.hor = (int16_t)(best_ss_x >> 2);
    out->mv_int.ver = (int16_t)(best_ss_y >> 2);
    out->best_index = best_index;
  }
}

/* ------------------------------------------------------------------------------------------------
 * Exhaustive sweep -- TLibEncoder/TEncSearch.cpp:4989-5091, the reference's IT_GT_SEARCH == 1
 * ("square unit search", N = 2 => 25^4 corner sets) as it compiles with IT_GT_GRID_SIZE 1
 * (SURVEY.md §8e): 1x grid (calcParamProjective / ProjectiveTransform get iCols,iRows :5026-5030,
 * TComPrediction.cpp:915-944 branch), "valid GT location" test re-enabled (:5020-5023), one pass from
 * the integer vector pcMvInt (half = quarter = 0), no AMVP start vectors.
 * cand_begin/cand_end restrict the evaluation to a slice of the AFFINE candidates in loop order (the
 * unit the multi-GPU sweep shards); key = cost << 32 | flat loop index, ~0 when nothing was scored.
 * ---------------------------------------------------------------------------------------------- */
static void projective_transform_grid1(const int16_t* ref, int16_t* aux_out, const double h[9],
                                       int W, int H, int stride, int nss_window)
{
  for (int y = 0; y < H; y++) {                       /* TComPrediction.cpp:916-917 */
    for (int x = 0; x < W; x++) {
      double Fx = (h[0] * x + h[3] * y + h[6]) / (h[2] * x + h[5] * y + h[8]);
      double Fy = (h[1] * x + h[4] * y + h[7]) / (h[2] * x + h[5] * y + h[8]);
      int Y = (int)Fy;                                /* :929-944 */
      int X = (int)Fx;
      double q = (Fy - (double)Y);
      double p = (Fx - (double)X);
      if (Y < -nss_window) Y = -nss_window;
      if (X < -nss_window) X = -nss_window;
      if (Y > nss_window + H - 1) Y = nss_window + H - 1;
      if (X > nss_window + W - 1) X = nss_window + W - 1;
      if (Y + 1 > nss_window + H - 1) Y = nss_window + H - 2;
      if (X + 1 > nss_window + W - 1) X = nss_window + W - 2;
      const int16_t* pa = ref + Y * stride;
      double aux = (1.0 - q) * ((1.0 - p) * (double)(pa[X]) + p * (double)(pa[X + 1]));
      pa = ref + (Y + 1) * stride;
      aux += q * ((1.0 - p) * (double)(pa[X]) + p * (double)(pa[X + 1]));
      if (aux > 255) aux = 255;
      if (aux < 0) aux = 0;
      aux_out[x] = (int16_t)(aux + 0.5);              /* :1023 */
    }
    aux_out += W;
  }
}

uint64_t orc_gt_sweep_key(const HopGtJob* job, const int16_t* org_buf, const int16_t* ref_buf,
                          int cand_begin, int cand_end, uint32_t* n_scored)
{
  const int cols = job->cols, rows = job->rows, N = 2;
  const int16_t* org = org_buf + job->org_off;
  const int16_t* ref_y = ref_buf + job->ref_off;
  const int win_w = cols * 2, win_h = rows * 2;
  int16_t* window = (int16_t*)malloc(sizeof(int16_t) * win_w * win_h);
  int16_t* aux = (int16_t*)malloc(sizeof(int16_t) * rows * cols);
  const int mvx = job->ss_cand.hor, mvy = job->ss_cand.ver;          /* pcMvInt */
  const int16_t Hor = (int16_t)(mvx << 2), Ver = (int16_t)(mvy << 2);  /* :4713-4724, half = qter = 0 */
  const int nss_window = ((rows < cols) ? rows : cols) >> 1;          /* :4756, grid 1 */
  const int i_offset = mvx - cols / 2 + (mvy - rows / 2) * job->ref_stride;   /* :4728 */
  orc_stage_window(ref_y + i_offset, job->ref_stride, window, win_w, win_h, job->bit_depth);
  const int16_t* ref_srch = window + (cols / 2) + (rows / 2) * win_w;  /* :4746 */
  const int cnx[4] = {0, cols - 1, cols - 1, 0}, cny[4] = {0, 0, rows - 1, rows - 1};   /* :4781-4784 */
  uint64_t best = ~(uint64_t)0;
  uint32_t scored = 0;
  int affine_idx = 0;
  double proj[9];
  int cx[4], cy[4];
  for (int y0 = -N; y0 <= N; y0++) { cy[0] = cny[0] + y0;
  for (int x0 = -N; x0 <= N; x0++) { cx[0] = cnx[0] + x0;
  for (int y1 = -N; y1 <= N; y1++) { cy[1] = cny[1] + y1;
  for (int x1 = -N; x1 <= N; x1++) { cx[1] = cnx[1] + x1;
  for (int y2 = -N; y2 <= N; y2++) { cy[2] = cny[2] + y2;
  for (int x2 = -N; x2 <= N; x2++) { cx[2] = cnx[2] + x2;
  for (int y3 = -N; y3 <= N; y3++) { cy[3] = cny[3] + y3;
  for (int x3 = -N; x3 <= N; x3++) { cx[3] = cnx[3] + x3;
    if (x0 == x1 && x0 == x2 && x0 == x3 && y0 == y1 && y0 == y2 && y0 == y3) continue;     /* :5017 */
    /* the multi-GPU shard unit: position among the parallelogram (affine) offset patterns */
    if (!(x0 - x1 + x2 - x3 == 0 && y0 - y1 + y2 - y3 == 0)) {
      /* not a parallelogram: the reference's double test below rejects it (h[2], h[5] != 0) */
      orc_calc_param_projective(cx, cy, proj, cols, rows);
      if (proj[2] == 0.0 && proj[5] == 0.0) { free(window); free(aux); return 1; }   /* cannot happen */
      continue;
    }
    const int my_idx = affine_idx++;
    if (my_idx < cand_begin || my_idx >= cand_end) continue;
    /* valid GT location, marginX = marginY = 0 (:5020-5023) */
    if (!(((x0 + mvx < 0 && y0 + mvy <= 0) || (x0 + mvx >= 0 && y0 + mvy < 0)) &&
          ((x1 + mvx + cols < 0 && y1 + mvy <= 0) || (x1 + mvx + cols >= 0 && y1 + mvy < 0)) &&
          ((x2 + mvx + cols < 0 && y2 + mvy + rows <= 0) || (x2 + mvx + cols >= 0 && y2 + mvy + rows < 0)) &&
          ((x3 + mvx < 0 && y3 + mvy + rows <= 0) || (x3 + mvx >= 0 && y3 + mvy + rows < 0)))) continue;
    orc_calc_param_projective(cx, cy, proj, cols, rows);                                   /* :5026 */
    if (!(proj[2] == 0.0 && proj[5] == 0.0)) continue;                                     /* :5028 */
    projective_transform_grid1(ref_srch, aux, proj, cols, rows, win_w, nss_window);        /* :5030 */
    uint32_t dist;
    if (job->use_had) dist = orc_hads(org, job->org_stride, aux, cols, cols, rows, job->bit_depth);
    else dist = orc_sad(org, job->org_stride, aux, cols, cols, rows, 0, job->bit_depth);
    dist += orc_get_cost_xy(&job->cost, Hor, Ver);                                         /* :5035 */
    dist += orc_get_cost_bits(&job->cost, orc_get_bits_gt(cx[0], cy[0], cx[1] - cols + 1, cy[1],
                                                          cx[2] - cols + 1, cy[2] - rows + 1));   /* :5036-5041 */
    scored++;
    const uint32_t flat = (uint32_t)((((((((y0 + N) * 5 + (x0 + N)) * 5 + (y1 + N)) * 5 + (x1 + N)) * 5 + (y2 + N)) * 5 +
                                       (x2 + N)) * 5 + (y3 + N)) * 5 + (x3 + N));
    const uint64_t key = ((uint64_t)dist << 32) | flat;
    if (key < best) best = key;          /* first strict minimum in loop order == min (cost, flat) */
  }}}}}}}}
  free(window); free(aux);
  if (n_scored) *n_scored = scored;
  return best;
}

/* key -> what xPatternSearchGT (mode 1) leaves in its outputs (:5070-5090) */
void orc_gt_sweep_finalize(const HopGtJob* job, uint64_t key, HopGtResult* out)
{
  const int cols = job->cols, rows = job->rows, N = 2;
  memset(out, 0, sizeof(*out));
  out->cost = job->threshold;
  out->best_index = -1;
  if (key == ~(uint64_t)0 || (uint32_t)(key >> 32) >= job->threshold) return;   /* uiDist < uiDistBest never true */
  uint32_t flat = (uint32_t)key;
  int o[8];
  for (int k = 7; k >= 0; k--) { o[k] = (int)(flat % 5) - N; flat /= 5; }     /* y0,x0,y1,x1,y2,x2,y3,x3 */
  const int bx[4] = {0 + o[1], cols - 1 + o[3], cols - 1 + o[5], 0 + o[7]};
  const int by[4] = {0 + o[0], 0 + o[2], rows - 1 + o[4], rows - 1 + o[6]};
  int any = 0;
  for (int k = 0; k < 4; k++) any |= (bx[k] != 0) | (by[k] != 0);
  if (!any) return;
  out->gt_flag = 1;
  out->gt[0].hor = (int16_t)bx[0];              out->gt[0].ver = (int16_t)by[0];
  out->gt[1].hor = (int16_t)(bx[1] - cols + 1); out->gt[1].ver = (int16_t)by[1];
  out->gt[2].hor = (int16_t)(bx[2] - cols + 1); out->gt[2].ver = (int16_t)(by[2] - rows + 1);
  out->gt[3].hor = (int16_t)bx[3];              out->gt[3].ver = (int16_t)(by[3] - rows + 1);
  out->cost = (uint32_t)(key >> 32);
  out->best_index = (int32_t)(uint32_t)key;
}

void orc_gt_sweep_batch(int n, const HopGtJob* jobs, const int16_t* org, const int16_t* ref, HopGtResult* out)
{
  for (int i = 0; i < n; i++) {
    uint32_t scored = 0;
    uint64_t key = orc_gt_sweep_key(&jobs[i], org, ref, 0, 1 << 30, &scored);
    orc_gt_sweep_finalize(&jobs[i], key, &out[i]);
    out[i].n_candidates = scored;
  }
}

void orc_gt_sweep_keys_batch(int n, const HopGtJob* jobs, const int16_t* org, const int16_t* ref,
                             int cand_begin, int cand_end, uint64_t* keys)
{
  for (int i = 0; i < n; i++) keys[i] = orc_gt_sweep_key(&jobs[i], org, ref, cand_begin, cand_end, NULL);
}

/* ------------------------------------------------------------------------------------------------
 * K5 -- fractional-pel refinement.  The reference materialises 16 planes (xExtDIFUpSamplingH/Q) and
 * indexes them with pointer fix-ups (xPatternRefinement :728-741); what every tested position reads is
 * the standard two-stage HEVC luma interpolation at that quarter-pel displacement: horizontal 8-tap
 * (isFirst, 14-bit intermediate, TComInterpolationFilter.cpp:173-254 / filterCopy :113-132) then
 * vertical 8-tap (isLast, rounding + clip).  Restated per position; pinned against the compiled
 * reference in tests/test_frac.py.
 * ---------------------------------------------------------------------------------------------- */
static const int16_t LUMA_FILTER[4][8] = {                 /* m_lumaFilter, TComInterpolationFilter.cpp:55-61 */
  {0, 0, 0, 64, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0}, {-1, 4, -11, 40, 40, -11, 4, -1}, {0, 1, -5, 17, 58, -10, 4, -1}};

void orc_interp_block(const int16_t* src, int stride, int qx, int qy, int cols, int rows, int bit_depth, int16_t* dst)
{
  const int ix = qx >> 2, fx = qx & 3, iy = qy >> 2, fy = qy & 3;   /* floor / fraction of the displacement */
  const int head = 14 - bit_depth;                                  /* IF_INTERNAL_PREC - bitDepth */
  const int16_t max_val = (int16_t)((1 << bit_depth) - 1);
  int16_t* tmp = (int16_t*)malloc(sizeof(int16_t) * (rows + 7) * cols);
  for (int r = -3; r < rows + 4; r++) {                             /* horizontal stage, isFirst && !isLast */
    const int16_t* s = src + (r + iy) * stride + ix;
    for (int c = 0; c < cols; c++) {
      int16_t v;
      if (fx == 0) {
        v = (int16_t)(s[c] << head);                                /* filterCopy isFirst :115-127 */
        v = (int16_t)(v - (int16_t)8192);
      } else {
        int sum = 0;
        for (int k = 0; k < 8; k++) sum += s[c + k - 3] * LUMA_FILTER[fx][k];
        const int shift = 6 - head;
        const int offset = -8192 << shift;
        v = (int16_t)((sum + offset) >> shift);
      }
      tmp[(r + 3) * cols + c] = v;
    }
  }
  for (int r = 0; r < rows; r++) {                                  /* vertical stage, !isFirst && isLast */
    for (int c = 0; c < cols; c++) {
      int16_t v;
      if (fy == 0) {
        int16_t off = (int16_t)8192;                                /* filterCopy isLast :135-149 */
        off = (int16_t)(off + (head ? (1 << (head - 1)) : 0));
        v = (int16_t)((tmp[(r + 3) * cols + c] + off) >> head);
      } else {
        int sum = 0;
        for (int k = 0; k < 8; k++) sum += tmp[(r + k) * cols + c] * LUMA_FILTER[fy][k];
        const int shift = 6 + head;
        const int offset = (1 << (shift - 1)) + (8192 << 6);
        v = (int16_t)((sum + offset) >> shift);
      }
      if (v < 0) v = 0;
      if (v > max_val) v = max_val;
      dst[r * cols + c] = v;
    }
  }
  free(tmp);
}

static const int8_t REFINE_H[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {1, -1}, {-1, 1}, {1, 1}};   /* s_acMvRefineH :46-57 */
static const int8_t REFINE_Q[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {1, 1}};   /* s_acMvRefineQ :59-70 */

void orc_frac_search(const HopFracJob* job, const int16_t* org_buf, const int16_t* ref_buf, HopFracResult* out)
{
  const int cols = job->cols, rows = job->rows;
  const int16_t* org = org_buf + job->org_off;
  const int16_t* src = ref_buf + job->ref_off + job->mv_int.hor + job->mv_int.ver * job->ref_stride;   /* :6579 */
  int16_t* blk = (int16_t*)malloc(sizeof(int16_t) * rows * cols);
  HopCostState cs = job->cost;
  uint32_t best = HOP_MAX_UINT;
  int best_i = 0;
  /* half-pel stage: baseRefMv (0,0), iFrac 2, rcMvFrac = int << 1, cost scale 1 (:4615, 6594-6598) */
  cs.cost_scale = 1;
  for (int i = 0; i < 9; i++) {
    const int hx = REFINE_H[i][0], hy = REFINE_H[i][1];
    orc_interp_block(src, job->ref_stride, hx * 2, hy * 2, cols, rows, job->bit_depth, blk);
    uint32_t d = job->use_had ? orc_hads(org, job->org_stride, blk, cols, cols, rows, job->bit_depth)
                              : orc_sad(org, job->org_stride, blk, cols, cols, rows, 0, job->bit_depth);
    d += orc_get_cost_xy(&cs, (job->mv_int.hor << 1) + hx, (job->mv_int.ver << 1) + hy);   /* :747 */
    if (d < best) { best = d; best_i = i; }
  }
  const int half_x = REFINE_H[best_i][0], half_y = REFINE_H[best_i][1];
  out->half.hor = (int16_t)half_x; out->half.ver = (int16_t)half_y;
  out->cost_half = best;
  /* quarter-pel stage: baseRefMv = half << 1, iFrac 1, rcMvFrac = ((int << 1) + half) << 1, cost scale 0 (:6600-6608) */
  cs.cost_scale = 0;
  best = HOP_MAX_UINT; best_i = 0;
  const int base_x = half_x << 1, base_y = half_y << 1;
  const int mvq_x = (((job->mv_int.hor << 1) + half_x) << 1), mvq_y = (((job->mv_int.ver << 1) + half_y) << 1);
  for (int i = 0; i < 9; i++) {
    const int qx = REFINE_Q[i][0], qy = REFINE_Q[i][1];
    orc_interp_block(src, job->ref_stride, base_x + qx, base_y + qy, cols, rows, job->bit_depth, blk);
    uint32_t d = job->use_had ? orc_hads(org, job->org_stride, blk, cols, cols, rows, job->bit_depth)
                              : orc_sad(org, job->org_stride, blk, cols, cols, rows, 0, job->bit_depth);
    d += orc_get_cost_xy(&cs, mvq_x + qx, mvq_y + qy);
    if (d < best) { best = d; best_i = i; }
  }
  out->qter.hor = REFINE_Q[best_i][0]; out->qter.ver = REFINE_Q[best_i][1];
  out->cost = best;
  free(blk);
}

void orc_frac_search_batch(int n, const HopFracJob* jobs, const int16_t* org, const int16_t* ref, HopFracResult* out)
{
  for (int i = 0; i < n; i++) orc_frac_search(&jobs[i], org, ref, &out[i]);
}

void orc_motion_search_batch(int n, const HopMotionJob* jobs, const int16_t* org, const int16_t* ref, HopMotionResult* out)
{
  for (int i = 0; i < n; i++) {
    const HopMotionJob* mj = &jobs[i];
    HopMotionResult* r = &out[i];
    memset(r, 0, sizeof(*r));
    orc_pattern_search(&mj->search, org, ref, &r->search);                      /* :4582 */
    r->gt.cost = 0; r->gt.best_index = -1;
    if (!r->search.found || (r->search.mv.hor == 0 && r->search.mv.ver == 0)) continue;   /* :4603-4611 */
    HopFracJob fj;
    memset(&fj, 0, sizeof(fj));
    fj.org_off = mj->search.org_off; fj.ref_off = mj->search.ref_off;
    fj.org_stride = mj->search.org_stride; fj.ref_stride = mj->search.ref_stride;
    fj.cols = mj->search.cols; fj.rows = mj->search.rows;
    fj.mv_int = r->search.mv; fj.use_had = mj->use_had; fj.bit_depth = mj->search.bit_depth;
    fj.cost = mj->search.cost;
    orc_frac_search(&fj, org, ref, &r->frac);                                    /* :4617 */
    r->refined = 1;
    r->gt.cost = r->frac.cost;
    if (!mj->use_gt) continue;                                                   /* :4627 */
    HopGtJob gj;
    memset(&gj, 0, sizeof(gj));
    gj.org_off = fj.org_off; gj.ref_off = fj.ref_off; gj.org_stride = fj.org_stride; gj.ref_stride = fj.ref_stride;
    gj.cols = fj.cols; gj.rows = fj.rows;
    gj.ss_cand = r->search.mv;                /* pcCU->getSSBestCand()[0] == the integer winner */
    gj.num_pred = mj->num_pred;
    for (int k = 0; k < HOP_MAX_PRED; k++) gj.amvp[k] = mj->amvp[k];
    gj.threshold = r->frac.cost; gj.use_had = mj->use_had; gj.bit_depth = fj.bit_depth;
    gj.cost = mj->search.cost; gj.cost.cost_scale = 0;                           /* :4619 */
    orc_pattern_search_gt(&gj, org, ref, &r->gt);
  }
}

void orc_pattern_search_batch(int n, const HopSearchJob* jobs, const int16_t* org, const int16_t* ref,
                              HopSearchResult* out)
{
  for (int i = 0; i < n; i++) orc_pattern_search(&jobs[i], org, ref, &out[i]);
}

void orc_pattern_search_gt_batch(int n, const HopGtJob* jobs, const int16_t* org, const int16_t* ref,
                                 HopGtResult* out)
{
  for (int i = 0; i < n; i++) orc_pattern_search_gt(&jobs[i], org, ref, &out[i]);
}

/* ------------------------------------------------------------------------------------------------
 * K4 -- TLibCommon/TComPicYuv.cpp:247-274 (xExtendPicCompBorder, luma)
 * ---------------------------------------------------------------------------------------------- */
void orc_extend_border(int16_t* origin, int stride, int pic_w, int pic_h, int margin)
{
  int16_t* pi = origin;
  for (int y = 0; y < pic_h; y++) {
    for (int x = 0; x < margin; x++) {
      pi[-margin + x] = pi[0];
      pi[pic_w + x] = pi[pic_w - 1];
    }
    pi += stride;
  }
  pi -= (stride + margin);
  for (int y = 0; y < margin; y++) memcpy(pi + (y + 1) * stride, pi, sizeof(int16_t) * (pic_w + (margin << 1)));
  pi -= ((pic_h - 1) * stride);
  for (int y = 0; y < margin; y++) memcpy(pi - (y + 1) * stride, pi, sizeof(int16_t) * (pic_w + (margin << 1)));
}

/* ------------------------------------------------------------------------------------------------
 * K6 -- motion-compensated prediction (uni-prediction) + what the encoder derives from it.
 * xPredInterLumaBlk / xPredInterChromaBlk (TLibCommon/TComPrediction.cpp:639-720, 1235-1347) with the DCT-IF
 * (TComInterpolationFilter.cpp:92-420), xPredGTLuma / xPredGTChroma (:723-805, 1351-1420),
 * calcParamProjectiveC (:834-859), xGetInterPredictionError (TLibEncoder/TEncSearch.cpp:2951-2977),
 * xGetTemplateCost (:4390-4477), isValidPattern (TLibCommon/TComRdCost.cpp:430-442), calcRdCost DF_SAD (:59-111).
 * ---------------------------------------------------------------------------------------------- */
static const int16_t CHROMA_FILTER[8][4] = {               /* m_chromaFilter, TComInterpolationFilter.cpp:63-73 */
  {0, 64, 0, 0}, {-2, 58, 10, -2}, {-4, 54, 16, -2}, {-6, 46, 28, -4}, {-4, 36, 36, -4}, {-4, 28, 46, -6}, {-2, 16, 54, -4}, {-2, 10, 58, -2}};

/* One sample of the block the three-way branch of xPredInter{Luma,Chroma}Blk writes for bi == false:
 *   yFrac == 0 : filterHor(isLast)            (frac 0: filterCopy first == last, a plain copy: NOT_VALID stays -1)
 *   xFrac == 0 : filterVer(isFirst, isLast)
 *   else       : filterHor(!isLast) into the 14-bit intermediate, filterVer(!isFirst, isLast)
 * `src` points at the integer position of the block's sample (0,0); taps = 8 (luma) or 4 (chroma). */
static int16_t mc_sample(const int16_t* src, int stride, int x, int y, int fx, int fy, int taps, int bit_depth)
{
  const int half = taps / 2 - 1;
  const int head = 14 - bit_depth;                            /* IF_INTERNAL_PREC - bitDepth */
  const int16_t max_val = (int16_t)((1 << bit_depth) - 1);
  const int16_t* cx = taps == 8 ? LUMA_FILTER[fx] : CHROMA_FILTER[fx];
  const int16_t* cy = taps == 8 ? LUMA_FILTER[fy] : CHROMA_FILTER[fy];
  const int16_t* s = src + y * stride + x;
  if (fy == 0) {
    if (fx == 0) return s[0];
    int sum = 0;
    for (int k = 0; k < taps; k++) sum += s[k - half] * cx[k];
    int16_t v = (int16_t)((sum + 32) >> 6);                    /* isFirst && isLast: shift 6, offset 1 << 5 */
    if (v < 0) v = 0;
    if (v > max_val) v = max_val;
    return v;
  }
  if (fx == 0) {
    int sum = 0;
    for (int k = 0; k < taps; k++) sum += s[(k - half) * stride] * cy[k];
    int16_t v = (int16_t)((sum + 32) >> 6);
    if (v < 0) v = 0;
    if (v > max_val) v = max_val;
    return v;
  }
  int sum2 = 0;
  for (int r = 0; r < taps; r++) {
    const int16_t* row = s + (r - half) * stride;
    int sum = 0;
    for (int k = 0; k < taps; k++) sum += row[k - half] * cx[k];
    const int shift = 6 - head;                                /* isFirst && !isLast */
    const int offset = -8192 << shift;
    const int16_t t = (int16_t)((sum + offset) >> shift);
    sum2 += t * cy[r];
  }
  const int shift = 6 + head;                                  /* !isFirst && isLast */
  const int offset = (1 << (shift - 1)) + (8192 << 6);
  int16_t v = (int16_t)((sum2 + offset) >> shift);
  if (v < 0) v = 0;
  if (v > max_val) v = max_val;
  return v;
}

/* calcParamProjectiveC, TComPrediction.cpp:834-859 (corners in binary64) */
static void calc_param_projective_c(const double x[4], const double y[4], double h[9], int width, int height)
{
  double H, W, dx[4], dy[4];
  W = (double)width - 1.0;
  H = (double)height - 1.0;
  dx[1] = x[1] - x[2];
  dx[2] = x[3] - x[2];
  dx[3] = x[0] - x[1] + x[2] - x[3];
  dy[1] = y[1] - y[2];
  dy[2] = y[3] - y[2];
  dy[3] = y[0] - y[1] + y[2] - y[3];
  h[2] = ((dx[3] * dy[2] - dx[2] * dy[3]) / (dx[1] * dy[2] - dx[2] * dy[1])) / W;
  h[5] = ((dx[1] * dy[3] - dx[3] * dy[1]) / (dx[1] * dy[2] - dx[2] * dy[1])) / H;
  h[0] = (x[1] - x[0]) / W + h[2] * x[1];
  h[3] = (x[3] - x[0]) / H + h[5] * x[3];
  h[6] = x[0];
  h[1] = (y[1] - y[0]) / W + h[2] * y[1];
  h[4] = (y[3] - y[0]) / H + h[5] * y[3];
  h[7] = y[0];
  h[8] = 1.0;
}

void orc_predict(const HopPredJob* job, const int16_t* org_buf, const int16_t* ref_buf, int16_t* dst_buf, HopPredResult* out)
{
  const int chroma = job->comp != 0;
  const int bw = chroma ? job->cols >> 1 : job->cols, bh = chroma ? job->rows >> 1 : job->rows;   /* block of this component */
  const int taps = chroma ? 4 : 8;
  const int fmask = chroma ? 7 : 3, fshift = chroma ? 3 : 2;
  const int16_t* ref = ref_buf + job->ref_off;
  const int stride = job->ref_stride;
  int16_t* pred = (int16_t*)malloc(sizeof(int16_t) * bw * bh);
  out->valid = 1; out->dist = 0; out->cost = 0;

  if (job->template_cost && job->is_ss) {                                 /* TEncSearch.cpp:4420-4436, TComRdCost.cpp:430-442 */
    const int16_t* cur = ref + (job->mv_probe.hor >> 2) + (job->mv_probe.ver >> 2) * stride;
    const int16_t* lb = cur + (job->rows + 4) * stride;
    const int16_t* rb = lb + (job->cols + 4);
    if (*lb == HOP_NOT_VALID || *rb == HOP_NOT_VALID) {
      out->valid = 0; out->cost = 0x7fffffffu;                            /* MAX_INT */
      free(pred);
      return;
    }
  }
  const int gt_any = job->gt[0].hor | job->gt[1].hor | job->gt[2].hor | job->gt[3].hor |
                     job->gt[0].ver | job->gt[1].ver | job->gt[2].ver | job->gt[3].ver;
  const int fx = job->mv.hor & fmask, fy = job->mv.ver & fmask;
  if (!job->gt_flag || !gt_any || job->template_cost) {                   /* plain branch (:651, 1246) */
    const int16_t* src = ref + (job->mv.hor >> fshift) + (job->mv.ver >> fshift) * stride;
    for (int y = 0; y < bh; y++)
      for (int x = 0; x < bw; x++) pred[y * bw + x] = mc_sample(src, stride, x, y, fx, fy, taps, job->bit_depth);
  } else {
    /* the 2W x 2H region around the block (:683-712 luma, :1292-1330 chroma; width / 4 is on the LUMA width) */
    const int ox = chroma ? job->cols / 4 : job->cols / 2, oy = chroma ? job->rows / 4 : job->rows / 2;
    const int16_t* src = ref + (job->mv.hor >> fshift) - ox + ((job->mv.ver >> fshift) - oy) * stride;
    const int W2 = 2 * bw, H2 = 2 * bh;
    int16_t* reg = (int16_t*)malloc(sizeof(int16_t) * W2 * H2);
    for (int y = 0; y < H2; y++)
      for (int x = 0; x < W2; x++) reg[y * W2 + x] = mc_sample(src, stride, x, y, fx, fy, taps, job->bit_depth);
    /* xPredGTLuma / xPredGTChroma on the component's block size (:723-805, 1351-1420) */
    int nss = ((bh < bw) ? bh : bw) >> 1;
    nss *= 2;                                                             /* IT_GT_GRID_SIZE */
    int last_step = nss >> 6;                                             /* IT_MAX_NSS_Iteration */
    if (last_step == 0) last_step = 1;
    double h[9];
    if (!chroma) {
      int32_t cx[4], cy[4];
      cx[0] = job->gt[0].hor * last_step;              cy[0] = job->gt[0].ver * last_step;
      cx[1] = job->gt[1].hor * last_step + bw * 2 - 1; cy[1] = job->gt[1].ver * last_step;
      cx[2] = job->gt[2].hor * last_step + bw * 2 - 1; cy[2] = job->gt[2].ver * last_step + bh * 2 - 1;
      cx[3] = job->gt[3].hor * last_step;              cy[3] = job->gt[3].ver * last_step + bh * 2 - 1;
      orc_calc_param_projective(cx, cy, h, bw * 2, bh * 2);
    } else {
      const double ls = (double)last_step;                                /* Double lastIterationStep (:1366) */
      double cx[4], cy[4];
      cx[0] = ((double)job->gt[0].hor / 2) * ls;                cy[0] = ((double)job->gt[0].ver / 2) * ls;
      cx[1] = (((double)job->gt[1].hor / 2) * ls) + bw * 2 - 1; cy[1] = ((double)job->gt[1].ver / 2) * ls;
      cx[2] = (((double)job->gt[2].hor / 2) * ls) + bw * 2 - 1; cy[2] = (((double)job->gt[2].ver / 2) * ls) + bh * 2 - 1;
      cx[3] = ((double)job->gt[3].hor / 2) * ls;                cy[3] = (((double)job->gt[3].ver / 2) * ls) + bh * 2 - 1;
      calc_param_projective_c(cx, cy, h, bw * 2, bh * 2);
    }
    orc_projective_transform(reg + bw / 2 + (bh / 2) * W2, pred, h, bw * 2, bh * 2, W2, (((bh < bw) ? bh : bw) >> 1) * 2);
    free(reg);
  }
  if (job->dst_off >= 0 && dst_buf)
    for (int y = 0; y < bh; y++)
      for (int x = 0; x < bw; x++) dst_buf[job->dst_off + y * job->dst_stride + x] = pred[y * bw + x];
  if (job->template_cost) {
    /* getDistPart(.., DF_SAD) then calcRdCost(bits, dist, false, DF_SAD) (TEncSearch.cpp:4473-4474) */
    const uint32_t sad = orc_sad(org_buf + job->org_off, job->org_stride, pred, bw, bw, bh, 0, job->bit_depth);
    const double lambda = (double)job->lambda_sad;
    double c = ((double)sad + (double)((int)(job->mvp_bits * lambda + .5) >> 16));
    c = (double)(uint32_t)floor(c);
    out->dist = sad;
    out->cost = (uint32_t)c;
  } else if (job->dist_func) {
    out->dist = job->dist_func == HOP_DF_HADS ? orc_hads(org_buf + job->org_off, job->org_stride, pred, bw, bw, bh, job->bit_depth)
                                              : orc_sad(org_buf + job->org_off, job->org_stride, pred, bw, bw, bh, 0, job->bit_depth);
    out->cost = out->dist;
  }
  free(pred);
}

void orc_predict_batch(int n, const HopPredJob* jobs, const int16_t* org, const int16_t* ref, int16_t* dst, HopPredResult* out)
{
  for (int i = 0; i < n; i++) orc_predict(&jobs[i], org, ref, dst, &out[i]);
}

/* ------------------------------------------------------------------------------------------------
 * K7 -- intra mode pre-screen: predIntraLumaAng for the 35 modes + calcHAD
 * (TLibCommon/TComPrediction.cpp:129-170, 192-348, 1468-1546; TComPattern.cpp:49-56, 583-607;
 *  TComRdCost.cpp:391-425; the caller's loop TLibEncoder/TEncSearch.cpp:2451-2464).
 * above[k] = pSrc[k - srcStride - 1], left[k] = pSrc[(k - 1) * srcStride - 1]  (k = 0: the corner sample).
 * ---------------------------------------------------------------------------------------------- */
static const unsigned char INTRA_FILTER[5] = {10, 7, 1, 0, 10};      /* m_aucIntraFilter, TComPattern.cpp:49-56 */

static int ilog2(int n) { int l = 0; while ((1 << l) < n) l++; return l; }

void orc_intra_predict(const int32_t* above, const int32_t* left, int n, int mode, int bit_depth, int above_avail, int left_avail,
                       int16_t* dst /* n x n, stride n */)
{
  const int filter_edges = n <= 16;                                   /* predIntraLumaAng :333-346 */
  if (n < 4 || n > 64) return;
  if (mode == 0) {                                                    /* xPredIntraPlanar :1468-1510 */
    int left_col[65], top_row[65], bottom_row[64], right_col[64];
    const int shift1 = ilog2(n), shift2 = shift1 + 1;
    for (int k = 0; k < n + 1; k++) { top_row[k] = above[k + 1]; left_col[k] = left[k + 1]; }
    const int bottom_left = left_col[n], top_right = top_row[n];
    for (int k = 0; k < n; k++) {
      bottom_row[k] = bottom_left - top_row[k];
      right_col[k] = top_right - left_col[k];
      top_row[k] <<= shift1;
      left_col[k] <<= shift1;
    }
    for (int k = 0; k < n; k++) {
      int hor = left_col[k] + n;
      for (int l = 0; l < n; l++) {
        hor += right_col[k];
        top_row[l] += bottom_row[l];
        dst[k * n + l] = (int16_t)((hor + top_row[l]) >> shift2);
      }
    }
    return;
  }
  if (mode == 1) {                                                    /* DC: predIntraGetPredValDC :129-170 */
    int sum = 0;
    int16_t dc;
    if (above_avail) for (int i = 0; i < n; i++) sum += above[i + 1];
    if (left_avail) for (int i = 0; i < n; i++) sum += left[i + 1];
    if (above_avail && left_avail) dc = (int16_t)((sum + n) / (n + n));
    else if (above_avail) dc = (int16_t)((sum + n / 2) / n);
    else if (left_avail) dc = (int16_t)((sum + n / 2) / n);
    else dc = (int16_t)left[1];                                        /* pSrc[-1] */
    for (int k = 0; k < n * n; k++) dst[k] = dc;
    if (filter_edges && above_avail && left_avail) {                  /* xDCPredFiltering :1524-1546 */
      dst[0] = (int16_t)((above[1] + left[1] + 2 * dst[0] + 2) >> 2);
      for (int x = 1; x < n; x++) dst[x] = (int16_t)((above[x + 1] + 3 * dst[x] + 2) >> 2);
      for (int y = 1; y < n; y++) dst[y * n] = (int16_t)((left[y + 1] + 3 * dst[y * n] + 2) >> 2);
    }
    return;
  }
  /* xPredIntraAng :192-314 */
  static const int ang_table[9] = {0, 2, 5, 9, 13, 17, 21, 26, 32};
  static const int inv_ang_table[9] = {0, 4096, 1638, 910, 630, 482, 390, 315, 256};
  const int mode_hor = mode < 18, mode_ver = !mode_hor;
  int angle = mode_ver ? mode - 26 : -(mode - 10);
  int abs_ang = abs(angle);
  const int sign = angle < 0 ? -1 : 1;
  const int inv_angle = inv_ang_table[abs_ang];
  abs_ang = ang_table[abs_ang];
  angle = sign * abs_ang;
  int16_t ref_above[2 * 64 + 1], ref_left[2 * 64 + 1];
  int16_t *ref_main, *ref_side;
  if (angle < 0) {
    for (int k = 0; k < n + 1; k++) ref_above[k + n - 1] = (int16_t)above[k];
    for (int k = 0; k < n + 1; k++) ref_left[k + n - 1] = (int16_t)left[k];
    ref_main = (mode_ver ? ref_above : ref_left) + (n - 1);
    ref_side = (mode_ver ? ref_left : ref_above) + (n - 1);
    int inv_sum = 128;
    for (int k = -1; k > (n * angle) >> 5; k--) {
      inv_sum += inv_angle;
      ref_main[k] = ref_side[inv_sum >> 8];
    }
  } else {
    for (int k = 0; k < 2 * n + 1; k++) ref_above[k] = (int16_t)above[k];
    for (int k = 0; k < 2 * n + 1; k++) ref_left[k] = (int16_t)left[k];
    ref_main = mode_ver ? ref_above : ref_left;
    ref_side = mode_ver ? ref_left : ref_above;
  }
  if (angle == 0) {
    for (int k = 0; k < n; k++)
      for (int l = 0; l < n; l++) dst[k * n + l] = ref_main[l + 1];
    if (filter_edges) {
      const int max_val = (1 << bit_depth) - 1;
      for (int k = 0; k < n; k++) {
        int v = dst[k * n] + ((ref_side[k + 1] - ref_side[0]) >> 1);
        dst[k * n] = (int16_t)(v < 0 ? 0 : (v > max_val ? max_val : v));
      }
    }
  } else {
    int delta_pos = 0;
    for (int k = 0; k < n; k++) {
      delta_pos += angle;
      const int delta_int = delta_pos >> 5, delta_fract = delta_pos & 31;
      if (delta_fract) {
        for (int l = 0; l < n; l++) {
          const int idx = l + delta_int + 1;
          dst[k * n + l] = (int16_t)(((32 - delta_fract) * ref_main[idx] + delta_fract * ref_main[idx + 1] + 16) >> 5);
        }
      } else {
        for (int l = 0; l < n; l++) dst[k * n + l] = ref_main[l + delta_int + 1];
      }
    }
  }
  if (mode_hor)                                                       /* flip */
    for (int k = 0; k < n - 1; k++)
      for (int l = k + 1; l < n; l++) { int16_t t = dst[k * n + l]; dst[k * n + l] = dst[l * n + k]; dst[l * n + k] = t; }
}

void orc_intra_prescreen(const HopIntraJob* job, const int16_t* org_buf, const int32_t* refs, uint32_t* out /* 35 */)
{
  const int n = job->size, sw = 2 * n + 1, lg = ilog2(n);
  const int32_t* r = refs + job->refs_off;
  int16_t* pred = (int16_t*)malloc(sizeof(int16_t) * n * n);
  for (int mode = 0; mode < HOP_INTRA_MODES; mode++) {
    /* TComPattern::getPredictorPtr :583-607 */
    int diff = abs(mode - 10) < abs(mode - 26) ? abs(mode - 10) : abs(mode - 26);
    int filt = diff > INTRA_FILTER[lg - 2] ? 1 : 0;
    if (mode == 1) filt = 0;
    const int32_t* above = r + (filt ? 2 : 0) * sw;
    const int32_t* left = above + sw;
    orc_intra_predict(above, left, n, mode, job->bit_depth, job->above_avail, job->left_avail, pred);
    /* calcHAD :391-425: 8x8 tiles when both sizes are multiples of 8, else 4x4 */
    const int t = (n % 8 == 0) ? 8 : 4;
    uint32_t sum = 0;
    const int16_t* org = org_buf + job->org_off;
    for (int y = 0; y < n; y += t)
      for (int x = 0; x < n; x += t) sum += had_tile(org + y * job->org_stride + x, pred + y * n + x, job->org_stride, n, t);
    out[mode] = sum >> DIST_SHIFT(job->bit_depth);
  }
  free(pred);
}

void orc_intra_prescreen_batch(int n, const HopIntraJob* jobs, const int16_t* org, const int32_t* refs, uint32_t* out)
{
  for (int i = 0; i < n; i++) orc_intra_prescreen(&jobs[i], org, refs, out + (size_t)i * HOP_INTRA_MODES);
}
