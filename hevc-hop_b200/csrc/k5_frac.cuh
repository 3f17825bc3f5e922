// k5_frac.cuh -- K5: half-/quarter-pel refinement of the SS vector (sm_100a), device code shared by the
// stand-alone kernel (k5_frac_search) and the fused motion kernel.
//
// Replaces TEncSearch::xPatternSearchFracDIF (TLibEncoder/TEncSearch.cpp:6564-6610): xExtDIFUpSamplingH/Q
// (:7818-8011, 16 interpolated planes on the host) + xPatternRefinement (:709-761, 2 x 9 Hadamard / SAD
// costs).  What every tested position reads is the two-stage HEVC luma interpolation at that quarter-pel
// displacement (TLibCommon/TComInterpolationFilter.cpp:92-254): horizontal 8-tap into a 14-bit
// intermediate (isFirst), vertical 8-tap with rounding and clip (isLast); integer arithmetic only.
//
// One CTA per PU.  Per stage the 9 positions are 3 horizontal x 3 vertical displacements: the three
// horizontal intermediates are built once in shared memory, then a task = (position, Hadamard tile)
// filters its tile vertically into registers, subtracts the original block and runs the butterflies.
#pragma once
#include "hop_common.cuh"
#include "hop_internal.h"

namespace hop {

__constant__ int8_t c_luma_filter[4][8] = {   // m_lumaFilter, TComInterpolationFilter.cpp:55-61
  {0, 0, 0, 64, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0}, {-1, 4, -11, 40, 40, -11, 4, -1}, {0, 1, -5, 17, 58, -10, 4, -1}};
__constant__ int8_t c_refine_h[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {1, -1}, {-1, 1}, {1, 1}};   // s_acMvRefineH :46-57
__constant__ int8_t c_refine_q[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {1, 1}};   // s_acMvRefineQ :59-70

struct FracShared {
  uint32_t dist[9];
  int32_t  best_i;
  uint32_t best_cost;
};

__host__ __device__ inline size_t frac_smem_bytes(int cols, int rows)
{
  // [source region (rows+8) x (cols+8) int16][3 horizontal intermediates (rows+8) x cols int16]
  return (sizeof(int16_t) * ((size_t)(rows + 8) * (cols + 8) + 3 * (size_t)(rows + 8) * cols) + 15) & ~(size_t)15;
}

template <int N>
struct FracTile {
  int d[N * N];
  __device__ __forceinline__ uint32_t satd()
  {
#pragma unroll
    for (int y = 0; y < N; y++)
#pragma unroll
      for (int len = 1; len < N; len <<= 1)
#pragma unroll
        for (int i = 0; i < N; i += len << 1)
#pragma unroll
          for (int j = i; j < i + len; j++) { int a = d[y * N + j], b = d[y * N + j + len]; d[y * N + j] = a + b; d[y * N + j + len] = a - b; }
#pragma unroll
    for (int x = 0; x < N; x++)
#pragma unroll
      for (int len = 1; len < N; len <<= 1)
#pragma unroll
        for (int i = 0; i < N; i += len << 1)
#pragma unroll
          for (int j = i; j < i + len; j++) { int a = d[j * N + x], b = d[(j + len) * N + x]; d[j * N + x] = a + b; d[(j + len) * N + x] = a - b; }
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < N * N; k++) s = __sad(d[k], 0, s);
    return N == 8 ? (s + 2) >> 2 : (s + 1) >> 1;     // xCalcHADs8x8 / xCalcHADs4x4 rounding
  }
  __device__ __forceinline__ uint32_t sad()
  {
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < N * N; k++) s = __sad(d[k], 0, s);
    return s;
  }
};

// one refinement stage: positions base + 2^scale_shift * refine[i] (quarter-pel units), i = 0..8
//   s_org : original block, int32, stride cols          s_src : staged source region
//   s_tmp : 3 planes, plane index = refine x-component + 1
template <int N>
__device__ __forceinline__ void frac_stage(FracShared& fs, const int* __restrict__ s_org, const int16_t* __restrict__ s_src,
                                           int16_t* __restrict__ s_tmp, int cols, int rows, int bit_depth, int use_had,
                                           int base_qx, int base_qy, int step, const int8_t (*refine)[2])
{
  const int head = 14 - bit_depth;                     // IF_INTERNAL_PREC - bitDepth
  const int src_w = cols + 8, plane = (rows + 8) * cols;
  if (threadIdx.x < 9) fs.dist[threadIdx.x] = 0;
  // horizontal intermediates for the three x displacements (isFirst && !isLast)
  for (int i = threadIdx.x; i < 3 * plane; i += blockDim.x) {
    const int xi = i / plane, rem = i - xi * plane, rr = rem / cols, c = rem - rr * cols;
    const int qx = base_qx + step * (xi - 1);
    const int ix = qx >> 2, fx = qx & 3;
    const int16_t* s = s_src + rr * src_w + (c + ix + 4);
    int16_t v;
    if (fx == 0) {
      v = (int16_t)(s[0] << head);                     // filterCopy isFirst, :115-127
      v = (int16_t)(v - (int16_t)8192);
    } else {
      int sum = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) sum += (int)s[k - 3] * c_luma_filter[fx][k];
      const int shift = 6 - head;
      v = (int16_t)((sum + (-8192 << shift)) >> shift);
    }
    s_tmp[i] = v;
  }
  __syncthreads();
  const int tiles_x = cols / N, ntiles = tiles_x * (rows / N);
  const int max_val = (1 << bit_depth) - 1;
  for (int t = threadIdx.x; t < 9 * ntiles; t += blockDim.x) {
    const int i = t % 9, tile = t / 9;
    const int tx = (tile % tiles_x) * N, ty = (tile / tiles_x) * N;
    const int qy = base_qy + step * refine[i][1];
    const int iy = qy >> 2, fy = qy & 3;
    const int16_t* tp = s_tmp + (refine[i][0] + 1) * plane;
    FracTile<N> ft;
#pragma unroll
    for (int r = 0; r < N; r++) {
#pragma unroll
      for (int c = 0; c < N; c++) {
        // tmp row index of source row (ty + r + iy + k - 3) is that + 4
        const int16_t* col = tp + (ty + r + iy + 1) * cols + tx + c;
        int v;
        if (fy == 0) {
          int16_t off = (int16_t)8192;                 // filterCopy isLast, :135-149
          off = (int16_t)(off + (head ? (1 << (head - 1)) : 0));
          v = (int16_t)((col[3 * cols] + off) >> head);
        } else {
          int sum = 0;
#pragma unroll
          for (int k = 0; k < 8; k++) sum += (int)col[k * cols] * c_luma_filter[fy][k];
          const int shift = 6 + head;
          v = (int16_t)((sum + (1 << (shift - 1)) + (8192 << 6)) >> shift);
        }
        v = min(max(v, 0), max_val);
        ft.d[r * N + c] = s_org[(ty + r) * cols + tx + c] - v;
      }
    }
    atomicAdd(&fs.dist[i], use_had ? ft.satd() : ft.sad());
  }
  __syncthreads();
}

// Whole xPatternSearchFracDIF for one PU by one CTA.  s_org must hold the original block (int32, stride
// cols); `scratch` provides frac_smem_bytes(cols, rows) bytes.  Result valid in every thread on return.
__device__ __forceinline__ HopFracResult frac_search_cta(FracShared& fs, const int* __restrict__ s_org,
                                                         unsigned char* scratch, const int16_t* __restrict__ ref_pos,
                                                         int ref_stride, int cols, int rows, int bit_depth, int use_had,
                                                         HopCostState cs, HopMv mv_int)
{
  int16_t* s_src = reinterpret_cast<int16_t*>(scratch);
  int16_t* s_tmp = s_src + (rows + 8) * (cols + 8);
  const int src_w = cols + 8;
  // source region: x in [-4, cols+4), y in [-4, rows+4) around the integer position (:6579)
  for (int i = threadIdx.x; i < (rows + 8) * src_w; i += blockDim.x) {
    const int y = i / src_w, x = i - y * src_w;
    s_src[i] = ref_pos[(long long)(y - 4) * ref_stride + (x - 4)];
  }
  __syncthreads();
  const int tile_n = ((rows % 8 == 0) && (cols % 8 == 0)) ? 8 : 4;
  const int dist_shift = bit_depth - 8;
  HopFracResult res;
  // half-pel stage: baseRefMv (0,0), iFrac 2, mv cost in half-pel units, cost scale 1 (:4615, 6594-6598)
  if (tile_n == 8) frac_stage<8>(fs, s_org, s_src, s_tmp, cols, rows, bit_depth, use_had, 0, 0, 2, c_refine_h);
  else             frac_stage<4>(fs, s_org, s_src, s_tmp, cols, rows, bit_depth, use_had, 0, 0, 2, c_refine_h);
  if (threadIdx.x == 0) {
    cs.cost_scale = 1;
    uint32_t best = HOP_MAX_UINT; int bi = 0;
    for (int i = 0; i < 9; i++) {
      const uint32_t d = (fs.dist[i] >> dist_shift) + mv_cost(cs, (mv_int.hor << 1) + c_refine_h[i][0], (mv_int.ver << 1) + c_refine_h[i][1]);
      if (d < best) { best = d; bi = i; }
    }
    fs.best_i = bi; fs.best_cost = best;
  }
  __syncthreads();
  const int half_x = c_refine_h[fs.best_i][0], half_y = c_refine_h[fs.best_i][1];
  res.half.hor = (int16_t)half_x; res.half.ver = (int16_t)half_y;
  res.cost_half = fs.best_cost;
  __syncthreads();
  // quarter-pel stage: baseRefMv = half << 1, iFrac 1, mv cost in quarter-pel units, cost scale 0 (:6600-6608)
  if (tile_n == 8) frac_stage<8>(fs, s_org, s_src, s_tmp, cols, rows, bit_depth, use_had, half_x << 1, half_y << 1, 1, c_refine_q);
  else             frac_stage<4>(fs, s_org, s_src, s_tmp, cols, rows, bit_depth, use_had, half_x << 1, half_y << 1, 1, c_refine_q);
  if (threadIdx.x == 0) {
    cs.cost_scale = 0;
    const int mx = ((mv_int.hor << 1) + half_x) << 1, my = ((mv_int.ver << 1) + half_y) << 1;
    uint32_t best = HOP_MAX_UINT; int bi = 0;
    for (int i = 0; i < 9; i++) {
      const uint32_t d = (fs.dist[i] >> dist_shift) + mv_cost(cs, mx + c_refine_q[i][0], my + c_refine_q[i][1]);
      if (d < best) { best = d; bi = i; }
    }
    fs.best_i = bi; fs.best_cost = best;
  }
  __syncthreads();
  res.qter.hor = c_refine_q[fs.best_i][0]; res.qter.ver = c_refine_q[fs.best_i][1];
  res.cost = fs.best_cost;
  __syncthreads();
  return res;
}

}  // namespace hop
