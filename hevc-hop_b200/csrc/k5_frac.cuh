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
// horizontal intermediates are built once in shared memory; the vertical filter, the difference to the
// original block and the Hadamard butterflies run in two fine-grained phases (see frac_stage).
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

// positions handled per shared-memory round: all 9 at once for PUs up to 1024 samples (one round, two barriers),
// 3 rounds of 3 beyond that (keeps the scratch of a 64x64 PU at 24 KB)
__host__ __device__ inline int frac_chunk(int cols, int rows) { return cols * rows <= 1024 ? 9 : 3; }

__host__ __device__ inline size_t frac_smem_bytes(int cols, int rows)
{
  // [source region (rows+8) x (cols+8) int16][3 horizontal intermediates (rows+8) x cols int16]
  // [row-transformed differences of frac_chunk() positions, rows x cols int16]
  return (sizeof(int16_t) * ((size_t)(rows + 8) * (cols + 8) + 3 * (size_t)(rows + 8) * cols +
                             (size_t)frac_chunk(cols, rows) * rows * cols) + 15) & ~(size_t)15;
}

// one refinement stage: positions base + step * refine[i] (quarter-pel units), i = 0..8.
//   s_org : original block, int32, stride cols          s_src : staged source region
//   s_tmp : 3 horizontal intermediates, plane index = refine x-component + 1
//   s_rt  : row-transformed differences of frac_chunk() positions
// Two phases per round so that even an 8x8 PU keeps ~70 threads busy: (A) one thread per tile ROW filters
// its N pixels vertically, subtracts the original and runs the horizontal N-point butterflies; (B) one
// thread per tile COLUMN runs the vertical butterflies; the N column sums of a tile meet by shuffles for
// the per-tile rounding.
template <int N>
__device__ __forceinline__ void frac_stage(FracShared& fs, const int* __restrict__ s_org, const int16_t* __restrict__ s_src,
                                           int16_t* __restrict__ s_tmp, int16_t* __restrict__ s_rt, int cols, int rows,
                                           int bit_depth, int use_had, int base_qx, int base_qy, int step,
                                           const int8_t (*refine)[2])
{
  const int head = 14 - bit_depth;                     // IF_INTERNAL_PREC - bitDepth
  const int src_w = cols + 8, plane = (rows + 8) * cols, blk = rows * cols;
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
  const int tiles_x = cols / N;
  const int max_val = (1 << bit_depth) - 1;
  const int FRAC_CHUNK = frac_chunk(cols, rows);
  for (int i0 = 0; i0 < 9; i0 += FRAC_CHUNK) {
    // phase A: (position, row y, tile column tx) -> N pixels
    for (int t = threadIdx.x; t < FRAC_CHUNK * rows * tiles_x; t += blockDim.x) {
      const int tx = t % tiles_x, y = (t / tiles_x) % rows, il = t / (tiles_x * rows), i = i0 + il;
      const int qy = base_qy + step * refine[i][1];
      const int iy = qy >> 2, fy = qy & 3;
      const int16_t* tp = s_tmp + (refine[i][0] + 1) * plane + (y + iy + 1) * cols + tx * N;   // source row y+iy-3 is tmp row +4
      int d[N];
#pragma unroll
      for (int c = 0; c < N; c++) {
        int v;
        if (fy == 0) {
          int16_t off = (int16_t)8192;                 // filterCopy isLast, :135-149
          off = (int16_t)(off + (head ? (1 << (head - 1)) : 0));
          v = (int16_t)((tp[3 * cols + c] + off) >> head);
        } else {
          int sum = 0;
#pragma unroll
          for (int k = 0; k < 8; k++) sum += (int)tp[k * cols + c] * c_luma_filter[fy][k];
          const int shift = 6 + head;
          v = (int16_t)((sum + (1 << (shift - 1)) + (8192 << 6)) >> shift);
        }
        v = min(max(v, 0), max_val);
        d[c] = s_org[y * cols + tx * N + c] - v;
      }
      if (use_had) {
#pragma unroll
        for (int len = 1; len < N; len <<= 1)
#pragma unroll
          for (int a = 0; a < N; a += len << 1)
#pragma unroll
            for (int j = a; j < a + len; j++) { const int u = d[j], w = d[j + len]; d[j] = u + w; d[j + len] = u - w; }
#pragma unroll
        for (int c = 0; c < N; c++) s_rt[il * blk + y * cols + tx * N + c] = (int16_t)d[c];
      } else {
        unsigned sum = 0;
#pragma unroll
        for (int c = 0; c < N; c++) sum = __sad(d[c], 0, sum);
        atomicAdd(&fs.dist[i], sum);
      }
    }
    __syncthreads();
    if (use_had) {
      // phase B: (position, tile row ty, column x): vertical butterflies; x is the fastest index, so the N
      // columns of a tile sit in N adjacent lanes
      const int ntask = FRAC_CHUNK * (rows / N) * cols;
      for (int t0 = 0; t0 < ntask; t0 += blockDim.x) {
        const int t = t0 + threadIdx.x;
        unsigned sum = 0;
        int i = 0;
        if (t < ntask) {
          const int x = t % cols, ty = (t / cols) % (rows / N), il = t / (cols * (rows / N));
          i = i0 + il;
          int d[N];
#pragma unroll
          for (int r = 0; r < N; r++) d[r] = s_rt[il * blk + (ty * N + r) * cols + x];
#pragma unroll
          for (int len = 1; len < N; len <<= 1)
#pragma unroll
            for (int a = 0; a < N; a += len << 1)
#pragma unroll
              for (int j = a; j < a + len; j++) { const int u = d[j], w = d[j + len]; d[j] = u + w; d[j + len] = u - w; }
#pragma unroll
          for (int r = 0; r < N; r++) sum = __sad(d[r], 0, sum);
        }
        // tile sum over its N columns (lanes x .. x+N-1 are aligned to N because cols % N == 0)
#pragma unroll
        for (int o = 1; o < N; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (t < ntask && (threadIdx.x & (N - 1)) == 0)
          atomicAdd(&fs.dist[i], N == 8 ? (sum + 2) >> 2 : (sum + 1) >> 1);   // xCalcHADs8x8 / 4x4 rounding
      }
      __syncthreads();
    }
  }
}

// Whole xPatternSearchFracDIF for one PU by one CTA.  s_org must hold the original block (int32, stride
// cols); `scratch` provides frac_smem_bytes(cols, rows) bytes.  Result valid in every thread on return.
__device__ __forceinline__ HopFracResult frac_search_cta(FracShared& fs, const int* __restrict__ s_org,
                                                         unsigned char* scratch, const int16_t* __restrict__ ref_pos,
                                                         int ref_stride, int cols, int rows, int bit_depth, int use_had,
                                                         HopCostState cs, HopMv mv_int, bool src_staged = false)
{
  int16_t* s_src = reinterpret_cast<int16_t*>(scratch);
  int16_t* s_tmp = s_src + (rows + 8) * (cols + 8);
  int16_t* s_rt = s_tmp + 3 * (rows + 8) * cols;
  const int src_w = cols + 8;
  // source region: x in [-4, cols+4), y in [-4, rows+4) around the integer position (:6579); the fused
  // single-PU kernel stages it together with the first HOP window (one global-memory latency for both)
  if (!src_staged)
    for (int i = threadIdx.x; i < (rows + 8) * src_w; i += blockDim.x) {
      const int y = i / src_w, x = i - y * src_w;
      s_src[i] = ref_pos[(long long)(y - 4) * ref_stride + (x - 4)];
    }
  __syncthreads();
  const int tile_n = ((rows % 8 == 0) && (cols % 8 == 0)) ? 8 : 4;
  const int dist_shift = bit_depth - 8;
  HopFracResult res;
  // half-pel stage: baseRefMv (0,0), iFrac 2, mv cost in half-pel units, cost scale 1 (:4615, 6594-6598)
  if (tile_n == 8) frac_stage<8>(fs, s_org, s_src, s_tmp, s_rt, cols, rows, bit_depth, use_had, 0, 0, 2, c_refine_h);
  else             frac_stage<4>(fs, s_org, s_src, s_tmp, s_rt, cols, rows, bit_depth, use_had, 0, 0, 2, c_refine_h);
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
  if (tile_n == 8) frac_stage<8>(fs, s_org, s_src, s_tmp, s_rt, cols, rows, bit_depth, use_had, half_x << 1, half_y << 1, 1, c_refine_q);
  else             frac_stage<4>(fs, s_org, s_src, s_tmp, s_rt, cols, rows, bit_depth, use_had, half_x << 1, half_y << 1, 1, c_refine_q);
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
