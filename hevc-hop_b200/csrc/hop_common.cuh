// hop_common.cuh -- shared device helpers of libhopgpu (sm_100a).
//
// Bit-cost helpers restate TComRdCost::xGetComponentBits / getBits / getCost / getBitsGT
// (TLibCommon/TComRdCost.cpp:270-284, TComRdCost.h:185-216) in closed form.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/hop_gpu.h"

namespace hop {

// exp-Golomb style length: 1 + 2*floor(log2(temp)), temp = (v<=0) ? (-v<<1)+1 : (v<<1)
__device__ __forceinline__ uint32_t component_bits(int v)
{
  uint32_t t = (v <= 0) ? (uint32_t)((-v << 1) + 1) : (uint32_t)(v << 1);
  return 2u * (31u - (uint32_t)__clz((int)t)) + 1u;
}

__device__ __forceinline__ uint32_t mv_bits(const HopCostState& cs, int x, int y)
{
  return component_bits((x << cs.cost_scale) - cs.pred.hor) +
         component_bits((y << cs.cost_scale) - cs.pred.ver);
}

// (m_uiCost * bits) >> 16 in UInt arithmetic
__device__ __forceinline__ uint32_t mv_cost(const HopCostState& cs, int x, int y)
{
  return (cs.lambda_cost * mv_bits(cs, x, y)) >> 16;
}

__device__ __forceinline__ uint32_t bits_cost(const HopCostState& cs, uint32_t bits)
{
  return (cs.lambda_cost * bits) >> 16;
}

// IT_GT_CODING 0, IT_GT_AFFINE 1, W_GT 1: GT0, GT1, GT2 x/y
__device__ __forceinline__ uint32_t gt_bits(int x0, int y0, int x1, int y1, int x2, int y2)
{
  return component_bits(x0) + component_bits(y0) + component_bits(x1) + component_bits(y1) +
         component_bits(x2) + component_bits(y2);
}

__host__ __device__ __forceinline__ bool sad_width_has_subshift(int cols)
{
  return cols == 4 || cols == 8 || cols == 16 || cols == 32 || cols == 64 || cols == 12 ||
         cols == 24 || cols == 48;
}

// One N x N Hadamard tile of the difference org - cur with the reference's per-tile rounding
// (xCalcHADs2x2 / 4x4 / 8x8, TComRdCost.cpp:1366-1575): generic form for the stand-alone kernels (K3, K6).
template <int N>
__device__ __forceinline__ uint32_t had_tile(const int16_t* __restrict__ org, int so,
                                             const int16_t* __restrict__ cur, int sc)
{
  int d[N * N];
#pragma unroll
  for (int y = 0; y < N; y++)
#pragma unroll
    for (int x = 0; x < N; x++) d[y * N + x] = (int)org[y * so + x] - (int)cur[y * sc + x];
#pragma unroll
  for (int y = 0; y < N; y++)
#pragma unroll
    for (int len = 1; len < N; len <<= 1)
#pragma unroll
      for (int i = 0; i < N; i += len << 1)
#pragma unroll
        for (int j = i; j < i + len; j++) {
          int a = d[y * N + j], b = d[y * N + j + len];
          d[y * N + j] = a + b; d[y * N + j + len] = a - b;
        }
#pragma unroll
  for (int x = 0; x < N; x++)
#pragma unroll
    for (int len = 1; len < N; len <<= 1)
#pragma unroll
      for (int i = 0; i < N; i += len << 1)
#pragma unroll
        for (int j = i; j < i + len; j++) {
          int a = d[j * N + x], b = d[(j + len) * N + x];
          d[j * N + x] = a + b; d[(j + len) * N + x] = a - b;
        }
  int s = 0;
#pragma unroll
  for (int k = 0; k < N * N; k++) s += abs(d[k]);
  if (N == 8) return (uint32_t)((s + 2) >> 2);
  if (N == 4) return (uint32_t)((s + 1) >> 1);
  return (uint32_t)s;
}

// block-wide min of a 64-bit key; result valid in every thread. `scratch` holds >= 32 entries.
__device__ __forceinline__ unsigned long long block_min_u64(unsigned long long v,
                                                            unsigned long long* scratch)
{
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = (w < v) ? w : v;
  }
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  const unsigned nwarp = (blockDim.x + 31u) >> 5;
  unsigned long long r = (lane < nwarp) ? scratch[lane] : ~0ull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long w = __shfl_xor_sync(0xffffffffu, r, o);
    r = (w < r) ? w : r;
  }
  __syncthreads();
  return r;
}

}  // namespace hop
