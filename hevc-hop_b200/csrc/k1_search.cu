// k1_search.cu -- K1: self-similarity full-search block match (sm_100a).
//
// Replaces TEncSearch::xPatternSearch (TLibEncoder/TEncSearch.cpp:6262-6371) with the SAD family of
// TComRdCost (TLibCommon/TComRdCost.cpp:513-1010), isValidPattern (:444-458) and getCost
// (TComRdCost.h:185-202).
//
// Parallel restatement: every (x,y) of the causal window is independent; the serial loop's "first
// strict minimum in raster order" equals the minimum of the 64-bit key (cost << 32 | rasterIndex)
// over the positions that pass the SS gates.  The window of one PU is split over `slices` CTAs (so a
// single in-encoder call can use many SMs); CTAs merge through atomicMin on one 64-bit word per PU.
#include "hop_common.cuh"
#include "hop_internal.h"

namespace hop {

__global__ void k1_init_keys(int n, unsigned long long* keys)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = ~0ull;
}

// v1: one thread per search position, original block in shared memory, reference rows read through
// the read-only path (adjacent threads read overlapping rows, so L1 serves most of it).
__global__ void __launch_bounds__(K1_THREADS)
k1_search(int n_jobs, const HopSearchJob* __restrict__ jobs, const int16_t* __restrict__ org_buf,
          const int16_t* __restrict__ ref_buf, unsigned long long* __restrict__ keys)
{
  __shared__ int16_t s_org[HOP_MAX_PU * HOP_MAX_PU];
  __shared__ unsigned long long s_red[32];
  const int job_id = blockIdx.x;
  const HopSearchJob job = jobs[job_id];
  const int cols = job.cols, rows = job.rows;
  const int nx = job.rng_right - job.rng_left + 1;
  const int ny = job.rng_bottom - job.rng_top + 1;
  if (nx <= 0 || ny <= 0) return;
  int sub_shift = (job.fast_enc && rows > 8) ? 1 : 0;           // TEncSearch.cpp:6303-6309
  if (!sad_width_has_subshift(cols)) sub_shift = 0;             // generic xGetSAD ignores it
  const int step = 1 << sub_shift;
  const int dist_shift = job.bit_depth - 8;
  const int16_t* org = org_buf + job.org_off;
  const int16_t* ref_y = ref_buf + job.ref_off;
  const int stride = job.ref_stride;

  for (int i = threadIdx.x; i < rows * cols; i += blockDim.x)
    s_org[i] = org[(i / cols) * job.org_stride + (i % cols)];
  __syncthreads();

  // rows of the window handled by this slice
  const int slices = gridDim.y, slice = blockIdx.y;
  const int rows_per = (ny + slices - 1) / slices;
  const int y_lo = slice * rows_per, y_hi = min(ny, y_lo + rows_per);
  unsigned long long best = ~0ull;
  for (int idx = y_lo * nx + threadIdx.x; idx < y_hi * nx; idx += blockDim.x) {
    const int py = idx / nx, px = idx - py * nx;
    const int x = job.rng_left + px, y = job.rng_top + py;
    const int16_t* srch = ref_y + (long long)y * stride + x;
    if (job.is_ss) {
      if ((x >= job.offset_x) && (y > job.offset_y)) continue;              // :6328
      const int16_t* lb = srch + (rows + 4) * stride;                        // isValidPattern
      if (__ldg(lb) == HOP_NOT_VALID || __ldg(lb + cols + 4) == HOP_NOT_VALID) continue;   // :6330
    }
    uint32_t sum = 0;
    for (int r = 0; r < rows; r += step) {
      const int16_t* rr = srch + r * stride;
      const int16_t* oo = s_org + r * cols;
#pragma unroll 4
      for (int c = 0; c < cols; c++) sum = __sad((int)oo[c], (int)__ldg(rr + c), sum);
    }
    sum <<= sub_shift;
    sum >>= dist_shift;
    sum += mv_cost(job.cost, x, y);                                          // :6336
    const unsigned long long key = ((unsigned long long)sum << 32) | (unsigned)idx;
    best = key < best ? key : best;
  }
  best = block_min_u64(best, s_red);
  if (threadIdx.x == 0 && best != ~0ull) atomicMin(&keys[job_id], best);
}

__global__ void k1_finalize(int n, const HopSearchJob* __restrict__ jobs,
                            const unsigned long long* __restrict__ keys, HopSearchResult* __restrict__ out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const HopSearchJob job = jobs[i];
  const unsigned long long key = keys[i];
  HopSearchResult r;
  r.mv.hor = 0; r.mv.ver = 0;
  if (key == ~0ull) {                                            // :6356-6360
    r.found = 0; r.sad = HOP_MAX_UINT; r.cost = HOP_MAX_UINT;
  } else {
    const int nx = job.rng_right - job.rng_left + 1;
    const unsigned idx = (unsigned)(key & 0xffffffffu);
    const int x = job.rng_left + (int)(idx % (unsigned)nx), y = job.rng_top + (int)(idx / (unsigned)nx);
    r.found = 1;
    r.mv.hor = (int16_t)x; r.mv.ver = (int16_t)y;
    r.cost = (uint32_t)(key >> 32);
    r.sad = r.cost - mv_cost(job.cost, x, y);                    // :6365
  }
  out[i] = r;
}

cudaError_t search_launch(int n, const HopSearchJob* d_jobs, const int16_t* d_org, const int16_t* d_ref,
                          HopSearchResult* d_out, unsigned long long* d_keys, int slices,
                          cudaStream_t stream, int* launches)
{
  if (slices < 1) slices = 1;
  if (slices > K1_MAX_SLICES) slices = K1_MAX_SLICES;
  k1_init_keys<<<(n + 255) / 256, 256, 0, stream>>>(n, d_keys);
  k1_search<<<dim3(n, slices), K1_THREADS, 0, stream>>>(n, d_jobs, d_org, d_ref, d_keys);
  k1_finalize<<<(n + 255) / 256, 256, 0, stream>>>(n, d_jobs, d_keys, d_out);
  if (launches) *launches += 3;
  return cudaGetLastError();
}

}  // namespace hop
