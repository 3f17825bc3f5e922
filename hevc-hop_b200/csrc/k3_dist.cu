// k3_dist.cu -- K3: stand-alone batched distortion kernels (sm_100a).
//
// Replaces the entries of TComRdCost::m_afpDistortFunc that the HOP search uses
// (TLibCommon/TComRdCost.cpp:177-221): the SAD family (:513-1010, row sub-sampling, <<iSubShift,
// >>(bitDepth-8)) and xGetHADs with 8x8 / 4x4 / 2x2 tiles and per-tile rounding (:1366-1708).
// One warp per job; lanes own tiles (HAD) or pixels (SAD).
#include "hop_common.cuh"
#include "hop_internal.h"

namespace hop {

__global__ void __launch_bounds__(128)
k3_dist(int n_jobs, const HopDistJob* __restrict__ jobs, const int16_t* __restrict__ org_buf,
        const int16_t* __restrict__ cur_buf, uint32_t* __restrict__ out)
{
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_jobs) return;
  const HopDistJob job = jobs[warp];
  const int16_t* org = org_buf + job.org_off;
  const int16_t* cur = cur_buf + job.cur_off;
  const int cols = job.cols, rows = job.rows;
  uint32_t sum = 0;
  if (job.func == HOP_DF_HADS) {
    int n = 0;
    if ((rows % 8 == 0) && (cols % 8 == 0)) n = 8;
    else if ((rows % 4 == 0) && (cols % 4 == 0)) n = 4;
    else if ((rows % 2 == 0) && (cols % 2 == 0)) n = 2;
    if (n == 0) { if (lane == 0) out[warp] = HOP_MAX_UINT; return; }   // assert(false) in the reference
    const int tx = cols / n, nt = tx * (rows / n);
    for (int t = lane; t < nt; t += 32) {
      const int x = (t % tx) * n, y = (t / tx) * n;
      const int16_t* o = org + y * job.org_stride + x;
      const int16_t* c = cur + y * job.cur_stride + x;
      if (n == 8) sum += had_tile<8>(o, job.org_stride, c, job.cur_stride);
      else if (n == 4) sum += had_tile<4>(o, job.org_stride, c, job.cur_stride);
      else sum += had_tile<2>(o, job.org_stride, c, job.cur_stride);
    }
  } else {
    const int sub_shift = sad_width_has_subshift(cols) ? job.sub_shift : 0;
    const int step = 1 << sub_shift;
    const int nrow = rows / step;
    for (int i = lane; i < nrow * cols; i += 32) {
      const int r = (i / cols) * step, c = i % cols;
      sum = __sad((int)org[r * job.org_stride + c], (int)cur[r * job.cur_stride + c], sum);
    }
    sum <<= sub_shift;   // distributes over the lane partial sums
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) out[warp] = sum >> (job.bit_depth - 8);
}

cudaError_t dist_launch(int n, const HopDistJob* d_jobs, const int16_t* d_org, const int16_t* d_cur,
                        uint32_t* d_out, cudaStream_t stream, int* launches)
{
  const int warps_per_block = 4;
  k3_dist<<<(n + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, stream>>>(
      n, d_jobs, d_org, d_cur, d_out);
  if (launches) (*launches)++;
  return cudaGetLastError();
}

}  // namespace hop
