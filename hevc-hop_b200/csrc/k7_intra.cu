// k7_intra.cu -- K7: intra mode pre-screen of a PU (sm_100a).  SURVEY.md 8f-4.
//
// Replaces the loop of TEncSearch::estIntraPredQT that ranks the 35 intra modes by prediction signal only
// (TLibEncoder/TEncSearch.cpp:2451-2464): TComPrediction::predIntraLumaAng (TLibCommon/TComPrediction.cpp:316-348)
// = xPredIntraPlanar (:1468-1510) / xPredIntraAng (:192-314, DC value :129-170, edge filter of the pure horizontal
// and vertical modes) / xDCPredFiltering (:1524-1546), each followed by TComRdCost::calcHAD
// (TLibCommon/TComRdCost.cpp:391-425).  Which of the two reference-sample buffers a mode reads follows
// TComPattern::getPredictorPtr (TLibCommon/TComPattern.cpp:49-56, 583-607).
//
// One CTA per PU; the 35 modes are independent, a warp owns a mode at a time: its lanes build the (extended) main
// reference in shared memory, evaluate every prediction sample in closed form (the reference's running sums of the
// planar mode and the flip of the horizontal modes become index arithmetic), then run the Hadamard tiles of the
// warp's private prediction against the original block.  Integer arithmetic only: bit-exact by construction.
#include "hop_common.cuh"
#include "hop_internal.h"

namespace hop {

__constant__ unsigned char c_intra_filter[5] = {10, 7, 1, 0, 10};          // m_aucIntraFilter
__constant__ short c_ang_table[9] = {0, 2, 5, 9, 13, 17, 21, 26, 32};
__constant__ short c_inv_ang_table[9] = {0, 4096, 1638, 910, 630, 482, 390, 315, 256};

constexpr int K7_WARPS = 8;

__host__ __device__ inline size_t k7_smem_bytes(int n)
{
  // [refs 4 x (2n+1) int32][org n x n int16][per warp: main reference 3n+2 int16, prediction n x n int16]
  size_t b = (sizeof(int32_t) * 4 * (size_t)(2 * n + 1) + 15) & ~(size_t)15;
  b += (sizeof(int16_t) * (size_t)n * n + 15) & ~(size_t)15;
  b += K7_WARPS * (((sizeof(int16_t) * (size_t)(3 * n + 2) + 15) & ~(size_t)15) + ((sizeof(int16_t) * (size_t)n * n + 15) & ~(size_t)15));
  return b;
}

__global__ void __launch_bounds__(K7_WARPS * 32)
k7_intra_prescreen(int n_jobs, const HopIntraJob* __restrict__ jobs, const int16_t* __restrict__ org_buf,
                   const int32_t* __restrict__ refs_buf, uint32_t* __restrict__ out)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const int job_id = blockIdx.x;
  if (job_id >= n_jobs) return;
  const HopIntraJob job = jobs[job_id];
  const int n = job.size, sw = 2 * n + 1;
  const int lg = 31 - __clz(n);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t* s_refs = reinterpret_cast<int32_t*>(smem);
  int16_t* s_org = reinterpret_cast<int16_t*>(smem + ((sizeof(int32_t) * 4 * (size_t)sw + 15) & ~(size_t)15));
  unsigned char* per_warp = reinterpret_cast<unsigned char*>(s_org) + ((sizeof(int16_t) * (size_t)n * n + 15) & ~(size_t)15);
  const size_t main_bytes = (sizeof(int16_t) * (size_t)(3 * n + 2) + 15) & ~(size_t)15;
  const size_t pred_bytes = (sizeof(int16_t) * (size_t)n * n + 15) & ~(size_t)15;
  int16_t* s_main = reinterpret_cast<int16_t*>(per_warp + warp * (main_bytes + pred_bytes));
  int16_t* s_pred = reinterpret_cast<int16_t*>(reinterpret_cast<unsigned char*>(s_main) + main_bytes);

  for (int i = threadIdx.x; i < 4 * sw; i += blockDim.x) s_refs[i] = refs_buf[job.refs_off + i];
  {
    const int16_t* org = org_buf + job.org_off;
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) s_org[i] = org[(i / n) * job.org_stride + (i % n)];
  }
  __syncthreads();

  const bool filter_edges = n <= 16;                                   // predIntraLumaAng :333-346
  const int max_val = (1 << job.bit_depth) - 1;
  for (int mode = warp; mode < HOP_INTRA_MODES; mode += K7_WARPS) {
    // getPredictorPtr: filtered or unfiltered reference samples
    const int d10 = abs(mode - 10), d26 = abs(mode - 26);
    int filt = (d10 < d26 ? d10 : d26) > (int)c_intra_filter[lg - 2] ? 1 : 0;
    if (mode == 1) filt = 0;
    const int32_t* above = s_refs + (filt ? 2 : 0) * sw;
    const int32_t* left = above + sw;

    if (mode == 0) {                                                   // planar, closed form of :1497-1509
      const int shift1 = lg, shift2 = lg + 1;
      const int bottom_left = left[n + 1], top_right = above[n + 1];
      for (int i = lane; i < n * n; i += 32) {
        const int k = i / n, l = i - k * n;
        const int lc = left[k + 1], tr = above[l + 1];
        const int hor = (lc << shift1) + n + (l + 1) * (top_right - lc);
        const int ver = (tr << shift1) + (k + 1) * (bottom_left - tr);
        s_pred[i] = (int16_t)((hor + ver) >> shift2);
      }
    } else if (mode == 1) {                                            // DC
      int sum = 0;
      if (job.above_avail) for (int i = lane; i < n; i += 32) sum += above[i + 1];
      if (job.left_avail) for (int i = lane; i < n; i += 32) sum += left[i + 1];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      int dc;
      if (job.above_avail && job.left_avail) dc = (sum + n) / (n + n);
      else if (job.above_avail || job.left_avail) dc = (sum + n / 2) / n;
      else dc = left[1];
      dc = (int)(int16_t)dc;
      const bool f = filter_edges && job.above_avail && job.left_avail;   // xDCPredFiltering
      for (int i = lane; i < n * n; i += 32) {
        const int k = i / n, l = i - k * n;
        int v = dc;
        if (f) {
          if (k == 0 && l == 0) v = (above[1] + left[1] + 2 * dc + 2) >> 2;
          else if (k == 0) v = (above[l + 1] + 3 * dc + 2) >> 2;
          else if (l == 0) v = (left[k + 1] + 3 * dc + 2) >> 2;
        }
        s_pred[i] = (int16_t)v;
      }
    } else {                                                           // angular, xPredIntraAng :192-314
      const bool mode_hor = mode < 18, mode_ver = !mode_hor;
      int angle = mode_ver ? mode - 26 : -(mode - 10);
      const int abs_idx = abs(angle), sign = angle < 0 ? -1 : 1;
      const int inv_angle = c_inv_ang_table[abs_idx];
      angle = sign * (int)c_ang_table[abs_idx];
      const int32_t* main_src = mode_ver ? above : left;
      const int32_t* side_src = mode_ver ? left : above;
      // M(j) = s_main[j + n], j in [-n, 2n]
      if (angle < 0) {
        for (int j = lane; j <= n; j += 32) s_main[j + n] = (int16_t)main_src[j];
        const int last = (n * angle) >> 5;                             // extend the main reference to the left (:253-259)
        for (int k = -1 - lane; k > last; k -= 32) s_main[k + n] = (int16_t)side_src[(128 + (-k) * inv_angle) >> 8];
      } else {
        for (int j = lane; j <= 2 * n; j += 32) s_main[j + n] = (int16_t)main_src[j];
      }
      __syncwarp();
      const int16_t* M = s_main + n;
      for (int i = lane; i < n * n; i += 32) {
        const int k = i / n, l = i - k * n;
        int v;
        if (angle == 0) {
          v = M[l + 1];
          if (filter_edges && l == 0) v = min(max(v + (((int)(int16_t)side_src[k + 1] - (int)(int16_t)side_src[0]) >> 1), 0), max_val);
        } else {
          const int pos = (k + 1) * angle, di = pos >> 5, fr = pos & 31, idx = l + di + 1;
          v = fr ? (int)(int16_t)(((32 - fr) * M[idx] + fr * M[idx + 1] + 16) >> 5) : M[idx];
        }
        s_pred[mode_hor ? l * n + k : i] = (int16_t)v;                  // horizontal modes: the flip of :300-313
      }
    }
    __syncwarp();
    // calcHAD: 8x8 tiles when the size is a multiple of 8, 4x4 otherwise
    unsigned int sum = 0;
    if ((n & 7) == 0) {
      const int tx = n >> 3, nt = tx * tx;
      for (int t = lane; t < nt; t += 32) {
        const int x = (t % tx) * 8, y = (t / tx) * 8;
        sum += had_tile<8>(s_org + y * n + x, n, s_pred + y * n + x, n);
      }
    } else {
      const int tx = n >> 2, nt = tx * tx;
      for (int t = lane; t < nt; t += 32) {
        const int x = (t % tx) * 4, y = (t / tx) * 4;
        sum += had_tile<4>(s_org + y * n + x, n, s_pred + y * n + x, n);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) out[(size_t)job_id * HOP_INTRA_MODES + mode] = sum >> (job.bit_depth - 8);
    __syncwarp();
  }
}

cudaError_t intra_launch(int n, const HopIntraJob* d_jobs, const int16_t* d_org, const int32_t* d_refs, uint32_t* d_out,
                         int max_size, cudaStream_t stream, int* launches)
{
  static SmemOptIn opt_in;
  {
    cudaError_t e = opt_in.ensure(k7_intra_prescreen, (int)k7_smem_bytes(HOP_MAX_PU));
    if (e != cudaSuccess) return e;
  }
  k7_intra_prescreen<<<n, K7_WARPS * 32, k7_smem_bytes(max_size), stream>>>(n, d_jobs, d_org, d_refs, d_out);
  if (launches) (*launches)++;
  return cudaGetLastError();
}

}  // namespace hop
