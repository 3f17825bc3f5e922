// k6_pred.cu -- K6: motion-compensated prediction of a PU (uni-prediction) and the costs the encoder derives
// from it (sm_100a).  SURVEY.md 8f-3.
//
// Replaces TComPrediction::xPredInterLumaBlk / xPredInterChromaBlk (TLibCommon/TComPrediction.cpp:639-720,
// 1235-1347) with the DCT-IF of TComInterpolationFilter (TComInterpolationFilter.cpp:92-420), the GT branches
// xPredGTLuma / xPredGTChroma (:723-805, 1351-1420) with calcParamProjective / calcParamProjectiveC (:807-859)
// and ProjectiveTransform (:904-1030), and on top of the prediction
//   * TEncSearch::xGetInterPredictionError (TLibEncoder/TEncSearch.cpp:2951-2977): SAD / Hadamard against the original
//   * TEncSearch::xGetTemplateCost (:4390-4477): isValidPattern gate (TComRdCost.cpp:430-442), SAD, calcRdCost DF_SAD.
//
// One CTA per job, three phases in shared memory: (1) the source region the reference reads -- block or 2W x 2H
// region plus the filter support that the vector's fraction needs, nothing more -- is staged as int16; (2) the
// interpolation runs per output sample with the reference's two roundings (14-bit intermediate, then the final
// shift and clip; an integer vector is a plain copy, so NOT_VALID samples stay -1 exactly as in the reference);
// (3) for a GT the region is warped with the reference's LITERAL binary64 operation sequence (this kernel produces
// the samples that are coded, not a cost to be compared, so nothing is restructured), incl. the general projective
// division.  Prediction, distortion and template cost leave through the result / dst buffers.
#include "hop_common.cuh"
#include "hop_internal.h"

namespace hop {

__constant__ int8_t c_mc_luma[4][8] = {     // m_lumaFilter, TComInterpolationFilter.cpp:55-61
  {0, 0, 0, 64, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0}, {-1, 4, -11, 40, 40, -11, 4, -1}, {0, 1, -5, 17, 58, -10, 4, -1}};
__constant__ int8_t c_mc_chroma[8][4] = {   // m_chromaFilter, :63-73
  {0, 64, 0, 0}, {-2, 58, 10, -2}, {-4, 54, 16, -2}, {-6, 46, 28, -4}, {-4, 36, 36, -4}, {-4, 28, 46, -6}, {-2, 16, 54, -4}, {-2, 10, 58, -2}};

struct PredShared {
  double h[9];
  unsigned int dist;
  int invalid;
};

// one output sample of the three-way branch of xPredInter{Luma,Chroma}Blk (bi == false); s points at the staged
// sample that corresponds to the output sample's integer position
template <int TAPS>
__device__ __forceinline__ int mc_sample(const int16_t* __restrict__ s, int sw, int fx, int fy, int bit_depth)
{
  constexpr int half = TAPS / 2 - 1;
  const int head = 14 - bit_depth;
  const int max_val = (1 << bit_depth) - 1;
  if (fy == 0) {
    if (fx == 0) return s[0];                                              // filterCopy, isFirst == isLast
    int sum = 0;
#pragma unroll
    for (int k = 0; k < TAPS; k++) sum += s[k - half] * (TAPS == 8 ? c_mc_luma[fx][k] : c_mc_chroma[fx][k]);
    const int v = (int)(int16_t)((sum + 32) >> 6);                         // Short val = (sum + offset) >> shift
    return min(max(v, 0), max_val);
  }
  if (fx == 0) {
    int sum = 0;
#pragma unroll
    for (int k = 0; k < TAPS; k++) sum += s[(k - half) * sw] * (TAPS == 8 ? c_mc_luma[fy][k] : c_mc_chroma[fy][k]);
    const int v = (int)(int16_t)((sum + 32) >> 6);
    return min(max(v, 0), max_val);
  }
  int sum2 = 0;
  const int shift1 = 6 - head, off1 = -8192 << shift1;
#pragma unroll
  for (int r = 0; r < TAPS; r++) {
    const int16_t* row = s + (r - half) * sw;
    int sum = 0;
#pragma unroll
    for (int k = 0; k < TAPS; k++) sum += row[k - half] * (TAPS == 8 ? c_mc_luma[fx][k] : c_mc_chroma[fx][k]);
    const int t = (int)(int16_t)((sum + off1) >> shift1);                  // isFirst && !isLast, stored as Short
    sum2 += t * (TAPS == 8 ? c_mc_luma[fy][r] : c_mc_chroma[fy][r]);
  }
  const int shift2 = 6 + head, off2 = (1 << (shift2 - 1)) + (8192 << 6);   // !isFirst && isLast
  const int v = (int)(int16_t)((sum2 + off2) >> shift2);
  return min(max(v, 0), max_val);
}

// calcParamProjective / calcParamProjectiveC in the reference's operation order (binary64, no contraction);
// x, y are the corners as doubles (exact integers for luma, multiples of 0.5 for chroma)
__device__ void calc_param(const double x[4], const double y[4], double h[9], int width, int height, bool chroma_form,
                           const int xi[4], const int yi[4])
{
  const double W = __dsub_rn((double)width, 1.0), H = __dsub_rn((double)height, 1.0);
  double dx1 = __dsub_rn(x[1], x[2]), dx2 = __dsub_rn(x[3], x[2]);
  double dx3 = __dsub_rn(__dadd_rn(__dsub_rn(x[0], x[1]), x[2]), x[3]);
  double dy1 = __dsub_rn(y[1], y[2]), dy2 = __dsub_rn(y[3], y[2]);
  double dy3 = __dsub_rn(__dadd_rn(__dsub_rn(y[0], y[1]), y[2]), y[3]);
  const double den = __dsub_rn(__dmul_rn(dx1, dy2), __dmul_rn(dx2, dy1));
  h[2] = __ddiv_rn(__ddiv_rn(__dsub_rn(__dmul_rn(dx3, dy2), __dmul_rn(dx2, dy3)), den), W);
  h[5] = __ddiv_rn(__ddiv_rn(__dsub_rn(__dmul_rn(dx1, dy3), __dmul_rn(dx3, dy1)), den), H);
  // luma: (Double)(x[1] - x[0]) is an INTEGER subtraction then a conversion (:824); chroma subtracts doubles (:851)
  const double x10 = chroma_form ? __dsub_rn(x[1], x[0]) : (double)(xi[1] - xi[0]);
  const double x30 = chroma_form ? __dsub_rn(x[3], x[0]) : (double)(xi[3] - xi[0]);
  const double y10 = chroma_form ? __dsub_rn(y[1], y[0]) : (double)(yi[1] - yi[0]);
  const double y30 = chroma_form ? __dsub_rn(y[3], y[0]) : (double)(yi[3] - yi[0]);
  h[0] = __dadd_rn(__ddiv_rn(x10, W), __dmul_rn(h[2], x[1]));
  h[3] = __dadd_rn(__ddiv_rn(x30, H), __dmul_rn(h[5], x[3]));
  h[6] = x[0];
  h[1] = __dadd_rn(__ddiv_rn(y10, W), __dmul_rn(h[2], y[1]));
  h[4] = __dadd_rn(__ddiv_rn(y30, H), __dmul_rn(h[5], y[3]));
  h[7] = y[0];
  h[8] = 1.0;
}

// ProjectiveTransform (IT_GT_GRID_SIZE 2, bilinear), one output sample (xo, yo) of the bw x bh block, literal
// operation sequence of TComPrediction.cpp:925-972, 1025.  reg: the 2bw x 2bh region, stride W2.
__device__ __forceinline__ int warp_literal(const int16_t* __restrict__ reg, int W2, const double* h, int bw, int bh, int xo, int yo)
{
  const int W = 2 * bw, H = 2 * bh;
  const int off_x = W / 2 - (W / 2 / 2), off_y = H / 2 - (H / 2 / 2);
  const int nss = (((bh < bw) ? bh : bw) >> 1) * 2;
  const double x = (double)(off_x + xo), y = (double)(off_y + yo);
  const double den = __dadd_rn(__dadd_rn(__dmul_rn(h[2], x), __dmul_rn(h[5], y)), h[8]);
  const double Fx = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(h[0], x), __dmul_rn(h[3], y)), h[6]), den);
  const double Fy = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(h[1], x), __dmul_rn(h[4], y)), h[7]), den);
  int Y = __double2int_rz(Fy) - off_y;
  int X = __double2int_rz(Fx) - off_x;
  const double q = __dsub_rn(__dsub_rn(Fy, (double)off_y), (double)Y);
  const double p = __dsub_rn(__dsub_rn(Fx, (double)off_x), (double)X);
  const int lo = -nss / 2, hy = nss / 2 + H / 2 - 1, hx = nss / 2 + W / 2 - 1;
  if (Y < lo) Y = lo;
  if (X < lo) X = lo;
  if (Y > hy) Y = hy;
  if (X > hx) X = hx;
  if (Y + 1 > hy) Y = hy - 1;
  if (X + 1 > hx) X = hx - 1;
  // piRefSrch = region + (bw/2, bh/2)
  const int16_t* pa = reg + (bh / 2 + Y) * W2 + bw / 2 + X;
  const double omp = __dsub_rn(1.0, p), omq = __dsub_rn(1.0, q);
  double aux = __dmul_rn(omq, __dadd_rn(__dmul_rn(omp, (double)pa[0]), __dmul_rn(p, (double)pa[1])));
  aux = __dadd_rn(aux, __dmul_rn(q, __dadd_rn(__dmul_rn(omp, (double)pa[W2]), __dmul_rn(p, (double)pa[W2 + 1]))));
  if (aux > 255) aux = 255;                          // hard-coded 8-bit clip (:969-972)
  if (aux < 0) aux = 0;
  return (int)(int16_t)__double2int_rz(__dadd_rn(aux, 0.5));
}

__host__ __device__ inline size_t k6_smem_bytes(int bw, int bh, int taps, bool gt)
{
  const int rw = gt ? 2 * bw : bw, rh = gt ? 2 * bh : bh;
  size_t b = sizeof(PredShared) + 16;
  b += (sizeof(int16_t) * (size_t)(rw + taps - 1) * (rh + taps - 1) + 15) & ~(size_t)15;   // staged source
  if (gt) b += (sizeof(int16_t) * (size_t)rw * rh + 15) & ~(size_t)15;                      // interpolated region
  b += (sizeof(int16_t) * (size_t)bw * bh + 15) & ~(size_t)15;                              // prediction
  return b;
}

__global__ void __launch_bounds__(256)
k6_predict(int n_jobs, const HopPredJob* __restrict__ jobs, const int16_t* __restrict__ org_buf,
           const int16_t* __restrict__ ref_buf, int16_t* __restrict__ dst_buf, HopPredResult* __restrict__ out, RefBounds rb)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const int job_id = blockIdx.x;
  if (job_id >= n_jobs) return;
  const HopPredJob job = jobs[job_id];
  PredShared& sh = *reinterpret_cast<PredShared*>(smem);
  const bool chroma = job.comp != 0;
  const int bw = chroma ? job.cols >> 1 : job.cols, bh = chroma ? job.rows >> 1 : job.rows;
  const int taps = chroma ? 4 : 8, half = taps / 2 - 1;
  const int fmask = chroma ? 7 : 3, fshift = chroma ? 3 : 2;
  const int fx = job.mv.hor & fmask, fy = job.mv.ver & fmask;
  const int gt_any = job.gt[0].hor | job.gt[1].hor | job.gt[2].hor | job.gt[3].hor |
                     job.gt[0].ver | job.gt[1].ver | job.gt[2].ver | job.gt[3].ver;
  const bool gt = job.gt_flag && gt_any && !job.template_cost;            // :651 / :1246
  const int stride = job.ref_stride;

  if (threadIdx.x == 0) {
    sh.dist = 0;
    int bad = 0;
    if (job.template_cost && job.is_ss) {                                  // TEncSearch.cpp:4420-4436
      long long o = job.ref_off + (job.mv_probe.hor >> 2) + (long long)((job.mv_probe.ver >> 2) + job.rows + 4) * stride;
      long long o2 = o + job.cols + 4;
      o = o < rb.lo ? rb.lo : (o > rb.hi ? rb.hi : o);
      o2 = o2 < rb.lo ? rb.lo : (o2 > rb.hi ? rb.hi : o2);
      bad = ref_buf[o] == HOP_NOT_VALID || ref_buf[o2] == HOP_NOT_VALID;
    }
    sh.invalid = bad;
  }
  __syncthreads();
  if (sh.invalid) {
    if (threadIdx.x == 0) { HopPredResult r; r.valid = 0; r.dist = 0; r.cost = 0x7fffffffu; out[job_id] = r; }
    return;
  }

  // (1) stage what the reference reads: block or 2W x 2H region, plus the filter support the fraction needs
  const int rw = gt ? 2 * bw : bw, rh = gt ? 2 * bh : bh;
  const int ox = gt ? (chroma ? job.cols / 4 : job.cols / 2) : 0, oy = gt ? (chroma ? job.rows / 4 : job.rows / 2) : 0;
  const int left = fx ? half : 0, top = fy ? half : 0;
  const int sw = rw + (fx ? taps - 1 : 0), shh = rh + (fy ? taps - 1 : 0);
  int16_t* s_src = reinterpret_cast<int16_t*>(smem + ((sizeof(PredShared) + 15) & ~(size_t)15));
  int16_t* s_reg = s_src + ((((size_t)(rw + taps - 1) * (rh + taps - 1)) + 7) & ~(size_t)7);
  int16_t* s_pred = gt ? s_reg + (((size_t)rw * rh + 7) & ~(size_t)7) : s_reg;
  {
    const long long base = job.ref_off + (job.mv.hor >> fshift) - ox - left + (long long)((job.mv.ver >> fshift) - oy - top) * stride;
    for (int i = threadIdx.x; i < sw * shh; i += blockDim.x) {
      const int y = i / sw, x = i - y * sw;
      long long o = base + (long long)y * stride + x;
      o = o < rb.lo ? rb.lo : (o > rb.hi ? rb.hi : o);        // a vector may point past the buffer: keep the read inside
      s_src[i] = ref_buf[o];
    }
  }
  if (gt && threadIdx.x == 0) {
    // corners of xPredGTLuma / xPredGTChroma (:748-774, 1370-1396) on the component's block size
    int nss = (((bh < bw) ? bh : bw) >> 1) * 2;
    int last_step = nss >> 6;
    if (last_step == 0) last_step = 1;
    double cx[4], cy[4];
    int xi[4] = {0, 0, 0, 0}, yi[4] = {0, 0, 0, 0};
    if (!chroma) {
      xi[0] = job.gt[0].hor * last_step;              yi[0] = job.gt[0].ver * last_step;
      xi[1] = job.gt[1].hor * last_step + bw * 2 - 1; yi[1] = job.gt[1].ver * last_step;
      xi[2] = job.gt[2].hor * last_step + bw * 2 - 1; yi[2] = job.gt[2].ver * last_step + bh * 2 - 1;
      xi[3] = job.gt[3].hor * last_step;              yi[3] = job.gt[3].ver * last_step + bh * 2 - 1;
      for (int k = 0; k < 4; k++) { cx[k] = (double)xi[k]; cy[k] = (double)yi[k]; }
    } else {
      const double ls = (double)last_step;
      const double wx = (double)(bw * 2 - 1), hy = (double)(bh * 2 - 1);
      cx[0] = __dmul_rn(__ddiv_rn((double)job.gt[0].hor, 2.0), ls);                cy[0] = __dmul_rn(__ddiv_rn((double)job.gt[0].ver, 2.0), ls);
      cx[1] = __dadd_rn(__dmul_rn(__ddiv_rn((double)job.gt[1].hor, 2.0), ls), wx); cy[1] = __dmul_rn(__ddiv_rn((double)job.gt[1].ver, 2.0), ls);
      cx[2] = __dadd_rn(__dmul_rn(__ddiv_rn((double)job.gt[2].hor, 2.0), ls), wx); cy[2] = __dadd_rn(__dmul_rn(__ddiv_rn((double)job.gt[2].ver, 2.0), ls), hy);
      cx[3] = __dmul_rn(__ddiv_rn((double)job.gt[3].hor, 2.0), ls);                cy[3] = __dadd_rn(__dmul_rn(__ddiv_rn((double)job.gt[3].ver, 2.0), ls), hy);
    }
    calc_param(cx, cy, sh.h, bw * 2, bh * 2, chroma, xi, yi);
  }
  __syncthreads();

  // (2) interpolation (or copy) of the region / block
  int16_t* s_out = gt ? s_reg : s_pred;
  for (int i = threadIdx.x; i < rw * rh; i += blockDim.x) {
    const int y = i / rw, x = i - y * rw;
    const int16_t* s = s_src + (y + top) * sw + x + left;
    s_out[i] = (int16_t)(chroma ? mc_sample<4>(s, sw, fx, fy, job.bit_depth) : mc_sample<8>(s, sw, fx, fy, job.bit_depth));
  }
  __syncthreads();

  // (3) GT: warp the region into the block
  if (gt) {
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) {
      const int y = i / bw, x = i - y * bw;
      s_pred[i] = (int16_t)warp_literal(s_reg, rw, sh.h, bw, bh, x, y);
    }
    __syncthreads();
  }

  if (job.dst_off >= 0 && dst_buf)
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) {
      const int y = i / bw, x = i - y * bw;
      dst_buf[job.dst_off + (long long)y * job.dst_stride + x] = s_pred[i];
    }

  const int func = job.template_cost ? HOP_DF_SAD : job.dist_func;
  if (func) {
    const int16_t* org = org_buf + job.org_off;
    unsigned int sum = 0;
    if (func == HOP_DF_HADS) {
      int n = 0;
      if ((bh % 8 == 0) && (bw % 8 == 0)) n = 8;
      else if ((bh % 4 == 0) && (bw % 4 == 0)) n = 4;
      else if ((bh % 2 == 0) && (bw % 2 == 0)) n = 2;
      const int tx = n ? bw / n : 0, nt = n ? tx * (bh / n) : 0;
      for (int t = threadIdx.x; t < nt; t += blockDim.x) {
        const int x = (t % tx) * n, y = (t / tx) * n;
        const int16_t* o = org + y * job.org_stride + x;
        const int16_t* c = s_pred + y * bw + x;
        if (n == 8) sum += had_tile<8>(o, job.org_stride, c, bw);
        else if (n == 4) sum += had_tile<4>(o, job.org_stride, c, bw);
        else sum += had_tile<2>(o, job.org_stride, c, bw);
      }
    } else {
      for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) {
        const int y = i / bw, x = i - y * bw;
        sum = __sad((int)org[y * job.org_stride + x], (int)s_pred[i], sum);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd(&sh.dist, sum);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    HopPredResult r;
    r.valid = 1;
    r.dist = func ? sh.dist >> (job.bit_depth - 8) : 0;
    r.cost = r.dist;
    if (job.template_cost) {
      // calcRdCost(bits, dist, false, DF_SAD), TComRdCost.cpp:94-98
      const double lambda = (double)job.lambda_sad;
      const int rate = __double2int_rz(__dadd_rn(__dmul_rn((double)job.mvp_bits, lambda), .5)) >> 16;
      r.cost = __double2uint_rz(floor(__dadd_rn((double)r.dist, (double)rate)));
    }
    out[job_id] = r;
  }
}

cudaError_t predict_launch(int n, const HopPredJob* d_jobs, const int16_t* d_org, const int16_t* d_ref, int16_t* d_dst,
                           HopPredResult* d_out, int max_cols, int max_rows, bool any_gt, cudaStream_t stream, int* launches, RefBounds rb)
{
  static SmemOptIn opt_in;
  const size_t worst = k6_smem_bytes(HOP_MAX_PU, HOP_MAX_PU, 8, true);
  {
    cudaError_t e = opt_in.ensure(k6_predict, (int)worst);
    if (e != cudaSuccess) return e;
  }
  const size_t smem = k6_smem_bytes(max_cols, max_rows, 8, any_gt);
  k6_predict<<<n, 256, smem, stream>>>(n, d_jobs, d_org, d_ref, d_dst, d_out, rb);
  if (launches) (*launches)++;
  return cudaGetLastError();
}

}  // namespace hop
