// k2_gt.cu -- K2: the HOP / geometric-transform candidate search (sm_100a).
//
// Replaces TEncSearch::xPatternSearchGT, diamond branch (TLibEncoder/TEncSearch.cpp:4686-4790,
// 5093-5467) together with calcParamProjective / ProjectiveTransform
// (TLibCommon/TComPrediction.cpp:807-832, 904-1030), xGetHADs / xCalcHADs4x4/8x8 and the SAD family
// (TLibCommon/TComRdCost.cpp:513-1010, 1366-1708) and the bit-cost helpers.
//
// Parallel restatement:
//   * one CTA per PU (a cluster of CTAs for a multi-tile PU on the single-call path); start vectors and diamond
//     passes are walked sequentially inside the kernel (they are data dependent), the 56 affine corner sets of a
//     pass are evaluated in parallel;
//   * the 620 corner sets of a pass reduce to a fixed table of 56 offset patterns (in reference loop order)
//     because the pass centres always form a parallelogram; every candidate is still put through the
//     reference's own affine test, decided exactly in integers;
//   * a task = (candidate, Hadamard tile): two lanes warp one 8x8 tile (one lane a 4x4 tile; eight lanes, a row
//     each, on the single-call path) -- sample POSITIONS with the reference's binary64 operation sequence
//     (no FMA contraction, __d*_rn intrinsics), the bilinear VALUE by three lerps whose rounded result provably
//     equals the reference's (DESIGN.md "warp rounding") -- take the difference to the original block, run the
//     2-D Walsh-Hadamard butterflies in registers / shuffles and add the rounded tile SATD to the candidate's
//     accumulator in shared memory;
//   * ordered argmin: the minimum of (cost, loop index) over the pass, accepted iff it beats the running best,
//     which carries across passes and start vectors exactly as uiDistBest does.
#include "hop_common.cuh"
#include <cstdlib>
#include <cooperative_groups.h>
#include "hop_internal.h"
#include "k5_frac.cuh"

namespace hop {

namespace cg = cooperative_groups;

// x0,y0,x1,y1,x2,y2,x3,y3 in {-1,0,1}, loop order; global memory: every thread reads its own entry (a per-thread
// index into constant memory would be serialised by the constant cache)
__device__ __align__(8) int8_t d_gt_offsets[GT_CANDS][8];

void gt_build_offset_table(int8_t table[GT_CANDS][8], int* count)
{
  // the 8 nested loops of TEncSearch.cpp:5217-5287, each running (s, 0, -s); diamond points only;
  // translations skipped (:5289); affine <=> o0 - o1 + o2 - o3 == 0 (centres are parallelograms).
  int n = 0;
  const int v[3] = {1, 0, -1};
  for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { int y0 = v[a], x0 = v[b]; if (y0 && x0) continue;
  for (int c = 0; c < 3; c++) for (int d = 0; d < 3; d++) { int y1 = v[c], x1 = v[d]; if (y1 && x1) continue;
  for (int e = 0; e < 3; e++) for (int f = 0; f < 3; f++) { int y2 = v[e], x2 = v[f]; if (y2 && x2) continue;
  for (int g = 0; g < 3; g++) for (int h = 0; h < 3; h++) { int y3 = v[g], x3 = v[h]; if (y3 && x3) continue;
    if (x0 == x1 && x0 == x2 && x0 == x3 && y0 == y1 && y0 == y2 && y0 == y3) continue;
    if (x0 - x1 + x2 - x3 != 0 || y0 - y1 + y2 - y3 != 0) continue;
    if (n < GT_CANDS) {
      int8_t* t = table[n];
      t[0] = x0; t[1] = y0; t[2] = x1; t[3] = y1; t[4] = x2; t[5] = y2; t[6] = x3; t[7] = y3;
    }
    n++;
  }}}}
  *count = n;
}

cudaError_t gt_upload_offset_table(const int8_t table[GT_CANDS][8])
{
  return cudaMemcpyToSymbol(d_gt_offsets, table, sizeof(int8_t) * GT_CANDS * 8);
}

#ifdef HOP_TRACE
__device__ unsigned long long g_trace_k2[HOP_TRACE_SLOTS];
void trace_read_k2(unsigned long long* out) { cudaMemcpyFromSymbol(out, g_trace_k2, sizeof(g_trace_k2)); }
#endif

struct GtShared {
  // per-pass candidate table (SoA, loop order)
  double   h0[GT_CANDS], h3[GT_CANDS], h6[GT_CANDS];   // Fx = h0*x + h3*y + h6
  double   h1[GT_CANDS], h4[GT_CANDS], h7[GT_CANDS];   // Fy = h1*x + h4*y + h7
  // what the argmin reads is double buffered by pass parity: the set-up of the next pass (by the thread that
  // owns the candidate) may overtake another warp that is still reducing the current one
  uint32_t add_cost[2][GT_CANDS];                      // getCost(Hor,Ver) + getCost(getBitsGT(..))
  uint32_t dist[2][GT_CANDS];                          // sum of tile SATDs / SADs
  int32_t  valid[2][GT_CANDS];
  int16_t  corner[2][GT_CANDS][8];
  // search state, one replica per head warp (identical contents): written by lane 0 after the argmin of a
  // pass, read by the warp's own lanes when they set up the next pass -- registers stay free for the tiles
  struct State {
    int4     cc;                                       // iBestNSSCenterX/Y, packed (y << 16) | (x & 0xffff) like corner[]
    int4     bc;                                       // iBestCornerX/Y (:4776)
    uint32_t dist_best;
    int32_t  best_ss_x, best_ss_y, best_index;
    uint32_t n_cand;
  } st[2];
  int2     offs[GT_CANDS];                             // the 56 corner-offset patterns (8 bytes each)
  unsigned long long red_key[2];                       // sweep kernel: cross-warp argmin
  uint32_t red_cnt[2];
};
constexpr size_t GT_SHARED_BYTES = (sizeof(GtShared) + 15) / 16 * 16;

// Quotient tables of calcParamProjective: h0 = fl(n / (W-1)) and h3 = fl(n / (H-1)) (W, H on the 2x grid) for every
// numerator a corner set can produce -- corners stay within +-nss of the initial rectangle, so
// n in [-2 nss, W - 1 + 2 nss].  Filled once per PU with the same IEEE division the reference executes, so that
// the per-pass candidate set-up is four table reads instead of four divisions (0.45 us per pass on the latency path).
__host__ __device__ inline int gt_div_entries(int dim2x, int nss) { return dim2x + 4 * nss + 1; }
__host__ __device__ inline size_t gt_div_bytes(int cols, int rows)
{
  const int nss = ((rows < cols ? rows : cols) >> 1) * 2;
  return (sizeof(double) * (size_t)(gt_div_entries(2 * cols, nss) + gt_div_entries(2 * rows, nss)) + 15) & ~(size_t)15;
}

// ---- in-register Walsh-Hadamard tiles ---------------------------------------------------------
template <int N>
struct Tile {
  int d[N * N];
  __device__ __forceinline__ void row_transform(int r)
  {
    int* p = d + r * N;
#pragma unroll
    for (int len = 1; len < N; len <<= 1)
#pragma unroll
      for (int i = 0; i < N; i += len << 1)
#pragma unroll
        for (int j = i; j < i + len; j++) { int a = p[j], b = p[j + len]; p[j] = a + b; p[j + len] = a - b; }
  }
  // column butterflies + sum of magnitudes; the last butterfly stage is folded into the magnitude sum:
  // |a+b| + |a-b| == 2*max(|a|,|b|)
  __device__ __forceinline__ uint32_t satd()
  {
#pragma unroll
    for (int x = 0; x < N; x++)
#pragma unroll
      for (int len = 1; len < N / 2; len <<= 1)
#pragma unroll
        for (int i = 0; i < N; i += len << 1)
#pragma unroll
          for (int j = i; j < i + len; j++) {
            int a = d[j * N + x], b = d[(j + len) * N + x];
            d[j * N + x] = a + b; d[(j + len) * N + x] = a - b;
          }
    unsigned s = 0;
#pragma unroll
    for (int x = 0; x < N; x++)
#pragma unroll
      for (int j = 0; j < N / 2; j++) s += (unsigned)max(abs(d[j * N + x]), abs(d[(j + N / 2) * N + x]));
    s <<= 1;
    return N == 8 ? (s + 2) >> 2 : (s + 1) >> 1;   // xCalcHADs8x8 :1572 / xCalcHADs4x4 :1476
  }
  __device__ __forceinline__ uint32_t sad()
  {
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < N * N; k++) s = __sad(d[k], 0, s);
    return s;
  }
};

// exact (double)n for 0 <= n < 2^31 without a conversion instruction: 2^52 + n is representable
__device__ __forceinline__ double small_int_to_double(int n)
{
  return __dsub_rn(__hiloint2double(0x43300000, n), 4503599627370496.0);
}

// Per-thread constants of the sampling loop.  wox/woy go through an opaque move so that they live in ordinary
// registers: the fused add-min-max instruction takes its addend from a register, and a value the compiler knows
// to be CTA-uniform would be re-materialised from the uniform register file for every pixel.
struct WarpCtx {
  unsigned win_sa;          // shared-memory address of the window
  int wox, woy;             // w - off_x, w - off_y: window coordinate of X = trunc(Fx) - off_x is trunc(Fx) + wox
  int lim_xw, lim_yw;       // (w + cols - 1) - 1 + w, likewise for rows: the upper clamp in window coordinates
  int off_x, off_y;
  __device__ __forceinline__ WarpCtx(const uint32_t* win, int w, int cols, int rows, int ox, int oy)
  {
    win_sa = (unsigned)__cvta_generic_to_shared(win);
    asm("mov.b32 %0, %1;" : "=r"(wox) : "r"(w - ox));
    asm("mov.b32 %0, %1;" : "=r"(woy) : "r"(w - oy));
    lim_xw = 2 * w + cols - 2; lim_yw = 2 * w + rows - 2;
    off_x = ox; off_y = oy;
  }
};

// One warped sample: ProjectiveTransform body for the affine case (h2 == h5 == 0 => denominator is
// exactly 1.0, so the reference's division returns its numerator unchanged), TComPrediction.cpp:925-972,1025.
//   win    : the clamped window samples as the HIGH 32-bit word of their binary64 value (integers < 2^20
//            have a zero low word), so reading a sample costs no int->double conversion;
//            win[Yc * WS + Xc] with Yc = Y + w, Xc = X + w
//
// What is bit-for-bit the reference's: Fx, Fy (computed by the callers with the reference's operation order)
// and their C truncation -- the sample POSITION is a discontinuous function of them.  What is NOT replayed
// operation by operation is the interpolation weight and the bilinear blend, because only the rounded 8-bit
// result is observable and it cannot depend on the last bits ("warp rounding"):
//  (1) p, q are rationals with the ODD denominator D = (2*cols-1)(2*rows-1) <= 127^2, so the exact blend of
//      four integers has denominator D^2 and stays >= 1/(2*D^2) >= 1.9e-9 away from every half-integer;
//  (2) the reference's own binary64 evaluation and the one below both stay within 2e-10 of that exact value
//      (weights within 8e-14 of the exact rationals, samples <= 1023, a handful of roundings at 2^-44);
//  (3) hence round-to-nearest of either is the same integer, and the 0/255 clip commutes with the rounding.
//  The reference's literal sequence is kept under HOP_WARP_REFERENCE_OPS for A/B parity runs.
// DXP = true (batched kernels, packed tile pairs): every window row is followed by a row of horizontal differences
// dx[X] = sample[X + 1] - sample[X] (exact small integers, stored like the samples as the high word of their binary64
// value), so that the horizontal lerps are one fused multiply-add each: top = A + p * dxA -- the SAME values as
// A + p * (B - A), since B - A is computed exactly either way -- two additions per warped pixel less.
template <int WS, bool DXP = false>
__device__ __forceinline__ int warp_sample(const WarpCtx& wc, double Fx, double Fy)
{
  constexpr int RS = DXP ? 2 * WS + 1 : WS;      // row stride of the window in words: odd (== 7 mod 32) either way
  const int wox = wc.wox, woy = wc.woy, lim_xw = wc.lim_xw, lim_yw = wc.lim_yw;
  const int Yt = __double2int_rz(Fy);            // C truncation toward zero
  const int Xt = __double2int_rz(Fx);
  // the six ordered clamps of :950-961 collapse to [-w, lim-1]: after the first four the value is in
  // [-w, lim]; the last two map lim to lim-1.  In window coordinates: max(min(Y + w, lim_yw), 0).
  const int Yc = __vimin_s32_relu(Yt + woy, lim_yw);
  const int Xc = __vimin_s32_relu(Xt + wox, lim_xw);
  // one IMAD + one LEA: the window's 32-bit shared-memory address is kept as an integer so that the word
  // offset of the window inside the CTA's dynamic shared memory is folded into the base once per pass
  const unsigned a = wc.win_sa + 4u * (unsigned)(Yc * RS + Xc);
  unsigned hA, hB, hC, hD;                        // DXP: hB, hD are the differences dxA, dxC
  asm("ld.shared.u32 %0, [%1];" : "=r"(hA) : "r"(a));
  asm("ld.shared.u32 %0, [%1+%2];" : "=r"(hB) : "r"(a), "n"(DXP ? WS * 4 : 4));
  asm("ld.shared.u32 %0, [%1+%2];" : "=r"(hC) : "r"(a), "n"(RS * 4));
  asm("ld.shared.u32 %0, [%1+%2];" : "=r"(hD) : "r"(a), "n"(DXP ? RS * 4 + WS * 4 : RS * 4 + 4));
  const double A = __hiloint2double((int)hA, 0), C = __hiloint2double((int)hC, 0);
#ifdef HOP_WARP_REFERENCE_OPS
  const double B = DXP ? __dadd_rn(A, __hiloint2double((int)hB, 0)) : __hiloint2double((int)hB, 0);   // exact
  const double D = DXP ? __dadd_rn(C, __hiloint2double((int)hD, 0)) : __hiloint2double((int)hD, 0);
#else
  const double dAB = DXP ? __hiloint2double((int)hB, 0) : __dsub_rn(__hiloint2double((int)hB, 0), A);
  const double dCD = DXP ? __hiloint2double((int)hD, 0) : __dsub_rn(__hiloint2double((int)hD, 0), C);
#endif
  // (double)Yt / (double)Xt stay I2F conversions: the kernel is issue bound, and the conversion-free form
  // (2^52 magic: LOP + MOV + DADD) measured 4 % slower (profiles/r01_k2_experiments.txt)
#ifdef HOP_WARP_REFERENCE_OPS
  const double off_xd = small_int_to_double(wc.off_x), off_yd = small_int_to_double(wc.off_y);
  const double q = __dsub_rn(__dsub_rn(Fy, off_yd), (double)(Yt - wc.off_y));
  const double p = __dsub_rn(__dsub_rn(Fx, off_xd), (double)(Xt - wc.off_x));
  const double omp = __dsub_rn(1.0, p), omq = __dsub_rn(1.0, q);
  double aux = __dmul_rn(omq, __dadd_rn(__dmul_rn(omp, A), __dmul_rn(p, B)));
  aux = __dadd_rn(aux, __dmul_rn(q, __dadd_rn(__dmul_rn(omp, C), __dmul_rn(p, D))));
#else
  // fractional parts: Fx - trunc(Fx) is exact (Sterbenz); the reference's (Fx - off) - X differs from it by
  // the rounding of Fx - off only (<= 2^-46)
  const double q = __dsub_rn(Fy, (double)Yt);
  const double p = __dsub_rn(Fx, (double)Xt);
  // three lerps: top = A + p(B-A), bot = C + p(D-C), aux = top + q(bot-top)   (6 fp64 ops instead of 11)
  const double top = __fma_rn(p, dAB, A);
  const double bot = __fma_rn(p, dCD, C);
  const double aux = __fma_rn(q, __dsub_rn(bot, top), top);
#endif
  // Adding 1.5*2^52 leaves round-to-nearest(aux) in the low word: no double->int conversion.
  const int v = __double2loint(__dadd_rn(aux, 6755399441055744.0));
  return __vimin_s32_relu(v, 255);
}

// ---- 4x4 tiles (PU shapes with a dimension of 4 or 12): one thread per tile -------------------
template <int WS, bool HAD>
__device__ __forceinline__ uint32_t eval_tile4(double h0, double h3, double h6, double h1, double h4, double h7,
                                               int tx, int ty, const int* __restrict__ org,
                                               const WarpCtx& wc, int cols)
{
  constexpr int N = 4;   // off_x/off_y: W/2 - W/4 with W = 2*cols on the 2x grid, 0 on the 1x grid (sweep)
  const int off_x = wc.off_x, off_y = wc.off_y;
  double h0x[N], h1x[N];
#pragma unroll
  for (int k = 0; k < N; k++) {
    const double xd = small_int_to_double(off_x + tx + k);
    h0x[k] = __dmul_rn(h0, xd);
    h1x[k] = __dmul_rn(h1, xd);
  }
  Tile<N> t;
#pragma unroll
  for (int r = 0; r < N; r++) {
    const double yd = small_int_to_double(off_y + ty + r);
    const double h3y = __dmul_rn(h3, yd), h4y = __dmul_rn(h4, yd);
    const int4 a = *reinterpret_cast<const int4*>(org + (ty + r) * cols + tx);
    const int o[N] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int k = 0; k < N; k++) {
      const double Fx = __dadd_rn(__dadd_rn(h0x[k], h3y), h6);    // (h0*x + h3*y) + h6, left to right
      const double Fy = __dadd_rn(__dadd_rn(h1x[k], h4y), h7);
      t.d[r * N + k] = o[k] - warp_sample<WS>(wc, Fx, Fy);
    }
    if (HAD) t.row_transform(r);
  }
  return HAD ? t.satd() : t.sad();
}

// ---- 8x8 tiles: two adjacent lanes per tile, 4 rows x 8 columns each ----------------------------
// Keeps the unrolled body at 32 pixels (instruction-cache resident) and the register tile at 32 words.
// Lane `half` owns rows 4*half .. 4*half+3: horizontal 8-point and vertical 4-point butterflies run in
// registers; the last vertical stage pairs coefficient (j,k) of the two halves and is folded into the
// magnitude sum: |a+b| + |a-b| == 2*max(|a|,|b|).  Each lane handles 16 of the 32 coefficient pairs
// after one shuffle per pair.  Returns the rounded tile SATD in both lanes (or the half-tile SAD).
template <int WS, bool HAD>
__device__ __forceinline__ uint32_t eval_half_tile8(double h0, double h3, double h6, double h1, double h4, double h7,
                                                    int tx, int ty, int half, const int* __restrict__ org,
                                                    const WarpCtx& wc, int cols)
{
  const int off_x = wc.off_x, off_y = wc.off_y;
  const int y0 = ty + 4 * half;
  double h3y[4], h4y[4];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const double yd = small_int_to_double(off_y + y0 + r);
    h3y[r] = __dmul_rn(h3, yd);
    h4y[r] = __dmul_rn(h4, yd);
  }
  int d[32];
#pragma unroll
  for (int r = 0; r < 4; r++) {                    // original block: two 16-byte loads per row
    const int4 a = *reinterpret_cast<const int4*>(org + (y0 + r) * cols + tx);
    const int4 b = *reinterpret_cast<const int4*>(org + (y0 + r) * cols + tx + 4);
    d[r * 8 + 0] = a.x; d[r * 8 + 1] = a.y; d[r * 8 + 2] = a.z; d[r * 8 + 3] = a.w;
    d[r * 8 + 4] = b.x; d[r * 8 + 5] = b.y; d[r * 8 + 6] = b.z; d[r * 8 + 7] = b.w;
  }
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const double xd = small_int_to_double(off_x + tx + k);
    const double h0x = __dmul_rn(h0, xd), h1x = __dmul_rn(h1, xd);
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const double Fx = __dadd_rn(__dadd_rn(h0x, h3y[r]), h6);    // (h0*x + h3*y) + h6, left to right
      const double Fy = __dadd_rn(__dadd_rn(h1x, h4y[r]), h7);
      d[r * 8 + k] -= warp_sample<WS>(wc, Fx, Fy);
    }
  }
  if (!HAD) {
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) s = __sad(d[i], 0, s);
    return s;
  }
#pragma unroll
  for (int r = 0; r < 4; r++)                      // horizontal 8-point
#pragma unroll
    for (int len = 1; len < 8; len <<= 1)
#pragma unroll
      for (int i = 0; i < 8; i += len << 1)
#pragma unroll
        for (int j = i; j < i + len; j++) {
          const int a = d[r * 8 + j], b = d[r * 8 + j + len];
          d[r * 8 + j] = a + b; d[r * 8 + j + len] = a - b;
        }
#pragma unroll
  for (int k = 0; k < 8; k++)                      // vertical 4-point inside the half
#pragma unroll
    for (int len = 1; len < 4; len <<= 1)
#pragma unroll
      for (int i = 0; i < 4; i += len << 1)
#pragma unroll
        for (int j = i; j < i + len; j++) {
          const int a = d[j * 8 + k], b = d[(j + len) * 8 + k];
          d[j * 8 + k] = a + b; d[(j + len) * 8 + k] = a - b;
        }
  // only the two lanes of the pair take part: other pairs of the warp may run a different number of
  // tiles (or none, for a rejected candidate), so a full-warp mask would wait for lanes that left
  const unsigned pair_mask = 3u << (threadIdx.x & 30u);
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int mine = half ? d[16 + i] : d[i];
    const int send = half ? d[i] : d[16 + i];
    const int recv = __shfl_xor_sync(pair_mask, send, 1);
    s += (unsigned)max(abs(mine), abs(recv));
  }
  s += __shfl_xor_sync(pair_mask, s, 1);
  return (s + 1) >> 1;                             // (2*sum + 2) >> 2, xCalcHADs8x8 TComRdCost.cpp:1572
}

// ---- 8x8 tiles, two tiles per register tile (batched kernel) ------------------------------------
// The Hadamard butterflies are plain integer adds, and every intermediate of an 8x8 SATD of <= 10-bit residuals
// stays below 2^15 (|d| <= 1023, x8 horizontally, x4 inside the half): two tiles of the SAME candidate therefore
// share one register tile, one per 16-bit lane.  Lanes hold value + 0x8000 (unsigned, never 0), so that a 32-bit
//   a + b - 0x80008000   /   a - b + 0x80008000
// is the lane-wise sum / difference with no carry or borrow between the lanes (each lane result stays inside
// [1, 65535]) -- one IADD3 per butterfly output for TWO tiles.  The original block is staged already packed
// (tile A in the low lane, its partner B -- half a block further down, or to the right -- in the high lane), the
// warped sample of tile A is subtracted with multiplier 1, that of tile B with multiplier 65536 (one IMAD each,
// as before); |x| = max(lane, 0x10000 - lane) - 0x8000 on the packed word (VIMNMX3.U16x2), the two per-tile sums
// leave the packed domain through IDP.2A.  Same integers as the one-tile form, ~4 instructions per warped pixel less.
constexpr unsigned PK_BIAS = 0x80008000u;
__device__ __forceinline__ unsigned pk_add(unsigned a, unsigned b) { return a + b - PK_BIAS; }
__device__ __forceinline__ unsigned pk_sub(unsigned a, unsigned b) { return a - b + PK_BIAS; }

template <int WS, bool DXP>
__device__ __forceinline__ uint32_t eval_half_tile8_pair(double h0, double h3, double h6, double h1, double h4, double h7,
                                                         int txA, int tyA, int dxB, int dyB, int half,
                                                         const unsigned* __restrict__ orgp, int pw, const WarpCtx& wc)
{
  const int off_x = wc.off_x, off_y = wc.off_y;
  unsigned d[32];
#pragma unroll
  for (int r = 0; r < 4; r++) {                    // packed original block: two 16-byte loads per row, both tiles
    const uint4 a = *reinterpret_cast<const uint4*>(orgp + (tyA + 4 * half + r) * pw + txA);
    const uint4 b = *reinterpret_cast<const uint4*>(orgp + (tyA + 4 * half + r) * pw + txA + 4);
    d[r * 8 + 0] = a.x; d[r * 8 + 1] = a.y; d[r * 8 + 2] = a.z; d[r * 8 + 3] = a.w;
    d[r * 8 + 4] = b.x; d[r * 8 + 5] = b.y; d[r * 8 + 6] = b.z; d[r * 8 + 7] = b.w;
  }
#pragma unroll 1
  for (int t = 0; t < 2; t++) {                    // tile A, then tile B: the same 32-pixel body (instruction cache)
    const int tx = txA + (t ? dxB : 0), y0 = tyA + (t ? dyB : 0) + 4 * half;
    int negmul;
    asm("mov.b32 %0, %1;" : "=r"(negmul) : "r"(t ? -65536 : -1));   // opaque: keeps one IMAD per pixel for both tiles
    double h3y[4], h4y[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const double yd = small_int_to_double(off_y + y0 + r);
      h3y[r] = __dmul_rn(h3, yd);
      h4y[r] = __dmul_rn(h4, yd);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const double xd = small_int_to_double(off_x + tx + k);
      const double h0x = __dmul_rn(h0, xd), h1x = __dmul_rn(h1, xd);
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const double Fx = __dadd_rn(__dadd_rn(h0x, h3y[r]), h6);    // (h0*x + h3*y) + h6, left to right
        const double Fy = __dadd_rn(__dadd_rn(h1x, h4y[r]), h7);
        d[r * 8 + k] += (unsigned)(warp_sample<WS, DXP>(wc, Fx, Fy) * negmul);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; r++)                      // horizontal 8-point, both tiles at once
#pragma unroll
    for (int len = 1; len < 8; len <<= 1)
#pragma unroll
      for (int i = 0; i < 8; i += len << 1)
#pragma unroll
        for (int j = i; j < i + len; j++) {
          const unsigned a = d[r * 8 + j], b = d[r * 8 + j + len];
          d[r * 8 + j] = pk_add(a, b); d[r * 8 + j + len] = pk_sub(a, b);
        }
#pragma unroll
  for (int k = 0; k < 8; k++)                      // vertical 4-point inside the half
#pragma unroll
    for (int len = 1; len < 4; len <<= 1)
#pragma unroll
      for (int i = 0; i < 4; i += len << 1)
#pragma unroll
        for (int j = i; j < i + len; j++) {
          const unsigned a = d[j * 8 + k], b = d[(j + len) * 8 + k];
          d[j * 8 + k] = pk_add(a, b); d[(j + len) * 8 + k] = pk_sub(a, b);
        }
  const unsigned pair_mask = 3u << (threadIdx.x & 30u);
  // lane sums start at -16 * 0x8000: every accumulated maximum carries the lane bias once
  unsigned sA = 0u - 16u * 0x8000u, sB = 0u - 16u * 0x8000u;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const unsigned mine = half ? d[16 + i] : d[i];
    const unsigned send = half ? d[i] : d[16 + i];
    const unsigned recv = __shfl_xor_sync(pair_mask, send, 1);
    // max(|a|, |b|) + 0x8000 per lane: a lane and its negation 0x10000 - lane (no borrow: lanes are never 0)
    unsigned m = __vimax3_u16x2(mine, 0x00010000u - mine, recv);
    m = __vmaxu2(m, 0x00010000u - recv);
    sA = __dp2a_lo(m, 0x00000001u, sA);
    sB = __dp2a_lo(m, 0x00000100u, sB);
  }
  sA += __shfl_xor_sync(pair_mask, sA, 1);
  sB += __shfl_xor_sync(pair_mask, sB, 1);
  return ((sA + 1) >> 1) + ((sB + 1) >> 1);        // per tile (2*sum + 2) >> 2, xCalcHADs8x8 TComRdCost.cpp:1572
}

// ---- task loops -----------------------------------------------------------------------------------
// dynamic shared memory: [GtShared][org rows*cols int32][window (rows+2w) x WS uint32 high words]
// N == 4: blockDim.x = 56 * groups, thread -> (c = tid % 56, g = tid / 56), tiles g, g+groups, ...
// N == 8: blockDim.x = 112 * groups, thread -> (half = tid & 1, c = (tid >> 1) % 56, g = (tid >> 1) / 56)
// A thread keeps its candidate (and the six map coefficients) for the whole pass.
template <int WS, bool HAD>
__device__ __forceinline__ void run_tasks4(GtShared& sh, const int* s_org, const uint32_t* s_win,
                                           int w, int cols, int rows, int off_x, int off_y, int crank, int csize, int pb)
{
  const int tiles_x = cols / 4, ntiles = tiles_x * (rows / 4);
  // tile groups of this CTA: as many as it has lanes for, but no more than its share of the tiles (a cluster CTA
  // may be launched wider than its tiles need -- the extra warps serve the staging and the fractional stage)
  const int per_cta = (ntiles + csize - 1) / csize, avail = blockDim.x / GT_CANDS;
  const int c = threadIdx.x % GT_CANDS, g = threadIdx.x / GT_CANDS, groups = avail < per_cta ? avail : per_cta;
  if (g >= groups || !sh.valid[pb][c]) return;
  const double h0 = sh.h0[c], h3 = sh.h3[c], h6 = sh.h6[c];
  const double h1 = sh.h1[c], h4 = sh.h4[c], h7 = sh.h7[c];
  const WarpCtx wc(s_win, w, cols, rows, off_x, off_y);
  uint32_t acc = 0;
  for (int tile = g + groups * crank; tile < ntiles; tile += groups * csize) {   // cluster: CTAs interleave tiles
    const int tx = (tile % tiles_x) * 4, ty = (tile / tiles_x) * 4;
    acc += eval_tile4<WS, HAD>(h0, h3, h6, h1, h4, h7, tx, ty, s_org, wc, cols);
  }
  atomicAdd(&sh.dist[pb][c], acc);
}

template <int WS, bool HAD>
__device__ __forceinline__ void run_tasks8(GtShared& sh, const int* s_org, const uint32_t* s_win,
                                           int w, int cols, int rows, int off_x, int off_y, int crank, int csize, int pb)
{
  const int half = threadIdx.x & 1, pair = threadIdx.x >> 1;
  const int tiles_x = cols / 8, ntiles = tiles_x * (rows / 8);
  const int per_cta = (ntiles + csize - 1) / csize, avail = blockDim.x / (2 * GT_CANDS);
  const int c = pair % GT_CANDS, g = pair / GT_CANDS, groups = avail < per_cta ? avail : per_cta;
  // lanes of a pair always agree on (c, g), so the shuffles inside eval_half_tile8 see both lanes
  if (g >= groups || !sh.valid[pb][c]) return;
  const double h0 = sh.h0[c], h3 = sh.h3[c], h6 = sh.h6[c];
  const double h1 = sh.h1[c], h4 = sh.h4[c], h7 = sh.h7[c];
  const WarpCtx wc(s_win, w, cols, rows, off_x, off_y);
  uint32_t acc = 0;
  for (int tile = g + groups * crank; tile < ntiles; tile += groups * csize) {   // cluster: CTAs interleave tiles
    const int tx = (tile % tiles_x) * 8, ty = (tile / tiles_x) * 8;
    acc += eval_half_tile8<WS, HAD>(h0, h3, h6, h1, h4, h7, tx, ty, half, s_org, wc, cols);
  }
  if (!HAD || half == 0) atomicAdd(&sh.dist[pb][c], acc);   // HAD: both lanes hold the tile sums, count once
}

// Pairing of the 8x8 tiles of a PU for the packed form: tile (tx, ty) with the tile half a block further down
// (rows a multiple of 16), else with the tile half a block to the right (cols a multiple of 16).  Every PU shape
// with 8x8 tiles except 8x8 itself pairs.  The packed original block is indexed by tile A's coordinates.
struct PairGeom { int on, pw, ph, dxB, dyB; };
__host__ __device__ inline PairGeom gt_pair_geom(int cols, int rows)
{
  PairGeom g = {0, cols, rows, 0, 0};
  if ((cols & 7) || (rows & 7)) return g;
  if ((rows & 15) == 0)      { g.on = 1; g.ph = rows >> 1; g.dyB = rows >> 1; }
  else if ((cols & 15) == 0) { g.on = 1; g.pw = cols >> 1; g.dxB = cols >> 1; }
  return g;
}

template <int WS, bool DXP>
__device__ __forceinline__ void run_tasks8_pair(GtShared& sh, const unsigned* s_orgp, const uint32_t* s_win, const PairGeom pg,
                                                int w, int cols, int rows, int off_x, int off_y, int pb)
{
  const int half = threadIdx.x & 1, pair = threadIdx.x >> 1;
  const int tiles_x = pg.pw / 8, npairs = tiles_x * (pg.ph / 8);
  const int avail = blockDim.x / (2 * GT_CANDS);
  const int c = pair % GT_CANDS, g = pair / GT_CANDS, groups = avail < npairs ? avail : npairs;
  if (g >= groups || !sh.valid[pb][c]) return;
  const double h0 = sh.h0[c], h3 = sh.h3[c], h6 = sh.h6[c];
  const double h1 = sh.h1[c], h4 = sh.h4[c], h7 = sh.h7[c];
  const WarpCtx wc(s_win, w, cols, rows, off_x, off_y);
  uint32_t acc = 0;
  for (int pr = g; pr < npairs; pr += groups) {
    const int tx = (pr % tiles_x) * 8, ty = (pr / tiles_x) * 8;
    acc += eval_half_tile8_pair<WS, DXP>(h0, h3, h6, h1, h4, h7, tx, ty, pg.dxB, pg.dyB, half, s_orgp, pg.pw, wc);
  }
  if (half == 0) atomicAdd(&sh.dist[pb][c], acc);   // both lanes hold the tile sums, count once
}

// ---- latency form: one lane per tile ROW ----------------------------------------------------------
// A single PU (the in-encoder call) cannot fill the machine, so the length of the dependent chain decides: N
// adjacent lanes share a tile, each warps one row of N pixels (a quarter of the chain of the half-tile form),
// runs the horizontal butterflies in registers and meets the other rows through log2(N) shuffle stages; the
// last stage is folded into the magnitude sum as above.  ~20 % more instructions per pixel than the register
// tiles, which is why the batched (throughput) launches keep those.
template <int WS, bool HAD, int N>
__device__ __forceinline__ uint32_t eval_tile_row(double h0, double h3, double h6, double h1, double h4, double h7,
                                                  int tx, int ty, int j, unsigned group_mask, const int* __restrict__ org,
                                                  const WarpCtx& wc, int cols)
{
  const int off_x = wc.off_x, off_y = wc.off_y;
  const int y = ty + j;
  const double yd = small_int_to_double(off_y + y);
  const double h3y = __dmul_rn(h3, yd), h4y = __dmul_rn(h4, yd);
  int d[N];
#pragma unroll
  for (int k = 0; k < N; k += 4) {
    const int4 a = *reinterpret_cast<const int4*>(org + y * cols + tx + k);
    d[k] = a.x; d[k + 1] = a.y; d[k + 2] = a.z; d[k + 3] = a.w;
  }
#pragma unroll
  for (int k = 0; k < N; k++) {
    const double xd = small_int_to_double(off_x + tx + k);
    const double Fx = __dadd_rn(__dadd_rn(__dmul_rn(h0, xd), h3y), h6);    // (h0*x + h3*y) + h6, left to right
    const double Fy = __dadd_rn(__dadd_rn(__dmul_rn(h1, xd), h4y), h7);
    d[k] -= warp_sample<WS>(wc, Fx, Fy);
  }
  unsigned s = 0;
  if (!HAD) {
#pragma unroll
    for (int k = 0; k < N; k++) s = __sad(d[k], 0, s);
    return s;                                        // per-lane share of the tile SAD
  }
#pragma unroll
  for (int len = 1; len < N; len <<= 1)              // horizontal N-point
#pragma unroll
    for (int i = 0; i < N; i += len << 1)
#pragma unroll
      for (int q = i; q < i + len; q++) { const int a = d[q], b = d[q + len]; d[q] = a + b; d[q + len] = a - b; }
#pragma unroll
  for (int m = 1; m < N / 2; m <<= 1)                // vertical stages across the lanes, all but the last
#pragma unroll
    for (int k = 0; k < N; k++) {
      const int recv = __shfl_xor_sync(group_mask, d[k], m);
      d[k] = (j & m) ? recv - d[k] : d[k] + recv;
    }
#pragma unroll
  for (int k = 0; k < N; k++) {                      // last stage: |a+b| + |a-b| == 2*max(|a|,|b|), seen from both lanes
    const int recv = __shfl_xor_sync(group_mask, d[k], N / 2);
    s += (unsigned)max(abs(d[k]), abs(recv));
  }
#pragma unroll
  for (int m = 1; m < N; m <<= 1) s += __shfl_xor_sync(group_mask, s, m);   // sum over the N lanes == sum |T|
  return N == 8 ? (s + 2) >> 2 : (s + 1) >> 1;       // xCalcHADs8x8 :1572 / xCalcHADs4x4 :1476
}

template <int WS, bool HAD, int N>
__device__ __forceinline__ void run_tasks_rows(GtShared& sh, const int* s_org, const uint32_t* s_win,
                                               int w, int cols, int rows, int off_x, int off_y, int crank, int csize, int pb)
{
  const int tiles_x = cols / N, ntiles = tiles_x * (rows / N);
  const int per_cta = (ntiles + csize - 1) / csize, avail = blockDim.x / (N * GT_CANDS);
  const int j = threadIdx.x & (N - 1), slot = threadIdx.x / N;
  const int c = slot % GT_CANDS, g = slot / GT_CANDS, groups = avail < per_cta ? avail : per_cta;
  // the N lanes of a tile agree on (c, g): they leave, loop and shuffle together
  if (g >= groups || !sh.valid[pb][c]) return;
  const unsigned group_mask = ((1u << N) - 1u) << (threadIdx.x & 31 & ~(N - 1));
  const double h0 = sh.h0[c], h3 = sh.h3[c], h6 = sh.h6[c];
  const double h1 = sh.h1[c], h4 = sh.h4[c], h7 = sh.h7[c];
  const WarpCtx wc(s_win, w, cols, rows, off_x, off_y);
  uint32_t acc = 0;
  for (int tile = g + groups * crank; tile < ntiles; tile += groups * csize) {
    const int tx = (tile % tiles_x) * N, ty = (tile / tiles_x) * N;
    acc += eval_tile_row<WS, HAD, N>(h0, h3, h6, h1, h4, h7, tx, ty, j, group_mask, s_org, wc, cols);
  }
  if (!HAD || j == 0) atomicAdd(&sh.dist[pb][c], acc);   // HAD: every lane holds the tile sums, count once
}

template <int WS>
__device__ __forceinline__ void run_tasks(GtShared& sh, const int* s_org, const uint32_t* s_win, int w,
                                          int cols, int rows, int off_x, int off_y, int tile_n, int use_had,
                                          int crank = 0, int csize = 1, int pb = 0, bool fine = false)
{
  if (fine) {
    if (tile_n == 8) {
      if (use_had) run_tasks_rows<WS, true, 8>(sh, s_org, s_win, w, cols, rows, off_x, off_y, crank, csize, pb);
      else         run_tasks_rows<WS, false, 8>(sh, s_org, s_win, w, cols, rows, off_x, off_y, crank, csize, pb);
    } else {
      if (use_had) run_tasks_rows<WS, true, 4>(sh, s_org, s_win, w, cols, rows, off_x, off_y, crank, csize, pb);
      else         run_tasks_rows<WS, false, 4>(sh, s_org, s_win, w, cols, rows, off_x, off_y, crank, csize, pb);
    }
    return;
  }
  if (tile_n == 8) {
    if (use_had) run_tasks8<WS, true>(sh, s_org, s_win, w, cols, rows, off_x, off_y, crank, csize, pb);
    else         run_tasks8<WS, false>(sh, s_org, s_win, w, cols, rows, off_x, off_y, crank, csize, pb);
  } else {
    if (use_had) run_tasks4<WS, true>(sh, s_org, s_win, w, cols, rows, off_x, off_y, crank, csize, pb);
    else         run_tasks4<WS, false>(sh, s_org, s_win, w, cols, rows, off_x, off_y, crank, csize, pb);
  }
}

// Window row stride classes (in 32-bit words), compile-time so that the 2x2 footprint is one address
// plus immediates.  The 32 lanes of a warp are 16 neighbouring candidates x 2 half tiles reading the SAME
// pixel of their tile: their window positions differ by the candidates' corner offsets, i.e. by small
// (dx, dy) with |dx|, |dy| of the same size.  WS == 3 (mod 32) maps (X, Y) to bank (X + 3Y) mod 32, which
// separates (dx, dy) from (dx +- 1, dy -+ 1): measured 4.0 shared-memory wavefronts per request with
// WS == 1 (mod 32), 2.1 simulated with 3 (profiles/r01_k2_experiments.txt).  win_w = cols + min(cols, rows) <= 128.
constexpr int WS_A = 35, WS_B = 67, WS_C = 99, WS_D = 131;
__host__ __device__ constexpr int gt_stride_class(int win_w)
{
  return win_w <= 32 ? WS_A : win_w <= 64 ? WS_B : win_w <= 96 ? WS_C : WS_D;
}
constexpr int gt_class_threads(int) { return GT_THREADS; }

// launch configurations (threads per CTA, min CTAs per SM); HOP_K2_CFG selects one (tuning knob)
template <int CFG> struct GtCfg;
template <> struct GtCfg<0> { static constexpr int T = 336, B = 2; };
template <> struct GtCfg<1> { static constexpr int T = 224, B = 2; };
template <> struct GtCfg<2> { static constexpr int T = 448, B = 1; };
template <> struct GtCfg<3> { static constexpr int T = 224, B = 3; };
template <> struct GtCfg<4> { static constexpr int T = 448, B = 2; };
template <> struct GtCfg<5> { static constexpr int T = 224, B = 4; };
template <> struct GtCfg<6> { static constexpr int T = 896, B = 1; };

// Start vector b of the diamond search in integer pels (TEncSearch.cpp:5106-5153); false = the reference skips it.
__device__ __forceinline__ bool gt_start_vector(const HopGtJob& job, int b, int* Hx, int* Hy)
{
  if (b == 0) {
    if (job.ss_cand.hor == 0 && job.ss_cand.ver == 0) return false;   // :5116
    *Hx = job.ss_cand.hor; *Hy = job.ss_cand.ver;
  } else {
    const HopMv e = job.amvp[b - 1];
    if (e.hor == 0 && e.ver == 0) return false;                       // :5144
    *Hx = (int16_t)(e.hor >> 2); *Hy = (int16_t)(e.ver >> 2);         // :5148-5149
  }
  return true;
}

// Sample (wx, wy) of the window of start vector (Hx, Hy): [Hx - w, Hx + w + cols) x [Hy - w, Hy + w + rows) relative
// to the PU, clamped to [0, 2^bd - 1] (filterCopy first+last, TComInterpolationFilter.cpp:113-154), as the high
// word of its binary64 value (converted once, exact).  AMVP start vectors are raw neighbour vectors: their
// window may leave the reference buffer (the reference then reads whatever lies beyond its plane); the read is
// kept inside the buffer.
__device__ __forceinline__ long long gt_window_offset(long long ref_off, int ref_stride, RefBounds rb, int Hx, int Hy, int w, int wx, int wy)
{
  long long o = ref_off + (long long)(Hy - w + wy) * ref_stride + (Hx - w + wx);
  return o < rb.lo ? rb.lo : (o > rb.hi ? rb.hi : o);
}
__device__ __forceinline__ uint32_t gt_window_word(int v, int max_val)
{
  v = min(max(v, 0), max_val);
  return (uint32_t)__double2hiint((double)v);
}
template <int WS>
__device__ __forceinline__ void gt_stage_window(const HopGtJob& job, const int16_t* __restrict__ ref_buf, RefBounds rb,
                                                int Hx, int Hy, int w, int win_w, int win_h, int max_val, uint32_t* win,
                                                bool with_dx = false)
{
  if (with_dx) {     // sample rows interleaved with rows of horizontal differences (warp_sample<WS, true>)
    for (int i = threadIdx.x; i < win_w * win_h; i += blockDim.x) {
      const int wy = i / win_w, wx = i - wy * win_w;
      const int v0 = min(max((int)ref_buf[gt_window_offset(job.ref_off, job.ref_stride, rb, Hx, Hy, w, wx, wy)], 0), max_val);
      const int v1 = wx + 1 < win_w ? min(max((int)ref_buf[gt_window_offset(job.ref_off, job.ref_stride, rb, Hx, Hy, w, wx + 1, wy)], 0), max_val) : v0;
      win[wy * (2 * WS + 1) + wx] = (uint32_t)__double2hiint((double)v0);
      win[wy * (2 * WS + 1) + WS + wx] = (uint32_t)__double2hiint((double)(v1 - v0));
    }
    return;
  }
  for (int i = threadIdx.x; i < win_w * win_h; i += blockDim.x) {
    const int wy = i / win_w, wx = i - wy * win_w;
    win[wy * WS + wx] = gt_window_word(ref_buf[gt_window_offset(job.ref_off, job.ref_stride, rb, Hx, Hy, w, wx, wy)], max_val);
  }
}

// The whole xPatternSearchGT of one PU, executed by one CTA.  `out` is written by thread 0.
// smem_raw: [GtShared][org rows*cols int32][window (rows+2w) x WS uint32]; when `org_staged` the int32
// original block is already in place (the fused motion kernel stages it once for all stages).
//
// CL = true: the PU is searched by a thread-block CLUSTER (single-call latency path for large PUs): every
// CTA stages the same window and walks the same passes, but evaluates only its share of the Hadamard tiles;
// after a pass the per-candidate partial sums are gathered through distributed shared memory and every CTA
// takes the same argmin decision, so no state has to be broadcast.  Rank 0 writes the result.
template <int WS, bool CL = false, bool PAIR = false>
__device__ __forceinline__ void gt_search_cta(const HopGtJob& job, const int16_t* __restrict__ org_buf,
                                              const int16_t* __restrict__ ref_buf, unsigned char* smem_raw,
                                              HopGtResult* __restrict__ out, bool org_staged, RefBounds rb,
                                              uint32_t* win_slots = nullptr, int slot_words = 0, unsigned pre_mask = 0)
{
  GtShared& sh = *reinterpret_cast<GtShared*>(smem_raw);
  int crank = 0, csize = 1;
  if (CL) { cg::cluster_group cl = cg::this_cluster(); crank = (int)cl.block_rank(); csize = (int)cl.num_blocks(); }
  const int cols = job.cols, rows = job.rows;
  const int G = 2;                                              // IT_GT_GRID_SIZE
  const int nss_window = ((rows < cols ? rows : cols) >> 1) * G;  // TEncSearch.cpp:4756-4758
  const int w = nss_window / G;                                 // clamp radius inside ProjectiveTransform
  int last_step = nss_window >> 6;                              // :4763 (IT_MAX_NSS_Iteration 6)
  if (last_step == 0) last_step = 1;
  const int win_w = cols + 2 * w, win_h = rows + 2 * w;
  double* s_div_w = reinterpret_cast<double*>(smem_raw + GT_SHARED_BYTES);
  double* s_div_h = s_div_w + gt_div_entries(cols * G, nss_window);
  int* s_org = reinterpret_cast<int*>(smem_raw + GT_SHARED_BYTES + gt_div_bytes(cols, rows));
  uint32_t* s_win = reinterpret_cast<uint32_t*>(s_org + ((rows * cols + 3) & ~3));
  if (win_w > WS) return;   // host picked the wrong stride class (cannot happen through the ABI)
  const int16_t* org = org_buf + job.org_off;
  const int max_val = (1 << job.bit_depth) - 1;
  const int dist_shift = job.bit_depth - 8;
  const int tile_n = ((rows % 8 == 0) && (cols % 8 == 0)) ? 8 : 4;

  // batched kernel, Hadamard cost, 8x8 tiles that pair up: the original block is staged packed, two tiles per word
  // (eval_half_tile8_pair); it takes the place of the int32 copy, which nothing else reads in that kernel
  const PairGeom pg = gt_pair_geom(cols, rows);
  // a batch (the launch sized the window for the difference rows); <= 10-bit residuals keep every lane below 2^15
  const bool packed = PAIR && !org_staged && pg.on && job.use_had && job.bit_depth <= 10 && (int)gridDim.x > 1;
  if (packed) {
    unsigned* s_orgp = reinterpret_cast<unsigned*>(s_org);
    for (int i = threadIdx.x; i < pg.ph * pg.pw; i += blockDim.x) {
      const int y = i / pg.pw, x = i - y * pg.pw;
      const unsigned a = (unsigned)org[y * job.org_stride + x], b = (unsigned)org[(y + pg.dyB) * job.org_stride + x + pg.dxB];
      s_orgp[i] = (a + 0x8000u) | ((b + 0x8000u) << 16);
    }
  } else if (!org_staged)
    for (int i = threadIdx.x; i < rows * cols; i += blockDim.x)
      s_org[i] = org[(i / cols) * job.org_stride + (i % cols)];
  {
    const double Wd = __dsub_rn((double)(cols * G), 1.0), Hd = __dsub_rn((double)(rows * G), 1.0);   // :811-812
    const int nw = gt_div_entries(cols * G, nss_window), nh = gt_div_entries(rows * G, nss_window);
    for (int i = threadIdx.x; i < nw + nh; i += blockDim.x) {
      if (i < nw) s_div_w[i] = __ddiv_rn((double)(i - 2 * nss_window), Wd);
      else        s_div_h[i - nw] = __ddiv_rn((double)(i - nw - 2 * nss_window), Hd);
    }
  }
  __syncthreads();   // quotient tables (and the original block) visible before the first candidate set-up

  // Search state is replicated per head warp (threads 0..63 = two warps): both warps reduce all 56 candidate
  // costs of a pass and take the same decision, so a pass needs two CTA barriers (table ready, tiles done) and
  // no serial section.  Thread c < 56 also owns candidate c of the per-pass table.
  // single PU in the grid (the in-encoder call): row-per-lane tiles when that takes a single trip, i.e. the CTA
  // has 8 lanes for every (candidate, 8x8 tile) it owns -- measured 1.9 us per pass against 3.0 us for the
  // half-tile form; with two trips, or with 4x4 tiles (1.25 us either way), the register tiles stay
  const int my_tiles = ((cols / tile_n) * (rows / tile_n) + csize - 1) / csize;
  const bool fine = (int)gridDim.x == csize && tile_n == 8 && my_tiles * 8 * GT_CANDS <= (int)blockDim.x;
  const bool head = threadIdx.x < 64;
  const int lane = threadIdx.x & 31;
  const int c_own = threadIdx.x;                                // table entry built by this thread (if < 56)
  GtShared::State& st = sh.st[(threadIdx.x >> 5) & 1];
  if (c_own < GT_CANDS) sh.offs[c_own] = __ldg(reinterpret_cast<const int2*>(d_gt_offsets) + c_own);   // read back by the same thread
  if (head && lane == 0) {
    st.bc = make_int4(0, 0, 0, 0);
    st.dist_best = job.threshold;                               // :4769
    st.best_ss_x = 0; st.best_ss_y = 0; st.best_index = -1; st.n_cand = 0;
  }
  int pb = 0;                                                   // parity of the double-buffered arrays

  const int n_start = 1 + job.num_pred;
  for (int b = 0; b < n_start; b++) {                           // :5106-5110
    int Hx, Hy;                                                 // integer-pel start vector
    if (!gt_start_vector(job, b, &Hx, &Hy)) continue;
    const int Hor = (int16_t)(Hx << 2), Ver = (int16_t)(Hy << 2);   // Short, quarter-pel
    // latency path: one window slot per start vector, staged ahead by the caller (pre_mask); otherwise one slot,
    // restaged per start
    uint32_t* const win_b = win_slots ? win_slots + (size_t)b * slot_words : s_win;
    if (!((pre_mask >> b) & 1u)) {
      if (!win_slots) __syncthreads();   // previous start done with the slot
      gt_stage_window<WS>(job, ref_buf, rb, Hx, Hy, w, win_w, win_h, max_val, win_b, PAIR && packed);
    }
    const uint32_t mv_add = mv_cost(job.cost, Hor, Ver);        // :5345
    HOP_STAMP(g_trace_k2, 8 + b * 16);   // window loads of start b issued
    // first pass of a start vector: the centres are the initial rectangle (:5183-5203)
    if (head && lane == 0) st.cc = make_int4(0, cols * G - 1, ((rows * G - 1) << 16) | (cols * G - 1), (rows * G - 1) << 16);
    if (head) __syncwarp();

    int pass = 0;
    for (int j0 = nss_window; j0 > 1 && pass < 6; j0 /= 2, pass++) {   // :5181
      const int s = j0 / 2;
      pb ^= 1;
      if (c_own < GT_CANDS) {
        const int c = c_own;
        int cx[4], cy[4];
        const int4 cc = st.cc;
        const int ccw[4] = {cc.x, cc.y, cc.z, cc.w};
        const int2 offw = sh.offs[c];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int w2 = k < 2 ? offw.x : offw.y;
          cx[k] = (int)(int16_t)(ccw[k] & 0xffff) + s * (int)(int8_t)(w2 >> (16 * (k & 1)));
          cy[k] = (ccw[k] >> 16) + s * (int)(int8_t)(w2 >> (16 * (k & 1) + 8));
        }
        // calcParamProjective, TComPrediction.cpp:807-832.  The affine test "h[2] == 0.0 && h[5] == 0.0"
        // (:5323) is decided exactly in integers: the numerators of h[2], h[5] are products of small
        // integers (exact in binary64) and vanish together iff dx3 == dy3 == 0 when den != 0; den == 0
        // gives 0/0 = NaN, which fails the reference's comparison.  For an accepted candidate
        // h[2] = h[5] = +-0, so h[0] = fl((x1-x0)/W) + (+-0) = fl((x1-x0)/W) etc. -- same values, no
        // division chain through den.
        const int idx3 = cx[0] - cx[1] + cx[2] - cx[3], idy3 = cy[0] - cy[1] + cy[2] - cy[3];
        const int iden = (cx[1] - cx[2]) * (cy[3] - cy[2]) - (cx[3] - cx[2]) * (cy[1] - cy[2]);
        const int ok = (idx3 == 0 && idy3 == 0 && iden != 0) ? 1 : 0;
        if (b == 0 && pass == 1) HOP_STAMP(g_trace_k2, 57);
        sh.valid[pb][c] = ok;
        sh.dist[pb][c] = 0;
        if (ok) {
          const int nb = 2 * nss_window;                        // table index of numerator 0
          sh.h0[c] = s_div_w[cx[1] - cx[0] + nb];               // fl((x1 - x0) / (W - 1))
          sh.h3[c] = s_div_h[cx[3] - cx[0] + nb];               // fl((x3 - x0) / (H - 1))
          sh.h6[c] = (double)cx[0];
          sh.h1[c] = s_div_w[cy[1] - cy[0] + nb];
          sh.h4[c] = s_div_h[cy[3] - cy[0] + nb];
          sh.h7[c] = (double)cy[0];
          int g0x = cx[0], g0y = cy[0], g1x = cx[1] - cols * G + 1, g1y = cy[1];
          int g2x = cx[2] - cols * G + 1, g2y = cy[2] - rows * G + 1;
          if (last_step != 1) {                                 // only PUs beyond 64x64 have a coarser last step
            g0x /= last_step; g0y /= last_step; g1x /= last_step; g1y /= last_step; g2x /= last_step; g2y /= last_step;
          }
          sh.add_cost[pb][c] = mv_add + bits_cost(job.cost, gt_bits(g0x, g0y, g1x, g1y, g2x, g2y));   // :5345-5358
#pragma unroll
          for (int k = 0; k < 4; k++) { sh.corner[pb][c][2 * k] = (int16_t)cx[k]; sh.corner[pb][c][2 * k + 1] = (int16_t)cy[k]; }
        }
        if (b == 0 && pass == 1) HOP_STAMP(g_trace_k2, 58);
      }
      __syncthreads();
      HOP_STAMP(g_trace_k2, 9 + b * 16 + 2 * pass);    // candidate table of the pass built
      if (PAIR && packed) run_tasks8_pair<WS, true>(sh, reinterpret_cast<const unsigned*>(s_org), win_b, pg, w, cols, rows, cols >> 1, rows >> 1, pb);
      else run_tasks<WS>(sh, s_org, win_b, w, cols, rows, cols >> 1, rows >> 1, tile_n, job.use_had, crank, csize, pb, fine);
      if (CL) cg::this_cluster().sync(); else __syncthreads();
      HOP_STAMP(g_trace_k2, 10 + b * 16 + 2 * pass);   // tiles of the pass evaluated
      if (head) {
        // ordered argmin with the carried threshold: the serial loop keeps the FIRST strict minimum in
        // loop order (:5361) == the minimum of (cost, loop index) over the pass, accepted iff it beats
        // the running best.  Lane l looks at candidates l and l + 32.
        unsigned long long key = ~0ull;
        unsigned n_ok = 0;
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
          const int c = lane + 32 * hh;
          if (c < GT_CANDS && sh.valid[pb][c]) {
            uint32_t dsum = 0;
            if (CL) {
              cg::cluster_group cl = cg::this_cluster();
              // tile sums of all CTAs: independent remote loads (one DSMEM latency, not csize of them)
#pragma unroll
              for (int r = 0; r < 8; r++) if (r < csize) dsum += *cl.map_shared_rank(&sh.dist[pb][c], r);
            } else {
              dsum = sh.dist[pb][c];
            }
            const unsigned long long k2 = ((unsigned long long)((dsum >> dist_shift) + sh.add_cost[pb][c]) << 8) | (unsigned)c;
            key = k2 < key ? k2 : key;
            n_ok++;
          }
        }
        // warp minimum of the 64-bit key with two redux operations: cost first, then the loop index among
        // the lanes that hold that cost
        const unsigned kcost = (unsigned)(key >> 8) | (key == ~0ull ? 0xffffffffu : 0u);
        const unsigned mcost = __reduce_min_sync(0xffffffffu, kcost);
        const unsigned mc = __reduce_min_sync(0xffffffffu, (key != ~0ull && kcost == mcost) ? (unsigned)(key & 0xff) : 0xffffffffu);
        n_ok = __reduce_add_sync(0xffffffffu, n_ok);
        key = mc == 0xffffffffu ? ~0ull : ((unsigned long long)mcost << 8) | mc;
        __syncwarp();                       // every lane has read st.cc for the table of this pass
        if (lane == 0) {
          st.n_cand += n_ok;
          if (key != ~0ull && (uint32_t)(key >> 8) < st.dist_best) {   // :5363-5383
            const int best_c = (int)(key & 0xff);
            st.dist_best = (uint32_t)(key >> 8);
            const int4 cr = *reinterpret_cast<const int4*>(sh.corner[pb][best_c]);   // 8 x int16
            st.cc = cr; st.bc = cr;
            st.best_ss_x = Hor; st.best_ss_y = Ver;
            st.best_index = (b * 8 + pass) * 64 + best_c;
          }
        }
        __syncwarp();
      }
      if (b == 0 && pass == 0) HOP_STAMP(g_trace_k2, 56);
      // cluster: the partial sums of this pass are zeroed again two passes later (other parity), after the
      // next cluster barrier -- no CTA can still be gathering them
    }
  }
  if (threadIdx.x == 0 && crank == 0) {
    HopGtResult r;
    r.gt_flag = 0;
    for (int k = 0; k < 4; k++) { r.gt[k].hor = 0; r.gt[k].ver = 0; }
    r.cost = job.threshold;
    r.mv_int.hor = 0; r.mv_int.ver = 0;
    r.best_index = -1;
    r.n_candidates = st.n_cand;
    const int4 bc = st.bc;
    const int bcw[4] = {bc.x, bc.y, bc.z, bc.w};
    int bcx[4], bcy[4];
    for (int k = 0; k < 4; k++) { bcx[k] = (int)(int16_t)(bcw[k] & 0xffff); bcy[k] = bcw[k] >> 16; }
    const int any = bc.x | bc.y | bc.z | bc.w;
    if (any) {                                                  // :5436-5459
      r.gt_flag = 1;
      r.gt[0].hor = (int16_t)(bcx[0] / last_step);                  r.gt[0].ver = (int16_t)(bcy[0] / last_step);
      r.gt[1].hor = (int16_t)((bcx[1] - cols * G + 1) / last_step); r.gt[1].ver = (int16_t)(bcy[1] / last_step);
      r.gt[2].hor = (int16_t)((bcx[2] - cols * G + 1) / last_step); r.gt[2].ver = (int16_t)((bcy[2] - rows * G + 1) / last_step);
      r.gt[3].hor = (int16_t)(bcx[3] / last_step);                  r.gt[3].ver = (int16_t)((bcy[3] - rows * G + 1) / last_step);
      r.cost = st.dist_best;
      r.mv_int.hor = (int16_t)(st.best_ss_x >> 2);
      r.mv_int.ver = (int16_t)(st.best_ss_y >> 2);
      r.best_index = st.best_index;
    }
    *out = r;
  }
  if (CL) cg::this_cluster().sync();   // no CTA may leave while another still reads its partial sums
}

template <int WS, int CFG>
__global__ void __launch_bounds__(GtCfg<CFG>::T, GtCfg<CFG>::B)
k2_gt_search(int n_jobs, const HopGtJob* __restrict__ jobs, const int16_t* __restrict__ org_buf,
             const int16_t* __restrict__ ref_buf, HopGtResult* __restrict__ out,
             unsigned* done_flag, unsigned seq, RefBounds rb)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int job_id = blockIdx.x;
  if (job_id >= n_jobs) return;
  const HopGtJob job = jobs[job_id];
  gt_search_cta<WS, false, true>(job, org_buf, ref_buf, smem_raw, &out[job_id], false, rb);
  if (threadIdx.x == 0 && done_flag) {     // single-call path: result and flag live in mapped host memory
    __threadfence_system();
    *(volatile unsigned*)done_flag = seq;
  }
}

// ---- exhaustive sweep (reference mode IT_GT_SEARCH 1 + IT_GT_GRID_SIZE 1) ------------------------
// TEncSearch.cpp:4989-5091: one pass over the 25^4 corner sets around the initial rectangle of the 1x
// grid; the 7200 parallelogram patterns (in loop order, with their flat loop index) sit in a device
// table, a slice [cand_begin, cand_end) of which is evaluated -- the unit the multi-GPU sweep shards.
// Candidates are processed in batches of 56 with the task machinery of the diamond search; the CTA's
// partial argmin (cost << 32 | flat loop index) is merged with atomicMin into one word per PU.
__device__ SweepCand d_sweep_table[SWEEP_CANDS];

cudaError_t sweep_upload_table(const SweepCand* table)
{
  return cudaMemcpyToSymbol(d_sweep_table, table, sizeof(SweepCand) * SWEEP_CANDS);
}

void sweep_build_table(SweepCand* table, int* count)
{
  int n = 0;
  const int N = 2;
  for (int y0 = -N; y0 <= N; y0++) for (int x0 = -N; x0 <= N; x0++)
  for (int y1 = -N; y1 <= N; y1++) for (int x1 = -N; x1 <= N; x1++)
  for (int y2 = -N; y2 <= N; y2++) for (int x2 = -N; x2 <= N; x2++)
  for (int y3 = -N; y3 <= N; y3++) for (int x3 = -N; x3 <= N; x3++) {
    if (x0 == x1 && x0 == x2 && x0 == x3 && y0 == y1 && y0 == y2 && y0 == y3) continue;   // :5017
    if (x0 - x1 + x2 - x3 != 0 || y0 - y1 + y2 - y3 != 0) continue;                        // not affine
    if (n < SWEEP_CANDS) {
      SweepCand& c = table[n];
      c.o[0] = x0; c.o[1] = y0; c.o[2] = x1; c.o[3] = y1; c.o[4] = x2; c.o[5] = y2; c.o[6] = x3; c.o[7] = y3;
      c.flat = (uint32_t)((((((((y0 + N) * 5 + (x0 + N)) * 5 + (y1 + N)) * 5 + (x1 + N)) * 5 + (y2 + N)) * 5 +
                            (x2 + N)) * 5 + (y3 + N)) * 5 + (x3 + N));
    }
    n++;
  }
  *count = n;
}

// Sharded form (SweepXchg.world > 0): the partial argmin of a PU leaves this GPU the moment the PU's last CTA is
// done -- one 64-bit atomicMin (and one 32-bit atomicAdd of the candidate count) per rank into the merge words that
// every rank keeps in its own HBM, straight through peer-mapped memory over NVLink, overlapped with the CTAs
// still computing.  The last CTA of the whole grid then bumps an arrival counter on every rank.  No collective
// call, no extra launch: the all-reduce-min of SURVEY.md 8e is done by the kernel that produces the keys.
template <int WS>
__global__ void __launch_bounds__(gt_class_threads(WS), 2)
k2_gt_sweep(int n_jobs, const HopGtJob* __restrict__ jobs, const int16_t* __restrict__ org_buf,
            const int16_t* __restrict__ ref_buf, int cand_begin, int cand_end,
            unsigned long long* __restrict__ keys, unsigned int* __restrict__ counts, RefBounds rb,
            const __grid_constant__ SweepXchg xc)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GtShared& sh = *reinterpret_cast<GtShared*>(smem_raw);
  __shared__ uint32_t s_flat[GT_CANDS];
  const int job_id = blockIdx.x;
  if (job_id >= n_jobs) return;
  const HopGtJob job = jobs[job_id];
  const int cols = job.cols, rows = job.rows;
  const int w = (rows < cols ? rows : cols) >> 1;               // iNSSWindow on the 1x grid (:4756)
  const int win_w = cols + 2 * w, win_h = rows + 2 * w;
  int* s_org = reinterpret_cast<int*>(smem_raw + GT_SHARED_BYTES);
  uint32_t* s_win = reinterpret_cast<uint32_t*>(s_org + ((rows * cols + 3) & ~3));
  if (win_w > WS) return;
  const int16_t* org = org_buf + job.org_off;
  const int max_val = (1 << job.bit_depth) - 1;
  const int dist_shift = job.bit_depth - 8;
  const int tile_n = ((rows % 8 == 0) && (cols % 8 == 0)) ? 8 : 4;
  const int mvx = job.ss_cand.hor, mvy = job.ss_cand.ver;       // pcMvInt
  const int Hor = (int16_t)(mvx << 2), Ver = (int16_t)(mvy << 2);   // :4713-4724 with half = quarter = 0

  const PairGeom pg = gt_pair_geom(cols, rows);
  const bool packed = pg.on && job.use_had && job.bit_depth <= 10;   // two tiles per register tile (eval_half_tile8_pair)
  constexpr bool SWEEP_DX = WS != WS_D;                         // difference rows: not for the 64-wide class (two CTAs per SM stay resident)
  if (packed) {
    unsigned* s_orgp = reinterpret_cast<unsigned*>(s_org);
    for (int i = threadIdx.x; i < pg.ph * pg.pw; i += blockDim.x) {
      const int y = i / pg.pw, x = i - y * pg.pw;
      const unsigned a = (unsigned)org[y * job.org_stride + x], b = (unsigned)org[(y + pg.dyB) * job.org_stride + x + pg.dxB];
      s_orgp[i] = (a + 0x8000u) | ((b + 0x8000u) << 16);
    }
  } else
    for (int i = threadIdx.x; i < rows * cols; i += blockDim.x)
      s_org[i] = org[(i / cols) * job.org_stride + (i % cols)];
  for (int i = threadIdx.x; i < win_w * win_h; i += blockDim.x) {
    const int wy = i / win_w, wx = i - wy * win_w;
    long long o = job.ref_off + (long long)(mvy - w + wy) * job.ref_stride + (mvx - w + wx);
    o = o < rb.lo ? rb.lo : (o > rb.hi ? rb.hi : o);
    int v = ref_buf[o];
    v = min(max(v, 0), max_val);
    if (packed && SWEEP_DX) {                                   // with the row of horizontal differences (warp_sample<WS, true>)
      long long o1 = job.ref_off + (long long)(mvy - w + wy) * job.ref_stride + (mvx - w + wx + 1);
      o1 = o1 < rb.lo ? rb.lo : (o1 > rb.hi ? rb.hi : o1);
      const int v1 = wx + 1 < win_w ? min(max((int)ref_buf[o1], 0), max_val) : v;
      s_win[wy * (2 * WS + 1) + wx] = (uint32_t)__double2hiint((double)v);
      s_win[wy * (2 * WS + 1) + WS + wx] = (uint32_t)__double2hiint((double)(v1 - v));
    } else
      s_win[wy * WS + wx] = (uint32_t)__double2hiint((double)v);
  }
  const uint32_t mv_add = mv_cost(job.cost, Hor, Ver);          // :5035
  unsigned long long best = ~0ull;
  unsigned int scored = 0;

  const int n_batches = (cand_end - cand_begin + GT_CANDS - 1) / GT_CANDS;
  for (int batch = blockIdx.y; batch < n_batches; batch += gridDim.y) {
    __syncthreads();
    if (threadIdx.x < GT_CANDS) {
      const int c = threadIdx.x, idx = cand_begin + batch * GT_CANDS + c;
      int ok = 0;
      if (idx < cand_end) {
        const SweepCand sc = d_sweep_table[idx];
        const int ox[4] = {sc.o[0], sc.o[2], sc.o[4], sc.o[6]}, oy[4] = {sc.o[1], sc.o[3], sc.o[5], sc.o[7]};
        const int cx[4] = {ox[0], cols - 1 + ox[1], cols - 1 + ox[2], ox[3]};               // :4781-4784
        const int cy[4] = {oy[0], oy[1], rows - 1 + oy[2], rows - 1 + oy[3]};
        // valid GT location, marginX = marginY = 0 (:5020-5023)
        const int ax[4] = {ox[0] + mvx, ox[1] + mvx + cols, ox[2] + mvx + cols, ox[3] + mvx};
        const int ay[4] = {oy[0] + mvy, oy[1] + mvy, oy[2] + mvy + rows, oy[3] + mvy + rows};
        ok = 1;
#pragma unroll
        for (int k = 0; k < 4; k++) ok &= ((ax[k] < 0 && ay[k] <= 0) || (ax[k] >= 0 && ay[k] < 0)) ? 1 : 0;
        if (ok) {
          // calcParamProjective(iCurrCornerX, iCurrCornerY, dProjective, iCols, iRows), TComPrediction.cpp:807-832;
          // affine test (:5028) decided exactly in integers, see k2_gt_search
          const double Wd = __dsub_rn((double)cols, 1.0), Hd = __dsub_rn((double)rows, 1.0);
          const int idx3 = cx[0] - cx[1] + cx[2] - cx[3], idy3 = cy[0] - cy[1] + cy[2] - cy[3];
          const int iden = (cx[1] - cx[2]) * (cy[3] - cy[2]) - (cx[3] - cx[2]) * (cy[1] - cy[2]);
          ok = (idx3 == 0 && idy3 == 0 && iden != 0) ? 1 : 0;
          if (ok) {
            sh.h0[c] = __ddiv_rn((double)(cx[1] - cx[0]), Wd);
            sh.h3[c] = __ddiv_rn((double)(cx[3] - cx[0]), Hd);
            sh.h6[c] = (double)cx[0];
            sh.h1[c] = __ddiv_rn((double)(cy[1] - cy[0]), Wd);
            sh.h4[c] = __ddiv_rn((double)(cy[3] - cy[0]), Hd);
            sh.h7[c] = (double)cy[0];
            const uint32_t gb = gt_bits(cx[0], cy[0], cx[1] - cols + 1, cy[1], cx[2] - cols + 1, cy[2] - rows + 1);
            sh.add_cost[0][c] = mv_add + bits_cost(job.cost, gb);                                // :5035-5041
            s_flat[c] = sc.flat;
          }
        }
      }
      sh.valid[0][c] = ok;
      sh.dist[0][c] = 0;
    }
    __syncthreads();
    if (packed) run_tasks8_pair<WS, SWEEP_DX>(sh, reinterpret_cast<const unsigned*>(s_org), s_win, pg, w, cols, rows, 0, 0, 0);
    else run_tasks<WS>(sh, s_org, s_win, w, cols, rows, 0, 0, tile_n, job.use_had);
    __syncthreads();
    if (threadIdx.x < 64) {
      const int c = threadIdx.x;
      const bool ok = c < GT_CANDS && sh.valid[0][c];
      unsigned long long key = ok ? ((unsigned long long)((sh.dist[0][c] >> dist_shift) + sh.add_cost[0][c]) << 32) | s_flat[c] : ~0ull;
      const unsigned n_ok = __popc(__ballot_sync(0xffffffffu, ok));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other < key ? other : key;
      }
      if ((threadIdx.x & 31) == 0) { sh.red_key[threadIdx.x >> 5] = key; sh.red_cnt[threadIdx.x >> 5] = n_ok; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned long long key = sh.red_key[0] < sh.red_key[1] ? sh.red_key[0] : sh.red_key[1];
      best = key < best ? key : best;
      scored += sh.red_cnt[0] + sh.red_cnt[1];
    }
  }
  if (threadIdx.x == 0) {
    if (best != ~0ull) atomicMin(&keys[job_id], best);
    if (counts && scored) atomicAdd(&counts[job_id], scored);
    if (xc.world > 0) {
      __threadfence();
      if (atomicAdd(&xc.pu_done[job_id], 1u) == gridDim.y - 1) {          // last CTA of this PU on this GPU
        __threadfence();
        const unsigned long long key = atomicExch(&keys[job_id], ~0ull);  // merge words return to their idle state
        const unsigned int cnt = atomicExch(&counts[job_id], 0u);
        xc.pu_done[job_id] = 0;
        const size_t slot = (size_t)xc.parity * xc.max_pus + job_id;
        for (int r = 0; r < xc.world; r++) {
          if (key != ~0ull) atomicMin_system(xc.gkeys[r] + slot, key);
          if (cnt) atomicAdd_system(xc.gcounts[r] + slot, cnt);
        }
        __threadfence_system();
        if (atomicAdd(xc.grid_done, 1u) == gridDim.x - 1) {               // last PU of the grid: everything is out
          *xc.grid_done = 0;
          __threadfence_system();
          for (int r = 0; r < xc.world; r++) atomicAdd_system(xc.arrived[r], 1u);
        }
      }
    }
  }
}

__global__ void k2_sweep_init(int n, unsigned long long* keys, unsigned int* counts)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { keys[i] = ~0ull; if (counts) counts[i] = 0; }
}

// key -> the outputs xPatternSearchGT (mode 1) leaves (:5070-5090)
__device__ __forceinline__ void sweep_result(const HopGtJob& job, unsigned long long key, unsigned int count, HopGtResult* out)
{
  const int cols = job.cols, rows = job.rows, N = 2;
  HopGtResult r;
  r.gt_flag = 0;
  for (int k = 0; k < 4; k++) { r.gt[k].hor = 0; r.gt[k].ver = 0; }
  r.cost = job.threshold;
  r.mv_int.hor = 0; r.mv_int.ver = 0;
  r.best_index = -1;
  r.n_candidates = count;
  if (key != ~0ull && (uint32_t)(key >> 32) < job.threshold) {                               // uiDist < uiDistBest
    uint32_t flat = (uint32_t)key;
    int o[8];
    for (int k = 7; k >= 0; k--) { o[k] = (int)(flat % 5) - N; flat /= 5; }                 // y0,x0,...,y3,x3
    const int bx[4] = {o[1], cols - 1 + o[3], cols - 1 + o[5], o[7]};
    const int by[4] = {o[0], o[2], rows - 1 + o[4], rows - 1 + o[6]};
    int any = 0;
    for (int k = 0; k < 4; k++) any |= bx[k] | by[k];
    if (any) {
      r.gt_flag = 1;
      r.gt[0].hor = (int16_t)bx[0];              r.gt[0].ver = (int16_t)by[0];
      r.gt[1].hor = (int16_t)(bx[1] - cols + 1); r.gt[1].ver = (int16_t)by[1];
      r.gt[2].hor = (int16_t)(bx[2] - cols + 1); r.gt[2].ver = (int16_t)(by[2] - rows + 1);
      r.gt[3].hor = (int16_t)bx[3];              r.gt[3].ver = (int16_t)(by[3] - rows + 1);
      r.cost = (uint32_t)(key >> 32);
      r.best_index = (int32_t)(uint32_t)key;
    }
  }
  *out = r;
}

__global__ void k2_sweep_finalize(int n, const HopGtJob* __restrict__ jobs, const unsigned long long* __restrict__ keys,
                                  const unsigned int* __restrict__ counts, HopGtResult* __restrict__ out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  sweep_result(jobs[i], keys[i], counts ? counts[i] : 0, &out[i]);
}

// Second and last kernel of the sharded sweep: wait until every rank's grid has reported, then turn the merged
// words into results (every rank computes the same ones) and hand the words back in their idle state for the
// sweep after next (they are double buffered by sweep parity: a fast peer may already be pushing the next sweep).
__global__ void k2_sweep_finalize_x(int n, const HopGtJob* __restrict__ jobs, unsigned long long* gkeys, unsigned int* gcounts,
                                    const unsigned int* arrived, unsigned int target, HopGtResult* __restrict__ out)
{
  __shared__ int s_timed_out;
  if (threadIdx.x == 0) {
    // a peer that never reports (its process died) must not hang this GPU: give up after 10 s and say so
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    int late = 0;
    while ((int)(*(volatile const unsigned int*)arrived - target) < 0) {
      __nanosleep(200);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 10000000000ull) { late = 1; break; }
    }
    s_timed_out = late;
    __threadfence_system();
  }
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = *(volatile unsigned long long*)(gkeys + i);
  const unsigned int cnt = *(volatile unsigned int*)(gcounts + i);
  gkeys[i] = ~0ull;
  gcounts[i] = 0;
  sweep_result(jobs[i], key, cnt, &out[i]);
  if (s_timed_out) { out[i].gt_flag = -1; out[i].best_index = -2; }    // incomplete exchange: never a silent partial result
}

static size_t gt_smem_bytes(int ws, int max_cols, int max_rows, bool with_dx = false)
{
  const int w = (max_cols < max_rows ? max_cols : max_rows) >> 1;
  const size_t org = ((size_t)max_cols * max_rows + 3) & ~(size_t)3;
  return GT_SHARED_BYTES + gt_div_bytes(max_cols, max_rows) + sizeof(int) * org +
         sizeof(uint32_t) * (size_t)(with_dx ? 2 * ws + 1 : ws) * (max_rows + 2 * w);
}

template <int WS, int CFG>
static cudaError_t gt_launch_cfg(int n, const HopGtJob* d_jobs, const int16_t* d_org, const int16_t* d_ref,
                                 HopGtResult* d_out, int max_cols, int max_rows, cudaStream_t stream,
                                 unsigned* done_flag, unsigned seq, RefBounds rb)
{
  static SmemOptIn opt_in;
  {
    cudaError_t e = opt_in.ensure(k2_gt_search<WS, CFG>, (int)gt_smem_bytes(WS, HOP_MAX_PU, HOP_MAX_PU, true));
    if (e != cudaSuccess) return e;
  }
  // CTA = 56 candidates x `groups` tile groups x (2 lanes per 8x8 tile | 1 lane per 4x4 tile)
  const int tile = ((max_rows % 8 == 0) && (max_cols % 8 == 0)) ? 8 : 4;
  const int per_group = tile == 8 ? 2 * GT_CANDS : GT_CANDS;
  const int ntiles = (max_cols / tile) * (max_rows / tile) / (gt_pair_geom(max_cols, max_rows).on && n > 1 ? 2 : 1);   // tile pairs
  const int max_groups = GtCfg<CFG>::T / per_group;
  int groups = ntiles < max_groups ? ntiles : max_groups;
  // fewest loop trips wins; on a tie the smaller CTA (less idle lanes in the last trip)
  for (int g = groups - 1; g >= 1; g--)
    if ((ntiles + g - 1) / g <= (ntiles + groups - 1) / groups) groups = g;
  int threads = per_group * groups;
  if (threads < 64) threads = 64;   // set-up and argmin use the first 64 threads
  if (n == 1) threads = GtCfg<CFG>::T;   // single PU: all lanes, for the row-per-lane tiles (gt_search_cta: fine)
  k2_gt_search<WS, CFG><<<n, threads, gt_smem_bytes(WS, max_cols, max_rows, n > 1), stream>>>(n, d_jobs, d_org, d_ref, d_out, done_flag, seq, rb);
  return cudaGetLastError();
}

template <int WS>
static cudaError_t gt_launch_class(int n, const HopGtJob* d_jobs, const int16_t* d_org, const int16_t* d_ref,
                                   HopGtResult* d_out, int max_cols, int max_rows, cudaStream_t stream,
                                   unsigned* done_flag, unsigned seq, RefBounds rb)
{
  // measured on B200 (profiles/r01_k2_launch_cfg.txt): the kernel wants resident warps more than registers --
  // 4 CTAs x 224 threads (72 registers, a few spills outside the pixel loop) win for the small stride classes,
  // 2 CTAs x 448 threads for the 64x64 class (shared memory allows only two CTAs there)
  static int env_cfg = -2;
  if (env_cfg == -2) { const char* e = getenv("HOP_K2_CFG"); env_cfg = e ? atoi(e) : -1; }
  // a single PU (latency path) takes the widest CTA: all of its tiles at once
  const int cfg = env_cfg >= 0 ? env_cfg : (n == 1 ? 4 : (WS == WS_D ? 6 : 5));
  switch (cfg) {
    case 6:  return gt_launch_cfg<WS, 6>(n, d_jobs, d_org, d_ref, d_out, max_cols, max_rows, stream, done_flag, seq, rb);
    case 1:  return gt_launch_cfg<WS, 1>(n, d_jobs, d_org, d_ref, d_out, max_cols, max_rows, stream, done_flag, seq, rb);
    case 2:  return gt_launch_cfg<WS, 2>(n, d_jobs, d_org, d_ref, d_out, max_cols, max_rows, stream, done_flag, seq, rb);
    case 3:  return gt_launch_cfg<WS, 3>(n, d_jobs, d_org, d_ref, d_out, max_cols, max_rows, stream, done_flag, seq, rb);
    case 4:  return gt_launch_cfg<WS, 4>(n, d_jobs, d_org, d_ref, d_out, max_cols, max_rows, stream, done_flag, seq, rb);
    case 5:  return gt_launch_cfg<WS, 5>(n, d_jobs, d_org, d_ref, d_out, max_cols, max_rows, stream, done_flag, seq, rb);
    default: return gt_launch_cfg<WS, 0>(n, d_jobs, d_org, d_ref, d_out, max_cols, max_rows, stream, done_flag, seq, rb);
  }
}

cudaError_t gt_launch(int n, const HopGtJob* d_jobs, const int16_t* d_org, const int16_t* d_ref,
                      HopGtResult* d_out, int max_cols, int max_rows, cudaStream_t stream, int* launches,
                      RefBounds rb, unsigned* done_flag, unsigned seq)
{
  const int win_w = max_cols + (max_cols < max_rows ? max_cols : max_rows);
  if (launches) (*launches)++;
  switch (gt_stride_class(win_w)) {
    case WS_A:  return gt_launch_class<WS_A>(n, d_jobs, d_org, d_ref, d_out, max_cols, max_rows, stream, done_flag, seq, rb);
    case WS_B:  return gt_launch_class<WS_B>(n, d_jobs, d_org, d_ref, d_out, max_cols, max_rows, stream, done_flag, seq, rb);
    case WS_C:  return gt_launch_class<WS_C>(n, d_jobs, d_org, d_ref, d_out, max_cols, max_rows, stream, done_flag, seq, rb);
    default:  return gt_launch_class<WS_D>(n, d_jobs, d_org, d_ref, d_out, max_cols, max_rows, stream, done_flag, seq, rb);
  }
}

template <int WS>
static cudaError_t sweep_launch_class(int n, const HopGtJob* d_jobs, const int16_t* d_org, const int16_t* d_ref,
                                      int max_cols, int max_rows, int cand_begin, int cand_end, int chunks,
                                      unsigned long long* d_keys, unsigned int* d_counts, cudaStream_t stream, RefBounds rb,
                                      const SweepXchg& xc)
{
  static SmemOptIn opt_in;
  {
    cudaError_t e = opt_in.ensure(k2_gt_sweep<WS>, (int)gt_smem_bytes(WS, HOP_MAX_PU, HOP_MAX_PU, WS != WS_D));
    if (e != cudaSuccess) return e;
  }
  const int tile = ((max_rows % 8 == 0) && (max_cols % 8 == 0)) ? 8 : 4;
  const int per_group = tile == 8 ? 2 * GT_CANDS : GT_CANDS;
  const int ntiles = (max_cols / tile) * (max_rows / tile) / (gt_pair_geom(max_cols, max_rows).on ? 2 : 1);   // tile pairs
  const int max_groups = gt_class_threads(WS) / per_group;
  int groups = ntiles < max_groups ? ntiles : max_groups;
  for (int g = groups - 1; g >= 1; g--)
    if ((ntiles + g - 1) / g <= (ntiles + groups - 1) / groups) groups = g;
  int threads = per_group * groups;
  if (threads < 64) threads = 64;
  k2_gt_sweep<WS><<<dim3(n, chunks), threads, gt_smem_bytes(WS, max_cols, max_rows, WS != WS_D), stream>>>(
      n, d_jobs, d_org, d_ref, cand_begin, cand_end, d_keys, d_counts, rb, xc);
  return cudaGetLastError();
}

cudaError_t sweep_init_launch(int n, unsigned long long* d_keys, unsigned int* d_counts, cudaStream_t stream, int* launches)
{
  k2_sweep_init<<<(n + 255) / 256, 256, 0, stream>>>(n, d_keys, d_counts);
  if (launches) (*launches)++;
  return cudaGetLastError();
}

cudaError_t sweep_keys_launch(int n, const HopGtJob* d_jobs, const int16_t* d_org, const int16_t* d_ref,
                              int max_cols, int max_rows, int cand_begin, int cand_end, int chunks,
                              unsigned long long* d_keys, unsigned int* d_counts, cudaStream_t stream, int* launches, RefBounds rb,
                              const SweepXchg* xchg)
{
  static const SweepXchg none = {};
  const SweepXchg& xc = xchg ? *xchg : none;
  const int win_w = max_cols + (max_cols < max_rows ? max_cols : max_rows);
  if (launches) (*launches)++;
  switch (gt_stride_class(win_w)) {
    case WS_A:  return sweep_launch_class<WS_A>(n, d_jobs, d_org, d_ref, max_cols, max_rows, cand_begin, cand_end, chunks, d_keys, d_counts, stream, rb, xc);
    case WS_B:  return sweep_launch_class<WS_B>(n, d_jobs, d_org, d_ref, max_cols, max_rows, cand_begin, cand_end, chunks, d_keys, d_counts, stream, rb, xc);
    case WS_C:  return sweep_launch_class<WS_C>(n, d_jobs, d_org, d_ref, max_cols, max_rows, cand_begin, cand_end, chunks, d_keys, d_counts, stream, rb, xc);
    default:  return sweep_launch_class<WS_D>(n, d_jobs, d_org, d_ref, max_cols, max_rows, cand_begin, cand_end, chunks, d_keys, d_counts, stream, rb, xc);
  }
}

cudaError_t sweep_finalize_x_launch(int n, const HopGtJob* d_jobs, unsigned long long* d_gkeys, unsigned int* d_gcounts,
                                    const unsigned int* d_arrived, unsigned int target, HopGtResult* d_out, cudaStream_t stream, int* launches)
{
  k2_sweep_finalize_x<<<(n + 255) / 256, 256, 0, stream>>>(n, d_jobs, d_gkeys, d_gcounts, d_arrived, target, d_out);
  if (launches) (*launches)++;
  return cudaGetLastError();
}

cudaError_t sweep_finalize_launch(int n, const HopGtJob* d_jobs, const unsigned long long* d_keys,
                                  const unsigned int* d_counts, HopGtResult* d_out, cudaStream_t stream, int* launches)
{
  k2_sweep_finalize<<<(n + 255) / 256, 256, 0, stream>>>(n, d_jobs, d_keys, d_counts, d_out);
  if (launches) (*launches)++;
  return cudaGetLastError();
}

// ---- K5 stand-alone and the fused motion search ------------------------------------------------------
__global__ void __launch_bounds__(128)
k5_frac_search(int n_jobs, const HopFracJob* __restrict__ jobs, const int16_t* __restrict__ org_buf,
               const int16_t* __restrict__ ref_buf, HopFracResult* __restrict__ out)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int job_id = blockIdx.x;
  if (job_id >= n_jobs) return;
  const HopFracJob job = jobs[job_id];
  FracShared& fs = *reinterpret_cast<FracShared*>(smem_raw);
  int* s_org = reinterpret_cast<int*>(smem_raw + 64);
  const int16_t* org = org_buf + job.org_off;
  for (int i = threadIdx.x; i < job.rows * job.cols; i += blockDim.x)
    s_org[i] = org[(i / job.cols) * job.org_stride + (i % job.cols)];
  unsigned char* scratch = reinterpret_cast<unsigned char*>(s_org + ((job.rows * job.cols + 3) & ~3));
  const int16_t* ref_pos = ref_buf + job.ref_off + job.mv_int.hor + (long long)job.mv_int.ver * job.ref_stride;
  const HopFracResult r = frac_search_cta(fs, s_org, scratch, ref_pos, job.ref_stride, job.cols, job.rows,
                                          job.bit_depth, job.use_had, job.cost, job.mv_int);
  if (threadIdx.x == 0) out[job_id] = r;
}

cudaError_t frac_launch(int n, const HopFracJob* d_jobs, const int16_t* d_org, const int16_t* d_ref,
                        HopFracResult* d_out, int max_cols, int max_rows, cudaStream_t stream, int* launches)
{
  static SmemOptIn opt_in;
  const size_t worst = 64 + sizeof(int) * HOP_MAX_PU * HOP_MAX_PU + frac_smem_bytes(HOP_MAX_PU, HOP_MAX_PU);
  {
    cudaError_t e = opt_in.ensure(k5_frac_search, (int)worst);
    if (e != cudaSuccess) return e;
  }
  const size_t smem = 64 + sizeof(int) * (((size_t)max_cols * max_rows + 3) & ~(size_t)3) + frac_smem_bytes(max_cols, max_rows);
  k5_frac_search<<<n, 128, smem, stream>>>(n, d_jobs, d_org, d_ref, d_out);
  if (launches) (*launches)++;
  return cudaGetLastError();
}

// Second kernel of the fused motion search: one CTA per PU picks up the integer result of k1_search and
// runs xPatternSearchFracDIF and xPatternSearchGT back to back (TEncSearch.cpp:4601-4642), no host in between.
template <int WS, int CFG, bool CL>
__device__ __forceinline__ void motion_tail_body(int n_jobs, const HopMotionJob* __restrict__ jobs,
                                                 const int16_t* __restrict__ org_buf, const int16_t* __restrict__ ref_buf,
                                                 const HopSearchResult* __restrict__ k1, HopMotionResult* __restrict__ out,
                                                 unsigned* done_flag, unsigned seq, RefBounds rb, const InlinePu& ipu)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HOP_STAMP(g_trace_k2, 0);
  int crank = 0, csize = 1;
  if (CL) { cg::cluster_group cl = cg::this_cluster(); crank = (int)cl.block_rank(); csize = (int)cl.num_blocks(); }
  const int job_id = blockIdx.x / csize;
  if (job_id >= n_jobs) return;
  const HopMotionJob mj = ipu.use ? ipu.job : jobs[job_id];
  const HopSearchJob& sj = mj.search;
  const int cols = sj.cols, rows = sj.rows;
  const int w = (rows < cols ? rows : cols) >> 1, win_w = cols + 2 * w, win_h = rows + 2 * w;
  const int max_val = (1 << sj.bit_depth) - 1;
  FracShared& fs = *reinterpret_cast<FracShared*>(smem_raw);               // aliases GtShared, used before it
  int* s_org = reinterpret_cast<int*>(smem_raw + GT_SHARED_BYTES + gt_div_bytes(cols, rows));
  unsigned char* scratch = reinterpret_cast<unsigned char*>(s_org + ((rows * cols + 3) & ~3));
  // Single PU with room in shared memory (the launcher asks for it): the fractional stage and the window of every
  // start vector get their own space, so that all global-memory reads of the call happen in two rounds -- the
  // AMVP windows before K1 has finished, the K1-dependent ones (fractional source, window of the SS vector)
  // together right after -- instead of one exposed latency per stage.
  unsigned dyn_smem;
  asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_smem));
  const int slot_words = WS * win_h;
  const size_t fr_bytes = (frac_smem_bytes(cols, rows) + 15) & ~(size_t)15;
  const size_t multi_bytes = (size_t)(scratch - smem_raw) + fr_bytes + sizeof(uint32_t) * (size_t)slot_words * (1 + HOP_MAX_PRED);
  const bool multi = (int)gridDim.x == csize && mj.use_gt && win_w <= WS && multi_bytes <= dyn_smem;
  uint32_t* win_slots = multi ? reinterpret_cast<uint32_t*>(scratch + fr_bytes) : nullptr;
  HopGtJob gj;
  gj.org_off = sj.org_off; gj.ref_off = sj.ref_off; gj.org_stride = sj.org_stride; gj.ref_stride = sj.ref_stride;
  gj.cols = cols; gj.rows = rows;
  gj.num_pred = mj.num_pred;
  for (int k = 0; k < HOP_MAX_PRED; k++) gj.amvp[k] = mj.amvp[k];
  gj.use_had = mj.use_had; gj.bit_depth = sj.bit_depth;
  gj.cost = sj.cost; gj.cost.cost_scale = 0;   // :4619
  unsigned pre_mask = 0;
  // nothing below depends on k1_search: it runs before the wait for that grid (the single-call path launches this
  // kernel with programmatic stream serialisation, so it is resident while K1 still works)
  {
    const int16_t* org0 = org_buf + sj.org_off;
    if (ipu.use == 2) {
      for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) s_org[i] = ipu.org[i];
    } else {
      for (int i = threadIdx.x; i < rows * cols; i += blockDim.x)
        s_org[i] = org0[(i / cols) * sj.org_stride + (i % cols)];
    }
    if (multi)
      for (int b = 1; b <= gj.num_pred; b++) {
        int Hx, Hy;
        if (!gt_start_vector(gj, b, &Hx, &Hy)) continue;
        gt_stage_window<WS>(gj, ref_buf, rb, Hx, Hy, w, win_w, win_h, max_val, win_slots + (size_t)b * slot_words);
        pre_mask |= 1u << b;
      }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");   // no-op unless launched as a programmatic dependent
  HOP_STAMP(g_trace_k2, 1);
  const HopSearchResult sr = k1[job_id];
  HopMotionResult* res = &out[job_id];
  const bool go = sr.found == 1 && !(sr.mv.hor == 0 && sr.mv.ver == 0);      // :4603-4611
  if (threadIdx.x == 0 && crank == 0) {
    res->search = sr;
    res->refined = go ? 1 : 0;
    if (!go) {
      res->frac.half.hor = 0; res->frac.half.ver = 0; res->frac.qter.hor = 0; res->frac.qter.ver = 0;
      res->frac.cost = 0; res->frac.cost_half = 0;
    }
    if (!go || !mj.use_gt) {
      res->gt.gt_flag = 0;
      for (int k = 0; k < 4; k++) { res->gt.gt[k].hor = 0; res->gt.gt[k].ver = 0; }
      res->gt.cost = 0; res->gt.mv_int.hor = 0; res->gt.mv_int.ver = 0;
      res->gt.best_index = -1; res->gt.n_candidates = 0;
    }
  }
  if (go) {
    const int16_t* ref_pos = ref_buf + sj.ref_off + sr.mv.hor + (long long)sr.mv.ver * sj.ref_stride;
    gj.ss_cand = sr.mv;                            // pcCU->getSSBestCand()[0]
    if (multi) {
      // fractional source region and the window of the SS vector: both loads of an iteration are issued before
      // either value is used
      int16_t* s_src = reinterpret_cast<int16_t*>(scratch);
      const int src_w = cols + 8, n_src = (rows + 8) * src_w, n_win = win_w * win_h;
      for (int i = threadIdx.x; i < (n_src > n_win ? n_src : n_win); i += blockDim.x) {
        int a = 0, v = 0;
        const int wy = i / win_w, wx = i - wy * win_w;
        if (i < n_src) { const int y = i / src_w, x = i - y * src_w; a = ref_pos[(long long)(y - 4) * sj.ref_stride + (x - 4)]; }
        if (i < n_win) v = ref_buf[gt_window_offset(gj.ref_off, gj.ref_stride, rb, sr.mv.hor, sr.mv.ver, w, wx, wy)];
        if (i < n_src) s_src[i] = (int16_t)a;
        if (i < n_win) win_slots[wy * WS + wx] = gt_window_word(v, max_val);
      }
      pre_mask |= 1u;
    }
    const HopFracResult fr = frac_search_cta(fs, s_org, scratch, ref_pos, sj.ref_stride, cols, rows, sj.bit_depth,
                                             mj.use_had, sj.cost, sr.mv, multi);
    HOP_STAMP(g_trace_k2, 2);   // fractional refinement done
    if (threadIdx.x == 0 && crank == 0) { res->frac = fr; if (!mj.use_gt) res->gt.cost = fr.cost; }
    if (mj.use_gt) {
      gj.threshold = fr.cost;                      // ruiCost coming out of the frac stage (:4769)
      gt_search_cta<WS, CL>(gj, org_buf, ref_buf, smem_raw, &res->gt, true, rb, win_slots, slot_words, pre_mask);
    }
  }
  HOP_STAMP(g_trace_k2, 3);     // GT search done
  if (threadIdx.x == 0 && crank == 0 && done_flag) {
    __threadfence_system();
    *(volatile unsigned*)done_flag = seq;
  }
}

template <int WS, int CFG>
__global__ void __launch_bounds__(GtCfg<CFG>::T, GtCfg<CFG>::B)
k_motion_tail(int n_jobs, const HopMotionJob* __restrict__ jobs, const int16_t* __restrict__ org_buf,
              const int16_t* __restrict__ ref_buf, const HopSearchResult* __restrict__ k1,
              HopMotionResult* __restrict__ out, unsigned* done_flag, unsigned seq, RefBounds rb,
              const __grid_constant__ InlinePu ipu)
{
  motion_tail_body<WS, CFG, false>(n_jobs, jobs, org_buf, ref_buf, k1, out, done_flag, seq, rb, ipu);
}

// cluster forms (single-call latency path for PUs with several Hadamard tiles): launched with a cluster
// dimension of 2, 4 or 8 CTAs per PU
template <int WS>
__global__ void __launch_bounds__(GtCfg<4>::T, 1)
k_motion_tail_cl(int n_jobs, const HopMotionJob* __restrict__ jobs, const int16_t* __restrict__ org_buf,
                 const int16_t* __restrict__ ref_buf, const HopSearchResult* __restrict__ k1,
                 HopMotionResult* __restrict__ out, unsigned* done_flag, unsigned seq, RefBounds rb,
              const __grid_constant__ InlinePu ipu)
{
  motion_tail_body<WS, 0, true>(n_jobs, jobs, org_buf, ref_buf, k1, out, done_flag, seq, rb, ipu);
}

template <int WS>
__global__ void __launch_bounds__(GtCfg<4>::T, 1)
k2_gt_search_cl(int n_jobs, const HopGtJob* __restrict__ jobs, const int16_t* __restrict__ org_buf,
                const int16_t* __restrict__ ref_buf, HopGtResult* __restrict__ out, unsigned* done_flag, unsigned seq,
                RefBounds rb)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cl = cg::this_cluster();
  const int job_id = blockIdx.x / (int)cl.num_blocks();
  if (job_id >= n_jobs) return;
  const HopGtJob job = jobs[job_id];
  gt_search_cta<WS, true>(job, org_buf, ref_buf, smem_raw, &out[job_id], false, rb);
  if (threadIdx.x == 0 && cl.block_rank() == 0 && done_flag) {
    __threadfence_system();
    *(volatile unsigned*)done_flag = seq;
  }
}

// cluster size and CTA size for a PU shape on the latency path
static void cluster_geometry(int cols, int rows, int* csize, int* threads)
{
  // Measured on the single-call path (profiles/r01_latency_trace.txt): a cluster costs ~0.9 us per pass (cluster
  // barrier + remote gathers), one 8x8 tile of all 56 candidates ~0.5 us of an SM's conversion (XU) pipe.  So: one
  // CTA per 8x8 tile (or per four 4x4 tiles) up to the portable cluster size of 8, and no cluster for a PU that
  // is a single 8x8 tile's worth of work (8x8, 8x4, 4x8, 16x4, 4x16).
  constexpr int T = GtCfg<4>::T;
  const int tile = ((rows % 8 == 0) && (cols % 8 == 0)) ? 8 : 4;
  const int per_group = tile == 8 ? 2 * GT_CANDS : GT_CANDS;
  const int ntiles = (cols / tile) * (rows / tile);
  const int work = tile == 8 ? ntiles : ntiles / 4;       // in 8x8 tiles
  const int cs = work >= 8 ? 8 : work >= 4 ? 4 : work >= 2 ? 2 : 1;
  const int per_cta = (ntiles + cs - 1) / cs;
  const int max_groups = T / per_group;
  int groups = per_cta < max_groups ? per_cta : max_groups;
  for (int g = groups - 1; g >= 1; g--)
    if ((per_cta + g - 1) / g <= (per_cta + groups - 1) / groups) groups = g;
  int t = per_group * groups;
  if (t < 64) t = 64;
  if (cs > 1 && t < 448) t = 448;   // every CTA of a cluster runs the whole fractional stage: give it the warps
  *csize = cs; *threads = t;
}

// csize > 1: thread-block cluster; pdl: programmatic dependent launch -- the grid may become resident while the
// preceding kernel of the stream still runs and waits for it at griddepcontrol.wait
template <typename K, typename... Args>
static cudaError_t launch_ex(K kernel, int n, int csize, int threads, size_t smem, cudaStream_t stream, bool pdl, Args... args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n * csize);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (csize > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = csize; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    na++;
  }
  if (pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    na++;
  }
  cfg.attrs = at; cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
template <typename K, typename... Args>
static cudaError_t launch_cluster(K kernel, int n, int csize, int threads, size_t smem, cudaStream_t stream, Args... args)
{
  return launch_ex(kernel, n, csize, threads, smem, stream, false, args...);
}

static size_t motion_smem_bytes(int ws, int max_cols, int max_rows)
{
  const int w = (max_cols < max_rows ? max_cols : max_rows) >> 1;
  const size_t org = ((size_t)max_cols * max_rows + 3) & ~(size_t)3;
  const size_t win = sizeof(uint32_t) * (size_t)ws * (max_rows + 2 * w);
  const size_t fr = frac_smem_bytes(max_cols, max_rows);
  return GT_SHARED_BYTES + gt_div_bytes(max_cols, max_rows) + sizeof(int) * org + (win > fr ? win : fr);
}

// single-PU launch: room for the fractional scratch AND one window per start vector (motion_tail_body: multi), when
// that stays below MOTION_SMEM_CAP; otherwise the shared layout of the batched launches
constexpr size_t MOTION_SMEM_CAP = 200 * 1024;
static size_t motion_smem_single(int ws, int cols, int rows)
{
  const int w = (cols < rows ? cols : rows) >> 1;
  const size_t org = ((size_t)cols * rows + 3) & ~(size_t)3;
  const size_t fr = (frac_smem_bytes(cols, rows) + 15) & ~(size_t)15;
  const size_t multi = GT_SHARED_BYTES + gt_div_bytes(cols, rows) + sizeof(int) * org + fr +
                       sizeof(uint32_t) * (size_t)ws * (rows + 2 * w) * (1 + HOP_MAX_PRED);
  return multi <= MOTION_SMEM_CAP ? multi : motion_smem_bytes(ws, cols, rows);
}

template <int WS, int CFG>
static cudaError_t motion_tail_cfg(int n, const HopMotionJob* d_jobs, const int16_t* d_org, const int16_t* d_ref,
                                   const HopSearchResult* d_k1, HopMotionResult* d_out, int max_cols, int max_rows,
                                   cudaStream_t stream, unsigned* done_flag, unsigned seq, RefBounds rb, const InlinePu& ipu)
{
  static SmemOptIn opt_in;
  {
    cudaError_t e = opt_in.ensure(k_motion_tail<WS, CFG>, (int)MOTION_SMEM_CAP);
    if (e != cudaSuccess) return e;
  }
  const int tile = ((max_rows % 8 == 0) && (max_cols % 8 == 0)) ? 8 : 4;
  const int per_group = tile == 8 ? 2 * GT_CANDS : GT_CANDS;
  const int ntiles = (max_cols / tile) * (max_rows / tile);
  const int max_groups = GtCfg<CFG>::T / per_group;
  int groups = ntiles < max_groups ? ntiles : max_groups;
  for (int g = groups - 1; g >= 1; g--)
    if ((ntiles + g - 1) / g <= (ntiles + groups - 1) / groups) groups = g;
  int threads = per_group * groups;
  if (threads < 64) threads = 64;
  if (n == 1) threads = GtCfg<CFG>::T;   // single PU: all lanes, for the row-per-lane tiles and the fractional stage
  return launch_ex(k_motion_tail<WS, CFG>, n, 1, threads,
                   n == 1 ? motion_smem_single(WS, max_cols, max_rows) : motion_smem_bytes(WS, max_cols, max_rows), stream, ipu.use != 0,
                   n, d_jobs, d_org, d_ref, d_k1, d_out, done_flag, seq, rb, ipu);
}

cudaError_t motion_tail_launch(int n, const HopMotionJob* d_jobs, const int16_t* d_org, const int16_t* d_ref,
                               const HopSearchResult* d_k1, HopMotionResult* d_out, int max_cols, int max_rows,
                               cudaStream_t stream, int* launches, RefBounds rb, unsigned* done_flag, unsigned seq,
                               const InlinePu* inl)
{
  static const InlinePu no_inline = {};
  const InlinePu& ipu = inl ? *inl : no_inline;
  const int win_w = max_cols + (max_cols < max_rows ? max_cols : max_rows);
  if (launches) (*launches)++;
  switch (gt_stride_class(win_w)) {
    case WS_A:  return n == 1 ? motion_tail_cfg<WS_A, 4>(n, d_jobs, d_org, d_ref, d_k1, d_out, max_cols, max_rows, stream, done_flag, seq, rb, ipu)
                              : motion_tail_cfg<WS_A, 5>(n, d_jobs, d_org, d_ref, d_k1, d_out, max_cols, max_rows, stream, done_flag, seq, rb, ipu);
    case WS_B:  return n == 1 ? motion_tail_cfg<WS_B, 4>(n, d_jobs, d_org, d_ref, d_k1, d_out, max_cols, max_rows, stream, done_flag, seq, rb, ipu)
                              : motion_tail_cfg<WS_B, 5>(n, d_jobs, d_org, d_ref, d_k1, d_out, max_cols, max_rows, stream, done_flag, seq, rb, ipu);
    case WS_C:  return n == 1 ? motion_tail_cfg<WS_C, 4>(n, d_jobs, d_org, d_ref, d_k1, d_out, max_cols, max_rows, stream, done_flag, seq, rb, ipu)
                              : motion_tail_cfg<WS_C, 5>(n, d_jobs, d_org, d_ref, d_k1, d_out, max_cols, max_rows, stream, done_flag, seq, rb, ipu);
    default:  return motion_tail_cfg<WS_D, 4>(n, d_jobs, d_org, d_ref, d_k1, d_out, max_cols, max_rows, stream, done_flag, seq, rb, ipu);
  }
}

template <int WS>
static cudaError_t gt_cluster_class(const HopGtJob* d_job, const int16_t* d_org, const int16_t* d_ref, HopGtResult* d_out,
                                    int cols, int rows, int csize, int threads, cudaStream_t stream, unsigned* done_flag, unsigned seq, RefBounds rb)
{
  static SmemOptIn opt_in;
  {
    cudaError_t e = opt_in.ensure(k2_gt_search_cl<WS>, (int)gt_smem_bytes(WS, HOP_MAX_PU, HOP_MAX_PU));
    if (e != cudaSuccess) return e;
  }
  return launch_cluster(k2_gt_search_cl<WS>, 1, csize, threads, gt_smem_bytes(WS, cols, rows), stream,
                        1, d_job, d_org, d_ref, d_out, done_flag, seq, rb);
}

template <int WS>
static cudaError_t motion_cluster_class(const HopMotionJob* d_job, const int16_t* d_org, const int16_t* d_ref,
                                        const HopSearchResult* d_k1, HopMotionResult* d_out, int cols, int rows, int csize,
                                        int threads, cudaStream_t stream, unsigned* done_flag, unsigned seq, RefBounds rb,
                                        const InlinePu& ipu)
{
  static SmemOptIn opt_in;
  {
    cudaError_t e = opt_in.ensure(k_motion_tail_cl<WS>, (int)MOTION_SMEM_CAP);
    if (e != cudaSuccess) return e;
  }
  return launch_ex(k_motion_tail_cl<WS>, 1, csize, threads, motion_smem_single(WS, cols, rows), stream, ipu.use != 0,
                   1, d_job, d_org, d_ref, d_k1, d_out, done_flag, seq, rb, ipu);
}

// Latency path: ONE PU, searched by a cluster of CTAs when it has at least two Hadamard tiles.
// Returns cudaErrorNotSupported when the shape is too small for a cluster (the caller uses the plain kernel).
cudaError_t gt_single_launch(const HopGtJob* d_job, const int16_t* d_org, const int16_t* d_ref, HopGtResult* d_out,
                             int cols, int rows, cudaStream_t stream, int* launches, RefBounds rb, unsigned* done_flag, unsigned seq)
{
  int csize, threads;
  cluster_geometry(cols, rows, &csize, &threads);
  if (csize < 2) return cudaErrorNotSupported;
  if (launches) (*launches)++;
  switch (gt_stride_class(cols + (cols < rows ? cols : rows))) {
    case WS_A:  return gt_cluster_class<WS_A>(d_job, d_org, d_ref, d_out, cols, rows, csize, threads, stream, done_flag, seq, rb);
    case WS_B:  return gt_cluster_class<WS_B>(d_job, d_org, d_ref, d_out, cols, rows, csize, threads, stream, done_flag, seq, rb);
    case WS_C:  return gt_cluster_class<WS_C>(d_job, d_org, d_ref, d_out, cols, rows, csize, threads, stream, done_flag, seq, rb);
    default:  return gt_cluster_class<WS_D>(d_job, d_org, d_ref, d_out, cols, rows, csize, threads, stream, done_flag, seq, rb);
  }
}

cudaError_t motion_single_launch(const HopMotionJob* d_job, const int16_t* d_org, const int16_t* d_ref,
                                 const HopSearchResult* d_k1, HopMotionResult* d_out, int cols, int rows,
                                 cudaStream_t stream, int* launches, RefBounds rb, unsigned* done_flag, unsigned seq,
                                 const InlinePu* inl)
{
  static const InlinePu no_inline = {};
  const InlinePu& ipu = inl ? *inl : no_inline;
  int csize, threads;
  cluster_geometry(cols, rows, &csize, &threads);
  if (csize < 2) return cudaErrorNotSupported;
  if (launches) (*launches)++;
  switch (gt_stride_class(cols + (cols < rows ? cols : rows))) {
    case WS_A:  return motion_cluster_class<WS_A>(d_job, d_org, d_ref, d_k1, d_out, cols, rows, csize, threads, stream, done_flag, seq, rb, ipu);
    case WS_B:  return motion_cluster_class<WS_B>(d_job, d_org, d_ref, d_k1, d_out, cols, rows, csize, threads, stream, done_flag, seq, rb, ipu);
    case WS_C:  return motion_cluster_class<WS_C>(d_job, d_org, d_ref, d_k1, d_out, cols, rows, csize, threads, stream, done_flag, seq, rb, ipu);
    default:  return motion_cluster_class<WS_D>(d_job, d_org, d_ref, d_k1, d_out, cols, rows, csize, threads, stream, done_flag, seq, rb, ipu);
  }
}

}  // namespace hop
