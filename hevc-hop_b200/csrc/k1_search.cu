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
//
// Two paths inside one kernel, both on the GPU, both exact:
//   * byte path (8-bit content): the slice's window is staged ONCE into shared memory as packed bytes
//     (NOT_VALID -> 0), a thread owns 4 adjacent x positions (the four byte alignments of a word, made
//     with funnel shifts) x 4 y positions that share reference rows, and accumulates with
//     VABSDIFF4.U8.ACC (4 pixel-SADs + accumulate per instruction).  The sentinel is handled exactly:
//     while staging, the CTA records per row where the NOT_VALID samples start and checks that they
//     form the monotone staircase the SS reference always has (valid set closed to the left and
//     upwards).  Then a position that passes isValidPattern has no NOT_VALID sample in its footprint,
//     and positions that fail are discarded by the reference too -- so aliasing -1 with 0 is invisible.
//   * generic path (any bit depth, any sentinel layout, widths that are no multiple of 4): one thread
//     per position on the int16 samples.  Taken when the staircase check fails, for 10-bit content, and
//     for non-SS searches over windows that contain NOT_VALID samples.
#include "hop_common.cuh"
#include "hop_internal.h"

namespace hop {

// keys[] is all-ones and done[] all-zero between launches: the last CTA of a PU (a ticket counter decides)
// reads the merged key, writes the result and restores both words, so one search is ONE kernel launch.
__device__ __forceinline__ void k1_write_result(const HopSearchJob& job, unsigned long long key, HopSearchResult* out)
{
  HopSearchResult r;
  r.mv.hor = 0; r.mv.ver = 0;
  if (key == ~0ull) {                                            // TEncSearch.cpp:6356-6360
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
  *out = r;
}

struct K1Geom {            // per-CTA geometry of the slice, identical on host and device
  int nx, ny, y_lo, y_hi;  // window size, position rows [y_lo, y_hi) of this slice
  int step, rused;         // row sub-sampling step and number of block rows used
  int st_rows, st_cols;    // staged rows / columns that the reference itself would touch
  int sw;                  // staged row stride in bytes (multiple of 16)
};

__host__ __device__ inline int k1_sub_shift(const HopSearchJob& job)
{
  int s = (job.fast_enc && job.rows > 8) ? 1 : 0;               // TEncSearch.cpp:6303-6309
  if (!sad_width_has_subshift(job.cols)) s = 0;                 // generic xGetSAD ignores iSubShift
  return s;
}

__host__ __device__ inline K1Geom k1_geom(const HopSearchJob& job, int slice, int slices)
{
  K1Geom g;
  g.nx = job.rng_right - job.rng_left + 1;
  g.ny = job.rng_bottom - job.rng_top + 1;
  const int rows_per = (g.ny + slices - 1) / slices;
  g.y_lo = slice * rows_per;
  g.y_hi = g.y_lo + rows_per < g.ny ? g.y_lo + rows_per : g.ny;
  g.step = 1 << k1_sub_shift(job);
  g.rused = job.rows / g.step;
  const int ny_s = g.y_hi - g.y_lo;
  // last block row read is (rused-1)*step; the SS probes add row rows+4 and column cols+4
  g.st_rows = ny_s + (job.is_ss ? job.rows + 4 : (g.rused - 1) * g.step);
  g.st_cols = g.nx + (job.is_ss ? job.cols + 4 : job.cols - 1);
  const int ngx = (g.nx + 3) / 4;
  g.sw = ((4 * ngx + job.cols + 8) + 15) & ~15;
  return g;
}

__host__ __device__ inline size_t k1_smem_bytes(const HopSearchJob& job, const K1Geom& g)
{
  // [staged window bytes][org bytes][first_invalid + invalid count per staged row][bits tables]
  size_t b = (size_t)g.st_rows * g.sw + 16;
  b += ((size_t)g.rused * job.cols + 15) & ~(size_t)15;
  b += sizeof(int) * 2 * (size_t)g.st_rows;
  b += sizeof(int) * (size_t)(g.nx + (g.y_hi - g.y_lo));
  return b + 64;
}

// ---- generic path: one thread per position on the int16 samples -------------------------------
__device__ __forceinline__ unsigned long long k1_generic(const HopSearchJob& job, const K1Geom& g,
                                                         const int16_t* __restrict__ org,
                                                         const int16_t* __restrict__ ref_y)
{
  const int cols = job.cols, rows = job.rows, stride = job.ref_stride;
  const int sub_shift = k1_sub_shift(job), dist_shift = job.bit_depth - 8;
  unsigned long long best = ~0ull;
  for (int idx = g.y_lo * g.nx + threadIdx.x; idx < g.y_hi * g.nx; idx += blockDim.x) {
    const int py = idx / g.nx, px = idx - py * g.nx;
    const int x = job.rng_left + px, y = job.rng_top + py;
    const int16_t* srch = ref_y + (long long)y * stride + x;
    if (job.is_ss) {
      if ((x >= job.offset_x) && (y > job.offset_y)) continue;              // :6328
      const int16_t* lb = srch + (rows + 4) * stride;                        // isValidPattern
      if (__ldg(lb) == HOP_NOT_VALID || __ldg(lb + cols + 4) == HOP_NOT_VALID) continue;   // :6330
    }
    uint32_t sum = 0;
    for (int r = 0; r < rows; r += g.step) {
      const int16_t* rr = srch + r * stride;
      const int16_t* oo = org + r * job.org_stride;
      for (int c = 0; c < cols; c++) sum = __sad((int)oo[c], (int)__ldg(rr + c), sum);   // org may sit in shared memory (inline path)
    }
    sum <<= sub_shift;
    sum >>= dist_shift;
    sum += mv_cost(job.cost, x, y);                                          // :6336
    const unsigned long long key = ((unsigned long long)sum << 32) | (unsigned)idx;
    best = key < best ? key : best;
  }
  return best;
}

__device__ __forceinline__ unsigned vsad4_acc(unsigned a, unsigned b, unsigned c)
{
  unsigned d;
  asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

constexpr int K1_Q = 4;   // y positions per thread (they share reference rows); 1 on the single-PU latency path

// ---- byte path ----------------------------------------------------------------------------------
// WORDS = cols / 4 (compile time so that the per-row word loop unrolls)
template <int WORDS, int K1_Q>
__device__ __forceinline__ unsigned long long k1_bytes_q(const HopSearchJob& job, const K1Geom& g,
                                                       const unsigned char* __restrict__ s_win,
                                                       const unsigned* __restrict__ s_org,
                                                       const int* __restrict__ s_first_invalid,
                                                       const int* __restrict__ s_bits_x,
                                                       const int* __restrict__ s_bits_y)
{
  const int S = g.step, R = g.rused;
  const int ny_s = g.y_hi - g.y_lo;
  const int ngx = (g.nx + 3) / 4;
  const int sub_shift = k1_sub_shift(job);
  unsigned long long best = ~0ull;
  // task = (x group of 4, parity class, group of K1_Q positions of that class)
  int ntask_y = 0;
  for (int pi = 0; pi < S; pi++) {
    const int t_cnt = (ny_s - pi + S - 1) / S;
    ntask_y += t_cnt > 0 ? (t_cnt + K1_Q - 1) / K1_Q : 0;
  }
  const int ntask = ngx * ntask_y;
  for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
    const int xg = task % ngx;
    int ty = task / ngx, pi = 0;
    {   // locate (parity, group) from the flattened y-task index
      const int t0 = (ny_s + S - 1) / S;
      const int g0 = t0 > 0 ? (t0 + K1_Q - 1) / K1_Q : 0;
      if (ty >= g0) { ty -= g0; pi = 1; }
    }
    const int q0 = pi + S * (ty * K1_Q);          // first position row (relative to the slice)
    unsigned acc[4][K1_Q];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int j = 0; j < K1_Q; j++) acc[a][j] = 0;
    const unsigned char* col0 = s_win + 4 * xg;
    for (int m = 0; m < R + K1_Q - 1; m++) {
      const int srow = q0 + S * m;                // staged row of this reference row
      if (srow >= g.st_rows) break;               // only rows of positions beyond the slice remain
      const unsigned* rw = reinterpret_cast<const unsigned*>(col0 + (size_t)srow * g.sw);
      unsigned sh[4][WORDS];
      unsigned w0 = rw[0];
#pragma unroll
      for (int k = 0; k < WORDS; k++) {
        const unsigned w1 = rw[k + 1];
        sh[0][k] = w0;
        sh[1][k] = __funnelshift_r(w0, w1, 8);
        sh[2][k] = __funnelshift_r(w0, w1, 16);
        sh[3][k] = __funnelshift_r(w0, w1, 24);
        w0 = w1;
      }
#pragma unroll
      for (int j = 0; j < K1_Q; j++) {
        const int r = m - j;                      // block row (in sub-sampled rows) of position j
        if (r < 0 || r >= R) continue;            // uniform across the CTA
        const unsigned* ow = s_org + r * WORDS;
#pragma unroll
        for (int k = 0; k < WORDS; k++) {
          const unsigned o = ow[k];
#pragma unroll
          for (int a = 0; a < 4; a++) acc[a][j] = vsad4_acc(o, sh[a][k], acc[a][j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < K1_Q; j++) {
      const int q = q0 + S * j;                   // position row relative to the slice
      if (q >= ny_s) continue;
      const int py = g.y_lo + q, y = job.rng_top + py;
      const int prow = q + job.rows + 4;          // staged row of the isValidPattern probes
#pragma unroll
      for (int a = 0; a < 4; a++) {
        const int px = 4 * xg + a;
        if (px >= g.nx) continue;
        const int x = job.rng_left + px;
        if (job.is_ss) {
          if ((x >= job.offset_x) && (y > job.offset_y)) continue;          // :6328
          // staircase window: both probes valid <=> the right one lies left of the row's first NOT_VALID
          if (px + job.cols + 4 >= s_first_invalid[prow]) continue;         // :6330
        }
        unsigned sum = acc[a][j] << sub_shift;
        sum += (job.cost.lambda_cost * (unsigned)(s_bits_x[px] + s_bits_y[q])) >> 16;   // :6336
        const unsigned long long key = ((unsigned long long)sum << 32) | (unsigned)(py * g.nx + px);
        best = key < best ? key : best;
      }
    }
  }
  return best;
}

// A batch keeps K1_Q = 4 positions per thread (each staged reference row serves four block rows).  A single PU
// (the in-encoder call) has only a few position rows per slice and leaves most of the CTA idle with that: one
// position row per thread makes the per-thread chain 2-4x shorter and uses 2-4x more lanes.
template <int WORDS>
__device__ __forceinline__ unsigned long long k1_bytes(const HopSearchJob& job, const K1Geom& g,
                                                       const unsigned char* __restrict__ s_win,
                                                       const unsigned* __restrict__ s_org,
                                                       const int* __restrict__ s_first_invalid,
                                                       const int* __restrict__ s_bits_x,
                                                       const int* __restrict__ s_bits_y)
{
  if (gridDim.x == 1) return k1_bytes_q<WORDS, 1>(job, g, s_win, s_org, s_first_invalid, s_bits_x, s_bits_y);
  return k1_bytes_q<WORDS, K1_Q>(job, g, s_win, s_org, s_first_invalid, s_bits_x, s_bits_y);
}

#ifdef HOP_TRACE
__device__ unsigned long long g_trace_k1[HOP_TRACE_SLOTS];
void trace_read_k1(unsigned long long* out) { cudaMemcpyFromSymbol(out, g_trace_k1, sizeof(g_trace_k1)); }
#endif

__global__ void __launch_bounds__(K1_THREADS)
k1_search(int n_jobs, const HopSearchJob* __restrict__ jobs, const int16_t* __restrict__ org_buf,
          const int16_t* __restrict__ ref_buf, unsigned long long* __restrict__ keys,
          unsigned int* __restrict__ done, HopSearchResult* __restrict__ out, int smem_limit,
          unsigned* done_flag, unsigned seq, int job_stride, const __grid_constant__ InlinePu ipu)
{
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ unsigned long long s_red[32];
  __shared__ int s_unclean;
  __shared__ int16_t s_inl_org[INLINE_ORG_SAMPLES];
  HOP_STAMP(g_trace_k1, 0);
  // lets a programmatic dependent (the tail of the fused single-call search) become resident right away; it
  // still waits for this whole grid before it reads the result
  asm volatile("griddepcontrol.launch_dependents;");
  const int job_id = blockIdx.x;
  // job_stride: bytes between jobs (HopSearchJob arrays, or the leading member of HopMotionJob arrays)
  const HopSearchJob job = ipu.use ? ipu.job.search
                                   : *reinterpret_cast<const HopSearchJob*>(reinterpret_cast<const char*>(jobs) + (size_t)job_id * job_stride);
  const K1Geom g = k1_geom(job, blockIdx.y, gridDim.y);
  const bool empty = g.nx <= 0 || g.ny <= 0 || g.y_lo >= g.y_hi;    // degenerate window / slice beyond it
  const int16_t* org = org_buf + job.org_off;
  const int16_t* ref_y = ref_buf + job.ref_off;
  const int cols = job.cols, rows = job.rows;
  if (ipu.use == 2) {                 // original block from the parameter bank (contiguous, stride cols)
    for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) s_inl_org[i] = ipu.org[i];
    org = s_inl_org;
    __syncthreads();
  }

  bool bytes_ok = !empty && job.bit_depth == 8 && (cols % 4) == 0 && cols <= HOP_MAX_PU && rows <= HOP_MAX_PU &&
                  k1_smem_bytes(job, g) <= (size_t)smem_limit;
  unsigned long long best;
  if (bytes_ok) {
    unsigned char* s_win = smem;
    unsigned* s_org = reinterpret_cast<unsigned*>(smem + (((size_t)g.st_rows * g.sw + 15) & ~(size_t)15));
    int* s_first_invalid = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(s_org) +
                                                  (((size_t)g.rused * cols + 15) & ~(size_t)15));
    int* s_cnt_invalid = s_first_invalid + g.st_rows;
    int* s_bits_x = s_cnt_invalid + g.st_rows;
    int* s_bits_y = s_bits_x + g.nx;
    const int ny_s = g.y_hi - g.y_lo;
    if (threadIdx.x == 0) s_unclean = 0;
    for (int i = threadIdx.x; i < g.st_rows; i += blockDim.x) { s_first_invalid[i] = 0x7fffffff; s_cnt_invalid[i] = 0; }
    // bit-length tables of the motion cost (TComRdCost.h:196-199)
    for (int i = threadIdx.x; i < g.nx; i += blockDim.x)
      s_bits_x[i] = (int)component_bits(((job.rng_left + i) << job.cost.cost_scale) - job.cost.pred.hor);
    for (int i = threadIdx.x; i < ny_s; i += blockDim.x)
      s_bits_y[i] = (int)component_bits(((job.rng_top + g.y_lo + i) << job.cost.cost_scale) - job.cost.pred.ver);
    __syncthreads();
    HOP_STAMP(g_trace_k1, 1);   // job arrived, bit tables built
    // original block: the sub-sampled rows, packed to bytes
    int bad = 0;
    for (int i = threadIdx.x; i < g.rused * (cols / 4); i += blockDim.x) {
      const int r = i / (cols / 4), k = i - r * (cols / 4);
      const int16_t* o = org + (r * g.step) * job.org_stride + 4 * k;
      unsigned w = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) { const int v = o[b]; bad |= (v < 0) | (v > 255); w |= (unsigned)(v & 255) << (8 * b); }
      s_org[i] = w;
    }
    // reference window of the slice: rows [rng_top + y_lo, ...), columns [rng_left, ...)
    const int16_t* win0 = ref_y + (long long)(job.rng_top + g.y_lo) * job.ref_stride + job.rng_left;
    const int wpr = g.sw / 4;
    for (int i = threadIdx.x; i < g.st_rows * wpr; i += blockDim.x) {
      const int r = i / wpr, k = i - r * wpr;
      const int16_t* p = win0 + (long long)r * job.ref_stride + 4 * k;
      unsigned w = 0;
      int n_inv = 0, first = 0x7fffffff;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const int c = 4 * k + b;
        int v = 0;
        if (c < g.st_cols) {                                   // never read what the reference would not
          v = __ldg(p + b);
          if (v == HOP_NOT_VALID) { n_inv++; first = first < c ? first : c; v = 0; }
          else bad |= (v < 0) | (v > 255);
        }
        w |= (unsigned)(v & 255) << (8 * b);
      }
      reinterpret_cast<unsigned*>(s_win + (size_t)r * g.sw)[k] = w;
      if (n_inv) { atomicMin(&s_first_invalid[r], first); atomicAdd(&s_cnt_invalid[r], n_inv); }
    }
    if (bad) s_unclean = 1;
    __syncthreads();
    HOP_STAMP(g_trace_k1, 2);   // block and window staged
    // staircase check: NOT_VALID samples form a suffix of every row, starting no later than in the row above
    for (int r = threadIdx.x; r < g.st_rows; r += blockDim.x) {
      const int first = s_first_invalid[r], cnt = s_cnt_invalid[r];
      bool ok = cnt == 0 || cnt == g.st_cols - first;
      if (r > 0 && first > s_first_invalid[r - 1]) ok = false;
      if (!job.is_ss && cnt != 0) ok = false;                  // no validity gate: footprints may hold -1
      if (!ok) s_unclean = 1;
    }
    __syncthreads();
    HOP_STAMP(g_trace_k1, 3);   // staircase checked
    bytes_ok = s_unclean == 0;
    if (bytes_ok) {
      switch (cols / 4) {
        case 1:  best = k1_bytes<1>(job, g, s_win, s_org, s_first_invalid, s_bits_x, s_bits_y); break;
        case 2:  best = k1_bytes<2>(job, g, s_win, s_org, s_first_invalid, s_bits_x, s_bits_y); break;
        case 3:  best = k1_bytes<3>(job, g, s_win, s_org, s_first_invalid, s_bits_x, s_bits_y); break;
        case 4:  best = k1_bytes<4>(job, g, s_win, s_org, s_first_invalid, s_bits_x, s_bits_y); break;
        case 6:  best = k1_bytes<6>(job, g, s_win, s_org, s_first_invalid, s_bits_x, s_bits_y); break;
        case 8:  best = k1_bytes<8>(job, g, s_win, s_org, s_first_invalid, s_bits_x, s_bits_y); break;
        case 12: best = k1_bytes<12>(job, g, s_win, s_org, s_first_invalid, s_bits_x, s_bits_y); break;
        case 16: best = k1_bytes<16>(job, g, s_win, s_org, s_first_invalid, s_bits_x, s_bits_y); break;
        default: bytes_ok = false; break;
      }
    }
  }
  if (!bytes_ok) best = empty ? ~0ull : k1_generic(job, g, org, ref_y);
  HOP_STAMP(g_trace_k1, 4);     // positions searched
  best = block_min_u64(best, s_red);
  HOP_STAMP(g_trace_k1, 5);
  if (threadIdx.x == 0) {
    if (best != ~0ull) atomicMin(&keys[job_id], best);
    __threadfence();
    if (atomicAdd(&done[job_id], 1u) == gridDim.y - 1) {          // last slice of this PU
      __threadfence();
      const unsigned long long key = atomicExch(&keys[job_id], ~0ull);
      done[job_id] = 0;
      k1_write_result(job, key, &out[job_id]);
      HOP_STAMP_ANY(g_trace_k1, 6);   // last slice published the result
      if (done_flag) {               // single-call path: result and flag live in mapped host memory
        __threadfence_system();
        *(volatile unsigned*)done_flag = seq;
      }
    }
  }
}

// =====================================================================================================================
// Batched form (n > 1 PUs of one width): k1_batch<W>, W = cols / 4 at compile time.
//
// Same exact byte path as above, re-balanced for throughput -- the VABSDIFF4 instruction shares its (half-rate) ALU
// pipe with every compare, shift and logic instruction, so everything that is not a SAD is cut down or moved off it:
//   * one kernel per width: the register tile is sized for THAT width (4-wide .. 16-wide PUs take 40-64 registers
//     instead of the 128 of the all-widths kernel), and the launch asks for the shared memory the jobs need, so 4
//     CTAs (32 warps) are resident per SM for narrow PUs instead of 2 (16 warps);
//   * staging reads 4 samples per 8-byte load and packs / classifies them with SIMD-in-register operations (the
//     window is shifted by its misalignment `mis` so that every load is aligned; the 0-3 extra positions on the left
//     are masked by the per-row bound);
//   * the SS gates become ONE compare per position: the causal gate (x >= offset_x && y > offset_y) and the staircase
//     form of isValidPattern both say "x below a bound that depends on the row", precomputed per row;
//   * the motion cost is (lbx[x] + lby[y]) >> 16 with lambda folded into the two tables (u32 arithmetic distributes);
//   * the reference-row loop is split into head / steady / tail, so the steady state has no per-row validity tests;
//   * positions of one row are compared on 32-bit sums (their raster index grows with x), the 64-bit ordered key is
//     touched once per row of four positions.
struct K1bGeom {
  int nx, ny, y_lo, y_hi, step, rused, st_rows, st_cols, mis, ngx, sw;
};
constexpr int K1B_PAD_ROWS = 32;

__host__ __device__ inline K1bGeom k1b_geom(const HopSearchJob& job, int slice, int slices, int mis)
{
  K1bGeom g;
  g.nx = job.rng_right - job.rng_left + 1;
  g.ny = job.rng_bottom - job.rng_top + 1;
  const int rows_per = (g.ny + slices - 1) / slices;
  g.y_lo = slice * rows_per;
  g.y_hi = g.y_lo + rows_per < g.ny ? g.y_lo + rows_per : g.ny;
  g.step = 1 << k1_sub_shift(job);
  g.rused = job.rows / g.step;
  const int ny_s = g.y_hi - g.y_lo;
  g.st_rows = ny_s + (job.is_ss ? job.rows + 4 : (g.rused - 1) * g.step);
  g.st_cols = g.nx + (job.is_ss ? job.cols + 4 : job.cols - 1);
  g.mis = mis;
  g.ngx = (g.nx + mis + 3) / 4;
  g.sw = ((4 * g.ngx + job.cols + 8) + 15) & ~15;
  return g;
}

__host__ __device__ inline size_t k1b_smem_bytes(const HopSearchJob& job, const K1bGeom& g)
{
  // [window bytes][org words][first_invalid, invalid count per staged row][lbx per shifted x][lby, bound per position row]
  // K1B_PAD_ROWS rows of slack behind the window: the padded positions of a parity class's last group read up to
  // 2 * step * (rows per task) rows past it (their sums are never used), and must stay inside the allocation
  size_t b = (size_t)(g.st_rows + K1B_PAD_ROWS) * g.sw + 16;
  b += ((size_t)g.rused * job.cols + 15) & ~(size_t)15;
  b += sizeof(int) * 2 * (size_t)g.st_rows;
  b += sizeof(int) * ((size_t)4 * g.ngx + 2 * (size_t)(g.y_hi - g.y_lo));
  return b + 64;
}

template <int W>
__device__ __forceinline__ void k1b_load_row(const unsigned char* row, unsigned (&sh)[4][W])
{
  const unsigned* rw = reinterpret_cast<const unsigned*>(row);
  unsigned w0 = rw[0];
#pragma unroll
  for (int k = 0; k < W; k++) {
    const unsigned w1 = rw[k + 1];
    sh[0][k] = w0;
    sh[1][k] = __funnelshift_r(w0, w1, 8);
    sh[2][k] = __funnelshift_r(w0, w1, 16);
    sh[3][k] = __funnelshift_r(w0, w1, 24);
    w0 = w1;
  }
}

template <int W>
__device__ __forceinline__ void k1b_accum(const unsigned* __restrict__ ow, const unsigned (&sh)[4][W], unsigned (&acc)[4])
{
#pragma unroll
  for (int k = 0; k < W; k++) {
    const unsigned o = ow[k];
#pragma unroll
    for (int a = 0; a < 4; a++) acc[a] = vsad4_acc(o, sh[a][k], acc[a]);
  }
}

template <int W, int Q>
__device__ __forceinline__ unsigned long long k1b_scan(const HopSearchJob& job, const K1bGeom& g,
                                                       const unsigned char* __restrict__ s_win, const unsigned* __restrict__ s_org,
                                                       const unsigned* __restrict__ s_lbx, const unsigned* __restrict__ s_lby,
                                                       const int* __restrict__ s_bound)
{
  constexpr int IDX_BITS = Q == 8 ? 5 : 4;            // position-in-task index: 4 j + a
  const int S = g.step, R = g.rused, ny_s = g.y_hi - g.y_lo, ngx = g.ngx;
  const int sub_shift = k1_sub_shift(job);
  const int t0 = (ny_s + S - 1) / S, g0 = (t0 + Q - 1) / Q;              // position rows of parity class 0 / their groups
  const int t1 = S > 1 ? (ny_s - 1 + S - 1) / S : 0, g1 = (t1 + Q - 1) / Q;
  const int ntask = ngx * (g0 + g1);
  unsigned long long best = ~0ull;
  // (xg, ty) of the flat task index, advanced without a division per task
  int xg = (int)threadIdx.x % ngx, tyf = (int)threadIdx.x / ngx;
  const int dx = (int)blockDim.x % ngx, dy = (int)blockDim.x / ngx;
  for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
    const int pi = tyf >= g0 ? 1 : 0, ty = tyf - (pi ? g0 : 0);
    const int q0 = pi + S * (ty * Q);             // first position row of the task (relative to the slice)
    unsigned acc[Q][4];
#pragma unroll
    for (int j = 0; j < Q; j++)
#pragma unroll
      for (int a = 0; a < 4; a++) acc[j][a] = 0;
    // running pointers: the reference row of step m and the block row that position row Q-1 meets at step m
    const unsigned char* rowp = s_win + 4 * xg + (size_t)q0 * g.sw;
    const size_t rstep = (size_t)S * g.sw;
    const unsigned* op = s_org - (Q - 1) * W;
    unsigned sh[4][W];
    // reference row m serves block row m - j of position row j; rows past the staged window only feed padded positions
    // head: m = 0 .. Q-2, position rows j <= m
#pragma unroll
    for (int m = 0; m < Q - 1; m++) {
      k1b_load_row<W>(rowp, sh);
#pragma unroll
      for (int j = 0; j <= m; j++) k1b_accum<W>(op + (Q - 1 - j) * W, sh, acc[j]);
      rowp += rstep; op += W;
    }
    // steady state: every position row takes part
    for (int m = Q - 1; m < R; m++) {
      k1b_load_row<W>(rowp, sh);
#pragma unroll
      for (int j = 0; j < Q; j++) k1b_accum<W>(op + (Q - 1 - j) * W, sh, acc[j]);
      rowp += rstep; op += W;
    }
    // tail: m = R .. R+Q-2, position rows j > m - R
#pragma unroll
    for (int t = 0; t < Q - 1; t++) {
      k1b_load_row<W>(rowp, sh);
#pragma unroll
      for (int j = t + 1; j < Q; j++) k1b_accum<W>(op + (Q - 1 - j) * W, sh, acc[j]);
      rowp += rstep; op += W;
    }
    // epilogue.  (acc << sub) + (t >> 16) == hi32(t * 65536) + acc * (1 << sub) runs on the multiply-add pipe.
    // Inside a task the raster index grows with the row and with x, so "first strict minimum" is the minimum of
    // (sum, position-in-task): the pair is packed into one word, sum * 4Q + (4 j + a) < 2^26 (8-bit SADs of at most
    // 64 x 32 x 2 samples), invalid positions become all-ones.
    // the two multipliers go through an opaque move: known powers of two would be strength-reduced to shifts / LEAs,
    // which execute on the ALU pipe the SADs need
    unsigned mul, sixteen, two16;
    asm("mov.u32 %0, %1;" : "=r"(mul) : "r"(1u << sub_shift));
    asm("mov.u32 %0, %1;" : "=r"(sixteen) : "r"(1u << IDX_BITS));
    asm("mov.u32 %0, %1;" : "=r"(two16) : "r"(65536u));
    const unsigned px0 = (unsigned)(4 * xg - g.mis);          // real x of alignment 0; negative (= huge) left of the window
    const uint4 lbx4 = *reinterpret_cast<const uint4*>(s_lbx + 4 * xg);
    const unsigned lbx[4] = {lbx4.x, lbx4.y, lbx4.z, lbx4.w};
    unsigned bs = 0xffffffffu;
    int bidx = 0;
    {
      unsigned k32 = 0xffffffffu;
#pragma unroll
      for (int j = 0; j < Q; j++) {
        const int q = q0 + S * j;
        const unsigned bound = q < ny_s ? (unsigned)s_bound[q] : 0u;     // padded rows: nothing is valid
        const unsigned lby = s_lby[q < ny_s ? q : 0];
#pragma unroll
        for (int a = 0; a < 4; a++) {
          const unsigned sum = __umulhi(lbx[a] + lby, two16) + acc[j][a] * mul;
          const unsigned cand = sum * sixteen + (unsigned)(4 * j + a);
          k32 = min(k32, px0 + a < bound ? cand : 0xffffffffu);          // one unsigned compare decides both gates
        }
      }
      if (k32 != 0xffffffffu) {
        bs = k32 >> IDX_BITS;
        bidx = (g.y_lo + q0 + S * (int)((k32 >> 2) & (unsigned)(Q - 1))) * g.nx + (int)px0 + (int)(k32 & 3u);
      }
    }
    if (bs != 0xffffffffu) {
      const unsigned long long key = ((unsigned long long)bs << 32) | (unsigned)bidx;
      best = key < best ? key : best;
    }
    xg += dx; tyf += dy;
    if (xg >= ngx) { xg -= ngx; tyf++; }
  }
  return best;
}

template <int W> struct K1bCfg { static constexpr int MINB = W <= 4 ? 4 : (W <= 8 ? 3 : 2); };

template <int W>
__global__ void __launch_bounds__(K1_THREADS, K1bCfg<W>::MINB)
k1_batch(int n_jobs, const HopSearchJob* __restrict__ jobs, const int16_t* __restrict__ org_buf,
         const int16_t* __restrict__ ref_buf, unsigned long long* __restrict__ keys,
         unsigned int* __restrict__ done, HopSearchResult* __restrict__ out, int smem_limit, int job_stride)
{
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ unsigned long long s_red[32];
  __shared__ int s_unclean;
  const int job_id = blockIdx.x;
  const HopSearchJob job = *reinterpret_cast<const HopSearchJob*>(reinterpret_cast<const char*>(jobs) + (size_t)job_id * job_stride);
  const int cols = job.cols, rows = job.rows;
  const int16_t* org = org_buf + job.org_off;
  const int16_t* ref_y = ref_buf + job.ref_off;
  // window origin of this slice and its misalignment to 8 bytes (4 samples); rows keep it when the stride is a multiple of 4
  const K1Geom g_old = k1_geom(job, blockIdx.y, gridDim.y);
  const bool empty = g_old.nx <= 0 || g_old.ny <= 0 || g_old.y_lo >= g_old.y_hi;
  const int16_t* win0 = ref_y + (long long)(job.rng_top + g_old.y_lo) * job.ref_stride + job.rng_left;
  const bool vec = (job.ref_stride & 3) == 0;
  const int mis = vec ? (int)((reinterpret_cast<unsigned long long>(win0) >> 1) & 3ull) : 0;
  const K1bGeom g = k1b_geom(job, blockIdx.y, gridDim.y, mis);
  bool bytes_ok = !empty && job.bit_depth == 8 && cols == 4 * W && rows <= HOP_MAX_PU && k1b_smem_bytes(job, g) <= (size_t)smem_limit;
  unsigned long long best = ~0ull;
  if (bytes_ok) {
    unsigned char* s_win = smem;
    unsigned* s_org = reinterpret_cast<unsigned*>(smem + (((size_t)(g.st_rows + K1B_PAD_ROWS) * g.sw + 15) & ~(size_t)15));
    unsigned* s_lbx = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(s_org) + (((size_t)g.rused * cols + 15) & ~(size_t)15));   // 16-byte aligned
    unsigned* s_lby = s_lbx + 4 * g.ngx;
    int* s_bound = reinterpret_cast<int*>(s_lby + (g.y_hi - g.y_lo));
    int* s_first_invalid = s_bound + (g.y_hi - g.y_lo);
    int* s_cnt_invalid = s_first_invalid + g.st_rows;
    const int ny_s = g.y_hi - g.y_lo;
    if (threadIdx.x == 0) s_unclean = 0;
    for (int i = threadIdx.x; i < g.st_rows; i += blockDim.x) { s_first_invalid[i] = 0x7fffffff; s_cnt_invalid[i] = 0; }
    // lambda * bits tables (TComRdCost.h:196-199; (l * (bx + by)) >> 16 == (l*bx + l*by) >> 16 in UInt arithmetic)
    for (int i = threadIdx.x; i < 4 * g.ngx; i += blockDim.x)
      s_lbx[i] = job.cost.lambda_cost * component_bits(((job.rng_left + i - mis) << job.cost.cost_scale) - job.cost.pred.hor);
    for (int i = threadIdx.x; i < ny_s; i += blockDim.x)
      s_lby[i] = job.cost.lambda_cost * component_bits(((job.rng_top + g.y_lo + i) << job.cost.cost_scale) - job.cost.pred.ver);
    __syncthreads();
    int bad = 0;
    for (int i = threadIdx.x; i < g.rused * W; i += blockDim.x) {       // original block: the sub-sampled rows, packed to bytes
      const int r = i / W, k = i - r * W;
      const int16_t* o = org + (r * g.step) * job.org_stride + 4 * k;
      unsigned w = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) { const int v = o[b]; bad |= (v < 0) | (v > 255); w |= (unsigned)(v & 255) << (8 * b); }
      s_org[i] = w;
    }
    // window: staged row r holds, at byte cs, the sample of real column cs - mis; aligned groups of 4 samples are read
    // with one 8-byte load and packed / classified with SIMD-in-register operations
    // A warp takes a row at a time, a lane up to three groups of it (all loads of a row are in flight together).  Groups
    // that lie completely inside the window are read with one aligned 8-byte load; fast path: no sample of the group
    // has a high byte set, i.e. four plain 8-bit samples -- one OR-AND, one compare, one byte permute.  The (at most
    // two) groups per row that straddle a window edge are filled in afterwards, one sample per thread, so that nothing
    // outside the window is ever read.
    const int wpr = g.sw / 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int k_first = (mis + 3) / 4, k_end = (g.st_cols + mis) / 4;      // groups [k_first, k_end) are complete
    // what a lane does with its (up to three) groups of a row does not depend on the row: decided once.  The row loop
    // is straight-line code -- loads, one test, one byte permute and one store per group through the 32-bit
    // shared-memory address -- with ONE branch for the rare groups that hold NOT_VALID or out-of-range samples.
    bool in_g[3], st_g[3];
#pragma unroll
    for (int t = 0; t < 3; t++) {
      const int k = lane + 32 * t;
      in_g[t] = vec && k >= k_first && k < k_end;
      st_g[t] = k < wpr;
    }
    // opaque moves: the lane's shared-memory base and the row pitch stay in registers (the compiler would otherwise
    // rebuild the shared-window address from the CTA's special registers in front of every store)
    unsigned win_sa, sw_r;
    asm volatile("mov.b32 %0, %1;" : "=r"(win_sa) : "r"((unsigned)__cvta_generic_to_shared(s_win) + 4u * (unsigned)lane));
    asm volatile("mov.b32 %0, %1;" : "=r"(sw_r) : "r"((unsigned)g.sw));
    for (int r = warp; r < g.st_rows; r += nwarp) {
      const uint2* prow = reinterpret_cast<const uint2*>(win0 + (long long)r * job.ref_stride - mis) + lane;
      uint2 u[3];
#pragma unroll
      for (int t = 0; t < 3; t++) {
        u[t] = make_uint2(0u, 0u);
        if (in_g[t]) u[t] = __ldg(prow + 32 * t);
      }
      unsigned w[3];
      bool slow = false;
#pragma unroll
      for (int t = 0; t < 3; t++) {
        w[t] = __byte_perm(u[t].x, u[t].y, 0x6420);                      // four plain 8-bit samples (zeros where nothing was read)
        slow |= ((u[t].x | u[t].y) & 0xff00ff00u) != 0;
      }
      if (slow) {
#pragma unroll
        for (int t = 0; t < 3; t++) {
          if (((u[t].x | u[t].y) & 0xff00ff00u) == 0) continue;
          const int k = lane + 32 * t;
          const unsigned e0 = __vcmpeq2(u[t].x, 0xffffffffu), e1 = __vcmpeq2(u[t].y, 0xffffffffu);   // 0xffff per NOT_VALID sample
          bad |= (((u[t].x & ~e0) | (u[t].y & ~e1)) & 0xff00ff00u) != 0;
          w[t] = __byte_perm(u[t].x & ~e0, u[t].y & ~e1, 0x6420);
          const int n_inv = __popc(e0 & 0x00010001u) + __popc(e1 & 0x00010001u);
          if (n_inv) {
            const int first = 4 * k - mis + ((e0 & 0xffffu) ? 0 : (e0 ? 1 : ((e1 & 0xffffu) ? 2 : 3)));
            atomicMin(&s_first_invalid[r], first);
            atomicAdd(&s_cnt_invalid[r], n_inv);
          }
        }
      }
      const unsigned row_sa = win_sa + (unsigned)r * sw_r;
      // edge and outside groups: zero for now
      if (st_g[0]) asm volatile("st.shared.u32 [%0], %1;" :: "r"(row_sa), "r"(w[0]) : "memory");
      if (st_g[1]) asm volatile("st.shared.u32 [%0+128], %1;" :: "r"(row_sa), "r"(w[1]) : "memory");
      if (st_g[2]) asm volatile("st.shared.u32 [%0+256], %1;" :: "r"(row_sa), "r"(w[2]) : "memory");
    }
    __syncthreads();
    {
      // samples of the window that no complete group covers: columns [0, 4 k_first - mis) and [4 k_end - mis, st_cols)
      // (every column when the rows are not 8-byte aligned)
      const int lo_n = vec ? min(4 * k_first - mis, g.st_cols) : g.st_cols;
      const int hi_0 = vec ? max(4 * k_end - mis, lo_n) : g.st_cols;
      const int per_row = lo_n + (g.st_cols - hi_0);
      for (int i = threadIdx.x; i < g.st_rows * per_row; i += blockDim.x) {
        const int r = i / per_row, e = i - r * per_row;
        const int c = e < lo_n ? e : hi_0 + (e - lo_n);
        int v = __ldg(win0 + (long long)r * job.ref_stride + c);
        if (v == HOP_NOT_VALID) { atomicMin(&s_first_invalid[r], c); atomicAdd(&s_cnt_invalid[r], 1); v = 0; }
        else bad |= (v < 0) | (v > 255);
        s_win[(size_t)r * g.sw + c + mis] = (unsigned char)v;
      }
    }
    if (bad) s_unclean = 1;
    __syncthreads();
    // staircase check (see the header of this file) and the per-row bound of the valid positions
    for (int r = threadIdx.x; r < g.st_rows; r += blockDim.x) {
      const int first = s_first_invalid[r], cnt = s_cnt_invalid[r];
      bool ok = cnt == 0 || cnt == g.st_cols - first;
      if (r > 0 && first > s_first_invalid[r - 1]) ok = false;
      if (!job.is_ss && cnt != 0) ok = false;
      if (!ok) s_unclean = 1;
    }
    for (int q = threadIdx.x; q < ny_s; q += blockDim.x) {
      int b = g.nx;
      if (job.is_ss) {
        const int y = job.rng_top + g.y_lo + q;
        const int fi = s_first_invalid[q + rows + 4];                    // row of the isValidPattern probes (:6330)
        if (fi != 0x7fffffff) b = min(b, fi - cols - 4);                 // valid iff px + cols + 4 < first NOT_VALID of that row
        if (y > job.offset_y) b = min(b, job.offset_x - job.rng_left);   // causal gate (:6328): x < offset_x
      }
      s_bound[q] = max(b, 0);                                            // valid iff 0 <= px < bound (one unsigned compare)
    }
    __syncthreads();
    bytes_ok = s_unclean == 0;
    // (eight position rows per thread, k1b_scan<W, 8>, were measured for the narrow PUs: no gain over four -- the
    //  larger register tile costs the fourth resident CTA)
    if (bytes_ok) best = k1b_scan<W, 4>(job, g, s_win, s_org, s_lbx, s_lby, s_bound);
  }
  if (!bytes_ok) best = empty ? ~0ull : k1_generic(job, g_old, org, ref_y);
  best = block_min_u64(best, s_red);
  if (threadIdx.x == 0) {
    if (best != ~0ull) atomicMin(&keys[job_id], best);
    __threadfence();
    if (atomicAdd(&done[job_id], 1u) == gridDim.y - 1) {          // last slice of this PU
      __threadfence();
      const unsigned long long key = atomicExch(&keys[job_id], ~0ull);
      done[job_id] = 0;
      k1_write_result(job, key, &out[job_id]);
    }
  }
}

template <int W>
static cudaError_t k1_batch_launch(int n, const HopSearchJob* d_jobs, const int16_t* d_org, const int16_t* d_ref, HopSearchResult* d_out,
                                   unsigned long long* d_keys, unsigned int* d_done, int slices, int smem_bytes, cudaStream_t stream, int job_stride)
{
  static SmemOptIn opt_in;
  cudaError_t e = opt_in.ensure(k1_batch<W>, 160 * 1024);
  if (e != cudaSuccess) return e;
  k1_batch<W><<<dim3(n, slices), K1_THREADS, smem_bytes, stream>>>(n, d_jobs, d_org, d_ref, d_keys, d_done, d_out, smem_bytes, job_stride);
  return cudaGetLastError();
}

// Shared memory k1_batch wants for `job` cut into `slices` (worst misalignment), host side helper.
size_t search_batch_smem_bytes(const HopSearchJob& job, int slices)
{
  size_t worst = 0;
  for (int s = 0; s < slices; s++) {
    const K1bGeom g = k1b_geom(job, s, slices, 3);
    if (g.y_lo >= g.y_hi) continue;
    const size_t b = k1b_smem_bytes(job, g);
    if (b > worst) worst = b;
  }
  return worst;
}

cudaError_t search_launch(int n, const HopSearchJob* d_jobs, const int16_t* d_org, const int16_t* d_ref,
                          HopSearchResult* d_out, unsigned long long* d_keys, unsigned int* d_done, int slices,
                          int smem_bytes, cudaStream_t stream, int* launches, unsigned* done_flag, unsigned seq,
                          int job_stride, const InlinePu* inl, int words_hint)
{
  if (words_hint > 0 && !inl && !done_flag && n > 1) {
    // a batch of PUs of one width (8-bit content): the per-width throughput kernel
    const int js = job_stride ? job_stride : (int)sizeof(HopSearchJob);
    if (slices < 1) slices = 1;
    if (slices > K1_MAX_SLICES) slices = K1_MAX_SLICES;
    if (smem_bytes > 160 * 1024) smem_bytes = 160 * 1024;
    if (smem_bytes < 1024) smem_bytes = 1024;
    cudaError_t e = cudaErrorInvalidValue;
    switch (words_hint) {
      case 1:  e = k1_batch_launch<1>(n, d_jobs, d_org, d_ref, d_out, d_keys, d_done, slices, smem_bytes, stream, js); break;
      case 2:  e = k1_batch_launch<2>(n, d_jobs, d_org, d_ref, d_out, d_keys, d_done, slices, smem_bytes, stream, js); break;
      case 3:  e = k1_batch_launch<3>(n, d_jobs, d_org, d_ref, d_out, d_keys, d_done, slices, smem_bytes, stream, js); break;
      case 4:  e = k1_batch_launch<4>(n, d_jobs, d_org, d_ref, d_out, d_keys, d_done, slices, smem_bytes, stream, js); break;
      case 6:  e = k1_batch_launch<6>(n, d_jobs, d_org, d_ref, d_out, d_keys, d_done, slices, smem_bytes, stream, js); break;
      case 8:  e = k1_batch_launch<8>(n, d_jobs, d_org, d_ref, d_out, d_keys, d_done, slices, smem_bytes, stream, js); break;
      case 12: e = k1_batch_launch<12>(n, d_jobs, d_org, d_ref, d_out, d_keys, d_done, slices, smem_bytes, stream, js); break;
      case 16: e = k1_batch_launch<16>(n, d_jobs, d_org, d_ref, d_out, d_keys, d_done, slices, smem_bytes, stream, js); break;
      default: break;
    }
    if (e != cudaErrorInvalidValue) { if (launches) *launches += 1; return e; }
  }
  static SmemOptIn opt_in;
  static const InlinePu no_inline = {};
  const int smem_max = 160 * 1024;
  {
    cudaError_t e = opt_in.ensure(k1_search, smem_max);
    if (e != cudaSuccess) return e;
  }
  if (slices < 1) slices = 1;
  if (slices > K1_MAX_SLICES) slices = K1_MAX_SLICES;
  if (smem_bytes > smem_max) smem_bytes = smem_max;
  if (smem_bytes < 1024) smem_bytes = 1024;
  k1_search<<<dim3(n, slices), K1_THREADS, smem_bytes, stream>>>(n, d_jobs, d_org, d_ref, d_keys, d_done, d_out, smem_bytes, done_flag, seq,
                                                                      job_stride ? job_stride : (int)sizeof(HopSearchJob), inl ? *inl : no_inline);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

// Shared memory the byte path wants for `job` when its window is cut into `slices` (host side helper).
size_t search_smem_bytes(const HopSearchJob& job, int slices)
{
  size_t worst = 0;
  for (int s = 0; s < slices; s++) {
    const K1Geom g = k1_geom(job, s, slices);
    if (g.y_lo >= g.y_hi) continue;
    const size_t b = k1_smem_bytes(job, g);
    if (b > worst) worst = b;
  }
  return worst;
}

}  // namespace hop
