// hop_internal.h -- declarations shared by the translation units of libhopgpu (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "../../include/hop_gpu.h"

namespace hop {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE property of a kernel, and one process may hold
// contexts on several GPUs (hop_ctx_create(device)): one bit per device ordinal, one object per kernel instantiation.
// Setting the attribute twice is harmless, so a race between two host threads only repeats the call.
struct SmemOptIn {
  std::atomic<unsigned long long> done{0};
  template <typename K>
  cudaError_t ensure(K kernel, int bytes)
  {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
  }
};

// addressable sample offsets of the reference buffer handed to a kernel (relative to its base pointer)
struct RefBounds { long long lo, hi; };
constexpr RefBounds REF_UNBOUNDED = {-(1ll << 62), (1ll << 62)};

constexpr int GT_CANDS   = 56;    // affine corner sets per diamond pass
constexpr int GT_THREADS = 448;   // sweep CTA size upper bound: 56 candidates x 2 lanes x 4 tile groups (72 registers, 2 CTAs per SM)
constexpr int K1_THREADS = 256;
constexpr int K1_MAX_SLICES = 32; // CTAs cooperating on one PU's search window

// K2
void        gt_build_offset_table(int8_t table[GT_CANDS][8], int* count);
cudaError_t gt_upload_offset_table(const int8_t table[GT_CANDS][8]);
cudaError_t gt_launch(int n, const HopGtJob* d_jobs, const int16_t* d_org, const int16_t* d_ref,
                      HopGtResult* d_out, int max_cols, int max_rows, cudaStream_t stream, int* launches,
                      RefBounds rb = REF_UNBOUNDED, unsigned* done_flag = nullptr, unsigned seq = 0);
// exhaustive sweep (reference mode IT_GT_SEARCH 1, N = 2): 85^2 - 25 parallelogram offset patterns
constexpr int SWEEP_CANDS = 7200;
struct SweepCand { int8_t o[8]; uint32_t flat; };   // x0,y0,...,x3,y3 in [-2,2]; flat 8-deep loop index
void        sweep_build_table(SweepCand* table, int* count);
cudaError_t sweep_upload_table(const SweepCand* table);
// Peer-memory exchange of the sharded sweep: per rank, device pointers (valid on THIS GPU, peers' memory mapped over
// NVLink) to the merge words [2][max_pus] every rank keeps in its own HBM, and to its arrival counter.  world == 0: off.
constexpr int SWEEP_MAX_RANKS = 8;
struct SweepXchg {
  unsigned long long* gkeys[SWEEP_MAX_RANKS];
  unsigned int*       gcounts[SWEEP_MAX_RANKS];
  unsigned int*       arrived[SWEEP_MAX_RANKS];
  unsigned int*       pu_done;      // local: tickets per PU, zero between sweeps
  unsigned int*       grid_done;    // local: ticket of the grid
  int world, rank, max_pus;
  unsigned parity;                  // sweep number & 1: which half of the merge words this sweep uses
};
cudaError_t sweep_init_launch(int n, unsigned long long* d_keys, unsigned int* d_counts, cudaStream_t stream, int* launches);
cudaError_t sweep_keys_launch(int n, const HopGtJob* d_jobs, const int16_t* d_org, const int16_t* d_ref,
                              int max_cols, int max_rows, int cand_begin, int cand_end, int chunks,
                              unsigned long long* d_keys, unsigned int* d_counts, cudaStream_t stream, int* launches,
                              RefBounds rb = REF_UNBOUNDED, const SweepXchg* xchg = nullptr);
cudaError_t sweep_finalize_x_launch(int n, const HopGtJob* d_jobs, unsigned long long* d_gkeys, unsigned int* d_gcounts,
                                    const unsigned int* d_arrived, unsigned int target, HopGtResult* d_out, cudaStream_t stream, int* launches);
cudaError_t sweep_finalize_launch(int n, const HopGtJob* d_jobs, const unsigned long long* d_keys,
                                  const unsigned int* d_counts, HopGtResult* d_out, cudaStream_t stream, int* launches);
struct InlinePu;
// latency path: one PU searched by a thread-block cluster; cudaErrorNotSupported for single-tile shapes
cudaError_t gt_single_launch(const HopGtJob* d_job, const int16_t* d_org, const int16_t* d_ref, HopGtResult* d_out,
                             int cols, int rows, cudaStream_t stream, int* launches, RefBounds rb, unsigned* done_flag, unsigned seq);
cudaError_t motion_single_launch(const HopMotionJob* d_job, const int16_t* d_org, const int16_t* d_ref,
                                 const HopSearchResult* d_k1, HopMotionResult* d_out, int cols, int rows,
                                 cudaStream_t stream, int* launches, RefBounds rb, unsigned* done_flag, unsigned seq,
                                 const InlinePu* inl = nullptr);
// Single-call latency path: the job and (for PUs up to 512 samples) the original block travel as kernel
// parameters -- the constant bank is filled by the launch itself, so the kernels start without a dependent
// read of host memory over PCIe (measured 2.4-5 us per read, twice per kernel).  use == 0: read jobs[] / org[].
constexpr int INLINE_ORG_SAMPLES = 512;
struct InlinePu {
  int use;                            // 0 = off, 1 = job inline, 2 = job and original block inline
  int pad;
  HopMotionJob job;                   // .search is the K1 job; org_off 0, org_stride cols when the block is inline
  int16_t org[INLINE_ORG_SAMPLES];
};
// K1
cudaError_t search_launch(int n, const HopSearchJob* d_jobs, const int16_t* d_org, const int16_t* d_ref,
                          HopSearchResult* d_out, unsigned long long* d_keys, unsigned int* d_done, int slices,
                          int smem_bytes, cudaStream_t stream, int* launches,
                          unsigned* done_flag = nullptr, unsigned seq = 0, int job_stride = 0,
                          const InlinePu* inl = nullptr, int words_hint = 0);
// words_hint = cols / 4 when every job of a batch (n > 1) has that width and 8-bit content: the per-width throughput
// kernel k1_batch<W> runs, with the shared memory search_batch_smem_bytes() says; 0 = the all-widths kernel
size_t      search_batch_smem_bytes(const HopSearchJob& job, int slices);
// K5 + fused motion search
cudaError_t frac_launch(int n, const HopFracJob* d_jobs, const int16_t* d_org, const int16_t* d_ref,
                        HopFracResult* d_out, int max_cols, int max_rows, cudaStream_t stream, int* launches);
cudaError_t motion_tail_launch(int n, const HopMotionJob* d_jobs, const int16_t* d_org, const int16_t* d_ref,
                               const HopSearchResult* d_k1, HopMotionResult* d_out, int max_cols, int max_rows,
                               cudaStream_t stream, int* launches, RefBounds rb = REF_UNBOUNDED,
                               unsigned* done_flag = nullptr, unsigned seq = 0, const InlinePu* inl = nullptr);
size_t      search_smem_bytes(const HopSearchJob& job, int slices);
constexpr int K1_DEFAULT_SMEM = 96 * 1024;   // byte-path budget when the job shapes are not known on the host
// K3
cudaError_t dist_launch(int n, const HopDistJob* d_jobs, const int16_t* d_org, const int16_t* d_cur,
                        uint32_t* d_out, cudaStream_t stream, int* launches);
// K6
cudaError_t predict_launch(int n, const HopPredJob* d_jobs, const int16_t* d_org, const int16_t* d_ref, int16_t* d_dst,
                           HopPredResult* d_out, int max_cols, int max_rows, bool any_gt, cudaStream_t stream, int* launches,
                           RefBounds rb = REF_UNBOUNDED);
// K7
cudaError_t intra_launch(int n, const HopIntraJob* d_jobs, const int16_t* d_org, const int32_t* d_refs, uint32_t* d_out,
                         int max_size, cudaStream_t stream, int* launches);
// K4
cudaError_t ref_fill_launch(int16_t* d_plane, size_t samples, int value, cudaStream_t stream, int* launches);
cudaError_t ref_extend_launch(int16_t* d_origin, int stride, int pic_w, int pic_h, int margin,
                              int x, int y, int w, int h, cudaStream_t stream, int* launches);
// probes
cudaError_t probe_launch(int what, int blocks, int threads, int iters, unsigned* d_sink,
                         cudaStream_t stream, double* lane_ops_per_thread_iter);


// ---- latency tracing (debug builds only: tools/build_variant.sh trace -DHOP_TRACE) ---------------------------
// HOP_STAMP(i) records %globaltimer (ns) of one thread at a phase boundary of the single-call path; every
// translation unit keeps its own table, read back by trace_read_k1 / trace_read_k2 (tools/latency_trace.py).
#ifdef HOP_TRACE
#define HOP_TRACE_SLOTS 64
#define HOP_STAMP_ANY(tbl, i) do { if (threadIdx.x == 0) { unsigned long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); (tbl)[(i)] = t_; } } while (0)
#define HOP_STAMP(tbl, i) do { if (blockIdx.x == 0 && blockIdx.y == 0) HOP_STAMP_ANY(tbl, i); } while (0)
void trace_read_k1(unsigned long long* out);
void trace_read_k2(unsigned long long* out);
#else
#define HOP_STAMP_ANY(tbl, i) do {} while (0)
#define HOP_STAMP(tbl, i) do {} while (0)
#endif

}  // namespace hop
