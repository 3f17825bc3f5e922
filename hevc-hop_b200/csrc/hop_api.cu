// hop_api.cu -- the C ABI of libhopgpu (include/hop_gpu.h): contexts, scratch management, the host
// (copy-in / run / copy-out) and device (HBM-resident) entry points, the SS reference mirror and the
// ALU probes.  No CPU fallback exists anywhere in this file: every path ends in a kernel launch or in
// an error status.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <new>
#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define HOP_CPU_RELAX() _mm_pause()
#else
#define HOP_CPU_RELAX() ((void)0)
#endif

#include "hop_internal.h"

using namespace hop;

namespace {

thread_local char g_err[512] = "";

int fail(int status, const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return status;
}

#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(HOP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),  \
                  __FILE__, __LINE__);                                                    \
  } while (0)

struct Scratch {
  void*  p = nullptr;
  size_t cap = 0;
};

}  // namespace

#ifdef HOP_TRACE
#include <time.h>
static unsigned long long g_host_trace[8];
static inline unsigned long long host_ns() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (unsigned long long)t.tv_sec * 1000000000ull + t.tv_nsec; }
#define HOST_STAMP(i) (g_host_trace[(i)] = host_ns())
#else
#define HOST_STAMP(i) ((void)0)
#endif

// Single-PU (in-encoder) searches run through SLOTS.  A slot owns a piece of mapped pinned host memory
// [job | original block | result | completion flag] (the GPU reads the block from it and writes result + flag back,
// zero-copy), the K1 merge words of its launch and, except slot 0, a side stream.  Slot 0 serves the synchronous
// call on the context stream; slots 1.. hold SPECULATIVE searches (hop_motion_search_prefetch): the complete
// request (job, block, version of the SS mirror) is the cache key, so a later hop_motion_search_batch(n = 1) that
// asks for exactly that search picks the finished result up instead of launching.
constexpr int PU_SLOTS = 1 + HOP_PREFETCH_SLOTS;
struct PuSlot {
  unsigned char* h = nullptr;       // mapped pinned host memory
  unsigned char* d = nullptr;       // its device alias
  unsigned       seq = 0;           // sequence number of the slot's last launch == flag value once it has finished
  bool           cached = false;    // holds a speculative search that nobody has consumed yet
  uint64_t       ref_version = 0;   // version of the SS mirror the search ran against
  uint64_t       synced_version = ~0ull;   // last mirror version the side stream has been ordered behind
  HopMotionJob   key;               // normalised job (org_off 0, org_stride cols, unused AMVP entries zero)
  cudaStream_t   stream = nullptr;
  unsigned long long* d_key = nullptr;     // K1 merge word, ticket counter and integer result of this slot's launch
  unsigned int*       d_done = nullptr;
  HopSearchResult*    d_k1 = nullptr;
};

struct HopCtx {
  int          device = 0;
  cudaStream_t stream = nullptr;
  int          sm_count = 0;
  uint64_t     launches = 0;
  Scratch      jobs, org, ref, out, keys, done, sweep_keys, k1res, sink;
  PuSlot       slots_pu[PU_SLOTS];
  unsigned char* pin_base_h = nullptr;     // one allocation for all slots
  unsigned char* slot_words = nullptr;     // device: PU_SLOTS x 64 B of merge words / K1 results
  int          next_pu_slot = 1;
  uint64_t     ref_version = 0;            // bumped by every change of the SS mirror
  cudaEvent_t  ref_event = nullptr;        // recorded on `stream` after the latest mirror change ...
  uint64_t     ref_event_version = ~0ull;  // ... of this version
  // sharded sweep: this rank's exchange block (IPC-exported), the peers' blocks mapped here, local merge words
  unsigned char* xch_block = nullptr;
  unsigned char* xch_peer[SWEEP_MAX_RANKS] = {nullptr};
  unsigned char* xch_local = nullptr;      // keys[max_pus] | counts[max_pus] | pu_done[max_pus] | grid_done
  int          xch_max_pus = 0, xch_world = 0, xch_rank = 0;
  unsigned     xch_epoch = 0;
  HopCtxStats  stats = {};
  double       spin_timeout_s = 20.0;      // HOP_TIMEOUT_MS: a kernel that never publishes its flag is an error, not a hang
  bool           use_clusters = true;   // HOP_CLUSTERS=0 turns the cluster form of the latency path off
  // asynchronous batches: a copy stream and a ring of scratch sets
  cudaStream_t copy_stream = nullptr;
  struct Slot { Scratch jobs, org, ref, out; cudaEvent_t h2d = nullptr, done = nullptr; bool busy = false; };
  Slot         slots[HOP_ASYNC_SLOTS];
  unsigned     next_slot = 0;
  // SS reference mirror
  int16_t*     plane = nullptr;
  int          pic_w = 0, pic_h = 0, margin = 0, stride = 0;
  bool         plane_valid = false;
};

namespace {

int ensure(HopCtx* ctx, Scratch& s, size_t bytes)
{
  if (bytes <= s.cap) return HOP_OK;
  if (s.p) CU(cudaFree(s.p));
  s.p = nullptr; s.cap = 0;
  size_t cap = bytes + bytes / 4 + 4096;
  cudaError_t e = cudaMalloc(&s.p, cap);
  if (e != cudaSuccess) return fail(HOP_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", cap, cudaGetErrorString(e));
  s.cap = cap;
  (void)ctx;
  return HOP_OK;
}

int bind(HopCtx* ctx)
{
  if (!ctx) return fail(HOP_ERR_ARG, "NULL context");
  CU(cudaSetDevice(ctx->device));
  return HOP_OK;
}

int shape_ok(int cols, int rows)
{
  auto ok = [](int v) { return v == 4 || v == 8 || v == 12 || v == 16 || v == 24 || v == 32 || v == 48 || v == 64; };
  return ok(cols) && ok(rows) && !(cols == 4 && rows == 4);
}

// __device__ tables are per device; contexts may be created from several host threads
std::mutex g_table_mutex;
bool g_table_ready[64] = {false};
bool g_sweep_ready[64] = {false};

int sweep_table_ready(HopCtx* ctx)
{
  std::lock_guard<std::mutex> lock(g_table_mutex);
  if (ctx->device < 64 && g_sweep_ready[ctx->device]) return HOP_OK;
  SweepCand* table = new (std::nothrow) SweepCand[SWEEP_CANDS];
  if (!table) return fail(HOP_ERR_NOMEM, "out of host memory");
  int count = 0;
  sweep_build_table(table, &count);
  if (count != SWEEP_CANDS) { delete[] table; return fail(HOP_ERR_STATE, "sweep table has %d entries, expected %d", count, SWEEP_CANDS); }
  cudaError_t e = sweep_upload_table(table);
  delete[] table;
  if (e != cudaSuccess) return fail(HOP_ERR_CUDA, "sweep table upload: %s", cudaGetErrorString(e));
  if (ctx->device < 64) g_sweep_ready[ctx->device] = true;
  return HOP_OK;
}

}  // namespace

extern "C" {

int hop_abi_version(void) { return HOP_ABI_VERSION; }
const char* hop_last_error(void) { return g_err; }

int hop_device_count(void)
{
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return fail(HOP_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  return n;
}

int hop_shape_supported(int cols, int rows) { return shape_ok(cols, rows); }

int hop_ctx_create(int device, HopCtx** out)
{
  if (!out) return fail(HOP_ERR_ARG, "out == NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return fail(HOP_ERR_CUDA, "no CUDA device (%s); libhopgpu has no CPU fallback",
                e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(HOP_ERR_ARG, "device %d out of range (%d devices)", device, n);
  CU(cudaSetDevice(device));
  // three attributes instead of cudaGetDeviceProperties (which queries everything and costs tens of ms at start-up)
  int cc_major = 0, cc_minor = 0, sm_count = 0;
  CU(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
  CU(cudaDeviceGetAttribute(&cc_minor, cudaDevAttrComputeCapabilityMinor, device));
  CU(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
  if (cc_major != 10)
    return fail(HOP_ERR_CUDA, "device %d is sm_%d%d; libhopgpu carries sm_100a code only", device, cc_major, cc_minor);
  HopCtx* ctx = new (std::nothrow) HopCtx();
  if (!ctx) return fail(HOP_ERR_NOMEM, "out of host memory");
  ctx->device = device;
  ctx->sm_count = sm_count;
  { const char* e = getenv("HOP_CLUSTERS"); ctx->use_clusters = !(e && e[0] == '0'); }
  { const char* e = getenv("HOP_TIMEOUT_MS"); if (e && atof(e) > 0) ctx->spin_timeout_s = atof(e) * 1e-3; }
  e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete ctx; return fail(HOP_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
  std::lock_guard<std::mutex> lock(g_table_mutex);
  if (device >= 64 || !g_table_ready[device]) {
    int8_t table[GT_CANDS][8];
    int count = 0;
    gt_build_offset_table(table, &count);
    if (count != GT_CANDS) { delete ctx; return fail(HOP_ERR_STATE, "GT offset table has %d entries, expected %d", count, GT_CANDS); }
    e = gt_upload_offset_table(table);
    if (e != cudaSuccess) { delete ctx; return fail(HOP_ERR_CUDA, "offset table upload: %s", cudaGetErrorString(e)); }
    if (device < 64) g_table_ready[device] = true;
  }
  *out = ctx;
  return HOP_OK;
}

void hop_ctx_destroy(HopCtx* ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  Scratch* all[] = {&ctx->jobs, &ctx->org, &ctx->ref, &ctx->out, &ctx->keys, &ctx->done, &ctx->sweep_keys, &ctx->k1res, &ctx->sink};
  for (Scratch* s : all) if (s->p) cudaFree(s->p);
  for (int i = 1; i < PU_SLOTS; i++)
    if (ctx->slots_pu[i].stream) { cudaStreamSynchronize(ctx->slots_pu[i].stream); cudaStreamDestroy(ctx->slots_pu[i].stream); }
  if (ctx->ref_event) cudaEventDestroy(ctx->ref_event);
  if (ctx->pin_base_h) cudaFreeHost(ctx->pin_base_h);
  if (ctx->slot_words) cudaFree(ctx->slot_words);
  for (int r = 0; r < SWEEP_MAX_RANKS; r++)
    if (ctx->xch_peer[r] && ctx->xch_peer[r] != ctx->xch_block) cudaIpcCloseMemHandle(ctx->xch_peer[r]);
  if (ctx->xch_block) cudaFree(ctx->xch_block);
  if (ctx->xch_local) cudaFree(ctx->xch_local);
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  for (auto& sl : ctx->slots) {
    Scratch* ss[] = {&sl.jobs, &sl.org, &sl.ref, &sl.out};
    for (Scratch* q : ss) if (q->p) cudaFree(q->p);
    if (sl.h2d) cudaEventDestroy(sl.h2d);
    if (sl.done) cudaEventDestroy(sl.done);
  }
  if (ctx->plane) cudaFree(ctx->plane);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int hop_ctx_sync(HopCtx* ctx)
{
  int st = bind(ctx);
  if (st) return st;
  if (ctx->copy_stream) CU(cudaStreamSynchronize(ctx->copy_stream));
  CU(cudaStreamSynchronize(ctx->stream));
  for (int i = 1; i < PU_SLOTS; i++) if (ctx->slots_pu[i].stream) CU(cudaStreamSynchronize(ctx->slots_pu[i].stream));
  for (auto& sl : ctx->slots) sl.busy = false;
  return HOP_OK;
}

void* hop_ctx_stream(HopCtx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t hop_ctx_launch_count(HopCtx* ctx) { return ctx ? ctx->launches : 0; }

int hop_ctx_stats(HopCtx* ctx, HopCtxStats* out)
{
  if (!ctx || !out) return fail(HOP_ERR_ARG, "NULL argument");
  *out = ctx->stats;
  return HOP_OK;
}

// ---------------------------------------------------------------------------------------------
// SS reference mirror (K4)
// ---------------------------------------------------------------------------------------------
int hop_ref_create(HopCtx* ctx, int pic_w, int pic_h, int margin)
{
  int st = bind(ctx);
  if (st) return st;
  if (pic_w <= 0 || pic_h <= 0 || margin < 0) return fail(HOP_ERR_ARG, "bad picture geometry %dx%d margin %d", pic_w, pic_h, margin);
  if (ctx->plane) { CU(cudaFree(ctx->plane)); ctx->plane = nullptr; }
  ctx->pic_w = pic_w; ctx->pic_h = pic_h; ctx->margin = margin; ctx->stride = pic_w + 2 * margin;
  size_t samples = (size_t)ctx->stride * (pic_h + 2 * margin);
  cudaError_t e = cudaMalloc((void**)&ctx->plane, samples * sizeof(int16_t));
  if (e != cudaSuccess) return fail(HOP_ERR_NOMEM, "cudaMalloc(plane %zu samples): %s", samples, cudaGetErrorString(e));
  ctx->plane_valid = false;
  ctx->ref_version++;
  return HOP_OK;
}

int hop_ref_reset(HopCtx* ctx, int value)
{
  int st = bind(ctx);
  if (st) return st;
  if (!ctx->plane) return fail(HOP_ERR_STATE, "hop_ref_reset before hop_ref_create");
  size_t samples = (size_t)ctx->stride * (ctx->pic_h + 2 * ctx->margin);
  int l = 0;
  CU(ref_fill_launch(ctx->plane, samples, value, ctx->stream, &l));
  ctx->launches += l;
  ctx->plane_valid = true;
  ctx->ref_version++;
  return HOP_OK;
}

int hop_ref_upload(HopCtx* ctx, const int16_t* plane, size_t plane_samples)
{
  int st = bind(ctx);
  if (st) return st;
  if (!ctx->plane) return fail(HOP_ERR_STATE, "hop_ref_upload before hop_ref_create");
  size_t samples = (size_t)ctx->stride * (ctx->pic_h + 2 * ctx->margin);
  if (!plane || plane_samples != samples) return fail(HOP_ERR_ARG, "plane has %zu samples, the mirror %zu", plane_samples, samples);
  CU(cudaMemcpyAsync(ctx->plane, plane, samples * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));   // the host plane may change right after the call
  ctx->plane_valid = true;
  ctx->ref_version++;
  return HOP_OK;
}

int hop_ref_update(HopCtx* ctx, int x, int y, int w, int h, const int16_t* src, int src_stride)
{
  int st = bind(ctx);
  if (st) return st;
  if (!ctx->plane || !ctx->plane_valid) return fail(HOP_ERR_STATE, "hop_ref_update before hop_ref_create/hop_ref_reset");
  if (!src || x < 0 || y < 0 || w <= 0 || h <= 0 || x + w > ctx->pic_w || y + h > ctx->pic_h || src_stride < w)
    return fail(HOP_ERR_ARG, "block (%d,%d %dx%d) outside the %dx%d picture", x, y, w, h, ctx->pic_w, ctx->pic_h);
  int16_t* origin = ctx->plane + (size_t)ctx->margin * ctx->stride + ctx->margin;
  // pageable source: the copy is staged by the runtime before the call returns, so `src` is not retained
  CU(cudaMemcpy2DAsync(origin + (size_t)y * ctx->stride + x, ctx->stride * sizeof(int16_t), src,
                       src_stride * sizeof(int16_t), w * sizeof(int16_t), h, cudaMemcpyHostToDevice, ctx->stream));
  int l = 0;
  CU(ref_extend_launch(origin, ctx->stride, ctx->pic_w, ctx->pic_h, ctx->margin, x, y, w, h, ctx->stream, &l));
  ctx->launches += l;
  ctx->ref_version++;      // speculative searches against the previous state can no longer be handed out
  return HOP_OK;
}

int hop_ref_download(HopCtx* ctx, int16_t* dst, size_t dst_samples)
{
  int st = bind(ctx);
  if (st) return st;
  if (!ctx->plane) return fail(HOP_ERR_STATE, "no reference plane");
  size_t samples = (size_t)ctx->stride * (ctx->pic_h + 2 * ctx->margin);
  if (!dst || dst_samples < samples) return fail(HOP_ERR_ARG, "destination too small (%zu < %zu)", dst_samples, samples);
  CU(cudaMemcpyAsync(dst, ctx->plane, samples * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return HOP_OK;
}

int hop_ref_stride(HopCtx* ctx) { return ctx ? ctx->stride : 0; }

const int16_t* hop_ref_origin_dev(HopCtx* ctx)
{
  if (!ctx || !ctx->plane) return nullptr;
  return ctx->plane + (size_t)ctx->margin * ctx->stride + ctx->margin;
}

// ---------------------------------------------------------------------------------------------
// device entry points
// ---------------------------------------------------------------------------------------------
namespace {
// sample offsets (relative to the pointer the kernels get) that may be dereferenced
RefBounds bounds_for(const HopCtx* ctx, bool mirror, size_t ref_samples)
{
  RefBounds rb;
  if (mirror) {
    const long long before = (long long)ctx->margin * ctx->stride + ctx->margin;       // samples in front of (0,0)
    rb.lo = -before;
    rb.hi = (long long)ctx->stride * (ctx->pic_h + 2 * ctx->margin) - 1 - before;
  } else {
    rb.lo = 0;
    rb.hi = (long long)ref_samples - 1;
  }
  return rb;
}

// bounds of a caller-supplied device reference buffer: the mirror's own geometry when d_ref is the mirror origin,
// else [0, ref_samples) -- AMVP start vectors are raw neighbour vectors, their windows may leave the buffer and the
// kernels keep every read inside these bounds
int dev_bounds(const HopCtx* ctx, const int16_t* d_ref, size_t ref_samples, RefBounds* rb)
{
  if (ctx->plane && d_ref == ctx->plane + (size_t)ctx->margin * ctx->stride + ctx->margin) { *rb = bounds_for(ctx, true, 0); return HOP_OK; }
  if (ref_samples == 0) return fail(HOP_ERR_ARG, "ref_samples == 0: the size of the device reference buffer is needed to bound the window reads");
  *rb = bounds_for(ctx, false, ref_samples);
  return HOP_OK;
}

int gt_dev(HopCtx* ctx, int n, const HopGtJob* d_jobs, const int16_t* d_org, const int16_t* d_ref, HopGtResult* d_out,
           int max_cols, int max_rows, cudaStream_t s, RefBounds rb)
{
  int l = 0;
  CU(gt_launch(n, d_jobs, d_org, d_ref, d_out, max_cols, max_rows, s, &l, rb));
  ctx->launches += l;
  return HOP_OK;
}

int k1_slices(const HopCtx* ctx, int n)
{
  // enough CTAs to cover the machine a few times: a single in-encoder call spreads one PU's window
  // over many SMs, a large batch needs no extra split
  int slices = (2 * ctx->sm_count + n - 1) / n;
  return slices > K1_MAX_SLICES ? K1_MAX_SLICES : (slices < 1 ? 1 : slices);
}
// A batch (n > 1) whose jobs all have one width and 8-bit content runs the per-width kernel k1_batch<W>:
// returns W (= cols / 4) and the shared memory it needs, or 0 (the all-widths kernel) for mixed batches.
int batch_width_hint(const HopCtx* ctx, int n, const HopSearchJob* first, size_t job_stride, size_t* smem_out)
{
  if (n <= 1) return 0;
  const int cols = first->cols;
  if (cols % 4 != 0 || cols < 4 || cols > HOP_MAX_PU) return 0;
  const int slices = (2 * ctx->sm_count + n - 1) / n > K1_MAX_SLICES ? K1_MAX_SLICES : ((2 * ctx->sm_count + n - 1) / n < 1 ? 1 : (2 * ctx->sm_count + n - 1) / n);
  size_t smem = 0;
  for (int i = 0; i < n; i++) {
    const HopSearchJob& j = *reinterpret_cast<const HopSearchJob*>(reinterpret_cast<const char*>(first) + (size_t)i * job_stride);
    if (j.cols != cols || j.bit_depth != 8) return 0;
    const size_t b = search_batch_smem_bytes(j, slices);
    if (b > smem) smem = b;
  }
  *smem_out = smem > (size_t)(160 * 1024) ? (size_t)(160 * 1024) : smem;
  return cols / 4;
}

int search_dev(HopCtx* ctx, int n, const HopSearchJob* d_jobs, const int16_t* d_org, const int16_t* d_ref,
               HopSearchResult* d_out, int smem_bytes, cudaStream_t s, unsigned* done_flag = nullptr, unsigned seq = 0,
               int job_stride = 0, const InlinePu* inl = nullptr, const PuSlot* slot = nullptr, int words_hint = 0)
{
  if (slot) {        // single-PU launch of a slot: its own merge words, so that slots may run concurrently
    int l = 0;
    CU(search_launch(n, d_jobs, d_org, d_ref, d_out, slot->d_key, slot->d_done, k1_slices(ctx, n), smem_bytes, s, &l,
                     done_flag, seq, job_stride, inl));
    ctx->launches += l;
    return HOP_OK;
  }
  // merge words: all-ones keys / zero tickets between launches (the kernel restores them itself).  They belong to
  // the context: batched searches of one context must be issued on one stream at a time.
  const size_t kcap = ctx->keys.cap, dcap = ctx->done.cap;
  int st = ensure(ctx, ctx->keys, sizeof(unsigned long long) * (size_t)n);
  if (st) return st;
  if ((st = ensure(ctx, ctx->done, sizeof(unsigned int) * (size_t)n))) return st;
  if (ctx->keys.cap != kcap) CU(cudaMemsetAsync(ctx->keys.p, 0xFF, ctx->keys.cap, s));
  if (ctx->done.cap != dcap) CU(cudaMemsetAsync(ctx->done.p, 0, ctx->done.cap, s));
  int l = 0;
  CU(search_launch(n, d_jobs, d_org, d_ref, d_out, (unsigned long long*)ctx->keys.p, (unsigned int*)ctx->done.p,
                   k1_slices(ctx, n), smem_bytes, s, &l, done_flag, seq, job_stride, inl, words_hint));
  ctx->launches += l;
  return HOP_OK;
}
}  // namespace

int hop_pattern_search_batch_dev(HopCtx* ctx, int n, const HopSearchJob* d_jobs, const int16_t* d_org,
                                 const int16_t* d_ref, HopSearchResult* d_out, int cols, int rows, int nx_max, int ny_max,
                                 void* stream)
{
  int st = bind(ctx);
  if (st) return st;
  if (n < 0 || (n > 0 && (!d_jobs || !d_org || !d_ref || !d_out))) return fail(HOP_ERR_ARG, "NULL device buffer");
  if (n == 0) return HOP_OK;
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  if (n > 1 && cols >= 4 && cols <= HOP_MAX_PU && cols % 4 == 0 && rows >= 4 && rows <= HOP_MAX_PU && nx_max > 0 && ny_max > 0) {
    // the caller vouches for one PU shape and a window bound: per-width kernel with exactly the shared memory it needs
    HopSearchJob shape;
    memset(&shape, 0, sizeof(shape));
    shape.cols = cols; shape.rows = rows; shape.is_ss = 1; shape.fast_enc = 1; shape.bit_depth = 8;
    shape.rng_right = nx_max - 1; shape.rng_bottom = ny_max - 1;
    size_t smem = search_batch_smem_bytes(shape, k1_slices(ctx, n));
    if (smem > (size_t)(160 * 1024)) smem = 160 * 1024;
    return search_dev(ctx, n, d_jobs, d_org, d_ref, d_out, (int)smem, s, nullptr, 0, 0, nullptr, nullptr, cols / 4);
  }
  return search_dev(ctx, n, d_jobs, d_org, d_ref, d_out, K1_DEFAULT_SMEM, s);
}

int hop_pattern_search_gt_batch_dev(HopCtx* ctx, int n, const HopGtJob* d_jobs, const int16_t* d_org,
                                    const int16_t* d_ref, size_t ref_samples, HopGtResult* d_out, int max_cols, int max_rows,
                                    void* stream)
{
  int st = bind(ctx);
  if (st) return st;
  if (n < 0 || (n > 0 && (!d_jobs || !d_org || !d_ref || !d_out))) return fail(HOP_ERR_ARG, "NULL device buffer");
  if (max_cols < 4 || max_cols > HOP_MAX_PU || max_rows < 4 || max_rows > HOP_MAX_PU)
    return fail(HOP_ERR_ARG, "shape bound %dx%d out of range", max_cols, max_rows);
  if (n == 0) return HOP_OK;
  RefBounds rb;
  if ((st = dev_bounds(ctx, d_ref, ref_samples, &rb))) return st;
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  int l = 0;
  CU(gt_launch(n, d_jobs, d_org, d_ref, d_out, max_cols, max_rows, s, &l, rb));
  ctx->launches += l;
  return HOP_OK;
}

int hop_dist_batch_dev(HopCtx* ctx, int n, const HopDistJob* d_jobs, const int16_t* d_org,
                       const int16_t* d_cur, uint32_t* d_out, void* stream)
{
  int st = bind(ctx);
  if (st) return st;
  if (n < 0 || (n > 0 && (!d_jobs || !d_org || !d_cur || !d_out))) return fail(HOP_ERR_ARG, "NULL device buffer");
  if (n == 0) return HOP_OK;
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  int l = 0;
  CU(dist_launch(n, d_jobs, d_org, d_cur, d_out, s, &l));
  ctx->launches += l;
  return HOP_OK;
}

// ---------------------------------------------------------------------------------------------
// host entry points: copy in, run, copy out
// ---------------------------------------------------------------------------------------------
namespace {

extern "C++" {
// In-encoder single call on the SS mirror (the latency path).  A slot's mapped pinned host buffer holds
// [job | W x H original block | result | completion flag]; the kernel reads the job and the block straight
// from host memory (a few KB over PCIe), writes the result and then the flag back, and the host spins on
// the flag: no copy calls, no stream synchronisation.
constexpr size_t PIN_JOB = 128;                                   // job slot (HopMotionJob, the largest job struct, is 104 B)
constexpr size_t PIN_ORG = HOP_MAX_PU * HOP_MAX_PU * sizeof(int16_t);
constexpr size_t PIN_OUT = 128;                                   // result slot (HopMotionResult is 72 B)
constexpr size_t PIN_FLAG = 64;
constexpr size_t PIN_BYTES = PIN_JOB + PIN_ORG + PIN_OUT + PIN_FLAG;
constexpr size_t SLOT_WORDS = 64;                                 // device bytes per slot: key 8 | ticket 4 | pad 4 | K1 result 16
static_assert(sizeof(HopMotionJob) <= PIN_JOB && sizeof(HopMotionResult) <= PIN_OUT, "slot layout");

int slots_ready(HopCtx* ctx)
{
  if (ctx->pin_base_h) return HOP_OK;
  unsigned char* h = nullptr;
  unsigned char* d = nullptr;
  CU(cudaHostAlloc((void**)&h, PIN_BYTES * PU_SLOTS, cudaHostAllocMapped));
  memset(h, 0, PIN_BYTES * PU_SLOTS);
  CU(cudaHostGetDevicePointer((void**)&d, h, 0));
  CU(cudaMalloc((void**)&ctx->slot_words, SLOT_WORDS * PU_SLOTS));
  CU(cudaMemsetAsync(ctx->slot_words, 0, SLOT_WORDS * PU_SLOTS, ctx->stream));
  for (int i = 0; i < PU_SLOTS; i++) {
    PuSlot& sl = ctx->slots_pu[i];
    sl.h = h + PIN_BYTES * i;
    sl.d = d + PIN_BYTES * i;
    unsigned char* w = ctx->slot_words + SLOT_WORDS * i;
    sl.d_key = (unsigned long long*)w;
    sl.d_done = (unsigned int*)(w + 8);
    sl.d_k1 = (HopSearchResult*)(w + 16);
    CU(cudaMemsetAsync(sl.d_key, 0xFF, sizeof(unsigned long long), ctx->stream));   // all-ones between launches
    if (i == 0) sl.stream = ctx->stream;
    else CU(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
  }
  CU(cudaEventCreateWithFlags(&ctx->ref_event, cudaEventDisableTiming));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->pin_base_h = h;
  return HOP_OK;
}

template <typename JOB>
int pack_single(HopCtx* ctx, PuSlot& sl, const JOB& job, const int16_t* org, size_t org_samples)
{
  const size_t need = (size_t)(job.rows - 1) * job.org_stride + job.cols;
  if (job.org_off < 0 || (size_t)job.org_off + need > org_samples) return fail(HOP_ERR_ARG, "original block outside the org buffer");
  JOB packed = job;
  packed.org_off = 0;                          // the block becomes contiguous behind the job
  packed.org_stride = job.cols;
  memcpy(sl.h, &packed, sizeof(JOB));
  int16_t* dst = (int16_t*)(sl.h + PIN_JOB);
  const int16_t* src = org + job.org_off;
  for (int r = 0; r < job.rows; r++) memcpy(dst + (size_t)r * job.cols, src + (size_t)r * job.org_stride, sizeof(int16_t) * job.cols);
  (void)ctx;
  return HOP_OK;
}

inline unsigned* slot_flag_dev(const PuSlot& sl) { return (unsigned*)(sl.d + PIN_JOB + PIN_ORG + PIN_OUT); }
inline const int16_t* slot_org_dev(const PuSlot& sl) { return (const int16_t*)(sl.d + PIN_JOB); }
inline unsigned char* slot_out_dev(const PuSlot& sl) { return sl.d + PIN_JOB + PIN_ORG; }

inline double mono_s() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }

// Wait for the slot's launch number sl.seq to publish its completion flag.  A kernel that dies, or that finishes
// without publishing, or that does not finish within the time-out is an error status -- never an endless spin.
int wait_slot_flag(HopCtx* ctx, PuSlot& sl)
{
  volatile unsigned* flag = (volatile unsigned*)(sl.h + PIN_JOB + PIN_ORG + PIN_OUT);
  const unsigned seq = sl.seq;
  unsigned long long spins = 0;
  double t0 = 0.0;
  while (*flag != seq) {
    HOP_CPU_RELAX();
    if ((++spins & 0xFFFF) == 0) {             // every 64K polls: did the kernel die, are we out of time?
      cudaError_t e = cudaStreamQuery(sl.stream);
      if (e != cudaSuccess && e != cudaErrorNotReady)
        return fail(HOP_ERR_CUDA, "kernel failed: %s", cudaGetErrorString(e));
      if (e == cudaSuccess && *flag != seq) return fail(HOP_ERR_CUDA, "kernel finished without publishing its result");
      const double now = mono_s();
      if (t0 == 0.0) t0 = now;
      else if (now - t0 > ctx->spin_timeout_s)
        return fail(HOP_ERR_CUDA, "no result within %.1f s (HOP_TIMEOUT_MS): kernel hung or launch lost", ctx->spin_timeout_s);
    }
  }
  __sync_synchronize();
  return HOP_OK;
}

template <typename RES>
int unpack_single(HopCtx* ctx, PuSlot& sl, RES* out)
{
  int st = wait_slot_flag(ctx, sl);
  if (st) return st;
  memcpy(out, sl.h + PIN_JOB + PIN_ORG, sizeof(RES));
  return HOP_OK;
}

// ---- the fused single-PU motion search of a slot ----------------------------------------------------------------
// normalised request: what the kernels see (contiguous block) with the unused AMVP entries cleared
HopMotionJob motion_key(const HopMotionJob& job)
{
  HopMotionJob k = job;
  k.search.org_off = 0;
  k.search.org_stride = job.search.cols;
  for (int i = 0; i < HOP_MAX_PRED; i++)
    if (i >= job.num_pred) { k.amvp[i].hor = 0; k.amvp[i].ver = 0; }
  return k;
}

int motion_job_ok(const HopMotionJob& mj, int i)
{
  const HopSearchJob& j = mj.search;
  if (!shape_ok(j.cols, j.rows) || j.bit_depth < 8 || j.bit_depth > 12 || mj.num_pred < 0 || mj.num_pred > HOP_MAX_PRED)
    return fail(HOP_ERR_ARG, "job %d: unsupported PU %dx%d / bit depth %d / num_pred %d", i, j.cols, j.rows, j.bit_depth, mj.num_pred);
  return HOP_OK;
}

// K1 -> (frac -> GT) for ONE PU against the SS mirror, enqueued on the slot's stream; the result lands in the
// slot's pinned buffer and its flag takes the value sl.seq.
int launch_motion_slot(HopCtx* ctx, PuSlot& sl, const HopMotionJob& job, const int16_t* org, size_t org_samples)
{
  const HopSearchJob& j = job.search;
  const size_t need = (size_t)(j.rows - 1) * j.org_stride + j.cols;
  if (j.org_off < 0 || (size_t)j.org_off + need > org_samples) return fail(HOP_ERR_ARG, "original block outside the org buffer");
  HOST_STAMP(0);
  // job (and a small block) as kernel parameters; larger blocks through the mapped buffer
  static thread_local InlinePu ipu;
  ipu.use = j.cols * j.rows <= INLINE_ORG_SAMPLES ? 2 : 1;
  ipu.job = motion_key(job);
  sl.key = ipu.job;
  int16_t* pin_org = (int16_t*)(sl.h + PIN_JOB);     // always filled: it is the block part of the cache key
  for (int r = 0; r < j.rows; r++) memcpy(pin_org + (size_t)r * j.cols, org + j.org_off + (size_t)r * j.org_stride, sizeof(int16_t) * j.cols);
  if (ipu.use == 2) memcpy(ipu.org, pin_org, sizeof(int16_t) * (size_t)j.cols * j.rows);
  const int slices = k1_slices(ctx, 1);
  size_t smem = search_smem_bytes(j, slices);
  if (smem > (size_t)(160 * 1024)) smem = 160 * 1024;
  const unsigned seq = ++sl.seq;
  sl.ref_version = ctx->ref_version;
  HOST_STAMP(1);
  int st = search_dev(ctx, 1, (const HopSearchJob*)sl.d, slot_org_dev(sl), hop_ref_origin_dev(ctx), sl.d_k1, (int)smem, sl.stream,
                      nullptr, 0, (int)sizeof(HopMotionJob), &ipu, &sl);
  if (st) return st;
  HOST_STAMP(2);
  int l = 0;
  cudaError_t ce = ctx->use_clusters
      ? motion_single_launch((const HopMotionJob*)sl.d, slot_org_dev(sl), hop_ref_origin_dev(ctx), sl.d_k1,
                             (HopMotionResult*)slot_out_dev(sl), j.cols, j.rows, sl.stream, &l,
                             bounds_for(ctx, true, 0), slot_flag_dev(sl), seq, &ipu)
      : cudaErrorNotSupported;
  if (ce == cudaErrorNotSupported)
    ce = motion_tail_launch(1, (const HopMotionJob*)sl.d, slot_org_dev(sl), hop_ref_origin_dev(ctx), sl.d_k1,
                            (HopMotionResult*)slot_out_dev(sl), j.cols, j.rows, sl.stream, &l,
                            bounds_for(ctx, true, 0), slot_flag_dev(sl), seq, &ipu);
  CU(ce);
  ctx->launches += l;
  HOST_STAMP(3);
  return HOP_OK;
}

// side streams read the mirror: order them behind the latest change of it (made on the context stream)
int order_behind_mirror(HopCtx* ctx, PuSlot& sl)
{
  if (sl.stream == ctx->stream || sl.synced_version == ctx->ref_version) return HOP_OK;
  if (ctx->ref_event_version != ctx->ref_version) {
    CU(cudaEventRecord(ctx->ref_event, ctx->stream));
    ctx->ref_event_version = ctx->ref_version;
  }
  CU(cudaStreamWaitEvent(sl.stream, ctx->ref_event, 0));
  sl.synced_version = ctx->ref_version;
  return HOP_OK;
}

// slot holding a finished-or-running speculative search for exactly this request, or -1
int find_cached(HopCtx* ctx, const HopMotionJob& key, const int16_t* org, size_t org_off, int org_stride)
{
  for (int i = 1; i < PU_SLOTS; i++) {
    const PuSlot& sl = ctx->slots_pu[i];
    if (!sl.cached || sl.ref_version != ctx->ref_version || memcmp(&sl.key, &key, sizeof(HopMotionJob)) != 0) continue;
    const int16_t* blk = (const int16_t*)(sl.h + PIN_JOB);
    const int cols = key.search.cols, rows = key.search.rows;
    bool same = true;
    for (int r = 0; r < rows && same; r++)
      same = memcmp(blk + (size_t)r * cols, org + org_off + (size_t)r * org_stride, sizeof(int16_t) * cols) == 0;
    if (same) return i;
  }
  return -1;
}
}  // extern "C++"

// upload jobs + org (+ ref unless the mirror is used); returns the device pointers
int stage_inputs(HopCtx* ctx, int n, const void* jobs, size_t job_size, const int16_t* org, size_t org_samples,
                 const int16_t* ref, size_t ref_samples, size_t out_bytes, const int16_t** d_ref_out)
{
  int st;
  if ((st = ensure(ctx, ctx->jobs, job_size * (size_t)n))) return st;
  if ((st = ensure(ctx, ctx->org, org_samples * sizeof(int16_t)))) return st;
  if ((st = ensure(ctx, ctx->out, out_bytes))) return st;
  CU(cudaMemcpyAsync(ctx->jobs.p, jobs, job_size * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->org.p, org, org_samples * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
  if (ref) {
    if ((st = ensure(ctx, ctx->ref, ref_samples * sizeof(int16_t)))) return st;
    CU(cudaMemcpyAsync(ctx->ref.p, ref, ref_samples * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
    *d_ref_out = (const int16_t*)ctx->ref.p;
  } else {
    if (!ctx->plane || !ctx->plane_valid) return fail(HOP_ERR_STATE, "ref == NULL but the context has no valid SS reference mirror");
    *d_ref_out = hop_ref_origin_dev(ctx);
  }
  return HOP_OK;
}

}  // namespace

int hop_pattern_search_batch(HopCtx* ctx, int n, const HopSearchJob* jobs, const int16_t* org, size_t org_samples,
                             const int16_t* ref, size_t ref_samples, HopSearchResult* out)
{
  int st = bind(ctx);
  if (st) return st;
  if (n < 0 || (n > 0 && (!jobs || !org || !out))) return fail(HOP_ERR_ARG, "NULL argument");
  if (n == 0) return HOP_OK;
  for (int i = 0; i < n; i++) {
    const HopSearchJob& j = jobs[i];
    if (j.cols < 1 || j.cols > HOP_MAX_PU || j.rows < 1 || j.rows > HOP_MAX_PU || j.bit_depth < 8 || j.bit_depth > 14)
      return fail(HOP_ERR_ARG, "job %d: unsupported block %dx%d / bit depth %d", i, j.cols, j.rows, j.bit_depth);
  }
  if (n == 1 && !ref && jobs[0].cols <= HOP_MAX_PU && jobs[0].rows <= HOP_MAX_PU) {
    // the encoder's call: one PU against the SS mirror
    if (!ctx->plane || !ctx->plane_valid) return fail(HOP_ERR_STATE, "ref == NULL but the context has no valid SS reference mirror");
    if ((st = slots_ready(ctx))) return st;
    PuSlot& sl = ctx->slots_pu[0];
    if ((st = pack_single(ctx, sl, jobs[0], org, org_samples))) return st;
    HopSearchJob packed;
    memcpy(&packed, sl.h, sizeof(packed));
    const int slices = k1_slices(ctx, 1);
    size_t smem = search_smem_bytes(packed, slices);
    const unsigned seq = ++sl.seq;
    st = search_dev(ctx, 1, (const HopSearchJob*)sl.d, slot_org_dev(sl), hop_ref_origin_dev(ctx),
                    (HopSearchResult*)slot_out_dev(sl), (int)(smem > (size_t)(160 * 1024) ? 160 * 1024 : smem), ctx->stream,
                    slot_flag_dev(sl), seq, 0, nullptr, &sl);
    if (st) return st;
    ctx->stats.single_calls++;
    return unpack_single(ctx, sl, out);
  }
  const int16_t* d_ref = nullptr;
  st = stage_inputs(ctx, n, jobs, sizeof(HopSearchJob), org, org_samples, ref, ref_samples,
                    sizeof(HopSearchResult) * (size_t)n, &d_ref);
  if (st) return st;
  size_t smem = 0;
  const int slices = k1_slices(ctx, n);
  size_t smem_w = 0;
  const int words = batch_width_hint(ctx, n, jobs, sizeof(HopSearchJob), &smem_w);
  if (words) smem = smem_w;
  else
    for (int i = 0; i < n; i++) {
      const size_t b = search_smem_bytes(jobs[i], slices);
      if (b > smem) smem = b;
    }
  st = search_dev(ctx, n, (const HopSearchJob*)ctx->jobs.p, (const int16_t*)ctx->org.p, d_ref,
                  (HopSearchResult*)ctx->out.p, (int)(smem > (size_t)(160 * 1024) ? 160 * 1024 : smem), ctx->stream,
                  nullptr, 0, 0, nullptr, nullptr, words);
  if (st) return st;
  CU(cudaMemcpyAsync(out, ctx->out.p, sizeof(HopSearchResult) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return HOP_OK;
}

int hop_pattern_search_gt_batch(HopCtx* ctx, int n, const HopGtJob* jobs, const int16_t* org, size_t org_samples,
                                const int16_t* ref, size_t ref_samples, HopGtResult* out)
{
  int st = bind(ctx);
  if (st) return st;
  if (n < 0 || (n > 0 && (!jobs || !org || !out))) return fail(HOP_ERR_ARG, "NULL argument");
  if (n == 0) return HOP_OK;
  int max_cols = 4, max_rows = 4;
  for (int i = 0; i < n; i++) {
    const HopGtJob& j = jobs[i];
    if (!shape_ok(j.cols, j.rows) || j.bit_depth < 8 || j.bit_depth > 14 || j.num_pred < 0 || j.num_pred > HOP_MAX_PRED)
      return fail(HOP_ERR_ARG, "job %d: unsupported PU %dx%d / bit depth %d / num_pred %d", i, j.cols, j.rows, j.bit_depth, j.num_pred);
    if (j.cols > max_cols) max_cols = j.cols;
    if (j.rows > max_rows) max_rows = j.rows;
  }
  if (n == 1 && !ref) {
    if (!ctx->plane || !ctx->plane_valid) return fail(HOP_ERR_STATE, "ref == NULL but the context has no valid SS reference mirror");
    if ((st = slots_ready(ctx))) return st;
    PuSlot& sl = ctx->slots_pu[0];
    if ((st = pack_single(ctx, sl, jobs[0], org, org_samples))) return st;
    const unsigned seq = ++sl.seq;
    int l = 0;
    cudaError_t ce = ctx->use_clusters
        ? gt_single_launch((const HopGtJob*)sl.d, slot_org_dev(sl), hop_ref_origin_dev(ctx),
                           (HopGtResult*)slot_out_dev(sl), max_cols, max_rows, ctx->stream, &l,
                           bounds_for(ctx, true, 0), slot_flag_dev(sl), seq)
        : cudaErrorNotSupported;
    if (ce == cudaErrorNotSupported)
      ce = gt_launch(1, (const HopGtJob*)sl.d, slot_org_dev(sl), hop_ref_origin_dev(ctx),
                     (HopGtResult*)slot_out_dev(sl), max_cols, max_rows, ctx->stream, &l,
                     bounds_for(ctx, true, 0), slot_flag_dev(sl), seq);
    CU(ce);
    ctx->launches += l;
    ctx->stats.single_calls++;
    if ((st = unpack_single(ctx, sl, out))) return st;
    ctx->stats.candidates += out->n_candidates;
    return HOP_OK;
  }
  const int16_t* d_ref = nullptr;
  st = stage_inputs(ctx, n, jobs, sizeof(HopGtJob), org, org_samples, ref, ref_samples,
                    sizeof(HopGtResult) * (size_t)n, &d_ref);
  if (st) return st;
  st = gt_dev(ctx, n, (const HopGtJob*)ctx->jobs.p, (const int16_t*)ctx->org.p, d_ref, (HopGtResult*)ctx->out.p,
              max_cols, max_rows, ctx->stream, bounds_for(ctx, ref == nullptr, ref_samples));
  if (st) return st;
  CU(cudaMemcpyAsync(out, ctx->out.p, sizeof(HopGtResult) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return HOP_OK;
}

int hop_pattern_search_gt_batch_async(HopCtx* ctx, int n, const HopGtJob* jobs, const int16_t* org, size_t org_samples,
                                      const int16_t* ref, size_t ref_samples, HopGtResult* out)
{
  int st = bind(ctx);
  if (st) return st;
  if (n < 0 || (n > 0 && (!jobs || !org || !ref || !out))) return fail(HOP_ERR_ARG, "NULL argument (the async form needs an explicit ref buffer)");
  if (n == 0) return HOP_OK;
  int max_cols = 4, max_rows = 4;
  for (int i = 0; i < n; i++) {
    const HopGtJob& j = jobs[i];
    if (!shape_ok(j.cols, j.rows) || j.bit_depth < 8 || j.bit_depth > 14 || j.num_pred < 0 || j.num_pred > HOP_MAX_PRED)
      return fail(HOP_ERR_ARG, "job %d: unsupported PU %dx%d / bit depth %d / num_pred %d", i, j.cols, j.rows, j.bit_depth, j.num_pred);
    if (j.cols > max_cols) max_cols = j.cols;
    if (j.rows > max_rows) max_rows = j.rows;
  }
  if (!ctx->copy_stream) CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  HopCtx::Slot& sl = ctx->slots[ctx->next_slot++ % HOP_ASYNC_SLOTS];
  if (!sl.h2d) { CU(cudaEventCreateWithFlags(&sl.h2d, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming)); }
  if (sl.busy) CU(cudaEventSynchronize(sl.done));          // the batch that used this slot HOP_ASYNC_SLOTS calls ago
  if ((st = ensure(ctx, sl.jobs, sizeof(HopGtJob) * (size_t)n))) return st;
  if ((st = ensure(ctx, sl.org, org_samples * sizeof(int16_t)))) return st;
  if ((st = ensure(ctx, sl.ref, ref_samples * sizeof(int16_t)))) return st;
  if ((st = ensure(ctx, sl.out, sizeof(HopGtResult) * (size_t)n))) return st;
  CU(cudaMemcpyAsync(sl.jobs.p, jobs, sizeof(HopGtJob) * (size_t)n, cudaMemcpyHostToDevice, ctx->copy_stream));
  CU(cudaMemcpyAsync(sl.org.p, org, org_samples * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->copy_stream));
  CU(cudaMemcpyAsync(sl.ref.p, ref, ref_samples * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->copy_stream));
  CU(cudaEventRecord(sl.h2d, ctx->copy_stream));
  CU(cudaStreamWaitEvent(ctx->stream, sl.h2d, 0));
  st = gt_dev(ctx, n, (const HopGtJob*)sl.jobs.p, (const int16_t*)sl.org.p, (const int16_t*)sl.ref.p,
              (HopGtResult*)sl.out.p, max_cols, max_rows, ctx->stream, bounds_for(ctx, false, ref_samples));
  if (st) return st;
  CU(cudaMemcpyAsync(out, sl.out.p, sizeof(HopGtResult) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaEventRecord(sl.done, ctx->stream));
  sl.busy = true;
  return HOP_OK;
}

int hop_dist_batch(HopCtx* ctx, int n, const HopDistJob* jobs, const int16_t* org, size_t org_samples,
                   const int16_t* cur, size_t cur_samples, uint32_t* out)
{
  int st = bind(ctx);
  if (st) return st;
  if (n < 0 || (n > 0 && (!jobs || !org || !cur || !out))) return fail(HOP_ERR_ARG, "NULL argument");
  if (n == 0) return HOP_OK;
  for (int i = 0; i < n; i++) {
    const HopDistJob& j = jobs[i];
    if (j.cols < 1 || j.rows < 1 || (j.func != HOP_DF_SAD && j.func != HOP_DF_HADS) || j.bit_depth < 8 || j.sub_shift < 0 || j.sub_shift > 3)
      return fail(HOP_ERR_ARG, "job %d: bad distortion job", i);
  }
  const int16_t* d_cur = nullptr;
  st = stage_inputs(ctx, n, jobs, sizeof(HopDistJob), org, org_samples, cur, cur_samples, sizeof(uint32_t) * (size_t)n, &d_cur);
  if (st) return st;
  st = hop_dist_batch_dev(ctx, n, (const HopDistJob*)ctx->jobs.p, (const int16_t*)ctx->org.p, d_cur,
                          (uint32_t*)ctx->out.p, ctx->stream);
  if (st) return st;
  CU(cudaMemcpyAsync(out, ctx->out.p, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return HOP_OK;
}

// ---------------------------------------------------------------------------------------------
// K5 and the fused motion search
// ---------------------------------------------------------------------------------------------
int hop_frac_search_batch(HopCtx* ctx, int n, const HopFracJob* jobs, const int16_t* org, size_t org_samples,
                          const int16_t* ref, size_t ref_samples, HopFracResult* out)
{
  int st = bind(ctx);
  if (st) return st;
  if (n < 0 || (n > 0 && (!jobs || !org || !out))) return fail(HOP_ERR_ARG, "NULL argument");
  if (n == 0) return HOP_OK;
  int max_cols = 4, max_rows = 4;
  for (int i = 0; i < n; i++) {
    const HopFracJob& j = jobs[i];
    if (!shape_ok(j.cols, j.rows) || j.bit_depth < 8 || j.bit_depth > 12)
      return fail(HOP_ERR_ARG, "job %d: unsupported PU %dx%d / bit depth %d", i, j.cols, j.rows, j.bit_depth);
    if (j.cols > max_cols) max_cols = j.cols;
    if (j.rows > max_rows) max_rows = j.rows;
  }
  const int16_t* d_ref = nullptr;
  st = stage_inputs(ctx, n, jobs, sizeof(HopFracJob), org, org_samples, ref, ref_samples, sizeof(HopFracResult) * (size_t)n, &d_ref);
  if (st) return st;
  int l = 0;
  CU(frac_launch(n, (const HopFracJob*)ctx->jobs.p, (const int16_t*)ctx->org.p, d_ref, (HopFracResult*)ctx->out.p,
                 max_cols, max_rows, ctx->stream, &l));
  ctx->launches += l;
  CU(cudaMemcpyAsync(out, ctx->out.p, sizeof(HopFracResult) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return HOP_OK;
}

// ---------------------------------------------------------------------------------------------
// K6: motion-compensated prediction (+ distortion / AMVP template cost)
// ---------------------------------------------------------------------------------------------
int hop_predict_batch(HopCtx* ctx, int n, const HopPredJob* jobs, const int16_t* org, size_t org_samples,
                      const int16_t* ref, size_t ref_samples, int16_t* dst, size_t dst_samples, HopPredResult* out)
{
  int st = bind(ctx);
  if (st) return st;
  if (n < 0 || (n > 0 && (!jobs || !out))) return fail(HOP_ERR_ARG, "NULL argument");
  if (n == 0) return HOP_OK;
  int max_cols = 4, max_rows = 4;
  bool any_gt = false, any_org = false, any_dst = false;
  for (int i = 0; i < n; i++) {
    const HopPredJob& j = jobs[i];
    if (!shape_ok(j.cols, j.rows) || j.bit_depth < 8 || j.bit_depth > 12 || (j.comp != 0 && j.comp != 1) ||
        (j.dist_func != 0 && j.dist_func != HOP_DF_SAD && j.dist_func != HOP_DF_HADS))
      return fail(HOP_ERR_ARG, "job %d: unsupported PU %dx%d / bit depth %d / component %d / distortion %d", i, j.cols, j.rows, j.bit_depth, j.comp, j.dist_func);
    if (j.comp && ((j.cols | j.rows) & 1)) return fail(HOP_ERR_ARG, "job %d: chroma of an odd-sized PU", i);
    if (j.template_cost && j.comp) return fail(HOP_ERR_ARG, "job %d: the template cost is a luma quantity", i);
    if (!ref && j.comp) return fail(HOP_ERR_ARG, "job %d: the SS mirror holds luma only; chroma jobs need an explicit plane", i);
    const int bw = j.comp ? j.cols >> 1 : j.cols, bh = j.comp ? j.rows >> 1 : j.rows;
    if (j.dist_func || j.template_cost) {
      any_org = true;
      if (!org || j.org_off < 0 || (size_t)j.org_off + (size_t)(bh - 1) * j.org_stride + bw > org_samples)
        return fail(HOP_ERR_ARG, "job %d: original block outside the org buffer", i);
    }
    if (j.dst_off >= 0) {
      any_dst = true;
      if (!dst || (size_t)j.dst_off + (size_t)(bh - 1) * j.dst_stride + bw > dst_samples)
        return fail(HOP_ERR_ARG, "job %d: prediction outside the dst buffer", i);
    }
    any_gt |= j.gt_flag != 0;
    if (j.cols > max_cols) max_cols = j.cols;
    if (j.rows > max_rows) max_rows = j.rows;
  }
  static const int16_t dummy_org = 0;
  const int16_t* d_ref = nullptr;
  st = stage_inputs(ctx, n, jobs, sizeof(HopPredJob), any_org ? org : &dummy_org, any_org ? org_samples : 1, ref, ref_samples,
                    sizeof(HopPredResult) * (size_t)n, &d_ref);
  if (st) return st;
  if (any_dst) {
    if ((st = ensure(ctx, ctx->sink, dst_samples * sizeof(int16_t) + 64))) return st;
    // samples no job writes (gaps, candidates the template gate rejected) come back as the caller's buffer holds them
    CU(cudaMemcpyAsync(ctx->sink.p, dst, dst_samples * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
  }
  int l = 0;
  CU(predict_launch(n, (const HopPredJob*)ctx->jobs.p, (const int16_t*)ctx->org.p, d_ref, any_dst ? (int16_t*)ctx->sink.p : nullptr,
                    (HopPredResult*)ctx->out.p, max_cols, max_rows, any_gt, ctx->stream, &l, bounds_for(ctx, ref == nullptr, ref_samples)));
  ctx->launches += l;
  CU(cudaMemcpyAsync(out, ctx->out.p, sizeof(HopPredResult) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  if (any_dst) CU(cudaMemcpyAsync(dst, ctx->sink.p, dst_samples * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return HOP_OK;
}

// ---------------------------------------------------------------------------------------------
// K7: intra mode pre-screen
// ---------------------------------------------------------------------------------------------
int hop_intra_prescreen_batch(HopCtx* ctx, int n, const HopIntraJob* jobs, const int16_t* org, size_t org_samples,
                              const int32_t* refs, size_t refs_count, uint32_t* out)
{
  int st = bind(ctx);
  if (st) return st;
  if (n < 0 || (n > 0 && (!jobs || !org || !refs || !out))) return fail(HOP_ERR_ARG, "NULL argument");
  if (n == 0) return HOP_OK;
  int max_size = 4;
  for (int i = 0; i < n; i++) {
    const HopIntraJob& j = jobs[i];
    if ((j.size != 4 && j.size != 8 && j.size != 16 && j.size != 32 && j.size != 64) || j.bit_depth < 8 || j.bit_depth > 12)
      return fail(HOP_ERR_ARG, "job %d: unsupported block %dx%d / bit depth %d", i, j.size, j.size, j.bit_depth);
    if (j.org_off < 0 || (size_t)j.org_off + (size_t)(j.size - 1) * j.org_stride + j.size > org_samples)
      return fail(HOP_ERR_ARG, "job %d: original block outside the org buffer", i);
    if (j.refs_off < 0 || (size_t)j.refs_off + 4 * (size_t)(2 * j.size + 1) > refs_count)
      return fail(HOP_ERR_ARG, "job %d: reference samples outside the refs buffer", i);
    if (j.size > max_size) max_size = j.size;
  }
  if ((st = ensure(ctx, ctx->jobs, sizeof(HopIntraJob) * (size_t)n))) return st;
  if ((st = ensure(ctx, ctx->org, org_samples * sizeof(int16_t)))) return st;
  if ((st = ensure(ctx, ctx->ref, refs_count * sizeof(int32_t)))) return st;
  if ((st = ensure(ctx, ctx->out, sizeof(uint32_t) * HOP_INTRA_MODES * (size_t)n))) return st;
  CU(cudaMemcpyAsync(ctx->jobs.p, jobs, sizeof(HopIntraJob) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->org.p, org, org_samples * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->ref.p, refs, refs_count * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  int l = 0;
  CU(intra_launch(n, (const HopIntraJob*)ctx->jobs.p, (const int16_t*)ctx->org.p, (const int32_t*)ctx->ref.p, (uint32_t*)ctx->out.p,
                  max_size, ctx->stream, &l));
  ctx->launches += l;
  CU(cudaMemcpyAsync(out, ctx->out.p, sizeof(uint32_t) * HOP_INTRA_MODES * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return HOP_OK;
}

int hop_motion_search_batch(HopCtx* ctx, int n, const HopMotionJob* jobs, const int16_t* org, size_t org_samples,
                            const int16_t* ref, size_t ref_samples, HopMotionResult* out)
{
  int st = bind(ctx);
  if (st) return st;
  if (n < 0 || (n > 0 && (!jobs || !org || !out))) return fail(HOP_ERR_ARG, "NULL argument");
  if (n == 0) return HOP_OK;
  int max_cols = 4, max_rows = 4;
  for (int i = 0; i < n; i++) {
    if ((st = motion_job_ok(jobs[i], i))) return st;
    const HopSearchJob& j = jobs[i].search;
    if (j.cols > max_cols) max_cols = j.cols;
    if (j.rows > max_rows) max_rows = j.rows;
  }
  if (n == 1 && !ref) {
    // the encoder's call: zero-copy job / block / result, two stream-ordered launches, one wait -- or no launch at
    // all when a speculative search for exactly this request is already running or done
    if (!ctx->plane || !ctx->plane_valid) return fail(HOP_ERR_STATE, "ref == NULL but the context has no valid SS reference mirror");
    if ((st = slots_ready(ctx))) return st;
    const HopSearchJob& j = jobs[0].search;
    const size_t need = (size_t)(j.rows - 1) * j.org_stride + j.cols;
    if (j.org_off < 0 || (size_t)j.org_off + need > org_samples) return fail(HOP_ERR_ARG, "original block outside the org buffer");
    const int hit = find_cached(ctx, motion_key(jobs[0]), org, (size_t)j.org_off, j.org_stride);
    PuSlot& sl = ctx->slots_pu[hit > 0 ? hit : 0];
    if (hit > 0) {
      sl.cached = false;
      ctx->stats.cache_hits++;
    } else {
      ctx->stats.cache_misses++;
      if ((st = launch_motion_slot(ctx, sl, jobs[0], org, org_samples))) return st;
    }
    ctx->stats.single_calls++;
    st = unpack_single(ctx, sl, out);
    HOST_STAMP(4);
    if (st == HOP_OK) ctx->stats.candidates += out->gt.n_candidates;
    return st;
  }
  size_t smem = 0;
  const int slices = k1_slices(ctx, n);
  for (int i = 0; i < n; i++) {
    const size_t b = search_smem_bytes(jobs[i].search, slices);
    if (b > smem) smem = b;
  }
  if (smem > (size_t)(160 * 1024)) smem = 160 * 1024;
  size_t smem_w = 0;
  const int words = batch_width_hint(ctx, n, &jobs[0].search, sizeof(HopMotionJob), &smem_w);
  if (words) smem = smem_w;
  if ((st = ensure(ctx, ctx->k1res, sizeof(HopSearchResult) * (size_t)n))) return st;
  HopSearchResult* d_k1 = (HopSearchResult*)ctx->k1res.p;
  const int16_t* d_ref = nullptr;
  st = stage_inputs(ctx, n, jobs, sizeof(HopMotionJob), org, org_samples, ref, ref_samples, sizeof(HopMotionResult) * (size_t)n, &d_ref);
  if (st) return st;
  st = search_dev(ctx, n, (const HopSearchJob*)ctx->jobs.p, (const int16_t*)ctx->org.p, d_ref, d_k1, (int)smem, ctx->stream,
                  nullptr, 0, (int)sizeof(HopMotionJob), nullptr, nullptr, words);
  if (st) return st;
  int l = 0;
  CU(motion_tail_launch(n, (const HopMotionJob*)ctx->jobs.p, (const int16_t*)ctx->org.p, d_ref, d_k1,
                        (HopMotionResult*)ctx->out.p, max_cols, max_rows, ctx->stream, &l,
                        bounds_for(ctx, ref == nullptr, ref_samples)));
  ctx->launches += l;
  CU(cudaMemcpyAsync(out, ctx->out.p, sizeof(HopMotionResult) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return HOP_OK;
}

// Speculative searches: enqueue the fused single-PU motion search of every job on a side stream and return at
// once.  The caller goes on with host work; when it later asks hop_motion_search_batch(n = 1, ref = NULL) for
// EXACTLY one of these requests (same job fields, same block samples, SS mirror unchanged since) it gets that
// launch's result -- bit for bit what the synchronous call computes, because it is the same kernels on the same
// inputs.  Any other request is simply a miss.  Results nobody asks for are dropped when their slot is reused.
int hop_motion_search_prefetch(HopCtx* ctx, int n, const HopMotionJob* jobs, const int16_t* org, size_t org_samples)
{
  int st = bind(ctx);
  if (st) return st;
  if (n < 0 || (n > 0 && (!jobs || !org))) return fail(HOP_ERR_ARG, "NULL argument");
  if (n == 0) return HOP_OK;
  if (!ctx->plane || !ctx->plane_valid) return fail(HOP_ERR_STATE, "prefetch needs a valid SS reference mirror");
  if ((st = slots_ready(ctx))) return st;
  for (int i = 0; i < n; i++) {
    if ((st = motion_job_ok(jobs[i], i))) return st;
    const HopSearchJob& j = jobs[i].search;
    const size_t need = (size_t)(j.rows - 1) * j.org_stride + j.cols;
    if (j.org_off < 0 || (size_t)j.org_off + need > org_samples) return fail(HOP_ERR_ARG, "original block outside the org buffer");
    if (find_cached(ctx, motion_key(jobs[i]), org, (size_t)j.org_off, j.org_stride) > 0) continue;   // already on its way
    PuSlot& sl = ctx->slots_pu[ctx->next_pu_slot];
    ctx->next_pu_slot = ctx->next_pu_slot + 1 < PU_SLOTS ? ctx->next_pu_slot + 1 : 1;
    if (sl.cached) {
      // oldest speculative search, never asked for: its kernels must be through with the slot's buffers
      if ((st = wait_slot_flag(ctx, sl))) return st;
      sl.cached = false;
      ctx->stats.prefetch_dropped++;
    }
    if ((st = order_behind_mirror(ctx, sl))) return st;
    if ((st = launch_motion_slot(ctx, sl, jobs[i], org, org_samples))) return st;
    sl.cached = true;
    ctx->stats.prefetched++;
  }
  return HOP_OK;
}

// ---------------------------------------------------------------------------------------------
// exhaustive sweep
// ---------------------------------------------------------------------------------------------
int hop_gt_sweep_keys_dev(HopCtx* ctx, int n, const HopGtJob* d_jobs, const int16_t* d_org, const int16_t* d_ref,
                          size_t ref_samples, int max_cols, int max_rows, int cand_begin, int cand_end,
                          uint64_t* d_keys, uint32_t* d_counts, void* stream)
{
  int st = bind(ctx);
  if (st) return st;
  if (n < 0 || (n > 0 && (!d_jobs || !d_org || !d_ref || !d_keys))) return fail(HOP_ERR_ARG, "NULL device buffer");
  if (max_cols < 4 || max_cols > HOP_MAX_PU || max_rows < 4 || max_rows > HOP_MAX_PU)
    return fail(HOP_ERR_ARG, "shape bound %dx%d out of range", max_cols, max_rows);
  if (cand_begin < 0 || cand_end > SWEEP_CANDS || cand_begin > cand_end)
    return fail(HOP_ERR_ARG, "candidate slice [%d,%d) outside [0,%d]", cand_begin, cand_end, SWEEP_CANDS);
  if (n == 0) return HOP_OK;
  if ((st = sweep_table_ready(ctx))) return st;
  RefBounds rb;
  if ((st = dev_bounds(ctx, d_ref, ref_samples, &rb))) return st;
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  int l = 0;
  CU(sweep_init_launch(n, (unsigned long long*)d_keys, d_counts, s, &l));
  if (cand_end > cand_begin) {
    const int batches = (cand_end - cand_begin + GT_CANDS - 1) / GT_CANDS;
    // spread one PU's candidate batches over the machine: at least ~6 waves of the 2 CTAs resident per SM, so that
    // the last, partly filled wave costs a few percent instead of up to half of the launch
    int chunks = (12 * ctx->sm_count + n - 1) / n;
    if (chunks > batches) chunks = batches;
    if (chunks < 1) chunks = 1;
    CU(sweep_keys_launch(n, d_jobs, d_org, d_ref, max_cols, max_rows, cand_begin, cand_end, chunks,
                         (unsigned long long*)d_keys, d_counts, s, &l, rb));
  }
  ctx->launches += l;
  return HOP_OK;
}

int hop_gt_sweep_finalize_dev(HopCtx* ctx, int n, const HopGtJob* d_jobs, const uint64_t* d_keys,
                              const uint32_t* d_counts, HopGtResult* d_out, void* stream)
{
  int st = bind(ctx);
  if (st) return st;
  if (n < 0 || (n > 0 && (!d_jobs || !d_keys || !d_out))) return fail(HOP_ERR_ARG, "NULL device buffer");
  if (n == 0) return HOP_OK;
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  int l = 0;
  CU(sweep_finalize_launch(n, d_jobs, (const unsigned long long*)d_keys, d_counts, d_out, s, &l));
  ctx->launches += l;
  return HOP_OK;
}

// ---- sharded sweep over peer memory (no collective call) ------------------------------------------------------
namespace {
inline size_t xch_gkeys_bytes(int max_pus)  { return sizeof(unsigned long long) * 2 * (size_t)max_pus; }
inline size_t xch_gcounts_bytes(int max_pus) { return (sizeof(unsigned int) * 2 * (size_t)max_pus + 63) & ~(size_t)63; }
inline size_t xch_block_bytes(int max_pus)  { return xch_gkeys_bytes(max_pus) + xch_gcounts_bytes(max_pus) + 64; }
}

int hop_sweep_exchange_create(HopCtx* ctx, int max_pus, HopSweepHandle* mine)
{
  int st = bind(ctx);
  if (st) return st;
  if (max_pus <= 0 || !mine) return fail(HOP_ERR_ARG, "bad exchange geometry");
  if (ctx->xch_block) return fail(HOP_ERR_STATE, "exchange already created");
  static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(HopSweepHandle), "handle size");
  const size_t bytes = xch_block_bytes(max_pus);
  CU(cudaMalloc((void**)&ctx->xch_block, bytes));
  CU(cudaMemsetAsync(ctx->xch_block, 0, bytes, ctx->stream));
  CU(cudaMemsetAsync(ctx->xch_block, 0xFF, xch_gkeys_bytes(max_pus), ctx->stream));          // merge words idle = all-ones
  const size_t local = (sizeof(unsigned long long) + 2 * sizeof(unsigned int)) * (size_t)max_pus + 64;
  CU(cudaMalloc((void**)&ctx->xch_local, local));
  CU(cudaMemsetAsync(ctx->xch_local, 0, local, ctx->stream));
  CU(cudaMemsetAsync(ctx->xch_local, 0xFF, sizeof(unsigned long long) * (size_t)max_pus, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, ctx->xch_block));
  memset(mine, 0, sizeof(*mine));
  memcpy(mine, &h, sizeof(h));
  ctx->xch_max_pus = max_pus;
  ctx->xch_world = 0;
  return HOP_OK;
}

int hop_sweep_exchange_connect(HopCtx* ctx, int world, int rank, const HopSweepHandle* all)
{
  int st = bind(ctx);
  if (st) return st;
  if (!ctx->xch_block) return fail(HOP_ERR_STATE, "hop_sweep_exchange_connect before hop_sweep_exchange_create");
  if (world < 1 || world > SWEEP_MAX_RANKS || rank < 0 || rank >= world || !all)
    return fail(HOP_ERR_ARG, "world %d / rank %d out of range (at most %d ranks)", world, rank, SWEEP_MAX_RANKS);
  for (int r = 0; r < world; r++) {
    if (r == rank) { ctx->xch_peer[r] = ctx->xch_block; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, &all[r], sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(HOP_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s (no peer access between the GPUs?)", r, cudaGetErrorString(e)); }
    ctx->xch_peer[r] = (unsigned char*)p;
  }
  ctx->xch_world = world;
  ctx->xch_rank = rank;
  ctx->xch_epoch = 0;
  return HOP_OK;
}

int hop_gt_sweep_sharded_dev(HopCtx* ctx, int n, const HopGtJob* d_jobs, const int16_t* d_org, const int16_t* d_ref,
                             size_t ref_samples, int max_cols, int max_rows, HopGtResult* d_out, void* stream)
{
  int st = bind(ctx);
  if (st) return st;
  if (ctx->xch_world < 1) return fail(HOP_ERR_STATE, "sharded sweep before hop_sweep_exchange_connect");
  if (n <= 0 || n > ctx->xch_max_pus) return fail(HOP_ERR_ARG, "%d PUs, the exchange holds %d", n, ctx->xch_max_pus);
  if (!d_jobs || !d_org || !d_ref || !d_out) return fail(HOP_ERR_ARG, "NULL device buffer");
  if (max_cols < 4 || max_cols > HOP_MAX_PU || max_rows < 4 || max_rows > HOP_MAX_PU)
    return fail(HOP_ERR_ARG, "shape bound %dx%d out of range", max_cols, max_rows);
  if ((st = sweep_table_ready(ctx))) return st;
  RefBounds rb;
  if ((st = dev_bounds(ctx, d_ref, ref_samples, &rb))) return st;
  cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
  const int world = ctx->xch_world, rank = ctx->xch_rank, mp = ctx->xch_max_pus;
  // contiguous slice of the flattened candidate range owned by this rank (sizes differ by at most one)
  const int base = SWEEP_CANDS / world, rem = SWEEP_CANDS % world;
  const int cand_begin = rank * base + (rank < rem ? rank : rem), cand_end = cand_begin + base + (rank < rem ? 1 : 0);
  const unsigned epoch = ++ctx->xch_epoch;
  SweepXchg xc = {};
  for (int r = 0; r < world; r++) {
    xc.gkeys[r] = (unsigned long long*)ctx->xch_peer[r];
    xc.gcounts[r] = (unsigned int*)(ctx->xch_peer[r] + xch_gkeys_bytes(mp));
    xc.arrived[r] = (unsigned int*)(ctx->xch_peer[r] + xch_gkeys_bytes(mp) + xch_gcounts_bytes(mp));
  }
  unsigned long long* l_keys = (unsigned long long*)ctx->xch_local;
  unsigned int* l_counts = (unsigned int*)(l_keys + mp);
  xc.pu_done = l_counts + mp;
  xc.grid_done = xc.pu_done + mp;
  xc.world = world; xc.rank = rank; xc.max_pus = mp; xc.parity = epoch & 1u;
  const int batches = (cand_end - cand_begin + GT_CANDS - 1) / GT_CANDS;
  int chunks = (12 * ctx->sm_count + n - 1) / n;
  if (chunks > batches) chunks = batches;
  if (chunks < 1) chunks = 1;
  int l = 0;
  CU(sweep_keys_launch(n, d_jobs, d_org, d_ref, max_cols, max_rows, cand_begin, cand_end, chunks, l_keys, l_counts, s, &l, rb, &xc));
  CU(sweep_finalize_x_launch(n, d_jobs, xc.gkeys[rank] + (size_t)xc.parity * mp, xc.gcounts[rank] + (size_t)xc.parity * mp,
                             xc.arrived[rank], epoch * (unsigned)world, d_out, s, &l));
  ctx->launches += l;
  return HOP_OK;
}

int hop_gt_sweep_batch(HopCtx* ctx, int n, const HopGtJob* jobs, const int16_t* org, size_t org_samples,
                       const int16_t* ref, size_t ref_samples, HopGtResult* out)
{
  int st = bind(ctx);
  if (st) return st;
  if (n < 0 || (n > 0 && (!jobs || !org || !out))) return fail(HOP_ERR_ARG, "NULL argument");
  if (n == 0) return HOP_OK;
  int max_cols = 4, max_rows = 4;
  for (int i = 0; i < n; i++) {
    const HopGtJob& j = jobs[i];
    if (!shape_ok(j.cols, j.rows) || j.bit_depth < 8 || j.bit_depth > 14)
      return fail(HOP_ERR_ARG, "job %d: unsupported PU %dx%d / bit depth %d", i, j.cols, j.rows, j.bit_depth);
    if (j.cols > max_cols) max_cols = j.cols;
    if (j.rows > max_rows) max_rows = j.rows;
  }
  const int16_t* d_ref = nullptr;
  st = stage_inputs(ctx, n, jobs, sizeof(HopGtJob), org, org_samples, ref, ref_samples,
                    sizeof(HopGtResult) * (size_t)n, &d_ref);
  if (st) return st;
  if ((st = ensure(ctx, ctx->sweep_keys, (sizeof(unsigned long long) + sizeof(unsigned int)) * (size_t)n + 16))) return st;
  uint64_t* d_keys = (uint64_t*)ctx->sweep_keys.p;
  uint32_t* d_counts = (uint32_t*)(d_keys + n);
  st = hop_gt_sweep_keys_dev(ctx, n, (const HopGtJob*)ctx->jobs.p, (const int16_t*)ctx->org.p, d_ref, ref_samples, max_cols, max_rows,
                             0, SWEEP_CANDS, d_keys, d_counts, ctx->stream);
  if (st) return st;
  st = hop_gt_sweep_finalize_dev(ctx, n, (const HopGtJob*)ctx->jobs.p, d_keys, d_counts, (HopGtResult*)ctx->out.p, ctx->stream);
  if (st) return st;
  CU(cudaMemcpyAsync(out, ctx->out.p, sizeof(HopGtResult) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return HOP_OK;
}

// ---------------------------------------------------------------------------------------------
// ALU probes
// ---------------------------------------------------------------------------------------------
int hop_probe_alu(HopCtx* ctx, int what, double* gops_out, double* ms_out)
{
  int st = bind(ctx);
  if (st) return st;
  if (what < 0 || what > 5 || !gops_out) return fail(HOP_ERR_ARG, "bad probe id %d", what);
  if ((st = ensure(ctx, ctx->sink, 64))) return st;
  const int blocks = ctx->sm_count * 8, threads = 256, iters = 1 << 14;
  double per = 0;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  CU(probe_launch(what, blocks, threads, iters / 16, (unsigned*)ctx->sink.p, ctx->stream, &per));   // warm-up
  CU(cudaEventRecord(e0, ctx->stream));
  CU(probe_launch(what, blocks, threads, iters, (unsigned*)ctx->sink.p, ctx->stream, &per));
  CU(cudaEventRecord(e1, ctx->stream));
  CU(cudaEventSynchronize(e1));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  ctx->launches += 2;
  const double ops = (double)blocks * threads * (double)iters * per;
  *gops_out = ops / (ms * 1e-3) / 1e9;
  if (ms_out) *ms_out = ms;
  return HOP_OK;
}

}  // extern "C"

#ifdef HOP_TRACE
// debug builds only: out[0..7] host stamps (ns, CLOCK_MONOTONIC), out[8..71] k1 table, out[72..135] tail table (%globaltimer ns)
extern "C" int hop_debug_trace(HopCtx* ctx, unsigned long long* out)
{
  int st = bind(ctx);
  if (st) return st;
  CU(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < 8; i++) out[i] = g_host_trace[i];
  hop::trace_read_k1(out + 8);
  hop::trace_read_k2(out + 8 + 64);
  return HOP_OK;
}
#endif
