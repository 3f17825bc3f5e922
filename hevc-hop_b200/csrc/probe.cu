// probe.cu -- ALU peak probes for the roofline: MEASURED_PEAKS.json only holds HBM
// and bf16 numbers, the HOP kernels are bound by the int32 and fp64 pipes.  Each probe runs a long
// chain-parallel instruction loop on every SM; the library reports lane-operations per second.
#include "hop_common.cuh"
#include "hop_internal.h"

namespace hop {

constexpr int ILP = 8;

template <int WHAT>
__global__ void __launch_bounds__(256) k_probe(int iters, unsigned* sink)
{
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (WHAT == 0 || WHAT == 1 || WHAT == 5) {
    unsigned a[ILP], b = t * 2654435761u + 12345u, c = t ^ 0x9e3779b9u;
#pragma unroll
    for (int k = 0; k < ILP; k++) a[k] = t + k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int k = 0; k < ILP; k++) {
        if (WHAT == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[k]) : "r"(b));
        if (WHAT == 1) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[k]) : "r"(b), "r"(c));
        if (WHAT == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b), "r"(c));
      }
    }
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s ^= a[k];
    if (s == 0x12345678u) sink[0] = s;
  } else {
    double a[ILP], b = 1.0000001 + t * 1e-12, c = 1e-9;
#pragma unroll
    for (int k = 0; k < ILP; k++) a[k] = 1.0 + k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int k = 0; k < ILP; k++) {
        if (WHAT == 2) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[k]) : "d"(c));
        if (WHAT == 3) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[k]) : "d"(b));
        if (WHAT == 4) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[k]) : "d"(b), "d"(c));
      }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s += a[k];
    if (s == 0.123) sink[0] = 1;
  }
}

cudaError_t probe_launch(int what, int blocks, int threads, int iters, unsigned* d_sink,
                         cudaStream_t stream, double* lane_ops_per_thread_iter)
{
  *lane_ops_per_thread_iter = (double)ILP;
  switch (what) {
    case 0: k_probe<0><<<blocks, threads, 0, stream>>>(iters, d_sink); break;
    case 1: k_probe<1><<<blocks, threads, 0, stream>>>(iters, d_sink); break;
    case 2: k_probe<2><<<blocks, threads, 0, stream>>>(iters, d_sink); break;
    case 3: k_probe<3><<<blocks, threads, 0, stream>>>(iters, d_sink); break;
    case 4: k_probe<4><<<blocks, threads, 0, stream>>>(iters, d_sink); break;
    case 5: k_probe<5><<<blocks, threads, 0, stream>>>(iters, d_sink); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace hop
