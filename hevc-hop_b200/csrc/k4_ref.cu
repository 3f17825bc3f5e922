// k4_ref.cu -- K4: device mirror of the SS reference luma plane (sm_100a).
//
// Replaces, for the mirror, TComPicYuv::setPicPel(NOT_VALID) (TLibCommon/TComPicYuv.cpp:200-206),
// the CU copy of TEncCu::xCopyYuv2SSRef (TLibEncoder/TEncCu.cpp:1694) and the border re-extension
// TComPicYuv::extendPicBorder / xExtendPicCompBorder (TComPicYuv.cpp:236-274).
//
// The reference re-extends the whole picture border after every CU; on the device only the margin
// samples that depend on the patched block change, so the extension is incremental and exact:
//   left/right margins replicate the row's edge sample; the top (bottom) margin rows replicate the
//   first (last) picture row INCLUDING its left/right margins.
#include "hop_common.cuh"
#include "hop_internal.h"

namespace hop {

__global__ void k4_fill(int16_t* plane, size_t samples, int16_t value)
{
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < samples; i += stride) plane[i] = value;
}

// origin = sample (0,0).  The block [x,x+w) x [y,y+h) has just been written.
__global__ void k4_extend(int16_t* origin, int stride, int pic_w, int pic_h, int margin,
                          int x, int y, int w, int h)
{
  // extended column range touched by this patch (block columns plus the side margins it feeds)
  const int ex0 = (x == 0) ? -margin : x;
  const int ex1 = (x + w == pic_w) ? pic_w + margin : x + w;     // exclusive
  const int ew = ex1 - ex0;
  const int top = (y == 0) ? margin : 0, bottom = (y + h == pic_h) ? margin : 0;
  // phase 1: side margins of the block rows
  const int side = (x == 0 ? margin : 0) + (x + w == pic_w ? margin : 0);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
  for (int i = tid; i < side * h; i += nthr) {
    const int r = y + i / side;
    int k = i % side;
    int16_t* row = origin + (long long)r * stride;
    if (x == 0 && k < margin) row[-margin + k] = row[0];
    else { if (x == 0) k -= margin; row[pic_w + k] = row[pic_w - 1]; }
  }
  // phase 2 needs phase 1 of rows 0 / pic_h-1 when the patch touches a corner: recompute the source
  // value directly instead of synchronising.
  for (int i = tid; i < ew * (top + bottom); i += nthr) {
    const int m = i / ew, cx = ex0 + i % ew;
    const int sx = cx < 0 ? 0 : (cx >= pic_w ? pic_w - 1 : cx);
    if (m < top) {
      origin[(long long)(-1 - m) * stride + cx] = origin[sx];
    } else {
      const int mm = m - top;
      origin[(long long)(pic_h + mm) * stride + cx] = origin[(long long)(pic_h - 1) * stride + sx];
    }
  }
}

cudaError_t ref_fill_launch(int16_t* d_plane, size_t samples, int value, cudaStream_t stream, int* launches)
{
  k4_fill<<<296, 256, 0, stream>>>(d_plane, samples, (int16_t)value);
  if (launches) (*launches)++;
  return cudaGetLastError();
}

cudaError_t ref_extend_launch(int16_t* d_origin, int stride, int pic_w, int pic_h, int margin,
                              int x, int y, int w, int h, cudaStream_t stream, int* launches)
{
  const bool touches = (x == 0) || (y == 0) || (x + w == pic_w) || (y + h == pic_h);
  if (!touches) return cudaSuccess;
  k4_extend<<<8, 256, 0, stream>>>(d_origin, stride, pic_w, pic_h, margin, x, y, w, h);
  if (launches) (*launches)++;
  return cudaGetLastError();
}

}  // namespace hop
