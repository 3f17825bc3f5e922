/*
 * hop_border.h -- incremental border re-extension of a picture plane (host side of K4).
 *
 * TEncCu::xCopyYuv2SSRef (TLibEncoder/TEncCu.cpp:1694-1696) re-extends the WHOLE border of the SS reference after
 * every CU it copies (TComPicYuv::xExtendPicCompBorder, TLibCommon/TComPicYuv.cpp:247-274: left/right margins
 * replicate each row's edge sample, the top/bottom margin rows replicate the first/last picture row INCLUDING its
 * side margins).  On a plane whose border was complete before the block [x, x+w) x [y, y+h) changed, only the
 * margin samples that are copies of block samples can differ afterwards; this function rewrites exactly those, so
 * the plane ends up identical to what the full re-extension leaves -- at a cost that depends on the block, not on
 * the picture (the full pass touches 6.4 MB per CU on a 7728x5368 picture, ~85 times per CTU).
 * The device mirror does the same in k4_extend (hevc-hop_b200/csrc/k4_ref.cu).  Plain C++, no encoder types.
 */
#ifndef HOP_BORDER_H
#define HOP_BORDER_H

#include <cstring>

template <typename T>
inline void hop_extend_patch_border(T* origin, int stride, int pic_w, int pic_h, int margin_x, int margin_y,
                                    int x, int y, int w, int h)
{
  const bool left = (x == 0), right = (x + w == pic_w);
  if (left || right) {
    for (int r = y; r < y + h; r++) {
      T* row = origin + (long)r * stride;
      if (left)  { const T v = row[0];         for (int k = 1; k <= margin_x; k++) row[-k] = v; }
      if (right) { const T v = row[pic_w - 1]; for (int k = 0; k < margin_x; k++) row[pic_w + k] = v; }
    }
  }
  const int ex0 = left ? -margin_x : x;                       /* extended column range fed by this block */
  const int ex1 = right ? pic_w + margin_x : x + w;
  const size_t bytes = sizeof(T) * (size_t)(ex1 - ex0);
  if (y == 0) {
    const T* src = origin + ex0;
    for (int m = 1; m <= margin_y; m++) std::memcpy(origin - (long)m * stride + ex0, src, bytes);
  }
  if (y + h == pic_h) {
    const T* src = origin + (long)(pic_h - 1) * stride + ex0;
    for (int m = 1; m <= margin_y; m++) std::memcpy(origin + (long)(pic_h - 1 + m) * stride + ex0, src, bytes);
  }
}

#endif /* HOP_BORDER_H */
