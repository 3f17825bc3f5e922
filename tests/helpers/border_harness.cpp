// Test helper: C entry point for include/hop_border.h (the host-side incremental border re-extension the encoder
// shim uses), so that the CPU suite can compare it with the full re-extension of the oracle.
#include <stdint.h>
#include "hop_border.h"

extern "C" void border_extend_patch(int16_t* origin, int stride, int pic_w, int pic_h, int margin, int x, int y, int w, int h)
{
  hop_extend_patch_border<int16_t>(origin, stride, pic_w, pic_h, margin, margin, x, y, w, h);
}
