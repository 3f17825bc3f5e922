"""Phase timeline of the fused single-PU motion search (the encoder's call), from a HOP_TRACE build.

  tools/build_variant.sh trace -DHOP_TRACE
  HOP_LIB=hevc-hop_b200/build/variants/libhopgpu_trace.so python tools/latency_trace.py

Host stamps are CLOCK_MONOTONIC, device stamps %globaltimer: only differences inside one clock are meaningful.
"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import conftest  # noqa: F401
import hevc_hop_b200 as hop
from hevc_hop_b200.workload import PuBatch


def main():
    ctx = hop.HopContext(0)
    fn = ctx.lib.hop_debug_trace
    fn.argtypes = [C.c_void_p, C.c_void_p]
    fn.restype = C.c_int
    buf = np.zeros(8 + 128, dtype=np.uint64)
    for c, r in [(8, 4), (8, 8), (16, 16), (32, 32)]:
        b = PuBatch(c, r, 1, seed=5, sr=64, n_start=3)
        mj = b.motion_jobs()
        ctx.ref_create(b.pw, b.ph, 0)
        ctx.ref_upload(np.ascontiguousarray(b.ref).reshape(b.ph, b.pw))
        rows = []
        for it in range(60):
            res = ctx.motion_search(mj, b.org, None)
            assert fn(ctx.h, buf.ctypes.data) == 0
            h, k1, k2 = buf[:8].astype(np.int64), buf[8:72].astype(np.int64), buf[72:].astype(np.int64)
            if it < 10:
                continue
            row = {"host pack": h[1] - h[0], "host launch k1": h[2] - h[1], "host launch tail": h[3] - h[2], "host wait": h[4] - h[3],
                   "host total": h[4] - h[0],
                   "k1 job+tables": k1[1] - k1[0], "k1 staging": k1[2] - k1[1], "k1 staircase": k1[3] - k1[2],
                   "k1 search": k1[4] - k1[3], "k1 reduce": k1[5] - k1[4], "k1 finalize (last slice)": k1[6] - k1[5],
                   "k1 total": k1[6] - k1[0], "gap k1 -> tail": k2[0] - k1[6],
                   "tail prologue + wait for k1": k2[1] - k2[0], "k1 end -> tail released": k2[1] - k1[6],
                   "tail frac": k2[2] - k2[1], "tail gt": k2[3] - k2[2], "tail total": k2[3] - k2[0],
                   "gpu span": k2[3] - k1[0]}
            row["fine: barrier -> argmin done"] = k2[56] - k2[8 + 2]
            row["fine: argmin done -> table math start"] = k2[57] - k2[56]
            row["fine: table math + stores"] = k2[58] - k2[57]
            row["fine: stores -> barrier passed"] = k2[8 + 3] - k2[58]
            for bb in range(3):
                base = 8 + 16 * bb
                if k2[base] >= k2[2]:
                    row["gt start %d window" % bb] = k2[base + 1] - k2[base]
                    p = 0
                    while p < 6 and k2[base + 2 + 2 * p] > k2[base + 1 + 2 * p] >= k2[base]:
                        row["gt start %d pass %d eval" % (bb, p)] = k2[base + 2 + 2 * p] - k2[base + 1 + 2 * p]
                        if p + 1 < 6 and k2[base + 3 + 2 * p] > k2[base + 2 + 2 * p]:
                            row["gt start %d pass %d argmin+table" % (bb, p)] = k2[base + 3 + 2 * p] - k2[base + 2 + 2 * p]
                        p += 1
            rows.append(row)
        print("== %dx%d  (found=%d refined=%d gt_flag=%d n_cand=%d)" % (c, r, res["search"]["found"][0], res["refined"][0],
                                                                    res["gt"]["gt_flag"][0], res["gt"]["n_candidates"][0]))
        keys = list(rows[-1].keys())
        for k in keys:
            v = [rw[k] for rw in rows if k in rw]
            print("   %-32s %8.2f us" % (k, float(np.median(v)) / 1e3))


if __name__ == "__main__":
    main()
