"""Generate tests/golden/*.npz from the compiled UNMODIFIED reference (oracle/_ref/libhopref.so).

Run in the build container (needs /root/reference to have been compiled by `make -C oracle ref`):
    python tests/golden/make_golden.py
Each fixture stores the seeded inputs (job arrays + sample buffers) and the outputs the reference's own
xPatternSearch / xPatternSearchGT / DistFunc / extendPicBorder produced for them.  The fixtures pin the
oracle (tests/test_oracle_cpu.py) and the CUDA path (tests/test_gpu_parity.py) on boxes where the
reference cannot be built.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import conftest  # noqa: F401  (registers hevc_hop_b200)
import _oracle
import hevc_hop_b200 as hop
from hevc_hop_b200.workload import PuBatch, dist_jobs

SHAPES = [(8, 8), (16, 16), (32, 32), (64, 64), (8, 4), (4, 8), (16, 8), (8, 16), (16, 4), (4, 16),
          (16, 12), (12, 16), (32, 16), (32, 8), (32, 24), (24, 32), (64, 32), (64, 16), (16, 64), (64, 48), (48, 64)]


def main():
    ref = _oracle.ref()
    assert ref is not None, "build oracle/_ref first: make -C oracle ref"
    out = {}
    k = 0
    for bit_depth in (8, 10):
        for (c, r) in SHAPES:
            if bit_depth == 10 and (c, r) not in [(8, 8), (16, 16), (16, 12), (32, 16), (8, 4)]:
                continue
            for use_had in (1, 0):
                if use_had == 0 and (c, r) not in [(8, 8), (16, 12), (32, 32), (8, 4)]:
                    continue
                n = 2 if c * r >= 2048 else 4
                b = PuBatch(c, r, n, seed=100 + k, bit_depth=bit_depth, sr=16, use_had=use_had, n_start=3,
                            threshold=0xFFFFFFFE if k % 3 else 1500)
                tag = "s%02d" % k
                out[tag + "_org"] = b.org
                out[tag + "_ref"] = b.ref
                out[tag + "_k1_jobs"] = b.search_jobs
                out[tag + "_k2_jobs"] = b.gt_jobs
                out[tag + "_k1_out"] = ref.pattern_search(b.search_jobs, b.org, b.ref)
                out[tag + "_k2_out"] = ref.pattern_search_gt(b.gt_jobs, b.org, b.ref)
                k += 1
    out["n_search"] = np.int64(k)
    np.savez_compressed(os.path.join(HERE, "search_golden.npz"), **out)

    dist = {}
    k = 0
    for bit_depth in (8, 10):
        for (c, r) in SHAPES + [(2, 2), (6, 2), (20, 6), (12, 8)]:
            for func, sub in ((hop.HOP_DF_HADS, 0), (hop.HOP_DF_SAD, 0), (hop.HOP_DF_SAD, 1)):
                if func == hop.HOP_DF_SAD and sub == 1 and r < 2:
                    continue
                jobs, org, cur = dist_jobs(c, r, 3, seed=500 + k, bit_depth=bit_depth, func=func, sub_shift=sub)
                tag = "d%03d" % k
                dist[tag + "_jobs"] = jobs
                dist[tag + "_org"] = org
                dist[tag + "_cur"] = cur
                dist[tag + "_out"] = ref.dist(jobs, org, cur)
                k += 1
    dist["n_dist"] = np.int64(k)
    np.savez_compressed(os.path.join(HERE, "dist_golden.npz"), **dist)

    # exhaustive sweep: outputs of the reference compiled with IT_GT_SEARCH 1 / IT_GT_GRID_SIZE 1
    rs = _oracle.ref_sweep()
    assert rs is not None, "build the sweep reference first: make -C oracle ref_sweep"
    sw = {}
    k = 0
    for bit_depth, shapes in ((8, [(8, 8), (16, 16), (8, 4), (4, 8), (16, 12), (32, 8), (32, 32)]), (10, [(8, 8), (16, 8)])):
        for (c, r) in shapes:
            for use_had in (1, 0):
                b = PuBatch(c, r, 2, seed=900 + k, bit_depth=bit_depth, sr=20, use_had=use_had, n_start=1,
                            threshold=0xFFFFFFFE if k % 2 else 3000)
                tag = "w%02d" % k
                sw[tag + "_org"], sw[tag + "_ref"], sw[tag + "_jobs"] = b.org, b.ref, b.gt_jobs
                sw[tag + "_out"] = rs.pattern_search_gt(b.gt_jobs, b.org, b.ref)
                k += 1
    sw["n_sweep"] = np.int64(k)
    np.savez_compressed(os.path.join(HERE, "sweep_golden.npz"), **sw)

    # fractional-pel refinement: outputs of the reference's xPatternSearchFracDIF
    fr = {}
    k = 0
    for bit_depth, shapes in ((8, [(8, 8), (16, 16), (8, 4), (4, 8), (16, 12), (32, 8), (32, 32), (64, 64), (24, 32)]), (10, [(8, 8), (16, 8)])):
        for (c, r) in shapes:
            for use_had in (1, 0):
                b = PuBatch(c, r, 3, seed=700 + k, bit_depth=bit_depth, sr=20, use_had=use_had, n_start=1)
                tag = "f%02d" % k
                fj = b.frac_jobs()
                fr[tag + "_org"], fr[tag + "_ref"], fr[tag + "_jobs"] = b.org, b.ref, fj
                fr[tag + "_out"] = _oracle.frac_search(fj, b.org, b.ref, "ref")
                k += 1
    fr["n_frac"] = np.int64(k)
    np.savez_compressed(os.path.join(HERE, "frac_golden.npz"), **fr)

    # border extension: random plane with -1 staircase, reference extendPicBorder (margin 80)
    rng = np.random.default_rng(7)
    pic_w, pic_h, m = 96, 72, 80
    plane = rng.integers(-1, 256, size=(pic_h + 2 * m, pic_w + 2 * m)).astype(np.int16)
    before = plane.copy()
    mm = ref.lib.ref_extend_border(plane.ctypes.data, pic_w, pic_h)
    assert mm == m
    np.savez_compressed(os.path.join(HERE, "border_golden.npz"), before=before, after=plane,
                        pic_w=np.int64(pic_w), pic_h=np.int64(pic_h), margin=np.int64(m))
    print("golden fixtures written:", k, "dist cases")


if __name__ == "__main__":
    main()
