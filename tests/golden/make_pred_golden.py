"""tests/golden/pred_golden.npz: inputs and outputs of the compiled UNMODIFIED reference's motion compensation
(xPredInterLumaBlk / xPredInterChromaBlk with the GT branches, TComPrediction.cpp:639-805, 1235-1420) and of the
distortion / AMVP template cost computed from it (TEncSearch.cpp:2951-2977, 4390-4477), driven through
oracle/ref_harness.cpp.  Run in the build container:  python tests/golden/make_pred_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import conftest  # noqa: F401
import _oracle
from hevc_hop_b200.workload import PredBatch

SHAPES = [(8, 8), (16, 16), (32, 32), (64, 64), (8, 4), (4, 8), (16, 12), (32, 8), (24, 32), (64, 16)]
CASES = [dict(comp=0, kind="gt", bit_depth=8, with_invalid=True), dict(comp=1, kind="gt", bit_depth=8, with_invalid=False),
         dict(comp=0, kind="dist", bit_depth=10, with_invalid=False), dict(comp=0, kind="template", bit_depth=8, with_invalid=True),
         dict(comp=1, kind="plain", bit_depth=10, with_invalid=True), dict(comp=0, kind="plain", bit_depth=8, with_invalid=False)]


def main():
    assert _oracle.ref() is not None, "build oracle/_ref first: make -C oracle ref"
    out = {}
    for k, case in enumerate(CASES):
        shapes = [s for s in SHAPES if not case["comp"] or (s[0] >= 8 and s[1] >= 8)]
        b = PredBatch(shapes, 2, seed=300 + k, **case)
        res, dst = _oracle.predict(b.jobs, b.org, b.ref, b.dst_samples, "ref")
        t = "p%02d" % k
        out[t + "_jobs"], out[t + "_org"], out[t + "_ref"], out[t + "_out"], out[t + "_dst"] = b.jobs, b.org, b.ref, res, dst
        print(t, case, "dist sum", int(res["dist"].sum()), "dst mean %.1f" % dst.mean(), "valid", int(res["valid"].sum()), "/", len(res))
    out["n_pred"] = np.int64(len(CASES))
    np.savez_compressed(os.path.join(HERE, "pred_golden.npz"), **out)


if __name__ == "__main__":
    main()
