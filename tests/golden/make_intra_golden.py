"""tests/golden/intra_golden.npz: inputs and outputs of the compiled UNMODIFIED reference's intra pre-screen
(predIntraLumaAng for the 35 modes + calcHAD, TComPrediction.cpp:129-348, 1468-1546; TComRdCost.cpp:391-425), driven
through oracle/ref_harness.cpp.  Run in the build container:  python tests/golden/make_intra_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import conftest  # noqa: F401
import _oracle
from hevc_hop_b200.workload import intra_jobs


def main():
    assert _oracle.ref() is not None, "build oracle/_ref first: make -C oracle ref"
    out = {}
    for k, bd in enumerate((8, 10)):
        jobs, org, refs = intra_jobs([4, 8, 16, 32, 64], 5, seed=60 + k, bit_depth=bd)
        t = "i%02d" % k
        out[t + "_jobs"], out[t + "_org"], out[t + "_refs"] = jobs, org, refs
        out[t + "_out"] = _oracle.intra_prescreen(jobs, org, refs, "ref")
        print(t, "bit depth", bd, "jobs", len(jobs), "cost sum", int(out[t + "_out"].sum()))
    out["n_intra"] = np.int64(2)
    np.savez_compressed(os.path.join(HERE, "intra_golden.npz"), **out)


if __name__ == "__main__":
    main()
