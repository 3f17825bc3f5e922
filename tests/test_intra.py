"""K7: intra mode pre-screen -- 35 x (predIntraLumaAng + calcHAD) per PU.  Oracle vs the compiled
reference's own functions and golden vectors on CPU; the CUDA path vs the oracle with -m gpu."""
import os

import numpy as np
import pytest

import conftest  # noqa: F401
import _oracle
import hevc_hop_b200 as hop
from hevc_hop_b200.workload import intra_jobs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SIZES = [4, 8, 16, 32, 64]


@pytest.mark.skipif(_oracle.ref() is None, reason="oracle/_ref/libhopref.so not built")
@pytest.mark.parametrize("bit_depth", [8, 10])
def test_oracle_intra_vs_compiled_reference(bit_depth):
    jobs, org, refs = intra_jobs(SIZES, 8, seed=11 * bit_depth, bit_depth=bit_depth)
    assert (_oracle.intra_prescreen(jobs, org, refs, "orc") == _oracle.intra_prescreen(jobs, org, refs, "ref")).all()


def test_oracle_intra_matches_golden():
    g = np.load(os.path.join(GOLD, "intra_golden.npz"))
    for k in range(int(g["n_intra"])):
        t = "i%02d" % k
        assert (_oracle.intra_prescreen(g[t + "_jobs"], g[t + "_org"], g[t + "_refs"]) == g[t + "_out"]).all(), k


def test_intra_modes_against_numpy():
    """Independent statement of three modes: vertical (26) copies the row above (32x32: no edge filter), horizontal
    (10) the column to the left, DC without neighbours repeats the default sample; cost = tiled Hadamard of the difference."""
    def hadamard(n):
        h = np.array([[1]])
        while h.shape[0] < n:
            h = np.block([[h, h], [h, -h]])
        return h
    jobs, org, refs = intra_jobs([32], 3, seed=5)
    jobs["above_avail"] = 0; jobs["left_avail"] = 0
    got = _oracle.intra_prescreen(jobs, org, refs)
    H = hadamard(8)
    for i in range(len(jobs)):
        n, sw = 32, 65
        r = refs[jobs["refs_off"][i]:jobs["refs_off"][i] + 4 * sw]
        o = org[jobs["org_off"][i]:jobs["org_off"][i] + n * n].reshape(n, n).astype(np.int64)
        # 32x32: modes 10 / 26 read the unfiltered samples (diff 0 > m_aucIntraFilter[3] = 0 is false)
        preds = {26: np.tile(r[1:n + 1], (n, 1)), 10: np.tile(r[sw + 1:sw + n + 1].reshape(n, 1), (1, n)), 1: np.full((n, n), r[sw + 1])}
        for mode, p in preds.items():
            d = o - p
            s = sum((np.abs(H @ d[y:y + 8, x:x + 8] @ H).sum() + 2) >> 2 for y in range(0, n, 8) for x in range(0, n, 8))
            assert got[i, mode] == s, (i, mode)


# ---- CUDA path ---------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("bit_depth", [8, 10])
def test_gpu_intra_matches_oracle(ctx, bit_depth):
    for seed in (1, 2, 3):
        jobs, org, refs = intra_jobs(SIZES, 12, seed=seed * 7 + bit_depth, bit_depth=bit_depth)
        assert (ctx.intra_prescreen(jobs, org, refs) == _oracle.intra_prescreen(jobs, org, refs)).all(), seed
    jobs, org, refs = intra_jobs([4], 300, seed=77, bit_depth=bit_depth)          # many tiny PUs in one launch
    assert (ctx.intra_prescreen(jobs, org, refs) == _oracle.intra_prescreen(jobs, org, refs)).all()


@pytest.mark.gpu
def test_gpu_intra_golden(ctx):
    g = np.load(os.path.join(GOLD, "intra_golden.npz"))
    for k in range(int(g["n_intra"])):
        t = "i%02d" % k
        assert (ctx.intra_prescreen(g[t + "_jobs"], g[t + "_org"], g[t + "_refs"]) == g[t + "_out"]).all(), k
