"""K6: motion-compensated prediction incl. the GT branches, the inter prediction error and the AMVP
template cost -- oracle vs the compiled reference's own xPredInterLumaBlk / xPredInterChromaBlk / isValidPattern /
getDistPart / calcRdCost (CPU), golden vectors, and the CUDA path vs the oracle (-m gpu)."""
import os

import numpy as np
import pytest

import conftest  # noqa: F401
import _oracle
import hevc_hop_b200 as hop
from hevc_hop_b200.workload import PredBatch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SHAPES = [(8, 8), (16, 16), (32, 32), (64, 64), (8, 4), (4, 8), (16, 8), (16, 4), (16, 12), (12, 16), (32, 24), (64, 32), (48, 64)]
CASES = [(comp, kind, bd, inv) for comp in (0, 1) for kind in ("plain", "gt", "dist") for bd in (8, 10) for inv in (False, True)] + \
        [(0, "template", bd, inv) for bd in (8, 10) for inv in (False, True)]


def batch_for(case, seed, shapes=SHAPES, n=2):
    comp, kind, bd, inv = case
    if comp:
        shapes = [s for s in shapes if s[0] >= 8 and s[1] >= 8]      # chroma blocks of at least 4x4
    return PredBatch(shapes, n, seed=seed, bit_depth=bd, comp=comp, kind=kind, with_invalid=inv)


def same(a, b):
    (ra, da), (rb, db) = a, b
    return ra.tobytes() == rb.tobytes() and da.tobytes() == db.tobytes()


@pytest.mark.skipif(_oracle.ref() is None, reason="oracle/_ref/libhopref.so not built")
@pytest.mark.parametrize("case", CASES)
def test_oracle_predict_vs_compiled_reference(case):
    b = batch_for(case, seed=17 + CASES.index(case))
    want = _oracle.predict(b.jobs, b.org, b.ref, b.dst_samples, "ref")
    got = _oracle.predict(b.jobs, b.org, b.ref, b.dst_samples, "orc")
    assert (got[0]["valid"] == want[0]["valid"]).all()
    assert (got[0]["dist"] == want[0]["dist"]).all() and (got[0]["cost"] == want[0]["cost"]).all()
    assert got[1].tobytes() == want[1].tobytes()


def test_oracle_predict_matches_golden():
    g = np.load(os.path.join(GOLD, "pred_golden.npz"))
    for k in range(int(g["n_pred"])):
        t = "p%02d" % k
        res, dst = _oracle.predict(g[t + "_jobs"], g[t + "_org"], g[t + "_ref"], int(g[t + "_dst"].size))
        assert res.tobytes() == g[t + "_out"].tobytes() and dst.tobytes() == g[t + "_dst"].tobytes(), k


def test_template_gate_and_gt_zero_vectors():
    """Edge cases: a candidate whose probe hits NOT_VALID costs MAX_INT and is not predicted; gt_flag with all-zero
    vectors is the plain branch; no-GT integer vectors copy NOT_VALID samples unchanged (filterCopy first == last)."""
    b = PredBatch([(8, 8)], 1, seed=3, kind="template", with_invalid=True)
    j = b.jobs.copy()
    j["is_ss"] = 1
    j["ref_off"] = 170 * 640 + 220                                                 # PU at (220, 170) of the 640 x 400 plane
    j["mv_probe"]["hor"], j["mv_probe"]["ver"] = 150 * 4, 60 * 4                   # probes land in the uncoded quadrant
    res, _ = _oracle.predict(j, b.org, b.ref, b.dst_samples)
    assert res["valid"][0] == 0 and res["cost"][0] == 0x7FFFFFFF
    b = PredBatch([(8, 8)], 1, seed=4, kind="plain", frac=False, with_invalid=True)
    j = b.jobs.copy()
    j["ref_off"] = 170 * 640 + 220
    j["mv"]["hor"], j["mv"]["ver"] = 120 * 4, 40 * 4
    _, dst = _oracle.predict(j, b.org, b.ref, b.dst_samples)
    assert (dst == -1).any()


# ---- CUDA path ---------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_predict_matches_oracle(ctx, case):
    for seed in (5, 6):
        b = batch_for(case, seed=seed * 31 + CASES.index(case), n=3)
        assert same(ctx.predict(b.jobs, b.org, b.ref, b.dst_samples), _oracle.predict(b.jobs, b.org, b.ref, b.dst_samples)), (case, seed)


@pytest.mark.gpu
def test_gpu_predict_golden(ctx):
    g = np.load(os.path.join(GOLD, "pred_golden.npz"))
    for k in range(int(g["n_pred"])):
        t = "p%02d" % k
        res, dst = ctx.predict(g[t + "_jobs"], g[t + "_org"], g[t + "_ref"], int(g[t + "_dst"].size))
        assert res.tobytes() == g[t + "_out"].tobytes() and dst.tobytes() == g[t + "_dst"].tobytes(), k


@pytest.mark.gpu
def test_gpu_predict_of_the_searched_candidate_reproduces_the_search_cost(ctx):
    """Consistency of K2 and K6: predicting the GT the diamond search accepted and taking the Hadamard cost gives the
    distortion part of the search's cost whenever the window holds no NOT_VALID sample (the search clamps those to 0,
    motion compensation copies them)."""
    from hevc_hop_b200.workload import PuBatch
    b = PuBatch(16, 16, 6, seed=8, sr=32, n_start=1)
    gt = ctx.pattern_search_gt(b.gt_jobs, b.org, b.ref)
    assert gt["gt_flag"].any()
    orc = _oracle.oracle()
    for i in np.nonzero(gt["gt_flag"])[0]:
        j = np.zeros(1, dtype=hop.PRED_JOB_DT)
        gj = b.gt_jobs[i]
        for k in ("ref_off", "ref_stride", "org_off", "org_stride", "cols", "rows", "bit_depth"):
            j[k] = gj[k]
        j["dst_off"] = -1
        j["mv"]["hor"], j["mv"]["ver"] = int(gt["mv_int"]["hor"][i]) * 4, int(gt["mv_int"]["ver"][i]) * 4
        j["gt_flag"] = 1
        j["gt"] = gt["gt"][i]
        j["dist_func"] = hop.HOP_DF_HADS
        res, _ = ctx.predict(j, b.org, b.ref, 0)
        cs = gj["cost"].copy()
        bits = orc.lib.orc_get_bits_gt(*[int(v) for v in (gt["gt"][i]["hor"][0], gt["gt"][i]["ver"][0], gt["gt"][i]["hor"][1],
                                                              gt["gt"][i]["ver"][1], gt["gt"][i]["hor"][2], gt["gt"][i]["ver"][2])])
        lam = int(cs["lambda_cost"])
        mvc = orc.lib.orc_get_cost_xy(np.ascontiguousarray(cs).ctypes.data, int(j["mv"]["hor"][0]), int(j["mv"]["ver"][0]))
        assert int(res["dist"][0]) + mvc + ((lam * bits) >> 16) == int(gt["cost"][i])
