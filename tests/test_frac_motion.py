"""K5 (xPatternSearchFracDIF) and the fused motion search K1 -> K5 -> K2: oracle vs the compiled reference and
its golden vectors on CPU; the CUDA path vs the oracle with -m gpu."""
import os

import numpy as np
import pytest

import conftest  # noqa: F401
import _oracle
import hevc_hop_b200 as hop
from hevc_hop_b200.workload import PuBatch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SHAPES = [(8, 8), (16, 16), (32, 32), (64, 64), (8, 4), (4, 8), (16, 8), (16, 4), (16, 12), (12, 16), (32, 24), (64, 32)]


def frac_same(a, b, extras=True):
    ok = (a["cost"] == b["cost"]).all() and a["half"].tobytes() == b["half"].tobytes() and a["qter"].tobytes() == b["qter"].tobytes()
    if extras:
        ok &= (a["cost_half"] == b["cost_half"]).all()
    return bool(ok)


def test_oracle_frac_matches_golden():
    g = np.load(os.path.join(GOLD, "frac_golden.npz"))
    for k in range(int(g["n_frac"])):
        t = "f%02d" % k
        assert frac_same(_oracle.frac_search(g[t + "_jobs"], g[t + "_org"], g[t + "_ref"]), g[t + "_out"], extras=False), k


@pytest.mark.skipif(_oracle.ref() is None, reason="oracle/_ref/libhopref.so not built")
@pytest.mark.parametrize("shape", [(8, 8), (16, 12), (4, 8), (32, 16), (48, 64)])
@pytest.mark.parametrize("bit_depth", [8, 10])
def test_oracle_frac_vs_compiled_reference(shape, bit_depth):
    c, r = shape
    for use_had in (1, 0):
        b = PuBatch(c, r, 3, seed=c * 3 + r + bit_depth, bit_depth=bit_depth, sr=20, use_had=use_had, n_start=1)
        fj = b.frac_jobs()
        assert frac_same(_oracle.frac_search(fj, b.org, b.ref, "orc"), _oracle.frac_search(fj, b.org, b.ref, "ref"), extras=False)


def test_oracle_motion_chain_is_the_three_stages():
    """orc_motion_search == xPatternSearch, then FracDIF on its vector, then GT with the frac cost as threshold."""
    b = PuBatch(8, 8, 4, seed=12, sr=32, n_start=3)
    m = _oracle.motion_search(b.motion_jobs(), b.org, b.ref)
    orc = _oracle.oracle()
    k1 = orc.pattern_search(b.search_jobs, b.org, b.ref)
    assert m["search"].tobytes() == k1.tobytes() and (m["refined"] == 1).all()
    fr = _oracle.frac_search(b.frac_jobs(mv_int=k1["mv"]), b.org, b.ref)
    assert frac_same(m["frac"], fr)
    gj = b.gt_jobs.copy(); gj["ss_cand"] = k1["mv"]; gj["threshold"] = fr["cost"]
    assert m["gt"].tobytes() == orc.pattern_search_gt(gj, b.org, b.ref).tobytes()


# ---- CUDA path ---------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES)
def test_gpu_frac_matches_oracle(ctx, shape):
    c, r = shape
    n = 2 if c * r >= 2048 else 5
    for bit_depth in (8, 10):
        for use_had in (1, 0):
            b = PuBatch(c, r, n, seed=c * 11 + r + use_had, bit_depth=bit_depth, sr=20, use_had=use_had, n_start=1)
            fj = b.frac_jobs()
            assert frac_same(ctx.frac_search(fj, b.org, b.ref), _oracle.frac_search(fj, b.org, b.ref))


@pytest.mark.gpu
def test_gpu_frac_golden(ctx):
    g = np.load(os.path.join(GOLD, "frac_golden.npz"))
    for k in range(int(g["n_frac"])):
        t = "f%02d" % k
        assert frac_same(ctx.frac_search(g[t + "_jobs"], g[t + "_org"], g[t + "_ref"]), g[t + "_out"], extras=False), k


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES)
def test_gpu_motion_search_matches_oracle(ctx, shape):
    c, r = shape
    n = 2 if c * r >= 2048 else 4
    for use_gt in (1, 0):
        b = PuBatch(c, r, n, seed=c * 13 + r + use_gt, sr=max(32, r + 16), n_start=3)
        mj = b.motion_jobs(use_gt=use_gt)
        assert _oracle.motion_search(mj, b.org, b.ref)["refined"].any()
        got, want = ctx.motion_search(mj, b.org, b.ref), _oracle.motion_search(mj, b.org, b.ref)
        assert got.tobytes() == want.tobytes()
    # searches that stop after the integer stage: nothing valid / zero vector only
    b = PuBatch(c, r, 2, seed=3, sr=20)
    mj = b.motion_jobs()
    mj["search"]["rng_top"] = mj["search"]["rng_bottom"] + 1
    assert ctx.motion_search(mj, b.org, b.ref).tobytes() == _oracle.motion_search(mj, b.org, b.ref).tobytes()
    mj = b.motion_jobs()
    for k in ("rng_left", "rng_right", "rng_top", "rng_bottom"):
        mj["search"][k] = 0
    mj["search"]["is_ss"] = 0
    got = ctx.motion_search(mj, b.org, b.ref)
    assert (got["refined"] == 0).all() and got.tobytes() == _oracle.motion_search(mj, b.org, b.ref).tobytes()


@pytest.mark.gpu
def test_gpu_motion_search_on_the_mirror_single_call(ctx):
    """The encoder's call shape: n == 1, ref == NULL (zero-copy job, device mirror)."""
    rng = np.random.default_rng(5)
    pic_w, pic_h, m = 128, 128, 80
    ctx.ref_create(pic_w, pic_h, m)
    host = np.full((pic_h + 2 * m, pic_w + 2 * m), -1, dtype=np.int16)
    img = rng.integers(0, 256, size=(pic_h, pic_w)).astype(np.int16)
    host[m:m + 64, m:m + 128] = img[:64]
    host[m + 64:m + 128, m:m + 64] = img[64:, :64]
    _oracle.extend_border_oracle(host, pic_w, pic_h, m)
    ctx.ref_upload(host)
    stride = pic_w + 2 * m
    # multi-tile shapes take the thread-block-cluster form of the latency path (2, 4 or 8 CTAs per PU)
    # every PU shape of the encoder: single CTA, row-per-lane tiles, clusters of 2/4/8, inline and mapped blocks
    every = [(8, 8), (16, 16), (32, 32), (64, 64), (8, 4), (4, 8), (16, 8), (8, 16), (16, 4), (4, 16), (16, 12), (12, 16),
             (32, 16), (16, 32), (32, 8), (8, 32), (32, 24), (24, 32), (64, 32), (32, 64), (64, 16), (16, 64), (64, 48), (48, 64)]
    for (c, r), seed in [(sh, sd) for sh in every for sd in (0, 1)]:
        if not hop.load_library().hop_shape_supported(c, r):
            continue
        b = PuBatch(c, r, 1, seed=c + 100 * seed, sr=16, n_start=2 + seed)
        mj = b.motion_jobs()
        s = mj["search"]
        s["ref_stride"] = stride; s["ref_off"] = 64 * stride + 64
        s["rng_left"], s["rng_right"], s["rng_top"], s["rng_bottom"] = -60, 30, -60, -4
        mj["search"] = s
        mj["amvp"]["hor"][:, 0] = -40 * 4; mj["amvp"]["ver"][:, 0] = -50 * 4
        mj["amvp"]["hor"][:, 1] = 0; mj["amvp"]["ver"][:, 1] = 0
        got = ctx.motion_search(mj, b.org, None)
        m2 = mj.copy(); s2 = m2["search"]; s2["ref_off"] += m * stride + m; m2["search"] = s2
        want = _oracle.motion_search(m2, b.org, host.reshape(-1))
        assert got.tobytes() == want.tobytes(), (c, r)
        # the split calls of the same PU (K2 alone through the cluster form)
        if want["refined"][0]:
            gj = b.gt_jobs.copy()
            gj["ref_stride"] = stride; gj["ref_off"] = 64 * stride + 64
            gj["ss_cand"] = want["search"]["mv"]; gj["threshold"] = want["frac"]["cost"]; gj["amvp"] = mj["amvp"]
            assert ctx.pattern_search_gt(gj, b.org, None).tobytes() == want["gt"].tobytes(), (c, r)


@pytest.mark.gpu
def test_gpu_single_call_random_start_vectors(ctx):
    """Single-PU fused calls with random AMVP lists: 0-3 predictors, zero vectors (skipped by the reference), vectors
    that leave the valid region, the windows-staged-ahead layout and its fall-back, with and without Hadamard."""
    rng = np.random.default_rng(11)
    pic_w, pic_h, m = 192, 160, 80
    ctx.ref_create(pic_w, pic_h, m)
    host = np.full((pic_h + 2 * m, pic_w + 2 * m), -1, dtype=np.int16)
    img = rng.integers(0, 256, size=(pic_h, pic_w)).astype(np.int16)
    host[m:m + 96, m:m + pic_w] = img[:96]                       # causal region: rows above, and the row band left of the PU
    host[m + 96:m + pic_h, m:m + 96] = img[96:, :96]
    _oracle.extend_border_oracle(host, pic_w, pic_h, m)
    ctx.ref_upload(host)
    stride = pic_w + 2 * m
    shapes = [(8, 4), (4, 8), (8, 8), (16, 8), (8, 16), (16, 16), (16, 12), (32, 16), (32, 32), (32, 64), (64, 32), (64, 64)]
    for it in range(72):
        c, r = shapes[it % len(shapes)]
        b = PuBatch(c, r, 1, seed=500 + it, sr=16, n_start=1)
        mj = b.motion_jobs()
        s = mj["search"]
        s["ref_stride"] = stride; s["ref_off"] = 96 * stride + 96   # PU at (96, 96) of the picture
        s["rng_left"], s["rng_right"], s["rng_top"], s["rng_bottom"] = -70, 40, -70, -4
        mj["search"] = s
        mj["num_pred"] = int(rng.integers(0, 4))
        for k in range(3):
            kind = rng.integers(0, 4)
            if kind == 0:
                hv = (0, 0)                                       # skipped start (:5144)
            elif kind == 1:
                hv = (int(rng.integers(-90, -10)) * 4, int(rng.integers(-90, -10)) * 4)
            elif kind == 2:
                hv = (int(rng.integers(-300, 150)), int(rng.integers(-300, 150)))   # raw quarter-pel, anywhere inside the plane (also into NOT_VALID samples)
            else:
                hv = (int(rng.integers(-8, 8)), int(rng.integers(-8, 8)))           # rounds to a (near) zero integer vector
            mj["amvp"]["hor"][:, k] = hv[0]; mj["amvp"]["ver"][:, k] = hv[1]
        mj["use_had"] = int(it % 3 != 0)
        got = ctx.motion_search(mj, b.org, None)
        m2 = mj.copy(); s2 = m2["search"]; s2["ref_off"] += m * stride + m; m2["search"] = s2
        want = _oracle.motion_search(m2, b.org, host.reshape(-1))
        assert got.tobytes() == want.tobytes(), (it, c, r, int(mj["num_pred"][0]))
