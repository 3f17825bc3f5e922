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


def _cu_mirror(ctx, seed, pic_w=192, pic_h=160, m=80, px=96, py=96):
    """A picture whose causal area (rows above (px, py) and the band to its left) is coded; returns host plane + stride."""
    rng = np.random.default_rng(seed)
    ctx.ref_create(pic_w, pic_h, m)
    host = np.full((pic_h + 2 * m, pic_w + 2 * m), -1, dtype=np.int16)
    img = rng.integers(0, 256, size=(pic_h, pic_w)).astype(np.int16)
    host[m:m + py, m:m + pic_w] = img[:py]
    host[m + py:m + pic_h, m:m + px] = img[py:, :px]
    _oracle.extend_border_oracle(host, pic_w, pic_h, m)
    ctx.ref_upload(host)
    return host, pic_w + 2 * m, m, img


def _cu_jobs(stride, seed, px=96, py=96):
    """The first-PU requests of one 32x32 CU's partition modes (2Nx2N, Nx2N, 2NxN, four AMP shapes) + two of an 8x8 CU."""
    out = []
    for k, (c, r) in enumerate([(32, 32), (16, 32), (32, 16), (32, 8), (32, 24), (8, 32), (24, 32), (8, 8), (4, 8), (8, 4)]):
        b = PuBatch(c, r, 1, seed=seed * 50 + k, sr=16, n_start=3)
        mj = b.motion_jobs()
        s = mj["search"]
        s["ref_stride"] = stride; s["ref_off"] = py * stride + px
        s["rng_left"], s["rng_right"], s["rng_top"], s["rng_bottom"] = -70, 40, -70, -4
        mj["search"] = s
        mj["amvp"]["hor"][:, 0] = -40 * 4 + k; mj["amvp"]["ver"][:, 0] = -50 * 4
        mj["amvp"]["hor"][:, 1] = -12 * 4; mj["amvp"]["ver"][:, 1] = -30 * 4 - k
        out.append((mj, b.org.copy()))
    return out


def _want(mj, org, host, stride, m):
    m2 = mj.copy(); s2 = m2["search"]; s2["ref_off"] += m * stride + m; m2["search"] = s2
    return _oracle.motion_search(m2, org, host.reshape(-1))


@pytest.mark.gpu
def test_gpu_speculative_searches_hit_only_identical_requests(ctx):
    """hop_motion_search_prefetch + hop_motion_search_batch(n = 1, ref = NULL): a request is answered from a
    speculative launch only when job, block and SS-mirror state are identical; everything else launches afresh;
    every answer equals the oracle."""
    host, stride, m, img = _cu_mirror(ctx, seed=21)
    for rep in range(6):
        jobs = _cu_jobs(stride, seed=rep)
        st0 = ctx.stats()
        for mj, org in jobs:                                   # the speculation window of a CU
            ctx.motion_prefetch(mj, org)
        order = list(range(len(jobs)))
        if rep % 2:
            order.reverse()                                    # consumed in any order
        for i in order:
            mj, org = jobs[i]
            assert ctx.motion_search(mj, org, None).tobytes() == _want(mj, org, host, stride, m).tobytes(), (rep, i)
        st1 = ctx.stats()
        assert st1["cache_hits"] - st0["cache_hits"] == len(jobs) and st1["cache_misses"] == st0["cache_misses"]
        assert st1["prefetched"] - st0["prefetched"] == len(jobs)
    # near misses: one job field, one block sample
    jobs = _cu_jobs(stride, seed=40)
    for mj, org in jobs:
        ctx.motion_prefetch(mj, org)
    st0 = ctx.stats()
    mj, org = jobs[0]
    mj2 = mj.copy(); c2 = mj2["search"]["cost"]; c2["pred"]["hor"] += 4; s2 = mj2["search"]; s2["cost"] = c2; mj2["search"] = s2
    assert ctx.motion_search(mj2, org, None).tobytes() == _want(mj2, org, host, stride, m).tobytes()
    org2 = org.copy(); org2[5] ^= 1
    mj, _ = jobs[1]
    assert ctx.motion_search(mj, org2, None).tobytes() == _want(mj, org2, host, stride, m).tobytes()
    st1 = ctx.stats()
    assert st1["cache_misses"] - st0["cache_misses"] == 2 and st1["cache_hits"] == st0["cache_hits"]
    # the mirror changes (a CU is committed): everything enqueued before is stale
    blk = img[96:104, 96:104].copy()
    ctx.ref_update(96, 96, blk)
    host[m + 96:m + 104, m + 96:m + 104] = blk
    _oracle.extend_border_oracle(host, 192, 160, m)
    st0 = ctx.stats()
    for i in (2, 3, 7):
        mj, org = jobs[i]
        assert ctx.motion_search(mj, org, None).tobytes() == _want(mj, org, host, stride, m).tobytes(), i
    st1 = ctx.stats()
    assert st1["cache_misses"] - st0["cache_misses"] == 3 and st1["cache_hits"] == st0["cache_hits"]
    # more requests than slots: the oldest are dropped, the answers stay right
    many = _cu_jobs(stride, seed=41) + _cu_jobs(stride, seed=42)
    for mj, org in many:
        ctx.motion_prefetch(mj, org)
    for mj, org in many:
        assert ctx.motion_search(mj, org, None).tobytes() == _want(mj, org, host, stride, m).tobytes()
    assert ctx.stats()["prefetch_dropped"] > 0
    # a duplicate request inside one window is enqueued once
    st0 = ctx.stats()
    mj, org = many[0]
    ctx.motion_prefetch(mj, org); ctx.motion_prefetch(mj, org)
    assert ctx.stats()["prefetched"] - st0["prefetched"] == 1
    assert ctx.motion_search(mj, org, None).tobytes() == _want(mj, org, host, stride, m).tobytes()


@pytest.mark.gpu
def test_gpu_two_contexts_on_two_devices_in_one_process():
    """One process, one context per GPU (the ABI allows it): kernels that opt in to more than 48 KB of dynamic shared
    memory must be configured on EVERY device they run on (per-device function attributes)."""
    lib = hop.load_library()
    if lib.hop_device_count() < 2:
        pytest.skip("needs two GPUs")
    ctxs = [hop.HopContext(d) for d in (0, 1)]
    try:
        for c, r in [(64, 64), (32, 32), (8, 8)]:
            b = PuBatch(c, r, 3, seed=c + r, sr=128 if c == 64 else 32, n_start=3)
            want_k1 = _oracle.oracle().pattern_search(b.search_jobs, b.org, b.ref)
            want_k2 = _oracle.oracle().pattern_search_gt(b.gt_jobs, b.org, b.ref)
            mj = b.motion_jobs()
            want_m = _oracle.motion_search(mj, b.org, b.ref)
            for cx in ctxs:
                assert cx.pattern_search(b.search_jobs, b.org, b.ref).tobytes() == want_k1.tobytes()
                assert cx.pattern_search_gt(b.gt_jobs, b.org, b.ref).tobytes() == want_k2.tobytes()
                assert cx.motion_search(mj, b.org, b.ref).tobytes() == want_m.tobytes()
                assert cx.gt_sweep(b.gt_jobs[:1], b.org, b.ref).tobytes() == _oracle.gt_sweep(b.gt_jobs[:1], b.org, b.ref).tobytes()
    finally:
        for cx in ctxs:
            cx.close()
