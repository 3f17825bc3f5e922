"""GPU suite (-m gpu): the CUDA path, called through the C ABI (libhopgpu.so), against the oracle on
the same seeded inputs (bit-exact: integer / index work), against the committed golden vectors of the
compiled reference, and -- at full microbench sizes -- through size-independent properties."""
import os

import numpy as np
import pytest

import _oracle
import hevc_hop_b200 as hop
from hevc_hop_b200.workload import PuBatch, dist_jobs, gt_passes

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

ALL_SHAPES = [(8, 8), (16, 16), (32, 32), (64, 64), (8, 4), (4, 8), (16, 8), (8, 16), (16, 4), (4, 16), (16, 12),
              (12, 16), (32, 16), (16, 32), (32, 8), (8, 32), (32, 24), (24, 32), (64, 32), (32, 64), (64, 16),
              (16, 64), (64, 48), (48, 64)]


def gt_assert_equal(a, b, extras=True):
    for k in ("gt_flag", "cost") + (("best_index", "n_candidates") if extras else ()):
        assert (a[k] == b[k]).all(), (k, a[k], b[k])
    assert a["gt"].tobytes() == b["gt"].tobytes()
    assert a["mv_int"].tobytes() == b["mv_int"].tobytes()


def test_library_loaded_and_device_is_blackwell(ctx):
    assert ctx.lib.hop_abi_version() == 2
    assert ctx.lib.hop_device_count() >= 1


@pytest.mark.parametrize("shape", ALL_SHAPES)
def test_k1_matches_oracle(ctx, shape):
    c, r = shape
    orc = _oracle.oracle()
    b = PuBatch(c, r, 4, seed=c * 31 + r, sr=24)
    got = ctx.pattern_search(b.search_jobs, b.org, b.ref)
    want = orc.pattern_search(b.search_jobs, b.org, b.ref)
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("shape", [(8, 8), (16, 16), (32, 32), (64, 64), (64, 16), (16, 12), (4, 8)])
@pytest.mark.parametrize("n", [1, 3, 300])
def test_k1_matches_oracle_full_range(ctx, shape, n):
    """SearchRange 128 (cfg/3DHencoder_intra_main.cfg:31), the window bench.py's K1 leg is quoted on: 257 x 125
    positions per PU.  n = 1 and 3 cut one PU's window into 32 row slices (the in-encoder geometry), n = 300 gives one
    slice per PU with the 160 KB shared-memory cap deciding between byte path and generic path."""
    c, r = shape
    if n == 300 and c * r >= 2048:
        n = 150                                  # oracle time on the box's host cores; still one slice per PU
    orc = _oracle.oracle()
    b = PuBatch(c, r, n, seed=c * 7 + r + n, sr=128)
    j = b.search_jobs[0]
    assert (j["rng_right"] - j["rng_left"] + 1, j["rng_bottom"] - j["rng_top"] + 1) == (257, 125)
    got = ctx.pattern_search(b.search_jobs, b.org, b.ref)
    want = orc.pattern_search(b.search_jobs, b.org, b.ref)
    assert got.tobytes() == want.tobytes()
    assert want["found"].all()


@pytest.mark.parametrize("shape", ALL_SHAPES)
@pytest.mark.parametrize("use_had", [1, 0])
def test_k2_matches_oracle(ctx, shape, use_had):
    c, r = shape
    orc = _oracle.oracle()
    n = 2 if c * r >= 2048 else 5
    b = PuBatch(c, r, n, seed=c * 17 + r + use_had, sr=24, use_had=use_had, n_start=3)
    got = ctx.pattern_search_gt(b.gt_jobs, b.org, b.ref)
    want = orc.pattern_search_gt(b.gt_jobs, b.org, b.ref)
    gt_assert_equal(got, want)


@pytest.mark.parametrize("shape", [(8, 8), (16, 16), (16, 12), (32, 16), (8, 4)])
def test_main10_matches_oracle(ctx, shape):
    """10-bit input keeps the reference's quirks: 255 clip inside the warp, >>2 distortion scaling."""
    c, r = shape
    orc = _oracle.oracle()
    b = PuBatch(c, r, 4, seed=c + r, bit_depth=10, sr=20, n_start=3)
    assert ctx.pattern_search(b.search_jobs, b.org, b.ref).tobytes() == \
        orc.pattern_search(b.search_jobs, b.org, b.ref).tobytes()
    gt_assert_equal(ctx.pattern_search_gt(b.gt_jobs, b.org, b.ref), orc.pattern_search_gt(b.gt_jobs, b.org, b.ref))


def test_golden_vectors_of_compiled_reference(ctx):
    g = np.load(os.path.join(GOLD, "search_golden.npz"))
    for k in range(int(g["n_search"])):
        t = "s%02d" % k
        org, ref = g[t + "_org"], g[t + "_ref"]
        assert ctx.pattern_search(g[t + "_k1_jobs"], org, ref).tobytes() == g[t + "_k1_out"].tobytes(), k
        gt_assert_equal(ctx.pattern_search_gt(g[t + "_k2_jobs"], org, ref), g[t + "_k2_out"], extras=False)
    g = np.load(os.path.join(GOLD, "dist_golden.npz"))
    for k in range(int(g["n_dist"])):
        t = "d%03d" % k
        assert (ctx.dist(g[t + "_jobs"], g[t + "_org"], g[t + "_cur"]) == g[t + "_out"]).all(), k


def test_k2_edge_cases(ctx):
    orc = _oracle.oracle()
    b = PuBatch(16, 16, 6, seed=21, sr=24, n_start=3)
    # zero start vectors: nothing searched
    j = b.gt_jobs.copy(); j["ss_cand"]["hor"] = 0; j["ss_cand"]["ver"] = 0; j["amvp"]["hor"] = 0; j["amvp"]["ver"] = 0
    r = ctx.pattern_search_gt(j, b.org, b.ref)
    assert (r["gt_flag"] == 0).all() and (r["n_candidates"] == 0).all() and (r["cost"] == j["threshold"]).all()
    # unbeatable threshold, thresholds near the optimum (ties resolved by strict '<' in loop order)
    full = orc.pattern_search_gt(b.gt_jobs, b.org, b.ref)
    for delta in (0, 1, 2, 40):
        j = b.gt_jobs.copy(); j["threshold"] = full["cost"] + delta
        gt_assert_equal(ctx.pattern_search_gt(j, b.org, b.ref), orc.pattern_search_gt(j, b.org, b.ref))
    j = b.gt_jobs.copy(); j["threshold"] = 0
    r = ctx.pattern_search_gt(j, b.org, b.ref)
    assert (r["gt_flag"] == 0).all() and (r["n_candidates"] == 3 * 4 * 56).all()
    # start vectors whose window overlaps NOT_VALID samples (clamped to 0 by the staging)
    j = b.gt_jobs.copy(); j["ss_cand"]["hor"] = -4; j["ss_cand"]["ver"] = -12
    gt_assert_equal(ctx.pattern_search_gt(j, b.org, b.ref), orc.pattern_search_gt(j, b.org, b.ref))
    # flat content: many exact ties
    flat_org = np.full_like(b.org, 100); flat_ref = np.where(b.ref < 0, b.ref, 100).astype(np.int16)
    gt_assert_equal(ctx.pattern_search_gt(b.gt_jobs, flat_org, flat_ref),
                    orc.pattern_search_gt(b.gt_jobs, flat_org, flat_ref))
    assert ctx.pattern_search(b.search_jobs, flat_org, flat_ref).tobytes() == \
        orc.pattern_search(b.search_jobs, flat_org, flat_ref).tobytes()


def test_k1_edge_cases(ctx):
    orc = _oracle.oracle()
    b = PuBatch(16, 8, 5, seed=2, sr=20)
    for mod in ("empty", "all_invalid", "no_ss", "no_fen", "ragged"):
        s, ref = b.search_jobs.copy(), b.ref
        if mod == "empty":
            s["rng_top"] = s["rng_bottom"] + 1
        elif mod == "all_invalid":
            ref = np.full_like(b.ref, -1)
        elif mod == "no_ss":
            s["is_ss"] = 0
        elif mod == "no_fen":
            s["fast_enc"] = 0
        elif mod == "ragged":
            s["rng_left"] += np.arange(5); s["rng_right"] -= 2 * np.arange(5); s["rng_top"] += 3
        assert ctx.pattern_search(s, b.org, ref).tobytes() == orc.pattern_search(s, b.org, ref).tobytes(), mod
    # sentinel inside a block whose probes are valid (cannot happen in the encoder; SAD must still be exact)
    ref = b.ref.copy().reshape(5, b.ph, b.pw)
    ref[:, 3:6, 10:14] = -1
    ref = ref.reshape(-1)
    assert ctx.pattern_search(b.search_jobs, b.org, ref).tobytes() == \
        orc.pattern_search(b.search_jobs, b.org, ref).tobytes()


@pytest.mark.parametrize("shape", [(8, 8), (64, 64), (16, 12), (4, 8), (2, 2), (6, 2), (20, 6), (48, 64)])
def test_k3_matches_oracle(ctx, shape):
    c, r = shape
    orc = _oracle.oracle()
    for bit_depth in (8, 10):
        for func, sub in ((hop.HOP_DF_HADS, 0), (hop.HOP_DF_SAD, 0), (hop.HOP_DF_SAD, 1)):
            if sub and r < 2:
                continue
            jobs, org, cur = dist_jobs(c, r, 7, seed=c * 3 + r, bit_depth=bit_depth, func=func, sub_shift=sub)
            assert (ctx.dist(jobs, org, cur) == orc.dist(jobs, org, cur)).all()


def test_k4_reference_mirror(ctx):
    """Patch-by-patch mirror update + incremental border extension == full extendPicBorder each time."""
    rng = np.random.default_rng(4)
    pic_w, pic_h, m = 192, 128, 80
    ctx.ref_create(pic_w, pic_h, m)
    ctx.ref_reset(-1)
    host = np.full((pic_h + 2 * m, pic_w + 2 * m), -1, dtype=np.int16)
    assert (ctx.ref_download() == host).all()
    for cy in range(0, pic_h, 64):
        for cx in range(0, pic_w, 64):
            for (bx, by, bs) in [(0, 0, 32), (32, 0, 32), (0, 32, 32), (32, 32, 16), (48, 32, 16), (32, 48, 16), (48, 48, 16),
                                 (0, 0, 64)]:
                blk = rng.integers(0, 256, size=(bs, bs)).astype(np.int16)
                ctx.ref_update(cx + bx, cy + by, blk)
                host[m + cy + by:m + cy + by + bs, m + cx + bx:m + cx + bx + bs] = blk
                _oracle.extend_border_oracle(host, pic_w, pic_h, m)
            assert (ctx.ref_download() == host).all(), (cx, cy)
    # golden vector of the compiled reference's extendPicBorder
    g = np.load(os.path.join(GOLD, "border_golden.npz"))
    pw, ph, mm = int(g["pic_w"]), int(g["pic_h"]), int(g["margin"])
    ctx.ref_create(pw, ph, mm)
    ctx.ref_reset(-1)
    ctx.ref_update(0, 0, g["before"][mm:mm + ph, mm:mm + pw])
    assert (ctx.ref_download() == g["after"]).all()


def test_search_on_the_mirror(ctx):
    """ref == NULL: jobs address the context's SS reference mirror (negative offsets reach the margin)."""
    orc = _oracle.oracle()
    rng = np.random.default_rng(8)
    pic_w, pic_h, m = 128, 128, 80
    ctx.ref_create(pic_w, pic_h, m)
    ctx.ref_reset(-1)
    host = np.full((pic_h + 2 * m, pic_w + 2 * m), -1, dtype=np.int16)
    img = rng.integers(0, 256, size=(pic_h, pic_w)).astype(np.int16)
    for (x, y, w, h) in [(0, 0, 128, 64), (0, 64, 64, 64)]:       # first CTU row + first CTU of the second
        ctx.ref_update(x, y, img[y:y + h, x:x + w])
        host[m + y:m + y + h, m + x:m + x + w] = img[y:y + h, x:x + w]
        _oracle.extend_border_oracle(host, pic_w, pic_h, m)
    stride = pic_w + 2 * m
    b = PuBatch(16, 16, 1, seed=1, sr=16)
    s = b.search_jobs.copy()
    s["ref_stride"] = stride
    s["ref_off"] = 64 * stride + 64            # PU at (64,64) relative to sample (0,0)
    s["rng_left"], s["rng_right"], s["rng_top"], s["rng_bottom"] = -72, 40, -72, -4
    got = ctx.pattern_search(s, b.org, None)
    s2 = s.copy(); s2["ref_off"] += m * stride + m
    want = orc.pattern_search(s2, b.org, host.reshape(-1))
    assert got.tobytes() == want.tobytes()
    gj = b.gt_jobs.copy()
    gj["ref_stride"] = stride; gj["ref_off"] = 64 * stride + 64
    gj["ss_cand"]["hor"] = got["mv"]["hor"]; gj["ss_cand"]["ver"] = got["mv"]["ver"]
    gj["amvp"]["hor"][:, 0] = -70 * 4; gj["amvp"]["ver"][:, 0] = -66 * 4      # window reaches into the margin
    g2 = gj.copy(); g2["ref_off"] += m * stride + m
    gt_assert_equal(ctx.pattern_search_gt(gj, b.org, None), orc.pattern_search_gt(g2, b.org, host.reshape(-1)))


def test_full_size_properties(ctx):
    """BASELINE sizes (SearchRange 128, up to 64x64): properties that need no CPU oracle run.
    (a) threshold consistency: re-running K2 with threshold = its own best cost either accepts nothing
        (cost == threshold) or finds a strictly better cost (the diamond walk is greedy, so a different
        threshold may steer it elsewhere) -- it can never return a worse one;
    (b) an accepted result always beats the threshold strictly;
    (c) K1 optimum is a fixed point: restricting the window to the winner returns the same SAD/cost;
    (d) every call scores exactly starts*passes*56 candidates."""
    for (c, r, n) in [(64, 64, 8), (32, 32, 16), (16, 16, 32), (8, 8, 64)]:
        b = PuBatch(c, r, n, seed=c, n_start=2)
        full = ctx.pattern_search_gt(b.gt_jobs, b.org, b.ref)
        assert (full["n_candidates"] == 2 * gt_passes(c, r) * 56).all()
        assert (full["gt_flag"] == 1).all()
        j = b.gt_jobs.copy(); j["threshold"] = full["cost"]
        again = ctx.pattern_search_gt(j, b.org, b.ref)
        assert (again["cost"] <= full["cost"]).all()
        assert ((again["gt_flag"] == 1) == (again["cost"] < full["cost"])).all()
        k1 = ctx.pattern_search(b.search_jobs, b.org, b.ref)
        assert (k1["found"] == 1).all()
        s = b.search_jobs.copy()
        s["rng_left"] = s["rng_right"] = k1["mv"]["hor"]; s["rng_top"] = s["rng_bottom"] = k1["mv"]["ver"]
        one = ctx.pattern_search(s, b.org, b.ref)
        assert one.tobytes() == k1.tobytes()
        # oracle spot check on one PU of the full-size batch
        orc = _oracle.oracle()
        assert orc.pattern_search_gt(b.gt_jobs[:1], b.org, b.ref).tobytes() == full[:1].tobytes()


def test_device_entry_points_with_torch_buffers(ctx):
    """HBM-resident path used by bench.py: torch owns the device buffers and the stream."""
    import torch
    orc = _oracle.oracle()
    b = PuBatch(16, 16, 9, seed=77, sr=24, n_start=2)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(a.view(np.uint8).copy()).to(dev)
    d_org, d_ref = t(b.org), t(b.ref)
    d_sj, d_gj = t(b.search_jobs), t(b.gt_jobs)
    d_so = torch.zeros(len(b.search_jobs) * hop.SEARCH_RES_DT.itemsize, dtype=torch.uint8, device=dev)
    d_go = torch.zeros(len(b.gt_jobs) * hop.GT_RES_DT.itemsize, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()                     # uploads done before the context's stream reads them
    stream = ctx.stream
    ctx.pattern_search_dev(b.n, d_sj.data_ptr(), d_org.data_ptr(), d_ref.data_ptr(), d_so.data_ptr(), stream)
    ctx.pattern_search_gt_dev(b.n, d_gj.data_ptr(), d_org.data_ptr(), d_ref.data_ptr(), b.ref.size, d_go.data_ptr(), 16, 16, stream)
    torch.cuda.synchronize()
    so = d_so.cpu().numpy().view(hop.SEARCH_RES_DT)
    go = d_go.cpu().numpy().view(hop.GT_RES_DT)
    assert so.tobytes() == orc.pattern_search(b.search_jobs, b.org, b.ref).tobytes()
    # the same batch with the shape / window hint (per-width throughput kernel), and with a hint that is too small
    # for the windows (the kernel then takes its generic path: slower, same results)
    for nx_max, ny_max in ((2 * 24 + 1, 24), (8, 8)):
        d_so.zero_()
        ctx.pattern_search_dev(b.n, d_sj.data_ptr(), d_org.data_ptr(), d_ref.data_ptr(), d_so.data_ptr(), stream,
                               cols=16, rows=16, nx_max=nx_max, ny_max=ny_max)
        torch.cuda.synchronize()
        assert d_so.cpu().numpy().view(hop.SEARCH_RES_DT).tobytes() == so.tobytes()
    assert go.tobytes() == orc.pattern_search_gt(b.gt_jobs, b.org, b.ref).tobytes()


def test_async_batches_match_sync(ctx):
    """hop_pattern_search_gt_batch_async: several batches in flight (more than the slot ring holds), results
    land in the callers' buffers after hop_ctx_sync and equal the synchronous call."""
    orc = _oracle.oracle()
    batches = [PuBatch(c, r, 6, seed=40 + i, sr=24, n_start=2) for i, (c, r) in enumerate(
        [(8, 8), (16, 16), (8, 4), (32, 32), (16, 8), (16, 12)])]
    outs = [np.zeros(b.n, dtype=hop.GT_RES_DT) for b in batches]
    for b, o in zip(batches, outs):
        ctx._check(ctx.lib.hop_pattern_search_gt_batch_async(ctx.h, b.n, b.gt_jobs.ctypes.data, b.org.ctypes.data, b.org.size,
                                                             b.ref.ctypes.data, b.ref.size, o.ctypes.data))
    ctx.sync()
    for b, o in zip(batches, outs):
        assert o.tobytes() == orc.pattern_search_gt(b.gt_jobs, b.org, b.ref).tobytes()


def test_wild_amvp_vector_cannot_leave_the_mirror(ctx):
    """AMVP start vectors are raw neighbour vectors; one that points far outside the SS plane (the
    reference would read beyond its buffer) must not fault on the device: reads are kept inside the mirror."""
    rng = np.random.default_rng(9)
    pic_w, pic_h, m = 128, 128, 80
    ctx.ref_create(pic_w, pic_h, m)
    host = rng.integers(0, 256, size=(pic_h + 2 * m, pic_w + 2 * m)).astype(np.int16)
    ctx.ref_upload(host)
    stride = pic_w + 2 * m
    b = PuBatch(16, 16, 1, seed=2, sr=16, n_start=3)
    gj = b.gt_jobs.copy()
    gj["ref_stride"] = stride; gj["ref_off"] = 64 * stride + 64
    gj["ss_cand"]["hor"] = -20; gj["ss_cand"]["ver"] = -30
    gj["amvp"]["hor"][:, 0] = -3000 * 4; gj["amvp"]["ver"][:, 0] = -3000 * 4       # far above / left of the plane
    gj["amvp"]["hor"][:, 1] = 3000 * 4; gj["amvp"]["ver"][:, 1] = 3000 * 4         # far below / right of it
    a = ctx.pattern_search_gt(gj, b.org, None)
    assert ctx.pattern_search_gt(gj, b.org, None).tobytes() == a.tobytes()
    assert a["n_candidates"][0] == 3 * 4 * 56
    mj = b.motion_jobs()
    s = mj["search"]; s["ref_stride"] = stride; s["ref_off"] = 64 * stride + 64
    s["rng_left"], s["rng_right"], s["rng_top"], s["rng_bottom"] = -40, 30, -40, -4
    s["is_ss"] = 0
    mj["search"] = s
    mj["amvp"] = gj["amvp"]
    r = ctx.motion_search(mj, b.org, None)
    assert r["refined"][0] == 1 and ctx.motion_search(mj, b.org, None).tobytes() == r.tobytes()


def test_randomised_content_and_parameters(ctx):
    """Many small random cases: extreme sample values (0 / max checkerboards, flat, noise), random lambda,
    predictors, thresholds, start vectors and staircase positions, both bit depths."""
    orc = _oracle.oracle()
    rng = np.random.default_rng(2026)
    shapes = [(8, 8), (8, 4), (4, 8), (16, 8), (16, 16), (16, 12), (12, 16), (32, 8)]
    for it in range(40):
        c, r = shapes[it % len(shapes)]
        bd = 8 if it % 3 else 10
        maxv = (1 << bd) - 1
        b = PuBatch(c, r, 3, seed=1000 + it, bit_depth=bd, sr=24, use_had=it % 2, n_start=1 + it % 3)
        kind = it % 4
        if kind == 0:      # 0 / max checkerboard
            ref = np.where(b.ref < 0, b.ref, ((np.arange(b.ref.size) + np.arange(b.ref.size) // b.pw) % 2) * maxv).astype(np.int16)
            org = rng.integers(0, 2, size=b.org.size).astype(np.int16) * maxv
        elif kind == 1:    # uniform noise over the full range
            ref = np.where(b.ref < 0, b.ref, rng.integers(0, maxv + 1, size=b.ref.size)).astype(np.int16)
            org = rng.integers(0, maxv + 1, size=b.org.size).astype(np.int16)
        elif kind == 2:    # flat
            ref = np.where(b.ref < 0, b.ref, maxv // 3).astype(np.int16)
            org = np.full(b.org.size, maxv // 3 + 1, dtype=np.int16)
        else:
            ref, org = b.ref, b.org
        sj, gj = b.search_jobs.copy(), b.gt_jobs.copy()
        lam = int(rng.integers(1000, 4000000))
        for j in (sj, gj):
            j["cost"]["lambda_cost"] = lam
            j["cost"]["pred"]["hor"] = rng.integers(-600, 600, size=3).astype(np.int16)
            j["cost"]["pred"]["ver"] = rng.integers(-600, 600, size=3).astype(np.int16)
        gj["threshold"] = rng.choice([0xFFFFFFFE, 500, 5000, 50000], size=3)
        assert ctx.pattern_search(sj, org, ref).tobytes() == orc.pattern_search(sj, org, ref).tobytes(), it
        gt_assert_equal(ctx.pattern_search_gt(gj, org, ref), orc.pattern_search_gt(gj, org, ref))
        fj = b.frac_jobs()
        fj["cost"] = gj["cost"]
        got, want = ctx.frac_search(fj, org, ref), _oracle.frac_search(fj, org, ref)
        assert got.tobytes() == want.tobytes(), it


def test_packed_tile_pairs_equal_single_calls(ctx):
    """A batch evaluates two 8x8 tiles of a candidate in one register tile (16-bit lanes) on a window that carries rows
    of horizontal differences; a single-PU call takes the one-tile forms (row per lane, half tiles).  Two
    independent GPU paths, same answers -- at full PU sizes, both bit depths, extreme content (0 / max checkerboards
    drive the Hadamard coefficients to their bound, 32 x 1023 inside a half tile) and in one mixed-shape batch."""
    from hevc_hop_b200.workload import GtBatch
    rng = np.random.default_rng(8)
    mixed_jobs = []
    for (c, r) in [(64, 64), (32, 32), (16, 16), (64, 16), (16, 64), (32, 24), (64, 48), (48, 64), (16, 8), (8, 16), (32, 8), (8, 8)]:
        for bd in (8, 10):
            n = 4 if c * r >= 2048 else 8
            maxv = (1 << bd) - 1
            b = GtBatch(c, r, n, seed=700 + c + 3 * r + bd, bit_depth=bd)
            org, ref = b.org.copy(), b.ref.copy()
            half = org.size // 2                  # second half of the PUs: checkerboard blocks against a noise window
            org[half:] = ((np.arange(org.size - half) // 3) % 2) * maxv
            ref[ref.size // 2:] = rng.integers(0, maxv + 1, size=ref.size - ref.size // 2)
            gj = b.gt_jobs.copy()
            gj["threshold"] = rng.choice([0xFFFFFFFE, 200000, 20000], size=n)
            got = ctx.pattern_search_gt(gj, org, ref)
            single = np.concatenate([ctx.pattern_search_gt(gj[k:k + 1], org, ref) for k in range(n)])
            assert got.tobytes() == single.tobytes(), (c, r, bd)
            if bd == 8:
                mixed_jobs.append((gj[:2].copy(), org, ref, got[:2].copy()))
    # one batch of every 8-bit shape (two PUs each): per-job shapes differ, the launch is sized for 64x64
    jobs, orgs, refs, want = [], [], [], []
    for (gj, org, ref, res) in mixed_jobs:
        j = gj.copy()
        j["org_off"] += sum(o.size for o in orgs)
        j["ref_off"] += sum(x.size for x in refs)
        jobs.append(j); orgs.append(org); refs.append(ref); want.append(res)
    got = ctx.pattern_search_gt(np.concatenate(jobs), np.concatenate(orgs), np.concatenate(refs))
    assert got.tobytes() == np.concatenate(want).tobytes()


def test_three_lerp_blend_equals_reference_operation_sequence_at_full_size(ctx):
    """DESIGN.md "warp rounding": K2 blends with three lerps; libhopgpu_refops.so is the same library with the
    reference's literal binary64 operation sequence (-DHOP_WARP_REFERENCE_OPS).  The oracle pins both at sizes
    it can finish in seconds; here the two GPU builds are compared at bench size (thousands of PUs up to 64x64,
    both bit depths, diamond search and exhaustive sweep) -- every result field must be identical."""
    import __graft_entry__ as graft
    from hevc_hop_b200.workload import GtBatch
    if not os.path.exists(graft.LIB_REFOPS):
        pytest.skip("libhopgpu_refops.so not built")
    twin = hop.HopContext(0, lib_path=graft.LIB_REFOPS)
    rng = np.random.default_rng(77)
    pixels = 0
    try:
        for (c, r, n) in [(8, 8, 2048), (16, 16, 1024), (32, 32, 512), (64, 64, 384), (16, 12, 512), (8, 4, 1024), (64, 32, 256)]:
            for bd in (8, 10):
                b = GtBatch(c, r, n, seed=900 + c + r + bd, bit_depth=bd)
                gj = b.gt_jobs.copy()
                gj["threshold"] = rng.choice([0xFFFFFFFE, 100000, 5000], size=n)
                gj["use_had"] = rng.integers(0, 2, size=n)
                got, want = ctx.pattern_search_gt(gj, b.org, b.ref), twin.pattern_search_gt(gj, b.org, b.ref)
                assert got.tobytes() == want.tobytes(), (c, r, bd)
                pixels += int(got["n_candidates"].sum()) * c * r
        for (c, r, n) in [(16, 16, 24), (8, 8, 48)]:
            b = PuBatch(c, r, n, seed=31, bit_depth=10, sr=32, n_start=1)
            got, want = ctx.gt_sweep(b.gt_jobs, b.org, b.ref), twin.gt_sweep(b.gt_jobs, b.org, b.ref)
            assert got.tobytes() == want.tobytes(), ("sweep", c, r)
            pixels += int(got["n_candidates"].sum()) * c * r
    finally:
        twin.close()
    assert pixels > 1.5e9   # warped pixels compared (1.85e9 for the diamond cases alone)
