"""CPU suite (-m "not gpu"): pins the oracle (oracle/hop_oracle.c) against the golden vectors produced
by the compiled reference, and -- when oracle/_ref/libhopref.so is present -- against the reference
itself on fresh seeded inputs.  Also covers the host-side logic and the C-ABI surface."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _oracle
import hevc_hop_b200 as hop
from hevc_hop_b200.workload import PuBatch, dist_jobs, gt_passes, CANDIDATES_PER_PASS, lambda_motion_sad

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

needs_ref = pytest.mark.skipif(_oracle.ref() is None, reason="oracle/_ref/libhopref.so not built")


def gt_equal(a, b, extras=True):
    """Compare K2 results; the reference cannot report best_index / n_candidates (extras)."""
    keys = ["gt_flag", "cost"]
    ok = all((a[k] == b[k]).all() for k in keys)
    ok &= a["gt"].tobytes() == b["gt"].tobytes() and a["mv_int"].tobytes() == b["mv_int"].tobytes()
    if extras:
        ok &= (a["best_index"] == b["best_index"]).all() and (a["n_candidates"] == b["n_candidates"]).all()
    return bool(ok)


def test_oracle_search_matches_golden():
    orc = _oracle.oracle()
    g = np.load(os.path.join(GOLD, "search_golden.npz"))
    n = int(g["n_search"])
    assert n >= 30
    for k in range(n):
        t = "s%02d" % k
        org, ref = g[t + "_org"], g[t + "_ref"]
        k1 = orc.pattern_search(g[t + "_k1_jobs"], org, ref)
        assert k1.tobytes() == g[t + "_k1_out"].tobytes(), "K1 case %d" % k
        k2 = orc.pattern_search_gt(g[t + "_k2_jobs"], org, ref)
        assert gt_equal(k2, g[t + "_k2_out"], extras=False), "K2 case %d" % k


def test_oracle_dist_matches_golden():
    orc = _oracle.oracle()
    g = np.load(os.path.join(GOLD, "dist_golden.npz"))
    for k in range(int(g["n_dist"])):
        t = "d%03d" % k
        got = orc.dist(g[t + "_jobs"], g[t + "_org"], g[t + "_cur"])
        assert (got == g[t + "_out"]).all(), "dist case %d" % k


def test_oracle_border_matches_golden():
    g = np.load(os.path.join(GOLD, "border_golden.npz"))
    plane = g["before"].copy()
    _oracle.extend_border_oracle(plane, int(g["pic_w"]), int(g["pic_h"]), int(g["margin"]))
    assert (plane == g["after"]).all()


def test_had_sad_against_numpy():
    """Independent numpy statement of SAD / tiled Hadamard (Sylvester matrix) vs the C oracle."""
    orc = _oracle.oracle()

    def hadamard(n):
        h = np.array([[1]])
        while h.shape[0] < n:
            h = np.block([[h, h], [h, -h]])
        return h

    for (c, r) in [(8, 8), (16, 8), (12, 16), (4, 8), (2, 2), (6, 2), (64, 64)]:
        jobs, org, cur = dist_jobs(c, r, 2, seed=c * 100 + r, func=hop.HOP_DF_HADS)
        got = orc.dist(jobs, org, cur)
        n = 8 if (r % 8 == 0 and c % 8 == 0) else 4 if (r % 4 == 0 and c % 4 == 0) else 2
        H = hadamard(n)
        for i in range(2):
            d = (org.reshape(2, r, c)[i].astype(np.int64) - cur.reshape(2, r, c)[i])
            s = 0
            for y in range(0, r, n):
                for x in range(0, c, n):
                    t = np.abs(H @ d[y:y + n, x:x + n] @ H).sum()
                    s += (t + 2) >> 2 if n == 8 else (t + 1) >> 1 if n == 4 else t
            assert got[i] == s
        jobs, org, cur = dist_jobs(c, r, 2, seed=c * 100 + r + 1, func=hop.HOP_DF_SAD)
        got = orc.dist(jobs, org, cur)
        want = np.abs(org.astype(np.int64) - cur).reshape(2, -1).sum(axis=1)
        assert (got == want).all()


def test_component_bits_closed_form():
    orc = _oracle.oracle().lib
    for v in list(range(-600, 600)) + [-32768, 32767, 4096, -4096]:
        t = (-v * 2 + 1) if v <= 0 else v * 2
        assert orc.orc_component_bits(v) == 2 * (t.bit_length() - 1) + 1


def test_affine_candidate_table_has_56_entries():
    """Host restatement of the 8-deep loop nest: 620 non-translation diamond combos, 56 parallelograms."""
    v = (1, 0, -1)
    diamond = [(y, x) for y in v for x in v if y == 0 or x == 0]
    total = affine = 0
    for (y0, x0) in diamond:
        for (y1, x1) in diamond:
            for (y2, x2) in diamond:
                for (y3, x3) in diamond:
                    if x0 == x1 == x2 == x3 and y0 == y1 == y2 == y3:
                        continue
                    total += 1
                    if x0 - x1 + x2 - x3 == 0 and y0 - y1 + y2 - y3 == 0:
                        affine += 1
    assert total == 620 and affine == CANDIDATES_PER_PASS == 56


def test_gt_passes_per_shape():
    assert [gt_passes(*s) for s in [(8, 8), (16, 16), (32, 32), (64, 64), (8, 4), (16, 12), (64, 16)]] == \
        [3, 4, 5, 6, 2, 3, 4]


def test_oracle_counts_all_candidates():
    orc = _oracle.oracle()
    b = PuBatch(16, 16, 3, seed=5, sr=24, n_start=3)
    r = orc.pattern_search_gt(b.gt_jobs, b.org, b.ref)
    assert (r["n_candidates"] == 3 * 4 * 56).all()


def test_oracle_edge_cases():
    orc = _oracle.oracle()
    b = PuBatch(8, 8, 4, seed=9, sr=16)
    # (a) zero start vector and zero predictors: nothing is searched, threshold returned untouched
    j = b.gt_jobs.copy()
    j["ss_cand"]["hor"] = 0; j["ss_cand"]["ver"] = 0
    r = orc.pattern_search_gt(j, b.org, b.ref)
    assert (r["gt_flag"] == 0).all() and (r["cost"] == j["threshold"]).all() and (r["n_candidates"] == 0).all()
    # (b) unbeatable threshold: every candidate scored, none accepted
    j = b.gt_jobs.copy(); j["threshold"] = 0
    r = orc.pattern_search_gt(j, b.org, b.ref)
    assert (r["gt_flag"] == 0).all() and (r["n_candidates"] == 3 * 56).all()
    # (c) K1 with an empty / degenerate window (first CTU: top = bottom + 1) => not found, MAX_UINT
    s = b.search_jobs.copy(); s["rng_top"] = s["rng_bottom"] + 1
    r = orc.pattern_search(s, b.org, b.ref)
    assert (r["found"] == 0).all() and (r["sad"] == hop.HOP_MAX_UINT).all()
    # (d) K1 where every probe is NOT_VALID
    ref = np.full_like(b.ref, -1)
    r = orc.pattern_search(b.search_jobs, b.org, ref)
    assert (r["found"] == 0).all()
    # (e) non-SS search ignores the gates and the sentinel
    s = b.search_jobs.copy(); s["is_ss"] = 0
    r = orc.pattern_search(s, b.org, ref)
    assert (r["found"] == 1).all()


@needs_ref
@pytest.mark.parametrize("shape", [(8, 8), (16, 16), (32, 32), (8, 4), (4, 8), (16, 12), (32, 8), (24, 32), (64, 16)])
@pytest.mark.parametrize("bit_depth", [8, 10])
def test_oracle_vs_compiled_reference(shape, bit_depth):
    orc, ref = _oracle.oracle(), _oracle.ref()
    c, r = shape
    for use_had in (1, 0):
        b = PuBatch(c, r, 3, seed=c * 7 + r + bit_depth, bit_depth=bit_depth, sr=20, use_had=use_had, n_start=3,
                    threshold=0xFFFFFFFE if use_had else 2500)
        assert orc.pattern_search(b.search_jobs, b.org, b.ref).tobytes() == \
            ref.pattern_search(b.search_jobs, b.org, b.ref).tobytes()
        assert gt_equal(orc.pattern_search_gt(b.gt_jobs, b.org, b.ref),
                        ref.pattern_search_gt(b.gt_jobs, b.org, b.ref), extras=False)


@needs_ref
def test_oracle_warp_vs_compiled_reference():
    """calcParamProjective + ProjectiveTransform on random parallelogram corner sets, incl. windows with clamps."""
    orc, ref = _oracle.oracle(), _oracle.ref()
    rng = np.random.default_rng(11)
    for (c, r) in [(8, 8), (16, 8), (32, 32), (12, 16)]:
        w = min(c, r)
        for _ in range(20):
            d0 = rng.integers(-w + 1, w, size=2); d1 = rng.integers(-w + 1, w, size=2); d2 = rng.integers(-w + 1, w, size=2)
            d3 = d0 - d1 + d2
            cx = np.array([0 + d0[0], 2 * c - 1 + d1[0], 2 * c - 1 + d2[0], 0 + d3[0]], dtype=np.int32)
            cy = np.array([0 + d0[1], 0 + d1[1], 2 * r - 1 + d2[1], 2 * r - 1 + d3[1]], dtype=np.int32)
            h1, h2 = orc.calc_param_projective(cx, cy, 2 * c, 2 * r), ref.calc_param_projective(cx, cy, 2 * c, 2 * r)
            assert h1.tobytes() == h2.tobytes()
            if not (h1[2] == 0.0 and h1[5] == 0.0):
                continue
            win = rng.integers(0, 256, size=(2 * r, 2 * c)).astype(np.int16)
            a1 = orc.projective_transform(win, c, r, h1, w)
            a2 = ref.projective_transform(win, c, r, h1, w)
            assert (a1 == a2).all()


@needs_ref
def test_bits_and_cost_vs_compiled_reference():
    orc, ref = _oracle.oracle().lib, _oracle.ref().lib
    rng = np.random.default_rng(3)
    for _ in range(300):
        v = [int(x) for x in rng.integers(-200, 200, size=6)]
        assert orc.orc_get_bits_gt(*v) == ref.ref_bits_gt(*v)
    cs = hop.HopCostState(lambda_motion_sad(32), 2, hop.HopMv(-37, 12))
    for _ in range(300):
        x, y = (int(t) for t in rng.integers(-140, 140, size=2))
        assert orc.orc_get_cost_xy(C.byref(cs), x, y) == ref.ref_get_cost_xy(C.byref(cs), x, y)


# ---- C-ABI surface (no compute without a GPU) ------------------------------------------------------
def test_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "hop_gpu.h")).read()
    declared = set(re.findall(r"\b(hop_[a-z0-9_]+)\s*\(", header))
    bound = {name for name, _, _ in hop.ABI}
    assert declared == bound, (declared ^ bound)
    lib = hop.load_library()       # raises if libhopgpu.so is missing or a symbol is not exported
    assert lib.hop_abi_version() == 2
    assert lib.hop_shape_supported(16, 12) == 1 and lib.hop_shape_supported(4, 4) == 0 and lib.hop_shape_supported(20, 8) == 0


def test_no_cpu_fallback_without_device():
    """Without a CUDA device context creation must fail loudly (status + message), never compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(hop.HopError) as e:
        hop.HopContext(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_reference_oracle():
    """The product tree (hevc-hop_b200/, include/) must not import, link or execute oracle/ code."""
    bad = []
    for base in ("hevc-hop_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            if os.path.basename(dp) == "build":
                continue
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dp, fn)).read()
                    if re.search(r"hop_oracle|libhoporacle|libhopref|_oracle\b|orc_[a-z]", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_host_incremental_border_equals_full_extension(tmp_path):
    """include/hop_border.h (what the shim runs on the HOST plane after every CU of xCopyYuv2SSRef) against the
    reference's full re-extension (oracle restatement of TComPicYuv.cpp:247-274) after every commit of a CTU walk."""
    import subprocess
    so = str(tmp_path / "libborder.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "helpers", "border_harness.cpp"), "-o", so])
    lib = C.CDLL(so)
    lib.border_extend_patch.argtypes = [C.c_void_p] + [C.c_int] * 8
    lib.border_extend_patch.restype = None
    rng = np.random.default_rng(3)
    for (pic_w, pic_h, m) in [(64, 64, 80), (136, 104, 80), (200, 72, 40)]:
        stride = pic_w + 2 * m
        inc = np.full((pic_h + 2 * m, stride), -1, dtype=np.int16)
        full = inc.copy()
        img = rng.integers(0, 256, size=(pic_h, pic_w)).astype(np.int16)
        # quadtree-like commits: every 8x8 block in raster-of-CTU order, then the enclosing 16 / 32 / 64 blocks again
        blocks = []
        for cy in range(0, pic_h, 64):
            for cx in range(0, pic_w, 64):
                for size in (8, 16, 32, 64):
                    for by in range(cy, min(cy + 64, pic_h), size):
                        for bx in range(cx, min(cx + 64, pic_w), size):
                            if bx + size <= pic_w and by + size <= pic_h:
                                blocks.append((bx, by, size))
        for (bx, by, size) in blocks:
            blk = img[by:by + size, bx:bx + size] ^ (size & 8)
            for p in (inc, full):
                p[m + by:m + by + size, m + bx:m + bx + size] = blk
            origin = inc.ctypes.data + 2 * (m * stride + m)
            lib.border_extend_patch(origin, stride, pic_w, pic_h, m, bx, by, size, size)
            _oracle.extend_border_oracle(full, pic_w, pic_h, m)
            assert (inc == full).all(), (pic_w, pic_h, bx, by, size)


def test_batch_front_end_over_the_unpatched_reference(tmp_path):
    """integration/hop_batch_main.cpp runs the reference's TAppEncTop once per job inside ONE process (the GPU build keeps
    its CUDA context that way).  Built over the UNPATCHED reference objects (oracle/_ref/TAppEncoderRefBatch,
    -DHOP_BATCH_NO_GPU) it must write, job after job, what a process per image writes -- bitstream, reconstruction and
    syntax trace -- also when an image comes by a second time and when the bit depth changes between jobs."""
    from hevc_hop_b200 import encoder
    exe = os.path.join(os.path.dirname(_oracle.REF_ENCODER), "TAppEncoderRefBatch")
    if not (os.path.exists(exe) and os.path.exists(_oracle.REF_ENCODER)):
        pytest.skip("reference encoders not built")
    cases = [dict(width=64, height=64, seed=1), dict(width=72, height=64, seed=2, qp=27),
             dict(width=64, height=64, seed=1), dict(width=64, height=64, seed=5, bit_depth=10)]
    w = encoder.EncoderWorker(binary=exe)
    try:
        outs = [w.encode(**c) for c in cases]
    finally:
        w.close()
    for c, o in zip(cases, outs):
        ref = _oracle.encode_reference(**c)
        assert o["bitstream"] == ref["bitstream"] and o["rec"] == ref["rec"] and o["trace"] == ref["trace"], c


def test_packed_lane_hadamard_model():
    """The arithmetic claim behind K2's two-tiles-per-register-tile form (k2_gt.cu, eval_half_tile8_pair), modelled in
    numpy uint32: two 8x8 residual tiles ride in the 16-bit lanes of one word with a bias of 0x8000 per lane;
    a + b - 0x80008000 / a - b + 0x80008000 are the lane-wise butterflies with no carry or borrow between the lanes for
    every residual of <= 10-bit content, |x| + 0x8000 = max(lane, 0x10000 - lane), and the rounded SATDs equal the
    plain 8x8 Hadamard SATD of each tile -- including the all-extreme tiles that drive a coefficient to 32 x 1023."""
    rng = np.random.default_rng(3)
    BIAS = np.uint32(0x80008000)

    def satd_plain(d):                       # xCalcHADs8x8 (TComRdCost.cpp:1478-1575): sum |H d H| with (sum + 2) >> 2
        h = np.array([[1]], dtype=np.int64)
        for _ in range(3):
            h = np.block([[h, h], [h, -h]])
        return (int(np.abs(h @ d.astype(np.int64) @ h).sum()) + 2) >> 2

    def satd_packed(da, db):
        w = ((da.astype(np.int64) + 0x8000) | ((db.astype(np.int64) + 0x8000) << 16)).astype(np.uint32)   # (8, 8) words
        halves = []
        for half in (0, 1):                  # the two lanes of a pair: rows 4*half .. 4*half+3
            t = w[4 * half:4 * half + 4].copy()
            for ln in (1, 2, 4):             # horizontal 8-point
                for i in range(0, 8, 2 * ln):
                    for j in range(i, i + ln):
                        a, b = t[:, j].copy(), t[:, j + ln].copy()
                        t[:, j], t[:, j + ln] = a + b - BIAS, a - b + BIAS
            for ln in (1, 2):                # vertical 4-point inside the half
                for i in range(0, 4, 2 * ln):
                    for j in range(i, i + ln):
                        a, b = t[j].copy(), t[j + ln].copy()
                        t[j], t[j + ln] = a + b - BIAS, a - b + BIAS
            assert ((t & 0xffff) != 0).all() and ((t >> 16) != 0).all()        # lanes never reach 0: negation below is exact
            halves.append(t)
        s = np.zeros(2, dtype=np.int64)
        for (mine, other) in ((halves[0][:2], halves[1][:2]), (halves[1][2:], halves[0][2:])):   # last vertical stage, folded
            for m, r in zip(mine.reshape(-1), other.reshape(-1)):
                cand = [np.uint32(m), np.uint32(0x00010000) - np.uint32(m), np.uint32(r), np.uint32(0x00010000) - np.uint32(r)]
                lo = max(int(c) & 0xffff for c in cand) - 0x8000
                hi = max(int(c) >> 16 for c in cand) - 0x8000
                s += (lo, hi)
        return ((int(s[0]) + 1) >> 1), ((int(s[1]) + 1) >> 1)

    with np.errstate(over="ignore"):
        for it in range(60):
            maxv = 1023 if it % 2 else 255
            if it < 8:                       # extreme tiles: every residual +-maxv in Hadamard-aligned sign patterns
                sign = np.where((np.add.outer(np.arange(8) * (it & 3), np.arange(8) * (it >> 1)) & 1) == 0, 1, -1)
                da, db = sign * maxv, -sign * maxv if it & 1 else np.full((8, 8), maxv)
            else:
                da, db = rng.integers(-maxv, maxv + 1, size=(8, 8)), rng.integers(-maxv, maxv + 1, size=(8, 8))
            assert satd_packed(np.asarray(da), np.asarray(db)) == (satd_plain(np.asarray(da)), satd_plain(np.asarray(db))), it


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` runs here (no GPU): one JSON line on stdout with the contract's keys, the same
    `config.workload` string as the GPU arm, and a cpu_baseline block that describes the bounded sample."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if _oracle.ref() is None:
        pytest.skip("oracle/_ref/libhopref.so not built")
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-500:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    sys.path.insert(0, root)
    import bench
    assert d["config"]["workload"] == bench.WORKLOAD % 4096


def test_batch_driver_on_cpu_workers_and_processes_agree():
    """hevc_hop_b200.batch.encode_batch over the CPU builds of the two front ends (unpatched reference): a pool of long-lived
    workers pulling from a queue writes the same bitstreams as one process per image; a task the encoder rejects is
    reported as an error and the rest of the queue is still encoded."""
    from hevc_hop_b200 import batch
    exe = os.path.join(os.path.dirname(_oracle.REF_ENCODER), "TAppEncoderRefBatch")
    if not (os.path.exists(exe) and os.path.exists(_oracle.REF_ENCODER)):
        pytest.skip("reference encoders not built")
    tasks = [dict(width=64, height=64, seed=30 + i) for i in range(4)]
    a, _ = batch.encode_batch(tasks, procs=2, persistent=True, binary=exe)
    b, _ = batch.encode_batch(tasks, procs=2, persistent=False, binary=_oracle.REF_ENCODER)
    assert all("error" not in r for r in a + b), (a, b)
    assert [r["md5"] for r in a] == [r["md5"] for r in b] and [r["rec_md5"] for r in a] == [r["rec_md5"] for r in b]
    assert sum(1 for r in a if r["worker_startup_s"] > 0) == 2
    bad = [dict(width=64, height=64, seed=1), dict(width=64, height=64, seed=2, cfg="/nonexistent/hop.cfg"), dict(width=64, height=64, seed=3)]
    c, _ = batch.encode_batch(bad, procs=1, persistent=True, binary=exe)
    assert "error" in c[1] and "error" not in c[0] and "error" not in c[2]
