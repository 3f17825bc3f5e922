"""Synthetic PU batches for the stand-alone HOP candidate-search microbench (BASELINE.json configs[2],
SURVEY.md §8d) and for parity tests.

Every PU gets its own little "causal neighbourhood": a plane of (2*ox + W + 8) x (oy + H + 8) int16
samples cut from the synthetic lenslet image, with the PU's own position at (ox, oy) =
(max(SR, W), max(SR, 2H + 4)).  Samples at x >= ox and y >= oy (the PU itself and everything right/below it) are NOT_VALID (-1), exactly the
staircase the SS reference has while a CU is being coded.  The search window, gate offsets and cost
state are what TEncSearch::xSetSearchRange / xMotionEstimation produce for a 2Nx2N PU at a CU origin
(TEncSearch.cpp:6224-6259, 4555-4560).
"""
import math

import numpy as np

from . import GT_JOB_DT, SEARCH_JOB_DT, DIST_JOB_DT, FRAC_JOB_DT, MOTION_JOB_DT, PRED_JOB_DT, INTRA_JOB_DT, HOP_DF_HADS, HOP_DF_SAD
from .lenslet import lenslet_luma

SEARCH_RANGE = 128
PROBE_PAD = 8


def lambda_motion_sad(qp=32):
    """m_uiLambdaMotionSAD for an intra(SS) slice at `qp`: TEncSlice.cpp:385-395 + TComRdCost.cpp:167-173."""
    lam = 0.57 * 2.0 ** ((qp - 12) / 3.0)
    return int(math.floor(65536.0 * math.sqrt(lam)))


def plane_origin(cols, rows, sr=SEARCH_RANGE):
    """Position of the PU inside its private plane: far enough from the plane edges for the K1 search
    range AND for the 2W x 2H window of any K2 start vector drawn by PuBatch."""
    return max(sr, cols), max(sr, 2 * rows + 4)


def plane_dims(cols, rows, sr=SEARCH_RANGE):
    ox, oy = plane_origin(cols, rows, sr)
    return 2 * ox + cols + PROBE_PAD, oy + rows + PROBE_PAD


class PuBatch:
    """org/ref sample buffers + K1 and K2 job arrays for `n` PUs of one shape."""

    def __init__(self, cols, rows, n, seed=0, bit_depth=8, qp=32, sr=SEARCH_RANGE, use_had=1,
                 n_start=1, threshold=0xFFFFFFFE, source=None):
        self.cols, self.rows, self.n, self.sr = cols, rows, n, sr
        pw, ph = plane_dims(cols, rows, sr)
        ox, oy = plane_origin(cols, rows, sr)
        self.pw, self.ph, self.ox, self.oy = pw, ph, ox, oy
        rng = np.random.default_rng(seed)
        if source is None:
            # one big lenslet picture, planes are random crops of it (cheap for thousands of PUs)
            side = max(1024, 2 * pw)
            source = lenslet_luma(side, side, seed=seed, bit_depth=bit_depth).astype(np.int16)
        sh, sw = source.shape
        ref = np.empty((n, ph, pw), dtype=np.int16)
        org = np.empty((n, rows, cols), dtype=np.int16)
        ys = rng.integers(0, sh - ph, size=n)
        xs = rng.integers(0, sw - pw, size=n)
        for i in range(n):
            crop = source[ys[i]:ys[i] + ph, xs[i]:xs[i] + pw]
            org[i] = crop[oy:oy + rows, ox:ox + cols]
            ref[i] = crop
        ref[:, oy:, ox:] = -1
        self.org = org.reshape(-1)
        self.ref = ref.reshape(-1)
        lam = lambda_motion_sad(qp)

        sj = np.zeros(n, dtype=SEARCH_JOB_DT)
        idx = np.arange(n, dtype=np.int64)
        sj["org_off"] = idx * (rows * cols)
        sj["ref_off"] = idx * (ph * pw) + oy * pw + ox
        sj["org_stride"] = cols
        sj["ref_stride"] = pw
        sj["cols"], sj["rows"] = cols, rows
        sj["rng_left"], sj["rng_right"] = -sr, sr
        sj["rng_top"], sj["rng_bottom"] = -sr, -4            # min(bottom, -offY-4), offY = 0
        sj["offset_x"], sj["offset_y"] = -cols - 4, -rows - 4
        sj["is_ss"], sj["fast_enc"], sj["bit_depth"] = 1, 1, bit_depth
        sj["cost"]["lambda_cost"] = lam
        sj["cost"]["cost_scale"] = 2
        # AMVP-like predictor: a quarter-pel vector pointing into the causal area
        sj["cost"]["pred"]["hor"] = (rng.integers(-sr // 2, sr // 2, size=n) * 4).astype(np.int16)
        sj["cost"]["pred"]["ver"] = (rng.integers(-sr, 0, size=n) * 4).astype(np.int16)
        self.search_jobs = sj

        gj = np.zeros(n, dtype=GT_JOB_DT)
        for k in ("org_off", "ref_off", "org_stride", "ref_stride", "cols", "rows", "bit_depth"):
            gj[k] = sj[k]
        # start vector: somewhere in the fully coded rows above the PU, 2W x 2H window inside the plane
        hy_lo = -oy + rows // 2
        hy_hi = -rows - rows // 2 - 4
        hx_lo, hx_hi = -ox + cols // 2, ox - cols // 2
        hx = rng.integers(hx_lo, hx_hi + 1, size=n)
        hy = rng.integers(hy_lo, hy_hi + 1, size=n)
        gj["ss_cand"]["hor"] = hx.astype(np.int16)
        gj["ss_cand"]["ver"] = hy.astype(np.int16)
        gj["num_pred"] = 2                                     # fillMvpCand always returns AMVP_MAX_NUM_CANDS
        if n_start > 1:
            for k in range(min(n_start - 1, 2)):
                ax = rng.integers(hx_lo, hx_hi + 1, size=n) * 4 + rng.integers(0, 4, size=n)
                ay = rng.integers(hy_lo, hy_hi + 1, size=n) * 4 + rng.integers(0, 4, size=n)
                gj["amvp"][:, k]["hor"] = ax.astype(np.int16)
                gj["amvp"][:, k]["ver"] = ay.astype(np.int16)
        gj["threshold"] = threshold
        gj["use_had"] = use_had
        gj["cost"]["lambda_cost"] = lam
        gj["cost"]["cost_scale"] = 0
        gj["cost"]["pred"] = sj["cost"]["pred"]
        self.gt_jobs = gj

    def passes(self):
        return gt_passes(self.cols, self.rows)

    def frac_jobs(self, mv_int=None):
        """xPatternSearchFracDIF jobs refining `mv_int` (default: the K2 start vectors, which lie in the
        fully coded rows above the PU so that the 8-tap support is valid)."""
        fj = np.zeros(self.n, dtype=FRAC_JOB_DT)
        for k in ("org_off", "ref_off", "org_stride", "ref_stride", "cols", "rows", "use_had", "bit_depth"):
            fj[k] = self.gt_jobs[k]
        mv = (self.gt_jobs["ss_cand"] if mv_int is None else mv_int).copy()
        if mv_int is None:
            # the 8-tap support reaches 4 samples beyond the block: keep it inside this PU's plane
            mv["hor"] = np.maximum(mv["hor"], -self.ox + 4)
            mv["ver"] = np.maximum(mv["ver"], -self.oy + 4)
        fj["mv_int"] = mv
        fj["cost"] = self.gt_jobs["cost"]
        return fj

    def motion_jobs(self, use_gt=1):
        mj = np.zeros(self.n, dtype=MOTION_JOB_DT)
        sj = self.search_jobs.copy()
        # the frac stage reads 4 samples beyond the matched block: keep the window 4 away from the plane edge
        sj["rng_left"] = np.maximum(sj["rng_left"], -self.ox + 4)
        sj["rng_top"] = np.maximum(sj["rng_top"], -self.oy + 4)
        mj["search"] = sj
        mj["use_had"] = self.gt_jobs["use_had"]
        mj["use_gt"] = use_gt
        mj["num_pred"] = self.gt_jobs["num_pred"]
        mj["amvp"] = self.gt_jobs["amvp"]
        return mj


def gt_passes(cols, rows):
    """Diamond passes per start vector: j0 = window, window/2, ... > 1, at most 6 (TEncSearch.cpp:5181)."""
    j0 = (min(cols, rows) >> 1) * 2
    p = 0
    while j0 > 1 and p < 6:
        p += 1
        j0 //= 2
    return p


CANDIDATES_PER_PASS = 56   # affine corner sets among the 620 diamond combinations


def dist_jobs(cols, rows, n, seed=0, bit_depth=8, func=HOP_DF_HADS, sub_shift=0):
    """n independent (org, cur) block pairs of one shape for the K3 stand-alone distortion kernels."""
    rng = np.random.default_rng(seed)
    maxv = (1 << bit_depth) - 1
    org = rng.integers(0, maxv + 1, size=(n, rows, cols)).astype(np.int16)
    cur = np.clip(org + rng.integers(-40, 41, size=org.shape), -1, maxv).astype(np.int16)
    jobs = np.zeros(n, dtype=DIST_JOB_DT)
    idx = np.arange(n, dtype=np.int64)
    jobs["org_off"] = idx * rows * cols
    jobs["cur_off"] = idx * rows * cols
    jobs["org_stride"] = cols
    jobs["cur_stride"] = cols
    jobs["cols"], jobs["rows"] = cols, rows
    jobs["func"], jobs["sub_shift"], jobs["bit_depth"] = func, sub_shift, bit_depth
    return jobs, org.reshape(-1), cur.reshape(-1)


class GtBatch:
    """K2-only microbench batch (BASELINE.json configs[2]): per PU the original block and the compact
    2W x 2H reference window of ONE start vector -- exactly the samples xPatternSearchGT stages into
    m_filteredBlock[0][0] (TEncSearch.cpp:5161-5165).  ref_off is set so that the start vector
    (ss_cand) lands on the window; the PU's own position lies outside the buffer and is never read."""

    def __init__(self, cols, rows, n, seed=0, bit_depth=8, qp=32, use_had=1, threshold=0xFFFFFFFE, source=None):
        self.cols, self.rows, self.n = cols, rows, n
        rng = np.random.default_rng(seed)
        if source is None:
            source = lenslet_luma(1024, 1024, seed=seed, bit_depth=bit_depth).astype(np.int16)
        sh, sw = source.shape
        ww, wh = 2 * cols, 2 * rows
        win = np.empty((n, wh, ww), dtype=np.int16)
        org = np.empty((n, rows, cols), dtype=np.int16)
        ys = rng.integers(0, sh - wh - 16, size=n)
        xs = rng.integers(0, sw - ww - 16, size=n)
        for i in range(n):
            win[i] = source[ys[i]:ys[i] + wh, xs[i]:xs[i] + ww]
            # the block one micro-image pitch (15 px) to the right of the window centre: self-similar content
            oy, ox = ys[i] + rows // 2, xs[i] + cols // 2 + 15
            org[i] = source[oy:oy + rows, ox:ox + cols]
        self.org = org.reshape(-1)
        self.ref = win.reshape(-1)
        hx = rng.integers(-120, -8, size=n)
        hy = rng.integers(-120, -8, size=n)
        gj = np.zeros(n, dtype=GT_JOB_DT)
        idx = np.arange(n, dtype=np.int64)
        gj["org_off"] = idx * (rows * cols)
        gj["org_stride"] = cols
        gj["ref_stride"] = ww
        gj["ref_off"] = idx * (ww * wh) + (rows // 2 - hy) * ww + (cols // 2 - hx)
        gj["cols"], gj["rows"] = cols, rows
        gj["ss_cand"]["hor"] = hx.astype(np.int16)
        gj["ss_cand"]["ver"] = hy.astype(np.int16)
        gj["num_pred"] = 2          # fillMvpCand pads with zero vectors, which the search skips
        gj["threshold"] = threshold
        gj["use_had"] = use_had
        gj["bit_depth"] = bit_depth
        gj["cost"]["lambda_cost"] = lambda_motion_sad(qp)
        gj["cost"]["cost_scale"] = 0
        gj["cost"]["pred"]["hor"] = ((hx + rng.integers(-3, 4, size=n)) * 4).astype(np.int16)
        gj["cost"]["pred"]["ver"] = ((hy + rng.integers(-3, 4, size=n)) * 4).astype(np.int16)
        self.gt_jobs = gj

    def candidates(self):
        return self.n * gt_passes(self.cols, self.rows) * CANDIDATES_PER_PASS

    def pixel_candidates(self):
        return self.candidates() * self.cols * self.rows

    def input_bytes(self):
        return self.org.nbytes + self.ref.nbytes + self.gt_jobs.nbytes


class PredBatch:
    """K6 jobs (motion-compensated prediction, GT and plain, luma and chroma, distortion / AMVP template cost) over
    one picture-like plane: every PU sits far enough from the plane edges for its 2W x 2H region, the vector and the
    8-tap support.  kind: "plain", "gt", "dist", "template"."""

    def __init__(self, shapes, n_per_shape, seed=0, bit_depth=8, comp=0, kind="gt", frac=True, with_invalid=False):
        rng = np.random.default_rng(seed)
        pw, ph = 640, 400
        plane = lenslet_luma(pw, ph, seed=seed, bit_depth=bit_depth).astype(np.int16)
        if with_invalid:
            plane[ph // 2:, pw // 2:] = -1            # not yet coded: the reference copies / filters these samples as they are
        self.ref = plane.reshape(-1)
        jobs, orgs, dst_off = [], [], 0
        for (c, r) in shapes:
            for _ in range(n_per_shape):
                j = np.zeros(1, dtype=PRED_JOB_DT)
                bw, bh = (c >> 1, r >> 1) if comp else (c, r)
                px = int(rng.integers(200, pw - 200)); py = int(rng.integers(160, ph - 160))
                j["ref_off"] = py * pw + px
                j["ref_stride"] = pw
                j["cols"], j["rows"], j["comp"], j["bit_depth"] = c, r, comp, bit_depth
                mvx = int(rng.integers(-60, 60)) * 4; mvy = int(rng.integers(-60, 60)) * 4
                if frac and kind != "gt":
                    mvx += int(rng.integers(0, 4)); mvy += int(rng.integers(0, 4))
                if kind == "gt" and comp:
                    pass                                   # integer luma vectors: chroma sees 0 or half-pel fractions
                if kind == "gt" and frac and rng.integers(0, 4) == 0:
                    mvx += int(rng.integers(0, 4)); mvy += int(rng.integers(0, 4))   # the reference handles it; the search never produces it
                j["mv"]["hor"], j["mv"]["ver"] = mvx, mvy
                if kind == "gt":
                    j["gt_flag"] = 1
                    w = min(c, r) >> 1
                    if rng.integers(0, 8) == 0:
                        g = np.zeros((4, 2), dtype=np.int64)       # flag set, all-zero vectors: plain branch
                    else:
                        d0 = rng.integers(-w + 1, w, size=2); d1 = rng.integers(-w + 1, w, size=2); d2 = rng.integers(-w + 1, w, size=2)
                        g = np.stack([d0, d1, d2, d0 - d1 + d2])   # parallelogram: what the affine search produces
                        g = np.clip(g, -w + 1, w - 1)
                    j["gt"]["hor"][0], j["gt"]["ver"][0] = g[:, 0], g[:, 1]
                org = np.clip(plane[py:py + bh, px:px + bw] + rng.integers(-12, 13, size=(bh, bw)), 0, (1 << bit_depth) - 1).astype(np.int16)
                j["org_off"] = sum(o.size for o in orgs)
                j["org_stride"] = bw
                orgs.append(org.reshape(-1))
                j["dst_off"] = dst_off
                j["dst_stride"] = bw
                dst_off += bw * bh
                if kind == "dist":
                    j["dist_func"] = HOP_DF_HADS if rng.integers(0, 2) else HOP_DF_SAD
                    if rng.integers(0, 2):
                        j["gt_flag"] = 1
                        j["gt"]["hor"][0] = [1, 2, 2, 1]; j["gt"]["ver"][0] = [-1, -1, 1, 1]
                        j["mv"]["hor"], j["mv"]["ver"] = (mvx >> 2) << 2, (mvy >> 2) << 2
                if kind == "template":
                    j["template_cost"], j["is_ss"] = 1, int(rng.integers(0, 2))
                    j["mv_probe"]["hor"] = mvx + int(rng.integers(-8, 9)); j["mv_probe"]["ver"] = mvy + int(rng.integers(-8, 9))
                    j["mvp_bits"] = int(rng.integers(1, 4))
                    j["lambda_sad"] = lambda_motion_sad(int(rng.integers(22, 38)))
                jobs.append(j)
        self.jobs = np.concatenate(jobs)
        self.org = np.concatenate(orgs)
        self.dst_samples = dst_off


def intra_jobs(sizes, n_per_size, seed=0, bit_depth=8):
    """K7 jobs: per PU an original block and the four reference-sample arrays (unfiltered / filtered, above / left,
    2N+1 entries each, entry 0 = corner) as TComPattern::initAdiPattern would leave them: the unfiltered samples come
    from the picture around the block, the filtered ones are their [1 2 1] smoothing (any values exercise the
    predictors; the pre-screen only reads them)."""
    rng = np.random.default_rng(seed)
    maxv = (1 << bit_depth) - 1
    src = lenslet_luma(512, 512, seed=seed, bit_depth=bit_depth).astype(np.int64)
    jobs, orgs, refs = [], [], []
    for n in sizes:
        for _ in range(n_per_size):
            j = np.zeros(1, dtype=INTRA_JOB_DT)
            y, x = int(rng.integers(1, 512 - 2 * n - 1)), int(rng.integers(1, 512 - 2 * n - 1))
            org = src[y:y + n, x:x + n]
            above = src[y - 1, x - 1:x + 2 * n].copy()            # corner + 2N above
            left = src[y - 1:y + 2 * n, x - 1].copy()             # corner + 2N left
            if rng.integers(0, 4) == 0:
                above[:] = left[:] = 1 << (bit_depth - 1)         # no neighbours: the default fill
            line = np.concatenate([left[:0:-1], above])           # bottom-left ... corner ... top-right
            fl = line.copy()
            fl[1:-1] = (line[:-2] + 2 * line[1:-1] + line[2:] + 2) >> 2
            f_left, f_above = fl[2 * n::-1].copy(), fl[2 * n:].copy()
            j["org_off"] = sum(o.size for o in orgs)
            j["org_stride"] = n
            j["refs_off"] = sum(r.size for r in refs)
            j["size"], j["bit_depth"] = n, bit_depth
            j["above_avail"], j["left_avail"] = int(rng.integers(0, 2)), int(rng.integers(0, 2))
            orgs.append(np.clip(org + rng.integers(-6, 7, size=org.shape), 0, maxv).astype(np.int16).reshape(-1))
            refs.append(np.concatenate([above, left, f_above, f_left]).astype(np.int32))
            jobs.append(j)
    return np.concatenate(jobs), np.concatenate(orgs), np.concatenate(refs)
