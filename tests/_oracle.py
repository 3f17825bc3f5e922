"""ctypes loaders for the CPU checkers (TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs may import this).

  oracle()  -> oracle/_build/libhoporacle.so, the C restatement (built on demand with gcc)
  ref()     -> oracle/_ref/libhopref.so, the compiled UNMODIFIED reference + harness, or None when it
               has not been built (it needs /root/reference at build time; the prebuilt file travels)
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_P = C.c_void_p
_cache = {}


def _np(a):
    return a.ctypes.data if a is not None else None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle"])


class _Checker:
    """Common call surface of the C restatement (prefix orc_) and the compiled reference (prefix ref_)."""

    def __init__(self, lib, prefix):
        import hevc_hop_b200 as hop
        self.hop = hop
        self.lib = lib
        self.prefix = prefix
        f = lambda n: getattr(lib, prefix + n)
        self._ps = f("pattern_search_batch"); self._ps.argtypes = [C.c_int, _P, _P, _P, _P]; self._ps.restype = None
        self._gt = f("pattern_search_gt_batch"); self._gt.argtypes = [C.c_int, _P, _P, _P, _P]; self._gt.restype = None
        self._dist = f("dist"); self._dist.argtypes = [_P, _P, _P]; self._dist.restype = C.c_uint32
        self._cpp = f("calc_param_projective"); self._cpp.argtypes = [_P, _P, _P, C.c_int, C.c_int]; self._cpp.restype = None
        self._pt = f("projective_transform"); self._pt.argtypes = [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int]; self._pt.restype = None

    def pattern_search(self, jobs, org, ref):
        hop = self.hop
        jobs = np.ascontiguousarray(jobs, dtype=hop.SEARCH_JOB_DT)
        out = np.zeros(len(jobs), dtype=hop.SEARCH_RES_DT)
        self._ps(len(jobs), _np(jobs), _np(org), _np(ref), _np(out))
        return out

    def pattern_search_gt(self, jobs, org, ref):
        hop = self.hop
        jobs = np.ascontiguousarray(jobs, dtype=hop.GT_JOB_DT)
        out = np.zeros(len(jobs), dtype=hop.GT_RES_DT)
        if self.prefix == "ref_" and len(jobs):
            # the reference's prologue interpolates 16 planes around the window (8-tap filters: a few rows and
            # columns beyond the 2W x 2H samples the search reads) although only plane [0][0] is used afterwards;
            # on compact per-PU windows those reads leave the buffer at its two ends, so give it guard bands
            guard = 8 * int(jobs["ref_stride"].max()) + 64
            ref = np.concatenate([np.zeros(guard, np.int16), np.ascontiguousarray(ref, dtype=np.int16).reshape(-1),
                                  np.zeros(guard, np.int16)])
            jobs = jobs.copy()
            jobs["ref_off"] += guard
        self._gt(len(jobs), _np(jobs), _np(org), _np(ref), _np(out))
        return out

    def dist(self, jobs, org, cur):
        jobs = np.ascontiguousarray(jobs, dtype=self.hop.DIST_JOB_DT)
        out = np.zeros(len(jobs), dtype=np.uint32)
        for i in range(len(jobs)):
            out[i] = self._dist(jobs[i:i + 1].ctypes.data, _np(org), _np(cur))
        return out

    def calc_param_projective(self, cx, cy, width, height):
        cx = np.ascontiguousarray(cx, dtype=np.int32); cy = np.ascontiguousarray(cy, dtype=np.int32)
        h = np.zeros(9, dtype=np.float64)
        self._cpp(_np(cx), _np(cy), _np(h), width, height)
        return h

    def projective_transform(self, window, cols, rows, h, nss_window):
        """window: (2*rows, 2*cols) int16 (already clamped); returns the rows x cols warped block."""
        window = np.ascontiguousarray(window, dtype=np.int16)
        aux = np.zeros((rows, cols), dtype=np.int16)
        h = np.ascontiguousarray(h, dtype=np.float64)
        stride = window.shape[1]
        base = window.ctypes.data + 2 * ((rows // 2) * stride + cols // 2)
        self._pt(base, _np(aux), _np(h), 2 * cols, 2 * rows, stride, nss_window)
        return aux


def oracle():
    if "orc" not in _cache:
        path = os.path.join(ORACLE_DIR, "_build", "libhoporacle.so")
        src = os.path.join(ORACLE_DIR, "hop_oracle.c")
        if os.environ.get("HOP_ORACLE_LIB"):          # e.g. the sanitizer build of `make -C oracle asan`
            path = os.path.abspath(os.environ["HOP_ORACLE_LIB"])
        elif not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build_oracle()
        lib = C.CDLL(path)
        chk = _Checker(lib, "orc_")
        lib.orc_component_bits.argtypes = [C.c_int32]; lib.orc_component_bits.restype = C.c_uint32
        lib.orc_get_bits_gt.argtypes = [C.c_int32] * 6; lib.orc_get_bits_gt.restype = C.c_uint32
        lib.orc_get_cost_xy.argtypes = [_P, C.c_int32, C.c_int32]; lib.orc_get_cost_xy.restype = C.c_uint32
        lib.orc_extend_border.argtypes = [_P, C.c_int, C.c_int, C.c_int, C.c_int]; lib.orc_extend_border.restype = None
        lib.orc_stage_window.argtypes = [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int]; lib.orc_stage_window.restype = None
        _cache["orc"] = chk
    return _cache["orc"]


def frac_search(jobs, org, ref, which="orc"):
    """xPatternSearchFracDIF: C restatement (which="orc") or the compiled reference (which="ref")."""
    chk = oracle() if which == "orc" else ref_lib()
    hop = oracle().hop
    jobs = np.ascontiguousarray(jobs, dtype=hop.FRAC_JOB_DT)
    out = np.zeros(len(jobs), dtype=hop.FRAC_RES_DT)
    fn = getattr(chk.lib, ("orc_" if which == "orc" else "ref_") + "frac_search_batch")
    fn.argtypes = [C.c_int, _P, _P, _P, _P]; fn.restype = None
    fn(len(jobs), _np(jobs), _np(org), _np(ref), _np(out))
    return out


def motion_search(jobs, org, ref):
    o = oracle()
    jobs = np.ascontiguousarray(jobs, dtype=o.hop.MOTION_JOB_DT)
    out = np.zeros(len(jobs), dtype=o.hop.MOTION_RES_DT)
    fn = o.lib.orc_motion_search_batch
    fn.argtypes = [C.c_int, _P, _P, _P, _P]; fn.restype = None
    fn(len(jobs), _np(jobs), _np(org), _np(ref), _np(out))
    return out


def predict(jobs, org, ref, dst_samples, which="orc"):
    """K6 (prediction + distortion / template cost): C restatement ("orc") or the compiled reference ("ref")."""
    chk = oracle() if which == "orc" else ref_lib()
    hop = oracle().hop
    jobs = np.ascontiguousarray(jobs, dtype=hop.PRED_JOB_DT)
    out = np.zeros(len(jobs), dtype=hop.PRED_RES_DT)
    dst = np.zeros(max(1, dst_samples), dtype=np.int16)
    fn = getattr(chk.lib, ("orc_" if which == "orc" else "ref_") + "predict_batch")
    fn.argtypes = [C.c_int, _P, _P, _P, _P, _P]; fn.restype = None
    fn(len(jobs), _np(jobs), _np(org), _np(ref), _np(dst), _np(out))
    return out, dst


def intra_prescreen(jobs, org, refs, which="orc"):
    """K7: 35 x (predIntraLumaAng + calcHAD) per PU: C restatement ("orc") or the compiled reference ("ref")."""
    chk = oracle() if which == "orc" else ref_lib()
    hop = oracle().hop
    jobs = np.ascontiguousarray(jobs, dtype=hop.INTRA_JOB_DT)
    refs = np.ascontiguousarray(refs, dtype=np.int32)
    out = np.zeros((len(jobs), hop.HOP_INTRA_MODES), dtype=np.uint32)
    fn = getattr(chk.lib, ("orc_" if which == "orc" else "ref_") + "intra_prescreen_batch")
    fn.argtypes = [C.c_int, _P, _P, _P, _P]; fn.restype = None
    fn(len(jobs), _np(jobs), _np(org), _np(refs), _np(out))
    return out


def ref_lib():
    return ref()


def gt_sweep(jobs, org, ref):
    """Oracle statement of the exhaustive sweep (reference mode IT_GT_SEARCH 1 / IT_GT_GRID_SIZE 1)."""
    o = oracle()
    hop = o.hop
    jobs = np.ascontiguousarray(jobs, dtype=hop.GT_JOB_DT)
    out = np.zeros(len(jobs), dtype=hop.GT_RES_DT)
    fn = o.lib.orc_gt_sweep_batch
    fn.argtypes = [C.c_int, _P, _P, _P, _P]; fn.restype = None
    fn(len(jobs), _np(jobs), _np(org), _np(ref), _np(out))
    return out


def gt_sweep_keys(jobs, org, ref, cand_begin, cand_end):
    o = oracle()
    jobs = np.ascontiguousarray(jobs, dtype=o.hop.GT_JOB_DT)
    keys = np.zeros(len(jobs), dtype=np.uint64)
    fn = o.lib.orc_gt_sweep_keys_batch
    fn.argtypes = [C.c_int, _P, _P, _P, C.c_int, C.c_int, _P]; fn.restype = None
    fn(len(jobs), _np(jobs), _np(org), _np(ref), cand_begin, cand_end, _np(keys))
    return keys


def gt_sweep_finalize(jobs, keys):
    o = oracle()
    jobs = np.ascontiguousarray(jobs, dtype=o.hop.GT_JOB_DT)
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    out = np.zeros(len(jobs), dtype=o.hop.GT_RES_DT)
    fn = o.lib.orc_gt_sweep_finalize
    fn.argtypes = [_P, C.c_uint64, _P]; fn.restype = None
    for i in range(len(jobs)):
        fn(jobs[i:i + 1].ctypes.data, int(keys[i]), out[i:i + 1].ctypes.data)
    return out


def ref_sweep():
    """The reference compiled in its exhaustive mode (oracle/_ref/libhopref_sweep.so): its
    xPatternSearchGT IS the sweep.  None when not built."""
    if "ref_sweep" not in _cache:
        p = os.path.join(ORACLE_DIR, "_ref", "libhopref_sweep.so")
        _cache["ref_sweep"] = _Checker(C.CDLL(p), "ref_") if os.path.exists(p) else None
    return _cache["ref_sweep"]


REF_ENCODER = os.path.join(ORACLE_DIR, "_ref", "TAppEncoderRef")   # the unmodified CPU reference encoder
REF_DECODER = os.path.join(ORACLE_DIR, "_ref", "TAppDecoderRef")


def encode_reference(width, height, **kw):
    """Encode a synthetic lenslet frame with the unmodified CPU reference (checker / baseline only).

    The unmodified reference dies with SIGSEGV now and then (tools/flake_probe.py: 1-2 of 45 runs of the 136x104
    case even with one retry; it reads beyond its buffers, e.g. the interpolation prologue of xPatternSearchGT and
    AMVP start vectors that leave the plane, so the outcome depends on the heap layout).  Its OUTPUT is stable
    (44 of 44 identical), so the checker is simply started again; the GPU-backed encoder gets no retries."""
    from hevc_hop_b200 import encoder
    return encoder.encode(REF_ENCODER, width, height, retries=5, **kw)


def ref_path():
    return os.path.join(ORACLE_DIR, "_ref", "libhopref.so")


def ref():
    if "ref" not in _cache:
        if not os.path.exists(ref_path()):
            _cache["ref"] = None
        else:
            lib = C.CDLL(ref_path())
            chk = _Checker(lib, "ref_")
            lib.ref_bits_gt.argtypes = [C.c_int] * 6; lib.ref_bits_gt.restype = C.c_uint32
            lib.ref_get_cost_xy.argtypes = [_P, C.c_int, C.c_int]; lib.ref_get_cost_xy.restype = C.c_uint32
            lib.ref_extend_border.argtypes = [_P, C.c_int, C.c_int]; lib.ref_extend_border.restype = C.c_int
            _cache["ref"] = chk
    return _cache["ref"]


def extend_border_oracle(plane, pic_w, pic_h, margin):
    """plane: (pic_h+2m, pic_w+2m) int16, modified in place."""
    stride = plane.shape[1]
    origin = plane.ctypes.data + 2 * (margin * stride + margin)
    oracle().lib.orc_extend_border(origin, stride, pic_w, pic_h, margin)
    return plane
