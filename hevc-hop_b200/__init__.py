"""hevc_hop_b200 -- Python host mirror of the libhopgpu C ABI (include/hop_gpu.h).

The product is the CUDA library `libhopgpu.so` (sources in csrc/); this module only binds it with
ctypes so that tests and bench.py can drive exactly the entry points the HM encoder shim binds.
There is NO CPU fallback: importing works without the library (so CPU-only unit tests of the host
logic can run), but every compute call raises HopError if the library or a CUDA device is missing.

The directory name carries a hyphen (`hevc-hop_b200/`, mandated layout); `load_package()` in
`__graft_entry__.py` / `tests/conftest.py` registers it under the importable name `hevc_hop_b200`.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HOP_LIB", os.path.join(_HERE, "libhopgpu.so"))   # HOP_LIB: experimental builds

HOP_MAX_UINT = 0xFFFFFFFF
HOP_NOT_VALID = -1
HOP_DF_SAD = 8
HOP_DF_HADS = 22
HOP_MAX_PRED = 3
HOP_SWEEP_CANDS = 7200


class HopError(RuntimeError):
    pass


class HopMv(C.Structure):
    _fields_ = [("hor", C.c_int16), ("ver", C.c_int16)]


class HopCostState(C.Structure):
    _fields_ = [("lambda_cost", C.c_uint32), ("cost_scale", C.c_int32), ("pred", HopMv)]


class HopSearchJob(C.Structure):
    _fields_ = [
        ("org_off", C.c_int64), ("ref_off", C.c_int64),
        ("org_stride", C.c_int32), ("ref_stride", C.c_int32),
        ("cols", C.c_int32), ("rows", C.c_int32),
        ("rng_left", C.c_int32), ("rng_top", C.c_int32), ("rng_right", C.c_int32), ("rng_bottom", C.c_int32),
        ("offset_x", C.c_int32), ("offset_y", C.c_int32),
        ("is_ss", C.c_int32), ("fast_enc", C.c_int32), ("bit_depth", C.c_int32),
        ("cost", HopCostState),
    ]


class HopSearchResult(C.Structure):
    _fields_ = [("found", C.c_int32), ("mv", HopMv), ("sad", C.c_uint32), ("cost", C.c_uint32)]


class HopGtJob(C.Structure):
    _fields_ = [
        ("org_off", C.c_int64), ("ref_off", C.c_int64),
        ("org_stride", C.c_int32), ("ref_stride", C.c_int32),
        ("cols", C.c_int32), ("rows", C.c_int32),
        ("ss_cand", HopMv), ("num_pred", C.c_int32), ("amvp", HopMv * HOP_MAX_PRED),
        ("threshold", C.c_uint32), ("use_had", C.c_int32), ("bit_depth", C.c_int32),
        ("cost", HopCostState),
    ]


class HopGtResult(C.Structure):
    _fields_ = [
        ("gt_flag", C.c_int32), ("gt", HopMv * 4), ("cost", C.c_uint32), ("mv_int", HopMv),
        ("best_index", C.c_int32), ("n_candidates", C.c_uint32),
    ]


class HopFracJob(C.Structure):
    _fields_ = [
        ("org_off", C.c_int64), ("ref_off", C.c_int64),
        ("org_stride", C.c_int32), ("ref_stride", C.c_int32),
        ("cols", C.c_int32), ("rows", C.c_int32),
        ("mv_int", HopMv), ("use_had", C.c_int32), ("bit_depth", C.c_int32), ("cost", HopCostState),
    ]


class HopFracResult(C.Structure):
    _fields_ = [("half", HopMv), ("qter", HopMv), ("cost", C.c_uint32), ("cost_half", C.c_uint32)]


class HopMotionJob(C.Structure):
    _fields_ = [("search", HopSearchJob), ("use_had", C.c_int32), ("use_gt", C.c_int32), ("num_pred", C.c_int32),
                ("amvp", HopMv * HOP_MAX_PRED)]


class HopMotionResult(C.Structure):
    _fields_ = [("search", HopSearchResult), ("refined", C.c_int32), ("frac", HopFracResult), ("gt", HopGtResult)]


class HopPredJob(C.Structure):
    _fields_ = [
        ("ref_off", C.c_int64), ("org_off", C.c_int64), ("dst_off", C.c_int64),
        ("ref_stride", C.c_int32), ("org_stride", C.c_int32), ("dst_stride", C.c_int32),
        ("cols", C.c_int32), ("rows", C.c_int32), ("comp", C.c_int32), ("mv", HopMv),
        ("gt_flag", C.c_int32), ("gt", HopMv * 4), ("bit_depth", C.c_int32), ("dist_func", C.c_int32),
        ("template_cost", C.c_int32), ("is_ss", C.c_int32), ("mv_probe", HopMv), ("mvp_bits", C.c_uint32),
        ("lambda_sad", C.c_uint32),
    ]


class HopPredResult(C.Structure):
    _fields_ = [("valid", C.c_int32), ("dist", C.c_uint32), ("cost", C.c_uint32)]


class HopIntraJob(C.Structure):
    _fields_ = [("org_off", C.c_int64), ("refs_off", C.c_int64), ("org_stride", C.c_int32), ("size", C.c_int32),
                ("above_avail", C.c_int32), ("left_avail", C.c_int32), ("bit_depth", C.c_int32), ("reserved", C.c_int32)]


class HopCtxStats(C.Structure):
    _fields_ = [("single_calls", C.c_uint64), ("cache_hits", C.c_uint64), ("cache_misses", C.c_uint64),
                ("prefetched", C.c_uint64), ("prefetch_dropped", C.c_uint64), ("candidates", C.c_uint64)]


class HopDistJob(C.Structure):
    _fields_ = [
        ("org_off", C.c_int64), ("cur_off", C.c_int64),
        ("org_stride", C.c_int32), ("cur_stride", C.c_int32),
        ("cols", C.c_int32), ("rows", C.c_int32),
        ("func", C.c_int32), ("sub_shift", C.c_int32), ("bit_depth", C.c_int32),
    ]


# numpy views of the same layouts (for bulk job construction and result comparison)
MV_DT = np.dtype([("hor", "<i2"), ("ver", "<i2")])
COST_DT = np.dtype([("lambda_cost", "<u4"), ("cost_scale", "<i4"), ("pred", MV_DT)])
SEARCH_JOB_DT = np.dtype([
    ("org_off", "<i8"), ("ref_off", "<i8"), ("org_stride", "<i4"), ("ref_stride", "<i4"),
    ("cols", "<i4"), ("rows", "<i4"),
    ("rng_left", "<i4"), ("rng_top", "<i4"), ("rng_right", "<i4"), ("rng_bottom", "<i4"),
    ("offset_x", "<i4"), ("offset_y", "<i4"), ("is_ss", "<i4"), ("fast_enc", "<i4"), ("bit_depth", "<i4"),
    ("cost", COST_DT)], align=True)
SEARCH_RES_DT = np.dtype([("found", "<i4"), ("mv", MV_DT), ("sad", "<u4"), ("cost", "<u4")], align=True)
GT_JOB_DT = np.dtype([
    ("org_off", "<i8"), ("ref_off", "<i8"), ("org_stride", "<i4"), ("ref_stride", "<i4"),
    ("cols", "<i4"), ("rows", "<i4"), ("ss_cand", MV_DT), ("num_pred", "<i4"), ("amvp", MV_DT, (HOP_MAX_PRED,)),
    ("threshold", "<u4"), ("use_had", "<i4"), ("bit_depth", "<i4"), ("cost", COST_DT)], align=True)
GT_RES_DT = np.dtype([("gt_flag", "<i4"), ("gt", MV_DT, (4,)), ("cost", "<u4"), ("mv_int", MV_DT),
                      ("best_index", "<i4"), ("n_candidates", "<u4")], align=True)
DIST_JOB_DT = np.dtype([
    ("org_off", "<i8"), ("cur_off", "<i8"), ("org_stride", "<i4"), ("cur_stride", "<i4"),
    ("cols", "<i4"), ("rows", "<i4"), ("func", "<i4"), ("sub_shift", "<i4"), ("bit_depth", "<i4")], align=True)

FRAC_JOB_DT = np.dtype([
    ("org_off", "<i8"), ("ref_off", "<i8"), ("org_stride", "<i4"), ("ref_stride", "<i4"),
    ("cols", "<i4"), ("rows", "<i4"), ("mv_int", MV_DT), ("use_had", "<i4"), ("bit_depth", "<i4"), ("cost", COST_DT)], align=True)
FRAC_RES_DT = np.dtype([("half", MV_DT), ("qter", MV_DT), ("cost", "<u4"), ("cost_half", "<u4")], align=True)
MOTION_JOB_DT = np.dtype([("search", SEARCH_JOB_DT), ("use_had", "<i4"), ("use_gt", "<i4"), ("num_pred", "<i4"),
                          ("amvp", MV_DT, (HOP_MAX_PRED,))], align=True)
MOTION_RES_DT = np.dtype([("search", SEARCH_RES_DT), ("refined", "<i4"), ("frac", FRAC_RES_DT), ("gt", GT_RES_DT)], align=True)

PRED_JOB_DT = np.dtype([
    ("ref_off", "<i8"), ("org_off", "<i8"), ("dst_off", "<i8"), ("ref_stride", "<i4"), ("org_stride", "<i4"), ("dst_stride", "<i4"),
    ("cols", "<i4"), ("rows", "<i4"), ("comp", "<i4"), ("mv", MV_DT), ("gt_flag", "<i4"), ("gt", MV_DT, (4,)),
    ("bit_depth", "<i4"), ("dist_func", "<i4"), ("template_cost", "<i4"), ("is_ss", "<i4"), ("mv_probe", MV_DT),
    ("mvp_bits", "<u4"), ("lambda_sad", "<u4")], align=True)
INTRA_JOB_DT = np.dtype([("org_off", "<i8"), ("refs_off", "<i8"), ("org_stride", "<i4"), ("size", "<i4"), ("above_avail", "<i4"),
                         ("left_avail", "<i4"), ("bit_depth", "<i4"), ("reserved", "<i4")], align=True)
HOP_INTRA_MODES = 35
assert INTRA_JOB_DT.itemsize == C.sizeof(HopIntraJob) == 40
PRED_RES_DT = np.dtype([("valid", "<i4"), ("dist", "<u4"), ("cost", "<u4")], align=True)
assert PRED_JOB_DT.itemsize == C.sizeof(HopPredJob) == 104, (PRED_JOB_DT.itemsize, C.sizeof(HopPredJob))
assert PRED_RES_DT.itemsize == C.sizeof(HopPredResult) == 12
assert FRAC_JOB_DT.itemsize == C.sizeof(HopFracJob), (FRAC_JOB_DT.itemsize, C.sizeof(HopFracJob))
assert FRAC_RES_DT.itemsize == C.sizeof(HopFracResult) == 16
assert MOTION_JOB_DT.itemsize == C.sizeof(HopMotionJob), (MOTION_JOB_DT.itemsize, C.sizeof(HopMotionJob))
assert MOTION_RES_DT.itemsize == C.sizeof(HopMotionResult), (MOTION_RES_DT.itemsize, C.sizeof(HopMotionResult))
assert SEARCH_JOB_DT.itemsize == C.sizeof(HopSearchJob) == 80
assert SEARCH_RES_DT.itemsize == C.sizeof(HopSearchResult) == 16
assert GT_JOB_DT.itemsize == C.sizeof(HopGtJob) == 80
assert GT_RES_DT.itemsize == C.sizeof(HopGtResult) == 36
assert DIST_JOB_DT.itemsize == C.sizeof(HopDistJob) == 48

# every symbol include/hop_gpu.h declares: (name, restype, argtypes)
_P = C.c_void_p
ABI = [
    ("hop_abi_version", C.c_int, []),
    ("hop_last_error", C.c_char_p, []),
    ("hop_device_count", C.c_int, []),
    ("hop_ctx_create", C.c_int, [C.c_int, C.POINTER(_P)]),
    ("hop_ctx_destroy", None, [_P]),
    ("hop_ctx_sync", C.c_int, [_P]),
    ("hop_ctx_stream", _P, [_P]),
    ("hop_shape_supported", C.c_int, [C.c_int, C.c_int]),
    ("hop_ref_create", C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    ("hop_ref_reset", C.c_int, [_P, C.c_int]),
    ("hop_ref_upload", C.c_int, [_P, _P, C.c_size_t]),
    ("hop_ref_update", C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int]),
    ("hop_ref_download", C.c_int, [_P, _P, C.c_size_t]),
    ("hop_ref_stride", C.c_int, [_P]),
    ("hop_ref_origin_dev", _P, [_P]),
    ("hop_pattern_search_batch", C.c_int, [_P, C.c_int, _P, _P, C.c_size_t, _P, C.c_size_t, _P]),
    ("hop_pattern_search_gt_batch", C.c_int, [_P, C.c_int, _P, _P, C.c_size_t, _P, C.c_size_t, _P]),
    ("hop_pattern_search_gt_batch_async", C.c_int, [_P, C.c_int, _P, _P, C.c_size_t, _P, C.c_size_t, _P]),
    ("hop_dist_batch", C.c_int, [_P, C.c_int, _P, _P, C.c_size_t, _P, C.c_size_t, _P]),
    ("hop_frac_search_batch", C.c_int, [_P, C.c_int, _P, _P, C.c_size_t, _P, C.c_size_t, _P]),
    ("hop_motion_search_batch", C.c_int, [_P, C.c_int, _P, _P, C.c_size_t, _P, C.c_size_t, _P]),
    ("hop_predict_batch", C.c_int, [_P, C.c_int, _P, _P, C.c_size_t, _P, C.c_size_t, _P, C.c_size_t, _P]),
    ("hop_intra_prescreen_batch", C.c_int, [_P, C.c_int, _P, _P, C.c_size_t, _P, C.c_size_t, _P]),
    ("hop_motion_search_prefetch", C.c_int, [_P, C.c_int, _P, _P, C.c_size_t]),
    ("hop_pattern_search_batch_dev", C.c_int, [_P, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    ("hop_pattern_search_gt_batch_dev", C.c_int, [_P, C.c_int, _P, _P, _P, C.c_size_t, _P, C.c_int, C.c_int, _P]),
    ("hop_dist_batch_dev", C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P]),
    ("hop_gt_sweep_keys_dev", C.c_int, [_P, C.c_int, _P, _P, _P, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    ("hop_gt_sweep_finalize_dev", C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P]),
    ("hop_sweep_exchange_create", C.c_int, [_P, C.c_int, _P]),
    ("hop_sweep_exchange_connect", C.c_int, [_P, C.c_int, C.c_int, _P]),
    ("hop_gt_sweep_sharded_dev", C.c_int, [_P, C.c_int, _P, _P, _P, C.c_size_t, C.c_int, C.c_int, _P, _P]),
    ("hop_gt_sweep_batch", C.c_int, [_P, C.c_int, _P, _P, C.c_size_t, _P, C.c_size_t, _P]),
    ("hop_ctx_launch_count", C.c_uint64, [_P]),
    ("hop_ctx_stats", C.c_int, [_P, _P]),
    ("hop_probe_alu", C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
]

_lib = None


def load_library(path=None):
    """dlopen libhopgpu.so and type every ABI symbol.  Raises HopError when the library is missing
    (run `python -c 'import __graft_entry__ as g; g.build()'` first) -- never falls back."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise HopError("libhopgpu.so not built (%s); run __graft_entry__.build(). There is no CPU fallback." % p)
    lib = C.CDLL(p)
    for name, res, args in ABI:
        fn = getattr(lib, name)      # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    return int(a)   # raw device / host address


class HopContext:
    """One encoder context (= one GPU).  Thin object wrapper over the hop_ctx_* / hop_* C functions."""

    def __init__(self, device=0, lib_path=None):
        self.lib = load_library(lib_path)   # lib_path: another build of the same ABI (tests: the refops twin)
        h = _P()
        self._check(self.lib.hop_ctx_create(int(device), C.byref(h)))
        self.h = h
        self.device = device

    def _check(self, st):
        if st != 0:
            msg = self.lib.hop_last_error()
            raise HopError("libhopgpu status %d: %s" % (st, msg.decode() if msg else "?"))

    def close(self):
        if getattr(self, "h", None):
            self.lib.hop_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._check(self.lib.hop_ctx_sync(self.h))

    @property
    def stream(self):
        return self.lib.hop_ctx_stream(self.h)

    @property
    def launch_count(self):
        return int(self.lib.hop_ctx_launch_count(self.h))

    # ---- SS reference mirror (K4) ----
    def ref_create(self, pic_w, pic_h, margin=80):
        self._check(self.lib.hop_ref_create(self.h, pic_w, pic_h, margin))
        self._ref_shape = (pic_h + 2 * margin, pic_w + 2 * margin)

    def ref_reset(self, value=HOP_NOT_VALID):
        self._check(self.lib.hop_ref_reset(self.h, value))

    def ref_upload(self, plane):
        plane = np.ascontiguousarray(plane, dtype=np.int16)
        self._check(self.lib.hop_ref_upload(self.h, _ptr(plane), plane.size))

    def ref_update(self, x, y, block):
        block = np.ascontiguousarray(block, dtype=np.int16)
        h, w = block.shape
        self._check(self.lib.hop_ref_update(self.h, x, y, w, h, _ptr(block), w))

    def ref_download(self):
        out = np.empty(self._ref_shape, dtype=np.int16)
        self._check(self.lib.hop_ref_download(self.h, _ptr(out), out.size))
        return out

    # ---- host entry points ----
    def pattern_search(self, jobs, org, ref):
        jobs = np.ascontiguousarray(jobs, dtype=SEARCH_JOB_DT)
        out = np.zeros(len(jobs), dtype=SEARCH_RES_DT)
        self._check(self.lib.hop_pattern_search_batch(
            self.h, len(jobs), _ptr(jobs), _ptr(org), org.size, _ptr(ref), 0 if ref is None else ref.size, _ptr(out)))
        return out

    def pattern_search_gt(self, jobs, org, ref):
        jobs = np.ascontiguousarray(jobs, dtype=GT_JOB_DT)
        out = np.zeros(len(jobs), dtype=GT_RES_DT)
        self._check(self.lib.hop_pattern_search_gt_batch(
            self.h, len(jobs), _ptr(jobs), _ptr(org), org.size, _ptr(ref), 0 if ref is None else ref.size, _ptr(out)))
        return out

    def frac_search(self, jobs, org, ref):
        jobs = np.ascontiguousarray(jobs, dtype=FRAC_JOB_DT)
        out = np.zeros(len(jobs), dtype=FRAC_RES_DT)
        self._check(self.lib.hop_frac_search_batch(
            self.h, len(jobs), _ptr(jobs), _ptr(org), org.size, _ptr(ref), 0 if ref is None else ref.size, _ptr(out)))
        return out

    def motion_search(self, jobs, org, ref):
        jobs = np.ascontiguousarray(jobs, dtype=MOTION_JOB_DT)
        out = np.zeros(len(jobs), dtype=MOTION_RES_DT)
        self._check(self.lib.hop_motion_search_batch(
            self.h, len(jobs), _ptr(jobs), _ptr(org), org.size, _ptr(ref), 0 if ref is None else ref.size, _ptr(out)))
        return out

    def predict(self, jobs, org, ref, dst_samples=0):
        """K6: motion-compensated prediction (+ distortion / template cost).  Returns (results, dst buffer)."""
        jobs = np.ascontiguousarray(jobs, dtype=PRED_JOB_DT)
        out = np.zeros(len(jobs), dtype=PRED_RES_DT)
        dst = np.zeros(max(1, dst_samples), dtype=np.int16)
        self._check(self.lib.hop_predict_batch(
            self.h, len(jobs), _ptr(jobs), _ptr(org), org.size, _ptr(ref), 0 if ref is None else ref.size,
            _ptr(dst) if dst_samples else None, dst_samples, _ptr(out)))
        return out, dst

    def intra_prescreen(self, jobs, org, refs):
        """K7: Hadamard cost of the 35 intra predictions of every PU -> (n, 35) uint32."""
        jobs = np.ascontiguousarray(jobs, dtype=INTRA_JOB_DT)
        refs = np.ascontiguousarray(refs, dtype=np.int32)
        out = np.zeros((len(jobs), HOP_INTRA_MODES), dtype=np.uint32)
        self._check(self.lib.hop_intra_prescreen_batch(self.h, len(jobs), _ptr(jobs), _ptr(org), org.size, _ptr(refs), refs.size, _ptr(out)))
        return out

    def motion_prefetch(self, jobs, org):
        """Enqueue speculative single-PU motion searches against the SS mirror (hop_motion_search_prefetch)."""
        jobs = np.ascontiguousarray(jobs, dtype=MOTION_JOB_DT)
        self._check(self.lib.hop_motion_search_prefetch(self.h, len(jobs), _ptr(jobs), _ptr(org), org.size))

    def stats(self):
        st = HopCtxStats()
        self._check(self.lib.hop_ctx_stats(self.h, C.byref(st)))
        return {k: int(getattr(st, k)) for k, _ in HopCtxStats._fields_}

    def dist(self, jobs, org, cur):
        jobs = np.ascontiguousarray(jobs, dtype=DIST_JOB_DT)
        out = np.zeros(len(jobs), dtype=np.uint32)
        self._check(self.lib.hop_dist_batch(
            self.h, len(jobs), _ptr(jobs), _ptr(org), org.size, _ptr(cur), cur.size, _ptr(out)))
        return out

    # ---- device entry points (addresses of HBM-resident buffers, e.g. torch tensors' data_ptr()) ----
    def pattern_search_dev(self, n, d_jobs, d_org, d_ref, d_out, stream=None, cols=0, rows=0, nx_max=0, ny_max=0):
        """cols / rows / nx_max / ny_max: the batch's single PU shape and window bound (zeros = mixed batch)."""
        self._check(self.lib.hop_pattern_search_batch_dev(self.h, n, d_jobs, d_org, d_ref, d_out, cols, rows, nx_max, ny_max, stream))

    def pattern_search_gt_dev(self, n, d_jobs, d_org, d_ref, ref_samples, d_out, max_cols, max_rows, stream=None):
        """ref_samples: int16 samples addressable behind d_ref (bounds the window reads)."""
        self._check(self.lib.hop_pattern_search_gt_batch_dev(self.h, n, d_jobs, d_org, d_ref, ref_samples, d_out,
                                                             max_cols, max_rows, stream))

    def dist_dev(self, n, d_jobs, d_org, d_cur, d_out, stream=None):
        self._check(self.lib.hop_dist_batch_dev(self.h, n, d_jobs, d_org, d_cur, d_out, stream))

    # ---- exhaustive sweep ----
    def gt_sweep(self, jobs, org, ref):
        jobs = np.ascontiguousarray(jobs, dtype=GT_JOB_DT)
        out = np.zeros(len(jobs), dtype=GT_RES_DT)
        self._check(self.lib.hop_gt_sweep_batch(
            self.h, len(jobs), _ptr(jobs), _ptr(org), org.size, _ptr(ref), 0 if ref is None else ref.size, _ptr(out)))
        return out

    def gt_sweep_keys_dev(self, n, d_jobs, d_org, d_ref, ref_samples, max_cols, max_rows, cand_begin, cand_end, d_keys,
                          d_counts=None, stream=None):
        self._check(self.lib.hop_gt_sweep_keys_dev(self.h, n, d_jobs, d_org, d_ref, ref_samples, max_cols, max_rows,
                                                   cand_begin, cand_end, d_keys, d_counts, stream))

    def gt_sweep_finalize_dev(self, n, d_jobs, d_keys, d_counts, d_out, stream=None):
        self._check(self.lib.hop_gt_sweep_finalize_dev(self.h, n, d_jobs, d_keys, d_counts, d_out, stream))

    def sweep_exchange_create(self, max_pus):
        """-> 64-byte handle of this rank's merge words (to be all-gathered over the ranks)."""
        h = (C.c_ubyte * 64)()
        self._check(self.lib.hop_sweep_exchange_create(self.h, int(max_pus), C.byref(h)))
        return bytes(h)

    def sweep_exchange_connect(self, world, rank, handles):
        buf = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles))
        self._check(self.lib.hop_sweep_exchange_connect(self.h, world, rank, C.byref(buf)))

    def gt_sweep_sharded_dev(self, n, d_jobs, d_org, d_ref, ref_samples, max_cols, max_rows, d_out, stream=None):
        self._check(self.lib.hop_gt_sweep_sharded_dev(self.h, n, d_jobs, d_org, d_ref, ref_samples, max_cols, max_rows, d_out, stream))

    def probe_alu(self, what):
        g, ms = C.c_double(), C.c_double()
        self._check(self.lib.hop_probe_alu(self.h, what, C.byref(g), C.byref(ms)))
        return g.value, ms.value
