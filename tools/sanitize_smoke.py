"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

hop = g.load_package()
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import _oracle
from hevc_hop_b200.workload import PuBatch, dist_jobs

ctx = hop.HopContext(0)
orc = _oracle.oracle()
for (c, r) in [(8, 8), (16, 12), (8, 4), (32, 32)]:
    b = PuBatch(c, r, 2, seed=c + r, sr=max(24, r + 16), n_start=3)
    assert ctx.pattern_search(b.search_jobs, b.org, b.ref).tobytes() == orc.pattern_search(b.search_jobs, b.org, b.ref).tobytes()
    assert ctx.pattern_search_gt(b.gt_jobs, b.org, b.ref).tobytes() == orc.pattern_search_gt(b.gt_jobs, b.org, b.ref).tobytes()
    fj = b.frac_jobs()
    assert (ctx.frac_search(fj, b.org, b.ref)["cost"] == _oracle.frac_search(fj, b.org, b.ref)["cost"]).all()
    mj = b.motion_jobs()
    assert ctx.motion_search(mj, b.org, b.ref).tobytes() == _oracle.motion_search(mj, b.org, b.ref).tobytes()
    b10 = PuBatch(c, r, 1, seed=1, sr=24, bit_depth=10)
    assert ctx.pattern_search(b10.search_jobs, b10.org, b10.ref).tobytes() == orc.pattern_search(b10.search_jobs, b10.org, b10.ref).tobytes()
b = PuBatch(8, 8, 2, seed=3, sr=20, n_start=1)
assert (ctx.gt_sweep(b.gt_jobs, b.org, b.ref)["cost"] == _oracle.gt_sweep(b.gt_jobs, b.org, b.ref)["cost"]).all()
# batched K2 forms: two tiles per register tile on a window with difference rows, incl. the 896-thread 64x64 class, the
# sweep on a pairing shape, and one mixed-shape batch sized for its largest PU
for (c, r) in [(64, 64), (64, 16), (16, 64), (32, 24), (16, 8)]:
    b = PuBatch(c, r, 2, seed=5 + c, sr=max(24, r + 16), n_start=2)
    batch = ctx.pattern_search_gt(b.gt_jobs, b.org, b.ref)
    single = np.concatenate([ctx.pattern_search_gt(b.gt_jobs[k:k + 1], b.org, b.ref) for k in range(2)])
    assert batch.tobytes() == single.tobytes(), (c, r)
b = PuBatch(16, 16, 2, seed=4, sr=20, n_start=1, bit_depth=10)
assert ctx.gt_sweep(b.gt_jobs, b.org, b.ref).tobytes() == _oracle.gt_sweep(b.gt_jobs, b.org, b.ref).tobytes()
parts = [PuBatch(c, r, 1, seed=9 + c, sr=24, n_start=2) for (c, r) in [(8, 8), (64, 32), (16, 16), (8, 4)]]
jobs, orgs, refs = [], [], []
for q in parts:
    j = q.gt_jobs.copy()
    j["org_off"] += sum(o.size for o in orgs); j["ref_off"] += sum(x.size for x in refs)
    jobs.append(j); orgs.append(q.org); refs.append(q.ref)
mixed = ctx.pattern_search_gt(np.concatenate(jobs), np.concatenate(orgs), np.concatenate(refs))
assert mixed.tobytes() == np.concatenate([orc.pattern_search_gt(q.gt_jobs, q.org, q.ref) for q in parts]).tobytes()
jobs, org, cur = dist_jobs(16, 8, 5, func=hop.HOP_DF_HADS)
assert (ctx.dist(jobs, org, cur) == orc.dist(jobs, org, cur)).all()
# mirror + single-call latency path (clusters)
rng = np.random.default_rng(1)
pic_w, pic_h, m = 128, 128, 80
ctx.ref_create(pic_w, pic_h, m)
ctx.ref_reset(-1)
img = rng.integers(0, 256, size=(pic_h, pic_w)).astype(np.int16)
ctx.ref_update(0, 0, img[:64])
ctx.ref_update(0, 64, img[64:, :64])
stride = pic_w + 2 * m
for (c, r) in [(16, 16), (8, 4), (32, 32), (64, 64)]:
    b = PuBatch(c, r, 1, seed=c, sr=16, n_start=2)
    mj = b.motion_jobs()
    s = mj["search"]; s["ref_stride"] = stride; s["ref_off"] = 64 * stride + 64
    s["rng_left"], s["rng_right"], s["rng_top"], s["rng_bottom"] = -60, 30, -60, -4
    mj["search"] = s
    ctx.motion_search(mj, b.org, None)
print("sanitize smoke ok, launches", ctx.launch_count)
