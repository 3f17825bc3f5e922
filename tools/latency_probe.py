"""Single-PU call latency through the host ABI (the path the encoder takes), per stage and PU shape.

  python tools/latency_probe.py [reps]

Each line: median microseconds of one call with one job for K1 only (hop_pattern_search_batch), the fractional
refinement only, K2 only (hop_pattern_search_gt_batch) and the fused motion search (hop_motion_search_batch).
"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import conftest  # noqa: F401  (registers the package)
import hevc_hop_b200 as hop
from hevc_hop_b200.workload import PuBatch


def med(fn, reps):
    for _ in range(20):
        fn()
    t = []
    for _ in range(reps):
        a = time.perf_counter()
        fn()
        t.append(time.perf_counter() - a)
    return 1e6 * float(np.median(t))


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    ctx = hop.HopContext(0)
    print("shape      K1     frac   K2     fused   (us per single-PU call, median of %d)" % reps)
    for c, r in [(8, 4), (4, 8), (8, 8), (16, 8), (16, 16), (32, 32), (64, 64)]:
        b = PuBatch(c, r, 1, seed=5, sr=64, n_start=3)
        sj, gj, fj, mj = b.search_jobs, b.gt_jobs, b.frac_jobs(), b.motion_jobs()
        # the encoder's situation: the reference plane is resident, only the job and the original block travel
        ctx.ref_create(b.pw, b.ph, 0)
        ctx.ref_upload(np.ascontiguousarray(b.ref).reshape(b.ph, b.pw))
        want = ctx.motion_search(mj, b.org, b.ref)
        assert ctx.motion_search(mj, b.org, None).tobytes() == want.tobytes()
        t1 = med(lambda: ctx.pattern_search(sj, b.org, None), reps)
        t2 = med(lambda: ctx.frac_search(fj, b.org, None), reps)
        t3 = med(lambda: ctx.pattern_search_gt(gj, b.org, None), reps)
        t4 = med(lambda: ctx.motion_search(mj, b.org, None), reps)
        print("%2dx%-2d   %6.1f %6.1f %6.1f %6.1f" % (c, r, t1, t2, t3, t4))


if __name__ == "__main__":
    main()
