"""One launch series per kernel family, small enough for `ncu --set full` (40 replays per launch):
    python tools/ncu_targets.py [k2] [k1] [sweep] [tail] [pred] [intra] [dist] [ref]      (default: all)
Every family runs `WARM` untimed launches first, then the launches ncu should capture; prints CUDA-event times."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import conftest  # noqa: F401
import numpy as np, torch
import hevc_hop_b200 as hop
from hevc_hop_b200.workload import PuBatch, GtBatch
from hevc_hop_b200.lenslet import lenslet_luma

WARM = int(os.environ.get("NCU_WARM", "1"))
what = set(sys.argv[1:]) or {"k2", "k1", "sweep", "tail", "pred", "intra", "dist", "ref"}
ctx = hop.HopContext(0)
dev = torch.device("cuda", 0)
ts = torch.cuda.ExternalStream(ctx.stream, device=dev)
src = lenslet_luma(1024, 1024, seed=99).astype(np.int16)
up = lambda a: torch.from_numpy(a.view(np.uint8)).to(dev)


def timed(name, fn, iters=WARM + 1):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    for i in range(iters):
        e[i].record(ts)
        fn()
    e[iters].record(ts)
    torch.cuda.synchronize()
    print(name, "ms per launch", [round(e[i].elapsed_time(e[i + 1]), 4) for i in range(iters)], flush=True)


if "k2" in what:
    for (c, r) in [(8, 8), (16, 16), (32, 32), (64, 64)]:
        b = GtBatch(c, r, 1184, seed=7, source=src)
        dj, do, dr = up(b.gt_jobs), up(b.org), up(b.ref)
        out = torch.zeros(b.n * hop.GT_RES_DT.itemsize, dtype=torch.uint8, device=dev)
        timed("k2 %dx%d x%d" % (c, r, b.n), lambda: ctx.pattern_search_gt_dev(
            b.n, dj.data_ptr(), do.data_ptr(), dr.data_ptr(), b.ref.size, out.data_ptr(), c, r, ctx.stream))

if "k1" in what:
    for (c, r) in [(8, 8), (16, 16), (32, 32), (64, 64)]:
        b = PuBatch(c, r, 592, seed=300, source=src)
        dj, do, dr = up(b.search_jobs), up(b.org), up(b.ref)
        out = torch.zeros(b.n * hop.SEARCH_RES_DT.itemsize, dtype=torch.uint8, device=dev)
        timed("k1 %dx%d x%d" % (c, r, b.n), lambda: ctx.pattern_search_dev(
            b.n, dj.data_ptr(), do.data_ptr(), dr.data_ptr(), out.data_ptr(), ctx.stream,
            cols=c, rows=r, nx_max=2 * b.sr + 1, ny_max=b.sr))

if "sweep" in what:
    for (c, r) in [(8, 8), (16, 16), (32, 32)]:
        b = PuBatch(c, r, 256, seed=4242, bit_depth=10, sr=32, n_start=1)
        dj, do, dr = up(b.gt_jobs), up(b.org), up(b.ref)
        keys = torch.empty(b.n, dtype=torch.int64, device=dev)
        cnt = torch.zeros(b.n, dtype=torch.int32, device=dev)
        out = torch.zeros(b.n * hop.GT_RES_DT.itemsize, dtype=torch.uint8, device=dev)

        def run():
            ctx.gt_sweep_keys_dev(b.n, dj.data_ptr(), do.data_ptr(), dr.data_ptr(), b.ref.size, c, r, 0, hop.HOP_SWEEP_CANDS,
                                  keys.data_ptr(), cnt.data_ptr(), ctx.stream)
            ctx.gt_sweep_finalize_dev(b.n, dj.data_ptr(), keys.data_ptr(), cnt.data_ptr(), out.data_ptr(), ctx.stream)
        timed("sweep %dx%d x%d" % (c, r, b.n), run)

if "tail" in what:
    # the encoder's call: one PU per host call against the resident mirror (k1_search + k_motion_tail)
    for (c, r) in [(8, 4), (8, 8), (16, 16), (32, 32)]:
        b = PuBatch(c, r, 1, seed=5, sr=64, n_start=3)
        ctx.ref_create(b.pw, b.ph, 0)
        ctx.ref_upload(np.ascontiguousarray(b.ref).reshape(b.ph, b.pw))
        mj = b.motion_jobs()
        timed("fused single-PU call %dx%d" % (c, r), lambda: ctx.motion_search(mj, b.org, None), iters=WARM + 2)

if "pred" in what:
    from hevc_hop_b200.workload import PredBatch
    pb = PredBatch([(8, 8), (16, 16), (32, 32), (64, 64), (16, 8), (32, 24)], 32, seed=11, kind="gt")
    timed("k6 predict x%d" % len(pb.jobs), lambda: ctx.predict(pb.jobs, pb.org, pb.ref, pb.dst_samples))

if "intra" in what:
    from hevc_hop_b200.workload import intra_jobs
    ij, io, ir = intra_jobs([4, 8, 16, 32, 64], 64, seed=12)
    timed("k7 intra x%d" % len(ij), lambda: ctx.intra_prescreen(ij, io, ir))

if "dist" in what:
    from hevc_hop_b200.workload import dist_jobs
    dj, dorg, dcur = dist_jobs(16, 16, 4096, func=hop.HOP_DF_HADS)
    timed("k3 dist x%d (16x16 HADs)" % len(dj), lambda: ctx.dist(dj, dorg, dcur))

if "ref" in what:
    # K4: the SS mirror of a 1024x1024 picture -- fill, then one 64x64 CU patch with its incremental border extension
    ctx.ref_create(1024, 1024, 80)
    blk = np.full((64, 64), 77, dtype=np.int16)
    def run():
        ctx.ref_reset(-1)
        ctx.ref_update(0, 0, blk)
        ctx.ref_update(960, 960, blk)
        ctx.sync()
    timed("k4 reset + 2 CU updates", run)
