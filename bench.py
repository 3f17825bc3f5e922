#!/usr/bin/env python
"""bench.py -- HOP candidate-search throughput on B200 (BASELINE.json metric, configs[2]).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--pus P]

One "step" = one pass of the HOP hot path (K2: xPatternSearchGT, all diamond passes, 1 start vector)
over one batch of synthetic PUs: P PUs for each of the shapes 8x8, 16x16, 32x32, 64x64 (P = 4096 by
default), inputs = original block + the 2W x 2H causal window of the start vector, threshold
MAX_UINT-1 so every candidate is warped and scored.  A HOP candidate = one affine corner set that is
warped (bilinear, IEEE binary64) and scored (Hadamard SATD + lambda*bits): 56 per pass.

  value      HOP candidates/s, whole job (all ranks), inputs already resident in HBM
  e2e        same metric through the host C-ABI call hop_pattern_search_gt_batch with HOST buffers:
             H2D copies of jobs/org/windows and the D2H read of the results are inside the timed region
  roofline   dominant kernel k2_gt_search against the fp64 pipe (the warp is binary64 by reference
             semantics) using a peak measured in this run by hop_probe_alu; the HBM view is reported
             beside it (`hbm`) because the contract asks for it, although the kernel is ALU bound
  cpu_baseline  the compiled reference (oracle/_ref/libhopref.so, kind "reference") or the C port of
             it (oracle/, kind "port") on a bounded sample of the same workload, host cores stated

N > 1 (torchrun): one process per GPU, each rank owns an independent batch (weak scaling, no data-path
collective: PUs / lenslet images are independent); NCCL is used for the barriers and the max-over-ranks
time only.  --impl reference times the reference's own CPU implementation on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

SHAPES = [(8, 8), (16, 16), (32, 32), (64, 64)]
# one string for both arms: the GPU arm runs the whole workload per step, the reference arm a bounded sample of it
WORKLOAD = ("configs[2]: standalone HOP candidate-search microbench, %d PUs of each of 8x8/16x16/32x32/64x64 "
            "per GPU, full diamond grid (3/4/5/6 passes x 56 affine candidates), 1 start vector, HadamardME, "
            "8-bit, QP32 lambda")
FP64_OPS_PER_PIXEL = 24      # Fx,Fy 8 + p,q 4 + bilinear 11 + rounding 1
INT_OPS_PER_PIXEL = 15       # Hadamard 9 + clamps 6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pus", type=int, default=4096, help="PUs per shape per GPU")
    ap.add_argument("--cpu-sample", type=int, default=512, help="PUs per shape in the CPU baseline sample (~16 s on one core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--encode-size", type=int, default=1024, help="edge of the square lenslet images of the encode leg; 0 = skip")
    ap.add_argument("--encode-procs", type=int, default=0, help="encoder processes per GPU (sharing it through MPS); 0 = host cores / 8, at least 1")
    ap.add_argument("--encode-images", type=int, default=0, help="images per GPU in the encode leg; 0 = one per encoder process")
    ap.add_argument("--sweep-pus", type=int, default=4096, help="PUs (Main10, 8x8/16x16/32x32/16x8 mix) of the sharded exhaustive sweep; 0 = skip")
    ap.add_argument("--k1-pus", type=int, default=2368, help="PUs per shape for the secondary K1 (SS full search) measurement (2368 = 16 per SM: 4 to 8 waves of resident CTAs); 0 = skip")
    return ap.parse_args()


def make_batches(hop, pus, seed):
    from hevc_hop_b200.workload import GtBatch
    from hevc_hop_b200.lenslet import lenslet_luma
    src = lenslet_luma(1024, 1024, seed=seed).astype(np.int16)
    return [GtBatch(c, r, pus, seed=seed * 16 + i, source=src) for i, (c, r) in enumerate(SHAPES)]


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for (t, line) in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                smax = float(f[2])
                if t0 <= t <= t1 + 0.1:
                    sm.append(float(f[1]))
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU baseline (the only place bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, shape_idx, n, seed = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    graft.load_package()
    import _oracle
    from hevc_hop_b200.workload import GtBatch
    chk = _oracle.ref() if kind == "reference" else _oracle.oracle()
    c, r = SHAPES[shape_idx]
    b = GtBatch(c, r, n, seed=seed)
    t = time.perf_counter()
    out = chk.pattern_search_gt(b.gt_jobs, b.org, b.ref)
    dt = time.perf_counter() - t
    return b.candidates(), dt, os.getpid(), int(out["cost"].astype(np.int64).sum() & 0xFFFFFFFF)


def cpu_baseline(sample_per_shape, procs):
    """Time the reference CPU path on `procs` processes; returns (candidates/s, kind, cores, sample text)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    graft.load_package()
    import _oracle
    import multiprocessing as mp
    kind = "reference" if _oracle.ref() is not None else "port"
    if kind == "port":
        _oracle.oracle()   # build once before forking
    # split every shape's sample over the workers so that all workers carry the same mix
    per = max(1, sample_per_shape // procs)
    tasks = [(kind, s, per, 1000 + 17 * w + s) for w in range(procs) for s in range(len(SHAPES))]
    if procs == 1:
        res = [_cpu_worker(t) for t in tasks]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_cpu_worker, tasks, chunksize=len(SHAPES))
    # search time only (input synthesis excluded); the processes run concurrently, one per core, so the
    # job's duration is the slowest process's summed search time
    per_pid = {}
    for (_, dt, pid, _) in res:
        per_pid[pid] = per_pid.get(pid, 0.0) + dt
    wall = max(per_pid.values())
    cands = sum(r[0] for r in res)
    sample = "%d PUs of each shape %s per process x %d processes, all diamond passes, 1 start vector" % (
        per, "/".join("%dx%d" % s for s in SHAPES), procs)
    return cands / wall, kind, procs, sample, wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    vals = []
    for i in range(args.warmup + args.steps):
        v, kind, c, sample, wall = cpu_baseline(8 * procs, procs)   # 8 PUs of each shape per process
        if i >= args.warmup:
            vals.append((v, wall))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([w for _, w in vals]) * 1e3)
    line = {
        "impl": "reference", "metric": "HOP candidates/s", "value": value, "unit": "candidates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+int32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD % args.pus},
        "cpu_baseline": {"value": value, "unit": "candidates/s", "cores": c, "kind": kind,
                         "sample": sample + " per step (a bounded sample of the workload: the CPU needs ~4 h per full step on one core)"},
        "e2e": {"value": value, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------
# lenslet encode s/image: the reference encoder with the drop-in patch, one process + context per GPU
# ---------------------------------------------------------------------------------------------------
def measure_encode(args, rank, world, local, dist, torch, dev):
    """Lenslet encode s/image (BASELINE configs[0] / configs[3] shape): every rank (= GPU) runs P encoder processes
    that share its GPU through MPS and encode `images` independent 1024x1024 frames; no collective on the data path."""
    import hashlib
    from hevc_hop_b200 import encoder, batch
    if args.encode_size <= 0 or not os.path.exists(encoder.HOP_ENCODER):
        return None
    n = args.encode_size
    cores = os.cpu_count() or 8
    procs = args.encode_procs or max(1, cores // 8)          # the 8-GPU box runs 8 ranks: never more processes than cores
    images = args.encode_images or procs
    mps = False
    if procs > 1:
        if local == 0:
            mps = batch.mps_start()
        if dist is not None:
            dist.barrier()
        mps = batch.mps_running()
        if not mps:
            procs = 1                                        # without MPS the driver time-slices whole contexts: no gain
    # rank 0's first image is the golden one (seed 0: its CPU bitstream is pinned in tests/golden/encode_golden.json)
    seeds = [0 if (rank == 0 and i == 0) else 100 + rank * images + i for i in range(images)]
    tasks = [dict(width=n, height=n, seed=sd) for sd in seeds]
    if dist is not None:
        dist.barrier()
    res, makespan = batch.encode_batch(tasks, device=local, procs=procs, use_mps=mps, stats=True)
    errors = [r["error"] for r in res if "error" in r]
    secs = [r["seconds"] for r in res if "error" not in r]
    # where an encoder process spends its wall clock (HOP_STATS lines of the shim): search calls on the GPU path, SS-mirror
    # updates, CUDA context creation; the rest is the reference's own host code (RDO, transforms, CABAC)
    split = np.zeros(3)
    startup = [r.get("worker_startup_s", 0.0) for r in res if "error" not in r]
    for r in res:
        if r.get("stats"):          # long-lived workers report per job; their CUDA start-up is `worker_startup_s`
            st = r["stats"]
            split += np.array([st["gpu_search_calls"], st["ss_mirror_updates"], st["context_and_mirror_create"]])
            continue
        for ln in (r.get("log") or "").splitlines():
            f = ln.split()
            if len(f) >= 5 and f[0] == "hopshim:" and f[3] == "calls":
                k = {"xPatternSearch": 0, "xPatternSearchGT": 0, "prefetch": 0, "refUpdate": 1, "hostBorder": 1, "refReset+create": 2}.get(f[1])
                if k is not None:
                    split[k] += float(f[4])
    # host probe: the same fixed single-threaded CPU job in as many processes as encoders ran, on all ranks at once --
    # how much slower does a host core get when N x procs of them are busy (all-core clocks, SMT siblings, memory)?
    if dist is not None:
        dist.barrier()
    probe_src = ("import time,numpy as np\nr=np.random.default_rng(1).random(1500000)\nt=time.perf_counter()\n"
                 "for _ in range(40): np.sort(r)\nprint(time.perf_counter()-t)")
    pp = [subprocess.Popen([sys.executable, "-c", probe_src], stdout=subprocess.PIPE, text=True) for _ in range(procs)]
    probe = [float(p_.communicate()[0].strip() or 0) for p_ in pp]
    t = torch.tensor([makespan, sum(secs), len(secs), max(secs) if secs else 0.0, len(errors)] + list(split) +
                     [sum(probe), len(probe), sum(startup), sum(1 for x in startup if x > 0)], dtype=torch.float64, device=dev)
    tmax = t.clone()
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.barrier()
    if mps and local == 0:
        batch.mps_stop()
    if rank != 0:
        return None
    total = int(t[2].item())
    ctus = ((n + 63) // 64) ** 2
    mk = float(tmax[0].item())
    enc = {"image": "%dx%d synthetic lenslet, HOP intra cfg, QP 32, 1 frame" % (n, n), "images": total,
           "gpus": world, "encoder_processes_per_gpu": procs, "mps": bool(mps), "host_cores": cores,
           "makespan_s": mk, "images_per_s": total / mk if mk > 0 else None,
           "s_per_image": float(t[1].item()) / max(1, total), "s_per_image_max": float(tmax[3].item()),
           "s_per_image_incl_worker_startup": (float(t[1].item()) + float(t[10].item())) / max(1, total),
           "ctus_per_image": ctus, "s_per_ctu": float(t[1].item()) / max(1, total) / ctus, "errors": int(t[4].item()),
           "workers": {"long_lived": bool(t[11].item() > 0), "count": int(t[11].item()),
                       "mean_startup_s": float(t[10].item()) / max(1.0, float(t[11].item())),
                       "what": "encoder processes that stay alive for their share of the queue (integration/hop_batch_main.cpp): "
                               "CUDA start-up once per worker, inside the makespan, not inside s_per_image"},
           "host_probe": {"seconds": float(t[8].item()) / max(1.0, float(t[9].item())), "processes": int(t[9].item()),
                          "what": "a fixed single-threaded numpy job run in as many concurrent processes as encoders, all ranks at once: "
                                  "its growth with the GPU count is the host's share of any per-image slowdown"},
           "s_per_image_split": {"gpu_search_calls": float(t[5].item()) / max(1, total), "ss_mirror_updates": float(t[6].item()) / max(1, total),
                                 "cuda_context_create": float(t[7].item()) / max(1, total),
                                 "host_rdo_and_rest": (float(t[1].item()) - float(t[5].item() + t[6].item() + t[7].item())) / max(1, total),
                                 "what": "mean over all images of all ranks, from the shim's HOP_STATS timers; host = the reference's own "
                                         "untouched code (RDO, transform/quant, CABAC, YUV I/O)"},
           "note": "s_per_image = mean wall clock of one image's encode over all images of all ranks (a process per image: incl. its "
                   "CUDA start-up; long-lived workers: start-up is paid once per worker and reported in `workers`, inside the "
                   "makespan); makespan = slowest rank; images are independent (no collective)"}
    gpath = os.path.join(ROOT, "tests", "golden", "encode_golden.json")
    if n == 1024 and os.path.exists(gpath) and "error" not in res[0]:
        g = json.load(open(gpath)).get("c0_1024x1024_qp32")
        if g:
            enc["bitstream_identical"] = bool(res[0]["md5"] == g["bitstream_md5"] and res[0]["bytes"] == g["bitstream_bytes"])
            enc["cpu_reference"] = {"s_per_image": g["cpu_seconds"], "s_per_ctu": g["cpu_seconds"] / ctus, "cores": 1,
                                    "measured": "unmodified reference encoder, same 1024x1024 frame (seed 0), build container CPU; "
                                                "md5 / size / seconds in tests/golden/encode_golden.json"}
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle
    if world == 1 and not args.no_cpu_baseline and os.path.exists(_oracle.REF_ENCODER):
        # bounded CPU sample on THIS box: the unmodified reference on a 256x256 image, the patched encoder on the same image
        try:
            ref = _oracle.encode_reference(256, 256, seed=7)
            hop = encoder.encode(encoder.HOP_ENCODER, 256, 256, seed=7, device=local)
            enc["cpu_reference_256x256"] = {"s_per_image": ref["seconds"], "s_per_ctu": ref["seconds"] / 16, "cores": 1,
                                            "gpu_s_per_image_same_input": hop["seconds"],
                                            "bitstream_identical": hashlib.md5(ref["bitstream"]).hexdigest() ==
                                            hashlib.md5(hop["bitstream"]).hexdigest()}
        except RuntimeError as e:     # the unmodified reference segfaults sporadically (tests/_oracle.py)
            enc["cpu_reference_256x256"] = {"error": str(e)[:200]}
    return enc


# ---------------------------------------------------------------------------------------------------
# exhaustive HOP parameter sweep (configs[4]): candidate range sharded over the ranks; the exchange is done by
# the sweep kernel itself (atomicMin into every rank's merge words over NVLink); NCCL all-reduce form beside it
# ---------------------------------------------------------------------------------------------------
SWEEP_MIX = [((8, 8), 0.5), ((16, 16), 0.375), ((32, 32), 0.109375), ((16, 8), 0.015625)]


def measure_sweep(hop, ctx, torch, dist, dev, tstream, pus, rank, world):
    if pus <= 0:
        return None
    from hevc_hop_b200 import sweep
    from hevc_hop_b200.workload import PuBatch
    up = lambda a: torch.from_numpy(a.view(np.uint8)).to(dev)
    groups = []
    for gi, ((c, r), share) in enumerate(SWEEP_MIX):
        n = max(1, int(round(pus * share)))
        b = PuBatch(c, r, n, seed=4242 + gi, bit_depth=10, sr=32, n_start=1)      # same inputs on every rank
        groups.append({"b": b, "jobs": up(b.gt_jobs), "org": up(b.org), "ref": up(b.ref),
                       "out": torch.zeros(n * hop.GT_RES_DT.itemsize, dtype=torch.uint8, device=dev)})
    total = sum(g["b"].n for g in groups)
    torch.cuda.synchronize()
    try:
        xch = sweep.SweepExchange(ctx, dist, max(g["b"].n for g in groups), rank, world)
        xerr = None
    except Exception as e:                      # no peer access between the GPUs: only the collective form runs
        xch, xerr = None, str(e)[:200]
    flags = torch.tensor([1 if xch is not None else 0], device=dev)
    if dist is not None:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if int(flags.item()) == 0:
        xch = None

    def run_peer():
        for g in groups:
            b = g["b"]
            xch.sweep(b.n, g["jobs"], g["org"], g["ref"], b.cols, b.rows, g["out"], ctx.stream)

    def run_coll():
        for g in groups:
            b = g["b"]
            sweep.sweep_on_device(ctx, torch, dist, b.n, g["jobs"], g["org"], g["ref"], b.cols, b.rows, g["out"], rank, world, ctx.stream)

    def timed(fn, iters=3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(iters + 1):
            if it == 1:
                torch.cuda.synchronize()
                if dist is not None:
                    dist.barrier()
                torch.cuda.synchronize()
                e0.record(tstream)
            fn()
        e1.record(tstream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle

    def check():
        ok = True
        for g in groups:
            b = g["b"]
            got = g["out"].cpu().numpy().view(hop.GT_RES_DT)[:2]
            want = _oracle.gt_sweep(b.gt_jobs[:2], b.org, b.ref)
            ok &= bool((got["cost"] == want["cost"]).all() and got["gt"].tobytes() == want["gt"].tobytes() and
                       (got["best_index"] == want["best_index"]).all() and (got["n_candidates"] == want["n_candidates"]).all())
        return ok

    ms_peer = ok_peer = None
    if xch is not None:
        ms_peer = timed(run_peer)
        ok_peer = check() if rank == 0 else None
    ms_coll = timed(run_coll)
    ok_coll = check() if rank == 0 else None
    if rank != 0:
        return None
    ms = ms_peer if ms_peer is not None else ms_coll
    return {"workload": "%d PUs Main10 (%s), reference mode IT_GT_SEARCH 1 (N=2): %d affine corner sets per PU, candidate range "
                        "sharded over %d GPU(s), one launch pair per shape" % (
                            total, ", ".join("%d x %dx%d" % (g["b"].n, g["b"].cols, g["b"].rows) for g in groups), hop.HOP_SWEEP_CANDS, world),
            "scaling": "strong", "ms": ms, "candidates_per_s": total * hop.HOP_SWEEP_CANDS / (ms * 1e-3),
            "exchange": ("in-kernel atomicMin into every rank's merge words over peer memory (NVLink), arrival counters; no collective call"
                         if ms_peer is not None else "collective (peer exchange unavailable: %s)" % xerr),
            "parity_spot_check": ok_peer if ms_peer is not None else ok_coll,
            "collective_form": {"ms": ms_coll, "candidates_per_s": total * hop.HOP_SWEEP_CANDS / (ms_coll * 1e-3),
                                "what": "same kernels + all_reduce(MIN) on keys and all_reduce(SUM) on counts (NCCL)" if world > 1 else "single rank, no exchange",
                                "parity_spot_check": ok_coll}}


# ---------------------------------------------------------------------------------------------------
# secondary: K1 full search (SAD positions/s), reported beside the headline
# ---------------------------------------------------------------------------------------------------
def measure_k1(hop, ctx, torch, dev, tstream, pus, peaks):
    from hevc_hop_b200.workload import PuBatch
    from hevc_hop_b200.lenslet import lenslet_luma
    src = lenslet_luma(1024, 1024, seed=99).astype(np.int16)
    res = {"unit": "SAD positions/s", "per_shape": {}}
    tot_pos, tot_px, tot_ms = 0, 0, 0.0
    for i, (c, r) in enumerate(SHAPES):
        b = PuBatch(c, r, pus, seed=300 + i, source=src)
        d_jobs = torch.from_numpy(b.search_jobs.view(np.uint8)).to(dev)
        d_org = torch.from_numpy(b.org.view(np.uint8)).to(dev)
        d_ref = torch.from_numpy(b.ref.view(np.uint8)).to(dev)
        d_out = torch.zeros(b.n * hop.SEARCH_RES_DT.itemsize, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(3):
            if it == 1:
                e0.record(tstream)
            ctx.pattern_search_dev(b.n, d_jobs.data_ptr(), d_org.data_ptr(), d_ref.data_ptr(), d_out.data_ptr(), ctx.stream,
                                   cols=c, rows=r, nx_max=2 * b.sr + 1, ny_max=b.sr)
        e1.record(tstream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        # parity spot check of what was just timed, at its own size (SearchRange 128): first PUs against the oracle
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import _oracle
        got = d_out.cpu().numpy().view(hop.SEARCH_RES_DT)[:2]
        res["parity_spot_check"] = bool(res.get("parity_spot_check", True) and
                                        got.tobytes() == _oracle.oracle().pattern_search(b.search_jobs[:2], b.org, b.ref).tobytes())
        j = b.search_jobs[0]
        npos = int((j["rng_right"] - j["rng_left"] + 1) * (j["rng_bottom"] - j["rng_top"] + 1)) * b.n
        px = npos * c * (r // 2 if r > 8 else r)
        res["per_shape"]["%dx%d" % (c, r)] = {"ms": ms, "positions_per_s": npos / (ms * 1e-3), "pixel_sads_per_s": px / (ms * 1e-3)}
        tot_pos += npos; tot_px += px; tot_ms += ms
    res["value"] = tot_pos / (tot_ms * 1e-3)
    res["pixel_sads_per_s"] = tot_px / (tot_ms * 1e-3)
    res["vabsdiff4_peak_pixel_sads_per_s"] = peaks.get("vabsdiff4", 0) * 4e9
    res["frac_of_vabsdiff4_peak"] = res["pixel_sads_per_s"] / max(1.0, res["vabsdiff4_peak_pixel_sads_per_s"])
    res["workload"] = "%d PUs of each of 8x8/16x16/32x32/64x64, SearchRange 128 causal window (257x125 positions), FEN row sub-sampling" % pus
    return res


def bind_to_gpu_numa_node(torch, local):
    """Several ranks on one host: run this rank -- and with it the first touch of its pinned host buffers -- on the CPUs
    of the NUMA node its GPU hangs off, so that the end-to-end copies do not cross the socket interconnect.  Returns a
    short description, or None when the topology is not visible (then nothing is changed).  HOP_BENCH_NUMA=0: off."""
    if os.environ.get("HOP_BENCH_NUMA", "1") == "0":
        return None
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return "gpu %s: no NUMA node reported" % bdf
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "gpu %s: node %d has no CPU this process may use" % (bdf, node)
        os.sched_setaffinity(0, cpus)
        return "gpu %s on NUMA node %d: rank bound to %d CPUs" % (bdf, node, len(cpus))
    except Exception as e:
        return "topology not visible (%s)" % str(e)[:80]


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libhopgpu has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cpus_before = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(torch, local) if world > 1 else None
    hop = graft.load_package()
    ctx = hop.HopContext(local)
    batches = make_batches(hop, args.pus, seed=rank + 1)

    def to_dev(a):
        return torch.from_numpy(a.view(np.uint8)).to(dev)

    def pinned(a):
        t = torch.from_numpy(a.view(np.uint8)).pin_memory()
        return t

    dbat = []
    for b in batches:
        d = {"b": b, "jobs": to_dev(b.gt_jobs), "org": to_dev(b.org), "ref": to_dev(b.ref),
             "out": torch.zeros(b.n * hop.GT_RES_DT.itemsize, dtype=torch.uint8, device=dev),
             "h_jobs": pinned(b.gt_jobs), "h_org": pinned(b.org), "h_ref": pinned(b.ref),
             "h_out": torch.zeros(b.n * hop.GT_RES_DT.itemsize, dtype=torch.uint8).pin_memory()}
        dbat.append(d)
    # all library work and all timing events live on the context's own stream (torch's default stream
    # has handle 0, which the ABI reads as "use the context stream")
    tstream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    torch.cuda.set_stream(tstream)
    stream = ctx.stream
    cands_per_step = sum(b.candidates() for b in batches)
    pix_per_step = sum(b.pixel_candidates() for b in batches)
    in_bytes = sum(b.input_bytes() for b in batches)
    out_bytes = sum(b.n * hop.GT_RES_DT.itemsize for b in batches)

    def step_resident():
        for d in dbat:
            b = d["b"]
            ctx.pattern_search_gt_dev(b.n, d["jobs"].data_ptr(), d["org"].data_ptr(), d["ref"].data_ptr(), b.ref.size,
                                      d["out"].data_ptr(), b.cols, b.rows, stream)

    # host calls of one end-to-end step: one hop_pattern_search_gt_batch_async call per shape, except that a
    # shape with more than 16 MB of input (32x32, 64x64) goes in chunks of 1184 PUs (8 waves of the 148 SMs for the
    # 64x64 class; each chunk a complete call on its own slice of the pinned host buffers, offsets rebased), so
    # that its copy pipelines with its own kernels instead of preceding them.  Small shapes first: their copies
    # are short and their kernels cover the copy of the next, larger call.
    e2e_calls = []
    for d in dbat:
        b = d["b"]
        org_per, ref_per = b.org.size // b.n, b.ref.size // b.n
        chunk = 1184 if 2 * b.n * (org_per + ref_per) > (16 << 20) else b.n
        for k0 in range(0, b.n, chunk):
            k1 = min(b.n, k0 + chunk)
            jobs = b.gt_jobs[k0:k1].copy()
            jobs["org_off"] -= k0 * org_per
            jobs["ref_off"] -= k0 * ref_per
            assert jobs["org_off"].min() >= 0 and jobs["ref_off"].min() >= 0
            hj = pinned(jobs)
            e2e_calls.append((k1 - k0, hj, d["h_org"].data_ptr() + 2 * k0 * org_per, (k1 - k0) * org_per,
                              d["h_ref"].data_ptr() + 2 * k0 * ref_per, (k1 - k0) * ref_per,
                              d["h_out"].data_ptr() + k0 * hop.GT_RES_DT.itemsize))

    os.sched_setaffinity(0, cpus_before)      # the pinned buffers exist (first touch done): the later legs use every core again

    def step_e2e():
        # the reference-facing host call: pinned HOST buffers in, results back on the host; the calls of a step
        # are streamed through the asynchronous form (the input copy of call k+1 overlaps the kernel of call
        # k; a ring of HOP_ASYNC_SLOTS device slots), then one sync -- all inside the timed region
        for (n, hj, p_org, n_org, p_ref, n_ref, p_out) in e2e_calls:
            ctx._check(ctx.lib.hop_pattern_search_gt_batch_async(ctx.h, n, hj.data_ptr(), p_org, n_org, p_ref, n_ref, p_out))
        ctx.sync()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ALU peaks for the roofline, measured in this run (burst, kernel alone)
    peaks = {}
    if rank == 0:
        for what, name in ((2, "fp64_add"), (3, "fp64_mul"), (0, "int32_add"), (5, "int32_lop3"), (1, "vabsdiff4")):
            peaks[name] = max(ctx.probe_alu(what)[0] for _ in range(2))

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = ctx.launch_count
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(dbat) + 1)] for _ in range(args.steps)]
    barrier()
    t_wall0 = time.time()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(tstream)
    for s in range(args.steps):
        for i, d in enumerate(dbat):
            b = d["b"]
            ev[s][i].record(tstream)
            ctx.pattern_search_gt_dev(b.n, d["jobs"].data_ptr(), d["org"].data_ptr(), d["ref"].data_ptr(), b.ref.size,
                                      d["out"].data_ptr(), b.cols, b.rows, stream)
        ev[s][len(dbat)].record(tstream)
    e1.record(tstream)
    barrier()
    t_wall1 = time.time()
    launches = ctx.launch_count - launches0
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = cands_per_step * world / (ms_per_step * 1e-3)
    # per-shape kernel durations (CUDA events on the launching stream)
    per_shape_ms = [float(np.mean([ev[s][i].elapsed_time(ev[s][i + 1]) for s in range(args.steps)])) for i in range(len(dbat))]

    # parity spot check of what was just timed (outside the timed region): first PUs of every shape
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    parity = None
    if rank == 0:
        import _oracle
        orc = _oracle.oracle()
        parity = True
        for d in dbat:
            b = d["b"]
            got = d["out"].cpu().numpy().view(hop.GT_RES_DT)[:2]
            want = orc.pattern_search_gt(b.gt_jobs[:2], b.org, b.ref)
            parity &= got.tobytes() == want.tobytes()

    # end-to-end through the host C-ABI call (warm-up: every slot of the ring has seen the largest chunk)
    for _ in range(5):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    e2e_same = all(bool((d["h_out"].numpy() == d["out"].cpu().numpy()).all()) for d in dbat)   # host results == resident results
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = cands_per_step * world / (e2e_ms * 1e-3)

    # whole-image encode through the patched reference encoder: one image per GPU (replicas, no collective)
    enc = measure_encode(args, rank, world, local, dist if world > 1 else None, torch, dev)

    # exhaustive sweep sharded over the ranks: the one path with a real exchange (NCCL all-reduce-min)
    swp = measure_sweep(hop, ctx, torch, dist if world > 1 else None, dev, tstream, args.sweep_pus, rank, world)

    # secondary measurement (rank 0, outside the headline region): K1 = xPatternSearch, SearchRange 128
    k1 = None
    if rank == 0 and args.k1_pus > 0:
        k1 = measure_k1(hop, ctx, torch, dev, tstream, args.k1_pus, peaks)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel (k2_gt_search, all four launches of a step)
    kern_ms = sum(per_shape_ms)
    fp64_ops = FP64_OPS_PER_PIXEL * pix_per_step
    int_ops = INT_OPS_PER_PIXEL * pix_per_step
    fp64_peak = max(peaks["fp64_add"], peaks["fp64_mul"])
    achieved = fp64_ops / (kern_ms * 1e-3) / 1e9
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, hbm_src = float(mp["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    hbm_achieved = (in_bytes + out_bytes) / (kern_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k2_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_step")
        except Exception:
            traffic = None
    pipes = None
    pp = os.path.join(ROOT, "profiles", "k2_pipes.json")
    if os.path.exists(pp):
        try:
            pipes = json.load(open(pp))
        except Exception:
            pipes = None
    roofline = {
        "kernel": "k2_gt_search", "bound": "fp64-alu", "achieved": achieved, "peak": fp64_peak, "unit": "Gop/s",
        "frac": achieved / fp64_peak, "traffic": traffic,
        # what the pipes really do (ncu capture of the same kernels, profiles/): `frac` counts the REFERENCE's 24 fp64
        # operations per warped pixel, the kernel executes about half as many
        "fp64_pipe_busy": pipes.get("fp64_pipe_busy") if pipes else None,
        "pipes_ncu": pipes,
        "peak_source": "hop_probe_alu DADD/DMUL loop, measured in this run (MEASURED_PEAKS.json has no fp64/int32 figure)",
        "algorithmic": {"fp64_ops_per_pixel_candidate": FP64_OPS_PER_PIXEL, "int32_ops_per_pixel_candidate": INT_OPS_PER_PIXEL,
                        "pixel_candidates_per_step": pix_per_step},
        "int32": {"achieved": int_ops / (kern_ms * 1e-3) / 1e9, "peak": peaks["int32_lop3"], "unit": "Gop/s",
                  "frac": int_ops / (kern_ms * 1e-3) / 1e9 / peaks["int32_lop3"], "peak_int32_add": peaks["int32_add"]},
        "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                "algorithmic_bytes_per_step": in_bytes + out_bytes, "peak_source": hbm_src},
        "kernel_ms_per_step": kern_ms, "per_shape_ms": dict(zip(["%dx%d" % s for s in SHAPES], per_shape_ms)),
    }

    cpu = None
    if not args.no_cpu_baseline:
        v, kind, cores, sample, wall = cpu_baseline(args.cpu_sample, 1)
        cpu = {"value": v, "unit": "candidates/s", "cores": cores, "kind": kind, "sample": sample, "seconds": wall}

    line = {
        "metric": "HOP candidates/s", "value": value, "unit": "candidates/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64+int32", "data": "synthetic",
        "config": {"workload": WORKLOAD % args.pus,
                   "l2": "inputs larger than L2 (%.0f MB per step per GPU)" % (in_bytes / 1e6),
                   "candidates_per_step_per_gpu": cands_per_step},
        "e2e": {"value": e2e_value, "unit": "candidates/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                "ms_per_step": e2e_ms, "api": "hop_pattern_search_gt_batch_async x%d (32x32 and 64x64 shapes in chunks of 1184 PUs) + hop_ctx_sync (host buffers, pinned)" % len(e2e_calls)},
        "e2e_results_equal_resident": e2e_same,
        "numa_binding_rank0": numa,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "parity_spot_check": parity,
        "k1_sad_search": k1,
        "encode": enc,
        "sweep": swp,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of the contract, on the process's real stdout."""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())
    else:
        print(json.dumps(line))


def main():
    global _REAL_STDOUT
    args = parse()
    # stdout carries exactly one JSON line: whatever libraries print while the bench runs (torch prints an
    # "NCCL version" banner at the first collective) is sent to stderr at the file-descriptor level
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
