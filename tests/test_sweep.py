"""Exhaustive HOP parameter sweep (configs[4]): oracle vs the reference compiled in its exhaustive mode,
the sharding / all-reduce-min host logic on 2 gloo ranks (CPU), and the CUDA path (-m gpu)."""
import os

import numpy as np
import pytest

import conftest  # noqa: F401  (registers hevc_hop_b200, also inside spawned ranks)
import _oracle
import hevc_hop_b200 as hop
from hevc_hop_b200 import sweep
from hevc_hop_b200.workload import PuBatch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def same(a, b, extras=False):
    ok = (a["gt_flag"] == b["gt_flag"]).all() and (a["cost"] == b["cost"]).all() and a["gt"].tobytes() == b["gt"].tobytes()
    if extras:
        ok &= (a["best_index"] == b["best_index"]).all() and (a["n_candidates"] == b["n_candidates"]).all()
    return bool(ok)


def test_oracle_sweep_matches_golden():
    g = np.load(os.path.join(GOLD, "sweep_golden.npz"))
    for k in range(int(g["n_sweep"])):
        t = "w%02d" % k
        assert same(_oracle.gt_sweep(g[t + "_jobs"], g[t + "_org"], g[t + "_ref"]), g[t + "_out"]), k


@pytest.mark.skipif(_oracle.ref_sweep() is None, reason="oracle/_ref/libhopref_sweep.so not built")
@pytest.mark.parametrize("shape", [(8, 8), (16, 8), (8, 4), (16, 12), (24, 32)])
def test_oracle_sweep_vs_compiled_reference(shape):
    c, r = shape
    b = PuBatch(c, r, 3, seed=c * 5 + r, sr=20, n_start=1)
    full = _oracle.gt_sweep(b.gt_jobs, b.org, b.ref)
    assert same(full, _oracle.ref_sweep().pattern_search_gt(b.gt_jobs, b.org, b.ref))
    j = b.gt_jobs.copy(); j["threshold"] = full["cost"]          # ties with the threshold are not accepted
    assert same(_oracle.gt_sweep(j, b.org, b.ref), _oracle.ref_sweep().pattern_search_gt(j, b.org, b.ref))


def test_shard_range_partitions_the_candidates():
    for world in (1, 2, 3, 4, 7, 8):
        edges = [sweep.shard_range(hop.HOP_SWEEP_CANDS, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == hop.HOP_SWEEP_CANDS
        assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
        sizes = [e - b for b, e in edges]
        assert max(sizes) - min(sizes) <= 1


def test_key_int64_round_trip_keeps_order():
    k = np.array([0xFFFFFFFFFFFFFFFF, 5 << 32 | 7, 5 << 32 | 6, 0x7FFFFFFF << 32], dtype=np.uint64)
    i = sweep.keys_to_int64(k)
    assert i[0] == sweep.NONE_KEY and i.argmin() == 2
    assert (sweep.int64_to_keys(i) == k).all()


def _gloo_rank(rank, world, port, result):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = PuBatch(8, 8, 4, seed=77, sr=16, n_start=1)
    begin, end = sweep.shard_range(hop.HOP_SWEEP_CANDS, rank, world)
    keys = sweep.keys_to_int64(_oracle.gt_sweep_keys(b.gt_jobs, b.org, b.ref, begin, end))
    t = torch.from_numpy(keys)
    sweep.allreduce_min_keys(t, dist)
    out = _oracle.gt_sweep_finalize(b.gt_jobs, sweep.int64_to_keys(t.numpy()))
    full = _oracle.gt_sweep(b.gt_jobs, b.org, b.ref)
    ok = same(out, full) and (out["best_index"] == full["best_index"]).all()
    r = torch.tensor([1 if ok else 0])
    dist.all_reduce(r, op=dist.ReduceOp.MIN)
    if rank == 0:
        result.put(int(r.item()))
    dist.destroy_process_group()


def test_two_rank_allreduce_min_reproduces_serial_sweep():
    """world_size 2 over gloo: each rank scores its slice of the candidates, one MIN all-reduce of the
    (cost << 32 | loop index) words, identical finalisation on every rank == the unsharded sweep."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_rank, args=(r, 2, 29631, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) == 1


# ---- CUDA path ---------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(8, 8), (16, 16), (8, 4), (4, 8), (16, 12), (32, 8), (32, 32), (64, 64)])
def test_gpu_sweep_matches_oracle(ctx, shape):
    c, r = shape
    n = 1 if c * r >= 1024 else 3
    for use_had in (1, 0):
        b = PuBatch(c, r, n, seed=c * 9 + r + use_had, sr=20, use_had=use_had, n_start=1)
        want = _oracle.gt_sweep(b.gt_jobs, b.org, b.ref)
        assert same(ctx.gt_sweep(b.gt_jobs, b.org, b.ref), want, extras=True)
        j = b.gt_jobs.copy(); j["threshold"] = want["cost"]
        assert same(ctx.gt_sweep(j, b.org, b.ref), _oracle.gt_sweep(j, b.org, b.ref), extras=True)


@pytest.mark.gpu
def test_gpu_sweep_golden_and_main10(ctx):
    g = np.load(os.path.join(GOLD, "sweep_golden.npz"))
    for k in range(int(g["n_sweep"])):
        t = "w%02d" % k
        assert same(ctx.gt_sweep(g[t + "_jobs"], g[t + "_org"], g[t + "_ref"]), g[t + "_out"]), k


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3, 8])
def test_gpu_sharded_keys_min_equals_full_sweep(ctx, world):
    """One GPU emulating `world` ranks: per-slice key kernels, MIN over slices, finalize == full sweep."""
    import torch
    b = PuBatch(16, 16, 5, seed=5, sr=20, n_start=1)
    dev = torch.device("cuda", 0)
    up = lambda a: torch.from_numpy(a.view(np.uint8).copy()).to(dev)
    d_jobs, d_org, d_ref = up(b.gt_jobs), up(b.org), up(b.ref)
    d_out = torch.zeros(b.n * hop.GT_RES_DT.itemsize, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    merged = None
    counts = torch.zeros(b.n, dtype=torch.int32, device=dev)
    for rank in range(world):
        begin, end = sweep.shard_range(hop.HOP_SWEEP_CANDS, rank, world)
        keys = torch.empty(b.n, dtype=torch.int64, device=dev)
        cnt = torch.zeros(b.n, dtype=torch.int32, device=dev)
        ctx.gt_sweep_keys_dev(b.n, d_jobs.data_ptr(), d_org.data_ptr(), d_ref.data_ptr(), b.ref.size, 16, 16, begin, end,
                              keys.data_ptr(), cnt.data_ptr(), ctx.stream)
        ctx.sync()
        keys = torch.where(keys < 0, torch.full_like(keys, int(sweep.NONE_KEY)), keys)
        merged = keys if merged is None else torch.minimum(merged, keys)
        counts += cnt
    merged = torch.where(merged == int(sweep.NONE_KEY), torch.full_like(merged, -1), merged)
    torch.cuda.synchronize()
    ctx.gt_sweep_finalize_dev(b.n, d_jobs.data_ptr(), merged.data_ptr(), counts.data_ptr(), d_out.data_ptr(), ctx.stream)
    ctx.sync()
    got = d_out.cpu().numpy().view(hop.GT_RES_DT)
    assert same(got, _oracle.gt_sweep(b.gt_jobs, b.org, b.ref), extras=True)


# ---- sharded sweep over peer memory: one process per GPU, no collective on the data path -----------------------
def _peer_rank(rank, world, port, result):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)      # carries the 64-byte handles and barriers only
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    cx = hop.HopContext(rank)
    xch = sweep.SweepExchange(cx, dist, 64, rank, world)
    ok = True
    up = lambda a: torch.from_numpy(a.view(np.uint8).copy()).to(dev)
    # different shapes, bit depths and batch sizes back to back (the merge words alternate by sweep parity)
    cases = [(8, 8, 5, 8), (16, 16, 7, 10), (16, 12, 3, 8), (32, 32, 2, 8), (8, 4, 64, 8), (16, 16, 1, 8), (4, 8, 9, 10)]
    for it, (c, r, n, bd) in enumerate(cases * 2):
        b = PuBatch(c, r, n, seed=900 + it, bit_depth=bd, sr=20, n_start=1, use_had=it % 3 != 0)
        if it % 5 == 4:
            b.gt_jobs["threshold"] = 0           # nothing can be accepted: flag 0, counts still exchanged
        d_jobs, d_org, d_ref = up(b.gt_jobs), up(b.org), up(b.ref)
        d_out = torch.zeros(n * hop.GT_RES_DT.itemsize, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        xch.sweep(n, d_jobs, d_org, d_ref, c, r, d_out, cx.stream)
        cx.sync()
        got = d_out.cpu().numpy().view(hop.GT_RES_DT)
        ok &= same(got, _oracle.gt_sweep(b.gt_jobs, b.org, b.ref), extras=True)
    r = torch.tensor([1 if ok else 0])
    dist.all_reduce(r, op=dist.ReduceOp.MIN)
    if rank == 0:
        result.put(int(r.item()))
    dist.barrier()
    cx.close()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_gpu_peer_memory_sweep_equals_serial_sweep():
    """One process per GPU: atomicMin of the partial keys into every rank's merge words over NVLink, arrival
    counters instead of a collective -- results equal the unsharded oracle sweep on every rank."""
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs two GPUs with peer access")
    import torch.multiprocessing as mp
    mctx = mp.get_context("spawn")
    q = mctx.Queue()
    procs = [mctx.Process(target=_peer_rank, args=(r, world, 29641, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=5) == 1


@pytest.mark.gpu
def test_gpu_peer_memory_sweep_single_rank(ctx):
    """world = 1 goes through the same two kernels (the rank pushes into its own merge words)."""
    import torch
    dev = torch.device("cuda", 0)
    xch = sweep.SweepExchange(ctx, None, 16, 0, 1)
    up = lambda a: torch.from_numpy(a.view(np.uint8).copy()).to(dev)
    for it, (c, r, n) in enumerate([(8, 8, 4), (16, 8, 16), (32, 16, 2), (8, 8, 1)]):
        b = PuBatch(c, r, n, seed=40 + it, sr=20, n_start=1)
        d_jobs, d_org, d_ref = up(b.gt_jobs), up(b.org), up(b.ref)
        d_out = torch.zeros(n * hop.GT_RES_DT.itemsize, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        xch.sweep(n, d_jobs, d_org, d_ref, c, r, d_out, ctx.stream)
        ctx.sync()
        assert same(d_out.cpu().numpy().view(hop.GT_RES_DT), _oracle.gt_sweep(b.gt_jobs, b.org, b.ref), extras=True), it
