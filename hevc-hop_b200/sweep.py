"""Sharded exhaustive HOP parameter sweep (BASELINE.json configs[4]).

The HOP_SWEEP_CANDS affine corner sets of the reference's IT_GT_SEARCH 1 mode are independent, so the
flattened candidate range is cut into `world` contiguous slices, one per GPU.  Every rank evaluates its
slice for ALL PUs of the batch and produces one 64-bit key per PU (cost << 32 | flat loop index); ONE
all-reduce with MIN over those words reproduces the serial "first strict minimum in loop order", after
which every rank finalises identical results.  The message is 8 B x #PUs: latency bound, so many PUs
are batched per all-reduce.

`torch.distributed` is the plumbing (NCCL over NVLink on the GPU box, gloo in the CPU tests); the
kernels are libhopgpu's.  Keys travel as int64 (costs stay far below 2^31, so the order is unchanged).
"""
import numpy as np

from . import HOP_SWEEP_CANDS

NONE_KEY = np.int64(0x7FFFFFFFFFFFFFFF)


def shard_range(total, rank, world):
    """Contiguous slice [begin, end) of `total` candidates owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def keys_to_int64(keys_u64):
    """uint64 keys (all-ones = nothing scored) -> int64 with the same ordering for MIN."""
    k = np.asarray(keys_u64, dtype=np.uint64)
    out = k.astype(np.int64)          # all-ones wraps to -1: map it to the largest int64 instead
    out[k == np.uint64(0xFFFFFFFFFFFFFFFF)] = NONE_KEY
    return out


def int64_to_keys(keys_i64):
    k = np.asarray(keys_i64, dtype=np.int64).astype(np.uint64)
    k[np.asarray(keys_i64) == NONE_KEY] = np.uint64(0xFFFFFFFFFFFFFFFF)
    return k


def allreduce_min_keys(t, dist=None, group=None):
    """In-place MIN all-reduce of an int64 key tensor (no-op without a process group)."""
    if dist is not None and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return t


def sweep_on_device(ctx, torch, dist, n, d_jobs, d_org, d_ref, max_cols, max_rows, d_out, rank=0, world=1, stream=None):
    """One sharded sweep over device-resident inputs (torch uint8 tensors) with the COLLECTIVE exchange
    (all_reduce MIN on the keys, SUM on the counts: NCCL on GPUs, gloo in the CPU tests); returns the key tensor.
    This is the baseline form; `SweepExchange` is the product path on an NVLink node.

    The kernels run on the context's stream; the torch ops and the collectives in between are issued on the
    same stream (ExternalStream), so the order is the program order whatever torch's current stream is."""
    dev = d_jobs.device
    begin, end = shard_range(HOP_SWEEP_CANDS, rank, world)
    ext = torch.cuda.ExternalStream(stream or ctx.stream, device=dev)
    with torch.cuda.stream(ext):
        keys = torch.empty(n, dtype=torch.int64, device=dev)
        counts = torch.zeros(n, dtype=torch.int32, device=dev)
        ctx.gt_sweep_keys_dev(n, d_jobs.data_ptr(), d_org.data_ptr(), d_ref.data_ptr(), d_ref.numel() // 2, max_cols, max_rows,
                              begin, end, keys.data_ptr(), counts.data_ptr(), ext.cuda_stream)
        if world > 1:
            # all-ones (nothing scored) is -1 as int64: lift it above every real key before the MIN
            keys = torch.where(keys < 0, torch.full_like(keys, int(NONE_KEY)), keys)
            allreduce_min_keys(keys, dist)
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)
            keys = torch.where(keys == int(NONE_KEY), torch.full_like(keys, -1), keys)
        ctx.gt_sweep_finalize_dev(n, d_jobs.data_ptr(), keys.data_ptr(), counts.data_ptr(), d_out.data_ptr(), ext.cuda_stream)
    return keys


class SweepExchange:
    """Sharded sweep WITHOUT a collective call: every rank exports one merge word per PU (CUDA IPC), the sweep
    kernel min-reduces its partial keys into all ranks' words through peer-mapped memory over NVLink while it is
    still computing, a second kernel waits for every rank's arrival counter and finalises (include/hop_gpu.h,
    hop_gt_sweep_sharded_dev).  torch.distributed only carries the 64-byte handles once, at set-up."""

    def __init__(self, ctx, dist, max_pus, rank, world):
        self.ctx, self.rank, self.world = ctx, rank, world
        mine = ctx.sweep_exchange_create(max_pus)
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, mine)      # also the barrier behind which every rank's words are initialised
        else:
            handles[0] = mine
        ctx.sweep_exchange_connect(world, rank, handles)
        if world > 1:
            dist.barrier()                             # nobody pushes before everybody has mapped everybody

    def sweep(self, n, d_jobs, d_org, d_ref, max_cols, max_rows, d_out, stream=None):
        """Two kernel launches on the context's stream; d_out holds the finalised results when the stream is done."""
        self.ctx.gt_sweep_sharded_dev(n, d_jobs.data_ptr(), d_org.data_ptr(), d_ref.data_ptr(), d_ref.numel() // 2,
                                      max_cols, max_rows, d_out.data_ptr(), stream)
