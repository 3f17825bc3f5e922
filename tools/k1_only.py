"""Run only the batched K1 search for one PU shape (ncu target):  python tools/k1_only.py cols rows [pus] [iters]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import conftest  # noqa: F401
import numpy as np, torch
import hevc_hop_b200 as hop
from hevc_hop_b200.workload import PuBatch
from hevc_hop_b200.lenslet import lenslet_luma

c, r = int(sys.argv[1]), int(sys.argv[2])
pus = int(sys.argv[3]) if len(sys.argv) > 3 else 592
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ctx = hop.HopContext(0)
dev = torch.device("cuda", 0)
src = lenslet_luma(1024, 1024, seed=99).astype(np.int16)
b = PuBatch(c, r, pus, seed=300, source=src)
up = lambda a: torch.from_numpy(a.view(np.uint8)).to(dev)
d_jobs, d_org, d_ref = up(b.search_jobs), up(b.org), up(b.ref)
d_out = torch.zeros(b.n * hop.SEARCH_RES_DT.itemsize, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
ts = torch.cuda.ExternalStream(ctx.stream, device=dev)
e = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
for i in range(iters):
    e[i].record(ts)
    ctx.pattern_search_dev(b.n, d_jobs.data_ptr(), d_org.data_ptr(), d_ref.data_ptr(), d_out.data_ptr(), ctx.stream,
                           cols=c, rows=r, nx_max=2 * b.sr + 1, ny_max=b.sr)
e[iters].record(ts)
torch.cuda.synchronize()
print("%dx%d x%d: ms per launch" % (c, r, pus), [round(e[i].elapsed_time(e[i + 1]), 4) for i in range(iters)])
