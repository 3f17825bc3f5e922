"""Start-up costs of the library on a GPU box (run from the repo root): import, hop_ctx_create (CUDA context
creation dominates: 0.6-1.9 s depending on the box), mirror creation / upload, first and second fused search."""
import time, sys, os
t0=time.perf_counter()
sys.path.insert(0,'tests'); import conftest
import numpy as np
import hevc_hop_b200 as hop
from hevc_hop_b200.workload import PuBatch
t1=time.perf_counter()
ctx=hop.HopContext(0)
t2=time.perf_counter()
ctx.ref_create(1024,1024,80); t3=time.perf_counter()
ctx.ref_reset(); ctx.sync(); t4=time.perf_counter()
b = PuBatch(8, 8, 1, seed=5, sr=64, n_start=3)
ctx.ref_create(b.pw, b.ph, 0); ctx.ref_upload(np.ascontiguousarray(b.ref).reshape(b.ph, b.pw)); t5=time.perf_counter()
r=ctx.motion_search(b.motion_jobs(), b.org, None); t6=time.perf_counter()
r=ctx.motion_search(b.motion_jobs(), b.org, None); t7=time.perf_counter()
print("import %.3f  ctx_create %.3f  ref_create %.3f  ref_reset %.3f  upload %.3f  first search %.3f  second %.6f" % (t1-t0,t2-t1,t3-t2,t4-t3,t5-t4,t6-t5,t7-t6))
