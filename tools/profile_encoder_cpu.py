import os, sys, shutil
sys.path.insert(0, 'tests'); import conftest
from hevc_hop_b200 import encoder
b = os.path.join(os.path.dirname(encoder.HOP_ENCODER), "TAppEncoderHop_pg")
r = encoder.encode(b, 512, 512, seed=100, device=0, keep=True)
print("seconds", r["seconds"], "dir", r["dir"], os.listdir(r["dir"]))
g = os.path.join(r["dir"], "gmon.out")
if os.path.exists(g): shutil.copy(g, "gpurun_out/gmon.out")
for line in r["log"].splitlines():
    if "Total Time" in line: print(line)
