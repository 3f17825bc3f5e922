"""gprof flat profile of the patched encoder's HOST code (run from the repo root on a GPU box).

Needs integration/_build/TAppEncoderHop_pg, the patched encoder linked with -pg (PC sampling covers every object
of the binary, call counts are not needed):
    OBJ=oracle/_ref/obj; LIBOBJS=$(find $OBJ/Lib -name '*.o' -not -path '*/TLibDecoder/*' | grep -v -E "TLibEncoder/(TEncSearch|TEncCu|TEncGOP)\\.o")
    g++ -pg -o integration/_build/TAppEncoderHop_pg $OBJ/App/TAppEncoder/*.o integration/_build/obj/*.o $LIBOBJS \\
        -Lhevc-hop_b200 -lhopgpu -lpthread -Wl,-rpath,'$ORIGIN/../../hevc-hop_b200'
then, with the gmon.out this script brings back in gpurun_out/:
    gprof -b -p integration/_build/TAppEncoderHop_pg gpurun_out/gmon.out
"""
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
