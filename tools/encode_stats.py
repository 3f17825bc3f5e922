"""Encode one synthetic lenslet image with the patched reference encoder and print the shim's per-shape call
statistics (HOP_STATS=1):  python tools/encode_stats.py [size] [seed]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import conftest  # noqa: F401
from hevc_hop_b200 import encoder

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 100
os.environ["HOP_STATS"] = "1"
r = encoder.encode(encoder.HOP_ENCODER, n, n, seed=seed, device=0)
print("encode %dx%d: %.2f s, %d bytes" % (n, n, r["seconds"], len(r["bitstream"])))
for line in (r.get("log") or "").splitlines():
    if "hopshim" in line or "Total Time" in line:
        print(line)
