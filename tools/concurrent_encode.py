"""K encoder processes at once on ONE GPU (different seeds): how many host-bound encoders can a B200 serve?
    python tools/concurrent_encode.py [size] [k1,k2,...] [--pin]
Prints per-process seconds, makespan and images/s for every K (run from the repo root on a GPU box)."""
import os, sys, time, threading
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import conftest  # noqa: F401
from hevc_hop_b200 import encoder

size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4, 8]
pin = "--pin" in sys.argv
cores = os.cpu_count()
for k in ks:
    res = [None] * k

    def run(i):
        extra = {"HOP_STATS": "1"}
        binary = encoder.HOP_ENCODER
        r = encoder.encode(binary, size, size, seed=200 + i, device=0, env_extra=extra,
                           launcher=(["taskset", "-c", str((2 * i) % cores)] if pin else None))
        res[i] = r
    t0 = time.perf_counter()
    th = [threading.Thread(target=run, args=(i,)) for i in range(k)]
    [t.start() for t in th]
    [t.join() for t in th]
    wall = time.perf_counter() - t0
    secs = [r["seconds"] for r in res]
    gpu = []
    for r in res:
        for line in r["log"].splitlines():
            if "xPatternSearchGT " in line and "calls" in line and "x" not in line.split("xPatternSearchGT")[1][:6]:
                gpu.append(line.split("calls")[1].split("s")[0].strip())
    print("K=%d pin=%s: per-process s %s  makespan %.2f s  images/s %.3f  GT-call s %s" % (
        k, pin, ["%.2f" % s for s in secs], wall, k / wall, gpu), flush=True)
