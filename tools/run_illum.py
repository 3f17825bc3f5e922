"""BASELINE.json configs[1]: the full Lytro-Illum-size 7728x5368 synthetic lenslet image, HOP intra cfg, at
QP 22/27/32/37 on ONE B200 -- four encoder processes sharing the GPU through MPS (the encoder is host bound).
    python tools/run_illum.py [--size WxH] [--qps 22,27,32,37] [--out gpurun_out/illum.json]
Parity: every bitstream is decoded by the unmodified CPU reference decoder and compared with the encoder's
reconstruction (round trip); bitstream identity itself is pinned on a region of this image by
tests/test_encoder_integration.py::test_golden_bitstreams (c1_illum_crop_*)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import conftest  # noqa: F401
import _oracle
from hevc_hop_b200 import batch

ap = argparse.ArgumentParser()
ap.add_argument("--size", default="7728x5368")
ap.add_argument("--qps", default="22,27,32,37")
ap.add_argument("--out", default="gpurun_out/illum.json")
ap.add_argument("--no-mps", action="store_true")
a = ap.parse_args()
w, h = [int(x) for x in a.size.split("x")]
qps = [int(x) for x in a.qps.split(",")]
mps = (not a.no_mps) and batch.mps_start()
from hevc_hop_b200.lenslet import lenslet_luma, write_yuv420
yuv = "/tmp/hop_illum_%dx%d.yuv" % (w, h)
write_yuv420(yuv, lenslet_luma(w, h, seed=0))       # one input file for the four QP points
tasks = [dict(width=w, height=h, seed=0, qp=q, input_yuv=yuv) for q in qps]
t0 = time.time()
res, makespan = batch.encode_batch(tasks, device=0, procs=len(tasks), use_mps=mps, keep_outputs=True, stats=True)
if mps:
    batch.mps_stop()
ctus = ((w + 63) // 64) * ((h + 63) // 64)
report = {"image": "%dx%d synthetic lenslet (seed 0), HOP intra cfg, 1 frame" % (w, h), "ctus": ctus, "qps": qps,
          "processes_on_one_gpu": len(tasks), "mps": bool(mps), "makespan_s": makespan, "host_cores": os.cpu_count(), "per_qp": {}}
for q, r in zip(qps, res):
    if "error" in r:
        report["per_qp"][str(q)] = r
        continue
    ok, msg = batch.decoder_round_trip(_oracle.REF_DECODER, r["dir"]) if os.path.exists(_oracle.REF_DECODER) else (None, "decoder not built")
    stats = [l for l in (r["log"] or "").splitlines() if (l.startswith("hopshim:") and "x" not in l.split()[2][:3]) or l.startswith("@@HOPBATCH stats")][:9]
    report["per_qp"][str(q)] = {"s_per_image": r["seconds"], "s_per_ctu": r["seconds"] / ctus, "bytes": r["bytes"], "md5": r["md5"],
                                "decoder_round_trip_identical": ok, "decoder_msg": msg, "shim_stats": stats}
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
json.dump(report, open(a.out, "w"), indent=1)
print(json.dumps(report, indent=1))
