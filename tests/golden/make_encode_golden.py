#!/usr/bin/env python
"""Golden bitstreams of the UNMODIFIED CPU reference encoder at the sizes BASELINE.json quotes.

    python tests/golden/make_encode_golden.py [-j PROCS] [case ...]

Runs oracle/_ref/TAppEncoderRef (the reference compiled from /root/reference by `make -C oracle ref`) with the
reference's OWN cfg (cfg/3DHencoder_intra_main.cfg) on the synthetic lenslet frames below and records, per case,
md5 + size of str.bin, md5 of rec.yuv and the CPU seconds, in tests/golden/encode_golden.json.  The `-m gpu`
test test_encoder_integration.py::test_golden_bitstreams re-encodes the same frames through the GPU-backed
encoder with integration/hop_intra.cfg and compares md5s, which pins (a) bitstream identity at 1024x1024
(BASELINE configs[0]) and on a region of the 7728x5368 image at QP 22/27/32/37 (configs[1]), (b) that the repo's
cfg is equivalent to the reference's.  Needs /root/reference (build container only); the json travels.
"""
import hashlib
import json
import multiprocessing
import os
import platform
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(HERE, "encode_golden.json")
REF_CFG = "/root/reference/cfg/3DHencoder_intra_main.cfg"
REF_CFG_P = "/root/reference/cfg/3dHencoder_lowdelay_P_main.cfg"   # used with --ScalableHolo=0: ISS frame, then PSS frames
ILLUM = (7728, 5368)

CASES = {
    # BASELINE configs[0]
    "c0_1024x1024_qp32": dict(width=1024, height=1024, qp=32, seed=0),
    # BASELINE configs[1]: top-left region of the Illum-size image, the four QPs of the sweep
    "c1_illum_crop_512x256_qp22": dict(width=512, height=256, qp=22, seed=0, crop_of=ILLUM),
    "c1_illum_crop_512x256_qp27": dict(width=512, height=256, qp=27, seed=0, crop_of=ILLUM),
    "c1_illum_crop_512x256_qp32": dict(width=512, height=256, qp=32, seed=0, crop_of=ILLUM),
    "c1_illum_crop_512x256_qp37": dict(width=512, height=256, qp=37, seed=0, crop_of=ILLUM),
    # mid size, another seed (BASELINE configs[3] uses seeds 0..63)
    "c3_512x512_seed5_qp32": dict(width=512, height=512, qp=32, seed=5),
    # Main10 input (configs[4] content)
    "c4_main10_256x128_qp32": dict(width=256, height=128, qp=32, seed=2, bit_depth=10),
    # low-delay P: frame 0 ISS, frames 1.. PSS (list 0 = previous pictures + the SS reference, reset per slice)
    "pss_3frames_128x64_qp32": dict(width=128, height=64, qp=32, seed=0, frames=3, lowdelay_p=True),
}


def run(name):
    import __graft_entry__
    __graft_entry__.load_package()
    import _oracle
    kw = dict(CASES[name])
    cfg = REF_CFG
    if kw.pop("lowdelay_p", False):
        cfg = REF_CFG_P
        kw["extra_args"] = ("--ScalableHolo=0",)
    r = _oracle.encode_reference(kw.pop("width"), kw.pop("height"), cfg=cfg, **kw)
    rec = dict(CASES[name])
    if "crop_of" in rec:
        rec["crop_of"] = list(rec["crop_of"])
    rec.update(bitstream_md5=hashlib.md5(r["bitstream"]).hexdigest(), bitstream_bytes=len(r["bitstream"]),
               rec_md5=hashlib.md5(r["rec"]).hexdigest(), cpu_seconds=round(r["seconds"], 2),
               cpu=platform.processor() or platform.machine(), cfg=os.path.relpath(cfg, "/root/reference") + " (reference's own)")
    return name, rec


def main():
    args = sys.argv[1:]
    procs = 4
    if args[:1] == ["-j"]:
        procs = int(args[1]); args = args[2:]
    names = args or list(CASES)
    golden = json.load(open(OUT)) if os.path.exists(OUT) else {}
    with multiprocessing.Pool(procs) as pool:
        for name, rec in pool.imap_unordered(run, names):
            golden[name] = rec
            json.dump(golden, open(OUT, "w"), indent=1, sort_keys=True)
            print(name, rec["bitstream_md5"], rec["bitstream_bytes"], "B", rec["cpu_seconds"], "s", flush=True)


if __name__ == "__main__":
    main()
