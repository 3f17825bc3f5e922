"""Drive the HM-15 / HEVC-HOP encoder binaries on synthetic lenslet images.

  TAppEncoderHop  integration/_build/  the reference encoder with the drop-in patch,
                                        linked against libhopgpu.so -- the product path
  TAppEncoderRef  oracle/_ref/         the unmodified reference (CPU), used only as checker / baseline

Both are built in the build container (they need /root/reference) and travel to the GPU box as files.
"""
import os
import subprocess
import tempfile
import time

from .lenslet import lenslet_luma, write_yuv420

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOP_ENCODER = os.path.join(ROOT, "integration", "_build", "TAppEncoderHop")
REF_ENCODER = os.path.join(ROOT, "oracle", "_ref", "TAppEncoderRef")
REF_DECODER = os.path.join(ROOT, "oracle", "_ref", "TAppDecoderRef")
CFG = os.path.join(ROOT, "integration", "hop_intra.cfg")


def encode(binary, width, height, seed=0, qp=32, bit_depth=8, device=0, workdir=None, keep=False):
    """Encode one synthetic lenslet frame; returns dict(bitstream=bytes, seconds=float, rec=bytes, log=str)."""
    if not os.path.exists(binary):
        raise FileNotFoundError(binary + " not built (make -C integration / make -C oracle ref)")
    tmp = workdir or tempfile.mkdtemp(prefix="hopenc_")
    yuv = os.path.join(tmp, "in.yuv")
    write_yuv420(yuv, lenslet_luma(width, height, seed=seed, bit_depth=bit_depth), bit_depth=bit_depth)
    cmd = [binary, "-c", CFG, "-i", yuv, "-wdt", str(width), "-hgt", str(height), "-fr", "30", "-f", "1",
           "-b", "str.bin", "-o", "rec.yuv", "--MIsize=15", "--QP=%d" % qp]
    if bit_depth != 8:
        cmd += ["--InputBitDepth=%d" % bit_depth, "--InternalBitDepth=%d" % bit_depth]
    if width % 8 or height % 8:
        cmd += ["--ConformanceMode=1"]
    env = dict(os.environ, HOP_DEVICE=str(device))
    t0 = time.perf_counter()
    p = subprocess.run(cmd, cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    dt = time.perf_counter() - t0
    if p.returncode != 0 and binary == REF_ENCODER:
        # the unmodified CPU reference was seen to die once right after start-up on a fresh box (signal, no
        # message); it is only the checker here, so it gets one more try -- the GPU encoder never does
        t0 = time.perf_counter()
        p = subprocess.run(cmd, cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        dt = time.perf_counter() - t0
    if p.returncode != 0 or not os.path.exists(os.path.join(tmp, "str.bin")):
        raise RuntimeError("encoder failed (rc=%d, cmd=%s):\n%s\n...\n%s" % (p.returncode, " ".join(cmd), p.stdout[:600], p.stdout[-600:]))
    out = {"bitstream": open(os.path.join(tmp, "str.bin"), "rb").read(),
           "rec": open(os.path.join(tmp, "rec.yuv"), "rb").read(), "seconds": dt, "log": p.stdout, "dir": tmp}
    trace = os.path.join(tmp, "TraceEnc.txt")
    out["trace"] = open(trace, "rb").read() if os.path.exists(trace) else b""
    if not keep and workdir is None:
        import shutil
        shutil.rmtree(tmp, ignore_errors=True)
    return out
