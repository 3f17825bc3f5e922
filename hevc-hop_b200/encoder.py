"""Drive an HM-15 / HEVC-HOP encoder binary on synthetic lenslet images.

  TAppEncoderHop  integration/_build/  the reference encoder with the drop-in patch,
                                        linked against libhopgpu.so -- the product path

It is built in the build container (it needs /root/reference) and travels to the GPU box as a file.  The
unmodified CPU reference binary used as checker / baseline is known to the test helpers only.
"""
import os
import subprocess
import tempfile
import time

from .lenslet import lenslet_luma, write_yuv420

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOP_ENCODER = os.path.join(ROOT, "integration", "_build", "TAppEncoderHop")
HOP_ENCODER_BATCH = os.path.join(ROOT, "integration", "_build", "TAppEncoderHopBatch")   # job queue on stdin, one CUDA context
CFG = os.path.join(ROOT, "integration", "hop_intra.cfg")
CFG_LOWDELAY_P = os.path.join(ROOT, "integration", "hop_lowdelay_p.cfg")   # ISS frame + PSS frames (temporal + SS reference)


def encode(binary, width, height, seed=0, qp=32, bit_depth=8, device=0, workdir=None, keep=False, retries=0,
           crop_of=None, cfg=None, env_extra=None, frames=1, extra_args=(), launcher=None, input_yuv=None):
    """Encode one synthetic lenslet frame; returns dict(bitstream=bytes, seconds=float, rec=bytes, log=str).

    crop_of=(W, H): the frame is the top-left width x height region of the W x H image with this seed;
    cfg: another encoder configuration file than integration/hop_intra.cfg; env_extra: extra environment;
    frames > 1: frame k is the lenslet image with seed + k (low-delay P cfg: ISS frame followed by PSS frames);
    input_yuv: an already written input file of that geometry (several encodes of one image, e.g. a QP sweep)."""
    if not os.path.exists(binary):
        raise FileNotFoundError(binary + " not built (python __graft_entry__.py in the build container)")
    tmp = workdir or tempfile.mkdtemp(prefix="hopenc_")
    yuv = input_yuv or os.path.join(tmp, "in.yuv")
    for k in range(0 if input_yuv else frames):
        write_yuv420(yuv, lenslet_luma(width, height, seed=seed + k, bit_depth=bit_depth, crop_of=crop_of), bit_depth=bit_depth,
                     append=k > 0)
    cmd = [binary, "-c", cfg or CFG, "-i", yuv, "-wdt", str(width), "-hgt", str(height), "-fr", "30", "-f", str(frames),
           "-b", "str.bin", "-o", "rec.yuv", "--MIsize=15", "--QP=%d" % qp]
    if bit_depth != 8:
        cmd += ["--InputBitDepth=%d" % bit_depth, "--InternalBitDepth=%d" % bit_depth]
    if width % 8 or height % 8:
        cmd += ["--ConformanceMode=1"]
    cmd += list(extra_args)
    if launcher:                      # e.g. ["taskset", "-c", "3"]
        cmd = list(launcher) + cmd
    env = _child_env(device, env_extra)
    t0 = time.perf_counter()
    p = subprocess.run(cmd, cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    dt = time.perf_counter() - t0
    for _ in range(retries):
        if p.returncode == 0:
            break
        # callers that run the unmodified CPU reference as a checker allow it one more try (it was seen to
        # die once right after start-up on a fresh box, by a signal, without a message)
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


def _child_env(device, env_extra):
    """Environment of an encoder process on GPU `device`.  The process sees only ITS GPU (HOP_ENC_PIN=0 turns that off):
    CUDA start-up enumerates and maps every visible device -- 7.1 s against 2.1 s for a lone start on a 4-GPU box
    (profiles/r02_ctx_create_probe_4gpu_box.txt)."""
    env = dict(os.environ, HOP_DEVICE=str(device))
    if os.environ.get("HOP_ENC_PIN", "1") != "0":
        vis = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v != ""]
        env["CUDA_VISIBLE_DEVICES"] = vis[device] if device < len(vis) else str(device)
        env["HOP_DEVICE"] = "0"
    env.update(env_extra or {})
    return env


class EncoderWorker:
    """One long-lived encoder process (integration/hop_batch_main.cpp): the reference's TAppEncTop run once per job
    inside a process that keeps libhopgpu's CUDA context, so a queue of images pays the CUDA start-up once.
    encode() takes the arguments of encoder.encode() and returns the same dict (plus `stats` = seconds the job spent
    in search calls / SS-mirror updates / context + mirror creation when HOP_STATS is set)."""

    def __init__(self, binary=None, device=0, env_extra=None):
        self.binary = binary or HOP_ENCODER_BATCH
        if not os.path.exists(self.binary):
            raise FileNotFoundError(self.binary + " not built (python __graft_entry__.py in the build container)")
        env = _child_env(device, env_extra)
        env.setdefault("CUDA_MODULE_LOADING", "EAGER")    # kernels are loaded while the worker starts, not inside its first image
        t0 = time.perf_counter()
        self.p = subprocess.Popen([self.binary], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                  text=True, bufsize=1, env=env)
        self.ready_line = self._read_until("@@HOPBATCH ready")[1]
        self.startup_seconds = time.perf_counter() - t0

    def _read_until(self, marker):
        log = []
        while True:
            line = self.p.stdout.readline()
            if line == "":
                raise RuntimeError("encoder worker died:\n" + "".join(log[-30:]))
            if line.startswith(marker):
                return "".join(log), line.strip()
            log.append(line)

    def encode(self, width, height, seed=0, qp=32, bit_depth=8, crop_of=None, cfg=None, frames=1, extra_args=(), keep=False,
               input_yuv=None):
        tmp = tempfile.mkdtemp(prefix="hopenc_")
        yuv = input_yuv or os.path.join(tmp, "in.yuv")
        for k in range(0 if input_yuv else frames):
            write_yuv420(yuv, lenslet_luma(width, height, seed=seed + k, bit_depth=bit_depth, crop_of=crop_of), bit_depth=bit_depth,
                         append=k > 0)
        args = ["-c", cfg or CFG, "-i", yuv, "-wdt", str(width), "-hgt", str(height), "-fr", "30", "-f", str(frames),
                "-b", "str.bin", "-o", "rec.yuv", "--MIsize=15", "--QP=%d" % qp]
        if bit_depth != 8:
            args += ["--InputBitDepth=%d" % bit_depth, "--InternalBitDepth=%d" % bit_depth]
        if width % 8 or height % 8:
            args += ["--ConformanceMode=1"]
        args += list(extra_args)
        t0 = time.perf_counter()
        self.p.stdin.write("\t".join([tmp] + args) + "\n")
        self.p.stdin.flush()
        log, done = self._read_until("@@HOPBATCH done")
        dt = time.perf_counter() - t0
        rc = int(done.split()[2])
        if rc != 0 or not os.path.exists(os.path.join(tmp, "str.bin")):
            raise RuntimeError("encoder worker job failed (rc=%d):\n%s" % (rc, log[-800:]))
        stats = None
        for ln in log.splitlines():
            if ln.startswith("@@HOPBATCH stats"):
                f = ln.split()
                stats = {"gpu_search_calls": float(f[2]), "ss_mirror_updates": float(f[3]), "context_and_mirror_create": float(f[4])}
        out = {"bitstream": open(os.path.join(tmp, "str.bin"), "rb").read(), "rec": open(os.path.join(tmp, "rec.yuv"), "rb").read(),
               "seconds": dt, "log": log, "dir": tmp, "stats": stats}
        trace = os.path.join(tmp, "TraceEnc.txt")
        out["trace"] = open(trace, "rb").read() if os.path.exists(trace) else b""
        if not keep:
            import shutil
            shutil.rmtree(tmp, ignore_errors=True)
        return out

    def close(self):
        if self.p and self.p.poll() is None:
            try:
                self.p.stdin.write("\n")
                self.p.stdin.flush()
                self.p.wait(timeout=30)
            except Exception:
                self.p.kill()
        self.p = None
