"""Batch encoding of independent lenslet images on one or several B200s (BASELINE.json configs[1] and configs[3],
SURVEY.md 8e).

One image keeps its causal dependency chain on one GPU, and the patched encoder spends about three quarters of
its time in the reference's untouched, single-threaded host RDO -- so the unit of parallelism is the encoder
PROCESS, and one B200 can serve many of them.  A rank (= one GPU) runs `procs` encoder processes that pull images
from the rank's queue.  Processes of one GPU share it through NVIDIA MPS when the control daemon can be started
(kernels of different processes then overlap; without it the driver time-slices whole contexts and a search call
waits for the other processes' slices: measured 0.65 s -> 3.2 s of GPU-call time per 512x512 image with two
processes).  No data moves between ranks: images are independent ("replicas only", no collective).
"""
import hashlib
import os
import queue
import shutil
import subprocess
import threading
import time

from . import encoder

MPS_DIR = "/tmp/hop_mps_%d" % os.getuid()


def mps_env():
    return {"CUDA_MPS_PIPE_DIRECTORY": os.path.join(MPS_DIR, "pipe"), "CUDA_MPS_LOG_DIRECTORY": os.path.join(MPS_DIR, "log")}


def mps_running():
    return os.path.exists(os.path.join(MPS_DIR, "pipe", "control"))


def mps_start():
    """Start the MPS control daemon for this node (idempotent).  Returns True when clients can attach."""
    exe = shutil.which("nvidia-cuda-mps-control")
    if exe is None:
        return False
    if mps_running():
        return True
    env = dict(os.environ, **mps_env())
    os.makedirs(env["CUDA_MPS_PIPE_DIRECTORY"], exist_ok=True)
    os.makedirs(env["CUDA_MPS_LOG_DIRECTORY"], exist_ok=True)
    try:
        subprocess.run([exe, "-d"], env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=30)
    except Exception:
        return False
    for _ in range(50):
        if mps_running():
            return True
        time.sleep(0.1)
    return False


def mps_stop():
    exe = shutil.which("nvidia-cuda-mps-control")
    if exe is None or not mps_running():
        return
    try:
        subprocess.run([exe], input="quit\n", text=True, env=dict(os.environ, **mps_env()), timeout=60,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except Exception:
        pass
    shutil.rmtree(MPS_DIR, ignore_errors=True)


def encode_batch(tasks, device=0, procs=1, use_mps=False, binary=None, keep_outputs=False, stats=False, persistent=None):
    """Encode `tasks` (dicts of encoder.encode keyword arguments incl. width/height) with `procs` concurrent
    encoder processes on GPU `device`.  Returns (results in task order, makespan seconds).

    persistent (default: when the batch front end is built and no other binary is asked for): every worker is ONE
    long-lived process (encoder.EncoderWorker, integration/hop_batch_main.cpp) that encodes its share of the queue, so
    the CUDA start-up -- serialised machine-wide by the driver, DESIGN.md section 4 -- is paid once per worker and not
    once per image; results carry `worker_startup_s`."""
    if persistent is None:
        persistent = binary is None and os.path.exists(encoder.HOP_ENCODER_BATCH) and os.environ.get("HOP_BATCH_WORKERS", "1") != "0"
    worker_args = {"width", "height", "seed", "qp", "bit_depth", "crop_of", "cfg", "frames", "extra_args", "input_yuv"}
    if persistent and any(set(t) - worker_args for t in tasks):
        persistent = False                     # a task asks for something only the one-process-per-image path offers
    if persistent:
        return _encode_batch_workers(tasks, device, procs, use_mps, keep_outputs, stats, binary)
    binary = binary or encoder.HOP_ENCODER
    q = queue.Queue()
    for i, t in enumerate(tasks):
        q.put((i, t))
    out = [None] * len(tasks)
    env_extra = dict(mps_env()) if use_mps else {}
    if stats:
        env_extra["HOP_STATS"] = "1"

    def worker():
        while True:
            try:
                i, t = q.get_nowait()
            except queue.Empty:
                return
            kw = dict(t)
            w, h = kw.pop("width"), kw.pop("height")
            t0 = time.perf_counter()
            try:
                r = encoder.encode(binary, w, h, device=device, env_extra=env_extra, keep=keep_outputs, **kw)
                out[i] = {"seconds": r["seconds"], "bytes": len(r["bitstream"]), "md5": hashlib.md5(r["bitstream"]).hexdigest(),
                          "rec_md5": hashlib.md5(r["rec"]).hexdigest(), "dir": r["dir"] if keep_outputs else None,
                          "log": r["log"] if stats else None, "start": t0}
            except Exception as e:      # an encoder failure is reported, never hidden
                out[i] = {"error": str(e)[:400], "seconds": time.perf_counter() - t0}

    t0 = time.perf_counter()
    threads = [threading.Thread(target=worker) for _ in range(max(1, min(procs, len(tasks))))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    return out, time.perf_counter() - t0


def _encode_batch_workers(tasks, device, procs, use_mps, keep_outputs, stats, binary=None):
    q = queue.Queue()
    for i, t in enumerate(tasks):
        q.put((i, t))
    out = [None] * len(tasks)
    env_extra = dict(mps_env()) if use_mps else {}
    if stats:
        env_extra["HOP_STATS"] = "1"

    def worker():
        t_start = time.perf_counter()
        try:
            w = encoder.EncoderWorker(binary=binary, device=device, env_extra=env_extra)
        except Exception as e:
            while True:                                  # a worker that cannot start fails its share loudly
                try:
                    i, _ = q.get_nowait()
                except queue.Empty:
                    return
                out[i] = {"error": "worker start: " + str(e)[:300], "seconds": time.perf_counter() - t_start}
        first = True
        while True:
            try:
                i, t = q.get_nowait()
            except queue.Empty:
                break
            kw = dict(t)
            wd, ht = kw.pop("width"), kw.pop("height")
            t0 = time.perf_counter()
            try:
                r = w.encode(wd, ht, keep=keep_outputs, **kw)
                out[i] = {"seconds": r["seconds"], "bytes": len(r["bitstream"]), "md5": hashlib.md5(r["bitstream"]).hexdigest(),
                          "rec_md5": hashlib.md5(r["rec"]).hexdigest(), "dir": r["dir"] if keep_outputs else None,
                          "log": r["log"] if stats else None, "start": t0, "stats": r["stats"],
                          "worker_startup_s": w.startup_seconds if first else 0.0}
            except Exception as e:                       # an encoder failure is reported, never hidden
                out[i] = {"error": str(e)[:400], "seconds": time.perf_counter() - t0}
                try:
                    w.close()
                    w = encoder.EncoderWorker(binary=binary, device=device, env_extra=env_extra)
                except Exception:
                    break
            first = False
        w.close()

    t0 = time.perf_counter()
    threads = [threading.Thread(target=worker) for _ in range(max(1, min(procs, len(tasks))))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    for i, o in enumerate(out):
        if o is None:
            out[i] = {"error": "not encoded (its worker died)", "seconds": 0.0}
    return out, time.perf_counter() - t0


def decoder_round_trip(decoder_binary, workdir, bit_depth=8):
    """Decode workdir/str.bin with a decoder binary and compare with the encoder's reconstruction rec.yuv."""
    dec = os.path.join(workdir, "dec.yuv")
    cmd = [decoder_binary, "-b", os.path.join(workdir, "str.bin"), "-o", dec]
    if bit_depth != 8:
        cmd += ["-d", str(bit_depth)]
    p = subprocess.run(cmd, cwd=workdir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0 or not os.path.exists(dec):
        return False, p.stdout[-400:]
    same = open(dec, "rb").read() == open(os.path.join(workdir, "rec.yuv"), "rb").read()
    os.remove(dec)
    return same, ""
