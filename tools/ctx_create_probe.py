"""How long does a process need from exec to a usable libhopgpu context, alone and with K of them starting at once?
    python tools/ctx_create_probe.py [K,K,...] [--mps]
Each child: dlopen libhopgpu.so + hop_ctx_create + hop_ctx_destroy (C ABI through ctypes, no torch).  Variants: every
GPU visible (HOP_DEVICE picks one) against CUDA_VISIBLE_DEVICES narrowed to the child's GPU.  Children are spread
round-robin over the GPUs of the box."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CHILD = ("import ctypes,time,os,sys\n"
         "t0=time.perf_counter()\n"
         "lib=ctypes.CDLL(%r)\n"
         "h=ctypes.c_void_p()\n"
         "st=lib.hop_ctx_create(int(os.environ.get('HOP_DEVICE','0')),ctypes.byref(h))\n"
         "t1=time.perf_counter()\n"
         "lib.hop_ctx_destroy.argtypes=[ctypes.c_void_p]\n"
         "lib.hop_ctx_destroy(h)\n"
         "print(st, t1-t0, time.perf_counter()-t1)\n") % os.path.join(ROOT, "hevc-hop_b200", "libhopgpu.so")


def run(k, ngpu, pin, env_extra):
    procs = []
    t0 = time.perf_counter()
    for i in range(k):
        env = dict(os.environ, **env_extra)
        g = i % ngpu
        if pin:
            env["CUDA_VISIBLE_DEVICES"], env["HOP_DEVICE"] = str(g), "0"
        else:
            env["HOP_DEVICE"] = str(g)
        procs.append(subprocess.Popen([sys.executable, "-c", CHILD], env=env, stdout=subprocess.PIPE, text=True))
    outs = [p.communicate()[0].split() for p in procs]
    wall = time.perf_counter() - t0
    cr = [float(o[1]) for o in outs if len(o) == 3 and o[0] == "0"]
    return wall, cr, len(outs) - len(cr)


def main():
    ks = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else [1, 4, 16]
    ngpu = int(subprocess.run(["nvidia-smi", "-L"], stdout=subprocess.PIPE, text=True).stdout.count("GPU "))
    envs = [("no MPS", {})]
    if "--mps" in sys.argv:
        from __graft_entry__ import load_package
        load_package()
        from hevc_hop_b200 import batch
        if batch.mps_start():
            envs.append(("MPS", batch.mps_env()))
    print("GPUs on the box:", ngpu, " host cores:", os.cpu_count())
    for name, ee in envs:
        for pin in (False, True):
            for k in ks:
                wall, cr, bad = run(k, ngpu, pin, ee)
                print("%-6s %-22s K=%2d  create: mean %.2f s  max %.2f s   all done after %.2f s  failed %d" % (
                    name, "one GPU visible" if pin else "all GPUs visible", k, sum(cr) / max(1, len(cr)), max(cr) if cr else 0, wall, bad), flush=True)
    if len(envs) > 1:
        batch.mps_stop()


if __name__ == "__main__":
    main()
