python - <<'PY'
import sys, time, threading
sys.path.insert(0,'tests'); import conftest
from hevc_hop_b200 import encoder, batch
mps = batch.mps_start()
env = dict(batch.mps_env()) if mps else {}
env["HOP_STATS"]="1"
def run(k):
    w = encoder.EncoderWorker(env_extra=env)
    print(k, 'startup', round(w.startup_seconds,2), w.ready_line, flush=True)
    for j in range(2):
        r = w.encode(512, 512, seed=10*k+j)
        print(k, j, 'seconds', round(r['seconds'],2), r['stats'], [l for l in r['log'].splitlines() if 'HOPBATCH' in l or 'Total Time' in l], flush=True)
    w.close()
th=[threading.Thread(target=run,args=(k,)) for k in range(4)]
[t.start() for t in th]; [t.join() for t in th]
batch.mps_stop()
PY
