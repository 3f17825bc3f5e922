O=gpurun_out/r2z
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
lscpu | grep -i "numa\|socket\|^CPU(s)" > $O/lscpu.txt
for m in 1 0; do
  HOP_BENCH_NUMA=$m python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$m bench.py --gpus 8 --steps 10 --no-cpu-baseline --encode-size 0 --sweep-pus 0 --k1-pus 0 > $O/bench8_numa$m.json 2> $O/bench8_numa$m.err
  python - <<PY
import json
d=json.loads(open('$O/bench8_numa$m.json').read().strip().splitlines()[-1])
print('numa=$m', d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ms_per_step'], d.get('numa_binding_rank0'))
PY
done
cat $O/lscpu.txt; head -12 $O/topo.txt | cut -c1-150
