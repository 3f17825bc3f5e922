# usage: bash tools/_run_n.sh N   -- the bench under torchrun on N GPUs with a queue of 8 images per GPU (64 at N = 8)
N=$1
O=gpurun_out/r2s
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader > $O/gpus_$N.txt
nproc >> $O/gpus_$N.txt
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --encode-images 8 > $O/bench_$N.json 2> $O/bench_$N.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --encode-images 8 > $O/bench_$N.json 2> $O/bench_$N.err
fi
tail -c 600 $O/bench_$N.err
python - <<PY
import json
d=json.loads(open('$O/bench_$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['e2e']['value'])
print(json.dumps(d['encode'])[:1500])
print(json.dumps(d['sweep'])[:900])
PY
