O=gpurun_out/r2x
mkdir -p $O
python -m pytest tests/test_encoder_integration.py -m gpu -x -q -k "long_lived or batch_driver" 2>&1 | tail -4 | cut -c1-300
python bench.py --steps 3 --no-cpu-baseline --sweep-pus 0 --k1-pus 0 --encode-images 4 > $O/bench_workers.json 2> $O/bench_workers.err
python - <<'PY'
import json
for f in ("workers",):
    d=json.loads(open('gpurun_out/r2x/bench_%s.json'%f).read().strip().splitlines()[-1])
    e=d['encode']; print(f, {k:e.get(k) for k in ('images','encoder_processes_per_gpu','makespan_s','images_per_s','s_per_image','bitstream_identical','errors','workers','s_per_image_split')})
PY
