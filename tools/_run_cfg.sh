for c in 4 6 3; do
  HOP_K2_CFG=$c python bench.py --steps 5 --no-cpu-baseline --encode-size 0 --k1-pus 0 --sweep-pus 0 > gpurun_out/cfg_$c.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('gpurun_out/cfg_$c.json').read().strip().splitlines()[-1])
print('cfg $c', {k:round(v,4) for k,v in d['roofline']['per_shape_ms'].items()}, d['parity_spot_check'])
PY
done
