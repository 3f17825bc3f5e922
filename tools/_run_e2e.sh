python bench.py --steps 10 --no-cpu-baseline --encode-size 0 --k1-pus 0 --sweep-pus 0 > gpurun_out/e2e.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/e2e.json').read().strip().splitlines()[-1])
print('e2e', d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ms_per_step'], d['e2e_results_equal_resident'], d['e2e']['api'])
PY
