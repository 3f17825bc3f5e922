# parity tests of the search kernels + a short bench (K2, sweep): the loop for kernel work
O=gpurun_out/r2q
mkdir -p $O
(time python -m pytest tests/test_gpu_parity.py tests/test_sweep.py tests/test_frac_motion.py -m gpu -x -q) > $O/pytest.txt 2>&1
grep -E "passed|failed|error" $O/pytest.txt | tail -3
python bench.py --steps 10 --no-cpu-baseline --encode-size 0 --k1-pus 0 > $O/bench.json 2> $O/bench.err
tail -3 $O/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['per_shape_ms'], d['parity_spot_check'], d['e2e_results_equal_resident'])
print(d['sweep']['ms'], d['sweep']['candidates_per_s'], d['sweep']['parity_spot_check'])
PY
