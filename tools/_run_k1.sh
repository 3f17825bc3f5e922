O=gpurun_out/r2q
mkdir -p $O
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "k1 or golden or mirror or randomised or device_entry or full_size" 2>&1 | tail -2
python bench.py --steps 3 --no-cpu-baseline --encode-size 0 --sweep-pus 0 > $O/bench_k1.json 2> $O/bench_k1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q/bench_k1.json').read().strip().splitlines()[-1])
k=d['k1_sad_search']; print('k1', k['frac_of_vabsdiff4_peak'], k['value'], k['parity_spot_check'], {s:round(v['pixel_sads_per_s']/k['vabsdiff4_peak_pixel_sads_per_s'],3) for s,v in k['per_shape'].items()}, {s:round(v['ms'],4) for s,v in k['per_shape'].items()})
PY
