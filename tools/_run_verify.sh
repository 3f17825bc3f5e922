# final state check on one B200: smoke, full GPU suite, default bench (both arms)
O=gpurun_out/r2y
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
if [ "$1" != "nobench-tests" ]; then (time python -m pytest tests -m gpu -x -q) > $O/pytest.txt 2>&1; grep -E "passed|failed|error" $O/pytest.txt | tail -2; fi
python bench.py > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks'], d['parity_spot_check'], d['e2e_results_equal_resident'])
print(d['k1_sad_search']['frac_of_vabsdiff4_peak'], d['k1_sad_search']['parity_spot_check'], d['sweep']['candidates_per_s'], d['sweep']['parity_spot_check'])
e=d['encode']; print({k:e.get(k) for k in ('images','makespan_s','s_per_image','bitstream_identical','errors')})
PY
