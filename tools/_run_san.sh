O=gpurun_out/r2v
mkdir -p $O
python tools/sanitize_smoke.py > $O/smoke_plain.txt 2>&1 && \
timeout 1200 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > $O/memcheck.txt 2>&1
tail -5 $O/smoke_plain.txt; tail -12 $O/memcheck.txt
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "packed_tile or device_entry or async" 2>&1 | tail -3
