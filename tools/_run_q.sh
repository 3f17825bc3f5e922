O=gpurun_out/r2t
mkdir -p $O
python tools/ctx_create_probe.py 1,4,16 --mps > $O/ctx_probe_1gpu.txt 2>&1
cat $O/ctx_probe_1gpu.txt
python -m pytest tests/test_encoder_integration.py -m gpu -x -q -k "identical_to_reference or speculative" 2>&1 | tail -3
