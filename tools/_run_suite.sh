(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_final.txt 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_final.txt | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
