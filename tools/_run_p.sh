set -x
O=gpurun_out/r2p
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt
(time python -m pytest tests -m gpu -x -q) > $O/pytest.txt 2>&1
tail -3 $O/pytest.txt
python bench.py > $O/bench.json 2> $O/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
export NCU_WARM=0
python tools/ncu_targets.py > $O/targets.txt 2>&1 && \
ncu --set full --clock-control none -k regex:'k2_gt_search|k1_batch|k2_gt_sweep|k_motion_tail|k1_search|k6_predict|k7_intra' -o /tmp/full python tools/ncu_targets.py > $O/ncu_full.log 2>&1
ls -la /tmp/full.ncu-rep
ncu -i /tmp/full.ncu-rep --page raw --csv > $O/full_raw.csv 2>$O/export.err
ncu -i /tmp/full.ncu-rep --page source --csv --kernel-name regex:'k2_gt_search' > $O/src_k2.csv 2>>$O/export.err
ncu -i /tmp/full.ncu-rep --page source --csv --kernel-name regex:'k1_batch' > $O/src_k1.csv 2>>$O/export.err
ncu -i /tmp/full.ncu-rep --page source --csv --kernel-name regex:'k2_gt_sweep' > $O/src_sweep.csv 2>>$O/export.err
ncu -i /tmp/full.ncu-rep --page source --csv --kernel-name regex:'k_motion_tail' > $O/src_tail.csv 2>>$O/export.err
gzip -9 $O/src_*.csv
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --encode-size 0 > $O/bench_short.json 2> $O/bench_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --encode-size 0 > $O/ncu_list.log 2>&1
du -sh gpurun_out; ls -la $O
