# round-end evidence on one B200: GPU suite, bench (both arms), ncu --set full of every kernel family, launch list, K2 DRAM traffic
set -x
O=gpurun_out/r2w
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt
(time python -m pytest tests -m gpu -x -q) > $O/pytest.txt 2>&1
tail -3 $O/pytest.txt
python bench.py --encode-images 4 > $O/bench.json 2> $O/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
export NCU_WARM=0
python tools/ncu_targets.py > $O/targets.txt 2>&1 && \
ncu --set full --clock-control none -k regex:'k2_gt_search|k1_batch|k2_gt_sweep|k_motion_tail|k1_search|k6_predict|k7_intra|k3_dist|k4_' -o /tmp/full python tools/ncu_targets.py > $O/ncu_full.log 2>&1
ncu -i /tmp/full.ncu-rep --page raw --csv > $O/full_raw.csv 2>$O/export.err
for k in k2_gt_search k1_batch k2_gt_sweep k_motion_tail; do
  ncu -i /tmp/full.ncu-rep --page source --csv --kernel-name regex:$k > $O/src_$k.csv 2>>$O/export.err
done
gzip -9 -f $O/src_*.csv
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --encode-size 0 > $O/bench_short.json 2> $O/bench_short.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --encode-size 0 > $O/ncu_list.log 2>&1
du -sh gpurun_out; ls -la $O
