# ncu --set full of the search kernel families (after a plain run of the same command), CSV pages exported on the box
O=gpurun_out/r2u
mkdir -p $O
export NCU_WARM=0
python tools/ncu_targets.py k2 k1 sweep > $O/targets.txt 2>&1 && \
ncu --set full --clock-control none -k regex:'k2_gt_search|k1_batch|k2_gt_sweep' -o /tmp/full python tools/ncu_targets.py k2 k1 sweep > $O/ncu_full.log 2>&1
ncu -i /tmp/full.ncu-rep --page raw --csv > $O/full_raw.csv 2>$O/export.err
ncu -i /tmp/full.ncu-rep --page source --csv --kernel-name regex:'k2_gt_search' > $O/src_k2.csv 2>>$O/export.err
ncu -i /tmp/full.ncu-rep --page source --csv --kernel-name regex:'k1_batch' > $O/src_k1.csv 2>>$O/export.err
ncu -i /tmp/full.ncu-rep --page source --csv --kernel-name regex:'k2_gt_sweep' > $O/src_sweep.csv 2>>$O/export.err
gzip -9 -f $O/src_*.csv
cat $O/targets.txt; du -sh $O
