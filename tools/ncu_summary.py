"""Print the ncu metrics that matter for the HOP kernels from an `ncu --page raw --csv` export."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
pat = (r'gpu__time_duration.sum|pipe_(fp64|alu|fma|xu|lsu)\.avg\.pct_of_peak_sustained_active|sm__warps_active.avg.pct|'
       r'smsp__issue_active.avg.pct|registers_per_thread$|occupancy_limit_(registers|shared_mem|warps)|'
       r'bank_conflicts_pipe_lsu_mem_shared.sum|wavefronts_mem_shared.sum$|smsp__inst_executed.sum$|'
       r'warps_eligible.avg|dram__bytes_(read|write).sum$|shared_mem_per_block$|lts__t_bytes.sum$|sm__throughput.avg.pct')
for k in h:
    if re.search(pat, k):
        i = h.index(k)
        print(k, [r[i] for r in rows[2:]])
st = []
for k in h:
    if re.search(r'smsp__average_warps_issue_stalled.*per_issue_active', k):
        i = h.index(k)
        st.append((float(rows[-1][i]), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
print('stalls (last launch):', [(round(v, 2), k) for v, k in sorted(st, reverse=True)[:8]])
