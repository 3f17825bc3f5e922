"""Readable summary of an `ncu --set full` capture exported on the GPU box:
    python tools/ncu_report.py raw.csv [source.csv[.gz] ...]  > profiles/rNN_ncu_summary.txt
raw.csv    = `ncu -i rep --page raw --csv`      -> one block per launch (duration, launch geometry, pipes, stalls, DRAM)
source.csv = `ncu -i rep --page source --csv`   -> per launch: executed warp instructions by opcode (SASS view)"""
import collections, csv, gzip, io, re, sys

csv.field_size_limit(1 << 30)


def col(h, pat):
    for i, k in enumerate(h):
        if re.search(pat, k):
            return i
    return None


def raw_report(path):
    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    want = [("duration", r"gpu__time_duration\.sum$"), ("regs", r"launch__registers_per_thread$"),
            ("smem/CTA", r"launch__shared_mem_per_block$"), ("occ lim regs", r"occupancy_limit_registers$"),
            ("occ lim smem", r"occupancy_limit_shared_mem$"), ("warps active %", r"sm__warps_active\.avg\.pct_of_peak_sustained_active$"),
            ("issue active %", r"smsp__issue_active\.avg\.pct_of_peak_sustained_active$"),
            ("eligible warps/cycle", r"smsp__warps_eligible\.avg\.per_cycle_active$"),
            ("warp instr", r"smsp__inst_executed\.sum$"),
            ("pipe alu %", r"sm__inst_executed_pipe_alu\.avg\.pct_of_peak_sustained_active$"),
            ("pipe fma %", r"sm__inst_executed_pipe_fma\.avg\.pct_of_peak_sustained_active$"),
            ("pipe fp64 %", r"sm__inst_executed_pipe_fp64\.avg\.pct_of_peak_sustained_active$"),
            ("pipe xu %", r"sm__inst_executed_pipe_xu\.avg\.pct_of_peak_sustained_active$"),
            ("pipe lsu %", r"sm__inst_executed_pipe_lsu\.avg\.pct_of_peak_sustained_active$"),
            ("sm throughput %", r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed$"),
            ("dram read", r"dram__bytes_read\.sum$"), ("dram write", r"dram__bytes_write\.sum$"),
            ("dram throughput %", r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$"),
            ("l2 bytes", r"lts__t_bytes\.sum$"),
            ("smem wavefronts", r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$"),
            ("smem bank conflicts", r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$")]
    idx = [(n, col(h, p)) for n, p in want]
    stall = [(i, k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
             for i, k in enumerate(h) if re.search(r"smsp__average_warps_issue_stalled.*per_issue_active", k)]
    ki, bi, gi = h.index("Kernel Name"), h.index("Block Size"), h.index("Grid Size")
    for r in rows[2:]:
        name = re.sub(r"\(int, const Hop.*", "", r[ki]).replace("void ", "")
        print("== launch %s  %s  block %s grid %s" % (r[0], name, r[bi], r[gi]))
        line = []
        for n, i in idx:
            if i is not None and r[i] != "":
                line.append("%s %s%s" % (n, r[i], (" " + units[i]) if units[i] and "%" not in units[i] else ""))
        for k in range(0, len(line), 6):
            print("   " + " | ".join(line[k:k + 6]))
        st = sorted(((float(r[i] or 0), k) for i, k in stall), reverse=True)[:7]
        print("   stalls per issue: " + ", ".join("%s %.2f" % (k, v) for v, k in st))


def source_report(path):
    f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
    rows = list(csv.reader(f))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    seen = set()
    for a, b in zip(starts[:-1], starts[1:]):
        sig = (rows[a][1], tuple(tuple(r[:8]) for r in rows[a + 2:a + 40]))
        if sig in seen:          # the export repeats every launch's listing
            continue
        seen.add(sig)
        name = re.sub(r"\(int, const Hop.*", "", rows[a][1]).replace("void hop::", "")
        h = rows[a + 1]
        si, ei, smp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        ops, samp, tot = collections.Counter(), collections.Counter(), 0
        for r in rows[a + 2:b]:
            if len(r) <= ei:
                continue
            parts = r[si].split()
            if not parts:
                continue
            tok = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
            op = tok.split(".")[0] + (".64" if "F64" in tok else "")
            n = int(r[ei] or 0)
            ops[op] += n
            tot += n
            samp[op] += int(r[smp] or 0)
        ts = max(1, sum(samp.values()))
        print("== %s: %d executed warp instructions (SASS view), %d static" % (name, tot, b - a - 2))
        print("   " + "  ".join("%s %.1f%% (%.1f%% of stall samples)" % (op, 100.0 * n / max(1, tot), 100.0 * samp[op] / ts)
                                for op, n in ops.most_common(16)))


if __name__ == "__main__":
    raw_report(sys.argv[1])
    for p in sys.argv[2:]:
        print()
        source_report(p)
