"""Opcode mix weighted by executed warp-instructions from an `ncu --page source --csv` export (SASS view).
usage: tools_ncu_opmix.py src.csv [warp_pixels]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]
si, ei = h.index('Source'), h.index('Instructions Executed')
smp = h.index('# Samples')
ops = collections.Counter(); samp = collections.Counter(); tot = 0
for r in rows[hi + 1:]:
    if len(r) <= ei: continue
    parts = r[si].split()
    if not parts: continue
    op = parts[0] if not parts[0].startswith('@') else parts[1]
    op = op.split('.')[0] + ('.64' if 'F64' in r[si].split()[0 if not parts[0].startswith('@') else 1] else '')
    n = int(r[ei] or 0)
    ops[op] += n; tot += n; samp[op] += int(r[smp] or 0)
wp = float(sys.argv[2]) if len(sys.argv) > 2 else None
print('total warp-instr', tot, 'per warp-pixel' if wp else '', tot / wp if wp else '')
ts = sum(samp.values())
for op, n in ops.most_common(28):
    print('%-12s %12d %6.2f%%  %s  samples %5.1f%%' % (op, n, 100.0 * n / tot, ('%.2f/px' % (n / wp)) if wp else '', 100.0 * samp[op] / ts))
