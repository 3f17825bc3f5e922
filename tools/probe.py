import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
hop = g.load_package()
ctx = hop.HopContext(0)
names = {0: "int32_add", 1: "vabsdiff4_acc", 2: "fp64_add", 3: "fp64_mul", 4: "fp64_fma", 5: "lop3"}
res = {}
for k, n in names.items():
    best = 0
    for _ in range(3):
        gops, ms = ctx.probe_alu(k)
        best = max(best, gops)
    res[n] = {"gops": best, "ms": ms}
print(json.dumps(res))
