"""Top stall sites of a kernel from an .ncu-rep (source page, SASS level).  Usage: ncu_hot.py rep [n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(raw))
hdr = None
data = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
ix = {k: i for i, k in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
tot_inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
print(f"total samples {tot}, warp instructions executed {tot_inst}")
agg = {}
for k in ("stall_long_sb", "stall_wait", "stall_math", "stall_no_inst", "stall_short_sb", "stall_branch_resolving", "stall_not_selected", "stall_selected", "stall_lg", "stall_dispatch", "stall_mio", "stall_barrier", "stall_membar", "stall_drain", "stall_sleep", "stall_misc"):
    if k in ix:
        agg[k] = sum(int(r[ix[k]] or 0) for r in data)
print({k: round(v / tot, 3) for k, v in agg.items() if v})
data.sort(key=lambda r: -int(r[ix["# Samples"]] or 0))
for r in data[:n]:
    st = {k[6:]: int(r[ix[k]] or 0) for k in agg if int(r[ix[k]] or 0)}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{r[ix['Address']][-5:]} {int(r[ix['# Samples']]):6d} ({100 * int(r[ix['# Samples']]) / tot:4.1f}%) exec {r[ix['Instructions Executed']]:>10}  {r[ix['Source']][:70]:70s} {top}")
