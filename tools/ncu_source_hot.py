"""Summarise the SASS page of one kernel of an .ncu-rep: opcode histogram (executed, stall samples) and the top stall sites.
usage: python tools/ncu_source_hot.py rep.ncu-rep [launch_index] [top_n]"""
import csv
import subprocess
import sys
from collections import Counter

rep, idx, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:150])
hdr = next(r for r in rows if r and r[0] == "Address")
data = [r for r in rows if r and r[0].startswith("0x")]
i_s, i_ex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
print("stall samples", sum(int(r[i_s]) for r in data), "warp instructions", sum(int(r[i_ex]) for r in data), "SASS lines", len(data))
c, cs = Counter(), Counter()
for r in data:
    t = r[1].split()
    op = t[1] if t[0].startswith("@") else t[0]
    c[op] += int(r[i_ex]); cs[op] += int(r[i_s])
print("opcode: executed / stall samples")
for op, n in c.most_common(28):
    print(f"  {op:28s} {n:>10d} {cs[op]:>8d}")
print("top stall sites (line, SASS, samples, executed)")
for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][i_s]))[:topn]):
    print(f"  {i:5d} {data[i][1][:100]:100s} {data[i][i_s]:>6s} {data[i][i_ex]:>9s}")
