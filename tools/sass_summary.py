"""Per-kernel SASS opcode counts of libymp_b200.so (cuobjdump -sass): which kernels use tcgen05 (UTCHMMA / LDTM /
STTM), TMA loads / stores / reductions (UTMALDG / UTMASTG / UTMAREDG), cp.async (LDGSTS), legacy HMMA, ...
Usage: python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "youku-mplug_b200", "ymp", "libymp_b200.so")
OPS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "LDGSTS", "HMMA.16816",
       "FFMA2", "MUFU.EX2", "SYNCS", "ACQBULK", "RED.E", "ATOM", "STL", "LDL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + ".") or (o == "UTCHMMA" and op.startswith("UTCHMMA") and ".2CTA" not in op):
                kernels[cur][o] += 1
                break
    names = list(kernels)
    dm = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS opcode counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a)\n")
    hdr = ["kernel", "instr"] + OPS
    print("| " + " | ".join(hdr) + " |")
    print("|" + "---|" * len(hdr))
    tot = collections.Counter()
    for n, d in zip(names, dm):
        c = kernels[n]
        short = re.sub(r"\(.*", "", d).replace("ymp::", "")
        print("| `" + short + "` | " + str(c["_total"]) + " | " + " | ".join(str(c[o]) if c[o] else "" for o in OPS) + " |")
        tot.update(c)
    print("| **all kernels** | " + str(tot["_total"]) + " | " + " | ".join(str(tot[o]) for o in OPS) + " |")


if __name__ == "__main__":
    sys.exit(main())
