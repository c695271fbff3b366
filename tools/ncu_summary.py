"""Condense an ncu report (`ncu -i X.ncu-rep --page raw --csv`) into a small markdown table."""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main(rep, notes=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(m, n) for m, n in METRICS if m in idx]
    print("| # | kernel | " + " | ".join(n for _, n in cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for k, r in enumerate(rows[2:]):
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ymp::", "")
        vals = []
        for m, _ in cols:
            v, u = r[idx[m]], units[idx[m]]
            try:
                v = f"{float(v):.1f}"
            except ValueError:
                pass
            vals.append(f"{v} {u}".strip())
        print(f"| {k} | `{name}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
