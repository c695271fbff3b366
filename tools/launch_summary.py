"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys


def summarize(path, top=25):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row["Metric Unit"]
        t = t / 1e3 if unit == "ns" else (t * 1e3 if unit == "ms" else t)
        short = re.sub(r"\(.*", "", row["Kernel Name"])[:64]
        agg[short][0] += 1
        agg[short][1] += t
        n += 1
    tot = sum(v[1] for v in agg.values())
    out = [f"launches {n}, total device time {tot / 1e3:.1f} ms (cold-cache, serialised: compare SHARES)", "",
           "| share | total us | launches | avg us | kernel |", "|---:|---:|---:|---:|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append(f"| {100 * v[1] / tot:.1f}% | {v[1]:.0f} | {v[0]} | {v[1] / v[0]:.1f} | `{k}` |")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarize(sys.argv[1]))
