"""DRAM traffic of the GEMM launches of one training step vs their algorithmic operand bytes.
usage: python tools/gemm_traffic.py launches.csv gemm_shapes.json out.json
launches.csv: `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:^gemm_ --launch-skip S -c N
               --csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph`
gemm_shapes.json: `bench.py --gemm-report` of the same commit (per-shape launch counts and operand flags)."""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
im, iu, iv = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
tot = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0, "gpu__time_duration.sum": 0.0}
ids = set()
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
for r in rows[hi + 1:]:
    if len(r) <= iv or r[im] not in tot:
        continue
    ids.add(r[0])
    tot[r[im]] += float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
shapes = json.load(open(sys.argv[2]))
alg, n = 0.0, 0
for s in shapes:
    M, N, K = s["M"], s["N"], s["K"]
    out_b = 4 if s.get("out_f32") else 2
    b = (M * K + N * K) * 2 + M * N * out_b          # A, B read once, D written once
    # (accumulating / split-K launches: the partial tiles meet in L2 through reduce-adds; D is counted once)
    if s["res"]:
        b += M * N * (4 if s.get("res_f32") else 2)
    if s["aux_out"]:
        b += M * N * 2
    if s["aux_in"]:
        b += M * N * 2
    alg += b * s["n"]
    n += s["n"]
launches = len(ids)
out = dict(source="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:^gemm_ over the GEMM launches of one "
                  "bench.py step (B=32, --no-graph); algorithmic bytes from the --gemm-report of the same commit",
           launches=launches, dram_read_bytes=tot["dram__bytes_read.sum"], dram_write_bytes=tot["dram__bytes_write.sum"],
           bytes_per_launch=(tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / max(1, launches),
           duration_us_total=tot["gpu__time_duration.sum"], shape_report_launches=n,
           algorithmic_bytes_per_launch=alg / max(1, n),
           algorithmic_note="A + B + D (+ act' / residual / aux operands) read or written exactly once; broadcast residual tables counted as full tensors")
out["measured_over_algorithmic"] = out["bytes_per_launch"] / out["algorithmic_bytes_per_launch"]
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(json.dumps(out, indent=1))
