"""Rewrite the table of profiles/r02_parity_elementwise.md from gpurun_out/parity_elementwise.jsonl (the rows
`test_flow_matches_reference_golden` records on the GPU box; the last run wins per (case, mode))."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = {}
for l in open(os.path.join(ROOT, "gpurun_out", "parity_elementwise.jsonl")):
    r = json.loads(l)
    rows[(r["case"], r["mode"])] = r
path = os.path.join(ROOT, "profiles", "r02_parity_elementwise.md")
text = open(path).read()
head = text[:text.index("| case | mode |")]
f = lambda v: f"{v:.1e}"
out = ["| case | mode | lp max | lp p99.9 | z max | z p99.9 | y max | y p99.9 | ref fp32-vs-fp64 z max | z p99.9 |",
       "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|"]
for (case, mode), r in rows.items():
    out.append(f"| {case} | {mode} | {f(r['lp_max'])} | {f(r['lp_p999'])} | {f(r['z_max'])} | {f(r['z_p999'])} | {f(r['y_max'])} | "
               f"{f(r['y_p999'])} | {f(r['ref_fp32_vs_fp64_z_max'])} | {f(r['ref_fp32_vs_fp64_z_p999'])} |")
open(path, "w").write(head + "\n".join(out) + "\n")
print(len(rows), "rows")
