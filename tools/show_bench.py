"""Print the key numbers of a bench.py JSON line.  Usage: python tools/show_bench.py <file>"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline") or {}
print(f"value {d.get('value'):.4g} {d.get('unit')}  ms/step {d.get('ms_per_step'):.3f}  launches {d.get('gpu_launches')}")
print(f"roofline {r.get('achieved', 0):.1f}/{r.get('peak', 0):.0f} {r.get('unit')} frac {r.get('frac', 0):.3f} "
      f"(mode ceiling frac {r.get('frac_of_mode_ceiling')})")
e = d.get("e2e") or {}
print(f"e2e {e.get('value', 0):.4g} ({e.get('ms_per_step', 0):.2f} ms)  cpu {((d.get('cpu_baseline') or {}).get('value'))}")
print("modes", {k: (round(v["value"]), round(v["tflops"], 1), v["max_rel_diff_vs_fp32_mode"]) for k, v in (d.get("modes") or {}).items()})
print("breakdown", {k: round(v, 3) for k, v in (d.get("breakdown_ms") or {}).items()})
print("clocks", d.get("clocks"))
print("sample", d.get("sample"))
print("train", d.get("train"))
