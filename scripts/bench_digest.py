"""Prints the few numbers of a bench.py JSON line one looks at first (used by the guarded GPU run scripts)."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("N", d["n_gpus"], "value", d["value"], "e2e", d["e2e"]["value"], "its", d["solve"]["outer_its"], "parity", d["parity"]["ok"], "a00_us", d["roofline"]["avg_launch_us"], "frac", d["roofline"]["frac"])
print("assembled", (d.get("assembled") or {}).get("value"), "comm", d.get("comm"))
print("profile", json.dumps(d.get("profile", {}).get("by_category_ms")), d.get("profile", {}).get("solve_ms"))
s = d.get("strong_128") or {}
print("128^3", s.get("value"), s.get("outer_its"), s.get("parity_ok"), "a00_us", s.get("element_kernel_avg_us"))
print("128^3 profile", json.dumps((s.get("profile") or {}).get("by_category_ms")), (s.get("profile") or {}).get("solve_ms"))
if d.get("cpu_baseline"):
    print("cpu", d["cpu_baseline"].get("value"), d["cpu_baseline"].get("cores"))
