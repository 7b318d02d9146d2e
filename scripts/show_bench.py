"""print the headline fields of a bench.py JSON line: python scripts/show_bench.py file"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
r = d.get("roofline") or {}
print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "gpu_launches")}, "e2e", (d.get("e2e") or {}).get("value"))
print({k: r.get(k) for k in ("frac", "achieved", "ms_per_launch", "ms_dense_per_launch", "ms_linearize_per_launch", "n_ipm_mean", "n_refine_rounds_mean")})
print("latency", d.get("latency_ms"), "solver", d.get("solver"), "clocks", d.get("clocks"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
