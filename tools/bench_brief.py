#!/usr/bin/env python
"""One-line summary of a bench.py JSON line (value, ms per step, per-class kernel times). Diagnostic helper."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
tag = sys.argv[2] if len(sys.argv) > 2 else ""
ks = {k: (round(v["ms_per_step"], 2), int(v["launches_per_step"])) for k, v in (d.get("kernels") or {}).items()}
print(tag, round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None, ks)
for k, v in (d.get("also") or {}).items():
    print("   also", k, round(v.get("value", 0)), v.get("ms_per_step"))
