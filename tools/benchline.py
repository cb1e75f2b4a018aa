#!/usr/bin/env python
"""Prints the essentials of a bench.py JSON line (value, ms/step, e2e, per-kernel ms)."""
import json, sys
lines = [l for l in open(sys.argv[1]) if l.startswith("{")]
if not lines:
    print("no JSON line in", sys.argv[1]); sys.exit(0)
d = json.loads(lines[-1])
print(d.get("n_gpus"), "gpus", d.get("value"), d.get("unit"), "ms/step", d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"),
      d.get("kernels_ms"), (d.get("grad_allreduce") or {}).get("mode"))
