#!/usr/bin/env python
"""Prints the handful of numbers compared across runs from bench.py's JSON line.
Usage: bench_summary.py FILE [tag]   (FILE "-" = stdin)."""
import json
import sys

if len(sys.argv) < 2:
    sys.exit(__doc__)
tag = sys.argv[2] if len(sys.argv) > 2 else ""
src = sys.stdin if sys.argv[1] == "-" else open(sys.argv[1])
for line in src:
    if line.startswith("{"):
        d = json.loads(line)
        print(tag, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("kernel_ms_per_step"), d.get("gpu_launches"))
