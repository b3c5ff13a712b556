#!/usr/bin/env python
"""Reads bench.py's JSON line on stdin and prints the handful of numbers compared across runs."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    if line.startswith("{"):
        d = json.loads(line)
        print(tag, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("kernel_ms_per_step"), d.get("gpu_launches"))
