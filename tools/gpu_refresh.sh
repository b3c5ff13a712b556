#!/bin/bash
# Secondary bench lines of HEAD for profiles/ (run via gpurun on one GPU).
set -u
TAG=${1:-r03}
B="python bench.py --no-cpu-baseline"
$B --utts-per-gpu 4096 --steps 3 --warmup 3 --no-extra-legs > gpurun_out/${TAG}_bench_4096.json 2>/dev/null; echo "4096: $?"
$B --workload frontend --steps 5 --warmup 3 > gpurun_out/${TAG}_frontend40.json 2>/dev/null; echo "fe40: $?"
$B --workload frontend --mel 80 --steps 5 --warmup 3 > gpurun_out/${TAG}_frontend80.json 2>/dev/null; echo "fe80: $?"
$B --workload streaming --steps 20 --warmup 5 > gpurun_out/${TAG}_streaming.json 2>/dev/null; echo "streaming: $?"
$B --workload longform --steps 10 --warmup 3 > gpurun_out/${TAG}_longform.json 2>/dev/null; echo "longform: $?"
$B --workload longform --feed topk --steps 10 --warmup 3 > gpurun_out/${TAG}_longform_feed.json 2>/dev/null; echo "longform feed: $?"
for f in 4096 frontend40 frontend80 streaming longform longform_feed; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_$f.json" if "$f"=="4096" else "gpurun_out/${TAG}_$f.json").read().strip().splitlines()[-1])
    print("$f", d.get("value"), d.get("unit"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"))
except Exception as e: print("$f", "ERR", e)
PY
done
