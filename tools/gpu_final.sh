#!/bin/bash
# Final evidence of a round (run via gpurun on one GPU): GPU tests, smoke, the default bench line, the secondary lines.
set -u
TAG=${1:-r03f}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit=$?"
B="python bench.py --no-cpu-baseline"
$B --workload longform --exact --steps 10 --warmup 3 > gpurun_out/${TAG}_longform_exact.json 2>/dev/null
$B --workload longform --steps 10 --warmup 3 > gpurun_out/${TAG}_longform.json 2>/dev/null
$B --workload streaming --steps 200 --warmup 56 > gpurun_out/${TAG}_streaming.json 2>/dev/null
$B --workload frontend --steps 5 --warmup 3 > gpurun_out/${TAG}_frontend40.json 2>/dev/null
for f in bench longform_exact longform streaming frontend40; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$f.json").read().strip().splitlines()[-1])
    print("$f", d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("kernel_ms_per_step"), d.get("abi_ms_per_call"))
except Exception as e: print("$f", "ERR", e)
PY
done
