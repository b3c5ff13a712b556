#!/bin/bash
# Runs on the GPU box (via gpurun): tests, the default bench line, the reference arm, an ncu launch list with the
# metrics of profiles/traffic.json, and one full ncu capture of a hidden-layer launch and the fused output layer.
set -u
TAG=${1:-r03}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit=$?"
timeout 300 python tools/prof_one.py > /dev/null 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -s 34 -c 17 --csv --log-file gpurun_out/${TAG}_ncu_launches.csv python tools/prof_one.py > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 19 -c 2 \
    -o gpurun_out/${TAG}_full -f python tools/prof_one.py > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full exit=$?"
