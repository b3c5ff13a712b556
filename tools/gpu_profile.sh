#!/bin/bash
# Runs on the GPU box (via gpurun): bench line, ncu launch list, ncu full capture of the top kernels.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
SMALL="python bench.py --utts-per-gpu 64 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench exit=$?"; tail -c 3000 gpurun_out/bench_${TAG}.json
timeout 300 $SMALL > gpurun_out/plain_${TAG}.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $SMALL > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list exit=$?"
timeout 300 $SMALL > gpurun_out/plain2_${TAG}.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 8 -c 3 \
    -o gpurun_out/prof_gemm_${TAG} -f $SMALL > gpurun_out/ncu_gemm_${TAG}.log 2>&1
echo "ncu gemm exit=$?"
timeout 300 $SMALL > gpurun_out/plain3_${TAG}.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fbank_kernel|quantize_rows_kernel|finalize_rowcache|cmvn_kernel" -c 12 \
    -o gpurun_out/prof_misc_${TAG} -f $SMALL > gpurun_out/ncu_misc_${TAG}.log 2>&1
echo "ncu misc exit=$?"
