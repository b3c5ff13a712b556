# Timing probes of the fused output layer (CE_GPU_GEMM_DEBUG bits, gemm.h): launch 21 of tools/prof_one.py
for d in ${LSM_PROBE_BITS:-0 1 8 16 24 2 32}; do
CE_GPU_GEMM_DEBUG=$d ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:gemm_kernel -s 20 -c 1 --csv --log-file gpurun_out/lsm_dbg$d.csv python tools/prof_one.py > /dev/null 2>&1
echo "debug=$d" $(grep -E "duration|tensor" gpurun_out/lsm_dbg$d.csv | awk -F, '{print $NF}')
done
