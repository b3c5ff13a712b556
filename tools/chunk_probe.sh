# A/B: rows per chunk of the layer stack (do the fp32 activations stay in L2 between a GEMM and its Quantize?)
for c in 8192 16384 24576 32768 65536; do
  CE_GPU_CHUNK_ROWS=$c python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk_rows',$c, d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_ms_per_step'])"
done
