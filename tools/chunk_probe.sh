# A/B: rows per chunk of the layer stack (CE_GPU_CHUNK_ROWS; device-resident input uses twice the cap)
for c in ${CHUNK_ROWS:-8192 16384 32768 65536 131072 262144}; do
  CE_GPU_CHUNK_ROWS=$c python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk_rows',$c, d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_ms_per_step'])"
done
