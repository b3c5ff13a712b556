#!/usr/bin/env python
"""Times the tensor-core GEMM on a TDNN-layer shape through ce_gpu_gemm_u8 / ce_gpu_gemm_f32
(kernel time from the library's own CUDA events).  CE_GPU_GEMM_DEBUG selects timing probes."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from catears_b200 import api  # noqa: E402


def main():
    m, n, k = [int(x) for x in (sys.argv[1:4] or (65536, 1024, 3072))]
    kind = sys.argv[4] if len(sys.argv) > 4 else "int8"
    rng = np.random.default_rng(0)
    if kind == "int8":
        a = rng.integers(0, 256, (m, k), dtype=np.uint8)
        b = rng.integers(0, 256, (k, n), dtype=np.uint8)
        run = lambda: api.gemm_u8(a, 0.01, 3, b, 0.02, 5, want_acc=False)
    else:
        a = rng.standard_normal((m, k)).astype(np.float32)
        b = rng.standard_normal((k, n)).astype(np.float32)
        run = lambda: api.gemm_f32(a, b, kind)
    run()
    api.profile_enable(True)
    for _ in range(3):
        run()
    t = api.profile_read()["gemm"]
    ms = t[0] / t[1]
    ops = 2.0 * m * n * k
    print("debug=%s %s m=%d n=%d k=%d: %.1f us per launch, %.1f TOP/s" %
          (os.environ.get("CE_GPU_GEMM_DEBUG", "0"), kind, m, n, k, ms * 1e3, ops / ms / 1e9))


if __name__ == "__main__":
    main()
