#!/usr/bin/env python
"""Times the memory-bound kernels of the AM path alone through the stage-level C ABI (kernel
time from the library's own CUDA events; staging copies are outside the timed scopes)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from catears_b200 import api  # noqa: E402


def main():
    rows, cols = [int(x) for x in (sys.argv[1:3] or (65536, 1024))]
    rng = np.random.default_rng(0)
    x = rng.standard_normal((rows, cols)).astype(np.float32)
    api.quantize(x)
    api.profile_enable(True)
    for _ in range(3):
        api.quantize(x)
    tr = api.profile_trace()
    api.profile_enable(False)
    per = len(tr) // 3
    for c, a, b in tr[-per:]:
        gb = rows * cols * 5 / 1e9
        print("  %-9s %8.1f us  (%.0f GB/s if this is the 4B-in/1B-out pass)" % (c, 1e3 * (b - a), gb / (b - a) * 1e3))


if __name__ == "__main__":
    main()
