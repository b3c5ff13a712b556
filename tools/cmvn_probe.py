#!/usr/bin/env python
"""Times cmvn_kernel alone through ce_gpu_cmvn (library CUDA events around the launch)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from catears_b200 import api, synth  # noqa: E402


def main():
    n_utts, T = [int(x) for x in (sys.argv[1:3] or (128, 998))]
    rng = np.random.default_rng(0)
    feats = (rng.standard_normal((n_utts * T, 40)) * 3 + 12).astype(np.float32)
    off = np.arange(n_utts + 1, dtype=np.int64) * T
    stats = synth.default_cmvn_stats()
    api.cmvn(stats, feats, off)
    api.profile_enable(True)
    for _ in range(3):
        api.cmvn(stats, feats, off)
    tr = api.profile_trace()
    api.profile_enable(False)
    for c, a, b in tr:
        print("  %-9s %8.1f us  (%d utts x %d frames: %.0f clk/frame at 1.9 GHz)" % (c, 1e3 * (b - a), n_utts, T, (b - a) * 1e-3 * 1.9e9 / T))


if __name__ == "__main__":
    main()
