#!/usr/bin/env python
"""Times cmvn_kernel alone through ce_gpu_cmvn (library CUDA events around the launch)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from catears_b200 import api, synth  # noqa: E402


def main():
    n_utts, T = [int(x) for x in (sys.argv[1:3] or (128, 998))]
    # realistic log-mel values (the front end's own output on synthetic speech-like audio, tiled): random
    # numbers would make x_t - x_{t-600} inexact all the time and measure the kernel's fp64 fallback
    pcm, _ = synth.synth_batch(1, 160000 * 4)
    fb = api.fbank(pcm)
    reps = (n_utts * T + fb.shape[0] - 1) // fb.shape[0]
    feats = np.ascontiguousarray(np.tile(fb, (reps, 1))[:n_utts * T])
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
