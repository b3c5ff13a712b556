#!/usr/bin/env python
"""Times fbank_kernel alone through ce_gpu_fbank on device-resident PCM (library CUDA events)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from catears_b200 import api, synth  # noqa: E402


def main():
    import torch
    n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    mel = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    pcm, off = synth.synth_batch(min(n_utts, 64), 160000)
    reps = (n_utts + 63) // 64
    pcm = np.tile(pcm, reps)[:n_utts * 160000]
    off = np.arange(n_utts + 1, dtype=np.int64) * 160000
    d_pcm = torch.from_numpy(pcm).cuda()
    frames = int(api.frame_offsets(off)[-1])
    out = torch.empty((frames, mel), dtype=torch.float32, device="cuda")
    api.fbank(d_pcm, off, num_mel=mel, out=out)
    torch.cuda.synchronize()
    api.profile_enable(True)
    for _ in range(3):
        api.fbank(d_pcm, off, num_mel=mel, out=out)
    torch.cuda.synchronize()
    tr = api.profile_trace()
    api.profile_enable(False)
    ms = min(b - a for _, a, b in tr)
    by = frames * (320 + 4 * mel)
    print("fbank %d utts mel %d: %.1f us, %.3f G frames/s, %.0f GB/s (%.2f%% of 6555)" %
          (n_utts, mel, ms * 1e3, frames / ms / 1e6, by / ms / 1e6, by / ms / 1e6 / 65.55))


if __name__ == "__main__":
    main()
