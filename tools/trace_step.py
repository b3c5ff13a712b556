#!/usr/bin/env python
"""Timeline of one overlapped step (ce_gpu_profile_trace): per category busy time, the union of
the busy intervals, and how much of the GEMM time had a memory-bound kernel running beside it."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from catears_b200 import api, synth  # noqa: E402


def union(iv):
    iv = sorted(iv)
    tot, cur_a, cur_b = 0.0, None, None
    for a, b in iv:
        if cur_b is None or a > cur_b:
            if cur_b is not None:
                tot += cur_b - cur_a
            cur_a, cur_b = a, b
        else:
            cur_b = max(cur_b, b)
    if cur_b is not None:
        tot += cur_b - cur_a
    return tot


def main():
    import torch
    n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    conf = os.path.join(bench.model_dir(), "tdnn.conf")
    pcm, off = synth.synth_batch(n_utts, 160000)
    model = api.AcousticModelGpu(config=conf, precision="int8")
    host = os.environ.get("TRACE_HOST_PCM") == "1"     # e2e flavour: pinned host PCM in, host argmax out
    h_pcm = torch.from_numpy(pcm).pin_memory()
    d_pcm = h_pcm.numpy() if host else h_pcm.cuda()
    frames = int(api.frame_offsets(off)[-1])
    d_ll = torch.empty((frames, model.num_pdfs), dtype=torch.float32, device="cuda")
    d_am = torch.empty(frames, dtype=torch.int32).pin_memory().numpy() if host else \
        torch.empty(frames, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream()
    for _ in range(3):
        model.forward(d_pcm, off, loglik=d_ll, argmax=d_am, stream=s)
    torch.cuda.synchronize()
    api.profile_enable(True)
    import time
    t0 = time.perf_counter()
    model.forward(d_pcm, off, loglik=d_ll, argmax=d_am, stream=s)
    torch.cuda.synchronize()
    print("host wall time of the call: %.3f ms" % (1e3 * (time.perf_counter() - t0)))
    tr = api.profile_trace()
    api.profile_enable(False)
    t_end = max(t[2] for t in tr)
    print("records %d, span %.3f ms" % (len(tr), t_end))
    by = {}
    for c, a, b in tr:
        by.setdefault(c, []).append((a, b))
    for c, iv in by.items():
        print("  %-9s n=%4d sum %.3f ms union %.3f ms" % (c, len(iv), sum(b - a for a, b in iv), union(iv)))
    print("  union of everything %.3f ms" % union([(a, b) for _, a, b in tr]))
    # where the GPU idles inside the call: the largest gaps between consecutive kernels
    ev = sorted(tr, key=lambda t: t[1])
    gaps, end = [], ev[0][2]
    print("  first kernel starts at %.3f ms, last ends at %.3f ms" % (ev[0][1], t_end))
    for i in range(1, len(ev)):
        if ev[i][1] > end:
            gaps.append((ev[i][1] - end, i))
        end = max(end, ev[i][2])
    print("  idle inside the span: %.3f ms in %d gaps; largest:" % (sum(g for g, _ in gaps), len(gaps)))
    for g, i in sorted(gaps, reverse=True)[:10]:
        print("    %.1f us before record %d (%s at %.3f ms, after %s)" % (1e3 * g, i, ev[i][0], ev[i][1], ev[i - 1][0]))
    if len(sys.argv) > 2:
        for c, a, b in tr[:int(sys.argv[2])]:
            print("    %-9s %9.3f -> %9.3f  (%.1f us)" % (c, a, b, 1e3 * (b - a)))


if __name__ == "__main__":
    main()
