#!/usr/bin/env python
"""Cost of the selected outputs (ce_gpu_model_set_output) on the bench batch: milliseconds of the
finalize category per step for dense rows, top-k pairs and a pdf subset (device buffers)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from catears_b200 import api, synth  # noqa: E402


def main():
    import torch
    n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    conf = os.path.join(bench.model_dir(), "tdnn.conf")
    pcm, off = synth.synth_batch(n_utts, 160000)
    model = api.AcousticModelGpu(config=conf, precision="int8")
    d_pcm = torch.from_numpy(pcm).cuda()
    frames = int(api.frame_offsets(off)[-1])
    d_am = torch.empty(frames, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream()
    rng = np.random.default_rng(1)
    cases = [("dense", {}, model.num_pdfs)]
    for k in (16, 64, 128, 256, 1024):
        cases.append(("topk", {"k": k}, 2 * k))
    for n in (128, 512, 2048):
        cases.append(("subset", {"pdf_ids": np.sort(rng.permutation(model.num_pdfs)[:n])}, n))
    only = os.environ.get("SELECT_CASES")              # e.g. "topk:64,subset:128"
    if only:
        cases = []
        for c in only.split(","):
            mode, n = c.split(":")
            if mode == "dense":
                cases.append(("dense", {}, model.num_pdfs))
            elif mode == "topk":
                cases.append(("topk", {"k": int(n)}, 2 * int(n)))
            else:
                cases.append(("subset", {"pdf_ids": np.sort(rng.permutation(model.num_pdfs)[:int(n)])}, int(n)))
    for mode, kw, width in cases:
        model.set_output(mode, **kw)
        d_out = torch.empty((frames, width), dtype=torch.float32, device="cuda")
        for _ in range(2):
            model.forward(d_pcm, off, loglik=d_out, argmax=d_am, stream=s)
        torch.cuda.synchronize()
        api.profile_enable(True)
        n = 3
        for _ in range(n):
            model.forward(d_pcm, off, loglik=d_out, argmax=d_am, stream=s)
        prof = api.profile_read()
        api.profile_enable(False)
        total = sum(v[0] for v in prof.values()) / n
        print("%-7s %-14s row %6d B  finalize %.3f ms/step  (all kernels %.3f ms/step)" % (
            mode, ",".join("%s=%s" % (a, b if np.isscalar(b) else len(b)) for a, b in kw.items()),
            4 * width, prof["finalize"][0] / n, total), flush=True)
        del d_out


if __name__ == "__main__":
    main()
