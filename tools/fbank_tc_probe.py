#!/usr/bin/env python
"""Measured prototype of a TENSOR-CORE DFT front end against fbank_kernel: the same
log-mel features computed as GEMMs on library kernels (torch / cuBLAS -- a prototype to decide whether a
hand-written version is worth building, not product code):

  frames [F x 512] (DC removed, pre-emphasised, Hamming-windowed, zero-padded)  x  [512 x 514] (cos | -sin)

  * direct  : one real DFT as a GEMM, operands split fp16 hi/lo (3 products, fp32 accumulate) or 3 x TF32
  * factored: 512 = 32 x 16 Cooley-Tukey as two batched small GEMMs with a twiddle multiply in between
              (complex arithmetic as real 2x2 blocks), same splits

then |X|^2 -> mel (dense [257 x mel] fp32 GEMM) -> floor -> log.  Prints frames/s of every stage and the
maximum deviation from the FFT kernel's features (bar: 1e-4)."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from catears_b200 import api, synth  # noqa: E402


def timed(fn, n=5):
    import torch
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


def main():
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    pcm, off = synth.synth_batch(n_utts, 160000)
    F = int(api.frame_offsets(off)[-1])
    d_pcm = torch.from_numpy(pcm).cuda()
    d_out = torch.empty((F, 40), dtype=torch.float32, device="cuda")
    ms_fft, _ = timed(lambda: api.fbank(d_pcm, off, out=d_out))
    want = d_out.clone()
    print("fbank_kernel (FFT): %.3f ms for %d frames = %.3f G frames/s" % (ms_fft, F, F / ms_fft / 1e6))

    # ---- windowing pass (what a fused kernel would do while staging the GEMM's A operand) ----
    ham = torch.from_numpy(np.hamming(400).astype(np.float64) * 0 + (0.54 - 0.46 * np.cos(2 * np.pi * np.arange(400) / 399))).cuda()

    def frames64():
        x = d_pcm.view(n_utts, 160000).double()
        fr = x.unfold(1, 400, 160)                            # [U, 998, 400]
        fr = fr - fr.mean(dim=2, keepdim=True)
        pre = torch.cat([fr[:, :, :1] - 0.97 * fr[:, :, :1], fr[:, :, 1:] - 0.97 * fr[:, :, :-1]], dim=2)
        return (pre * ham).reshape(-1, 400)
    X64 = frames64()
    X = torch.zeros((F, 512), dtype=torch.float32, device="cuda")
    X[:, :400] = X64.float()
    k = torch.arange(257, device="cuda", dtype=torch.float64)
    n = torch.arange(512, device="cuda", dtype=torch.float64)
    ang = 2 * math.pi * torch.outer(n, k) / 512
    W64 = torch.cat([torch.cos(ang), -torch.sin(ang)], dim=1)             # [512, 514]
    # (514 columns make cuBLAS take a slow path: the GEMMs below run on 576 zero-padded columns, 12 % more FLOPs)
    W64 = torch.cat([W64, torch.zeros((512, 62), dtype=torch.float64, device="cuda")], dim=1)
    mel = torch.from_numpy(mel_matrix(40)).cuda()                          # [257, 40] fp32

    def tail(Y):                                                           # power -> mel -> log (fp32 GEMM)
        keep = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        p = Y[:, :257] ** 2 + Y[:, 257:514] ** 2
        out = torch.log(torch.clamp(p @ mel, min=float(np.finfo(np.float32).eps)))
        torch.backends.cuda.matmul.allow_tf32 = keep
        return out

    def report(name, ms_gemm, Y):
        ms_tail, feats = timed(lambda: tail(Y))
        err = float((feats - want).abs().max())
        print("%-34s GEMM %.3f ms = %.3f G frames/s; + power/mel/log %.3f ms -> %.3f G frames/s; max |diff| %.2e"
              % (name, ms_gemm, F / ms_gemm / 1e6, ms_tail, F / (ms_gemm + ms_tail) / 1e6, err))

    # ---- direct DFT, 3 x TF32 ----
    def split_tf32(a):
        hi = (a.view(torch.int32) & ~0x1fff).view(torch.float32)          # truncate to 10 mantissa bits
        return hi, a - hi
    Xh, Xl = split_tf32(X)
    Wh, Wl = split_tf32(W64.float())
    torch.backends.cuda.matmul.allow_tf32 = True
    ms, Y = timed(lambda: Xh @ Wh + (Xh @ Wl + Xl @ Wh))
    report("direct DFT, 3 x TF32", ms, Y)
    ms, Y = timed(lambda: X @ W64.float())
    report("direct DFT, 1 x TF32 (inexact)", ms, Y)
    torch.backends.cuda.matmul.allow_tf32 = False

    # ---- direct DFT, fp16 hi/lo (fp32 accumulate and output) ----
    try:
        s = 1.0 / 1024.0                                                   # keep |x| inside the fp16 range
        Xs = X * s
        Xh16 = Xs.half()
        Xl16 = (Xs - Xh16.float()).half()
        Wh16 = W64.half()
        Wl16 = (W64 - Wh16.double()).half()
        f = lambda: (torch.mm(Xh16, Wh16, out_dtype=torch.float32) + (torch.mm(Xh16, Wl16, out_dtype=torch.float32) +
                                                                      torch.mm(Xl16, Wh16, out_dtype=torch.float32))) / s
        ms, Y = timed(f)
        report("direct DFT, fp16 hi/lo x 3", ms, Y)
    except Exception as e:                                                 # torch without mm(out_dtype=)
        print("direct DFT, fp16 hi/lo: not available here (%s)" % str(e).splitlines()[0][:80])

    # ---- factored 512 = 32 x 16 (n = 16 n1 + n2, k = k1 + 32 k2), TF32 x 3 per stage ----
    # stage 1: for every n2 (16 of them) a 32-point DFT over n1  -> [F, 16, 32] complex
    # twiddle: * exp(-2 pi i n2 k1 / 512);  stage 2: 16-point DFT over n2 -> X[k1 + 32 k2]
    n1 = torch.arange(32, device="cuda", dtype=torch.float64)
    k1 = torch.arange(32, device="cuda", dtype=torch.float64)
    a1 = 2 * math.pi * torch.outer(n1, k1) / 32
    C1, S1 = torch.cos(a1).float(), (-torch.sin(a1)).float()             # [32, 32]
    n2 = torch.arange(16, device="cuda", dtype=torch.float64)
    tw = 2 * math.pi * torch.outer(n2, k1) / 512
    Tc, Ts = torch.cos(tw).float(), (-torch.sin(tw)).float()             # [16, 32]
    k2 = torch.arange(16, device="cuda", dtype=torch.float64)
    a2 = 2 * math.pi * torch.outer(n2, k2) / 16
    C2, S2 = torch.cos(a2).float(), (-torch.sin(a2)).float()             # [16, 16]

    def mm3(a, b):
        ah, al = split_tf32(a)
        bh, bl = split_tf32(b)
        return ah @ bh + (ah @ bl + al @ bh)

    def factored():
        torch.backends.cuda.matmul.allow_tf32 = True
        x = X.view(F, 32, 16).transpose(1, 2).reshape(F * 16, 32)        # [F*16 (n2), 32 (n1)]  (real input)
        ar, ai = mm3(x, C1), mm3(x, S1)                                   # stage 1: [F*16, 32 (k1)]
        ar, ai = ar.view(F, 16, 32), ai.view(F, 16, 32)
        br, bi = ar * Tc - ai * Ts, ar * Ts + ai * Tc                     # twiddle
        br = br.transpose(1, 2).reshape(F * 32, 16)                       # [F*32 (k1), 16 (n2)]
        bi = bi.transpose(1, 2).reshape(F * 32, 16)
        yr = mm3(br, C2) - mm3(bi, S2)                                    # stage 2: [F*32, 16 (k2)]
        yi = mm3(br, S2) + mm3(bi, C2)
        torch.backends.cuda.matmul.allow_tf32 = False
        yr = yr.view(F, 32, 16).transpose(1, 2).reshape(F, 512)           # index k = k1 + 32 k2
        yi = yi.view(F, 32, 16).transpose(1, 2).reshape(F, 512)
        return torch.cat([yr[:, :257], yi[:, :257]], dim=1)
    ms, Y = timed(factored, n=3)
    report("factored 32 x 16, 3 x TF32 (torch)", ms, Y)


def mel_matrix(num_mel):
    """Kaldi mel banks (src/fbank.cc:103-163) as a dense [257 x num_mel] matrix, fp32 like the reference."""
    def mel_scale(f):
        return 1127.0 * math.log(1.0 + f / 700.0)
    lo, hi = mel_scale(20.0), mel_scale(8000.0)
    delta = (hi - lo) / (num_mel + 1)
    m = np.zeros((257, num_mel), np.float32)
    for b in range(num_mel):
        left, center, right = lo + b * delta, lo + (b + 1) * delta, lo + (b + 2) * delta
        for i in range(256):
            ml = mel_scale(16000.0 / 512 * i)
            if left < ml < right:
                m[i, b] = (ml - left) / (center - left) if ml <= center else (right - ml) / (right - center)
    return m


if __name__ == "__main__":
    main()
