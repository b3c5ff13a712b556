#!/usr/bin/env python
"""Generates tests/golden/* from the reference tree.  Run HERE (where
/root/reference exists), after `make -C oracle ref`:

    python tests/golden/make_golden.py

What it writes (all small, all committed):
  * verbatim DATA fixtures of the reference's own tests (no source code):
      en-us-hello.wav, en-us-cat.wav, cmvn_stats.bin,
      fbankmat_en-us-hello.wav.txt, fbankcmvnmat_en-us-hello.wav.txt   (test/data/)
  * srfft_kat.npz: the 128-point known-answer vector of test/srfft_test.cc:13-273
    (numbers extracted from the two float arrays; input at :144-273, expected at :13-142)
  * ref_vectors.npz: outputs of the UNMODIFIED reference (oracle/_ref/libce_ref*.so) on
    seeded inputs: fbank, online CMVN past the 600-frame window, a small TDNN in the
    float path (in-order sgemm) and in the int8 composition (log-likelihoods, accumulators).
"""
import os
import re
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("CE_REFERENCE", "/root/reference")

from catears_b200 import synth  # noqa: E402
from oracle.ref import Ref  # noqa: E402


def small_model(tmp):
    """A TDNN with the bench network's structure at 1/16 width (hidden 64, 96 pdfs)."""
    return synth.write_model(tmp, name="small", hidden=64, num_pdfs=96, seed=4321)


def main():
    for name in ("en-us-hello.wav", "en-us-cat.wav", "cmvn_stats.bin",
                 "fbankmat_en-us-hello.wav.txt", "fbankcmvnmat_en-us-hello.wav.txt"):
        shutil.copyfile(os.path.join(REF, "test", "data", name), os.path.join(HERE, name))

    src = open(os.path.join(REF, "test", "srfft_test.cc")).read()
    arrays = re.findall(r"=\s*\{([^}]*)\}", src)
    vals = [np.array([float(t.rstrip("f")) for t in re.findall(r"[-+0-9.eE]+f?", a) if t.strip("f")],
                     np.float32) for a in arrays]
    vals = [v for v in vals if v.size == 128]
    assert len(vals) == 2, [v.size for v in vals]
    # first array in the file = expected output (:13-142), second = input (:144-273)
    np.savez(os.path.join(HERE, "srfft_kat.npz"), expected=vals[0], input=vals[1])

    ref = Ref()
    out = {}
    import wave
    for wav in ("en-us-hello", "en-us-cat"):
        w = wave.open(os.path.join(HERE, wav + ".wav"))
        pcm = np.frombuffer(w.readframes(w.getnframes()), np.int16)
        out["fbank40_" + wav] = ref.fbank(pcm)
    pcm = synth.synth_utterance(0, 16000)          # 1 s of config-2 audio
    out["fbank40_synth0_1s"] = ref.fbank(pcm)

    # CMVN beyond the sliding window: 700 frames of N(14, 3^2) features, bundled global stats.
    g = np.frombuffer(open(os.path.join(HERE, "cmvn_stats.bin"), "rb").read()[12:], np.float32)
    rng = np.random.default_rng(99)
    feats = (14.0 + 3.0 * rng.standard_normal((700, 40))).astype(np.float32)
    out["cmvn_in_700"] = feats
    out["cmvn_out_700"] = ref.cmvn(g, feats)

    with tempfile.TemporaryDirectory() as tmp:
        m = small_model(tmp)
        x = rng.standard_normal((57, 40)).astype(np.float32)
        out["am_in"] = x
        ref.set_sgemm("inorder")
        out["am_float"] = ref.am_forward(m["conf"], x)
        y, acc = ref.u8_forward(m["nnet"], m["prior"], m["left"], m["right"], x, dump_layer=1)
        out["am_u8"] = y
        out["am_u8_acc_linear1"] = acc
        _, acc6 = ref.u8_forward(m["nnet"], m["prior"], m["left"], m["right"], x, dump_layer=6)
        out["am_u8_acc_linear6"] = acc6

    a = rng.uniform(-0.5, 0.5, (37, 70)).astype(np.float32)   # ranges of test/gemm_test.cc:97-98
    b = rng.uniform(1.0, 2.0, (70, 24)).astype(np.float32)
    qa, sa, za = ref.quantize(a)
    qb, sb, zb = ref.quantize(b)
    c, acc = ref.gemm_u8(qa, sa, za, qb, sb, zb)
    out.update(q_a=a, q_b=b, q_a8=qa, q_b8=qb, q_params=np.array([sa, za, sb, zb], np.float64),
               q_c=c, q_acc=acc)
    neg = -np.abs(a) - 1.0                                    # all-negative: FLT_MIN quirk (Q9)
    qn, sn, zn = ref.quantize(neg)
    out.update(q_neg=neg, q_neg8=qn, q_neg_params=np.array([sn, zn], np.float64))

    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
