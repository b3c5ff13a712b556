"""Parity of the BENCHMARKED network -- the 6 x 1024 TDNN with 3072 pdfs of SURVEY 8d (seed 1234,
the model bench.py times) -- against the unmodified reference compiled into oracle/_ref, on a 1 s
and a 10 s utterance:

  int8   the activation QuantizationParams in front of every Linear layer and the int32 accumulators
         of ALL seven Linear layers bit-exact (src/matrix.cc:348-420, gemmlowp unpack.h:118-125);
         log-likelihoods to 1e-5 (only the log-sum-exp summation order differs); argmax identical
         (first maximum wins) wherever the reference's own top-2 margin is not an fp32 rounding tie;
  float  fp32-class paths (3xTF32, bf16x3) within the north_star's 1e-3 absolute of the reference's
         float nnet; single-pass TF32 / bf16 are reported with their measured error (they are fast
         modes OUTSIDE the tolerance and labelled so);
  PCM    from int16 samples through the GPU fbank + CMVN + int8 AM against the reference pipeline:
         the per-frame argmax agreement rate is printed and bounded.
"""
import os

import numpy as np
import pytest

from catears_b200 import api, formats as F, synth

pytestmark = pytest.mark.gpu

LEFT = RIGHT = 13


@pytest.fixture(scope="module")
def tdnn(tmp_path_factory):
    d = tmp_path_factory.mktemp("tdnn_full")
    stats = synth.default_cmvn_stats()
    m = synth.write_model(str(d), name="tdnn", cmvn_stats=stats)
    m["stats"] = stats
    return m


@pytest.fixture(scope="module")
def ref_sse4():
    """The reference built with -msse4.1 (gemmlowp's SSE4 kernel; integer results identical)."""
    from oracle import ref as R
    variant = "_sse4" if R.available("_sse4") else ""
    if not R.available(variant):
        pytest.skip("oracle/_ref was never built")
    r = R.Ref(variant)
    if not hasattr(r.L, "ref_u8_forward_trace"):
        pytest.skip("oracle/_ref predates ref_u8_forward_trace: rebuild with `make -C oracle ref ref-sse4`")
    return r


@pytest.fixture(scope="module")
def utterances(port, tdnn):
    """(name, pcm, CMVN'd features of the ORACLE front end) for a 1 s and a 10 s utterance."""
    out = []
    for name, u, n in (("1s", 3, 16000), ("10s", 5, 160000)):
        pcm = synth.synth_utterance(u, n)
        feats = port.cmvn(tdnn["stats"], port.fbank(pcm))
        out.append((name, pcm, feats))
    return out


def argmax_agreement(got, ref_ll, tie=2e-6):
    """Fraction of frames with the reference's argmax, and whether every disagreement is a frame
    whose top-2 margin in the reference is within `tie` (relative to the magnitude)."""
    want = ref_ll.argmax(axis=1)
    top2 = np.sort(ref_ll, axis=1)[:, -2:]
    margin = top2[:, 1] - top2[:, 0]
    clear = margin > tie * np.maximum(1.0, np.abs(top2[:, 1]))
    return float(np.mean(got == want)), bool(np.array_equal(got[clear], want[clear]))


def test_int8_every_layer_bit_exact(tdnn, ref_sse4, utterances):
    am = api.AcousticModelGpu(nnet=tdnn["nnet"], prior=tdnn["prior"], left_context=LEFT,
                              right_context=RIGHT, precision="int8")
    for name, _, feats in utterances:
        T = feats.shape[0]
        want_ll, want_q, want_acc = ref_sse4.u8_forward_trace(tdnn["nnet"], tdnn["prior"], LEFT, RIGHT, feats)
        assert len(want_acc) == 7 and want_ll.shape == (T, 3072)
        ll = arg = None
        for ordinal in range(7):
            am.keep_acc(ordinal)
            ll, arg = am.nnet(feats)
            acc = am.get_acc(0)
            # the reference's matrix at this layer holds the rows the cumulative Narrow left over
            assert acc.shape == want_acc[ordinal].shape, (name, ordinal, acc.shape)
            assert np.array_equal(acc, want_acc[ordinal]), (name, ordinal)
        sc, zp = am.get_qparams(0)
        assert [int(z) for z in zp] == [q[1] for q in want_q], name
        assert np.array_equal(sc, np.array([q[0] for q in want_q], np.float32)), name
        am.keep_acc(-1)
        err = float(np.abs(ll - want_ll).max())
        rate, clear_ok = argmax_agreement(arg, want_ll)
        print("int8 %s: %d frames, 7/7 accumulator matrices and qparams bit-exact, loglik max err %.2e, "
              "argmax agreement %.5f" % (name, T, err, rate))
        assert err < 1e-5
        assert clear_ok and rate > 0.999
    am.close()


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16x3", 1e-3), ("tf32", 2e-2), ("bf16", 0.15)])
def test_float_paths_vs_reference_float_nnet(tdnn, ref, utterances, precision, tol):
    """north_star: float log-likelihoods within 1e-3 absolute.  fp32 (3xTF32) and bf16x3 (bf16 hi/lo,
    three products) are the in-tolerance paths; tf32 / bf16 single pass are fast modes whose error
    is printed and only sanity-bounded."""
    if ref is None:
        pytest.skip("oracle/_ref was never built")
    am = api.AcousticModelGpu(nnet=tdnn["nnet"], prior=tdnn["prior"], left_context=LEFT,
                              right_context=RIGHT, precision=precision)
    for name, _, feats in utterances:
        if name == "10s" and not ref.set_sgemm("openblas"):
            ref.set_sgemm("inorder")
        want = ref.am_forward(tdnn["conf"], feats)
        ref.set_sgemm("inorder")
        ll, arg = am.nnet(feats)
        err = float(np.abs(ll - want).max())
        rate, _ = argmax_agreement(arg, want)
        print("%s %s: loglik max abs err %.3e (bar %.0e), argmax agreement %.4f" % (precision, name, err, tol, rate))
        assert err < tol, (precision, name, err)
    am.close()


def test_pcm_to_argmax_agreement_int8(tdnn, ref_sse4, utterances):
    """The whole path from PCM.  The GPU fbank differs from the reference's by ~1e-6 relative (another
    FFT algorithm), so a few u8 codes of the first layer may move by one step; reported: the fraction
    of frames whose argmax pdf equals the reference pipeline's, and the log-likelihood error."""
    am = api.AcousticModelGpu(config=tdnn["conf"], precision="int8")
    for name, pcm, _ in utterances:
        feats = ref_sse4.cmvn(tdnn["stats"], ref_sse4.fbank(pcm))
        want_ll, _, _ = ref_sse4.u8_forward_trace(tdnn["nnet"], tdnn["prior"], LEFT, RIGHT, feats, want_acc=False)
        ll, arg, _ = am.forward(pcm)
        rate, _ = argmax_agreement(arg, want_ll)
        err = float(np.abs(ll - want_ll).max())
        print("PCM->int8 %s: argmax agreement %.5f over %d frames, loglik max abs err %.3e"
              % (name, rate, ll.shape[0], err))
        assert rate > 0.98
        assert err < 0.01 * float(want_ll.max() - want_ll.min())
    am.close()
