"""The output layer fused with LogSoftmax + prior + argmax (GemmArgs::lsm, catears_b200/csrc/gemm.cu) against the
same layer followed by the separate log-softmax kernel (CE_GPU_FUSED_OUTPUT=0): same accumulators, same fp32
chain up to the logit (src/nnet.cc:34, eight_bit_int_gemm.cc:389), so the rows may differ only by the order in
which the row's exponentials are summed (src/vector.cc:110-122) -- a few 1e-7 of the log-sum -- and the argmax
only where the two best entries are a rounding tie.  Both are compared with the oracle as well."""
import os

import numpy as np
import pytest

from catears_b200 import api, formats as F

pytestmark = pytest.mark.gpu


def load(conf, precision, fused):
    old = os.environ.get("CE_GPU_FUSED_OUTPUT")
    os.environ["CE_GPU_FUSED_OUTPUT"] = str(fused)         # read when the model is loaded
    try:
        return api.AcousticModelGpu(config=conf, precision=precision)
    finally:
        if old is None:
            del os.environ["CE_GPU_FUSED_OUTPUT"]
        else:
            os.environ["CE_GPU_FUSED_OUTPUT"] = old


def ragged_batch(rng, lengths, dim=40):
    off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    return (rng.standard_normal((int(off[-1]), dim)) * 2.0).astype(np.float32), off


@pytest.mark.parametrize("precision,fused", [("int8", 1), ("fp32", 2), ("bf16", 2), ("bf16x3", 2)])
def test_fused_equals_separate(small_model, port, precision, fused):
    """Ragged batches (1 ... 300 frames: one block, several 128-row tiles, the packed 32-row layout of short
    blocks), every precision's instantiation of the fused kernel."""
    rng = np.random.default_rng(11)
    a = load(small_model["conf"], precision, fused)
    b = load(small_model["conf"], precision, 0)
    prior = F.read_vector(small_model["prior"])
    try:
        for lengths in ([57], [1, 2, 3, 200, 5], [300, 129, 128, 127, 64, 31, 33], [7] * 40):
            x, off = ragged_batch(rng, lengths)
            la, aa = a.nnet(x, off)
            lb, ab = b.nnet(x, off)
            assert la.shape == lb.shape == (int(off[-1]), 96)
            assert np.abs(la - lb).max() < 5e-6, (precision, lengths, np.abs(la - lb).max())
            top2 = np.sort(lb, axis=1)[:, -2:]
            clear = (top2[:, 1] - top2[:, 0]) > 1e-5
            assert np.array_equal(aa[clear], ab[clear])
            assert np.array_equal(aa, la.argmax(axis=1))   # the argmax is the first maximum of the row it wrote
        if precision in ("int8", "fp32"):
            x, off = ragged_batch(rng, [61])
            want = port.am_forward(small_model["nnet"], prior, 13, 13, x, mode="u8" if precision == "int8" else "float")
            got, _ = a.nnet(x, off)
            assert np.abs(got - want).max() < (1e-5 if precision == "int8" else 1e-3)
    finally:
        a.close()
        b.close()


def test_fused_argmax_only_and_selected_rows(small_model):
    """loglik = None (argmax only: the second sweep writes nothing) and the selecting outputs (ONE sweep of the
    fused layer writes plain logits in row space plus every row's log-sum-exp, the selection kernels subtract
    that very number): bit-identical to the dense rows."""
    rng = np.random.default_rng(12)
    m = load(small_model["conf"], "int8", 1)
    try:
        x, off = ragged_batch(rng, [150, 3, 77])
        ll, am = m.nnet(x, off)
        _, am_only = m.nnet(x, off, want_loglik=False)
        assert np.array_equal(am, am_only)
        m.set_output("topk", k=8)
        best, _ = m.nnet(x, off)
        order = np.argsort(-ll, axis=1, kind="stable")[:, :8]
        assert best.shape == (len(ll), 8)                   # SCORED_PDF entries
        assert np.array_equal(best[best.dtype.names[0]], np.take_along_axis(ll, order, axis=1))
        assert np.array_equal(best[best.dtype.names[1]], order.astype(np.int32))
        ids = np.array([5, 0, 95, 17], np.int32)
        m.set_output("subset", pdf_ids=ids)
        sub, _ = m.nnet(x, off)
        assert np.array_equal(sub, ll[:, ids])
    finally:
        m.close()


@pytest.mark.parametrize("num_pdfs", [4, 52, 100, 260, 516, 98])
def test_fused_ragged_widths(tmp_path, port, num_pdfs):
    """Output widths that end inside a 16-column piece, inside a 64-column part and just behind a 256-column
    tile (the last piece's masks, the TMA store's column clipping, warps with nothing to do); 98 is not a
    multiple of 4 and takes the separate kernel.  int8, against the oracle and against the separate kernel."""
    from catears_b200 import synth
    m = synth.write_model(str(tmp_path / "w"), name="w%d" % num_pdfs, hidden=64, num_pdfs=num_pdfs, seed=500 + num_pdfs)
    prior = F.read_vector(m["prior"])
    rng = np.random.default_rng(num_pdfs)
    a = load(m["conf"], "int8", 1)
    b = load(m["conf"], "int8", 0)
    try:
        x, off = ragged_batch(rng, [70, 1, 140])
        la, aa = a.nnet(x, off)
        lb, ab = b.nnet(x, off)
        assert la.shape == (211, num_pdfs)
        assert np.abs(la - lb).max() < 5e-6
        assert np.array_equal(aa, la.argmax(axis=1))
        want = port.am_forward(m["nnet"], prior, m["left"], m["right"], x[:70], mode="u8")
        assert np.abs(la[:70] - want).max() < 1e-5
        if num_pdfs >= 8 and num_pdfs % 4 == 0:             # (the selecting outputs need num_pdfs % 4 == 0)
            a.set_output("topk", k=3)
            best, _ = a.nnet(x, off)
            order = np.argsort(-la, axis=1, kind="stable")[:, :3]
            assert np.array_equal(best["pdf"], order.astype(np.int32))
            assert np.array_equal(best["loglik"], np.take_along_axis(la, order, axis=1))
    finally:
        a.close()
        b.close()
