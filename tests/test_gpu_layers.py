"""Every layer type of src/nnet.h:21-30 through ce_gpu_nnet: the literals of the reference's own
test/nnet_test.cc:37-224 (Splice with its edge clamp, Linear, Softmax, LogSoftmax, ReLU, Normalize,
BatchNorm, Narrow), and layer stacks tool/convert_am.py never emits (Normalize / Softmax, a Splice
without its Narrow) against the compiled reference (Nnet::Propagate, and the int8 composition).
Single-layer networks are loaded with prior = 1 (log prior 0), so the rows ce_gpu_nnet returns are
Layer::Propagate's.  Tolerance 1e-3: CheckEq of test/nnet_test.cc:22-24."""
import numpy as np
import pytest

from catears_b200 import api, formats as F

pytestmark = pytest.mark.gpu


def run_layers(tmp_path, layers, x, left=0, right=0, precision="fp32", num_out=None, keep_acc=-1):
    nnet, prior = str(tmp_path / "l.nnet"), str(tmp_path / "l.prior")
    F.write_nnet(nnet, layers, left, right)
    if num_out is None:
        num_out = x.shape[1]
        for l in layers:
            if l["type"] == F.SPLICE:
                num_out *= len(l["indices"])
            elif l["type"] == F.LINEAR:
                num_out = l["W"].shape[1]
    F.write_vector(prior, np.ones(num_out, np.float32))
    m = api.AcousticModelGpu(nnet=nnet, prior=prior, left_context=left, right_context=right,
                             precision=precision)
    assert m.feat_dim == x.shape[1] and m.num_pdfs == num_out
    if keep_acc >= 0:
        m.keep_acc(keep_acc)
    y, am = m.nnet(np.ascontiguousarray(x, np.float32))
    acc = m.get_acc(0) if keep_acc >= 0 else None
    m.close()
    return (y, am, acc, nnet, prior)


def close(a, b):
    return np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() < 1e-3


def test_splice_layer_with_clamp(tmp_path):
    """test/nnet_test.cc:37-57: SpliceLayer({-2, 1}) on 4 rows, clamped at both edges."""
    x = np.array([[1, 1], [2, 2], [3, 3], [4, 4]], np.float32)
    y = run_layers(tmp_path, [{"type": F.SPLICE, "indices": [-2, 1]}], x)[0]
    assert y.shape == (4, 4)
    assert np.array_equal(y, np.array([[1, 1, 2, 2], [1, 1, 3, 3], [1, 1, 4, 4], [2, 2, 4, 4]], np.float32))


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_linear_layer(tmp_path, precision):
    """test/nnet_test.cc:59-90 (the constructor takes W as [out x in]; the file holds [in x out])."""
    W = np.array([[0.1, 0.8, 0.9], [0.4, 0.2, 0.7], [0.2, 0.1, 0.1], [0.4, 0.3, 0.2]], np.float32)
    b = np.array([0.1, -0.1, 0.2, -0.2], np.float32)
    x = np.array([[0.3, -0.1, 0.9]], np.float32)
    y = run_layers(tmp_path, [{"type": F.LINEAR, "W": W.T.copy(), "b": b}], x, precision=precision)[0]
    assert close(y, [[0.86, 0.63, 0.34, 0.07]])


def test_softmax_layer(tmp_path):
    """test/nnet_test.cc:93-109."""
    x = np.array([[0.3, -0.1, 0.9, 0.2]], np.float32)
    y = run_layers(tmp_path, [{"type": F.SOFTMAX}], x)[0]
    assert close(y, [[0.2274135, 0.15243983, 0.41437442, 0.20577225]])


def test_logsoftmax_layer(tmp_path):
    """test/nnet_test.cc:112-134."""
    x = np.array([[0.6926, 0.5312, 0.3551], [0.1014, 0.4569, 0.6337], [0.5657, 0.8495, 0.8210],
                  [0.0483, 0.1684, 0.9234]], np.float32)
    y, am = run_layers(tmp_path, [{"type": F.LOGSOFTMAX}], x)[:2]
    assert close(y, [[-0.9418, -1.1032, -1.2793], [-1.4182, -1.0627, -0.8859], [-1.2862, -1.0024, -1.0309],
                     [-1.5100, -1.3899, -0.6349]])
    assert list(am) == [0, 2, 1, 2]


def test_relu_layer(tmp_path):
    """test/nnet_test.cc:136-152."""
    x = np.array([[0.3, -0.1, 0.9, 0.2]], np.float32)
    y = run_layers(tmp_path, [{"type": F.RELU}], x)[0]
    assert np.array_equal(y, np.array([[0.3, 0.0, 0.9, 0.2]], np.float32))


def test_normalize_layer(tmp_path):
    """test/nnet_test.cc:154-170: the squared sum of the row becomes its dimension."""
    x = np.array([[0.3, -0.1, 0.9, 0.2]], np.float32)
    y = run_layers(tmp_path, [{"type": F.NORMALIZE}], x)[0]
    assert abs(float((y.astype(np.float64) ** 2).sum()) - 4.0) < 1e-4
    want = x * np.float32(np.sqrt(4.0 / float((x.astype(np.float32) ** 2).sum())))
    assert close(y, want)


def test_batchnorm_layer(tmp_path):
    """test/nnet_test.cc:172-193."""
    p = np.array([0.1, 0.2, 0.3], np.float32)
    x = np.array([[0.1, 0.1, 0.1], [0.2, 0.2, 0.2]], np.float32)
    y = run_layers(tmp_path, [{"type": F.BATCHNORM, "scale": p, "offset": p}], x)[0]
    assert close(y, [[0.11, 0.22, 0.33], [0.12, 0.24, 0.36]])


def test_narrow_layer(tmp_path):
    """test/nnet_test.cc:195-224: NarrowLayer(1, 2) keeps rows [1, rows - 2).  Through the AM the
    matrix is the replicate-padded one (src/am.cc:119-124,152-155), so with contexts (1, 2) the kept
    rows are the input frames themselves."""
    W = np.array([[0.1, 0.8, 0.9], [0.4, 0.2, 0.7], [0.2, 0.1, 0.1], [0.4, 0.3, 0.2], [0.5, 0.6, 0.7]], np.float32)
    y = run_layers(tmp_path, [{"type": F.NARROW, "left": 1, "right": 2}], W, left=1, right=2)[0]
    assert np.array_equal(y, W)
    with pytest.raises(api.CeGpuError, match="does not match the rows the nnet removes"):
        run_layers(tmp_path, [{"type": F.NARROW, "left": 1, "right": 2}], W, left=0, right=0)


def general_stack(rng, dim=24):
    """Splice without Narrow (edge clamp live), Normalize, Softmax in the middle, BatchNorm alone."""
    h = 40
    return [
        {"type": F.SPLICE, "indices": [-2, 0, 1]},
        {"type": F.LINEAR, "W": (rng.standard_normal((3 * dim, h)) / np.sqrt(3 * dim)).astype(np.float32),
         "b": (0.1 * rng.standard_normal(h)).astype(np.float32)},
        {"type": F.RELU},
        {"type": F.NORMALIZE},
        {"type": F.SPLICE, "indices": [-1, 1]},
        {"type": F.NARROW, "left": 1, "right": 1},
        {"type": F.BATCHNORM, "scale": rng.uniform(0.5, 1.5, 2 * h).astype(np.float32),
         "offset": (0.1 * rng.standard_normal(2 * h)).astype(np.float32)},
        {"type": F.LINEAR, "W": (rng.standard_normal((2 * h, 16)) / np.sqrt(2 * h)).astype(np.float32),
         "b": (0.1 * rng.standard_normal(16)).astype(np.float32)},
        {"type": F.SOFTMAX},
        {"type": F.LINEAR, "W": rng.standard_normal((16, 12)).astype(np.float32),
         "b": (0.1 * rng.standard_normal(12)).astype(np.float32)},
        {"type": F.LOGSOFTMAX},
    ]


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16x3", 1e-3)])
def test_general_stack_vs_reference_propagate(tmp_path, ref, precision, tol):
    """Nnet::Propagate of the compiled reference on the replicate-padded matrix of one utterance,
    against ce_gpu_nnet on a ragged batch of three."""
    if ref is None:
        pytest.skip("oracle/_ref was never built")
    rng = np.random.default_rng(77)
    layers = general_stack(rng)
    sizes = [5, 37, 1]
    x = rng.standard_normal((sum(sizes), 24)).astype(np.float32)
    off = np.concatenate([[0], np.cumsum(sizes)])
    nnet, prior = str(tmp_path / "g.nnet"), str(tmp_path / "g.prior")
    F.write_nnet(nnet, layers, 1, 1)
    F.write_vector(prior, np.ones(12, np.float32))
    m = api.AcousticModelGpu(nnet=nnet, prior=prior, left_context=1, right_context=1, precision=precision)
    y, am = m.nnet(x, off)
    m.close()
    for u, T in enumerate(sizes):
        xu = x[off[u]:off[u + 1]]
        padded = np.concatenate([xu[:1], xu, xu[-1:]])
        want = ref.nnet_propagate(nnet, padded)
        assert want.shape == (T, 12)
        assert np.abs(y[off[u]:off[u + 1]] - want).max() < tol, (u, np.abs(y[off[u]:off[u + 1]] - want).max())


def test_general_stack_int8_vs_reference_composition(tmp_path, ref):
    """int8: Quantize(in) + MatMat_U8U8F32 + AddVec(b) per Linear layer with the
    reference's own classes for every other layer.  The first Linear layer reads spliced copies of
    the input, so its accumulators are bit-exact; later layers sit behind Normalize / Softmax
    (another summation order), so the end result is compared at the quantisation error budget."""
    if ref is None:
        pytest.skip("oracle/_ref was never built")
    rng = np.random.default_rng(78)
    layers = general_stack(rng)
    x = rng.standard_normal((50, 24)).astype(np.float32)
    y, am, acc, nnet, prior = run_layers(tmp_path, layers, x, left=1, right=1, precision="int8", num_out=12,
                                         keep_acc=0)
    want, wacc = ref.u8_forward(nnet, prior, 1, 1, x, dump_layer=0)
    assert acc.shape == wacc.shape == (52, 40)
    assert np.array_equal(acc, wacc)
    assert np.abs(y - want).max() < 0.02 * max(1.0, float(want.max() - want.min()))
