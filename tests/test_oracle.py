"""Pins the CPU oracle (oracle/ce_oracle.c) before anything is compared with it:
against the reference's own known answers and Kaldi dumps (tests/golden), against the
literals of the reference's unit tests, and -- where oracle/_ref was built -- against the
reference itself.  CPU only."""
import numpy as np
import pytest

from catears_b200 import formats as F
from catears_b200 import synth


# -- FFT --------------------------------------------------------------------

def test_srfft_known_answer(golden, port):
    """test/srfft_test.cc:275-288: 128-point forward real FFT, |diff| < 1e-4."""
    out = port.srfft(golden["srfft"]["input"])
    assert np.abs(out - golden["srfft"]["expected"]).max() < 1e-4


def test_srfft_is_dft(port):
    rng = np.random.default_rng(0)
    x = rng.standard_normal(512).astype(np.float32)
    out = port.srfft(x)
    X = np.fft.rfft(x.astype(np.float64))
    assert abs(out[0] - X[0].real) < 1e-3 and abs(out[1] - X[256].real) < 1e-3
    got = out[2::2] + 1j * out[3::2]
    assert np.abs(got - X[1:256]).max() < 2e-3


def test_srfft_vs_reference(port, ref):
    if ref is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(1)
    for n in (128, 512):
        x = (1000 * rng.standard_normal(n)).astype(np.float32)
        a, b = port.srfft(x), ref.srfft(x)
        assert np.abs(a - b).max() <= 2e-6 * np.abs(b).max()


# -- fbank ------------------------------------------------------------------

def test_fbank_kaldi_golden(golden, port):
    """test/fbank_test.cc:24-60: en-us-hello.wav -> 47x40 vs Kaldi, < 1e-4 abs."""
    fb = port.fbank(golden["hello_pcm"])
    assert fb.shape == (47, 40)
    assert np.abs(fb - golden["kaldi_fbank"]).max() < 1e-4


def test_fbank_vs_reference_vectors(golden, port):
    for key, pcm in (("fbank40_en-us-hello", golden["hello_pcm"]),
                     ("fbank40_en-us-cat", golden["cat_pcm"]),
                     ("fbank40_synth0_1s", synth.synth_utterance(0, 16000))):
        want = golden["ref"][key]
        got = port.fbank(pcm)
        assert got.shape == want.shape
        assert np.abs(got - want).max() < 1e-4 * max(1.0, np.abs(want).max())


def test_fbank_tables_vs_reference_live(port, ref):
    """Hamming/mel tables are fp32-exact restatements: identical fbank on silence-free
    noise to within FFT rounding, and frame counts follow snip-edges (fbank.cc:35-42)."""
    if ref is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)
    for n in (400, 559, 560, 1234, 16000):
        pcm = rng.integers(-20000, 20000, n).astype(np.int16)
        a, b = port.fbank(pcm), ref.fbank(pcm)
        assert a.shape == b.shape == (synth.num_frames(n), 40)
        assert np.abs(a - b).max() < 5e-5
    assert port.fbank(np.zeros(399, np.int16)).shape[0] == 0


def test_fbank_edge_inputs(port, ref):
    """All-zero audio hits the FLT_EPSILON floor (fbank.cc:243); full-scale audio stays finite."""
    z = port.fbank(np.zeros(800, np.int16))
    assert np.allclose(z, np.log(np.float32(1.1920929e-7)))
    big = np.full(800, 32767, np.int16)
    big[::2] = -32768
    assert np.isfinite(port.fbank(big)).all()
    if ref is not None:
        # A full-scale Nyquist tone puts e^13 (6e5x) between the top and bottom mel bins: the
        # weakest bin is fp32 leakage of the strongest, so two FFT algorithms agree only to
        # ~1e-7 of the PEAK energy there (1e-4 in the log), and to fp32 rounding elsewhere.
        a, b = port.fbank(big), ref.fbank(big)
        assert np.abs(np.exp(a) - np.exp(b)).max() < 1e-6 * np.exp(b).max()
        assert np.abs(a - b).max() < 5e-4


def test_mel_filter_shape(port):
    """492 non-zero weights for 40 bins; first filter [1,4), last [225,256)."""
    filt = port.mel_filters(40)
    assert filt[0][0] == 1 and len(filt[0][1]) == 3
    assert filt[39][0] == 225 and len(filt[39][1]) == 31
    assert sum(len(w) for _, w in filt) == 492


# -- CMVN -------------------------------------------------------------------

def test_cmvn_kaldi_golden(golden, port):
    """test/cmvn_test.cc:38-79 (stale in the reference, restated here two-sided)."""
    fb = port.fbank(golden["hello_pcm"])
    out = port.cmvn(golden["cmvn_stats"], fb)
    assert np.abs(out - golden["kaldi_cmvn"]).max() < 1e-4


def test_cmvn_bit_exact_vs_reference(golden, port, ref):
    want = golden["ref"]["cmvn_out_700"]
    got = port.cmvn(golden["cmvn_stats"], golden["ref"]["cmvn_in_700"])
    assert np.array_equal(got, want)          # past the 600-frame window, bit for bit
    if ref is not None:
        rng = np.random.default_rng(5)
        x = (10 + 5 * rng.standard_normal((1500, 40))).astype(np.float32)
        assert np.array_equal(port.cmvn(golden["cmvn_stats"], x), ref.cmvn(golden["cmvn_stats"], x))


# -- nnet layers: the literals of test/nnet_test.cc ---------------------------

def _run_layers(port, tmp_path, layers, x, mode="float"):
    p = str(tmp_path / "net.nnet")
    F.write_nnet(p, layers, 0, 0)
    return port.nnet_propagate(p, np.asarray(x, np.float32), mode)


def test_nnet_literals(port, tmp_path):
    eq = lambda a, b: np.abs(np.asarray(a) - np.asarray(b, np.float32)).max() < 1e-3  # :23-25
    # Splice {-2, 1} with clamping (:37-57)
    y = _run_layers(port, tmp_path, [{"type": F.SPLICE, "indices": [-2, 1]}],
                    [[1, 1], [2, 2], [3, 3], [4, 4]])
    assert eq(y, [[1, 1, 2, 2], [1, 1, 3, 3], [1, 1, 4, 4], [2, 2, 4, 4]])
    # Linear 3->4 (:59-90); the test passes W as [out x in], the file stores [in x out]
    W = np.array([[0.1, 0.8, 0.9], [0.4, 0.2, 0.7], [0.2, 0.1, 0.1], [0.4, 0.3, 0.2]], np.float32)
    y = _run_layers(port, tmp_path, [{"type": F.LINEAR, "W": W.T.copy(),
                                      "b": [0.1, -0.1, 0.2, -0.2]}], [[0.3, -0.1, 0.9]])
    assert eq(y, [[0.86, 0.63, 0.34, 0.07]])
    # Softmax (:93-109)
    y = _run_layers(port, tmp_path, [{"type": F.SOFTMAX}], [[0.3, -0.1, 0.9, 0.2]])
    assert eq(y, [[0.2274135, 0.15243983, 0.41437442, 0.20577225]])
    # LogSoftmax 4x3 (:112-133)
    y = _run_layers(port, tmp_path, [{"type": F.LOGSOFTMAX}],
                    [[0.6926, 0.5312, 0.3551], [0.1014, 0.4569, 0.6337],
                     [0.5657, 0.8495, 0.8210], [0.0483, 0.1684, 0.9234]])
    assert eq(y, [[-0.9418, -1.1032, -1.2793], [-1.4182, -1.0627, -0.8859],
                  [-1.2862, -1.0024, -1.0309], [-1.5100, -1.3899, -0.6349]])
    # ReLU (:135-151)
    assert eq(_run_layers(port, tmp_path, [{"type": F.RELU}], [[0.3, -0.1, 0.9, 0.2]]),
              [[0.3, 0.0, 0.9, 0.2]])
    # Normalize: |y|^2 == D (:154-170)
    y = _run_layers(port, tmp_path, [{"type": F.NORMALIZE}], [[0.3, -0.1, 0.9, 0.2]])
    assert abs(float((y.astype(np.float64) ** 2).sum()) - 4.0) < 1e-4
    # BatchNorm (:172-193)
    y = _run_layers(port, tmp_path, [{"type": F.BATCHNORM, "scale": [0.1, 0.2, 0.3],
                                      "offset": [0.1, 0.2, 0.3]}],
                    [[0.1, 0.1, 0.1], [0.2, 0.2, 0.2]])
    assert eq(y, [[0.11, 0.22, 0.33], [0.12, 0.24, 0.36]])
    # Narrow(1,2), incl. the "too few rows -> passthrough" branch (:195-224)
    Wd = [[0.1, 0.8, 0.9], [0.4, 0.2, 0.7], [0.2, 0.1, 0.1], [0.4, 0.3, 0.2], [0.5, 0.6, 0.7]]
    nar = [{"type": F.NARROW, "left": 1, "right": 2}]
    assert eq(_run_layers(port, tmp_path, nar, Wd), Wd[1:3])
    assert eq(_run_layers(port, tmp_path, nar, Wd[:3]), Wd[:3])


def test_nnet_file_roundtrip(small_model):
    layers, left, right = F.read_nnet(small_model["nnet"])
    want, wl, wr, prior = synth.tdnn_layers(hidden=64, num_pdfs=96, seed=4321)
    assert (left, right) == (wl, wr) == (13, 13)
    assert [l["type"] for l in layers] == [l["type"] for l in want]
    assert np.array_equal(layers[2]["W"], want[2]["W"])
    assert np.array_equal(F.read_vector(small_model["prior"]), prior)
    assert synth.flops_per_frame(synth.tdnn_layers()[0]) == 38158336


# -- quantisation + u8 GEMM ----------------------------------------------------

def test_quantize_and_gemm_u8_vs_reference_vectors(golden, port):
    g = golden["ref"]
    qa, sa, za = port.quantize(g["q_a"])
    qb, sb, zb = port.quantize(g["q_b"])
    assert np.array_equal(qa, g["q_a8"]) and np.array_equal(qb, g["q_b8"])
    assert [float(sa), za, float(sb), zb] == list(g["q_params"])
    c, acc = port.gemm_u8(qa, sa, za, qb, sb, zb)
    assert np.array_equal(acc, g["q_acc"])
    assert np.array_equal(c, g["q_c"])
    # all-negative matrix: max is seeded with FLT_MIN (matrix.cc:330-331)
    qn, sn, zn = port.quantize(g["q_neg"])
    assert np.array_equal(qn, g["q_neg8"]) and [float(sn), zn] == list(g["q_neg_params"])


def test_gemm_u8_error_budget(port):
    """test/gemm_test.cc:106-120: quantised product within 1% of the value range."""
    rng = np.random.default_rng(11)
    for m, n, k in ((5, 3, 2), (100, 100, 1), (121, 233, 17), (64, 64, 80)):
        a = rng.uniform(-0.5, 0.5, (m, k)).astype(np.float32)
        b = rng.uniform(1, 2, (k, n)).astype(np.float32)
        qa, sa, za = port.quantize(a)
        qb, sb, zb = port.quantize(b)
        c, _ = port.gemm_u8(qa, sa, za, qb, sb, zb)
        cref = port.sgemm(a, b)
        assert np.abs(c - cref).max() / (cref.max() - cref.min()) < 0.01


# -- AM -----------------------------------------------------------------------

def test_am_float_vs_reference_vectors(golden, port, small_model):
    g = golden["ref"]
    prior = F.read_vector(small_model["prior"])
    y = port.am_forward(small_model["nnet"], prior, 13, 13, g["am_in"], "float")
    assert y.shape == g["am_float"].shape == (57, 96)
    assert np.abs(y - g["am_float"]).max() < 2e-5      # same summation order; expf/logf only
    assert np.allclose(np.exp(y + np.log(prior)).sum(1), 1.0, atol=1e-4)


def test_am_u8_vs_reference_vectors(golden, port, small_model):
    g = golden["ref"]
    prior = F.read_vector(small_model["prior"])
    y, acc1 = port.am_forward(small_model["nnet"], prior, 13, 13, g["am_in"], "u8", 1,
                              g["am_u8_acc_linear1"].shape)
    assert np.array_equal(acc1, g["am_u8_acc_linear1"])          # int32 accumulators, exact
    _, acc6 = port.am_forward(small_model["nnet"], prior, 13, 13, g["am_in"], "u8", 6,
                              g["am_u8_acc_linear6"].shape)
    assert np.array_equal(acc6, g["am_u8_acc_linear6"])
    assert np.array_equal(y.argmax(1), g["am_u8"].argmax(1))
    assert np.abs(y - g["am_u8"]).max() < 2e-5


def test_am_streaming_equals_whole(ref, port, small_model, tmp_path):
    """chunked streaming through AcousticModel::Process == one whole-utterance
    batch (valid dilated convolution + replicate padding), up to sgemm blocking."""
    if ref is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(21)
    x = rng.standard_normal((83, 40)).astype(np.float32)
    prior = F.read_vector(small_model["prior"])
    whole = port.am_forward(small_model["nnet"], prior, 13, 13, x, "float")
    conf = str(tmp_path / "chunk8.conf")
    F.write_am_config(conf, small_model["nnet"], small_model["prior"], 13, 13, 8, 96,
                      small_model["tid2pdf"])
    import shutil
    for k in ("nnet", "prior", "tid2pdf"):
        shutil.copy(small_model[k], str(tmp_path))
    streamed = ref.am_forward(conf, x)
    assert streamed.shape == whole.shape
    assert np.abs(streamed - whole).max() < 2e-5


def test_independent_fp64_fbank_pins_40_and_80_bins(golden, port):
    """tests/kaldi_fp64.py (numpy float64, numpy's FFT, no code shared with the port) reproduces the
    Kaldi golden dump at 40 bins to 1e-4 absolute, and the C restatement agrees with it at 40 and at
    80 bins -- the size the reference cannot run (src/fbank.cc:155) and that no reference fixture
    covers, so this independent statement of the one-parameter formula is its pin."""
    import kaldi_fp64
    from catears_b200 import synth
    f40 = kaldi_fp64.fbank(golden["hello_pcm"], 40)
    assert f40.shape == golden["kaldi_fbank"].shape
    assert np.abs(f40 - golden["kaldi_fbank"]).max() < 1e-4
    w = kaldi_fp64.mel_weights(40)
    nz = (w > 0).sum(axis=1)
    assert int(nz.sum()) == 492 and int(nz.max()) == 31          # SURVEY 8a row 8 [probe]
    assert w[:, 0].sum() == 0 and w[:, 256].sum() == 0
    for pcm in (golden["hello_pcm"], golden["cat_pcm"], synth.synth_utterance(3, 32000)):
        for mel in (40, 80):
            want = kaldi_fp64.fbank(pcm, mel)
            got = port.fbank(pcm, mel=mel)
            assert np.abs(got - want).max() / max(1.0, np.abs(want).max()) < 1e-4
            assert (np.abs(got - want) / np.maximum(np.abs(want), 1.0)).max() < 1e-4, mel
