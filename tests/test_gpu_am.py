"""GPU parity of the acoustic model (AcousticModel::Process/EndOfStream, src/am.cc:115-164 +
Nnet::Propagate) through ce_gpu_nnet / ce_gpu_forward: int8 accumulators bit-exact, float
log-likelihoods within 1e-3 absolute (north_star), argmax equal wherever the oracle's top-2
margin is not a rounding tie."""
import numpy as np
import pytest

from catears_b200 import api, formats as F, synth

pytestmark = pytest.mark.gpu


def check_argmax(am, loglik_ref, tie=1e-4):
    top2 = np.sort(loglik_ref, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > tie
    want = loglik_ref.argmax(axis=1)
    assert np.array_equal(am[clear], want[clear])
    return float(np.mean(am == want))


@pytest.fixture(scope="module")
def models(small_model):
    out = {}
    for prec in ("int8", "fp32", "tf32", "bf16"):
        out[prec] = api.AcousticModelGpu(config=small_model["conf"], precision=prec)
    yield out
    for m in out.values():
        m.close()


def test_model_info(models, small_model):
    m = models["int8"]
    assert (m.num_pdfs, m.left_context, m.right_context, m.feat_dim) == (96, 13, 13, 40)


def test_am_float_reference_vectors(golden, models):
    """Log-likelihoods of the unmodified reference (in-order sgemm) on 57 frames."""
    x, want = golden["ref"]["am_in"], golden["ref"]["am_float"]
    ll, am = models["fp32"].nnet(x)
    assert ll.shape == want.shape == (57, 96)
    assert np.abs(ll - want).max() < 1e-3                  # north_star tolerance
    assert check_argmax(am, want) > 0.98
    ll_tf32, _ = models["tf32"].nnet(x)
    assert np.abs(ll_tf32 - want).max() < 2e-2             # single-pass TF32: 10-bit mantissa
    ll_bf16, _ = models["bf16"].nnet(x)
    assert np.abs(ll_bf16 - want).max() < 0.15             # bf16 operands: 8-bit mantissa, 7 layers


def test_am_u8_reference_vectors_bit_exact(golden, models):
    """int8 composition: accumulators of Linear 1 and of the last Linear (6)
    bit-exact, hence every quantisation in between was bit-exact too."""
    r = golden["ref"]
    m = models["int8"]
    for ordinal, key in ((1, "am_u8_acc_linear1"), (6, "am_u8_acc_linear6")):
        m.keep_acc(ordinal)
        ll, am = m.nnet(r["am_in"])
        acc = m.get_acc(0)
        assert acc.shape == r[key].shape
        assert np.array_equal(acc, r[key]), key
    m.keep_acc(-1)
    assert np.abs(ll - r["am_u8"]).max() < 1e-5            # only the log-sum-exp order differs
    assert check_argmax(am, r["am_u8"]) > 0.98


@pytest.mark.parametrize("layout", ["auto", "32", "128"])
def test_am_u8_ragged_batch_bit_exact_vs_oracle(port, models, small_model, monkeypatch, layout):
    """Batch of utterances of T = 1, 2, 7, 57, 130, 300 (T=1..7 exercise the rows-actually-read
    rule of the fused FindMinMax); accumulators of the last layer and log-likelihoods per
    utterance against the oracle run one utterance at a time.  The row space packs these short
    blocks at multiples of 32 rows (granule-mode int8 epilogue: another utterance in every quadrant
    of a GEMM tile); both layouts are also forced."""
    if layout != "auto":
        monkeypatch.setenv("CE_GPU_ROW_GRAN", layout)
    rng = np.random.default_rng(21)
    sizes = [1, 57, 2, 130, 7, 300]
    feats = rng.standard_normal((sum(sizes), 40)).astype(np.float32) * 2.0
    off = np.concatenate([[0], np.cumsum(sizes)])
    prior = F.read_vector(small_model["prior"])
    m = models["int8"]
    m.keep_acc(6)
    ll, am = m.nnet(feats, off)
    for u, T in enumerate(sizes):
        x = feats[off[u]:off[u + 1]]
        want, wacc = port.am_forward(small_model["nnet"], prior, 13, 13, x, mode="u8", dump_layer=6,
                                     acc_shape=(T, 96))
        assert np.array_equal(m.get_acc(u), wacc), T
        assert np.abs(ll[off[u]:off[u + 1]] - want).max() < 1e-5, T
        check_argmax(am[off[u]:off[u + 1]], want)
    m.keep_acc(-1)


def test_am_float_ragged_batch_vs_oracle(port, models, small_model):
    rng = np.random.default_rng(22)
    sizes = [3, 0, 200, 1, 64]
    feats = rng.standard_normal((sum(sizes), 40)).astype(np.float32)
    off = np.concatenate([[0], np.cumsum(sizes)])
    prior = F.read_vector(small_model["prior"])
    ll, am = models["fp32"].nnet(feats, off)
    for u, T in enumerate(sizes):
        if T == 0:
            continue
        want = port.am_forward(small_model["nnet"], prior, 13, 13, feats[off[u]:off[u + 1]])
        assert np.abs(ll[off[u]:off[u + 1]] - want).max() < 1e-3, T


@pytest.mark.parametrize("layout", ["32", "128"])
def test_batch_equals_singles_and_chunking(models, monkeypatch, layout):
    """Property: batch-of-N == N singles, bit for bit (int8), in both row-space layouts."""
    monkeypatch.setenv("CE_GPU_ROW_GRAN", layout)
    rng = np.random.default_rng(23)
    sizes = [90, 150, 40, 260, 5, 31, 33]
    feats = rng.standard_normal((sum(sizes), 40)).astype(np.float32)
    off = np.concatenate([[0], np.cumsum(sizes)])
    m = models["int8"]
    ll, am = m.nnet(feats, off)
    for u in range(len(sizes)):
        l1, a1 = m.nnet(feats[off[u]:off[u + 1]])
        assert np.array_equal(l1, ll[off[u]:off[u + 1]])
        assert np.array_equal(a1, am[off[u]:off[u + 1]])


def test_forward_pcm_to_loglik(golden, port, small_model, tmp_path):
    """Config 1: bundled en-us-hello.wav -> fbank -> CMVN (bundled stats) -> TDNN; the GPU
    pipeline against the oracle pipeline stage by stage composed on the CPU."""
    stats = tmp_path / "cmvn.bin"
    F.write_vector(str(stats), golden["cmvn_stats"])
    prior = F.read_vector(small_model["prior"])
    pcm = np.concatenate([golden["hello_pcm"], golden["cat_pcm"]])
    off = np.array([0, golden["hello_pcm"].size, pcm.size], np.int64)
    for prec, tol in (("fp32", 1e-3), ("int8", None)):
        m = api.AcousticModelGpu(nnet=small_model["nnet"], prior=small_model["prior"], left_context=13,
                                 right_context=13, cmvn_stats=str(stats), precision=prec)
        ll, am, fo = m.forward(pcm, off)
        assert list(fo) == [0, 47, 47 + port.num_frames(golden["cat_pcm"].size)]
        for u, x in enumerate((golden["hello_pcm"], golden["cat_pcm"])):
            feats = port.cmvn(golden["cmvn_stats"], port.fbank(x))
            want = port.am_forward(small_model["nnet"], prior, 13, 13, feats,
                                   mode="u8" if prec == "int8" else "float")
            got = ll[fo[u]:fo[u + 1]]
            if tol is not None:
                assert np.abs(got - want).max() < tol
            else:
                # int8: GPU fbank differs from the oracle's by ~1e-6 relative, which may move a
                # handful of u8 codes by one step; bounded by the quantisation error budget
                # (test/gemm_test.cc:120: 1 % of the range)
                assert np.abs(got - want).max() < 0.01 * (want.max() - want.min())
        m.close()


def test_no_logsoftmax_and_plain_linear_stack(port, tmp_path):
    """A stack without Splice and without LogSoftmax (test/nnet_test.cc:59-90 style Linear)."""
    rng = np.random.default_rng(31)
    layers = [{"type": F.LINEAR, "W": rng.standard_normal((40, 24)).astype(np.float32),
               "b": rng.standard_normal(24).astype(np.float32)},
              {"type": F.RELU},
              {"type": F.LINEAR, "W": rng.standard_normal((24, 12)).astype(np.float32),
               "b": rng.standard_normal(12).astype(np.float32)}]
    nnet, prior = str(tmp_path / "p.nnet"), str(tmp_path / "p.prior")
    F.write_nnet(nnet, layers, 0, 0)
    pr = np.full(12, 1.0 / 12, np.float32)
    F.write_vector(prior, pr)
    x = rng.standard_normal((33, 40)).astype(np.float32)
    want = port.am_forward(nnet, pr, 0, 0, x)
    m = api.AcousticModelGpu(nnet=nnet, prior=prior, precision="fp32")
    ll, _ = m.nnet(x)
    assert np.abs(ll - want).max() < 1e-3
    m.close()


def test_splice_without_narrow_runs_as_general_program(tmp_path, port):
    """A Splice that is not followed by its Narrow (not the tool/convert_am.py pattern) no longer is
    CE_GPU_EUNSUPPORTED: it compiles to the general layer program (edge clamp of src/nnet.cc:64-66).
    With contexts the stack does not remove, loading fails like the reference's assert (am.cc:106)."""
    rng = np.random.default_rng(32)
    layers = [{"type": F.SPLICE, "indices": [-1, 0, 1]},
              {"type": F.LINEAR, "W": rng.standard_normal((120, 8)).astype(np.float32),
               "b": rng.standard_normal(8).astype(np.float32)}]
    nnet, prior = str(tmp_path / "u.nnet"), str(tmp_path / "u.prior")
    F.write_nnet(nnet, layers, 0, 0)
    pr = np.full(8, 0.125, np.float32)
    F.write_vector(prior, pr)
    with pytest.raises(api.CeGpuError, match="does not match the rows the nnet removes"):
        api.AcousticModelGpu(nnet=nnet, prior=prior, left_context=1, right_context=1)
    x = rng.standard_normal((9, 40)).astype(np.float32)
    m = api.AcousticModelGpu(nnet=nnet, prior=prior, precision="fp32")
    ll, _ = m.nnet(x)
    m.close()
    idx = np.clip(np.arange(9)[:, None] + np.array([-1, 0, 1])[None, :], 0, 8)
    want = x[idx].reshape(9, 120).astype(np.float64) @ layers[1]["W"].astype(np.float64) + layers[1]["b"]
    assert np.abs(ll - (want - np.log(pr))).max() < 1e-3


def test_two_handles_two_threads_concurrently(small_model, port):
    """include/ce_gpu.h: distinct handles are independent.  Two host threads, each with its own
    model handle and CUDA stream on the same device, run the whole path at the same time; both get
    the result of a lone run, bit for bit (int8)."""
    import threading

    import torch
    pcm, off = synth.synth_batch(6, 16000 * 3)
    lone = api.AcousticModelGpu(config=small_model["conf"], precision="int8")
    want_ll, want_am, _ = lone.forward(pcm, off)
    lone.close()
    results, errors = {}, []

    def work(tag):
        try:
            m = api.AcousticModelGpu(config=small_model["conf"], precision="int8")
            s = torch.cuda.Stream()
            for _ in range(5):
                ll, am, _ = m.forward(pcm, off, stream=s)
            results[tag] = (ll, am)
            m.close()
        except Exception as e:          # surfaced in the main thread below
            errors.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for tag in (0, 1):
        assert np.array_equal(results[tag][0], want_ll)
        assert np.array_equal(results[tag][1], want_am)


def test_nnet_chunks_is_compute_batch_int8_bit_exact(ref, models, small_model):
    """ce_gpu_nnet_chunks = AcousticModel::ComputeBatch (src/am.cc:82-113) of many chunks at once: every
    block carries its own context the way Process / EndOfStream stack the buffer (chunk_size 16 here),
    nothing is replicated, and for int8 every block is ONE Quantize matrix per layer -- the reference's
    int8 composition run block by block (contexts 0: the block is taken as it is) gives the same rows."""
    if ref is None:
        pytest.skip("oracle/_ref was never built")
    rng = np.random.default_rng(41)
    T, L, R, chunk = 100, 13, 13, 16
    x = (rng.standard_normal((T, 40)) * 2).astype(np.float32)
    padded = np.concatenate([np.repeat(x[:1], L, 0), x, np.repeat(x[-1:], R, 0)])
    blocks, a = [], 0
    while a + chunk + L + R <= L + T:                      # Process: a batch once chunk + L + R frames are buffered
        blocks.append(padded[a:a + chunk + L + R])
        a += chunk
    blocks.append(padded[a:])                              # EndOfStream: the rest with the right padding
    off = np.concatenate([[0], np.cumsum([b.shape[0] for b in blocks])])
    ll, am = models["int8"].nnet_chunks(np.concatenate(blocks), off)
    assert ll.shape == (T, 96)
    o = 0
    for b in blocks:
        want, _ = ref.u8_forward(small_model["nnet"], small_model["prior"], 0, 0, b)
        n = b.shape[0] - L - R
        assert want.shape == (n, 96)
        assert np.abs(ll[o:o + n] - want).max() < 1e-5
        check_argmax(am[o:o + n], want)
        o += n
    # float: the same rows as the whole utterance (chunked == whole)
    whole, _ = models["fp32"].nnet(x)
    ll32, _ = models["fp32"].nnet_chunks(np.concatenate(blocks), off)
    assert np.abs(ll32 - whole).max() < 1e-4
