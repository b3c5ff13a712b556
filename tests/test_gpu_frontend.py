"""GPU parity: fbank (K1), 512-point real FFT and online CMVN (K2) through the C ABI against
the oracle, the reference's Kaldi goldens and size-independent properties."""
import numpy as np
import pytest

from catears_b200 import api, synth

pytestmark = pytest.mark.gpu


def rel_err(got, want):
    """north_star: fbank/CMVN within 1e-4 relative; relative to max(|ref|, 1)."""
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)))


def test_rfft512_known_dft():
    rng = np.random.default_rng(0)
    x = (1000 * rng.standard_normal((37, 512))).astype(np.float32)
    out = api.rfft512(x)
    X = np.fft.rfft(x.astype(np.float64), axis=1)
    scale = np.abs(X).max()
    assert np.abs(out[:, 0] - X[:, 0].real).max() < 2e-6 * scale
    assert np.abs(out[:, 1] - X[:, 256].real).max() < 2e-6 * scale
    got = out[:, 2::2] + 1j * out[:, 3::2]
    assert np.abs(got - X[:, 1:256]).max() < 2e-6 * scale


def test_rfft512_vs_oracle(port):
    rng = np.random.default_rng(1)
    x = (1000 * rng.standard_normal((8, 512))).astype(np.float32)
    out = api.rfft512(x)
    for i in range(8):
        want = port.srfft(x[i])
        assert np.abs(out[i] - want).max() <= 1e-5 * np.abs(want).max()


def test_fbank_kaldi_golden(golden):
    """test/fbank_test.cc:24-60: en-us-hello.wav -> 47x40 vs Kaldi compute-fbank-feats, 1e-4 abs."""
    fb = api.fbank(golden["hello_pcm"])
    assert fb.shape == (47, 40)
    assert np.abs(fb - golden["kaldi_fbank"]).max() < 1e-4


def test_fbank_vs_reference_vectors(golden):
    for key, pcm in (("fbank40_en-us-hello", golden["hello_pcm"]),
                     ("fbank40_en-us-cat", golden["cat_pcm"]),
                     ("fbank40_synth0_1s", synth.synth_utterance(0, 16000))):
        want = golden["ref"][key]
        got = api.fbank(pcm)
        assert got.shape == want.shape
        assert rel_err(got, want) < 1e-4, key


def test_fbank_ragged_batch_vs_oracle(port):
    """Ragged batch incl. empty, < 1 frame, exactly 1 frame, chunk boundaries (32-frame CTAs)."""
    sizes = [0, 399, 400, 559, 560, 400 + 160 * 31, 400 + 160 * 32, 400 + 160 * 33, 16000, 7802, 1, 48000]
    rng = np.random.default_rng(5)
    utts = [np.clip(np.rint(3000 * rng.standard_normal(n)), -32768, 32767).astype(np.int16) for n in sizes]
    pcm = np.concatenate(utts)
    off = np.concatenate([[0], np.cumsum(sizes)])
    got = api.fbank(pcm, off)
    fo = api.frame_offsets(off)
    assert got.shape[0] == fo[-1]
    for u, x in enumerate(utts):
        T = port.num_frames(x.size)
        assert fo[u + 1] - fo[u] == T
        if T:
            assert rel_err(got[fo[u]:fo[u + 1]], port.fbank(x)) < 1e-4, sizes[u]


def test_fbank_extreme_inputs(port):
    """Silence (floor -> log(FLT_EPSILON)), full-scale square wave, DC."""
    n = 4000
    cases = {
        "zeros": np.zeros(n, np.int16),
        "dc": np.full(n, 12345, np.int16),
        "square": np.where(np.arange(n) % 50 < 25, 32767, -32768).astype(np.int16),
        "impulse": np.eye(1, n, 777, dtype=np.int16)[0] * 32767,
    }
    for name, x in cases.items():
        got, want = api.fbank(x), port.fbank(x)
        assert got.shape == want.shape
        if name in ("zeros", "dc"):
            assert np.allclose(got, np.log(np.float32(1.1920929e-7)), atol=1e-5), name
        else:
            assert rel_err(got, want) < 1e-4, name


def test_fbank_80_bins_vs_port(port):
    pcm = synth.synth_utterance(3, 16000)
    got = api.fbank(pcm, num_mel=80)
    want = port.fbank(pcm, mel=80)
    assert got.shape == want.shape == (98, 80)
    assert rel_err(got, want) < 1e-4


@pytest.mark.parametrize("mel", [40, 80])
def test_fbank_vs_independent_fp64_kaldi_formula(golden, mel):
    """The pin for 80 mel bins (the reference aborts there, src/fbank.cc:155): tests/kaldi_fp64.py is an
    independent float64 numpy statement of the Kaldi formula -- checked against the Kaldi golden dump
    at 40 bins (tests/test_oracle.py) -- and the kernel must agree with it at 40 AND 80 bins to the same
    1e-4 (relative to max(|ref|, 1)) it meets against the reference at 40."""
    import kaldi_fp64
    for pcm in (golden["hello_pcm"], golden["cat_pcm"], synth.synth_utterance(3, 32000)):
        got = api.fbank(pcm, num_mel=mel)
        want = kaldi_fp64.fbank(pcm, mel)
        assert got.shape == want.shape
        assert rel_err(got, want.astype(np.float32)) < 1e-4, mel


def test_fbank_device_buffers_and_batch_equals_singles():
    import torch
    pcm, off = synth.synth_batch(5, 16000)
    batch = api.fbank(pcm, off)
    d_pcm = torch.from_numpy(pcm).cuda()
    d_out = torch.zeros((batch.shape[0], 40), dtype=torch.float32, device="cuda")
    api.fbank(d_pcm, off, out=d_out)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), batch)
    for u in range(5):
        single = api.fbank(pcm[off[u]:off[u + 1]])
        assert np.array_equal(single, batch[u * 98:(u + 1) * 98])


def test_cmvn_kaldi_golden(golden):
    """test/cmvn_test.cc:38-79: online CMVN of the Kaldi fbank vs apply-cmvn-online, 1e-4."""
    out = api.cmvn(golden["cmvn_stats"], golden["kaldi_fbank"])
    assert np.abs(out - golden["kaldi_cmvn"]).max() < 1e-4


def test_cmvn_bit_exact_vs_reference_vectors(golden):
    """700 frames (past the 600-frame window): the fp32 chain is replayed exactly."""
    out = api.cmvn(golden["cmvn_stats"], golden["ref"]["cmvn_in_700"])
    assert np.array_equal(out, golden["ref"]["cmvn_out_700"])


def test_cmvn_ragged_batch_bit_exact_vs_oracle(port, golden):
    rng = np.random.default_rng(11)
    sizes = [1, 0, 599, 600, 601, 1300, 47]
    feats = (14.0 + 3.0 * rng.standard_normal((sum(sizes), 40))).astype(np.float32)
    off = np.concatenate([[0], np.cumsum(sizes)])
    out = api.cmvn(golden["cmvn_stats"], feats, off)
    for u, n in enumerate(sizes):
        if n:
            want = port.cmvn(golden["cmvn_stats"], feats[off[u]:off[u + 1]])
            assert np.array_equal(out[off[u]:off[u + 1]], want), n


def test_cmvn_chain_exact_on_adversarial_values(port, golden):
    """The kernel replaces the reference's fp64 accumulate by RN(S + (x - x_old)) whenever the
    difference is exact and falls back to fp64 otherwise: mixed signs and magnitudes spread over
    twelve decades (inexact differences, catastrophic cancellation, exact zeros) must still be
    bit-identical to the in-order double chain of src/cmvn.cc:42-67, well past the 600-frame window."""
    rng = np.random.default_rng(2026)
    T = 1500
    mag = 10.0 ** rng.uniform(-6, 6, size=(T, 40))
    feats = (mag * rng.choice([-1.0, 1.0], size=(T, 40))).astype(np.float32)
    feats[rng.random((T, 40)) < 0.05] = 0.0
    feats[700:720] = feats[100:120]                  # x_t == x_{t-600}: exact zero differences
    out = api.cmvn(golden["cmvn_stats"], feats)
    assert np.array_equal(out, port.cmvn(golden["cmvn_stats"], feats))
    # a long utterance of ordinary log-mel magnitudes (rounding drift over 4000 frames)
    feats = (12.0 + 4.0 * rng.standard_normal((4000, 40))).astype(np.float32)
    out = api.cmvn(golden["cmvn_stats"], feats)
    assert np.array_equal(out, port.cmvn(golden["cmvn_stats"], feats))


def test_cmvn_long_utterances_bins_over_several_ctas(port, golden):
    """A few long utterances: the kernel gives every CTA a group of 4 (fewer than 4 utterances) or 8 bins and
    long tiles; the workers make x_t - x_{t-600} and its exactness test, one inexact element sends its whole
    tile through the fp64 form.  Bit-identical to the in-order chain of src/cmvn.cc:42-67 either way."""
    rng = np.random.default_rng(31)
    stats = golden["cmvn_stats"]
    # (a) log-mel-like values in one binade pair: every difference exact, the fast path all the way
    feats = rng.uniform(9.0, 26.0, size=(9000, 40)).astype(np.float32)
    feats = np.round(feats * 64.0) / 64.0                # coarse mantissas: exact differences
    out = api.cmvn(stats, feats.astype(np.float32))
    assert np.array_equal(out, port.cmvn(stats, feats.astype(np.float32)))
    # (b) full-mantissa values: inexact differences in most tiles (the fp64 form), 6 utterances -> groups of 8
    sizes = [4500, 4100, 5000, 4096, 4200, 6000]
    feats = (13.0 + 4.0 * rng.standard_normal((sum(sizes), 40))).astype(np.float32)
    feats[rng.random(feats.shape) < 0.02] *= 1e-3        # a few small values: certainly inexact against 13
    off = np.concatenate([[0], np.cumsum(sizes)])
    out = api.cmvn(stats, feats, off)
    for u in range(len(sizes)):
        assert np.array_equal(out[off[u]:off[u + 1]], port.cmvn(stats, feats[off[u]:off[u + 1]])), u
    # (c) one long utterance, mostly exact with isolated inexact elements (mixed fast / fp64 tiles), groups of 4
    feats = (np.round(rng.uniform(9.0, 26.0, size=(7000, 40)) * 64.0) / 64.0).astype(np.float32)
    for t in (650, 1999, 2000, 5555):
        feats[t, rng.integers(0, 40)] = np.float32(1e-3) * np.float32(rng.uniform(1, 2))
    out = api.cmvn(stats, feats)
    assert np.array_equal(out, port.cmvn(stats, feats))


def test_cmvn_in_place_on_device(golden, port):
    import torch
    rng = np.random.default_rng(12)
    feats = (14.0 + 3.0 * rng.standard_normal((900, 40))).astype(np.float32)
    d = torch.from_numpy(feats).cuda()
    api.cmvn(golden["cmvn_stats"], d, out=d)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy(), port.cmvn(golden["cmvn_stats"], feats))


def test_full_size_properties():
    """Config-2 shape at a size the oracle cannot finish quickly: 64 x 10 s utterances.
    Properties: shape, finiteness, batch == shifted copies (time-shift by one frame shift moves
    the features by exactly one row)."""
    pcm, off = synth.synth_batch(64, 160000)
    fb = api.fbank(pcm, off)
    assert fb.shape == (64 * 998, 40) and np.isfinite(fb).all()
    shifted = api.fbank(pcm[160:160000])
    assert np.array_equal(shifted, fb[1:998][:shifted.shape[0]])
