"""include/ce_host.hpp: the C++ host side that mirrors the reference's operator surface
(Fbank::Process, CMVN::GetFrame, AcousticModel::Read/Process/EndOfStream) over the C ABI.
The test program (tests/cpp/host_mirror_test.cc) drives it like src/ce_stt.cc drives the
reference: 1 KB PCM pieces, one frame at a time into the AM, rows collected for the decoder."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from catears_b200 import api, formats as F, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    if not os.path.exists(api.LIB_PATH):
        pytest.fail("libce_gpu.so is missing: run `make lib`")
    out = str(tmp_path_factory.mktemp("host_mirror") / "host_mirror_test")
    libdir = os.path.dirname(api.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cc"), "-o", out,
                           "-L" + libdir, "-lce_gpu", "-Wl,-rpath," + libdir])
    return out


def test_host_mirror_error_behaviour(exe):
    """Missing files, calls before Read, non-int16 samples and (here, without a GPU) the absence
    of any CPU fallback all come back as a Status with a message."""
    r = subprocess.run([exe, "errors"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == "OK"


def _stream(exe, tmp_path, model, pcm, precision, stats_path=None):
    pcm_path = str(tmp_path / "utt.s16le")
    out_path = str(tmp_path / ("out_%d.bin" % precision))
    pcm.astype("<i2").tofile(pcm_path)
    cmd = [exe, "stream", model["conf"], pcm_path, out_path, str(precision)]
    if stats_path:
        cmd.append(stats_path)
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(out_path, np.int32, 3)
    rows, cols, batches = int(raw[0]), int(raw[1]), int(raw[2])
    data = np.fromfile(out_path, np.float32, offset=12).reshape(rows, cols)
    return data, batches


@pytest.mark.gpu
def test_host_mirror_streaming_equals_batch_and_oracle(exe, tmp_path, golden, port):
    """chunk_size 16 forces many Process() batches; the concatenated rows must equal the whole-
    utterance evaluation (chunked == whole) and the oracle's float path (1e-3)."""
    stats = golden["cmvn_stats"]
    m = synth.write_model(str(tmp_path / "m"), name="small", hidden=64, num_pdfs=96, seed=4321,
                          chunk_size=16)
    stats_path = str(tmp_path / "stats.vec0")
    F.write_vector(stats_path, stats)
    pcm = golden["hello_pcm"]
    got, batches = _stream(exe, tmp_path, m, pcm, api.PRECISION_FP32, stats_path)
    assert got.shape == (47, 96) and batches == 2            # 47 frames = 2 x 16 + 15 at end of stream
    feats = port.cmvn(stats, port.fbank(pcm))
    prior = F.read_vector(m["prior"])
    want = port.am_forward(m["nnet"], prior, m["left"], m["right"], feats, mode="float")
    assert np.abs(got - want).max() < 1e-3
    am = api.AcousticModelGpu(config=m["conf"], precision="fp32")
    whole, _ = am.nnet(api.cmvn(stats, api.fbank(pcm)))
    am.close()
    assert np.abs(got - whole).max() < 1e-5


@pytest.mark.gpu
def test_host_mirror_int8_matches_batch_api(exe, tmp_path, golden):
    """int8 through the mirror: with chunk_size > T the buffered frames (which carry the reference's
    explicit edge replication, src/am.cc:119-124,152-155) are evaluated as ONE matrix; the rows must
    be bit-identical to the batch API fed the same explicitly padded matrix."""
    m = synth.write_model(str(tmp_path / "m8"), name="small", hidden=64, num_pdfs=96, seed=4321)
    pcm = golden["hello_pcm"]
    got, batches = _stream(exe, tmp_path, m, pcm, api.PRECISION_INT8)
    assert got.shape == (47, 96) and batches == 0
    feats = api.fbank(pcm)
    L, R = m["left"], m["right"]
    padded = np.concatenate([np.repeat(feats[:1], L, 0), feats, np.repeat(feats[-1:], R, 0)])
    am = api.AcousticModelGpu(config=m["conf"], precision="int8")
    whole, _ = am.nnet(padded)
    am.close()
    assert np.array_equal(got, whole[L:L + 47])


@pytest.mark.gpu
def test_stream_batch_micro_batches_equal_whole_utterances(exe, tmp_path, golden):
    """five live utterances (0.5 s to 12 s, one shorter than a frame) arrive in
    random-sized pieces and are evaluated together, one fbank / CMVN / AM pass per call, with the
    sample remainder, the CMVN sums + 600-frame history and the AM context carried between calls.
    Per stream the concatenated rows must equal the whole-utterance evaluation: the CMVN is
    bit-identical by construction, the fp32 log-likelihoods agree to 1e-5."""
    stats = golden["cmvn_stats"]
    m = synth.write_model(str(tmp_path / "m"), name="small", hidden=64, num_pdfs=96, seed=4321, cmvn_stats=stats)
    stats_path = str(tmp_path / "stats.vec0")
    F.write_vector(stats_path, stats)
    lengths = [8000, 192000, 300, 47001, 112345]              # 12 s > 600 frames: the CMVN window slides
    pcms, paths = [], []
    for i, n in enumerate(lengths):
        pcm = synth.synth_utterance(40 + i, n)
        pcms.append(pcm)
        paths.append(str(tmp_path / ("s%d.s16le" % i)))
        pcm.astype("<i2").tofile(paths[-1])
    prefix = str(tmp_path / "rows")
    for seed in (1, 2):
        r = subprocess.run([exe, "streams", m["conf"], str(api.PRECISION_FP32), stats_path, str(seed), prefix] + paths,
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        am = api.AcousticModelGpu(config=m["conf"], precision="fp32")
        for i, pcm in enumerate(pcms):
            raw = np.fromfile("%s.%d.bin" % (prefix, i), np.int32, 3)
            got = np.fromfile("%s.%d.bin" % (prefix, i), np.float32, offset=12).reshape(int(raw[0]), int(raw[1]))
            frames = 0 if pcm.size < 400 else 1 + (pcm.size - 400) // 160
            assert got.shape == (frames, 96), (i, got.shape)
            assert int(raw[2]) > 3                               # really several micro-batches
            if frames:
                whole, _, _ = am.forward(pcm)
                assert np.abs(got - whole).max() < 1e-5, i
        am.close()


@pytest.mark.gpu
def test_stream_batch_with_selected_outputs(exe, tmp_path, golden):
    """SURVEY 8f rank 4 through the serving form: the same micro-batched streams with rows gathered
    to a pdf subset, and with the 8 best (loglik, pdf) pairs per frame."""
    m = synth.write_model(str(tmp_path / "m"), name="small", hidden=64, num_pdfs=96, seed=4321)
    pcms, paths = [], []
    for i, n in enumerate([8000, 47001, 64000]):
        pcms.append(synth.synth_utterance(60 + i, n))
        paths.append(str(tmp_path / ("s%d.s16le" % i)))
        pcms[-1].astype("<i2").tofile(paths[-1])
    prefix = str(tmp_path / "rows")
    am = api.AcousticModelGpu(config=m["conf"], precision="fp32")
    whole = [am.forward(p)[0] for p in pcms]
    am.close()
    ids = [90, 3, 4, 41, 3]
    for sel in ("subset:" + ",".join(map(str, ids)), "topk:8"):
        env = dict(os.environ, HOST_MIRROR_SELECT=sel)
        r = subprocess.run([exe, "streams", m["conf"], str(api.PRECISION_FP32), "-", "3", prefix] + paths,
                           capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stdout + r.stderr
        for i, w in enumerate(whole):
            raw = np.fromfile("%s.%d.bin" % (prefix, i), np.int32, 3)
            got = np.fromfile("%s.%d.bin" % (prefix, i), np.float32, offset=12).reshape(int(raw[0]), int(raw[1]))
            assert got.shape[0] == w.shape[0]
            if sel.startswith("subset"):
                assert got.shape[1] == len(ids)
                assert np.abs(got - w[:, ids]).max() < 1e-5
            else:
                best = got.view(api.SCORED_PDF)
                assert best.shape == (w.shape[0], 8)
                assert np.abs(best["loglik"] - np.sort(w, axis=1)[:, ::-1][:, :8]).max() < 1e-5
                assert np.abs(best["loglik"] - np.take_along_axis(w, best["pdf"], axis=1)).max() < 1e-5
                assert (np.diff(best["loglik"], axis=1) <= 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("prec,select", [("fp32", None), ("int8", None), ("fp32", "topk:8")])
def test_device_resident_streams_equal_host_resident_bit_for_bit(exe, tmp_path, golden, prec, select):
    """ce_gpu_streams_* keeps the sample remainder, the CMVN sums + history and the AM context of
    every live utterance in device buffers; fed with the same random schedule it must produce the
    very rows of ce_host::StreamBatch (which carries that state on the host): same kernels, same
    inputs."""
    stats = golden["cmvn_stats"]
    m = synth.write_model(str(tmp_path / "m"), name="small", hidden=64, num_pdfs=96, seed=4321, cmvn_stats=stats)
    stats_path = str(tmp_path / "stats.vec0")
    F.write_vector(stats_path, stats)
    paths = []
    for i, n in enumerate([8000, 192000, 300, 47001, 112345, 399, 400]):   # 12 s: the CMVN window slides
        paths.append(str(tmp_path / ("s%d.s16le" % i)))
        synth.synth_utterance(40 + i, n).astype("<i2").tofile(paths[-1])
    out = {}
    for mode in ("host", "device"):
        env = dict(os.environ)
        if mode == "device":
            env["HOST_MIRROR_DEVICE_STATE"] = "1"
        if select:
            env["HOST_MIRROR_SELECT"] = select
        prefix = str(tmp_path / ("rows_" + mode))
        r = subprocess.run([exe, "streams", m["conf"], str(api.PRECISIONS[prec]), stats_path, "5", prefix] + paths,
                           capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stdout + r.stderr
        out[mode] = [np.fromfile("%s.%d.bin" % (prefix, i), np.int32) for i in range(len(paths))]
    for i, (a, b) in enumerate(zip(out["host"], out["device"])):
        assert a[2] > 3 and np.array_equal(a, b), i            # header and every row bit
    assert out["host"][1][0] == 1198 and out["host"][2][0] == 0 and out["host"][6][0] == 1


@pytest.mark.gpu
def test_stream_set_slot_reuse_and_row_capacity(tmp_path, golden):
    """A slot that ended is free again and starts from nothing; a row buffer that is too small is
    refused before any state changes."""
    stats = golden["cmvn_stats"]
    m = synth.write_model(str(tmp_path / "m"), name="small", hidden=64, num_pdfs=96, seed=4321, cmvn_stats=stats)
    am = api.AcousticModelGpu(config=m["conf"], precision="fp32")
    s = api.StreamSet(am, 2)
    try:
        a, b, c = (synth.synth_utterance(70 + i, n) for i, n in enumerate((30000, 52000, 41000)))
        sa, sb = s.open(), s.open()
        assert (sa, sb) == (0, 1)
        with pytest.raises(api.CeGpuError):
            s.open()                                         # both slots taken
        got_a = [s.process([sa, sb], [a[:20000], b[:7000]], [False, False])]
        assert s.rows_ready([sa], [a[20000:]], [True]) == 186 - got_a[0][0].shape[0]
        with pytest.raises(api.CeGpuError, match="nothing was changed"):
            s.process([sa], [a[20000:]], [True], rows_cap=3)
        got_a.append(s.process([sa, sb], [a[20000:], b[7000:30000]], [True, False]))
        sc = s.open()
        assert sc == sa                                      # the ended slot is free again
        got_c = s.process([sc, sb], [c, b[30000:]], [True, True])
        rows_a = np.concatenate([got_a[0][0], got_a[1][0]])
        rows_b = np.concatenate([got_a[0][1], got_a[1][1], got_c[1]])
        for rows, pcm in ((rows_a, a), (rows_b, b), (got_c[0], c)):
            whole, _, _ = am.forward(pcm)
            assert rows.shape == whole.shape
            assert np.abs(rows - whole).max() < 1e-5
        with pytest.raises(api.CeGpuError):
            s.process([sb], [None], [False])                 # ended: not open any more
        big = synth.synth_utterance(75, 192000)              # 12 s in ONE call: 460 KB of rows, copied in pieces
        sd = s.open()
        rows_big = s.process([sd], [big], [True])[0]
        whole, _, _ = am.forward(big)
        assert rows_big.shape == whole.shape == (1198, 96)
        assert np.abs(rows_big - whole).max() < 1e-5
    finally:
        s.close()
        am.close()


@pytest.mark.gpu
def test_cmvn_stream_is_bit_identical_for_any_split(golden):
    """ce_gpu_cmvn_stream: 1500 frames normalised in pieces of every size (1 .. 700 frames), two
    utterances at once, equal bit for bit to one ce_gpu_cmvn call."""
    rng = np.random.default_rng(8)
    stats = golden["cmvn_stats"]
    feats = [(12.0 + 4.0 * rng.standard_normal((1500, 40))).astype(np.float32),
             (9.0 + 2.0 * rng.standard_normal((777, 40))).astype(np.float32)]
    want = [api.cmvn(stats, f) for f in feats]
    state = np.zeros((2, 40), np.float32)
    t = [0, 0]
    got = [[], []]
    while t[0] < 1500 or t[1] < 777:
        step = [int(rng.integers(0, 700)), int(rng.integers(0, 300))]
        parts, nh, off = [], [], [0]
        for u in range(2):
            step[u] = min(step[u], feats[u].shape[0] - t[u])
            h = min(t[u], 600)
            parts.append(feats[u][t[u] - h:t[u] + step[u]])
            nh.append(h)
            off.append(off[-1] + h + step[u])
        out = api.cmvn_stream(stats, np.concatenate(parts), off, nh, t, state)
        a = 0
        for u in range(2):
            got[u].append(out[a:a + step[u]])
            a += step[u]
            t[u] += step[u]
    for u in range(2):
        assert np.array_equal(np.concatenate(got[u]), want[u]), u
