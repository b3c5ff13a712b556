"""SURVEY 8f rank 1 -- the drop-in itself: the reference's own `pocketkaldi` binary built twice from
the reference sources where they lie (oracle/Makefile `dropin`):

  oracle/_ref/pocketkaldi_ref   every reference source unchanged (CPU fbank + AM + decoder)
  oracle/_ref/pocketkaldi_gpu   the same sources, except that src/ce_stt.cc is replaced by
                                integration/ce_stt_gpu.cc (front end + AM through libce_gpu.so);
                                decoder, FST, hash table, symbol table and WAV reader are the
                                reference's unchanged code, its CPU fbank is not linked and its
                                cblas_sgemm aborts if reached.

Both decode the same audio with the same synthetic model and graph (the reference bundles neither,
SURVEY D5); the recognised word sequences must be identical."""
import os
import subprocess
import wave

import numpy as np
import pytest

from catears_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
BIN_REF = os.path.join(REF_DIR, "pocketkaldi_ref")
BIN_GPU = os.path.join(REF_DIR, "pocketkaldi_gpu")
MAKE_GRAPH = os.path.join(REF_DIR, "make_graph")
STREAM_REF = os.path.join(REF_DIR, "stream_ref")
STREAM_GPU = os.path.join(REF_DIR, "stream_gpu")
BIN_BATCH = os.path.join(REF_DIR, "pocketkaldi_batch")
GOLDEN = os.path.join(ROOT, "tests", "golden")

pytestmark = pytest.mark.skipif(
    not all(os.path.exists(p) for p in (BIN_REF, BIN_GPU, MAKE_GRAPH, STREAM_REF, STREAM_GPU)),
    reason="oracle/_ref drop-in binaries were never built (needs /root/reference: make -C oracle dropin)")


@pytest.fixture(scope="module")
def setup(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("dropin"))
    m = input_sensitive_model(d)
    subprocess.check_call([MAKE_GRAPH, d, "96", "12", "7"], stdout=subprocess.DEVNULL)
    os.replace(os.path.join(d, "tid2pdf.bin"), m["tid2pdf"])      # the graph's transition-id map
    with open(m["conf"], "a") as f:
        f.write("fst = HCLG.fst\nsymbol_table = words.txt\n")
    wav = os.path.join(d, "synth10s.wav")
    with wave.open(wav, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(16000)
        # 10 s of a gated frequency sweep: a non-stationary spectrum, so that the best path really
        # moves through the graph (stationary noise keeps the decoder on one word)
        rng = np.random.default_rng(5)
        t = np.arange(160000) / 16000.0
        f = 300 + 3000 * (0.5 + 0.5 * np.sin(2 * np.pi * 0.7 * t))
        ph = 2 * np.pi * np.cumsum(f) / 16000.0
        pcm = 6000 * np.sin(ph) * (0.3 + 0.7 * (np.sin(2 * np.pi * 3 * t) > 0)) + 50 * rng.standard_normal(t.size)
        w.writeframes(np.clip(np.round(pcm), -32768, 32767).astype("<i2").tobytes())
    scp = os.path.join(d, "list.scp")
    with open(scp, "w") as f:
        f.write("hello %s\ncat %s\nsynth %s\n" % (os.path.join(GOLDEN, "en-us-hello.wav"),
                                                 os.path.join(GOLDEN, "en-us-cat.wav"), wav))
    return {"conf": m["conf"], "wav": wav, "scp": scp}


def input_sensitive_model(d, num_pdfs=96, hidden=64, seed=99):
    """A two-layer TDNN whose output really depends on the audio (a random deep stack on raw,
    un-normalised fbank gives almost the same distribution for every frame, and the decoder would
    then stay on one word): the first layer's bias removes the features' common offset, the last
    layer has a gain, the prior is flat.  Written in the reference's NN02 / VEC0 / config formats."""
    from catears_b200 import formats as F
    rng = np.random.default_rng(seed)
    k = 3 * 40
    w1 = (rng.standard_normal((k, hidden)) / np.sqrt(k)).astype(np.float32)
    b1 = (-14.0 * w1.sum(0)).astype(np.float32)
    w2 = (3.0 * rng.standard_normal((hidden, num_pdfs)) / np.sqrt(hidden)).astype(np.float32)
    layers = [{"type": F.SPLICE, "indices": [-1, 0, 1]}, {"type": F.NARROW, "left": 1, "right": 1},
              {"type": F.LINEAR, "W": w1, "b": b1}, {"type": F.RELU},
              {"type": F.BATCHNORM, "scale": rng.uniform(0.5, 1.5, hidden).astype(np.float32),
               "offset": (0.1 * rng.standard_normal(hidden)).astype(np.float32)},
              {"type": F.LINEAR, "W": w2, "b": np.zeros(num_pdfs, np.float32)}, {"type": F.LOGSOFTMAX}]
    p = {x: os.path.join(d, "toy." + x) for x in ("nnet", "prior", "tid2pdf", "conf")}
    F.write_nnet(p["nnet"], layers, 1, 1)
    F.write_vector(p["prior"], np.full(num_pdfs, 1.0 / num_pdfs, np.float32))
    F.write_vector(p["tid2pdf"], np.arange(num_pdfs, dtype=np.int32), dtype="<i4")
    F.write_am_config(p["conf"], p["nnet"], p["prior"], 1, 1, 16, num_pdfs, p["tid2pdf"], {})
    return p


def run(binary, conf, audio, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([binary, conf, audio], capture_output=True, text=True, env=e)


def test_reference_binary_decodes_with_synthetic_graph(setup):
    """The unmodified reference (CPU) runs end to end on the synthetic model + graph."""
    r = run(BIN_REF, setup["conf"], os.path.join(GOLDEN, "en-us-hello.wav"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().startswith("word")


def test_gpu_binary_has_no_cpu_fallback(setup):
    if api.device_count() > 0:
        pytest.skip("a GPU is visible")
    r = run(BIN_GPU, setup["conf"], os.path.join(GOLDEN, "en-us-hello.wav"))
    assert r.returncode != 0
    assert "no CUDA device" in (r.stdout + r.stderr)


@pytest.mark.gpu
def test_unchanged_decoder_on_gpu_loglikelihoods_matches_reference(setup):
    """Same words from the reference's CPU pipeline and from its unchanged decoder fed by the GPU
    front end + acoustic model: single wavs, a 10 s utterance, and an .scp list."""
    for audio in (os.path.join(GOLDEN, "en-us-hello.wav"), os.path.join(GOLDEN, "en-us-cat.wav"),
                  setup["wav"], setup["scp"]):
        ref = run(BIN_REF, setup["conf"], audio)
        gpu = run(BIN_GPU, setup["conf"], audio)
        assert ref.returncode == 0, ref.stdout + ref.stderr
        assert gpu.returncode == 0, gpu.stdout + gpu.stderr
        assert ref.stdout.strip() != ""
        assert gpu.stdout == ref.stdout, (audio, ref.stdout, gpu.stdout)
    words = run(BIN_REF, setup["conf"], setup["wav"]).stdout.split()
    assert len(words) >= 20 and len(set(words)) >= 3              # a real path through the graph


@pytest.mark.gpu
def test_unchanged_decoder_on_selected_gpu_rows(setup):
    """fewer bytes per frame to the host decoder.  The rows gathered to the pdfs
    the graph can reach (with the remapped transition-id map) and the top-k rows with k = num_pdfs
    decode to exactly the reference's words; a small k is an approximation that still decodes."""
    for audio in (setup["wav"], setup["scp"]):
        ref = run(BIN_REF, setup["conf"], audio)
        assert ref.returncode == 0 and ref.stdout.strip() != ""
        for sel in ("subset", "topk:96", "dense"):
            gpu = run(BIN_GPU, setup["conf"], audio, env={"CE_STT_GPU_OUTPUT": sel})
            assert gpu.returncode == 0, gpu.stdout + gpu.stderr
            assert gpu.stdout == ref.stdout, (sel, audio)
    approx = run(BIN_GPU, setup["conf"], setup["wav"], env={"CE_STT_GPU_OUTPUT": "topk:24"})
    assert approx.returncode == 0 and len(approx.stdout.split()) >= 20
    bad = run(BIN_GPU, setup["conf"], setup["wav"], env={"CE_STT_GPU_OUTPUT": "topk:0"})
    assert bad.returncode != 0


def write_wav(path, pcm, width):
    """Mono 16 kHz PCM with `width` bytes per sample (8-bit is written as the SIGNED bytes the
    reference reads, src/pcm_reader.cc:36-40)."""
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(width)
        w.setframerate(16000)
        dt = {1: "<i1", 2: "<i2", 4: "<i4"}[width]
        w.writeframes(np.asarray(pcm).astype(dt).tobytes())


@pytest.mark.gpu
@pytest.mark.parametrize("piece", [1024, 777, 100000])
def test_streaming_partial_hypotheses_match_reference(setup, piece):
    """src/main.cc:28-52 feeds ce_stt_process 1 KB at a time and the reference refreshes utt->hyp every 20
    decoded frames (src/ce_stt.cc:326-327).  The GPU shim streams through ce_gpu_streams_* and hands the
    decoder its rows in the reference's own batches (chunk_size 16 here), so after EVERY call utt->hyp
    must be what the all-CPU reference holds: the drivers print it whenever it changes -- the two
    transcripts (byte positions, partial texts, final text) must be identical.  Also an odd piece size
    (samples split across calls, src/pcm_reader.cc:161,185-187) and one piece for the whole file."""
    for audio in (setup["wav"], os.path.join(GOLDEN, "en-us-hello.wav")):
        ref = subprocess.run([STREAM_REF, setup["conf"], audio, str(piece)], capture_output=True, text=True)
        gpu = subprocess.run([STREAM_GPU, setup["conf"], audio, str(piece)], capture_output=True, text=True)
        assert ref.returncode == 0, ref.stdout + ref.stderr
        assert gpu.returncode == 0, gpu.stdout + gpu.stderr
        assert gpu.stdout == ref.stdout, (audio, piece)
        if audio == setup["wav"] and piece == 1024:
            lines = ref.stdout.strip().splitlines()
            assert sum(l.startswith("partial") for l in lines) >= 10      # the text really grows call by call
            assert lines[-1].startswith("final")


@pytest.mark.gpu
def test_delta_lm_rescoring_is_wired(setup, tmp_path):
    """large_lm / original_lm (src/ce_stt.cc:84-113): the unchanged DeltaLmFst rescoring on GPU rows gives
    the reference's words.  The graph's words are entries of the reference's test/data/lm.words.txt so
    that its G.pfst / lm.1order.bin fixtures (test/fst_test.cc:178-197) apply."""
    import shutil
    d = str(tmp_path)
    m = input_sensitive_model(d)
    subprocess.check_call([MAKE_GRAPH, d, "96", "12", "7", "3"], stdout=subprocess.DEVNULL)
    os.replace(os.path.join(d, "tid2pdf.bin"), m["tid2pdf"])
    for f in ("G.pfst", "lm.1order.bin", "lm.words.txt"):
        shutil.copy(os.path.join(GOLDEN, f), d)
    with open(m["conf"], "a") as f:
        f.write("fst = HCLG.fst\nsymbol_table = lm.words.txt\nlarge_lm = G.pfst\noriginal_lm = lm.1order.bin\n")
    ref = subprocess.run([STREAM_REF, m["conf"], setup["wav"]], capture_output=True, text=True)
    gpu = subprocess.run([STREAM_GPU, m["conf"], setup["wav"]], capture_output=True, text=True)
    assert ref.returncode == 0, ref.stdout + ref.stderr
    assert gpu.returncode == 0, gpu.stdout + gpu.stderr
    assert gpu.stdout == ref.stdout
    assert len(ref.stdout.strip().splitlines()[-1].split()) >= 10
    # and it changes the result: without the delta LM the same graph decodes differently
    plain = str(tmp_path / "plain.conf")
    with open(m["conf"]) as f:
        lines = [l for l in f if not l.startswith(("large_lm", "original_lm"))]
    with open(plain, "w") as f:
        f.writelines(lines)
    base = subprocess.run([STREAM_REF, plain, setup["wav"]], capture_output=True, text=True)
    assert base.returncode == 0 and base.stdout != ref.stdout


@pytest.mark.gpu
def test_pcm_widths(setup, tmp_path):
    """src/pcm_reader.cc:168-182 passes 8-, 16- and 32-bit sample VALUES on unscaled.  8-bit and
    16-bit-range 32-bit files decode to the reference's words; a 32-bit sample outside 16 bits is
    refused with a message instead of being clamped."""
    with wave.open(setup["wav"]) as w:
        pcm = np.frombuffer(w.readframes(w.getnframes()), np.int16)
    # (8-bit: the signal divided by 64 and clipped; at 1/256 of the amplitude the REFERENCE's own decoder
    # runs into its symbol-table assertion on this synthetic graph, so there is nothing to compare with)
    cases = {"w32": (pcm.astype(np.int32), 4), "w8": (np.clip(pcm.astype(np.int32) // 64, -128, 127), 1)}
    for name, (x, width) in cases.items():
        path = str(tmp_path / (name + ".wav"))
        write_wav(path, x, width)
        ref = subprocess.run([STREAM_REF, setup["conf"], path], capture_output=True, text=True)
        gpu = subprocess.run([STREAM_GPU, setup["conf"], path], capture_output=True, text=True)
        assert ref.returncode == 0, ref.stdout + ref.stderr
        assert gpu.returncode == 0, gpu.stdout + gpu.stderr
        assert gpu.stdout == ref.stdout, name
    loud = str(tmp_path / "loud32.wav")
    write_wav(loud, pcm.astype(np.int32) * 4096, 4)
    gpu = subprocess.run([STREAM_GPU, setup["conf"], loud], capture_output=True, text=True)
    assert gpu.returncode != 0 and "does not fit 16 bits" in gpu.stderr


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(BIN_BATCH), reason="oracle/_ref/pocketkaldi_batch was never built")
def test_batch_decode_over_all_gpus_matches_reference_scp(setup, tmp_path):
    """SURVEY 8e through the C++ surface: `pocketkaldi_batch <conf> <scp>` evaluates the whole list as one
    batch over every visible GPU (ce_host::ShardedModel: ce_gpu_partition, one handle + host thread per
    GPU, graph-reachable pdf columns into pinned host rows) and decodes every utterance with the
    reference's unchanged Decoder on a CPU thread pool as its rows arrive.  The "<name> <words>" lines
    must equal the all-CPU reference's scp mode (src/main.cc:55-84), utterance for utterance."""
    rng = np.random.default_rng(12)
    lines = []
    for i in range(14):                                   # ragged lengths, one shorter than a frame
        n = [160000, 48000, 8000, 200, 100000][i % 5]
        t = np.arange(n) / 16000.0
        f = 300 + 2500 * (0.5 + 0.5 * np.sin(2 * np.pi * (0.3 + 0.1 * i) * t))
        x = 5000 * np.sin(2 * np.pi * np.cumsum(f) / 16000.0) * (0.3 + 0.7 * (np.sin(2 * np.pi * 2.5 * t) > 0))
        x = x + 60 * rng.standard_normal(n)
        path = str(tmp_path / ("u%02d.wav" % i))
        write_wav(path, np.clip(np.round(x), -32768, 32767), 2)
        lines.append("utt%02d %s\n" % (i, path))
    lines.append("hello %s\n" % os.path.join(GOLDEN, "en-us-hello.wav"))
    scp = str(tmp_path / "many.scp")
    with open(scp, "w") as f:
        f.writelines(lines)
    ref = run(BIN_REF, setup["conf"], scp)
    assert ref.returncode == 0, ref.stdout + ref.stderr
    for threads in ("1", "6"):
        gpu = subprocess.run([BIN_BATCH, setup["conf"], scp, threads], capture_output=True, text=True)
        assert gpu.returncode == 0, gpu.stdout + gpu.stderr
        assert gpu.stdout == ref.stdout, threads
    assert len(ref.stdout.strip().splitlines()) == 15
