import os
import struct
import sys
import wave

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def read_wav(name):
    w = wave.open(os.path.join(GOLDEN, name))
    return np.frombuffer(w.readframes(w.getnframes()), np.int16).copy()


def read_vec0(path):
    d = open(path, "rb").read()
    assert d[:4] == b"VEC0"
    nbytes, dim = struct.unpack("<ii", d[4:12])
    assert nbytes == 4 * dim + 4
    return np.frombuffer(d[12:12 + 4 * dim], np.float32).copy()


@pytest.fixture(scope="session")
def golden():
    g = {
        "hello_pcm": read_wav("en-us-hello.wav"),
        "cat_pcm": read_wav("en-us-cat.wav"),
        "kaldi_fbank": np.loadtxt(os.path.join(GOLDEN, "fbankmat_en-us-hello.wav.txt"),
                                  dtype=np.float32).reshape(-1, 40),
        "kaldi_cmvn": np.loadtxt(os.path.join(GOLDEN, "fbankcmvnmat_en-us-hello.wav.txt"),
                                 dtype=np.float32).reshape(-1, 40),
        "cmvn_stats": read_vec0(os.path.join(GOLDEN, "cmvn_stats.bin")),
        "srfft": dict(np.load(os.path.join(GOLDEN, "srfft_kat.npz"))),
        "ref": dict(np.load(os.path.join(GOLDEN, "ref_vectors.npz"))),
    }
    return g


@pytest.fixture(scope="session")
def port():
    from oracle.port import Port
    return Port()


@pytest.fixture(scope="session")
def ref():
    """The reference compiled unmodified (oracle/_ref); None where it was never built."""
    from oracle import ref as R
    if not R.available():
        return None
    r = R.Ref()
    r.set_sgemm("inorder")
    return r


@pytest.fixture(scope="session")
def small_model(tmp_path_factory):
    """The 1/16-width TDNN the golden vectors were made with (tests/golden/make_golden.py)."""
    from catears_b200 import synth
    d = tmp_path_factory.mktemp("small_model")
    return synth.write_model(str(d), name="small", hidden=64, num_pdfs=96, seed=4321)
