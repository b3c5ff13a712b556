"""CPU-side checks of the drop-in boundary: libce_gpu.so loads, exports every symbol that
include/ce_gpu.h declares, and every compute entry point fails loudly (no CPU fallback) when
there is no CUDA device.  No compute calls succeed without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from catears_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ce_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ce_gpu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = api.lib()
    names = _declared()
    assert len(names) >= 19
    for n in names:
        assert hasattr(L, n), "libce_gpu.so does not export %s" % n
    assert sorted(api.EXPORTS) == names


def test_version_and_error_string():
    L = api.lib()
    assert L.ce_gpu_version() >= 100
    assert isinstance(api.last_error(), str)


def test_frame_offsets_host_logic():
    """CalcNumFrames, src/fbank.cc:35-42: T = n < 400 ? 0 : 1 + (n - 400) / 160."""
    sizes = [0, 1, 399, 400, 401, 559, 560, 7802, 160000]
    off = np.concatenate([[0], np.cumsum(sizes)])
    fo = api.frame_offsets(off)
    want = [0 if n < 400 else 1 + (n - 400) // 160 for n in sizes]
    assert list(np.diff(fo)) == want
    assert want[-1] == 998 and want[-2] == 47
    with pytest.raises(api.CeGpuError):
        api.frame_offsets([0, 10, 5])


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path cannot be exercised")
    assert api.device_count() == 0
    pcm = np.zeros(1600, np.int16)
    with pytest.raises(api.CeGpuError, match="no CUDA device|CUDA"):
        api.fbank(pcm)
    with pytest.raises(api.CeGpuError):
        api.cmvn(np.ones(41, np.float32), np.zeros((3, 40), np.float32))
    with pytest.raises(api.CeGpuError):
        api.gemm_u8(np.zeros((4, 4), np.uint8), 1.0, 0, np.zeros((4, 4), np.uint8), 1.0, 0)
    with pytest.raises(api.CeGpuError):
        api.quantize(np.zeros((4, 4), np.float32))


def test_model_load_errors(tmp_path, small_model):
    """Status::IOError / Corruption paths of AcousticModel::Read (src/am.cc:26-64) surface as a
    NULL handle + message; without a device the loader stops at the device check."""
    with pytest.raises(api.CeGpuError, match="unable to open"):
        api.AcousticModelGpu(nnet=str(tmp_path / "missing.nnet"), prior=small_model["prior"],
                             left_context=13, right_context=13)
    bad = tmp_path / "bad.nnet"
    bad.write_bytes(b"NN01" + b"\0" * 12)
    with pytest.raises(api.CeGpuError, match="section name mismatch"):
        api.AcousticModelGpu(nnet=str(bad), prior=small_model["prior"], left_context=13,
                             right_context=13)
    conf = tmp_path / "x.conf"
    conf.write_text("nnet = a = b\n")
    with pytest.raises(api.CeGpuError, match="Unexpected line"):
        api.AcousticModelGpu(config=str(conf))
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(api.CeGpuError, match="no CUDA device"):
            api.AcousticModelGpu(config=small_model["conf"])


def test_null_handles_are_rejected_not_dereferenced():
    """Every handle-taking entry point of the decoder-feed and streaming additions answers a NULL
    handle with an error code and a message (the reference asserts; here it is CE_GPU_EINVAL)."""
    L = api.lib()
    ids = (C.c_int32 * 2)(0, 1)
    assert L.ce_gpu_model_set_output(None, 1, ids, 2) < 0
    assert "null model" in api.last_error()
    assert L.ce_gpu_model_output_width(None) < 0
    assert L.ce_gpu_model_set_rows_callback(None, api.ROWS_READY_FN(0), None) < 0
    assert not L.ce_gpu_streams_create(None, 4)
    assert "bad arguments" in api.last_error()
    assert L.ce_gpu_streams_open(None) < 0
    slots = (C.c_int * 1)(0)
    cnt = (C.c_int * 1)(160)
    assert L.ce_gpu_streams_rows_ready(None, slots, 1, cnt, None) < 0
    off = (C.c_int64 * 2)()
    assert L.ce_gpu_streams_process(None, slots, 1, None, cnt, None, None, 0, off, None) < 0
    L.ce_gpu_streams_free(None)                              # like free(NULL)


def test_corrupt_model_files_are_errors_not_crashes(tmp_path):
    """Header dimensions a file cannot back (2^31-ish rows, truncated data) come back as a load error
    with a message -- no allocation is driven by them and no C++ exception crosses the C ABI."""
    import struct

    import numpy as np
    from catears_b200 import formats as F
    prior = str(tmp_path / "p.prior")
    F.write_vector(prior, np.full(8, 0.125, np.float32))
    good = str(tmp_path / "good.nnet")
    F.write_nnet(good, [{"type": F.LINEAR, "W": np.zeros((40, 8), np.float32), "b": np.zeros(8, np.float32)}], 0, 0)
    data = open(good, "rb").read()
    cases = {
        "huge_mat": data[:16] + b"LAY0" + struct.pack("<i", 0) + b"MAT0" + struct.pack("<iii", 8, 2000000000, 2000000000),
        "huge_vec": data[:16] + b"LAY0" + struct.pack("<i", 7) + b"VEC0" + struct.pack("<ii", 4 * 500000000 + 4, 500000000),
        "truncated": data[:len(data) // 2],
        "bad_magic": b"XXXX" + data[4:],
        "splice_neg": data[:16] + b"LAY0" + struct.pack("<ii", 6, -5),
    }
    for name, blob in cases.items():
        path = str(tmp_path / (name + ".nnet"))
        open(path, "wb").write(blob)
        with pytest.raises(api.CeGpuError):
            api.AcousticModelGpu(nnet=path, prior=prior)
        assert api.last_error() != ""
