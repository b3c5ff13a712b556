"""ctypes binding of catears_b200/libce_gpu.so (include/ce_gpu.h) -- the same stub an embedder of
the reference would write (INTEGRATION.md shows the C++ one).

There is no Python or CPU implementation of any stage here: every function forwards to the
CUDA library and raises CeGpuError with ce_gpu_last_error() when the library reports a failure
(including "no CUDA device").  Arrays may be numpy arrays (host memory) or torch CUDA tensors
(device memory, passed by data_ptr()).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CE_GPU_LIB: another build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("CE_GPU_LIB") or os.path.join(_HERE, "libce_gpu.so")

PRECISION_INT8, PRECISION_BF16, PRECISION_FP32, PRECISION_TF32, PRECISION_BF16X3 = 0, 1, 2, 3, 4
PRECISIONS = {"int8": 0, "bf16": 1, "fp32": 2, "tf32": 3, "bf16x3": 4}

FRAME_LEN, FRAME_SHIFT = 400, 160

EXPORTS = [
    "ce_gpu_last_error", "ce_gpu_device_count", "ce_gpu_version", "ce_gpu_model_load",
    "ce_gpu_model_load_config", "ce_gpu_model_free", "ce_gpu_model_info", "ce_gpu_frame_offsets",
    "ce_gpu_fbank", "ce_gpu_cmvn", "ce_gpu_rfft512", "ce_gpu_nnet", "ce_gpu_forward",
    "ce_gpu_nnet_keep_acc", "ce_gpu_nnet_get_acc", "ce_gpu_quantize", "ce_gpu_gemm_u8",
    "ce_gpu_gemm_f32", "ce_gpu_launch_count", "ce_gpu_profile_enable", "ce_gpu_profile_read",
    "ce_gpu_profile_trace", "ce_gpu_selftest_quantizer", "ce_gpu_cmvn_stream",
    "ce_gpu_partition", "ce_gpu_time_shards", "ce_gpu_model_set_output",
    "ce_gpu_model_output_width", "ce_gpu_model_set_rows_callback", "ce_gpu_streams_create",
    "ce_gpu_streams_free", "ce_gpu_streams_open", "ce_gpu_streams_rows_ready", "ce_gpu_streams_process",
    "ce_gpu_nnet_get_qparams", "ce_gpu_nnet_chunks", "ce_gpu_host_alloc", "ce_gpu_host_free",
    "ce_gpu_streams_call_stats",
]
ROWS_READY_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64)
OUTPUT_MODES = {"dense": 0, "subset": 1, "topk": 2}
# ce_gpu_scored_pdf_t: one entry of a top-k row
SCORED_PDF = np.dtype([("loglik", np.float32), ("pdf", np.int32)])
PROFILE_CATEGORIES = ["fbank", "cmvn", "gemm", "quantize", "finalize", "other"]


class CeGpuError(RuntimeError):
    pass


_lib = None


def lib():
    """Loads libce_gpu.so; fails loudly when it has not been built (`make lib`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CeGpuError("%s is missing: run `make lib` (or __graft_entry__.build()); there is no "
                         "fallback implementation" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i64p = C.c_void_p, C.POINTER(C.c_int64)
    L.ce_gpu_last_error.restype = C.c_char_p
    L.ce_gpu_model_load.restype = vp
    L.ce_gpu_model_load.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.ce_gpu_model_load_config.restype = vp
    L.ce_gpu_model_load_config.argtypes = [C.c_char_p, C.c_int, C.c_int]
    L.ce_gpu_model_free.argtypes = [vp]
    L.ce_gpu_model_info.argtypes = [vp] + [C.POINTER(C.c_int)] * 6
    L.ce_gpu_frame_offsets.restype = C.c_int64
    L.ce_gpu_frame_offsets.argtypes = [i64p, C.c_int, i64p]
    L.ce_gpu_fbank.argtypes = [vp, i64p, C.c_int, C.c_int, vp, C.c_int, vp]
    L.ce_gpu_cmvn.argtypes = [vp, vp, i64p, C.c_int, C.c_int, vp, C.c_int, vp]
    L.ce_gpu_cmvn_stream.argtypes = [vp, vp, i64p, C.POINTER(C.c_int32), i64p, vp, C.c_int, C.c_int, vp,
                                     C.c_int, vp]
    L.ce_gpu_rfft512.argtypes = [vp, C.c_int, vp, C.c_int, vp]
    L.ce_gpu_nnet.argtypes = [vp, vp, i64p, C.c_int, vp, vp, vp]
    L.ce_gpu_forward.argtypes = [vp, vp, i64p, C.c_int, vp, vp, i64p, vp]
    L.ce_gpu_nnet_chunks.argtypes = [vp, vp, i64p, C.c_int, vp, vp, vp]
    L.ce_gpu_nnet_keep_acc.argtypes = [vp, C.c_int]
    L.ce_gpu_model_set_output.argtypes = [vp, C.c_int, C.POINTER(C.c_int32), C.c_int]
    L.ce_gpu_model_output_width.argtypes = [vp]
    L.ce_gpu_model_set_rows_callback.argtypes = [vp, ROWS_READY_FN, vp]
    ip = C.POINTER(C.c_int)
    L.ce_gpu_streams_create.restype = vp
    L.ce_gpu_streams_create.argtypes = [vp, C.c_int]
    L.ce_gpu_streams_free.argtypes = [vp]
    L.ce_gpu_streams_open.argtypes = [vp]
    L.ce_gpu_streams_rows_ready.restype = C.c_int64
    L.ce_gpu_streams_rows_ready.argtypes = [vp, ip, C.c_int, ip, C.POINTER(C.c_ubyte)]
    L.ce_gpu_streams_process.argtypes = [vp, ip, C.c_int, C.POINTER(vp), ip, C.POINTER(C.c_ubyte), vp,
                                         C.c_int64, i64p, vp]
    L.ce_gpu_streams_call_stats.argtypes = [vp, i64p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
    L.ce_gpu_nnet_get_acc.argtypes = [vp, C.c_int, vp, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.ce_gpu_nnet_get_qparams.argtypes = [vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int]
    L.ce_gpu_quantize.argtypes = [vp, C.c_int64, C.c_int, vp, C.POINTER(C.c_float),
                                  C.POINTER(C.c_int32), C.c_int, vp]
    L.ce_gpu_gemm_u8.argtypes = [vp, C.c_float, C.c_int32, vp, C.c_float, C.c_int32, C.c_int, C.c_int,
                                 C.c_int, vp, vp, C.c_int, vp]
    L.ce_gpu_gemm_f32.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp]
    L.ce_gpu_launch_count.restype = C.c_int64
    L.ce_gpu_launch_count.argtypes = [C.c_int]
    L.ce_gpu_partition.argtypes = [i64p, C.c_int, C.c_int, C.POINTER(C.c_int32)]
    L.ce_gpu_time_shards.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, i64p, i64p, i64p, i64p]
    L.ce_gpu_profile_enable.argtypes = [C.c_int]
    L.ce_gpu_profile_read.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.ce_gpu_selftest_quantizer.restype = C.c_int64
    L.ce_gpu_selftest_quantizer.argtypes = [C.c_int64, C.c_uint64, C.c_int]
    L.ce_gpu_profile_trace.argtypes = [C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                       C.POINTER(C.c_double)]
    _lib = L
    return L


def last_error():
    return lib().ce_gpu_last_error().decode("utf-8", "replace")


def _check(rc, what):
    if rc < 0:
        raise CeGpuError("%s failed (%d): %s" % (what, rc, last_error()))
    return rc


def _ptr(x):
    """Raw address of a numpy array (host) or torch tensor (host or device); None -> NULL."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return x.data_ptr()
    raise TypeError("expected numpy array or torch tensor, got %r" % type(x))


def _offsets(off):
    off = np.ascontiguousarray(off, np.int64)
    return off, off.ctypes.data_as(C.POINTER(C.c_int64))


def _stream(stream):
    if stream is None:
        return None
    return getattr(stream, "cuda_stream", stream)


def device_count():
    return lib().ce_gpu_device_count()


def launch_count(reset=False):
    return lib().ce_gpu_launch_count(1 if reset else 0)


def profile_enable(on=True):
    _check(lib().ce_gpu_profile_enable(1 if on else 0), "ce_gpu_profile_enable")


def profile_read():
    """{category: (milliseconds, launches)} of the kernels recorded since the last read."""
    ms = (C.c_double * 6)()
    n = (C.c_int64 * 6)()
    _check(lib().ce_gpu_profile_read(ms, n), "ce_gpu_profile_read")
    return {k: (ms[i], n[i]) for i, k in enumerate(PROFILE_CATEGORIES)}


def selftest_quantizer(n=1 << 26, seed=1, device=0):
    """Disagreements between the production quantiser arithmetic and plain IEEE (0 = exact)."""
    return _check(lib().ce_gpu_selftest_quantizer(n, seed, device), "ce_gpu_selftest_quantizer")


def profile_trace(cap=65536):
    """[(category, begin_ms, end_ms)] of every timed launch scope since the last read, in launch
    order; times share one clock across the chunk streams."""
    cat = (C.c_int32 * cap)()
    t0 = (C.c_double * cap)()
    t1 = (C.c_double * cap)()
    n = _check(lib().ce_gpu_profile_trace(cap, cat, t0, t1), "ce_gpu_profile_trace")
    return [(PROFILE_CATEGORIES[cat[i]], t0[i], t1[i]) for i in range(n)]


def frame_offsets(sample_offsets):
    """Snip-edges frame bookkeeping (src/fbank.cc:35-42): offsets [n+1] -> frame offsets [n+1]."""
    off, p = _offsets(sample_offsets)
    out = np.zeros(off.size, np.int64)
    _check(lib().ce_gpu_frame_offsets(p, off.size - 1, out.ctypes.data_as(C.POINTER(C.c_int64))),
           "ce_gpu_frame_offsets")
    return out


def partition(frame_offsets_, n_parts):
    """Contiguous, frame-balanced utterance groups, one per GPU: returns part_begin [n_parts+1]."""
    off, p = _offsets(frame_offsets_)
    out = np.zeros(n_parts + 1, np.int32)
    _check(lib().ce_gpu_partition(p, off.size - 1, n_parts, out.ctypes.data_as(C.POINTER(C.c_int32))),
           "ce_gpu_partition")
    return out


def time_shards(total_frames, n_parts, left_context, right_context, cmvn_history=600):
    """Long-form time shards with halos: arrays keep_begin, keep_end, feed_begin, feed_end."""
    arrs = [np.zeros(n_parts, np.int64) for _ in range(4)]
    ptrs = [a.ctypes.data_as(C.POINTER(C.c_int64)) for a in arrs]
    _check(lib().ce_gpu_time_shards(total_frames, n_parts, left_context, right_context, cmvn_history, *ptrs),
           "ce_gpu_time_shards")
    return arrs


def fbank(pcm, sample_offsets=None, num_mel=40, out=None, device=0, stream=None):
    """Fbank::Process for a ragged batch of int16 PCM.  Returns feats [total_frames x num_mel]."""
    if sample_offsets is None:
        sample_offsets = [0, pcm.shape[0]]
    off, p = _offsets(sample_offsets)
    foff = frame_offsets(off)
    if out is None:
        out = np.zeros((int(foff[-1]), num_mel), np.float32)
    _check(lib().ce_gpu_fbank(_ptr(pcm), p, off.size - 1, num_mel, _ptr(out), device, _stream(stream)),
           "ce_gpu_fbank")
    return out


def cmvn(global_stats, feats, frame_offsets_=None, out=None, device=0, stream=None):
    """CMVN::GetFrame for frames 0..T-1 of every utterance."""
    g = np.ascontiguousarray(global_stats, np.float32)
    num_mel = g.size - 1
    if frame_offsets_ is None:
        frame_offsets_ = [0, feats.shape[0]]
    off, p = _offsets(frame_offsets_)
    if out is None:
        out = np.zeros((int(off[-1]), num_mel), np.float32)
    _check(lib().ce_gpu_cmvn(g.ctypes.data, _ptr(feats), p, off.size - 1, num_mel, _ptr(out), device,
                             _stream(stream)), "ce_gpu_cmvn")
    return out


def cmvn_stream(global_stats, feats, frame_offsets_, n_hist, t_base, state, device=0):
    """CMVN continued across calls: see ce_gpu_cmvn_stream.  `state` ([n x mel] float32) is updated in
    place; returns the normalised NEW frames, packed by utterance."""
    g = np.ascontiguousarray(global_stats, np.float32)
    num_mel = g.size - 1
    off, p = _offsets(frame_offsets_)
    nh = np.ascontiguousarray(n_hist, np.int32)
    tb, ptb = _offsets(t_base)
    assert state.dtype == np.float32 and state.flags["C_CONTIGUOUS"]
    n_new = int((off[1:] - off[:-1]).sum() - nh.sum())
    out = np.zeros((n_new, num_mel), np.float32)
    _check(lib().ce_gpu_cmvn_stream(g.ctypes.data, _ptr(feats), p, nh.ctypes.data_as(C.POINTER(C.c_int32)), ptb,
                                    state.ctypes.data, off.size - 1, num_mel, _ptr(out), device, None),
           "ce_gpu_cmvn_stream")
    return out


def rfft512(x, device=0):
    x = np.ascontiguousarray(x, np.float32).reshape(-1, 512)
    out = np.zeros_like(x)
    _check(lib().ce_gpu_rfft512(x.ctypes.data, x.shape[0], out.ctypes.data, device, None), "ce_gpu_rfft512")
    return out


def quantize(x, device=0):
    """Quantize (src/matrix.cc:366-387): returns (codes u8, scale, zero_point)."""
    x = np.ascontiguousarray(x, np.float32)
    q = np.zeros(x.shape, np.uint8)
    s, z = C.c_float(), C.c_int32()
    _check(lib().ce_gpu_quantize(x.ctypes.data, x.shape[0], x.shape[1], q.ctypes.data, C.byref(s),
                                 C.byref(z), device, None), "ce_gpu_quantize")
    return q, np.float32(s.value), int(z.value)


def gemm_u8(a, sa, za, b, sb, zb, want_acc=True, device=0):
    """MatMat_U8U8F32 (src/matrix.cc:389-420): returns (C fp32, int32 accumulators or None)."""
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    m, k = a.shape
    k2, n = b.shape
    assert k == k2
    c = np.zeros((m, n), np.float32)
    acc = np.zeros((m, n), np.int32) if want_acc else None
    _check(lib().ce_gpu_gemm_u8(a.ctypes.data, C.c_float(sa), za, b.ctypes.data, C.c_float(sb), zb, m, n,
                                k, c.ctypes.data, acc.ctypes.data if want_acc else None, device, None),
           "ce_gpu_gemm_u8")
    return c, acc


def gemm_f32(a, b, precision="fp32", device=0):
    """MatMat (src/matrix.cc:300-323) on the tensor cores."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    m, k = a.shape
    k2, n = b.shape
    assert k == k2
    c = np.zeros((m, n), np.float32)
    _check(lib().ce_gpu_gemm_f32(a.ctypes.data, b.ctypes.data, m, n, k, c.ctypes.data,
                                 PRECISIONS[precision], device, None), "ce_gpu_gemm_f32")
    return c


class AcousticModelGpu:
    """Python face of ce_gpu_model_t: AcousticModel (src/am.h:22-73) evaluated for whole
    utterances on one GPU.  `Read`-style construction from a config file or explicit paths."""

    def __init__(self, nnet=None, prior=None, left_context=0, right_context=0, cmvn_stats=None,
                 precision="int8", device=0, config=None):
        L = lib()
        prec = PRECISIONS[precision] if isinstance(precision, str) else int(precision)
        if config is not None:
            h = L.ce_gpu_model_load_config(config.encode(), prec, device)
        else:
            h = L.ce_gpu_model_load(nnet.encode(), prior.encode(),
                                    cmvn_stats.encode() if cmvn_stats else None, left_context,
                                    right_context, prec, device)
        if not h:
            raise CeGpuError("model load failed: %s" % last_error())
        self._h = h
        v = [C.c_int() for _ in range(6)]
        _check(L.ce_gpu_model_info(h, *[C.byref(x) for x in v]), "ce_gpu_model_info")
        (self.num_pdfs, self.left_context, self.right_context, self.feat_dim, self.precision,
         self.device) = [x.value for x in v]

    def close(self):
        if getattr(self, "_h", None):
            lib().ce_gpu_model_free(self._h)
            self._h = None

    __del__ = close

    def set_output(self, mode="dense", pdf_ids=None, k=0):
        """What a row of `loglik` is from now on (ce_gpu_model_set_output): all pdfs, the columns
        `pdf_ids`, or the `k` best as SCORED_PDF entries."""
        ids, n = None, int(k)
        if mode == "subset":
            ids = np.ascontiguousarray(pdf_ids, np.int32)
            n = ids.size
        _check(lib().ce_gpu_model_set_output(
            self._h, OUTPUT_MODES[mode], ids.ctypes.data_as(C.POINTER(C.c_int32)) if ids is not None else None,
            n), "ce_gpu_model_set_output")
        self._out_mode = mode

    def set_rows_callback(self, fn):
        """fn(first_utt, n_utts, first_frame, n_frames) is called per chunk as soon as its rows
        are complete in the caller's host buffers (ce_gpu_model_set_rows_callback); None removes
        it.  Runs on a CUDA runtime thread: no CUDA / ce_gpu calls inside."""
        cb = ROWS_READY_FN(lambda _u, a, b, c, d: fn(a, b, c, d)) if fn is not None else ROWS_READY_FN(0)
        _check(lib().ce_gpu_model_set_rows_callback(self._h, cb, None), "ce_gpu_model_set_rows_callback")
        self._rows_cb = cb                                 # keep the trampoline alive

    def output_width(self):
        return lib().ce_gpu_model_output_width(self._h)

    def _outputs(self, n_frames, want_loglik, want_argmax, loglik, argmax):
        if loglik is None and want_loglik:
            if getattr(self, "_out_mode", "dense") == "topk":
                loglik = np.zeros((n_frames, self.output_width() // 2), SCORED_PDF)
            else:
                loglik = np.zeros((n_frames, self.output_width()), np.float32)
        if argmax is None and want_argmax:
            argmax = np.zeros(n_frames, np.int32)
        return loglik, argmax

    def nnet(self, feats, frame_offsets_=None, want_loglik=True, want_argmax=True, loglik=None,
             argmax=None, stream=None):
        """Process + EndOfStream of every utterance: feats [frames x feat_dim] -> log-likelihoods."""
        if frame_offsets_ is None:
            frame_offsets_ = [0, feats.shape[0]]
        off, p = _offsets(frame_offsets_)
        loglik, argmax = self._outputs(int(off[-1]), want_loglik, want_argmax, loglik, argmax)
        _check(lib().ce_gpu_nnet(self._h, _ptr(feats), p, off.size - 1, _ptr(loglik), _ptr(argmax),
                                 _stream(stream)), "ce_gpu_nnet")
        return loglik, argmax

    def nnet_chunks(self, feats, block_offsets, want_argmax=True, stream=None):
        """ComputeBatch (src/am.cc:82-113) of many chunks: every block of rows carries its own context;
        returns the packed rows of all blocks (block rows - left - right each)."""
        off, p = _offsets(block_offsets)
        n_out = int(np.maximum(np.diff(off) - self.left_context - self.right_context, 0).sum())
        loglik, argmax = self._outputs(n_out, True, want_argmax, None, None)
        _check(lib().ce_gpu_nnet_chunks(self._h, _ptr(feats), p, off.size - 1, _ptr(loglik), _ptr(argmax),
                                        _stream(stream)), "ce_gpu_nnet_chunks")
        return loglik, argmax

    def forward(self, pcm, sample_offsets=None, want_loglik=True, want_argmax=True, loglik=None,
                argmax=None, stream=None):
        """PCM -> fbank -> [CMVN] -> AM.  Returns (loglik, argmax, frame_offsets)."""
        if sample_offsets is None:
            sample_offsets = [0, pcm.shape[0]]
        off, p = _offsets(sample_offsets)
        foff = frame_offsets(off)
        loglik, argmax = self._outputs(int(foff[-1]), want_loglik, want_argmax, loglik, argmax)
        fo = np.zeros(off.size, np.int64)
        _check(lib().ce_gpu_forward(self._h, _ptr(pcm), p, off.size - 1, _ptr(loglik), _ptr(argmax),
                                    fo.ctypes.data_as(C.POINTER(C.c_int64)), _stream(stream)),
               "ce_gpu_forward")
        return loglik, argmax, fo

    def keep_acc(self, linear_ordinal):
        _check(lib().ce_gpu_nnet_keep_acc(self._h, linear_ordinal), "ce_gpu_nnet_keep_acc")

    def get_acc(self, utt=0):
        r, c = C.c_int(), C.c_int()
        _check(lib().ce_gpu_nnet_get_acc(self._h, utt, None, 0, C.byref(r), C.byref(c)), "ce_gpu_nnet_get_acc")
        acc = np.zeros((r.value, c.value), np.int32)
        _check(lib().ce_gpu_nnet_get_acc(self._h, utt, acc.ctypes.data, acc.size, C.byref(r), C.byref(c)),
               "ce_gpu_nnet_get_acc")
        return acc


    def get_qparams(self, utt=0):
        """(scale[], zero_point[]) of the activation Quantize in front of every Linear layer."""
        sc = (C.c_float * 64)()
        zp = (C.c_int32 * 64)()
        n = _check(lib().ce_gpu_nnet_get_qparams(self._h, utt, sc, zp, 64), "ce_gpu_nnet_get_qparams")
        return np.array(sc[:n], np.float32), np.array(zp[:n], np.int32)


class StreamSet:
    """Live utterances with their state on the device (ce_gpu_streams_*)."""

    def __init__(self, model, max_streams):
        self._m = model                                    # keeps the model alive
        self._h = lib().ce_gpu_streams_create(model._h, max_streams)
        if not self._h:
            raise CeGpuError("ce_gpu_streams_create failed: %s" % last_error())

    def close(self):
        if getattr(self, "_h", None):
            lib().ce_gpu_streams_free(self._h)
            self._h = None

    __del__ = close

    def call_stats(self, reset=False):
        """(calls, enqueue_us, total_us) spent inside ce_gpu_streams_process so far."""
        n, a, b = C.c_int64(), C.c_double(), C.c_double()
        _check(lib().ce_gpu_streams_call_stats(self._h, C.byref(n), C.byref(a), C.byref(b), 1 if reset else 0),
               "ce_gpu_streams_call_stats")
        return n.value, a.value, b.value

    def open(self):
        slot = lib().ce_gpu_streams_open(self._h)
        _check(min(slot, 0), "ce_gpu_streams_open")
        return slot

    def _args(self, slots, pieces, eos):
        n = len(slots)
        sl = (C.c_int * n)(*slots)
        cnt = (C.c_int * n)(*[0 if p is None else int(p.size) for p in pieces])
        e = (C.c_ubyte * n)(*[1 if x else 0 for x in eos])
        return n, sl, cnt, e

    def rows_ready(self, slots, pieces, eos):
        n, sl, cnt, e = self._args(slots, pieces, eos)
        r = lib().ce_gpu_streams_rows_ready(self._h, sl, n, cnt, e)
        _check(min(int(r), 0), "ce_gpu_streams_rows_ready")
        return int(r)

    def process(self, slots, pieces, eos, rows_cap=None):
        """pieces[i]: int16 array of new samples of slots[i] (or None).  Returns one row matrix per
        slot (host)."""
        n, sl, cnt, e = self._args(slots, pieces, eos)
        keep = [None if p is None else np.ascontiguousarray(p, np.int16) for p in pieces]
        ptrs = (C.c_void_p * n)(*[None if p is None or p.size == 0 else p.ctypes.data for p in keep])
        cap = self.rows_ready(slots, pieces, eos) if rows_cap is None else rows_cap
        w = self._m.output_width()
        flat = np.zeros((max(cap, 1), w), np.float32)
        off = np.zeros(n + 1, np.int64)
        _check(lib().ce_gpu_streams_process(self._h, sl, n, ptrs, cnt, e, flat.ctypes.data, cap,
                                            off.ctypes.data_as(C.POINTER(C.c_int64)), None),
               "ce_gpu_streams_process")
        return [flat[off[i]:off[i + 1]] for i in range(n)]
