"""ctypes view of oracle/_ref/libce_ref*.so -- TEST INFRASTRUCTURE ONLY.

The library is the UNMODIFIED reference (ishine/CatEars) compiled from
/root/reference by oracle/Makefile behind the C ABI of oracle/ref_shim.cc.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product (catears_b200) never does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def lib_path(variant=""):
    return os.path.join(_REF_DIR, "libce_ref%s.so" % variant)


def available(variant=""):
    return os.path.exists(lib_path(variant))


def find_openblas():
    """The OpenBLAS 0.3.15 bundled in the opencv wheel, or None."""
    import glob
    import sys
    for sp in sys.path:
        hits = glob.glob(os.path.join(sp, "opencv_python_headless.libs", "libopenblas*.so"))
        if hits:
            return hits[0]
    return None


class Ref:
    """variant: "" (PK_FBANK_DIM=40), "80" (80 mel bins), "_sse4" (gemmlowp SSE4 kernel)."""

    def __init__(self, variant=""):
        path = lib_path(variant)
        if not os.path.exists(path):
            raise FileNotFoundError(
                "%s missing: run `make -C oracle ref` where /root/reference exists" % path)
        L = C.CDLL(path)
        self.L = L
        L.ref_fbank_dim.restype = C.c_int
        L.ref_srfft.argtypes = [_f32p, C.c_int]
        L.ref_fbank_pcm16.argtypes = [_i16p, C.c_int, _f32p, C.c_int]
        L.ref_fbank_stream.argtypes = [C.c_char_p, C.c_int, C.c_int, _f32p, C.c_int]
        L.ref_cmvn.argtypes = [_f32p, _f32p, C.c_int, _f32p]
        L.ref_nnet_open.restype = C.c_void_p
        L.ref_nnet_open.argtypes = [C.c_char_p]
        L.ref_nnet_close.argtypes = [C.c_void_p]
        L.ref_nnet_propagate.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, _f32p, C.c_long,
                                         C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_am_open.restype = C.c_void_p
        L.ref_am_open.argtypes = [C.c_char_p]
        L.ref_am_close.argtypes = [C.c_void_p]
        L.ref_am_num_pdfs.argtypes = [C.c_void_p]
        L.ref_am_forward.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, _f32p, C.c_int,
                                     C.POINTER(C.c_int)]
        L.ref_quantize.argtypes = [_f32p, C.c_int, C.c_int, _u8p, C.POINTER(C.c_float),
                                   C.POINTER(C.c_int32)]
        L.ref_gemm_u8.argtypes = [_u8p, C.c_float, C.c_int32, _u8p, C.c_float, C.c_int32,
                                  C.c_int, C.c_int, C.c_int, _f32p, C.c_void_p]
        L.ref_u8_open.restype = C.c_void_p
        L.ref_u8_open.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        L.ref_u8_close.argtypes = [C.c_void_p]
        L.ref_u8_forward.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, _f32p, C.c_long,
                                     C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int,
                                     C.c_void_p, C.c_long, C.POINTER(C.c_int),
                                     C.POINTER(C.c_int)]
        if hasattr(L, "ref_u8_forward_trace"):
            L.ref_u8_forward_trace.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, _f32p, C.c_long,
                                               C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_long,
                                               C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_set_sgemm_backend.argtypes = [C.c_int, C.c_char_p]
        self.mel = L.ref_fbank_dim()

    # -- front-end ---------------------------------------------------------
    def srfft(self, x):
        x = np.ascontiguousarray(x, np.float32).copy()
        self.L.ref_srfft(x, x.size)
        return x

    def fbank(self, pcm):
        pcm = np.ascontiguousarray(pcm, np.int16)
        cap = max(1, pcm.size // 160 + 1)
        out = np.zeros((cap, self.mel), np.float32)
        n = self.L.ref_fbank_pcm16(pcm, pcm.size, out, cap)
        if n < 0:
            raise RuntimeError("ref_fbank_pcm16 failed: %d" % n)
        return out[:n].copy()

    def fbank_stream(self, pcm_bytes, chunk):
        cap = max(1, len(pcm_bytes) // 320 + 1)
        out = np.zeros((cap, self.mel), np.float32)
        n = self.L.ref_fbank_stream(pcm_bytes, len(pcm_bytes), chunk, out, cap)
        if n < 0:
            raise RuntimeError("ref_fbank_stream failed: %d" % n)
        return out[:n].copy()

    def cmvn(self, global_stats, feats):
        feats = np.ascontiguousarray(feats, np.float32)
        g = np.ascontiguousarray(global_stats, np.float32)
        assert g.size == self.mel + 1 and feats.shape[1] == self.mel
        out = np.zeros_like(feats)
        self.L.ref_cmvn(g, feats, feats.shape[0], out)
        return out

    # -- nnet / AM -----------------------------------------------------------
    def set_sgemm(self, backend, path=None):
        """backend: "inorder" (deterministic fp32 k-ascending loop) or "openblas"."""
        if backend == "inorder":
            return self.L.ref_set_sgemm_backend(0, None) == 0
        path = path or find_openblas()
        if path is None:
            return False
        # The wheel's OpenBLAS depends on the libgfortran/libquadmath shipped beside it.
        import glob
        d = os.path.dirname(path)
        for dep in sorted(glob.glob(os.path.join(d, "libquadmath*"))) + \
                sorted(glob.glob(os.path.join(d, "libgfortran*"))):
            try:
                C.CDLL(dep, mode=C.RTLD_GLOBAL)
            except OSError:
                pass
        return self.L.ref_set_sgemm_backend(1, path.encode()) == 0

    def nnet_propagate(self, nnet_path, x, out_cols_max=8192):
        h = self.L.ref_nnet_open(nnet_path.encode())
        if not h:
            raise RuntimeError("ref_nnet_open failed for %s" % nnet_path)
        try:
            x = np.ascontiguousarray(x, np.float32)
            cap = x.shape[0] * out_cols_max
            out = np.zeros(cap, np.float32)
            r, c = C.c_int(), C.c_int()
            rc = self.L.ref_nnet_propagate(h, x, x.shape[0], x.shape[1], out, cap,
                                           C.byref(r), C.byref(c))
            if rc != 0:
                raise RuntimeError("ref_nnet_propagate: %d" % rc)
            return out[: r.value * c.value].reshape(r.value, c.value).copy()
        finally:
            self.L.ref_nnet_close(h)

    def am_forward(self, conf_path, feats):
        """AcousticModel::Process per frame + EndOfStream, rows concatenated."""
        h = self.L.ref_am_open(conf_path.encode())
        if not h:
            raise RuntimeError("ref_am_open failed for %s" % conf_path)
        try:
            feats = np.ascontiguousarray(feats, np.float32)
            npdf = self.L.ref_am_num_pdfs(h)
            out = np.zeros((feats.shape[0] + 1, max(npdf, 1)), np.float32)
            cols = C.c_int()
            n = self.L.ref_am_forward(h, feats, feats.shape[0], feats.shape[1], out,
                                      out.shape[0], C.byref(cols))
            if n < 0:
                raise RuntimeError("ref_am_forward: %d" % n)
            assert n == 0 or cols.value == out.shape[1], (cols.value, out.shape)
            return out[:n].copy()
        finally:
            self.L.ref_am_close(h)

    def quantize(self, x):
        x = np.ascontiguousarray(x, np.float32)
        q = np.zeros(x.shape, np.uint8)
        s, z = C.c_float(), C.c_int32()
        self.L.ref_quantize(x, x.shape[0], x.shape[1], q, C.byref(s), C.byref(z))
        return q, np.float32(s.value), int(z.value)

    def gemm_u8(self, a, sa, za, b, sb, zb, want_acc=True):
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        m, k = a.shape
        k2, n = b.shape
        assert k == k2
        c = np.zeros((m, n), np.float32)
        acc = np.zeros((m, n), np.int32) if want_acc else None
        self.L.ref_gemm_u8(a, C.c_float(sa), za, b, C.c_float(sb), zb, m, n, k, c,
                           acc.ctypes.data if want_acc else None)
        return c, acc

    def u8_forward(self, nnet_path, prior_path, left, right, feats, dump_layer=-1,
                   out_cols_max=8192):
        """int8 composition. Returns (loglik, acc or None)."""
        h = self.L.ref_u8_open(nnet_path.encode(), prior_path.encode(), left, right)
        if not h:
            raise RuntimeError("ref_u8_open failed")
        try:
            feats = np.ascontiguousarray(feats, np.float32)
            T = feats.shape[0]
            rows = T + left + right
            cap = rows * out_cols_max
            out = np.zeros(cap, np.float32)
            r, c = C.c_int(), C.c_int()
            ar, ac = C.c_int(), C.c_int()
            acc = np.zeros(cap, np.int32) if dump_layer >= 0 else None
            rc = self.L.ref_u8_forward(h, feats, T, feats.shape[1], out, cap, C.byref(r),
                                       C.byref(c), dump_layer,
                                       acc.ctypes.data if acc is not None else None, cap,
                                       C.byref(ar), C.byref(ac))
            if rc != 0:
                raise RuntimeError("ref_u8_forward: %d" % rc)
            y = out[: r.value * c.value].reshape(r.value, c.value).copy()
            if acc is not None:
                acc = acc[: ar.value * ac.value].reshape(ar.value, ac.value).copy()
            return y, acc
        finally:
            self.L.ref_u8_close(h)

    def u8_forward_trace(self, nnet_path, prior_path, left, right, feats, want_acc=True,
                         out_cols_max=8192, max_layers=64):
        """int8 composition with every Linear layer traced: returns (loglik, [(scale, zero_point)],
        [acc per Linear layer] or None)."""
        h = self.L.ref_u8_open(nnet_path.encode(), prior_path.encode(), left, right)
        if not h:
            raise RuntimeError("ref_u8_open failed")
        try:
            feats = np.ascontiguousarray(feats, np.float32)
            T = feats.shape[0]
            rows = T + left + right
            cap = rows * out_cols_max
            out = np.zeros(cap, np.float32)
            r, c = C.c_int(), C.c_int()
            sc = np.zeros(max_layers, np.float32)
            zp = np.zeros(max_layers, np.int32)
            off = np.zeros(max_layers + 1, np.int64)
            ar = np.zeros(max_layers, np.int32)
            ac = np.zeros(max_layers, np.int32)
            acc_cap = rows * out_cols_max * 3 if want_acc else 0
            acc = np.zeros(acc_cap, np.int32) if want_acc else None
            n = self.L.ref_u8_forward_trace(h, feats, T, feats.shape[1], out, cap, C.byref(r),
                                            C.byref(c), max_layers, sc.ctypes.data, zp.ctypes.data,
                                            acc.ctypes.data if want_acc else None, acc_cap,
                                            off.ctypes.data, ar.ctypes.data, ac.ctypes.data)
            if n < 0:
                raise RuntimeError("ref_u8_forward_trace: %d" % n)
            y = out[: r.value * c.value].reshape(r.value, c.value).copy()
            accs = None
            if want_acc:
                accs = [acc[off[i]:off[i + 1]].reshape(ar[i], ac[i]).copy() for i in range(n)]
            return y, list(zip(sc[:n].tolist(), zp[:n].tolist())), accs
        finally:
            self.L.ref_u8_close(h)
