"""ctypes view of oracle/libce_oracle.so (the plain-C restatement, oracle/ce_oracle.c).

TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this; the product never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libce_oracle.so")

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")


def build(force=False):
    src = os.path.join(_HERE, "ce_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
    return _LIB


class Port:
    def __init__(self):
        L = C.CDLL(build())
        self.L = L
        L.orc_srfft.argtypes = [_f32p, C.c_int]
        L.orc_fbank_create.restype = C.c_void_p
        L.orc_fbank_create.argtypes = [C.c_int]
        L.orc_fbank_create2.restype = C.c_void_p
        L.orc_fbank_create2.argtypes = [C.c_int, C.c_int]
        L.orc_fbank_destroy.argtypes = [C.c_void_p]
        L.orc_fbank_filter.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), _f32p, C.c_int]
        L.orc_fbank_hamming.restype = C.POINTER(C.c_float)
        L.orc_fbank_hamming.argtypes = [C.c_void_p]
        L.orc_num_frames.argtypes = [C.c_int]
        L.orc_fbank.argtypes = [C.c_void_p, _i16p, C.c_int, _f32p]
        L.orc_cmvn.argtypes = [_f32p, _f32p, C.c_int, C.c_int, _f32p]
        L.orc_sgemm.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, _f32p]
        L.orc_quantize.argtypes = [_f32p, C.c_long, _u8p, C.POINTER(C.c_float),
                                   C.POINTER(C.c_int32)]
        L.orc_gemm_u8.argtypes = [_u8p, C.c_float, C.c_int32, _u8p, C.c_float, C.c_int32,
                                  C.c_int, C.c_int, C.c_int, _f32p, C.c_void_p]
        L.orc_nnet_open.restype = C.c_void_p
        L.orc_nnet_open.argtypes = [C.c_char_p]
        L.orc_nnet_close.argtypes = [C.c_void_p]
        L.orc_nnet_propagate.restype = C.POINTER(C.c_float)
        L.orc_nnet_propagate.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int,
                                         C.c_void_p]
        L.orc_am_forward.restype = C.POINTER(C.c_float)
        L.orc_am_forward.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int, _f32p,
                                     C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                     C.POINTER(C.c_int), C.c_int, C.c_void_p]
        L.orc_free.argtypes = [C.c_void_p]
        self._fb = {}

    def _fbank(self, mel):
        if mel not in self._fb:
            # mel == 40 keeps the reference's >=2-bins-per-filter assertion; other sizes are an
            # extension the reference cannot run (see ce_oracle.c orc_fbank_create2).
            h = self.L.orc_fbank_create2(mel, 1 if mel == 40 else 0)
            if not h:
                raise RuntimeError("orc_fbank_create(%d) failed" % mel)
            self._fb[mel] = h
        return self._fb[mel]

    def srfft(self, x):
        x = np.ascontiguousarray(x, np.float32).copy()
        self.L.orc_srfft(x, x.size)
        return x

    def hamming(self):
        p = self.L.orc_fbank_hamming(self._fbank(40))
        return np.ctypeslib.as_array(p, shape=(400,)).copy()

    def mel_filters(self, mel=40):
        """list of (offset, weights) per mel bin."""
        out = []
        buf = np.zeros(256, np.float32)
        for b in range(mel):
            off = C.c_int()
            w = self.L.orc_fbank_filter(self._fbank(mel), b, C.byref(off), buf, 256)
            out.append((off.value, buf[:w].copy()))
        return out

    def num_frames(self, n):
        return self.L.orc_num_frames(n)

    def fbank(self, pcm, mel=40):
        pcm = np.ascontiguousarray(pcm, np.int16)
        T = self.num_frames(pcm.size)
        out = np.zeros((max(T, 1), mel), np.float32)
        n = self.L.orc_fbank(self._fbank(mel), pcm, pcm.size, out)
        return out[:n].copy()

    def cmvn(self, g, feats):
        feats = np.ascontiguousarray(feats, np.float32)
        g = np.ascontiguousarray(g, np.float32)
        mel = feats.shape[1]
        assert g.size == mel + 1
        out = np.zeros_like(feats)
        self.L.orc_cmvn(g, feats, feats.shape[0], mel, out)
        return out

    def sgemm(self, a, b):
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        c = np.zeros((a.shape[0], b.shape[1]), np.float32)
        self.L.orc_sgemm(a, b, a.shape[0], b.shape[1], a.shape[1], c)
        return c

    def quantize(self, x):
        x = np.ascontiguousarray(x, np.float32)
        q = np.zeros(x.shape, np.uint8)
        s, z = C.c_float(), C.c_int32()
        self.L.orc_quantize(x, x.size, q, C.byref(s), C.byref(z))
        return q, np.float32(s.value), int(z.value)

    def gemm_u8(self, a, sa, za, b, sb, zb, want_acc=True):
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        m, k = a.shape
        n = b.shape[1]
        c = np.zeros((m, n), np.float32)
        acc = np.zeros((m, n), np.int32) if want_acc else None
        self.L.orc_gemm_u8(a, C.c_float(sa), za, b, C.c_float(sb), zb, m, n, k, c,
                           acc.ctypes.data if want_acc else None)
        return c, acc

    def _take(self, ptr, rows, cols):
        if not ptr:
            raise RuntimeError("oracle nnet evaluation failed")
        out = np.ctypeslib.as_array(ptr, shape=(rows * cols,)).reshape(rows, cols).copy()
        self.L.orc_free(ptr)
        return out

    def nnet_propagate(self, nnet_path, x, mode="float"):
        h = self.L.orc_nnet_open(nnet_path.encode())
        if not h:
            raise RuntimeError("orc_nnet_open failed for %s" % nnet_path)
        try:
            x = np.ascontiguousarray(x, np.float32)
            r, c = C.c_int(), C.c_int()
            p = self.L.orc_nnet_propagate(h, x, x.shape[0], x.shape[1],
                                          1 if mode == "u8" else 0, C.byref(r), C.byref(c),
                                          -1, None)
            return self._take(p, r.value, c.value)
        finally:
            self.L.orc_nnet_close(h)

    def am_forward(self, nnet_path, prior, left, right, feats, mode="float", dump_layer=-1,
                   acc_shape=None):
        """Whole utterance as one batch. Returns loglik (and acc if dump_layer >= 0)."""
        h = self.L.orc_nnet_open(nnet_path.encode())
        if not h:
            raise RuntimeError("orc_nnet_open failed for %s" % nnet_path)
        try:
            feats = np.ascontiguousarray(feats, np.float32)
            prior = np.ascontiguousarray(prior, np.float32)
            r, c = C.c_int(), C.c_int()
            acc = np.zeros(acc_shape, np.int32) if dump_layer >= 0 else None
            p = self.L.orc_am_forward(h, prior, prior.size, left, right, feats, feats.shape[0],
                                      feats.shape[1], 1 if mode == "u8" else 0, C.byref(r),
                                      C.byref(c), dump_layer,
                                      acc.ctypes.data if acc is not None else None)
            y = self._take(p, r.value, c.value)
            return (y, acc) if acc is not None else y
        finally:
            self.L.orc_nnet_close(h)
