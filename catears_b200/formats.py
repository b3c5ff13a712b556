"""On-disk formats of the reference (little-endian, host-native), writer + reader.

Layouts follow the reference's readers, which are the specification:
  VEC0 | int32 bytes(=4*dim+4) | int32 dim | data           src/vector.cc:267-300
  MAT0 | int32(8) | int32 rows | int32 cols | rows x VEC0   src/matrix.cc:160-191
  NN02 | int32 L | int32 R | int32 n | n x (LAY0 | int32 type | payload)
                                                            src/nnet.cc:221-293
and the writer side in tool/convert_am.py:26-39,156-168.  Layer type ids are
src/nnet.h:21-30.  A Linear layer's matrix is stored [in x out]
(tool/convert_am.py:318-323).

These are host utilities (model preparation); the product's loader for the same
formats is C++ (catears_b200/csrc/model_io.cc).
"""
import os
import struct

import numpy as np

LINEAR, RELU, NORMALIZE, SOFTMAX, SPLICE, BATCHNORM, LOGSOFTMAX, NARROW = 0, 1, 2, 3, 6, 7, 8, 9


def _vec_bytes(v, dtype="<f4"):
    v = np.ascontiguousarray(v, dtype)
    return b"VEC0" + struct.pack("<ii", v.size * 4 + 4, v.size) + v.tobytes()


def write_vector(path, v, dtype="<f4"):
    with open(path, "wb") as fd:
        fd.write(_vec_bytes(v, dtype))


def _mat_bytes(m):
    m = np.ascontiguousarray(m, "<f4")
    assert m.ndim == 2
    rows, cols = m.shape
    # One big buffer: header + rows x (12-byte VEC0 header + data).
    rec = np.zeros((rows, 12 + 4 * cols), np.uint8)
    rec[:, 0:4] = np.frombuffer(b"VEC0", np.uint8)
    rec[:, 4:12] = np.frombuffer(struct.pack("<ii", cols * 4 + 4, cols), np.uint8)
    rec[:, 12:] = m.view(np.uint8).reshape(rows, 4 * cols)
    return b"MAT0" + struct.pack("<iii", 8, rows, cols) + rec.tobytes()


def layer_bytes(layer):
    """layer: dict with 'type' and the payload fields of that type."""
    t = layer["type"]
    out = b"LAY0" + struct.pack("<i", t)
    if t == LINEAR:
        out += _mat_bytes(layer["W"]) + _vec_bytes(layer["b"])   # W is [in x out]
    elif t == SPLICE:
        idx = list(layer["indices"])
        out += struct.pack("<i", len(idx)) + struct.pack("<%di" % len(idx), *idx)
    elif t == NARROW:
        out += struct.pack("<ii", layer["left"], layer["right"])
    elif t == BATCHNORM:
        out += _vec_bytes(layer["scale"]) + _vec_bytes(layer["offset"])
    elif t in (RELU, NORMALIZE, SOFTMAX, LOGSOFTMAX):
        pass
    else:
        raise ValueError("unknown layer type %r" % t)
    return out


def write_nnet(path, layers, left_context, right_context):
    with open(path, "wb") as fd:
        fd.write(b"NN02" + struct.pack("<iii", left_context, right_context, len(layers)))
        for layer in layers:
            fd.write(layer_bytes(layer))


def write_am_config(path, nnet, prior, left, right, chunk_size, num_pdfs, tid2pdf,
                    extra=None):
    """key = value file read by Configuration::Read (src/configuration.cc:14-54);
    keys used by AcousticModel::Read (src/am.cc:31-56). Paths relative to the file."""
    with open(path, "w") as fd:
        fd.write("nnet = %s\nprior = %s\n" % (os.path.basename(nnet), os.path.basename(prior)))
        fd.write("left_context = %d\nright_context = %d\nchunk_size = %d\n" %
                 (left, right, chunk_size))
        fd.write("num_pdfs = %d\ntid2pdf = %s\n" % (num_pdfs, os.path.basename(tid2pdf)))
        for k, v in (extra or {}).items():
            fd.write("%s = %s\n" % (k, v))


# ---------------------------------------------------------------------------
# Reader (used by tests to round-trip the writer)
# ---------------------------------------------------------------------------

class _Cursor:
    def __init__(self, data):
        self.d, self.p = data, 0

    def take(self, n):
        b = self.d[self.p:self.p + n]
        if len(b) != n:
            raise ValueError("truncated file")
        self.p += n
        return b

    def i32(self):
        return struct.unpack("<i", self.take(4))[0]

    def tag(self, t):
        b = self.take(4)
        if b != t:
            raise ValueError("expected %r, found %r" % (t, b))


def _read_vec(c, dtype="<f4"):
    c.tag(b"VEC0")
    nbytes, dim = c.i32(), c.i32()
    if dim * 4 + 4 != nbytes:
        raise ValueError("VEC0 section size mismatch")
    return np.frombuffer(c.take(4 * dim), dtype).copy()


def read_vector(path, dtype="<f4"):
    return _read_vec(_Cursor(open(path, "rb").read()), dtype)


def _read_mat(c):
    c.tag(b"MAT0")
    c.i32()
    rows, cols = c.i32(), c.i32()
    m = np.zeros((rows, cols), np.float32)
    for r in range(rows):
        m[r] = _read_vec(c)
    return m


def read_nnet(path):
    c = _Cursor(open(path, "rb").read())
    c.tag(b"NN02")
    left, right, n = c.i32(), c.i32(), c.i32()
    layers = []
    for _ in range(n):
        c.tag(b"LAY0")
        t = c.i32()
        layer = {"type": t}
        if t == LINEAR:
            layer["W"] = _read_mat(c)
            layer["b"] = _read_vec(c)
        elif t == SPLICE:
            k = c.i32()
            layer["indices"] = [c.i32() for _ in range(k)]
        elif t == NARROW:
            layer["left"], layer["right"] = c.i32(), c.i32()
        elif t == BATCHNORM:
            layer["scale"] = _read_vec(c)
            layer["offset"] = _read_vec(c)
        elif t not in (RELU, NORMALIZE, SOFTMAX, LOGSOFTMAX):
            raise ValueError("unexpected layer type %d" % t)
        layers.append(layer)
    return layers, left, right
