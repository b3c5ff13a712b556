"""GPU parity of the tensor-core GEMM (K3/K4) and quantisation (K5) through the matrix-level
C ABI: bit-exact int32 accumulators / u8 codes, tolerance for the float paths."""
import numpy as np
import pytest

from catears_b200 import api

pytestmark = pytest.mark.gpu


def test_quantize_reference_vectors(golden):
    """Outputs of the unmodified reference's Quantize (src/matrix.cc:366-387), incl. the
    all-negative matrix that shows the FLT_MIN quirk."""
    r = golden["ref"]
    q, s, z = api.quantize(r["q_a"])
    assert (s, z) == (np.float32(r["q_params"][0]), int(r["q_params"][1]))
    assert np.array_equal(q, r["q_a8"])
    q, s, z = api.quantize(r["q_neg"])
    assert (s, z) == (np.float32(r["q_neg_params"][0]), int(r["q_neg_params"][1]))
    assert np.array_equal(q, r["q_neg8"])


def test_quantize_vs_oracle_shapes(port):
    rng = np.random.default_rng(3)
    for shape in ((1, 1), (5, 3), (121, 233), (1024, 1024), (7, 1023)):
        x = (rng.standard_normal(shape) * rng.uniform(0.1, 50)).astype(np.float32)
        if shape == (1, 1):
            x[0, 0] = 3.0
        q, s, z = api.quantize(x)
        wq, ws, wz = port.quantize(x)
        assert (s, z) == (ws, wz), shape
        assert np.array_equal(q, wq), shape


def test_quantizer_arithmetic_selftest():
    """The production quantiser (reciprocal + two fused corrections, truncation of q + 0.49999997,
    saturating pack) against plain IEEE `roundf(min(max(v / scale + zp, 0), 255))` on the device:
    2^27 triples concentrated on the rounding ties, zero disagreements allowed."""
    for seed in (1, 2026):
        assert api.selftest_quantizer(1 << 26, seed) == 0


def test_quantize_wide_rows_vs_oracle(port):
    """Shapes that take the two-rows-in-flight kernel (cols % 4 == 0, <= 1024), incl. a degenerate
    all-zero matrix (denormal scale -> the plain-division branch) and ragged row counts."""
    rng = np.random.default_rng(11)
    for shape in ((1, 4), (17, 40), (33, 128), (250, 200), (1000, 1024), (3, 512)):
        x = (rng.standard_normal(shape) * rng.uniform(0.1, 50)).astype(np.float32)
        x[rng.random(shape) < 0.3] = 0.0
        q, s, z = api.quantize(x)
        wq, ws, wz = port.quantize(x)
        assert (s, z) == (ws, wz), shape
        assert np.array_equal(q, wq), shape
    x = np.zeros((9, 64), np.float32)
    q, s, z = api.quantize(x)
    wq, ws, wz = port.quantize(x)
    assert (s, z) == (ws, wz) and np.array_equal(q, wq)
    x = -np.abs(rng.standard_normal((40, 256))).astype(np.float32)
    q, s, z = api.quantize(x)
    wq, ws, wz = port.quantize(x)
    assert (s, z) == (ws, wz) and np.array_equal(q, wq)


def test_gemm_u8_reference_vectors(golden):
    r = golden["ref"]
    sa, za, sb, zb = r["q_params"]
    c, acc = api.gemm_u8(r["q_a8"], np.float32(sa), int(za), r["q_b8"], np.float32(sb), int(zb))
    assert np.array_equal(acc, r["q_acc"])
    assert np.array_equal(c, r["q_c"])


@pytest.mark.parametrize("m,n,k", [(5, 3, 2), (100, 100, 1), (121, 233, 17), (128, 256, 128),
                                   (129, 257, 129), (1024, 1024, 80), (300, 1024, 3072)])
def test_gemm_u8_bit_exact_vs_oracle(port, m, n, k):
    """Shapes of test/gemm_test.cc:79-92 plus tile-boundary and TDNN-layer shapes."""
    rng = np.random.default_rng(m * 7 + n)
    a = rng.integers(0, 256, (m, k), dtype=np.uint8)
    b = rng.integers(0, 256, (k, n), dtype=np.uint8)
    za, zb = int(rng.integers(0, 256)), int(rng.integers(0, 256))
    sa, sb = np.float32(0.0123), np.float32(0.0045)
    c, acc = api.gemm_u8(a, sa, za, b, sb, zb)
    want = (a.astype(np.int64) - za) @ (b.astype(np.int64) - zb)
    assert np.array_equal(acc, want.astype(np.int32))
    wc = (want.astype(np.int32).astype(np.float32) * np.float32(sa * sb)).astype(np.float32)
    assert np.array_equal(c, wc)
    if m * n * k <= 2 ** 24:
        oc, oacc = port.gemm_u8(a, sa, za, b, sb, zb)
        assert np.array_equal(acc, oacc) and np.array_equal(c, oc)


def test_gemm_u8_saturated_codes():
    """All-255 codes with zero points 0: the largest raw sums (K * 255^2) stay exact."""
    m, n, k = 130, 260, 3072
    a = np.full((m, k), 255, np.uint8)
    b = np.full((k, n), 255, np.uint8)
    c, acc = api.gemm_u8(a, np.float32(1), 0, b, np.float32(1), 0)
    assert (acc == k * 255 * 255).all()
    c, acc = api.gemm_u8(a, np.float32(1), 255, b, np.float32(1), 255)
    assert (acc == 0).all()


@pytest.mark.parametrize("prec,tol", [("fp32", 2e-5), ("bf16x3", 1e-4), ("tf32", 2e-3), ("bf16", 1.5e-2)])
@pytest.mark.parametrize("m,n,k", [(5, 3, 2), (121, 233, 17), (1024, 1024, 80), (257, 1024, 3072)])
def test_gemm_f32_vs_float64(prec, tol, m, n, k):
    """test/gemm_test.cc:96-104 compares MatMat with SimpleMatMat to 0.01 abs on U[-.5,.5] x
    U[1,2]; here against a float64 product, relative to sqrt(k) * |a|max * |b|max."""
    rng = np.random.default_rng(k)
    a = rng.uniform(-0.5, 0.5, (m, k)).astype(np.float32)
    b = rng.uniform(1.0, 2.0, (k, n)).astype(np.float32)
    c = api.gemm_f32(a, b, prec)
    want = a.astype(np.float64) @ b.astype(np.float64)
    err = np.abs(c - want).max()
    assert err < tol * np.sqrt(k) * 0.5 * 2.0, (prec, err)
    if prec == "fp32":
        assert err < 0.01                                  # the reference's own bar
