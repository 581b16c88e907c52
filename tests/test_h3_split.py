"""The arithmetic of the fp16-split forward path (csrc/tc_gemm.cu H3 branch, csrc/split16.cu), emulated in numpy:
x = hi + lo with hi = fp16(x), lo = fp16(x - hi); acc = A_hi.B_lo + A_lo.B_hi + A_hi.B_hi in fp32.  Shows on the CPU
that the scheme reaches fp32-level accuracy (where one tf32 pass carries ~3e-4), what the fp16 exponent range costs
for small operands, and why the weights are pre-scaled per output channel.  The kernel itself is checked against fp64
in tests/test_h3_gpu.py."""
import numpy as np

f64 = lambda x: np.asarray(x, np.float64)


def split(x):
    x = np.asarray(x, np.float32)
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


def h3_matmul(A, W):
    Ah, Al = split(A)
    Wh, Wl = split(W)
    return f64(Ah) @ f64(Wl) + f64(Al) @ f64(Wh) + f64(Ah) @ f64(Wh)


def rna_tf32(x):
    return ((np.ascontiguousarray(x, np.float32).view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def test_split_is_exact_to_22_bits_in_the_normal_range():
    rs = np.random.RandomState(0)
    x = (rs.randn(100000) * 3).astype(np.float32)
    x = x[np.abs(x) >= 0.125]                    # lo stays a normal fp16 number
    hi, lo = split(x)
    err = np.abs(f64(hi) + f64(lo) - f64(x))
    assert np.all(err <= np.abs(x) * 2.0 ** -21)
    # below 2^-3 the lo half is subnormal: the ABSOLUTE error is bounded by half its spacing
    y = (rs.rand(100000) * 0.125).astype(np.float32)
    hi, lo = split(y)
    assert np.abs(f64(hi) + f64(lo) - f64(y)).max() <= 2.0 ** -25 * 1.0001


def test_three_products_reach_fp32_accuracy():
    rs = np.random.RandomState(1)
    M, K, N = 128, 2304, 64
    A = np.maximum(rs.randn(M, K) * np.abs(rs.randn(M, K)), 0).astype(np.float32)        # ReLU-like activations
    W = (rs.randn(K, N) * 0.02).astype(np.float32)
    ref = f64(A) @ f64(W)
    rel = lambda x: np.linalg.norm(x - ref) / np.linalg.norm(ref)
    one_pass = rel(f64(rna_tf32(A)) @ f64(rna_tf32(W)))
    unscaled = rel(h3_matmul(A, W))
    sc = 2.0 ** (13 - np.floor(np.log2(np.abs(W).max(axis=0))))                         # per output channel, as split16.cu
    scaled = rel(h3_matmul(A, (W * sc).astype(np.float32)) / sc)
    assert 1e-4 < one_pass < 6e-4
    assert unscaled < 5e-6 and scaled < 5e-7 and scaled < unscaled
    assert np.abs(W * sc).max() < 2.0 ** 14


def test_tiny_weights_need_the_row_scale():
    rs = np.random.RandomState(2)
    A = np.abs(rs.randn(64, 1024)).astype(np.float32)
    W = (rs.randn(1024, 32) * 1e-4).astype(np.float32)          # hi halves already lose bits without scaling
    ref = f64(A) @ f64(W)
    rel = lambda x: np.linalg.norm(x - ref) / np.linalg.norm(ref)
    sc = 2.0 ** (13 - np.floor(np.log2(np.abs(W).max(axis=0))))
    assert rel(h3_matmul(A, W)) > 1e-4
    assert rel(h3_matmul(A, (W * sc).astype(np.float32)) / sc) < 5e-7
