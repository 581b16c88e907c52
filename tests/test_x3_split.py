"""The arithmetic of the 3xTF32 forward path (csrc/tc_gemm.cu::tc_gemm_x3_kernel), emulated in numpy: operands split
on chip as hi = top 19 bits (what kind::tf32 reads), lo = x - hi, three products per k-step.  Shows on the CPU that the
scheme reaches fp32-level accuracy where the single-pass product (operands rounded to nearest, as the default path
does) carries ~3e-4, and that nothing is gained by splitting one operand only.  The kernel itself is checked against
fp64 in tests/test_zz_late_additions_gpu.py."""
import numpy as np


def trunc(x):
    return (np.ascontiguousarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def rna(x):
    return ((np.ascontiguousarray(x, np.float32).view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def test_three_pass_split_reaches_fp32_accuracy():
    rs = np.random.RandomState(0)
    M, K, N = 128, 2304, 64
    A = np.maximum(rs.randn(M, K) * np.abs(rs.randn(M, K)), 0).astype(np.float32)        # ReLU-like activations
    W = (rs.randn(K, N) * 0.02).astype(np.float32)
    f64 = lambda x: x.astype(np.float64)
    ref = f64(A) @ f64(W)
    rel = lambda x: np.linalg.norm(x - ref) / np.linalg.norm(ref)
    one_pass = rel(f64(rna(A)) @ f64(rna(W)))
    one_pass_trunc = rel(f64(trunc(A)) @ f64(trunc(W)))
    Ah, Wh = trunc(A), trunc(W)
    Al, Wl = A - Ah, W - Wh
    assert np.array_equal(f64(Ah) + f64(Al), f64(A)) and np.array_equal(f64(Wh) + f64(Wl), f64(W))     # the split is exact
    assert np.all(np.abs(Al) <= np.abs(Ah) * 2.0 ** -10 + 1e-45)
    three = rel(f64(Ah) @ f64(Wh) + f64(trunc(Al)) @ f64(Wh) + f64(Ah) @ f64(trunc(Wl)))
    a_only = rel(f64(Ah) @ f64(rna(W)) + f64(trunc(Al)) @ f64(rna(W)))
    assert 1e-4 < one_pass < 6e-4 and one_pass_trunc > 1.5 * one_pass      # truncation is biased: hence the rna at producers
    assert three < 2e-6
    assert a_only > 0.5 * one_pass                                         # splitting one operand buys < 2x
    fp32 = rel(f64(A @ W))
    assert three < 10 * fp32
