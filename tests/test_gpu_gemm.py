"""GPU tests of the GEMM back ends in isolation (tcgen05 bf16x3 / bf16, fp32 SIMT) against float64 numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [(128, 128, 64), (256, 128, 128), (128, 64, 256), (200, 96, 200), (130, 17, 70), (512, 4, 2048),
          (1536, 2048, 512), (64, 300, 1152),
          (40000, 512, 128), (38000, 200, 192),    # >= 4 tiles per CTA-pair slot: the persistent tile loop, ragged edges
          (4096, 64, 1152), (80000, 48, 128)]      # N <= 64 on CTA pairs (32 B rows per CTA), persistent at the larger M


def _check(M, N, K, ta, tb, mode, swap, tol):
    from aocr.capi import selftest_gemm
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((K, N)).astype(np.float32)
    C = selftest_gemm(A, B, ta=ta, tb=tb, mode=mode, swap=swap)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    err = np.abs(C - ref).max() / np.abs(ref).max()
    assert err < tol, (M, N, K, ta, tb, mode, swap, err)
    return err


@pytest.mark.parametrize("shape", SHAPES)
def test_simt_gemm(shape):
    for ta in (False, True):
        for tb in (False, True):
            _check(*shape, ta, tb, 2, False, 2e-6)


@pytest.mark.parametrize("shape", SHAPES)
def test_tcgen05_bf16x3_gemm(shape):
    for ta, tb, swap in ((False, True, False), (True, False, False), (False, False, True), (True, True, True)):
        _check(*shape, ta, tb, 0, swap, 3e-5)


@pytest.mark.parametrize("shape", SHAPES[:4])
def test_tcgen05_bf16_gemm(shape):
    _check(*shape, False, True, 1, False, 2e-2)


@pytest.mark.parametrize("shape,splits", [((256, 64, 1024), 4), ((128, 16, 4096), 16), ((200, 96, 640), 3),
                                          ((1024, 64, 4096), 0), ((4096, 64, 2048), 0)])
def test_tcgen05_split_k(shape, splits):
    from aocr.capi import selftest_gemm
    M, N, K = shape
    rng = np.random.default_rng(K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((K, N)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    c1 = selftest_gemm(A, B, tb=True, mode=0, splits=splits)
    c2 = selftest_gemm(A, B, tb=True, mode=0, splits=splits)
    assert np.abs(c1 - ref).max() / np.abs(ref).max() < 3e-5
    assert np.array_equal(c1, c2), "split-K reduction must be deterministic"
