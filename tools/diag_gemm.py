"""GPU diagnostic: error of each GEMM back end on a few shapes (prints, never asserts)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-attention-ocr_b200")]
import numpy as np
from aocr.capi import selftest_gemm
rng = np.random.default_rng(0)
for (M, N, K) in [(128, 128, 64), (128, 64, 64), (128, 128, 128), (256, 256, 256), (200, 96, 200), (512, 130, 1000)]:
    A = rng.standard_normal((M, K)).astype(np.float32); B = rng.standard_normal((K, N)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    for mode in (2, 1, 0):
        for swap in ((0, 1, 2) if mode != 2 else (0,)):
            try:
                C = selftest_gemm(A, B, ta=False, tb=True, mode=mode, swap=swap)
                err = np.abs(C - ref).max() / np.abs(ref).max()
                print(f"M{M} N{N} K{K} mode{mode} swap{int(swap)} err {err:.3e}", flush=True)
            except Exception as e:
                print(f"M{M} N{N} K{K} mode{mode} swap{int(swap)} EXC {e}", flush=True)
