"""Warm-L2 micro-benchmark of the tcgen05 GEMM core on the shapes of the hot path (prints us and TFLOP/s)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-attention-ocr_b200")]
from aocr.capi import Lib
dll = Lib.get().dll
fn = dll.aocr_bench_gemm
fn.restype = C.c_int
fn.argtypes = [C.c_int] * 7 + [C.POINTER(C.c_float)]
shapes = [  # (name, M, N, K, mn)
    ("dec g1  4096x64x2048", 4096, 64, 2048, 0), ("dec q   1024x64x1024", 1024, 64, 1024, 0),
    ("dec u   1024x64x2048", 1024, 64, 2048, 0), ("dec dx  2048x64x4096", 2048, 64, 4096, 0),
    ("enc h2h 2048x64x512", 2048, 64, 512, 0), ("enc dh  512x64x2048", 512, 64, 2048, 0),
    ("enc i2h 1536x4096x512", 1536, 4096, 512, 0), ("conv6-like 6400x512x4608", 6400, 512, 4608, 0),
    ("wgrad dec 4096x2048x1280 mn", 4096, 2048, 1280, 1), ("square 4096^3", 4096, 4096, 4096, 0),
]
only = os.environ.get('BENCH_ONLY')
for name, M, N, K, mn in shapes:
    if only and only not in name:
        continue
    for terms in ((3, 1) if not only else (3,)):
        for splits in ((0, 1, 2, 4, 8) if N == 64 else (0,)):
            us = C.c_float()
            rc = fn(M, N, K, terms, mn, splits, 50, C.byref(us))
            fl = 2.0 * M * N * K
            print(f"{name:32s} terms={terms} splits={splits}: {us.value:8.2f} us  {fl/us.value/1e6:8.1f} TFLOP/s (algorithmic)"
                  f"  {fl*terms/us.value/1e6:8.1f} (MMA)", flush=True)
