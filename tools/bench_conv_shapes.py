"""Decomposes the tcgen05 GEMM time on the convolution shapes of config 2 (plain GEMM with the same M, N, K):
full kernel, without TMA/MMA (AOCR_TC_DBG=1), without epilogue stores (AOCR_TC_DBG=2)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-attention-ocr_b200")]
from aocr.capi import Lib
dll = Lib.get().dll
fn = dll.aocr_bench_gemm
fn.restype = C.c_int
fn.argtypes = [C.c_int] * 7 + [C.POINTER(C.c_float)]
shapes = [("conv2 fwd", 51200, 128, 576), ("conv3 fwd", 12800, 256, 1152), ("conv4 fwd", 12800, 256, 2304),
          ("conv5 fwd", 6400, 512, 2304), ("conv6 fwd", 6400, 512, 4608), ("conv2 dgrad", 51200, 64, 1152),
          ("conv6 dgrad", 6400, 512, 4608)]
for dbg in (0, 1, 2):
    os.environ["AOCR_TC_DBG"] = str(dbg)
    for name, M, N, K in shapes:
        us = C.c_float()
        fn(M, N, K, 3, 0, 0, 30, C.byref(us))
        fl = 2.0 * M * N * K
        tiles = ((M + 127) // 128) * ((N + 127) // 128)
        print(f"dbg={dbg} {name:12s} {M}x{N}x{K}: {us.value:7.1f} us  {fl/us.value/1e6:6.1f} T/s  tiles={tiles} waves={tiles/148:.2f}", flush=True)
