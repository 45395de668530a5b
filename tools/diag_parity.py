"""GPU diagnostic (run under gpurun): prints per-tensor error of libaocr vs the float64 oracle."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-attention-ocr_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402

from oracle import Config, make_batch  # noqa: E402
from parity_util import train_parity, decode_parity  # noqa: E402

mode = int(os.environ.get("AOCR_GEMM_MODE", "2"))
B = int(os.environ.get("DIAG_B", "4"))
W = int(os.environ.get("DIAG_W", "100"))
cfg = Config(batch_size=B, max_encoder_l=40, max_decoder_l=12, input_feed=not os.environ.get("DIAG_NOFEED"))
batch = make_batch(B, W, 7, seed=int(os.environ.get("DIAG_SEED", "3")))
t = time.time()
out, (lg, lo) = train_parity(cfg, batch, gemm_mode=mode, verbose=True)
print("train loss gpu/oracle", lg, lo, "time", time.time() - t)
res, g, o = decode_parity(cfg, batch, gemm_mode=mode)
print("decode", res)
print("labels gpu", g["labels"][0][:12], "oracle", o["labels"][0][:12])
