"""Runs a few device-resident training steps of BASELINE config 2 (for ncu launch lists / profiles)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-attention-ocr_b200"), os.path.join(ROOT, "tests")]
from oracle import Config, make_batch, init_params, init_bn_stats
from parity_util import make_handle
B = int(os.environ.get("STEP_B", "64")); W = int(os.environ.get("STEP_W", "100")); T = int(os.environ.get("STEP_T", "20"))
n = int(os.environ.get("STEP_N", "3")); mode = int(os.environ.get("AOCR_GEMM_MODE", "0"))
cfg = Config(batch_size=B, max_encoder_l=max(80, W // 4), max_decoder_l=max(50, T))
h = make_handle(cfg, init_params(cfg), init_bn_stats(cfg), gemm_mode=mode)
b = make_batch(B, W, T - 1, force_T=T, kind="noise")
h.stage_batch(b["images"], b["targets"], b["targets_eval"])
for i in range(n):
    l0 = h.launch_count()
    loss = h.train_step_staged(0.1, sync=True)
    print("step", i, "loss", loss, "launches", h.launch_count() - l0, flush=True)
if os.environ.get("STEP_DECODE"):
    h.decode_greedy_staged(sync=True)
    print("decode done", flush=True)
if os.environ.get("STEP_PROF"):      # AOCR_PROF_DUMP=1 prints every profiled call (class, time, work)
    h.prof_enable(True)
    h.train_step_staged(0.1, sync=True)
    print("profiled step:", [h.prof_read(c) for c in range(3)], flush=True)
    h.prof_enable(False)
import time
h.synchronize()
for i in range(3):
    t0 = time.perf_counter()
    h.train_step_staged(0.1, sync=False)
    t1 = time.perf_counter()
    h.synchronize()
    t2 = time.perf_counter()
    print(f"host enqueue {1e3*(t1-t0):.3f} ms, enqueue+drain {1e3*(t2-t0):.3f} ms", flush=True)
