"""GPU diagnostic: per-tensor CNN gradient error against the float64 oracle at a given batch / image kind / GEMM mode.
Prints max-abs-relative and L2-relative error per CNN tensor (tests/parity_util.py metrics)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-attention-ocr_b200"), os.path.join(ROOT, "tests")]
from oracle import Config, make_batch  # noqa: E402
from parity_util import train_parity  # noqa: E402

B = int(os.environ.get("DIAG_B", "64"))
T = int(os.environ.get("DIAG_T", "8"))
for kind in os.environ.get("DIAG_KINDS", "strokes,noise").split(","):
    for mode in [int(m) for m in os.environ.get("DIAG_MODES", "0,2").split(",")]:
        cfg = Config(batch_size=B, max_encoder_l=30, max_decoder_l=max(T, 8))
        batch = make_batch(B, 100, T - 1, seed=910820, force_T=T, kind=kind)
        out, _ = train_parity(cfg, batch, gemm_mode=mode, decompose_cnn=True)
        cnn = {k: v for k, v in out.items() if ".cnn" in k}
        worst_abs = max(v for k, v in cnn.items() if k.startswith("grad."))
        worst_l2 = max(v for k, v in cnn.items() if k.startswith("gradl2."))
        other = max(v for k, v in out.items() if k.startswith("grad.") and ".cnn." not in k)
        print(f"   dsrc rel err {out['dsrc']:.2e}")
        print(f"B={B} kind={kind} mode={mode}: loss {out['loss']:.1e} logp {out['logp']:.1e} | CNN worst max-abs {worst_abs:.2e} "
              f"worst L2 {worst_l2:.2e} gradnorm {out['gradnorm.cnn']:.2e} | non-CNN worst max-abs {other:.2e}", flush=True)
        dec = {k: v for k, v in out.items() if k.startswith(("act", "flip", "cnn_bwd."))}
        print("   forward act rel err: " + " ".join(f"{out[f'act{l}']:.1e}" for l in range(1, 8)) +
              f" | decisions differing {out['flips']:.2e} (margin {out['flip_margin']:.1e}) | backward given the library's "
              f"dsrc and decisions: worst L2 {max(v for k, v in dec.items() if k.startswith('cnn_bwd.')):.2e}")
        if os.environ.get("DIAG_VERBOSE"):
            for k, v in list(cnn.items()) + [(k, v) for k, v in dec.items() if k.startswith("cnn_bwd.")]:
                print(f"    {k:34s} {v:.3e}")
