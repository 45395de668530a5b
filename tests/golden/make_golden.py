"""Generates tests/golden/*.npz from the float64 oracle (the reference itself cannot run here: DESIGN.md §2).
Re-run only when the oracle changes on purpose:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import Config, GROUPS, Oracle, init_params, init_bn_stats, make_batch  # noqa: E402

CASES = {
    # name: (config kwargs, batch args)
    "train_small_He64": (dict(batch_size=3, max_encoder_l=20, max_decoder_l=8, encoder_num_hidden=64), (3, 48, 5, 7)),
    "train_default_b2": (dict(batch_size=2, max_encoder_l=30, max_decoder_l=10), (2, 100, 6, 8)),
    "train_nofeed_He64": (dict(batch_size=2, max_encoder_l=20, max_decoder_l=8, encoder_num_hidden=64, input_feed=False),
                          (2, 53, 4, 9)),
}


def run_case(name):
    ckw, (B, W, maxlen, seed) = CASES[name]
    cfg = Config(**ckw)
    batch = make_batch(B, W, maxlen, seed=seed)
    params, bn = init_params(cfg, 910820), init_bn_stats(cfg)
    orc = Oracle(cfg, params, bn)
    loss, grads, logp = orc.forward_backward(batch["images"], batch["targets"], batch["targets_eval"])
    dec = Oracle(cfg, params, bn).decode_greedy(batch["images"], batch["targets"], batch["targets_eval"])
    out = {"loss_sum": np.float64(loss), "logp": logp.astype(np.float64),
           "labels": dec["labels"], "pred_scores": dec["pred_scores"], "gold_scores": dec["gold_scores"],
           "decode_loss_sum": np.float64(dec["loss_sum"]), "num_correct": np.int64(dec["num_correct"]),
           "gaps": dec["gaps"]}
    for g in GROUPS:
        out[f"gradnorm_{g}"] = np.float64(np.linalg.norm(grads[g]))
        idx = np.linspace(0, grads[g].size - 1, 257).astype(np.int64)
        out[f"gradidx_{g}"] = idx
        out[f"gradval_{g}"] = grads[g][idx]
    return cfg, batch, out


if __name__ == "__main__":
    for name in CASES:
        _, _, out = run_case(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "loss", float(out["loss_sum"]))
