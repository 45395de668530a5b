"""Shared helpers of the GPU parity tests: run libaocr (through the C ABI) and the float64 oracle on the same
seeded inputs and random-init weights, return per-tensor errors.  The oracle is the checker only."""
import numpy as np

from oracle import Config, GROUPS, Oracle, init_params, init_bn_stats, make_batch, param_specs
from oracle.layout import unflatten


def rel_err(a, b):
    """max |a-b| / max |b|  (tensor-scale relative error, the bar of BASELINE.json's north_star)"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


# ---- bars -----------------------------------------------------------------------------------------------------
TOL = 1e-3          # north_star: logits and loss within 1e-3 relative
GRAD_TOL = 2e-3     # gradients (no bar in north_star): tensor-scale max-abs relative error
# CNN gradients in tensor-core mode (bf16x3 operands, ~2^-17 relative rounding): behind a batch-norm they are
# differences of nearly equal sums, so single elements of a tensor carry amplified rounding.  Measured on B200
# (tools/diag_parity.py, B = 4 .. 256): worst per-tensor max-abs error 4e-3, worst per-tensor L2 error 3e-3.
CNN_GRAD_TOL_TC = 1e-2      # per tensor, max-abs / tensor max
CNN_GRAD_L2_TOL_TC = 5e-3   # per tensor, ||diff||_2 / ||oracle||_2


def train_tol(key, gemm_mode):
    if key in ("loss", "logp"):
        return TOL
    cnn = key.startswith(("grad.cnn.", "gradl2.cnn.")) or key == "gradnorm.cnn"
    if gemm_mode != 2 and cnn:
        return CNN_GRAD_L2_TOL_TC if key.startswith("gradl2.") or key == "gradnorm.cnn" else CNN_GRAD_TOL_TC
    return GRAD_TOL


def check_train(out, gemm_mode=0, only=None):
    """every entry of a train_parity() result against its bar; NaN fails (`not v <= tol`)"""
    bad = {k: v for k, v in out.items() if (only is None or k.startswith(only)) and not (v <= train_tol(k, gemm_mode))}
    assert not bad, bad


def make_handle(cfg: Config, params, bn, gemm_mode=0, global_batch=0, device=0):
    from aocr.capi import AocrConfig, Handle
    c = AocrConfig(batch_size=cfg.batch_size, max_encoder_l=cfg.max_encoder_l, max_decoder_l=cfg.max_decoder_l,
                   encoder_num_hidden=cfg.encoder_num_hidden, encoder_num_layers=1, decoder_num_layers=2,
                   target_vocab_size=cfg.target_vocab_size, target_embedding_size=cfg.target_embedding_size,
                   input_feed=1 if cfg.input_feed else 0, dropout=0.0, learning_rate=cfg.learning_rate,
                   dp_rank=0, dp_world=1, global_batch=global_batch, gemm_mode=gemm_mode)
    h = Handle(c, device)
    for i, g in enumerate(GROUPS):
        h.set_params(i, params[g])
    for i, k in enumerate(("bn3", "bn5", "bn7")):
        h.set_bn_stats(i, *bn[k])
    return h


def train_parity(cfg, batch, seed=910820, gemm_mode=0, verbose=False):
    """returns dict name -> relative error for loss, log-probs, intermediate taps and every named gradient"""
    params, bn = init_params(cfg, seed), init_bn_stats(cfg)
    orc = Oracle(cfg, params, bn)
    loss_o, grads_o, logp_o = orc.forward_backward(batch["images"], batch["targets"], batch["targets_eval"])
    h = make_handle(cfg, params, bn, gemm_mode)
    loss_g = h.forward_backward(batch["images"], batch["targets"], batch["targets_eval"])
    B, T = batch["targets"].shape
    out = {"loss": abs(loss_g - loss_o) / abs(loss_o)}
    logp_g = h.get_logprobs(0, T * B).reshape(T, B, -1)
    out["logp"] = rel_err(logp_g, logp_o)
    for g_i, g in enumerate(GROUPS):
        gg = h.get_grads(g_i)
        ng = unflatten(cfg, g, gg)
        no = unflatten(cfg, g, grads_o[g])
        gmax = float(np.abs(grads_o[g]).max())
        for name, _ in param_specs(cfg)[g]:
            # tensor-scale relative error; tensors whose true gradient is ~0 (conv biases in front of a
            # batch-norm) are measured against 1e-3 of the group's largest gradient instead of against ~0
            den = max(float(np.abs(no[name]).max()), 1e-3 * gmax)
            diff = ng[name].astype(np.float64) - no[name]
            if den == 0.0:
                # the whole group's true gradient is exactly 0 (e.g. one image, one source column: batch-norm over a
                # single sample): the library must produce ~0 too - an absolute bar, never 0/0
                out[f"grad.{g}.{name}"] = float(np.abs(diff).max())
                out[f"gradl2.{g}.{name}"] = float(np.linalg.norm(diff))
                continue
            out[f"grad.{g}.{name}"] = float(np.abs(diff).max() / den)
            # L2-relative error of the tensor: insensitive to cancellation at single elements, so it can carry a
            # tight bar even where the max-abs figure cannot (a dropped filter tap or a wrong BN term moves it by >10 %)
            l2den = max(float(np.linalg.norm(no[name])), 1e-3 * gmax * np.sqrt(no[name].size))
            out[f"gradl2.{g}.{name}"] = float(np.linalg.norm(diff) / l2den)
        out[f"gradnorm.{g}"] = abs(np.linalg.norm(gg.astype(np.float64)) - np.linalg.norm(grads_o[g])) / (
            np.linalg.norm(grads_o[g]) + 1e-300)
    # BN running statistics after one training step
    for i, k in enumerate(("bn3", "bn5", "bn7")):
        m, v = h.get_bn_stats(i)
        out[f"{k}.running_mean"] = rel_err(m, orc.bn[k][0].numpy())
        out[f"{k}.running_var"] = rel_err(v, orc.bn[k][1].numpy())
    h.close()
    if verbose:
        for k, v in out.items():
            print(f"  {k:40s} {v:.3e}")
    return out, (loss_g, loss_o)


def decode_parity(cfg, batch, seed=910820, gemm_mode=0, tie_eps=1e-4):
    params, bn = init_params(cfg, seed), init_bn_stats(cfg)
    orc = Oracle(cfg, params, bn)
    o = orc.decode_greedy(batch["images"], batch["targets"], batch["targets_eval"])
    h = make_handle(cfg, params, bn, gemm_mode)
    g = h.decode_greedy(batch["images"], batch["targets"], batch["targets_eval"])
    B = batch["images"].shape[0]
    L = cfg.max_decoder_l
    lp_greedy = h.get_logprobs(1, L * B).reshape(L, B, -1)
    lp_gold = h.get_logprobs(2, L * B).reshape(L, B, -1)
    h.close()
    # token parity outside argmax ties: a row is compared up to (excluding) its first tie position
    mism, tie_rows, compared = 0, 0, 0
    for b in range(B):
        ties = np.where(o["gaps"][b] < tie_eps)[0]
        upto = int(ties[0]) if len(ties) else L
        tie_rows += int(len(ties) > 0)
        compared += upto
        mism += int((g["labels"][b, :upto] != o["labels"][b, :upto]).sum())
    return {"token_mismatch": mism, "tokens_compared": compared, "tie_rows": tie_rows,
            "gold_logp": rel_err(lp_gold, o["gold_logp"]),
            "greedy_logp_t0": rel_err(lp_greedy[0], o["greedy_logp"][0]),
            "loss": abs(g["loss_sum"] - o["loss_sum"]) / abs(o["loss_sum"]),
            "gold_scores": rel_err(g["gold_scores"], o["gold_scores"]),
            "pred_scores": rel_err(g["pred_scores"], o["pred_scores"]),
            "num_correct": (g["num_correct"], o["num_correct"]),
            "min_gap": float(o["gaps"].min())}, g, o
