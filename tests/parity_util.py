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
# CNN gradients in tensor-core mode.  An end-to-end comparison of CNN gradients mixes two things: the arithmetic of the
# backward, and the DISCRETE decisions of the forward (ReLU masks, max-pool winners).  The library's forward activations
# agree with the oracle's to <= 4e-5 (bf16x3, seven layers deep), so a fraction ~6e-6 of the decisions - those where the
# oracle's own pre-activation is a near-tie - fall the other way; a flipped ReLU moves one gradient element by its full
# value, so a fraction f of flips costs ~sqrt(f) of a tensor's L2 norm: 0.4-1.4e-2 end to end (up to 2.8e-2 on i.i.d.
# noise images at batch 8), and 5e-4 .. 1.2e-3 even in the fp32 SIMT mode (forward 4e-6, flips 5e-7).  Measured on B200,
# tools/diag_cnn_grads.py, B = 8 .. 256.  `cnn_decomposed_parity` separates the two: with the LIBRARY's decisions and the
# library's `dsrc`, the oracle's CNN backward reproduces the library's CNN gradients to 2-5e-4 per tensor (L2), and
# that is the bar that pins the convolution data / weight gradients and the batch-norm backward (CNN_BWD_TOL).  The
# end-to-end figures keep a loose bar as a sanity check only.
CNN_GRAD_TOL_TC = 8e-2      # per tensor, max-abs / tensor max (end to end, decisions included)
CNN_GRAD_L2_TOL_TC = 2e-2   # per tensor, ||diff||_2 / ||oracle||_2 (end to end, decisions included)
CNN_ACT_TOL = 2e-4          # forward activations per layer, tensor-scale (measured <= 4e-5)
CNN_FLIP_TOL = 5e-5         # fraction of ReLU / pool decisions that differ (measured <= 7e-6)
CNN_FLIP_MARGIN = 1e-3      # ... and only where the oracle's pre-activation is within this of zero, tensor-scale (measured <= 2e-5)
CNN_BWD_TOL = 1.5e-3        # CNN backward given the library's dsrc and decisions, per tensor L2 (measured <= 5e-4)
DSRC_TOL = 1e-4             # gradient at the CNN output, tensor-scale


def train_tol(key, gemm_mode):
    if key in ("loss", "logp"):
        return TOL
    if key == "dsrc":
        return DSRC_TOL
    if key.startswith("cnn_bwd."):
        return CNN_BWD_TOL
    if key.startswith("act"):
        return CNN_ACT_TOL
    if key == "flips":
        return CNN_FLIP_TOL
    if key == "flip_margin":
        return CNN_FLIP_MARGIN
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


def train_parity(cfg, batch, seed=910820, gemm_mode=0, verbose=False, decompose_cnn=True):
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
    # the gradient handed to the CNN (everything upstream of the CNN backward), held tightly
    S = (batch["images"].shape[3] // 2) // 2 - 1
    out["dsrc"] = rel_err(h.debug_read("dsrc", (S, B, 512)), orc.last_dsrc)
    # BN running statistics after one training step
    for i, k in enumerate(("bn3", "bn5", "bn7")):
        m, v = h.get_bn_stats(i)
        out[f"{k}.running_mean"] = rel_err(m, orc.bn[k][0].numpy())
        out[f"{k}.running_var"] = rel_err(v, orc.bn[k][1].numpy())
    if decompose_cnn:
        out.update(cnn_decomposed_parity(h, orc, batch))
    h.close()
    if verbose:
        for k, v in out.items():
            print(f"  {k:40s} {v:.3e}")
    return out, (loss_g, loss_o)


def cnn_decomposed_parity(h, orc, batch):
    """The CNN backward, separated from the two things a tensor-scale comparison of its gradients mixes in.

    After `h.forward_backward` and `orc.forward_backward` on the same batch:
      1. forward: every layer's activation (library fp32, NHWC) against the oracle's            -> `act{l}` (rel err)
      2. discrete decisions: the ReLU masks and max-pool choices of the two forwards differ only where the oracle's own
         pre-activation is a near-tie (|y| or the window gap within the forward error)           -> `flips`, `flip_margin`
      3. backward: the oracle's CNN backward run on the LIBRARY's gradient at the CNN output (`dsrc`) and the LIBRARY's
         decisions, against the library's CNN gradients: what is left is the arithmetic of the convolution data /
         weight gradients and batch-norm backward alone                                       -> `cnn_bwd.{name}` (L2 rel)
    A flipped ReLU moves one element of a gradient by its full value, so a fraction f of flips costs ~sqrt(f) of a
    tensor's L2 norm: forward agreement to 1e-5 gives ~1e-5 flips and a ~3e-3..1e-2 end-to-end gradient difference,
    which is not an error of the backward.  Oracle: cnn.lua:9-45 forward, model.lua:692 backward."""
    import torch
    from oracle.layout import CNN_LAYERS
    B, _, _, W = batch["images"].shape
    W1, W2 = W // 2, W // 4
    S = W2 - 1
    cache = orc._last_ccache
    hw = [(16, W1), (8, W2), (8, W2), (4, W2), (4, W2), (2, W2), (1, S)]      # output (post-pool) size of layer l+1
    out = {}
    flips, total, margin = 0, 0, 0.0
    for l, (name, cin, cout, k, pad, bn, pool) in enumerate(CNN_LAYERS):
        c = cache[l]
        Ho, Wo = hw[l]
        if l == 6:
            a = h.debug_read("act7", (S, B, 512)).transpose(1, 2, 0).reshape(B, 512, 1, S)
        else:
            a = h.debug_read(f"act{l + 1}", (B, Ho, Wo, cout)).transpose(0, 3, 1, 2)
        # the oracle's activation of this layer = the next layer's input (or the CNN output)
        a_o = cache[l + 1]["xin"].numpy() if l < 6 else orc._last_cnn_out.numpy().transpose(0, 2, 1).reshape(B, 512, 1, S)
        out[f"act{l + 1}"] = rel_err(a, a_o)
        pos = torch.from_numpy(np.ascontiguousarray(a > 0))
        if pool is None:
            y_o = None
            new_pos = pos
            d = (new_pos != c["r_pos"])
            flips += int(d.sum()); total += d.numel()
            if d.any():   # how far from zero the oracle's own pre-activation is where the masks differ
                yo = c["xhat"] * c["gamma"][None, :, None, None] + orc.P["cnn"][f"{c['key']}.beta"][None, :, None, None]
                margin = max(margin, float(yo[d].abs().max() / yo.abs().max()))
        else:
            kh, kw = pool
            pi = h.debug_read(f"pidx{l + 1}", (B, Ho, Wo, cout)).transpose(0, 3, 1, 2).astype(np.int64)
            Hp, Wp = c["pre_pool_shape"][2:]
            ph = np.arange(Ho)[None, None, :, None]
            pw = np.arange(Wo)[None, None, None, :]
            idx = (2 * ph + pi // 2) * Wp + (kw * pw + (pi % 2 if kw == 2 else 0))
            idx = torch.from_numpy(np.ascontiguousarray(idx))
            new_pos = torch.zeros(c["pre_pool_shape"], dtype=torch.bool).reshape(B, cout, -1)
            new_pos.scatter_(2, idx.reshape(B, cout, -1), pos.reshape(B, cout, -1))
            new_pos = new_pos.reshape(c["pre_pool_shape"])
            # a decision differs when the pooled output is positive in one forward and not the other, or both are
            # positive and the window winner differs
            pos_o = cache[l + 1]["xin"] > 0
            d = (pos != pos_o) | (pos & pos_o & (idx != c["pool_idx"]))
            flips += int(d.sum()); total += d.numel()
            c["pool_idx"] = idx
        c["r_pos"] = new_pos
    out["flips"] = flips / total
    out["flip_margin"] = margin
    dsrc_g = h.debug_read("dsrc", (S, B, 512))
    g_o = unflatten(orc.cfg, "cnn", orc.cnn_grads_given(dsrc_g))
    g_g = unflatten(orc.cfg, "cnn", h.get_grads(GROUPS.index("cnn")))
    gmax = max(float(np.abs(v).max()) for v in g_o.values())
    for name, _ in param_specs(orc.cfg)["cnn"]:
        den = max(float(np.linalg.norm(g_o[name])), 1e-3 * gmax * np.sqrt(g_o[name].size))
        diff = float(np.linalg.norm(g_g[name].astype(np.float64) - g_o[name]))
        out[f"cnn_bwd.{name}"] = diff / den if den > 0 else diff      # all-zero true gradient: absolute
    return out


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
    # gold log-probs exist for the batch's own target length T: past it the padded targets carry neither loss nor score
    # and the library does not run the gold rows (zeros there)
    T = batch["targets"].shape[1]
    assert not np.any(lp_gold[T:]) or rel_err(lp_gold[T:], o["gold_logp"][T:]) < TOL
    return {"token_mismatch": mism, "tokens_compared": compared, "tie_rows": tie_rows,
            "gold_logp": rel_err(lp_gold[:T], o["gold_logp"][:T]),
            "greedy_logp_t0": rel_err(lp_greedy[0], o["greedy_logp"][0]),
            "loss": abs(g["loss_sum"] - o["loss_sum"]) / abs(o["loss_sum"]),
            "gold_scores": rel_err(g["gold_scores"], o["gold_scores"]),
            "pred_scores": rel_err(g["pred_scores"], o["pred_scores"]),
            "num_correct": (g["num_correct"], o["num_correct"]),
            "min_gap": float(o["gaps"].min())}, g, o
