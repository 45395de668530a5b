"""GPU parity tests (run on the B200 via `pytest -m gpu`): libaocr through the C ABI vs the float64 oracle on the
same seeded inputs and random-init weights.  Bars (BASELINE.json north_star): log-probs and loss within 1e-3
relative; greedy tokens exact outside argmax ties (oracle top-1/top-2 gap < 1e-4)."""
import numpy as np
import pytest

from oracle import Config, make_batch
from parity_util import train_parity, decode_parity, check_train as _check_train, TOL

pytestmark = pytest.mark.gpu
# bars: parity_util.py (loss / log-probs 1e-3; gradients 2e-3 per tensor; CNN gradients in tensor-core mode 1e-2
# max-abs and 5e-3 L2 per tensor)


@pytest.mark.parametrize("gemm_mode", [2, 0])
def test_train_step_parity_small(gemm_mode):
    cfg = Config(batch_size=4, max_encoder_l=30, max_decoder_l=12)
    batch = make_batch(4, 100, 7, seed=3)
    out, _ = train_parity(cfg, batch, gemm_mode=gemm_mode)
    _check_train(out, gemm_mode)


@pytest.mark.parametrize("gemm_mode", [2, 0])
def test_train_step_parity_ragged_width_and_no_input_feed(gemm_mode):
    cfg = Config(batch_size=3, max_encoder_l=40, max_decoder_l=9, input_feed=False)
    batch = make_batch(3, 133, 5, seed=11)      # odd widths exercise floor-mode pooling
    out, _ = train_parity(cfg, batch, gemm_mode=gemm_mode)
    _check_train(out, gemm_mode)


def test_three_train_steps_follow_oracle():
    """forward+backward+clip+SGD three times: loss and log-probs of every step within the 1e-3 bar, i.e. the
    gradients (incl. the cancellation-prone CNN ones) move the weights the way the reference's would."""
    from oracle import Oracle, init_params, init_bn_stats
    from parity_util import make_handle, rel_err
    cfg = Config(batch_size=8, max_encoder_l=30, max_decoder_l=12)
    params, bn = init_params(cfg, 910820), init_bn_stats(cfg)
    orc = Oracle(cfg, params, bn)
    h = make_handle(cfg, params, bn, gemm_mode=0)
    for step in range(3):
        b = make_batch(8, 100, 7, seed=100 + step)
        lo, _, logp_o = orc.train_step(b["images"], b["targets"], b["targets_eval"], 0.1)
        lg = h.train_step(b["images"], b["targets"], b["targets_eval"], 0.1)
        T, B = b["targets"].shape[1], 8
        logp_g = h.get_logprobs(0, T * B).reshape(T, B, -1)
        assert abs(lg - lo) / abs(lo) < TOL, (step, lg, lo)
        assert rel_err(logp_g, logp_o) < TOL, (step, rel_err(logp_g, logp_o))
    # parameter DELTAS (p_after - p_before), not parameters: an update is ~1e-4 of the parameter scale, so a 10 % error
    # of a gradient would move `rel_err(params)` by 1e-5 only.  L2-relative per group.
    po = orc.flat_params()
    for i, g in enumerate(("cnn", "enc_fw", "enc_bw", "decoder", "proj")):
        d_o = po[g] - params[g].astype(np.float64)
        d_g = h.get_params(i).astype(np.float64) - params[g].astype(np.float64)
        e = np.linalg.norm(d_g - d_o) / np.linalg.norm(d_o)
        # fp32 parameters: the delta itself is only known to ~2^-24 * |p| / |delta| ~ 1e-3
        # CNN: its gradients carry ~1e-2 of cancellation-amplified rounding per step (parity_util.py), and three steps compound
        assert e < (1.5e-1 if g == "cnn" else 5e-3), (g, e)
    h.close()


@pytest.mark.parametrize("gemm_mode", [2, 0])
def test_greedy_decode_parity_config1(gemm_mode):
    """BASELINE config 1: default model, random init, greedy decode of batch 4, 32x100."""
    cfg = Config(batch_size=4, max_encoder_l=80, max_decoder_l=50)
    batch = make_batch(4, 100, 10, seed=910820)
    res, g, o = decode_parity(cfg, batch, gemm_mode=gemm_mode)
    assert res["token_mismatch"] == 0, res
    assert res["tokens_compared"] > 0
    assert res["gold_logp"] < TOL and res["loss"] < TOL and res["gold_scores"] < TOL, res
    assert res["num_correct"][0] == res["num_correct"][1]


def test_contract_errors_match_reference_asserts():
    from aocr.capi import AocrError
    from parity_util import make_handle
    from oracle import init_params, init_bn_stats
    cfg = Config(batch_size=2, max_encoder_l=10, max_decoder_l=5)
    h = make_handle(cfg, init_params(cfg, 1), init_bn_stats(cfg))
    b = make_batch(2, 100, 3, seed=1)
    with pytest.raises(AocrError, match=r"max_encoder_l \(10\) < source_l \(24\)!"):
        h.forward_backward(b["images"], b["targets"], b["targets_eval"])
    b = make_batch(2, 40, 8, seed=1, force_T=9)
    with pytest.raises(AocrError, match=r"max_decoder_l \(5\) < target_l \(9\)!"):
        h.forward_backward(b["images"], b["targets"], b["targets_eval"])
    with pytest.raises(AocrError):
        h.get_grads(0)
    h.close()
