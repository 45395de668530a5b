"""GPU parity tests (run on the B200 via `pytest -m gpu`): libaocr through the C ABI vs the float64 oracle on the
same seeded inputs and random-init weights.  Bars (BASELINE.json north_star): log-probs and loss within 1e-3
relative; greedy tokens exact outside argmax ties (oracle top-1/top-2 gap < 1e-4)."""
import numpy as np
import pytest

from oracle import Config, make_batch
from parity_util import train_parity, decode_parity

pytestmark = pytest.mark.gpu

TOL = 1e-3          # north_star: logits and loss within 1e-3 relative
GRAD_TOL = 2e-3     # gradients: same bar class, tensor-scale relative


def _check_train(out):
    bad = {k: v for k, v in out.items() if v > (TOL if k in ("loss", "logp") else GRAD_TOL)}
    assert not bad, bad


@pytest.mark.parametrize("gemm_mode", [2, 0])
def test_train_step_parity_small(gemm_mode):
    cfg = Config(batch_size=4, max_encoder_l=30, max_decoder_l=12)
    batch = make_batch(4, 100, 7, seed=3)
    out, _ = train_parity(cfg, batch, gemm_mode=gemm_mode)
    _check_train(out)


def test_train_step_parity_ragged_width_and_no_input_feed():
    cfg = Config(batch_size=3, max_encoder_l=40, max_decoder_l=9, input_feed=False)
    batch = make_batch(3, 133, 5, seed=11)      # odd widths exercise floor-mode pooling
    out, _ = train_parity(cfg, batch)
    _check_train(out)


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
