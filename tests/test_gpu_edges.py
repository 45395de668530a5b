"""Edge shapes of the reference's batch format (src/data/data_gen.lua:96-120, src/model/model.lua:255-290): the narrowest
image the CNN accepts (one source column), a single image, an empty label (targets = [GO], targets_eval = [EOS]), a
target as long as max_decoder_l, and a batch whose every other row is pure padding after the first token."""
import numpy as np
import pytest

from oracle import Config, make_batch
from oracle.synth import str2numlist
from parity_util import train_parity, decode_parity, check_train

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _with_labels(batch, labels):
    lists = [str2numlist(s) for s in labels]
    T = max(len(l) for l in lists) - 1
    B = len(labels)
    tg = np.ones((B, T), np.int32)
    te = np.ones((B, T), np.int32)
    for i, l in enumerate(lists):
        tg[i, :len(l) - 1] = l[:-1]
        te[i, :len(l) - 1] = l[1:]
    out = dict(batch)
    out.update(targets=tg, targets_eval=te, labels=labels, num_nonzeros=int(sum(len(l) - 1 for l in lists)))
    return out


def _check(cfg, batch):
    out, _ = train_parity(cfg, batch, gemm_mode=0)
    check_train(out, gemm_mode=0)
    res, _, _ = decode_parity(cfg, batch, gemm_mode=0)
    assert res["token_mismatch"] == 0 and res["gold_logp"] < TOL and res["loss"] < TOL, res


def test_single_source_column_single_image():
    """W = 8 -> W/4 - 1 = 1 source column: attention over one position, encoder of one step; batch of one"""
    cfg = Config(batch_size=1, max_encoder_l=4, max_decoder_l=6)
    _check(cfg, make_batch(1, 8, 3, seed=51))


def test_empty_label_and_padded_rows():
    """row 0 has the empty label (only GO -> EOS), rows 1 and 2 differ in length: PAD targets carry no loss"""
    cfg = Config(batch_size=3, max_encoder_l=12, max_decoder_l=8)
    batch = _with_labels(make_batch(3, 40, 5, seed=52), ["", "a", "hello7"])
    _check(cfg, batch)


def test_target_as_long_as_max_decoder_l():
    """T == max_decoder_l exactly (model.lua:264 asserts max_decoder_l >= target_l)"""
    cfg = Config(batch_size=2, max_encoder_l=12, max_decoder_l=7)
    batch = _with_labels(make_batch(2, 44, 5, seed=53), ["abcdef", "xy"])     # GO + 6 chars = 7 inputs
    assert batch["targets"].shape[1] == 7
    _check(cfg, batch)


@pytest.mark.parametrize("B,W", [(5, 106), (7, 50), (3, 210), (66, 102)])
def test_ragged_widths_and_batches(B, W):
    """widths whose halves / quarters are odd (floor-mode pooling drops a column: W = 106 -> 53 -> 26, W = 50 -> 25 -> 12,
    W = 102 -> 51 -> 25) and batch sizes that divide no pixel box evenly: the implicit-GEMM convolutions tile their output
    with TMA boxes chosen per shape (partial boxes at the right / bottom / last-image edge, CTA pairs padded to even
    counts, exact contraction covers falling back to power-of-two boxes), DESIGN.md 5.1"""
    S = (W // 2) // 2 - 1
    cfg = Config(batch_size=B, max_encoder_l=S + 2, max_decoder_l=8)
    batch = make_batch(B, W, 5, seed=60 + B)
    out, _ = train_parity(cfg, batch, gemm_mode=0)
    check_train(out, gemm_mode=0)
