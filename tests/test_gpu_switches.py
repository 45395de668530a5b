"""Every scheduling mechanism of the engine has an off switch; each must leave the results unchanged.
The switches are read when a handle is created, so each case builds its own handle under a patched environment and is
compared with the oracle at the same bar as the default configuration."""
import os

import numpy as np
import pytest

from oracle import Config, make_batch
from parity_util import train_parity, decode_parity, check_train

pytestmark = pytest.mark.gpu
TOL = 1e-3

CASES = [
    {"AOCR_PERSIST": "0"},                                          # per-kernel recurrences
    {"AOCR_FUSE": "0"},                                             # executor without the fused GEMM -> cell commands
    {"AOCR_CLUSTER": "1"},                                          # executor launched without thread-block clusters
    {"AOCR_DUAL": "0"},                                             # greedy and gold decode passes one after the other
    {"AOCR_SHORT_GOLD": "0"},                                       # gold rows for all max_decoder_l steps
    {"AOCR_GRAPHS": "0", "AOCR_LANES": "0", "AOCR_PDL": "0"},       # plain serial launches
    {"AOCR_CG2": "0"},                                              # single-CTA GEMM kernel instead of the cta_group::2 pairs
    {"AOCR_PERSIST_GEMM": "0"},                                     # one tile per CTA pair, no tile loop / double-buffered accumulator
    {"AOCR_CG2_BN": "128", "AOCR_BOX_POW2": "1"},                   # 256 x 128 pair tiles, power-of-two pixel boxes
]


@pytest.mark.parametrize("env", CASES, ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_switch_leaves_results_unchanged(env, monkeypatch):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    cfg = Config(batch_size=8, max_encoder_l=30, max_decoder_l=12)
    batch = make_batch(8, 100, 9, seed=41)
    out, _ = train_parity(cfg, batch, gemm_mode=0)
    check_train(out, gemm_mode=0)
    res, _, _ = decode_parity(cfg, batch, gemm_mode=0)
    assert res["token_mismatch"] == 0 and res["gold_logp"] < TOL and res["loss"] < TOL, res


def test_early_group_updates_equal_the_update_at_the_end(monkeypatch):
    """AOCR_EARLY_UPDATE=1: aocr_train_step updates the decoder / projector / encoder groups as soon as their gradients
    are final, on a stream of its own under the CNN backward; the parameters after four steps must be bit-identical to
    updating every group after the last kernel (the default), in eager mode and when the step is replayed as a graph."""
    from oracle import GROUPS, init_params, init_bn_stats
    from parity_util import make_handle
    cfg = Config(batch_size=8, max_encoder_l=30, max_decoder_l=12)
    batch = make_batch(8, 100, 9, seed=43)
    params, bn = init_params(cfg, 910820), init_bn_stats(cfg)
    outs = []
    for early in ("1", "0"):
        monkeypatch.setenv("AOCR_EARLY_UPDATE", early)
        h = make_handle(cfg, params, bn)
        losses = [h.train_step(batch["images"], batch["targets"], batch["targets_eval"], 0.1) for _ in range(4)]
        outs.append((losses, [h.get_params(i) for i in range(5)]))
        h.close()
    assert outs[0][0] == outs[1][0]
    for g, a, b in zip(GROUPS, outs[0][1], outs[1][1]):
        assert np.array_equal(a, b), g
