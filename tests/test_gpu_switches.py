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
    {"AOCR_GRAPHS": "0", "AOCR_LANES": "0", "AOCR_PDL": "0"},       # plain serial launches
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
