"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the float64 oracle).
CPU: the oracle still reproduces them (pins the oracle against accidental change).
GPU: libaocr reproduces them through the C ABI without the oracle in the loop."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import CASES, run_case  # noqa: E402
from oracle import GROUPS, init_params, init_bn_stats  # noqa: E402


def _load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    gold = _load(name)
    _, _, out = run_case(name)
    assert abs(float(out["loss_sum"]) - float(gold["loss_sum"])) < 1e-9 * abs(float(gold["loss_sum"]))
    np.testing.assert_allclose(out["logp"], gold["logp"], rtol=0, atol=1e-10)
    assert np.array_equal(out["labels"], gold["labels"])
    for g in GROUPS:
        np.testing.assert_allclose(out[f"gradval_{g}"], gold[f"gradval_{g}"], rtol=1e-8, atol=1e-14)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_library_reproduces_golden(name):
    from oracle import Config, make_batch
    from parity_util import make_handle, rel_err
    gold = _load(name)
    ckw, (B, W, maxlen, seed) = CASES[name]
    cfg = Config(**ckw)
    batch = make_batch(B, W, maxlen, seed=seed)
    h = make_handle(cfg, init_params(cfg, 910820), init_bn_stats(cfg))
    loss = h.forward_backward(batch["images"], batch["targets"], batch["targets_eval"])
    T = batch["targets"].shape[1]
    logp = h.get_logprobs(0, T * B).reshape(T, B, -1)
    assert abs(loss - float(gold["loss_sum"])) < 1e-3 * abs(float(gold["loss_sum"]))
    assert rel_err(logp, gold["logp"]) < 1e-3
    for i, g in enumerate(GROUPS):
        gr = h.get_grads(i)
        tol = 2e-2 if g == "cnn" else 2e-3
        assert abs(np.linalg.norm(gr.astype(np.float64)) - float(gold[f"gradnorm_{g}"])) < tol * float(gold[f"gradnorm_{g}"])
    dec = h.decode_greedy(batch["images"], batch["targets"], batch["targets_eval"])
    ties = np.cumsum(gold["gaps"] < 1e-4, axis=1) > 0
    assert np.all((dec["labels"] == gold["labels"]) | ties)
    assert rel_err(dec["gold_scores"], gold["gold_scores"]) < 1e-3
    assert dec["num_correct"] == int(gold["num_correct"])
    h.close()
