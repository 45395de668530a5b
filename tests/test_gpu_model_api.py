"""GPU tests of the reference-facing surface above the C ABI: `Model` (create / step / save / load / vis, the mirror of
src/model/model.lua) and the three entry points an UNMODIFIED optim.sgd_list reaches through the parameter proxies
(aocr_group_norms, aocr_grad_scale, aocr_param_axpy; src/optim/optim_sgd.lua:49-52,90)."""
import os

import numpy as np
import pytest

from oracle import Config, GROUPS, Oracle, init_params, init_bn_stats, make_batch
from parity_util import make_handle, rel_err, TOL

pytestmark = pytest.mark.gpu

OPT = dict(batch_size=8, max_encoder_l=30, max_decoder_l=12, input_feed=True, learning_rate=0.1)


def _batch_list(b):
    return [b["images"], b["targets"], b["targets_eval"], b["num_nonzeros"], [f"img{i}.png" for i in range(len(b["images"]))]]


def test_create_draws_trainable_parameters():
    """Model:create (model.lua:83-112) must yield an initialised model: Torch7 reset() distributions per tensor"""
    from aocr import Model
    from oracle.layout import param_specs, unflatten
    m = Model(log=lambda s: None).create(OPT)
    cfg = Config(batch_size=8, max_encoder_l=30, max_decoder_l=12)
    p = m.get_parameters()
    for g in GROUPS:
        named = unflatten(cfg, g, p[g])
        last_bound = None
        for name, shape in param_specs(cfg)[g]:
            v = named[name].astype(np.float64).ravel()
            if name == "emb":                       # nn.LookupTable: N(0,1)
                assert abs(v.mean()) < 0.15 and abs(v.std() - 1.0) < 0.15, (name, v.mean(), v.std())
            elif name.endswith("gamma"):            # SpatialBatchNormalization: U(0,1)
                assert v.min() >= 0 and v.max() <= 1 and abs(v.mean() - 0.5) < 0.1, name
            elif name.endswith("beta"):
                assert np.all(v == 0), name
            else:                                    # Linear / conv weight: U(+-1/sqrt(fan_in)); the bias shares the bound
                if len(shape) >= 2:
                    last_bound = 1.0 / np.sqrt(np.prod(shape[1:]))
                assert np.abs(v).max() <= last_bound * (1 + 1e-6), (name, np.abs(v).max(), last_bound)
                if v.size >= 512:
                    assert abs(v.std() - last_bound / np.sqrt(3)) < 0.1 * last_bound, (name, v.std(), last_bound)
    # another seed, other weights; the same seed, the same weights
    m2 = Model(log=lambda s: None).create(dict(OPT, seed=7))
    m3 = Model(log=lambda s: None).create(OPT)
    assert not np.array_equal(m2.get_parameters()["proj"], p["proj"])
    assert np.array_equal(m3.get_parameters()["decoder"], p["decoder"])
    # and it trains: the loss of a repeated batch goes down
    b = make_batch(8, 100, 7, seed=5)
    losses = [m.step(_batch_list(b), False)[0] for _ in range(6)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0], losses
    for mm in (m, m2, m3):
        mm.shutdown()


def test_step_refused_without_parameters():
    from aocr.capi import AocrConfig, Handle, AocrError
    c = AocrConfig(batch_size=2, max_encoder_l=30, max_decoder_l=8, encoder_num_hidden=512, encoder_num_layers=1,
                   decoder_num_layers=2, target_vocab_size=39, target_embedding_size=20, input_feed=1)
    h = Handle(c, 0)
    b = make_batch(2, 100, 4, seed=1)
    with pytest.raises(AocrError, match="no parameters"):
        h.forward_backward(b["images"], b["targets"], b["targets_eval"])
    h.close()


def test_model_step_matches_oracle_and_save_load_round_trip(tmp_path):
    from aocr import Model
    cfg = Config(batch_size=8, max_encoder_l=30, max_decoder_l=12)
    params, bn = init_params(cfg, 910820), init_bn_stats(cfg)
    m = Model(log=lambda s: None).create(OPT)
    m.set_parameters(params, bn)
    orc = Oracle(cfg, params, bn)
    for step in range(2):
        b = make_batch(8, 100, 7, seed=200 + step)
        lo, _, _ = orc.train_step(b["images"], b["targets"], b["targets_eval"], 0.1)
        lg, stats = m.step(_batch_list(b), False)
        assert abs(lg - lo) / abs(lo) < TOL and stats[0] == b["num_nonzeros"]
        m.global_step += 1
    m.optim_state["learningRate"] = 0.05
    path = str(tmp_path / "model.t7")
    m.save(path)
    m2 = Model(log=lambda s: None)
    m2.load(path, {"batch_size": 4})                     # model.lua:71-74: the caller may override these four
    assert m2.config["batch_size"] == 4 and m2.config["max_decoder_l"] == 12 and m2.config["input_feed"] is True
    assert m2.global_step == 2 and m2.optim_state["learningRate"] == 0.05
    pa, pb = m.get_parameters(), m2.get_parameters()
    for g in GROUPS:
        assert np.array_equal(pa[g], pb[g]), g
    for i in range(3):
        for x, y in zip(m.handle.get_bn_stats(i), m2.handle.get_bn_stats(i)):
            assert np.array_equal(x, y)
    # the reloaded model decodes like the original and like the oracle (BN running statistics travelled)
    b = make_batch(4, 100, 7, seed=300)
    o = orc.decode_greedy(b["images"], b["targets"], b["targets_eval"])
    vis_dir = str(tmp_path / "vis")
    m2.vis(vis_dir)
    l2, st2 = m2.step(_batch_list(b), True)
    l1, st1 = m.step(_batch_list(b), True)
    assert l1 == l2 and st1 == st2
    assert abs(l2 - o["loss_sum"]) / abs(o["loss_sum"]) < TOL and st2[1] == o["num_correct"]
    m2.shutdown()
    m.shutdown()
    rows = open(os.path.join(vis_dir, "results.txt")).read().strip().split("\n")       # model.lua:628-633
    assert len(rows) == 4 and all(len(r.split("\t")) == 5 for r in rows)
    assert rows[0].split("\t")[0] == "img0.png" and rows[0].split("\t")[1] == b["labels"][0]
    with pytest.raises(AssertionError, match="does not exist"):
        Model(log=lambda s: None).load(str(tmp_path / "missing.t7"))


def test_unmodified_sgd_list_over_the_proxies_follows_oracle():
    """use_lua_optim: feval + optim.sgd_list exactly as model.lua:698-701 drives them - norm() per group, mul() when the
    norm exceeds 5, add(-lr, g) - through aocr_group_norms / aocr_grad_scale / aocr_param_axpy.  Compared with the
    oracle's own sgd_update (optim_sgd.lua:49-52,90) AND with the fused aocr_train_step path."""
    from aocr import Model
    cfg = Config(batch_size=8, max_encoder_l=30, max_decoder_l=12)
    params, bn = init_params(cfg, 910820), init_bn_stats(cfg)
    # scale the generator so at least one group's gradient norm exceeds the clip threshold 5
    params = {g: v.copy() for g, v in params.items()}
    params["proj"] *= 40.0
    ma = Model(log=lambda s: None).create(OPT); ma.set_parameters(params, bn)
    mb = Model(log=lambda s: None).create(OPT); mb.set_parameters(params, bn)
    orc = Oracle(cfg, params, bn)
    b = make_batch(8, 100, 7, seed=400)
    # the norms the proxies report = the oracle's
    _, go, _ = Oracle(cfg, params, bn).forward_backward(b["images"], b["targets"], b["targets_eval"])
    ma.handle.forward_backward(b["images"], b["targets"], b["targets_eval"])
    pn, gn = ma.handle.group_norms()
    clipped = 0
    for i, g in enumerate(GROUPS):
        assert abs(gn[i] - np.linalg.norm(go[g])) < (2e-2 if g == "cnn" else 2e-3) * np.linalg.norm(go[g]), (g, gn[i])
        assert abs(pn[i] - np.linalg.norm(params[g].astype(np.float64))) < 1e-5 * pn[i], g
        assert abs(ma.params[i].norm() - pn[i]) < 1e-12 and abs(ma.grad_params[i].norm() - gn[i]) < 1e-12
        clipped += int(gn[i] > 5)
    assert clipped >= 1, gn
    # mul / add on one group: exactly the arithmetic of the two tensor methods
    g3, p3 = ma.handle.get_grads(3), ma.handle.get_params(3)
    ma.grad_params[3].mul(0.5)
    assert np.array_equal(ma.handle.get_grads(3), g3 * np.float32(0.5))
    ma.params[3].add(-0.25, ma.grad_params[3])
    assert np.allclose(ma.handle.get_params(3), p3 - np.float32(0.25) * (g3 * np.float32(0.5)), rtol=0, atol=1e-7)
    ma.set_parameters(params, bn)
    # one optimiser step, three ways
    lo, _, _ = orc.train_step(b["images"], b["targets"], b["targets_eval"], 0.1)
    la, _ = ma.step(_batch_list(b), False, use_lua_optim=True)
    lb, _ = mb.step(_batch_list(b), False)
    assert abs(la - lo) / abs(lo) < TOL and abs(lb - lo) / abs(lo) < TOL
    assert ma.optim_state[1]["evalCounter"] == 1                       # optim_sgd.lua:93 (per-group state)
    po = orc.flat_params()
    pa, pb = ma.get_parameters(), mb.get_parameters()
    for g in GROUPS:
        d_o = po[g] - params[g].astype(np.float64)
        for name, pp in (("sgd_list", pa), ("fused", pb)):
            d = pp[g].astype(np.float64) - params[g].astype(np.float64)
            e = np.linalg.norm(d - d_o) / np.linalg.norm(d_o)
            assert e < (3e-2 if g == "cnn" else 5e-3), (name, g, e)
        assert rel_err(pa[g], pb[g]) < 1e-6, g            # the two library paths agree to fp32 rounding
    ma.shutdown()
    mb.shutdown()
