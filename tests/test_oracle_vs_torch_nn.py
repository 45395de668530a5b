"""Independent cross-check of the oracle's building blocks against PyTorch's own modules (`torch.nn.Conv2d`,
`BatchNorm2d`, `MaxPool2d`, `LSTM`, `log_softmax` / `nll_loss`): ATen is the lineal descendant of the TH / THNN code
behind the Torch7 modules the reference composes, so agreement of the hand-written restatement with these
stock modules — forward values, running statistics, and gradients through autograd — is evidence that does not come
from the oracle itself.  It does NOT pin parity in the contract's sense (no vector of the reference exists), and the
decoder's attention graph has no stock module to compare with."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import Config, Oracle, init_params, init_bn_stats, make_batch
from oracle.layout import CNN_LAYERS, unflatten

DT = torch.float64


def _stock_cnn(named, bn, train):
    """cnn.lua:9-45 out of stock modules"""
    layers = []
    bns = {}
    for name, cin, cout, k, pad, isbn, pool in CNN_LAYERS:
        conv = nn.Conv2d(cin, cout, k, stride=1, padding=pad).to(DT)
        conv.weight.data.copy_(torch.from_numpy(named[name + ".W"]))
        conv.bias.data.copy_(torch.from_numpy(named[name + ".b"]))
        layers.append(conv)
        if isbn:
            key = "bn" + name[-1]
            b = nn.BatchNorm2d(cout, eps=1e-5, momentum=0.1).to(DT)          # nn.SpatialBatchNormalization defaults
            b.weight.data.copy_(torch.from_numpy(named[key + ".gamma"]))
            b.bias.data.copy_(torch.from_numpy(named[key + ".beta"]))
            b.running_mean.copy_(torch.as_tensor(np.asarray(bn[key][0]), dtype=DT))
            b.running_var.copy_(torch.as_tensor(np.asarray(bn[key][1]), dtype=DT))
            layers.append(b)
            bns[key] = b
        layers.append(nn.ReLU())
        if pool is not None:
            layers.append(nn.MaxPool2d(kernel_size=pool, stride=pool))      # floor mode
    net = nn.Sequential(*layers)
    net.train(train)
    return net, bns


def test_cnn_forward_backward_and_running_statistics_match_stock_modules():
    cfg = Config(batch_size=3, max_encoder_l=30, max_decoder_l=8)
    params, bn = init_params(cfg, 5), init_bn_stats(cfg)
    named = {k: np.asarray(v, np.float64) for k, v in unflatten(cfg, "cnn", params["cnn"]).items()}
    images = make_batch(3, 100, 5, seed=9)["images"]
    x = (torch.from_numpy(images).to(DT) - 128.0) / 128.0
    for train in (True, False):
        orc = Oracle(cfg, params, bn)
        out_o, cache = orc.cnn_forward(torch.from_numpy(images).to(DT), train=train)
        net, bns = _stock_cnn(named, bn, train)
        y = net(x)                                                   # (B,512,1,S)
        out_s = y.reshape(3, 512, -1).transpose(1, 2)
        assert torch.allclose(out_o, out_s, rtol=0, atol=1e-10), float((out_o - out_s).abs().max())
        if train:
            for key, b in bns.items():                               # momentum 0.1, unbiased running variance
                assert torch.allclose(orc.bn[key][0], b.running_mean, atol=1e-12)
                assert torch.allclose(orc.bn[key][1], b.running_var, atol=1e-12)
            # hand-written backward against autograd through the stock modules
            g = torch.from_numpy(np.random.default_rng(1).standard_normal(tuple(out_s.shape)))
            G = orc.zero_grads()
            orc.cnn_backward(g, cache, G)
            out_s.backward(g)
            it = iter(net)
            for name, cin, cout, k, pad, isbn, pool in CNN_LAYERS:
                conv = next(m for m in it if isinstance(m, nn.Conv2d))
                assert torch.allclose(G["cnn"][name + ".W"], conv.weight.grad, atol=1e-9), name
                assert torch.allclose(G["cnn"][name + ".b"], conv.bias.grad, atol=1e-9), name
            for key, b in bns.items():
                assert torch.allclose(G["cnn"][key + ".gamma"], b.weight.grad, atol=1e-9)
                assert torch.allclose(G["cnn"][key + ".beta"], b.bias.grad, atol=1e-9)


def _to_torch_gate_order(w, H):
    """reference rows [in | forget | out | candidate] (LSTM.lua:90-98) -> torch.nn.LSTM rows [in | forget | cell | out]"""
    i, f, o, g = w[0:H], w[H:2 * H], w[2 * H:3 * H], w[3 * H:4 * H]
    return np.concatenate([i, f, g, o], axis=0)


def test_encoder_matches_a_stock_bidirectional_lstm():
    cfg = Config(batch_size=4, max_encoder_l=30, max_decoder_l=8)
    params, bn = init_params(cfg, 6), init_bn_stats(cfg)
    He = cfg.He
    lstm = nn.LSTM(512, He, num_layers=1, bidirectional=True).to(DT)
    for grp, sfx in (("enc_fw", ""), ("enc_bw", "_reverse")):
        p = {k: np.asarray(v, np.float64) for k, v in unflatten(cfg, grp, params[grp]).items()}
        getattr(lstm, "weight_ih_l0" + sfx).data.copy_(torch.from_numpy(_to_torch_gate_order(p["i2h.W"], He)))
        getattr(lstm, "weight_hh_l0" + sfx).data.copy_(torch.from_numpy(_to_torch_gate_order(p["h2h.W"], He)))
        getattr(lstm, "bias_ih_l0" + sfx).data.copy_(torch.from_numpy(_to_torch_gate_order(p["i2h.b"], He)))
        getattr(lstm, "bias_hh_l0" + sfx).data.copy_(torch.from_numpy(_to_torch_gate_order(p["h2h.b"], He)))
    S, B = 7, 4
    src = torch.from_numpy(np.random.default_rng(2).standard_normal((S, B, 512)))
    orc = Oracle(cfg, params, bn)
    ctx, _, finals = orc.enc_forward(src)
    out, (hn, cn) = lstm(src)                                        # out (S,B,2He) = [fw | bw] per position
    assert torch.allclose(ctx, out.transpose(0, 1), atol=1e-12)      # context[:, t] = [h_fw(t) ; h_bw(t)] (model.lua:303,315)
    assert torch.allclose(finals["fw"][1], hn[0], atol=1e-12) and torch.allclose(finals["fw"][0], cn[0], atol=1e-12)
    assert torch.allclose(finals["bw"][1], hn[1], atol=1e-12) and torch.allclose(finals["bw"][0], cn[1], atol=1e-12)


def test_generator_and_criterion_match_stock_functions():
    cfg = Config(batch_size=5, max_encoder_l=30, max_decoder_l=8)
    params, bn = init_params(cfg, 7), init_bn_stats(cfg)
    orc = Oracle(cfg, params, bn)
    a = torch.from_numpy(np.random.default_rng(3).standard_normal((5, cfg.Hd)))
    y = torch.tensor([1, 2, 3, 17, 39], dtype=torch.int32)           # 1 = PAD carries no loss (criterion.lua:5)
    logp = orc.generator(a)
    P = orc.P["proj"]
    ref = F.log_softmax(F.linear(a, P["W"], P["b"]), dim=1)
    assert torch.allclose(logp, ref, atol=1e-12)
    w = torch.ones(cfg.target_vocab_size, dtype=DT)
    w[0] = 0
    assert torch.allclose(orc.nll(logp, y), F.nll_loss(ref, y.long() - 1, weight=w, reduction="sum"), atol=1e-12)
