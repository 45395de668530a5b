"""Oracle self-tests (CPU): hand-written backward vs autograd and finite differences,
quirk semantics, optimiser and greedy-decode rules.  The reference ships no tests or
golden vectors, so these pin the oracle's internal consistency."""
import numpy as np
import pytest
import torch

from oracle import Config, GROUPS, Oracle, init_params, init_bn_stats, make_batch, group_sizes, param_specs
from oracle.layout import source_len, unflatten
from oracle.model import _cut


def small_cfg(**kw):
    d = dict(batch_size=3, max_encoder_l=12, max_decoder_l=7, encoder_num_hidden=8)
    d.update(kw)
    return Config(**d)


def autograd_reference(orc, batch):
    """Same forward, gradients from autograd (Q14 expressed as h1 - h1.detach())."""
    cfg = orc.cfg
    for g in GROUPS:
        for v in orc.P[g].values():
            v.requires_grad_(True)
    img = torch.as_tensor(batch["images"], dtype=torch.float64)
    tgt = torch.as_tensor(batch["targets"]).long().T
    tev = torch.as_tensor(batch["targets_eval"]).long().T
    B = img.shape[0]
    saved = {k: [m.clone(), v.clone()] for k, (m, v) in orc.bn.items()}
    cnn_out, _ = orc.cnn_forward(img, train=True)
    orc.bn = saved
    ctx, _, finals = orc.enc_forward(cnn_out.transpose(0, 1))
    st = orc.dec_init(finals, B)
    if cfg.input_feed:
        hcat = torch.cat([finals["fw"][1], finals["bw"][1]], dim=1)
        st["h1"] = hcat - hcat.detach()
    loss = 0.0
    for t in range(tgt.shape[0]):
        st, _ = orc.dec_step(tgt[t], ctx, st)
        loss = loss + orc.nll(orc.generator(st["a"]), tev[t]) / B
    loss.backward()
    grads = {g: {k: v.grad.detach().clone() for k, v in orc.P[g].items()} for g in GROUPS}
    for g in GROUPS:
        for v in orc.P[g].values():
            v.requires_grad_(False)
            v.grad = None
    return float(loss.detach()) * B, grads


@pytest.mark.parametrize("input_feed", [True, False])
def test_backward_matches_autograd(input_feed):
    cfg = small_cfg(input_feed=input_feed)
    batch = make_batch(3, 40, 5, seed=1)
    orc = Oracle(cfg, init_params(cfg, 7), init_bn_stats(cfg))
    loss_a, ga = autograd_reference(orc, batch)
    loss, grads, logp, G = orc.forward_backward(batch["images"], batch["targets"], batch["targets_eval"],
                                                return_named=True)
    assert abs(loss - loss_a) < 1e-9 * max(1, abs(loss_a))
    for g in GROUPS:
        for k in G[g]:
            a, b = G[g][k], ga[g][k]
            scale = float(b.abs().max()) + 1e-30
            assert float((a - b).abs().max()) <= 1e-9 * scale + 1e-14, (g, k)


def test_finite_difference_on_decoder_and_proj():
    cfg = small_cfg()
    batch = make_batch(2, 32, 4, seed=3)
    params = init_params(cfg, 11)
    orc = Oracle(cfg, params, init_bn_stats(cfg))
    _, grads, _ = orc.forward_backward(batch["images"], batch["targets"], batch["targets_eval"])
    rng = np.random.default_rng(0)
    B = batch["images"].shape[0]
    for g in ("decoder", "proj", "enc_fw"):
        n = params[g].shape[0]
        for idx in rng.integers(0, n, size=4):
            vals = []
            for eps in (1e-5, -1e-5):
                p = {k: v.astype(np.float64).copy() for k, v in params.items()}
                p[g][idx] += eps
                o = Oracle(cfg, p, init_bn_stats(cfg))
                l, _, _ = o.forward_backward(batch["images"], batch["targets"], batch["targets_eval"])
                vals.append(l / B)
            fd = (vals[0] - vals[1]) / 2e-5
            # enc_fw differs from the true derivative by the Q14 gradient injection; skip there unless input_feed off
            if g == "enc_fw":
                continue
            assert abs(fd - grads[g][idx]) <= 1e-5 * max(1.0, abs(fd)), (g, idx, fd, grads[g][idx])


def test_param_counts_match_survey():
    cfg = Config()
    sizes = group_sizes(cfg)
    assert sizes == {"cnn": 5551360, "enc_fw": 2101248, "enc_bw": 2101248, "decoder": 20022028, "proj": 39975}
    assert sum(sizes.values()) == 29815859
    assert source_len(100) == 24 and source_len(400) == 99 and source_len(800) == 199


def test_quirk_q14_h1_zero_under_input_feed():
    cfg = small_cfg(input_feed=True)
    orc = Oracle(cfg, init_params(cfg, 5), init_bn_stats(cfg))
    fin = {"fw": (torch.ones(2, 8, dtype=torch.float64), 2 * torch.ones(2, 8, dtype=torch.float64)),
           "bw": (3 * torch.ones(2, 8, dtype=torch.float64), 4 * torch.ones(2, 8, dtype=torch.float64))}
    st = orc.dec_init(fin, 2)
    assert float(st["h1"].abs().max()) == 0.0
    assert torch.equal(st["c1"][:, :8], fin["fw"][0]) and torch.equal(st["c1"][:, 8:], fin["bw"][0])
    cfg2 = small_cfg(input_feed=False)
    orc2 = Oracle(cfg2, init_params(cfg2, 5), init_bn_stats(cfg2))
    st2 = orc2.dec_init(fin, 2)
    assert torch.equal(st2["h1"][:, :8], fin["fw"][1]) and torch.equal(st2["h1"][:, 8:], fin["bw"][1])


def test_sgd_clip_per_group():
    cfg = small_cfg()
    params = init_params(cfg, 2)
    orc = Oracle(cfg, params, init_bn_stats(cfg))
    grads = {g: np.ones_like(params[g], dtype=np.float64) for g in GROUPS}
    grads["proj"] *= 1e-6
    before = orc.flat_params()
    pn, gn = orc.sgd_update(grads, lr=0.1)
    after = orc.flat_params()
    for g in GROUPS:
        n = np.linalg.norm(grads[g])
        scale = 5.0 / n if n > 5 else 1.0
        np.testing.assert_allclose(after[g], before[g] - 0.1 * scale * grads[g], rtol=0, atol=1e-12)


def test_greedy_rules_sticky_pad_and_counts():
    cfg = small_cfg()
    batch = make_batch(3, 40, 5, seed=9)
    orc = Oracle(cfg, init_params(cfg, 13), init_bn_stats(cfg))
    out = orc.decode_greedy(batch["images"], batch["targets"], batch["targets_eval"])
    lab = out["labels"]
    assert lab.shape == (3, cfg.max_decoder_l)
    for b in range(3):
        seen = False
        for t in range(cfg.max_decoder_l):
            if seen:
                assert lab[b, t] == 1          # after PAD/EOS only PAD (log-prob 0 wins)
            if lab[b, t] in (1, 3):
                seen = True
    assert 0 <= out["num_correct"] <= 3
    assert np.isfinite(out["loss_sum"]) and out["loss_sum"] > 0
    assert _cut([5, 6, 3, 7]) == [5, 6] and _cut([5, 1, 1]) == [5, 1, 1]


def test_bn_running_stats_update():
    cfg = small_cfg()
    batch = make_batch(3, 40, 5, seed=4)
    orc = Oracle(cfg, init_params(cfg, 3), init_bn_stats(cfg))
    orc.forward_backward(batch["images"], batch["targets"], batch["targets_eval"])
    rm, rv = orc.bn["bn3"]
    assert float(rm.abs().max()) > 0 and float((rv - 1).abs().max()) > 0


# ---- beam search / dictionary trie (model.lua:380-387,405-536,573-585; utils.lua:177-218)
def test_beam_one_is_greedy_and_wider_beams_never_score_lower():
    from oracle import Config, Oracle, init_params, init_bn_stats, make_batch
    cfg = Config(batch_size=3, max_encoder_l=30, max_decoder_l=7)
    b = make_batch(3, 100, 5, seed=21)
    o = Oracle(cfg, init_params(cfg), init_bn_stats(cfg))
    g = o.decode_greedy(b["images"], b["targets"], b["targets_eval"])
    r1 = o.decode_beam(b["images"], b["targets"], b["targets_eval"], 1)
    assert np.array_equal(g["labels"], r1["labels"]) and np.allclose(g["pred_scores"], r1["pred_scores"], atol=1e-12)
    assert r1["num_correct"] == g["num_correct"] and abs(r1["loss_sum"] - g["loss_sum"]) < 1e-9
    prev = r1["pred_scores"]
    for k in (2, 4):
        rk = o.decode_beam(b["images"], b["targets"], b["targets_eval"], k)
        assert np.all(rk["pred_scores"] >= prev - 1e-9)        # the best hypothesis of a wider beam is at least as good
        prev = rk["pred_scores"]


def test_trie_decode_emits_dictionary_words_and_c_trie_matches():
    from oracle import Config, Oracle, init_params, init_bn_stats, make_batch, load_dictionary, flatten_trie
    from oracle.synth import numlist2str
    words = ["hello", "help", "h3", "abc", "a", "zz9", "0", "42"]
    cfg = Config(batch_size=3, max_encoder_l=30, max_decoder_l=8)
    b = make_batch(3, 100, 5, seed=22)
    o = Oracle(cfg, init_params(cfg), init_bn_stats(cfg))
    r = o.decode_beam(b["images"], b["targets"], b["targets_eval"], 3, trie=load_dictionary(words))
    for row in r["labels"]:
        ids = []
        for v in row.tolist():
            if v == 3:
                break
            ids.append(v)
        assert numlist2str(ids) in words
    # the library's loadDictionary restatement (host code, no GPU needed) numbers the same trie the same way
    from aocr import Trie
    for adp in (False, True):
        assert np.array_equal(Trie(words=words, allow_digit_prefix=adp).numpy(), flatten_trie(load_dictionary(words, adp)))
