"""float64 CPU restatement of the recognition hot path (test infrastructure, see oracle/__init__.py).

Every block cites the reference lines it follows.  Primitives are PyTorch-CPU ops
(ATen is the lineal descendant of TH/THNN: cross-correlation conv, floor-mode pool);
backward is HAND-WRITTEN in the reference's schedule (one cell evaluation per timestep,
src/model/model.lua:294-316,553-568,643-690) and cross-checked against autograd and
finite differences in tests/test_oracle.py.

Reference quirks replicated (SURVEY App. C plus one the survey missed):
  Q2  decode always runs max_decoder_l steps (model.lua:266-274,376)
  Q3  sticky PAD via log-prob[PAD] <- 0 once prev token is PAD/EOS (model.lua:448-449)
  Q5  test step = greedy pass + teacher-forced gold pass (model.lua:589-627)
  Q14 with -input_feed the "zero layers >= 2" loop (model.lua:549-552, 369-372, 599-602) indexes the
      state list without the input-feed offset, so it zeroes list slots 3,4 = (h1, c2): the decoder
      starts from c1(0)=[c_fw(S);c_bw(1)] but **h1(0)=0**.  Backward still hands d h1(0) to the
      encoder finals (model.lua:666-667,680-681) because the zeroing is outside the graph.
  Q4  (t=1 parent index for token 39) is NOT replicated: it is an index error in the reference.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .layout import Config, GROUPS, CNN_LAYERS, param_specs, unflatten, flatten, source_len

DT = torch.float64
BN_EPS = 1e-5      # nn.SpatialBatchNormalization default [T7]
BN_MOM = 0.1


def _t(a, dtype=DT):
    return torch.as_tensor(np.asarray(a), dtype=dtype).clone()


class Oracle:
    def __init__(self, cfg: Config, params, bn_stats, dtype=DT):
        """params: {group: flat float array}; bn_stats: {bnK: (running_mean, running_var)}"""
        self.cfg = cfg
        self.dtype = dtype
        self.P = {}
        for g in GROUPS:
            named = unflatten(cfg, g, np.asarray(params[g]))
            self.P[g] = {k: _t(v, dtype) for k, v in named.items()}
        self.bn = {k: [_t(m, dtype), _t(v, dtype)] for k, (m, v) in bn_stats.items()}

    # ------------------------------------------------------------------ params
    def flat_params(self):
        return {g: flatten(self.cfg, g, {k: v.detach().numpy() for k, v in self.P[g].items()}) for g in GROUPS}

    def _flat_grads(self, G):
        return {g: flatten(self.cfg, g, {k: G[g][k].numpy() for k, _ in param_specs(self.cfg)[g]}) for g in GROUPS}

    # ------------------------------------------------------------------ CNN  (src/model/cnn.lua:9-45)
    def cnn_forward(self, images, train):
        """images (B,1,32,W) raw 0..255 -> (B,S,512); cache for backward."""
        P = self.P["cnn"]
        x = (images - 128.0) * (1.0 / 128)                      # cnn.lua:9-10
        cache = []
        for name, cin, cout, k, pad, bn, pool in CNN_LAYERS:
            xin = x
            z = F.conv2d(xin, P[f"{name}.W"], P[f"{name}.b"], stride=1, padding=pad)
            c = {"name": name, "xin": xin, "pad": pad, "pool": pool, "bn": bn}
            if bn:
                key = f"bn{name[-1]}"
                gamma, beta = P[f"{key}.gamma"], P[f"{key}.beta"]
                if train:
                    n = z.shape[0] * z.shape[2] * z.shape[3]
                    mu = z.mean(dim=(0, 2, 3))
                    var = z.var(dim=(0, 2, 3), unbiased=False)
                    rm, rv = self.bn[key]
                    self.bn[key] = [(1 - BN_MOM) * rm + BN_MOM * mu.detach(),
                                    (1 - BN_MOM) * rv + BN_MOM * var.detach() * n / max(n - 1, 1)]
                else:
                    mu, var = self.bn[key]
                inv = 1.0 / torch.sqrt(var + BN_EPS)
                xhat = (z - mu[None, :, None, None]) * inv[None, :, None, None]
                y = xhat * gamma[None, :, None, None] + beta[None, :, None, None]
                c.update(xhat=xhat, inv=inv, gamma=gamma, key=key, train=train)
            else:
                y = z
            r = torch.relu(y)
            c["r_pos"] = (y > 0)
            if pool is not None:
                kh, kw = pool                                     # floor mode, stride = kernel
                p, idx = F.max_pool2d(r, kernel_size=(kh, kw), stride=(kh, kw), return_indices=True)
                c.update(pool_idx=idx, pre_pool_shape=r.shape)
                x = p
            else:
                x = r
            cache.append(c)
        B, C, H, S = x.shape
        assert H == 1 and C == 512
        out = x.reshape(B, 512, S).transpose(1, 2).contiguous()  # cnn.lua:44-45
        return out, cache

    def cnn_backward(self, dout, cache, G):
        """dout (B,S,512) -> accumulates parameter grads into G['cnn'] (model.lua:692)."""
        P = self.P["cnn"]
        B, S, C = dout.shape
        d = dout.transpose(1, 2).reshape(B, 512, 1, S)
        for c in reversed(cache):
            name = c["name"]
            if c["pool"] is not None:
                kh, kw = c["pool"]
                d = F.max_unpool2d(d, c["pool_idx"], kernel_size=(kh, kw), stride=(kh, kw),
                                   output_size=c["pre_pool_shape"][2:])
            d = d * c["r_pos"]
            if c["bn"]:
                key = c["key"]
                xhat, inv, gamma = c["xhat"], c["inv"], c["gamma"]
                G["cnn"][f"{key}.gamma"] += (d * xhat).sum(dim=(0, 2, 3))
                G["cnn"][f"{key}.beta"] += d.sum(dim=(0, 2, 3))
                if c["train"]:
                    n = d.shape[0] * d.shape[2] * d.shape[3]
                    m1 = d.sum(dim=(0, 2, 3)) / n
                    m2 = (d * xhat).sum(dim=(0, 2, 3)) / n
                    d = (gamma * inv)[None, :, None, None] * (d - m1[None, :, None, None] - xhat * m2[None, :, None, None])
                else:
                    d = (gamma * inv)[None, :, None, None] * d
            xin = c["xin"]
            W = P[f"{name}.W"]
            G["cnn"][f"{name}.W"] += torch.nn.grad.conv2d_weight(xin, W.shape, d, stride=1, padding=c["pad"])
            G["cnn"][f"{name}.b"] += d.sum(dim=(0, 2, 3))
            if name != "conv1":
                d = torch.nn.grad.conv2d_input(xin.shape, W, d, stride=1, padding=c["pad"])
        return None

    # ------------------------------------------------------------------ LSTM cell (src/model/LSTM.lua:79-105)
    @staticmethod
    def _cell(gates, c_prev, H):
        """gate order along 4H: [in | forget | out | candidate] (LSTM.lua:90-98)"""
        i = torch.sigmoid(gates[:, 0 * H:1 * H])
        f = torch.sigmoid(gates[:, 1 * H:2 * H])
        o = torch.sigmoid(gates[:, 2 * H:3 * H])
        g = torch.tanh(gates[:, 3 * H:4 * H])
        c = f * c_prev + i * g
        tc = torch.tanh(c)
        h = o * tc
        return c, h, (i, f, o, g, tc)

    @staticmethod
    def _cell_bwd(dc_next, dh, c_prev, acts):
        """returns d(gates pre-activation) (B,4H), dc_prev"""
        i, f, o, g, tc = acts
        do = dh * tc
        dc = dc_next + dh * o * (1 - tc * tc)
        di = dc * g
        df = dc * c_prev
        dg = dc * i
        dgates = torch.cat([di * i * (1 - i), df * f * (1 - f), do * o * (1 - o), dg * (1 - g * g)], dim=1)
        return dgates, dc * f

    # ------------------------------------------------------------------ encoder (model.lua:293-316)
    def enc_forward(self, src):
        """src (S,B,512) -> context (B,S,2He), per-direction step caches"""
        He = self.cfg.He
        S, B, _ = src.shape
        ctx = torch.zeros(B, S, 2 * He, dtype=self.dtype)
        caches = {}
        finals = {}
        for d, grp, order in (("fw", "enc_fw", range(S)), ("bw", "enc_bw", range(S - 1, -1, -1))):
            P = self.P[grp]
            c = torch.zeros(B, He, dtype=self.dtype)
            h = torch.zeros(B, He, dtype=self.dtype)
            steps = {}
            for t in order:
                gates = src[t] @ P["i2h.W"].T + P["i2h.b"] + h @ P["h2h.W"].T + P["h2h.b"]
                c_new, h_new, acts = self._cell(gates, c, He)
                steps[t] = (c, h, acts)          # inputs (prev c,h) + activations
                c, h = c_new, h_new
                if d == "fw":
                    ctx[:, t, :He] = h           # model.lua:303
                else:
                    ctx[:, t, He:] = h           # model.lua:315
            caches[d] = steps
            finals[d] = (c, h)
        return ctx, caches, finals

    def enc_backward(self, src, caches, D_ctx, dc1_0, dh1_0, G):
        """model.lua:662-690 -> d src (S,B,512)"""
        He = self.cfg.He
        S, B, _ = src.shape
        dsrc = torch.zeros_like(src)
        for d, grp, order, sl in (("fw", "enc_fw", range(S - 1, -1, -1), slice(0, He)),
                                  ("bw", "enc_bw", range(S), slice(He, 2 * He))):
            P = self.P[grp]
            dc = dc1_0[:, sl].clone()            # model.lua:666-667 / 680-681
            dh = dh1_0[:, sl].clone()
            for t in order:
                c_prev, h_prev, acts = caches[d][t]
                dh = dh + D_ctx[:, t, sl]        # model.lua:670 / 684
                dg, dc = self._cell_bwd(dc, dh, c_prev, acts)
                G[grp]["i2h.W"] += dg.T @ src[t]
                G[grp]["i2h.b"] += dg.sum(0)
                G[grp]["h2h.W"] += dg.T @ h_prev
                G[grp]["h2h.b"] += dg.sum(0)
                dsrc[t] += dg @ P["i2h.W"]       # model.lua:675 (copy) / 689 (add)
                dh = dg @ P["h2h.W"]
        return dsrc

    # ------------------------------------------------------------------ decoder step (LSTM.lua:18-162)
    def dec_init(self, finals, B):
        """model.lua:539-552 incl. quirk Q14 (h1(0) zeroed under input feed)."""
        Hd = self.cfg.Hd
        c1 = torch.cat([finals["fw"][0], finals["bw"][0]], dim=1)
        h1 = torch.cat([finals["fw"][1], finals["bw"][1]], dim=1)
        z = torch.zeros(B, Hd, dtype=self.dtype)
        if self.cfg.input_feed:
            h1 = z.clone()                       # Q14
        return {"a": z.clone(), "c1": c1, "h1": h1, "c2": z.clone(), "h2": z.clone()}

    def dec_step(self, y, ctx, st):
        """y (B,) 1-based ids; st = {a,c1,h1,c2,h2} -> new state, cache"""
        P = self.P["decoder"]
        Hd = self.cfg.Hd
        e = P["emb"][y.long() - 1]                                            # LSTM.lua:55-56
        x1 = torch.cat([e, st["a"]], dim=1) if self.cfg.input_feed else e     # LSTM.lua:61-65
        g1 = x1 @ P["l1.i2h.W"].T + P["l1.i2h.b"] + st["h1"] @ P["l1.h2h.W"].T + P["l1.h2h.b"]
        c1, h1, acts1 = self._cell(g1, st["c1"], Hd)
        x2 = h1                                                               # Dropout(0) = identity, LSTM.lua:67-70
        g2 = x2 @ P["l2.i2h.W"].T + P["l2.i2h.b"] + st["h2"] @ P["l2.h2h.W"].T + P["l2.h2h.b"]
        c2, h2, acts2 = self._cell(g2, st["c2"], Hd)
        q = h2 @ P["attn.Wa"].T                                               # LSTM.lua:131
        sc = torch.einsum("bsh,bh->bs", ctx, q)                               # LSTM.lua:135-138
        al = torch.softmax(sc, dim=1)                                         # LSTM.lua:139-141
        cv = torch.einsum("bs,bsh->bh", al, ctx)                              # LSTM.lua:142-150
        cat = torch.cat([cv, h2], dim=1)                                      # LSTM.lua:153 ([context ; h])
        a = torch.tanh(cat @ P["attn.Wc"].T)                                  # LSTM.lua:154-156
        new = {"a": a, "c1": c1, "h1": h1, "c2": c2, "h2": h2}
        cache = {"y": y, "x1": x1, "st": st, "acts1": acts1, "acts2": acts2, "h1": h1, "h2": h2,
                 "q": q, "al": al, "cat": cat, "a": a}
        return new, cache

    def dec_step_bwd(self, cache, ctx, d, G, D_ctx):
        """d = {a,c1,h1,c2,h2} grads wrt this step's outputs -> grads wrt its input state."""
        P = self.P["decoder"]
        Hd, E = self.cfg.Hd, self.cfg.target_embedding_size
        a, cat, al, q = cache["a"], cache["cat"], cache["al"], cache["q"]
        du = d["a"] * (1 - a * a)
        G["decoder"]["attn.Wc"] += du.T @ cat
        dcat = du @ P["attn.Wc"]
        dcv, dh2 = dcat[:, :Hd], d["h2"] + dcat[:, Hd:]
        dal = torch.einsum("bh,bsh->bs", dcv, ctx)
        D_ctx += al[:, :, None] * dcv[:, None, :]
        dsc = al * (dal - (al * dal).sum(1, keepdim=True))
        dq = torch.einsum("bs,bsh->bh", dsc, ctx)
        D_ctx += dsc[:, :, None] * q[:, None, :]
        G["decoder"]["attn.Wa"] += dq.T @ cache["h2"]
        dh2 = dh2 + dq @ P["attn.Wa"]
        st = cache["st"]
        dg2, dc2_prev = self._cell_bwd(d["c2"], dh2, st["c2"], cache["acts2"])
        G["decoder"]["l2.i2h.W"] += dg2.T @ cache["h1"]
        G["decoder"]["l2.i2h.b"] += dg2.sum(0)
        G["decoder"]["l2.h2h.W"] += dg2.T @ st["h2"]
        G["decoder"]["l2.h2h.b"] += dg2.sum(0)
        dh1 = d["h1"] + dg2 @ P["l2.i2h.W"]
        dh2_prev = dg2 @ P["l2.h2h.W"]
        dg1, dc1_prev = self._cell_bwd(d["c1"], dh1, st["c1"], cache["acts1"])
        G["decoder"]["l1.i2h.W"] += dg1.T @ cache["x1"]
        G["decoder"]["l1.i2h.b"] += dg1.sum(0)
        G["decoder"]["l1.h2h.W"] += dg1.T @ st["h1"]
        G["decoder"]["l1.h2h.b"] += dg1.sum(0)
        dx1 = dg1 @ P["l1.i2h.W"]
        dh1_prev = dg1 @ P["l1.h2h.W"]
        G["decoder"]["emb"].index_add_(0, cache["y"].long() - 1, dx1[:, :E])   # LookupTable, no paddingValue
        da_prev = dx1[:, E:] if self.cfg.input_feed else torch.zeros_like(d["a"])
        return {"a": da_prev, "c1": dc1_prev, "h1": dh1_prev, "c2": dc2_prev, "h2": dh2_prev}

    # ------------------------------------------------------------------ generator + criterion
    def generator(self, a):
        """output_projector.lua:5-6"""
        P = self.P["proj"]
        return torch.log_softmax(a @ P["W"].T + P["b"], dim=1)

    @staticmethod
    def nll(logp, y):
        """criterion.lua:4-7: weights[PAD]=0, sizeAverage=false -> sum over rows"""
        w = (y != 1).to(logp.dtype)
        return -(w * logp.gather(1, (y.long() - 1)[:, None])[:, 0]).sum()

    # ------------------------------------------------------------------ feval, train branch (model.lua:284-316,537-569,634-695)
    def zero_grads(self):
        return {g: {k: torch.zeros_like(v) for k, v in self.P[g].items()} for g in GROUPS}

    def forward_backward(self, images, targets, targets_eval, global_batch=None, return_named=False):
        """returns (loss_sum [= loss*batch_size, model.lua:701], {group: flat grad}, logp (T,B,V))"""
        cfg = self.cfg
        img = _t(images, self.dtype)
        tgt = torch.as_tensor(np.asarray(targets)).long().T           # (T,B) model.lua:289
        tev = torch.as_tensor(np.asarray(targets_eval)).long().T
        B = img.shape[0]
        Bn = float(global_batch or B)                                 # Q7: DP uses the global batch
        T = tgt.shape[0]
        assert T <= cfg.max_decoder_l, f"max_decoder_l ({cfg.max_decoder_l}) < target_l ({T})!"
        cnn_out, ccache = self.cnn_forward(img, train=True)
        S = cnn_out.shape[1]
        assert S <= cfg.max_encoder_l, f"max_encoder_l ({cfg.max_encoder_l}) < source_l ({S})!"
        src = cnn_out.transpose(0, 1)                                 # model.lua:288
        ctx, ecache, finals = self.enc_forward(src)
        st = self.dec_init(finals, B)
        dcaches, preds = [], []
        for t in range(T):
            st, c = self.dec_step(tgt[t], ctx, st)
            dcaches.append(c)
            preds.append(st["a"])
        G = self.zero_grads()                                         # model.lua:637-641
        D_ctx = torch.zeros_like(ctx)
        Hd = cfg.Hd
        z = lambda: torch.zeros(B, Hd, dtype=self.dtype)
        d = {"a": z(), "c1": z(), "h1": z(), "c2": z(), "h2": z()}
        loss = 0.0
        logps = [None] * T
        for t in range(T - 1, -1, -1):                                # model.lua:643-661
            logp = self.generator(preds[t])
            logps[t] = logp
            y = tev[t]
            loss = loss + float(self.nll(logp, y)) / Bn               # model.lua:645
            w = (y != 1).to(self.dtype)
            dlogp = torch.zeros_like(logp)
            dlogp[torch.arange(B), y - 1] = -w / Bn                   # model.lua:646-647
            dz = dlogp - torch.exp(logp) * dlogp.sum(1, keepdim=True)  # LogSoftMax backward
            G["proj"]["W"] += dz.T @ preds[t]
            G["proj"]["b"] += dz.sum(0)
            d["a"] = d["a"] + dz @ self.P["proj"]["W"]                # model.lua:648-649
            d = self.dec_step_bwd(dcaches[t], ctx, d, G, D_ctx)
        dsrc = self.enc_backward(src, ecache, D_ctx, d["c1"], d["h1"], G)
        self.last_dsrc = dsrc.numpy().copy()                          # (S,B,512): the gradient handed to the CNN
        self._last_ccache = ccache
        self._last_cnn_out = cnn_out.detach()                         # (B,S,512)
        self.cnn_backward(dsrc.transpose(0, 1), ccache, G)            # model.lua:692
        out = (loss * Bn, self._flat_grads(G), torch.stack(logps).numpy())
        if return_named:
            return out + (G,)
        return out

    def cnn_grads_given(self, dsrc):
        """CNN parameter gradients (flat, group layout) for a GIVEN gradient at the CNN output, dsrc (S,B,512), through the
        cache of the last forward_backward.  Lets a test check the CNN backward in isolation from the rounding of what
        feeds it (model.lua:692)."""
        G = self.zero_grads()
        self.cnn_backward(_t(dsrc, self.dtype).transpose(0, 1), self._last_ccache, G)
        return self._flat_grads(G)["cnn"]

    # ------------------------------------------------------------------ optimiser (optim_sgd.lua:40-95)
    def sgd_update(self, grads, lr, clip=5.0):
        """per group: clip L2 norm to `clip`, p -= lr*g.  Returns (param norms, grad norms) as printed at :49."""
        pn, gn = [], []
        for g in GROUPS:
            gr = _t(grads[g], self.dtype)
            flat = _t(flatten(self.cfg, g, {k: v.numpy() for k, v in self.P[g].items()}), self.dtype)
            pn.append(float(flat.norm()))
            n = float(gr.norm())
            gn.append(n)
            if n > clip:
                gr = gr * (clip / n)
            flat = flat - lr * gr
            named = unflatten(self.cfg, g, flat.numpy())
            self.P[g] = {k: _t(v, self.dtype) for k, v in named.items()}
        return pn, gn

    def train_step(self, images, targets, targets_eval, lr, global_batch=None):
        loss, grads, logp = self.forward_backward(images, targets, targets_eval, global_batch)
        self.sgd_update(grads, lr)
        return loss, grads, logp

    # ------------------------------------------------------------------ greedy decode (model.lua:360-536,570-627)
    def decode_greedy(self, images, targets, targets_eval):
        cfg = self.cfg
        L = cfg.max_decoder_l
        img = _t(images, self.dtype)
        B = img.shape[0]
        tg = np.ones((B, L), np.int64)                                # model.lua:266-274
        te = np.ones((B, L), np.int64)
        T0 = np.asarray(targets).shape[1]
        assert T0 <= L, f"max_decoder_l ({L}) < target_l ({T0})!"
        tg[:, :T0] = np.asarray(targets)
        te[:, :T0] = np.asarray(targets_eval)
        tgt, tev = torch.as_tensor(tg).T, torch.as_tensor(te).T
        with torch.no_grad():
            cnn_out, _ = self.cnn_forward(img, train=False)
            S = cnn_out.shape[1]
            assert S <= cfg.max_encoder_l, f"max_encoder_l ({cfg.max_encoder_l}) < source_l ({S})!"
            ctx, _, finals = self.enc_forward(cnn_out.transpose(0, 1))
            st = self.dec_init(finals, B)
            tok = tgt[0].clone()                                      # GO row, model.lua:388
            score = torch.zeros(B, dtype=self.dtype)
            labels = torch.ones(B, L, dtype=torch.long)
            gaps = torch.zeros(B, L, dtype=self.dtype)
            greedy_logp = []
            for t in range(L):
                st, _ = self.dec_step(tok, ctx, st)
                logp = self.generator(st["a"]).clone()
                if t > 0:
                    stick = (tok == 1) | (tok == 3)                   # model.lua:448-449
                    logp[stick, 0] = 0.0
                greedy_logp.append(logp.numpy().copy())
                top2 = logp.topk(2, dim=1)
                gaps[:, t] = top2.values[:, 0] - top2.values[:, 1]
                new_tok = top2.indices[:, 0] + 1
                score = score + top2.values[:, 0]                     # model.lua:450-452
                tok = new_tok
                labels[:, t] = tok
            num_correct = 0
            for b in range(B):                                        # utils.lua:136-175
                num_correct += int(_cut(labels[b].tolist()) == _cut(te[b].tolist()))
            st = self.dec_init(finals, B)                             # gold pass, model.lua:589-627
            loss = 0.0
            gold = torch.zeros(B, dtype=self.dtype)
            gold_logp = []
            for t in range(L):
                st, _ = self.dec_step(tgt[t], ctx, st)
                logp = self.generator(st["a"])
                gold_logp.append(logp.numpy().copy())
                y = tev[t]
                loss += float(self.nll(logp, y)) / B
                w = (y != 1).to(self.dtype)
                gold = gold + w * logp.gather(1, (y - 1)[:, None])[:, 0]
        return {"labels": labels.numpy().astype(np.int32), "pred_scores": score.numpy(),
                "gold_scores": gold.numpy(), "loss_sum": loss * B, "num_correct": num_correct,
                "gaps": gaps.numpy(), "greedy_logp": np.stack(greedy_logp), "gold_logp": np.stack(gold_logp)}

    # ------------------------------------------------------------------ beam search (model.lua:321-536,570-588)
    def decode_beam(self, images, targets, targets_eval, beam_size, trie=None):
        """forward_only step with beam_size > 1 and / or a dictionary trie (nested dicts, oracle/trie.py).
        Follows the reference step by step: first step on the B un-replicated rows (model.lua:376-445), then beam*B rows
        with sticky PAD (:448-449), top-k over beam*V totals (:451-459) or the sorted walk over trie-valid continuations
        (:460-514), parent re-gather of every state (:516-534), backtrack from the best final beam (:573-585), then the
        teacher-forced gold pass (:589-627).  Row order of the replicated state is the reference's b*beam + k.
        Q4 (t = 1 parent computed from the un-decremented id) only matters for beam 1 (every replica of an image holds
        the same state at t = 1); parent = 1 is used.  `tie_gaps` (B, L): margin between the last continuation taken
        and the best one left out; `final_gap` (B): margin between the best and the second-best final beam."""
        cfg = self.cfg
        L, V = cfg.max_decoder_l, cfg.target_vocab_size
        K = min(int(beam_size), V)                                    # model.lua:229
        img = _t(images, self.dtype)
        B = img.shape[0]
        tg = np.ones((B, L), np.int64)
        te = np.ones((B, L), np.int64)
        T0 = np.asarray(targets).shape[1]
        assert T0 <= L, f"max_decoder_l ({L}) < target_l ({T0})!"
        tg[:, :T0] = np.asarray(targets)
        te[:, :T0] = np.asarray(targets_eval)
        tgt, tev = torch.as_tensor(tg).T, torch.as_tensor(te).T
        NEG = float("-inf")
        with torch.no_grad():
            cnn_out, _ = self.cnn_forward(img, train=False)
            ctx, _, finals = self.enc_forward(cnn_out.transpose(0, 1))
            st = self.dec_init(finals, B)
            beam_scores = torch.zeros(B, K, dtype=self.dtype)
            cur_hist, par_hist = [], []
            tie_gaps = torch.full((B, L), float("inf"), dtype=self.dtype)
            locs = [None] * B                                           # trie node per (b, beam)
            ctxK = ctx[:, None].expand(B, K, *ctx.shape[1:]).reshape(B * K, *ctx.shape[1:])
            beam_input = tgt[0].clone()
            for t in range(L):
                st, _ = self.dec_step(beam_input, ctx if t == 0 else ctxK, st)
                logp = self.generator(st["a"]).clone()
                if t == 0:
                    total = logp                                       # (B, V)
                    nb = 1
                else:
                    stick = (beam_input == 1) | (beam_input == 3)      # model.lua:448-449
                    logp[stick, 0] = 0.0
                    total = (logp.view(B, K, V) + beam_scores[:, :, None]).reshape(B, K * V)
                    nb = K
                raw = torch.zeros(B, K, dtype=torch.long)
                new_scores = torch.zeros(B, K, dtype=self.dtype)
                if trie is None:
                    vals, idx = total.topk(min(K + 1, total.shape[1]), dim=1)      # sorted; Torch's topk order is unspecified
                    raw, new_scores = idx[:, :K].clone(), vals[:, :K].clone()
                    if vals.shape[1] > K:
                        tie_gaps[:, t] = vals[:, K - 1] - vals[:, K]
                else:
                    order = torch.argsort(total, dim=1, descending=True)
                    for b in range(B):
                        picked = []
                        first_rejected = None
                        for j in order[b].tolist():
                            vid, beam = j % V + 1, j // V
                            node = trie if t == 0 else locs[b][beam]
                            ok = (vid in node) if t == 0 else (vid == 1 or vid in node)   # model.lua:417,472
                            if ok and len(picked) < K:
                                picked.append(j)
                            elif ok and first_rejected is None:
                                first_rejected = j
                            if len(picked) == K and first_rejected is not None:
                                break
                        assert picked, "dictionary admits no first character"
                        if first_rejected is not None and len(picked) == K:
                            tie_gaps[b, t] = total[b, picked[-1]] - total[b, first_rejected]
                        while len(picked) < K:                          # model.lua:424-436: pad with the best valid one
                            picked.append(picked[0])
                        raw[b] = torch.as_tensor(picked)
                        new_scores[b] = total[b, raw[b]]
                        nl = []
                        for j in picked:                                 # model.lua:437-442,498-511
                            vid, beam = j % V + 1, j // V
                            node = trie if t == 0 else locs[b][beam]
                            nl.append(node if (t > 0 and vid == 1) else node[vid])
                        locs[b] = nl
                beam_scores = new_scores
                cur = raw % V + 1                                      # model.lua:455-458
                par = raw // V if t > 0 else torch.zeros_like(raw)
                cur_hist.append(cur.clone())
                par_hist.append(par.clone())
                rows = (par + torch.arange(B)[:, None] * nb).reshape(-1)   # model.lua:522-531
                st = {k: v[rows] for k, v in st.items()}
                beam_input = cur.reshape(-1)
            srt = beam_scores.sort(dim=1, descending=True)
            scores, idx = srt.values[:, 0], srt.indices[:, 0]          # model.lua:574-576 (first maximum)
            final_gap = srt.values[:, 0] - srt.values[:, 1] if K > 1 else torch.full((B,), float("inf"), dtype=self.dtype)
            labels = torch.ones(B, L, dtype=torch.long)
            ar = torch.arange(B)
            for t in range(L - 1, -1, -1):                             # model.lua:577-585
                labels[:, t] = cur_hist[t][ar, idx]
                idx = par_hist[t][ar, idx]
            num_correct = 0
            for b in range(B):
                num_correct += int(_cut(labels[b].tolist()) == _cut(te[b].tolist()))
            st = self.dec_init(finals, B)                               # gold pass, model.lua:589-627
            loss = 0.0
            gold = torch.zeros(B, dtype=self.dtype)
            for t in range(L):
                st, _ = self.dec_step(tgt[t], ctx, st)
                logp = self.generator(st["a"])
                y = tev[t]
                loss += float(self.nll(logp, y)) / B
                w = (y != 1).to(self.dtype)
                gold = gold + w * logp.gather(1, (y - 1)[:, None])[:, 0]
        return {"labels": labels.numpy().astype(np.int32), "pred_scores": scores.numpy(), "gold_scores": gold.numpy(),
                "loss_sum": loss * B, "num_correct": num_correct, "tie_gaps": tie_gaps.numpy(), "final_gap": final_gap.numpy()}


def _cut(ids):
    out = []
    for v in ids:
        if v == 3:
            break
        out.append(int(v))
    return out
