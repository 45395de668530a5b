"""Checkpoints in Torch7's container (src/model/model.lua:45-80 `load`, :720-725 `save`).

Two payloads are understood, both `torch.load`-able files:

* the reference's own checkpoint `{ {cnn_model, encoder_fw, encoder_bw, decoder, output_projector}, config,
  global_step, optim_state }` (model.lua:724): the five serialised module trees.  `load_checkpoint` walks them —
  `nn.Sequential.modules` in construction order for the CNN and the output projector, the `nngraph` node lists of the
  LSTM graphs — and files every weight under the tensor names of `aocr/layout.py`.  The LSTM Linear modules carry no
  names (memory.lua:55-61 names them only under `-prealloc`), so they are told apart by graph structure: `h2h` of layer
  L is the Linear fed directly by the graph input `prev_h[L]` (LSTM.lua:47), `i2h` its partner in the same `CAddTable`
  (LSTM.lua:82-84); `W_a` and `W_c` are the (H,H) and (H,2H) `LinearNoBias` of the nested attention graph
  (LSTM.lua:130,155).
* a named-tensor table `{format="aocr-params-v1", params={group={name=tensor}}, bn={...}, config, global_step,
  optim_state}`: what `save_checkpoint` writes.  `lua/t7_convert.lua` turns it into the reference's module trees (and
  back) under Torch7 by copying tensor by tensor into a model built by `model:create` — rebuilding nngraph's node
  objects outside Torch7 would be guesswork, copying named tensors is not.
"""
import sys

import numpy as np

from . import t7
from .layout import CNN_LAYERS, GROUPS, flatten, param_specs, unflatten

FORMAT = "aocr-params-v1"
BN_KEYS = ("bn3", "bn5", "bn7")


# ---------------------------------------------------------------------------------------------- named-tensor table
def save_checkpoint(path, config, params, bn_stats, global_step=0, optim_state=None, dtype=np.float64):
    """params: {group: flat float array in the external layout}; bn_stats: [(running_mean, running_var)] x 3.
    Tensors are written as DoubleTensors by default (the reference keeps its CPU model in double, model.lua:54-58)."""
    cfg = {k: v for k, v in dict(config).items() if isinstance(v, (int, float, str, bool))}
    named = {g: {n: np.asarray(a, dtype=dtype) for n, a in unflatten(config, g, np.asarray(params[g])).items()} for g in GROUPS}
    bn = {k: {"running_mean": np.asarray(m, dtype=dtype), "running_var": np.asarray(v, dtype=dtype)}
          for k, (m, v) in zip(BN_KEYS, bn_stats)}
    t7.save(path, {"format": FORMAT, "params": named, "bn": bn, "config": cfg, "global_step": int(global_step),
                   "optim_state": dict(optim_state or {})})
    return path


# ---------------------------------------------------------------------------------------------- module trees
def _cls(o):
    return o.cls if isinstance(o, t7.T7Object) else ""


def _modules(seq):
    m = seq.get("modules") or []
    return m if isinstance(m, list) else [m[k] for k in sorted(k for k in m if isinstance(k, int))]


def _conv_weight(m, cin, cout, k):
    w = np.asarray(m["weight"], dtype=np.float64)
    return w.reshape(cout, cin, k, k)          # cudnn / nn.SpatialConvolution: 4-D; SpatialConvolutionMM: (Cout, Cin*k*k)


def _cnn_named(cnn):
    mods = _modules(cnn)
    convs = [m for m in mods if _cls(m).endswith(("SpatialConvolution", "SpatialConvolutionMM"))]
    bns = [m for m in mods if _cls(m).endswith("SpatialBatchNormalization")]
    assert len(convs) == 7 and len(bns) == 3, f"unexpected CNN: {len(convs)} convolutions, {len(bns)} batch-norms (cnn.lua:12-42)"
    named, stats, bi = {}, [], 0
    for (name, cin, cout, k, pad, bn), m in zip(CNN_LAYERS, convs):
        named[f"{name}.W"] = _conv_weight(m, cin, cout, k)
        named[f"{name}.b"] = np.asarray(m["bias"], dtype=np.float64)
        if bn:
            b = bns[bi]
            bi += 1
            named[f"bn{name[-1]}.gamma"] = np.asarray(b["weight"], dtype=np.float64)
            named[f"bn{name[-1]}.beta"] = np.asarray(b["bias"], dtype=np.float64)
            if b.get("running_var") is not None:
                var = np.asarray(b["running_var"], dtype=np.float64)
            else:                                   # nn before 2016 kept running_std = 1 / sqrt(var + eps)
                std = np.asarray(b["running_std"], dtype=np.float64)
                var = 1.0 / (std * std) - float(b.get("eps", 1e-5))
            stats.append((np.asarray(b["running_mean"], dtype=np.float64), var))
    return named, stats


def _graph_nodes(g):
    nodes = g.get("forwardnodes") or []
    return nodes if isinstance(nodes, list) else list(nodes.values())


def _parents(data):
    mi = data.get("mapindex") or {}
    if isinstance(mi, list):
        return [p for p in mi if isinstance(p, dict)]
    return [mi[i] for i in sorted(k for k in mi if isinstance(k, int))]


def _is_linear(m):
    return _cls(m) in ("nn.Linear", "nn.LinearNoBias")


def _lstm_named(g, use_attention):
    """the Linear / LookupTable / attention weights of one LSTM graph under the names of layout.param_specs"""
    datas = [n["data"] for n in _graph_nodes(g) if isinstance(n, t7.T7Object) and isinstance(n.get("data"), dict)]
    n_in = int(g.get("nInputs") or 0)
    named = {}
    pairs = []                                      # (i2h module, h2h module, index of the prev_h input)
    for d in datas:
        if _cls(d.get("module")) != "nn.CAddTable":
            continue
        ps = _parents(d)
        if len(ps) != 2 or not all(_is_linear(p.get("module")) for p in ps):
            continue
        feeds = []
        for p in ps:
            gp = _parents(p)
            feeds.append(gp[0].get("selectindex") if len(gp) == 1 else None)
        # h2h = the Linear fed straight by a graph input with the LARGER input index (prev_h[L] > x; LSTM.lua:45-47)
        cand = [i for i, f in enumerate(feeds) if f is not None]
        assert cand, "LSTM graph: no Linear of a gate sum is fed by a graph input (LSTM.lua:78-84)"
        h = max(cand, key=lambda i: feeds[i])
        pairs.append((ps[1 - h]["module"], ps[h]["module"], int(feeds[h])))
    pairs.sort(key=lambda t: t[2])                  # prev_h[1] < prev_h[2]: layer order
    if not use_attention:
        assert len(pairs) == 1, f"encoder graph: expected one LSTM layer, found {len(pairs)}"
        i2h, h2h, _ = pairs[0]
        named.update({"i2h.W": i2h["weight"], "i2h.b": i2h["bias"], "h2h.W": h2h["weight"], "h2h.b": h2h["bias"]})
        return {k: np.asarray(v, dtype=np.float64) for k, v in named.items()}
    assert len(pairs) == 2, f"decoder graph: expected two LSTM layers, found {len(pairs)}"
    assert n_in == 0 or pairs[0][2] in (4, 5), "decoder graph: unexpected input numbering (LSTM.lua:29-41)"
    for L, (i2h, h2h, _) in enumerate(pairs, start=1):
        named.update({f"l{L}.i2h.W": i2h["weight"], f"l{L}.i2h.b": i2h["bias"], f"l{L}.h2h.W": h2h["weight"], f"l{L}.h2h.b": h2h["bias"]})
    emb = [d["module"] for d in datas if _cls(d.get("module")) == "nn.LookupTable"]
    assert len(emb) == 1, "decoder graph: expected one LookupTable (LSTM.lua:56)"
    named["emb"] = emb[0]["weight"]
    attn = [d["module"] for d in datas if _cls(d.get("module")) == "nn.gModule"]
    assert len(attn) == 1, "decoder graph: expected the nested attention graph (LSTM.lua:110-113)"
    lin = [n["data"]["module"] for n in _graph_nodes(attn[0])
           if isinstance(n, t7.T7Object) and isinstance(n.get("data"), dict) and _is_linear(n["data"].get("module"))]
    for m in lin:
        w = np.asarray(m["weight"])
        named["attn.Wa" if w.shape[0] == w.shape[1] else "attn.Wc"] = w
    assert "attn.Wa" in named and "attn.Wc" in named, "attention graph: W_a (H,H) and W_c (H,2H) not found (LSTM.lua:130,155)"
    return {k: np.asarray(v, dtype=np.float64) for k, v in named.items()}


def _proj_named(proj):
    lin = [m for m in _modules(proj) if _cls(m) == "nn.Linear"]
    assert len(lin) == 1, "output projector: expected one Linear (output_projector.lua:5)"
    return {"W": np.asarray(lin[0]["weight"], dtype=np.float64), "b": np.asarray(lin[0]["bias"], dtype=np.float64)}


def load_checkpoint(path):
    """-> dict(config, params {group: flat float32, external layout}, bn_stats [(mean, var)] x 3, global_step, optim_state)"""
    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 20000))          # module graphs are deep object graphs
    try:
        ck = t7.load(path)
    finally:
        sys.setrecursionlimit(old)
    if isinstance(ck, dict) and ck.get("format") == FORMAT:
        config = dict(ck.get("config") or {})
        named = ck["params"]
        bn_stats = [(np.asarray(ck["bn"][k]["running_mean"], np.float32), np.asarray(ck["bn"][k]["running_var"], np.float32))
                    for k in BN_KEYS]
        step, opt = int(ck.get("global_step") or 0), dict(ck.get("optim_state") or {})
    elif isinstance(ck, dict) and isinstance(ck.get("params"), list) and len(ck["params"]) == 5:
        # what lua/model.lua's model:save writes: the five flat groups as they come out of aocr_get_params
        config = dict(ck.get("config") or {})
        flat = {g: np.asarray(ck["params"][i], np.float32).ravel() for i, g in enumerate(GROUPS)}
        bn_stats = [(np.asarray(b[0], np.float32), np.asarray(b[1], np.float32)) for b in ck["bn"]]
        return {"config": config, "params": flat, "bn_stats": bn_stats, "global_step": int(ck.get("global_step") or 0),
                "optim_state": dict(ck.get("optim_state") or {})}
    else:
        assert isinstance(ck, list) and len(ck) >= 2 and isinstance(ck[0], list) and len(ck[0]) == 5, \
            "not a torch-Attention-OCR checkpoint: expected {{5 modules}, config, global_step, optim_state} (model.lua:724)"
        mods, config = ck[0], dict(ck[1] or {})
        cnn_named, bn_stats = _cnn_named(mods[0])
        named = {"cnn": cnn_named, "enc_fw": _lstm_named(mods[1], False), "enc_bw": _lstm_named(mods[2], False),
                 "decoder": _lstm_named(mods[3], True), "proj": _proj_named(mods[4])}
        step = int(ck[2]) if len(ck) > 2 and ck[2] is not None else 0
        opt = dict(ck[3]) if len(ck) > 3 and isinstance(ck[3], dict) else {}
        # the embedding / hidden sizes are properties of the tensors; the saved config may predate a renamed option
        config.setdefault("encoder_num_hidden", int(named["enc_fw"]["h2h.W"].shape[1]))
        config.setdefault("target_embedding_size", int(named["decoder"]["emb"].shape[1]))
        config.setdefault("target_vocab_size", int(named["decoder"]["emb"].shape[0]))
        in1 = int(named["decoder"]["l1.i2h.W"].shape[1])
        config["input_feed"] = bool(in1 > int(named["decoder"]["emb"].shape[1]))
        bn_stats = [(np.asarray(m, np.float32), np.asarray(v, np.float32)) for m, v in bn_stats]
    params = {g: flatten(config, g, named[g]) for g in GROUPS}
    return {"config": config, "params": params, "bn_stats": bn_stats, "global_step": step, "optim_state": opt}


def group_names(config):
    return {g: [n for n, _ in param_specs(config)[g]] for g in GROUPS}
