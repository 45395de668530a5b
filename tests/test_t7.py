"""§8f-3: Torch7 container reader / writer (aocr/t7.py) and the checkpoint converters (aocr/checkpoint.py;
src/model/model.lua:45-80,720-725).  No Torch7 exists in the image: the byte-level case below is assembled by hand from
the format of torch7's File.lua / Tensor.c / Storage.c, and the module-tree case builds the object graph that
`torch.save` writes for the reference's five modules (nn.Sequential.modules, nngraph forwardnodes / mapindex /
selectindex) around known tensors."""
import struct

import numpy as np

from oracle import Config, init_params, init_bn_stats
from oracle import layout as olayout


def _cfgdict(input_feed=True):
    return dict(encoder_num_hidden=512, encoder_num_layers=1, decoder_num_layers=2, target_vocab_size=39,
                target_embedding_size=20, input_feed=input_feed, dropout=0.0, batch_size=4, max_encoder_l=30, max_decoder_l=10)


def test_product_layout_equals_oracle_layout():
    from aocr import layout
    for feed in (True, False):
        o = olayout.param_specs(Config(input_feed=feed))
        p = layout.param_specs(_cfgdict(feed))
        assert list(o) == list(p) == list(layout.GROUPS)
        for g in o:
            assert [(n, tuple(s)) for n, s in o[g]] == [(n, tuple(s)) for n, s in p[g]], g


def test_reads_a_hand_assembled_torch_file(tmp_path):
    # { [1] = DoubleTensor(2,3) filled 0..5, a = "xy", b = true }   (File.lua:writeObject, Tensor.c:write, Storage.c:write)
    i32 = lambda v: struct.pack("<i", v)
    i64 = lambda v: struct.pack("<q", v)
    s = lambda t: i32(len(t)) + t.encode()
    tensor = (i32(4) + i32(2) + s("V 1") + s("torch.DoubleTensor") + i32(2) + i64(2) + i64(3) + i64(3) + i64(1) + i64(1) +
              i32(4) + i32(3) + s("V 1") + s("torch.DoubleStorage") + i64(6) + struct.pack("<6d", *range(6)))
    blob = (i32(3) + i32(1) + i32(3) +
            i32(1) + struct.pack("<d", 1.0) + tensor +
            i32(2) + s("a") + i32(2) + s("xy") +
            i32(2) + s("b") + i32(5) + i32(1))
    p = tmp_path / "hand.t7"
    p.write_bytes(blob)
    from aocr import t7
    v = t7.load(str(p))
    assert v["a"] == "xy" and v["b"] is True
    assert np.array_equal(v[1], np.arange(6, dtype=np.float64).reshape(2, 3))
    # the writer emits the same bytes for the same value
    q = tmp_path / "again.t7"
    t7.save(str(q), {1: np.arange(6, dtype=np.float64).reshape(2, 3), "a": "xy", "b": True})
    assert q.read_bytes() == blob


def test_round_trip_shares_objects_and_keeps_types(tmp_path):
    from aocr import t7
    w = np.random.default_rng(0).standard_normal((3, 4)).astype(np.float32)
    obj = t7.T7Object("nn.Linear", {"weight": w, "bias": np.zeros(3), "train": False})
    v = {"mods": [obj, obj], "n": 7, "x": 0.5, "none_list": [1, "two", None, 4.25], "ids": np.arange(5, dtype=np.int32)}
    p = tmp_path / "rt.t7"
    t7.save(str(p), v)
    r = t7.load(str(p))
    assert r["mods"][0] is r["mods"][1] and r["mods"][0].cls == "nn.Linear"
    assert r["mods"][0]["weight"].dtype == np.float32 and np.array_equal(r["mods"][0]["weight"], w)
    assert r["n"] == 7 and r["x"] == 0.5 and r["ids"].dtype == np.int32
    # a table with a nil hole is not a sequence: it comes back keyed (Lua's pairs skips the nil)
    assert r["none_list"] == {1: 1, 2: "two", 4: 4.25} or r["none_list"] == [1, "two", None, 4.25]


def test_named_tensor_checkpoint_round_trip(tmp_path):
    from aocr.checkpoint import load_checkpoint, save_checkpoint
    cfg = Config()
    params, bn = init_params(cfg, 3), init_bn_stats(cfg)
    p = str(tmp_path / "ck.t7")
    save_checkpoint(p, _cfgdict(), params, [bn[k] for k in ("bn3", "bn5", "bn7")], global_step=12, optim_state={"learningRate": 0.05})
    ck = load_checkpoint(p)
    assert ck["global_step"] == 12 and ck["optim_state"]["learningRate"] == 0.05 and ck["config"]["input_feed"] is True
    for g in params:
        assert np.array_equal(ck["params"][g], params[g]), g


# ---- the object graph torch.save writes for the reference's modules -------------------------------------------------
def _node(data, nodes):
    from aocr import t7
    n = t7.T7Object("nngraph.Node", {"data": data, "children": [], "id": len(nodes) + 1, "visited": False})
    nodes.append(n)
    return n


def _apply(module_cls, fields, parent_datas, nodes, extra=None):
    """node = module(parents): nngraph records the parents in data.mapindex, both ways (nngraph/node.lua)"""
    from aocr import t7
    data = {"module": t7.T7Object(module_cls, dict(fields)), "input": [], "mapindex": {}}
    for i, pd in enumerate(parent_datas, start=1):
        data["mapindex"][i] = pd
        data["mapindex"][t7._TableKey(pd)] = i
    data.update(extra or {})
    _node(data, nodes)
    return data


def _lstm_graph(named, prefix_layers, use_attention, input_feed, shuffle_seed):
    """createLSTM (LSTM.lua:18-122) as a serialised nn.gModule; node order shuffled (the order is a Torch artefact)"""
    from aocr import t7
    nodes = []
    offset = (2 if input_feed else 1) if use_attention else 0
    nlayers = len(prefix_layers)
    n_in = 1 + offset + 2 * nlayers
    inn = {"module": t7.T7Object("nn.Identity", {}), "input": [], "mapindex": {}}
    _node(inn, nodes)
    inputs = [_apply("nn.Identity", {}, [inn], nodes, {"selectindex": i}) for i in range(1, n_in + 1)]
    outputs = []
    x = None
    for L, pre in enumerate(prefix_layers, start=1):
        prev_h, prev_c = inputs[L * 2 + 1 + offset - 1], inputs[L * 2 + offset - 1]
        if L == 1:
            x = inputs[0]
            if use_attention:
                x = _apply("nn.LookupTable", {"weight": named["emb"]}, [x], nodes)
                if input_feed:
                    x = _apply("nn.JoinTable", {"dimension": 2}, [x, inputs[offset]], nodes)
        else:
            x = _apply("nn.Dropout", {"p": 0}, [outputs[-1]], nodes)
        i2h = _apply("nn.Linear", {"weight": named[pre + "i2h.W"], "bias": named[pre + "i2h.b"]}, [x], nodes)
        h2h = _apply("nn.Linear", {"weight": named[pre + "h2h.W"], "bias": named[pre + "h2h.b"]}, [prev_h], nodes)
        sums = _apply("nn.CAddTable", {}, [i2h, h2h], nodes)
        resh = _apply("nn.Reshape", {}, [sums], nodes)
        split = _apply("nn.SplitTable", {}, [resh], nodes)
        gates = [_apply("nn.Sigmoid" if k < 3 else "nn.Tanh", {}, [split], nodes, {"selectindex": k + 1}) for k in range(4)]
        fc = _apply("nn.CMulTable", {}, [gates[1], prev_c], nodes)
        ig = _apply("nn.CMulTable", {}, [gates[0], gates[3]], nodes)
        next_c = _apply("nn.CAddTable", {}, [fc, ig], nodes)          # a CAddTable whose parents are NOT Linear
        next_h = _apply("nn.CMulTable", {}, [gates[2], _apply("nn.Tanh", {}, [next_c], nodes)], nodes)
        outputs += [next_c, next_h]
    if use_attention:
        an = []
        a_in = {"module": t7.T7Object("nn.Identity", {}), "input": [], "mapindex": {}}
        _node(a_in, an)
        a1 = _apply("nn.Identity", {}, [a_in], an, {"selectindex": 1})
        a2 = _apply("nn.Identity", {}, [a_in], an, {"selectindex": 2})
        tt = _apply("nn.LinearNoBias", {"weight": named["attn.Wa"]}, [a1], an)
        mm = _apply("nn.MM", {}, [a2, _apply("nn.Replicate", {}, [tt], an)], an)
        sm = _apply("nn.SoftMax", {}, [_apply("nn.Sum", {}, [mm], an)], an)
        cc = _apply("nn.Sum", {}, [_apply("nn.MM", {}, [_apply("nn.Replicate", {}, [sm], an), a2], an)], an)
        jt = _apply("nn.JoinTable", {}, [cc, a1], an)
        _apply("nn.Tanh", {}, [_apply("nn.LinearNoBias", {"weight": named["attn.Wc"]}, [jt], an)], an)
        attn = {"forwardnodes": an, "nInputs": 2, "name": "decoder_attn"}
        _apply("nn.gModule", attn, [outputs[-1], inputs[1]], nodes)
    rng = np.random.default_rng(shuffle_seed)
    order = list(rng.permutation(len(nodes)))
    return t7.T7Object("nn.gModule", {"forwardnodes": [nodes[i] for i in order], "nInputs": n_in})


def _reference_checkpoint(cfgd, params, bn):
    from aocr import t7
    from aocr.layout import CNN_LAYERS, unflatten
    named = {g: {k: v.astype(np.float64) for k, v in unflatten(cfgd, g, params[g]).items()} for g in params}
    mods = [t7.T7Object("nn.AddConstant", {"constant_scalar": -128}), t7.T7Object("nn.MulConstant", {"constant_scalar": 1 / 128})]
    for name, cin, cout, k, pad, isbn in CNN_LAYERS:
        mods.append(t7.T7Object("cudnn.SpatialConvolution", {"weight": named["cnn"][name + ".W"], "bias": named["cnn"][name + ".b"],
                                                             "nInputPlane": cin, "nOutputPlane": cout, "kW": k, "kH": k}))
        if isbn:
            key = "bn" + name[-1]
            mods.append(t7.T7Object("nn.SpatialBatchNormalization", {
                "weight": named["cnn"][key + ".gamma"], "bias": named["cnn"][key + ".beta"],
                "running_mean": np.asarray(bn[key][0], np.float64), "running_var": np.asarray(bn[key][1], np.float64), "eps": 1e-5}))
        mods.append(t7.T7Object("cudnn.ReLU", {}))
    cnn = t7.T7Object("nn.Sequential", {"modules": mods})
    feed = bool(cfgd["input_feed"])
    enc_fw = _lstm_graph(named["enc_fw"], [""], False, False, 1)
    enc_bw = _lstm_graph(named["enc_bw"], [""], False, False, 2)
    dec = _lstm_graph(named["decoder"], ["l1.", "l2."], True, feed, 3)
    proj = t7.T7Object("nn.Sequential", {"modules": [t7.T7Object("nn.Linear", {"weight": named["proj"]["W"], "bias": named["proj"]["b"]}),
                                                     t7.T7Object("nn.LogSoftMax", {})]})
    cfg_saved = {k: v for k, v in cfgd.items() if k not in ("target_vocab_size",)}      # sizes come from the tensors
    return [[cnn, enc_fw, enc_bw, dec, proj], cfg_saved, 321, {"learningRate": 0.0125, "evalCounter": 321}]


def test_reference_module_tree_checkpoint_is_mapped_by_structure(tmp_path):
    from aocr import t7
    from aocr.checkpoint import load_checkpoint
    for feed in (True, False):
        cfg = Config(input_feed=feed)
        cfgd = _cfgdict(feed)
        params, bn = init_params(cfg, 11), init_bn_stats(cfg)
        bn = {k: (np.random.default_rng(5).standard_normal(len(v[0])), np.random.default_rng(6).uniform(0.5, 2, len(v[1]))) for k, v in bn.items()}
        p = str(tmp_path / f"ref_{feed}.t7")
        t7.save(p, _reference_checkpoint(cfgd, params, bn))
        ck = load_checkpoint(p)
        assert ck["global_step"] == 321 and ck["optim_state"]["learningRate"] == 0.0125
        assert ck["config"]["input_feed"] is feed and ck["config"]["target_vocab_size"] == 39
        for g in params:
            assert np.array_equal(ck["params"][g], params[g]), g
        for i, k in enumerate(("bn3", "bn5", "bn7")):
            assert np.allclose(ck["bn_stats"][i][0], bn[k][0], atol=1e-6) and np.allclose(ck["bn_stats"][i][1], bn[k][1], atol=1e-6)
