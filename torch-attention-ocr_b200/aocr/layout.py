"""External ("Torch") layout of the five parameter groups `{cnn, enc_fw, enc_bw, decoder, proj}` (src/model/model.lua:150)
that `aocr_set_params` / `aocr_get_params` exchange: tensors in module construction order (src/model/cnn.lua:9-45,
src/model/LSTM.lua:41-105,124-162, src/model/output_projector.lua:4-6), weight then bias, Torch tensor shapes
(convolutions `(Cout, Cin, kH, kW)`, Linear `(out, in)`), each group flattened row-major.  Used by the checkpoint
converters; the library itself permutes into its native layout on import."""
import numpy as np

GROUPS = ("cnn", "enc_fw", "enc_bw", "decoder", "proj")

# (name, Cin, Cout, kernel, pad, batch-norm?)   src/model/cnn.lua:12-42
CNN_LAYERS = [("conv1", 1, 64, 3, 1, False), ("conv2", 64, 128, 3, 1, False), ("conv3", 128, 256, 3, 1, True),
              ("conv4", 256, 256, 3, 1, False), ("conv5", 256, 512, 3, 1, True), ("conv6", 512, 512, 3, 1, False),
              ("conv7", 512, 512, 2, 0, True)]


def param_specs(config):
    """{group: [(tensor name, shape), ...]} for a model config (dict with the reference's option names)"""
    He = int(config.get("encoder_num_hidden", 512))
    Hd, E, V = 2 * He, int(config.get("target_embedding_size", 20)), int(config.get("target_vocab_size", 39))
    assert int(config.get("encoder_num_layers", 1)) == 1 and int(config.get("decoder_num_layers", 2)) == 2, \
        "only the reference defaults (1 encoder layer, 2 decoder layers) are supported"
    cnn = []
    for name, cin, cout, k, pad, bn in CNN_LAYERS:
        cnn += [(f"{name}.W", (cout, cin, k, k)), (f"{name}.b", (cout,))]
        if bn:
            cnn += [(f"bn{name[-1]}.gamma", (cout,)), (f"bn{name[-1]}.beta", (cout,))]
    enc = [("i2h.W", (4 * He, 512)), ("i2h.b", (4 * He,)), ("h2h.W", (4 * He, He)), ("h2h.b", (4 * He,))]
    in1 = E + (Hd if config.get("input_feed", True) else 0)
    dec = [("emb", (V, E)),
           ("l1.i2h.W", (4 * Hd, in1)), ("l1.i2h.b", (4 * Hd,)), ("l1.h2h.W", (4 * Hd, Hd)), ("l1.h2h.b", (4 * Hd,)),
           ("l2.i2h.W", (4 * Hd, Hd)), ("l2.i2h.b", (4 * Hd,)), ("l2.h2h.W", (4 * Hd, Hd)), ("l2.h2h.b", (4 * Hd,)),
           ("attn.Wa", (Hd, Hd)), ("attn.Wc", (Hd, 2 * Hd))]
    proj = [("W", (V, Hd)), ("b", (V,))]
    return {"cnn": cnn, "enc_fw": enc, "enc_bw": list(enc), "decoder": dec, "proj": proj}


def flatten(config, group, named):
    out = []
    for name, shape in param_specs(config)[group]:
        a = np.asarray(named[name])
        assert tuple(a.shape) == tuple(shape), f"{group}.{name}: shape {a.shape}, expected {shape}"
        out.append(a.astype(np.float32).ravel())
    return np.concatenate(out)


def unflatten(config, group, flat):
    out, off = {}, 0
    for name, shape in param_specs(config)[group]:
        n = int(np.prod(shape))
        out[name] = np.asarray(flat[off:off + n]).reshape(shape)
        off += n
    assert off == len(flat), f"{group}: {len(flat)} values for a layout of {off}"
    return out
