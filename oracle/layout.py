"""Parameter layout + random-init recipe of the oracle (test infrastructure).

Groups follow the reference's optimiser contract: `self.layers = {cnn, encoder_fw,
encoder_bw, decoder, output_projector}` (src/model/model.lua:150) each flattened
to one vector by `getParameters()` (src/model/model.lua:161-168).  The intra-group
element order of Torch7's `getParameters()` is a nngraph topological-sort artefact
that `optim.sgd_list` cannot observe (it only takes norms and axpys,
src/optim/optim_sgd.lua:49-51,90), so we fix our own: module construction order of
the reference builder files, weight then bias, every tensor in Torch layout
(conv: (Cout,Cin,kH,kW); Linear: (out,in)).

Init distributions restate Torch7's `reset()` methods [T7, SURVEY App. B]:
  nn.Linear / nn.LinearNoBias (src/utils/model_utils.lua:68-85): U(-1/sqrt(in), 1/sqrt(in))
  SpatialConvolution: U(-1/sqrt(kW*kH*Cin), +)
  SpatialBatchNormalization: gamma ~ U(0,1), beta = 0, running_mean = 0, running_var = 1
  nn.LookupTable: N(0,1)
Values are drawn in float64 from numpy PCG64(seed) and rounded to float32 so the
GPU's fp32 master weights and the fp64 oracle hold *identical* numbers.
"""
from dataclasses import dataclass, asdict
import numpy as np

GROUPS = ["cnn", "enc_fw", "enc_bw", "decoder", "proj"]


@dataclass
class Config:
    """Fields `Model:create` reads from `opt` (src/model/model.lua:83-110; defaults src/train.lua:41-63)."""
    batch_size: int = 64
    max_encoder_l: int = 80
    max_decoder_l: int = 50
    encoder_num_hidden: int = 512
    encoder_num_layers: int = 1
    decoder_num_layers: int = 2
    target_vocab_size: int = 39
    target_embedding_size: int = 20
    input_feed: bool = True
    dropout: float = 0.0
    learning_rate: float = 0.1
    cnn_feature_size: int = 512  # src/model/model.lua:84

    @property
    def He(self):
        return self.encoder_num_hidden

    @property
    def Hd(self):
        return 2 * self.encoder_num_hidden  # src/model/model.lua:88

    def asdict(self):
        return asdict(self)


# (name, Cin, Cout, k, pad, bn?, pool) -- src/model/cnn.lua:12-42
CNN_LAYERS = [
    ("conv1", 1, 64, 3, 1, False, (2, 2)),
    ("conv2", 64, 128, 3, 1, False, (2, 2)),
    ("conv3", 128, 256, 3, 1, True, None),
    ("conv4", 256, 256, 3, 1, False, (2, 1)),   # SpatialMaxPooling(kW=1,kH=2,dW=1,dH=2): pool H only
    ("conv5", 256, 512, 3, 1, True, None),
    ("conv6", 512, 512, 3, 1, False, (2, 1)),
    ("conv7", 512, 512, 2, 0, True, None),
]


def source_len(W):
    """S = floor(floor(W/2)/2) - 1 (two 2x2 floor-mode pools then the 2x2 valid conv7)."""
    return (W // 2) // 2 - 1


def param_specs(cfg: Config):
    He, Hd, E, V = cfg.He, cfg.Hd, cfg.target_embedding_size, cfg.target_vocab_size
    assert cfg.encoder_num_layers == 1 and cfg.decoder_num_layers == 2, \
        "only the reference defaults (1 encoder layer, 2 decoder layers) are in scope"
    cnn = []
    for name, cin, cout, k, pad, bn, pool in CNN_LAYERS:
        cnn.append((f"{name}.W", (cout, cin, k, k)))
        cnn.append((f"{name}.b", (cout,)))
        if bn:
            cnn.append((f"bn{name[-1]}.gamma", (cout,)))
            cnn.append((f"bn{name[-1]}.beta", (cout,)))
    F = cfg.cnn_feature_size
    enc = [("i2h.W", (4 * He, F)), ("i2h.b", (4 * He,)), ("h2h.W", (4 * He, He)), ("h2h.b", (4 * He,))]
    in1 = E + (Hd if cfg.input_feed else 0)
    dec = [("emb", (V, E)),
           ("l1.i2h.W", (4 * Hd, in1)), ("l1.i2h.b", (4 * Hd,)), ("l1.h2h.W", (4 * Hd, Hd)), ("l1.h2h.b", (4 * Hd,)),
           ("l2.i2h.W", (4 * Hd, Hd)), ("l2.i2h.b", (4 * Hd,)), ("l2.h2h.W", (4 * Hd, Hd)), ("l2.h2h.b", (4 * Hd,)),
           ("attn.Wa", (Hd, Hd)), ("attn.Wc", (Hd, 2 * Hd))]
    proj = [("W", (V, Hd)), ("b", (V,))]
    return {"cnn": cnn, "enc_fw": enc, "enc_bw": list(enc), "decoder": dec, "proj": proj}


def group_sizes(cfg: Config):
    return {g: int(sum(int(np.prod(s)) for _, s in specs)) for g, specs in param_specs(cfg).items()}


def _fan_in(name, shape):
    if len(shape) == 4:
        return shape[1] * shape[2] * shape[3]
    return shape[1] if len(shape) == 2 else None


def init_params(cfg: Config, seed: int = 910820):
    """Returns {group: flat float32 array} drawn group by group, tensor by tensor."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    for g in GROUPS:
        chunks = []
        last_fan = None
        for name, shape in param_specs(cfg)[g]:
            n = int(np.prod(shape))
            if name == "emb":
                v = rng.standard_normal(n)
            elif name.endswith("gamma"):
                v = rng.uniform(0.0, 1.0, n)
            elif name.endswith("beta"):
                v = np.zeros(n)
            elif len(shape) >= 2:
                last_fan = _fan_in(name, shape)
                s = 1.0 / np.sqrt(last_fan)
                v = rng.uniform(-s, s, n)
            else:  # bias of the preceding weight: same stdv [T7]
                s = 1.0 / np.sqrt(last_fan)
                v = rng.uniform(-s, s, n)
            chunks.append(v.astype(np.float32))
        out[g] = np.concatenate(chunks)
    return out


def init_bn_stats(cfg: Config):
    """BN running stats are not parameters (side buffer): {layer: (running_mean, running_var)}."""
    return {"bn3": (np.zeros(256, np.float32), np.ones(256, np.float32)),
            "bn5": (np.zeros(512, np.float32), np.ones(512, np.float32)),
            "bn7": (np.zeros(512, np.float32), np.ones(512, np.float32))}


def unflatten(cfg: Config, group: str, flat):
    out, off = {}, 0
    for name, shape in param_specs(cfg)[group]:
        n = int(np.prod(shape))
        out[name] = flat[off:off + n].reshape(shape)
        off += n
    assert off == flat.shape[0], (group, off, flat.shape)
    return out


def flatten(cfg: Config, group: str, named):
    return np.concatenate([np.asarray(named[name]).reshape(-1) for name, _ in param_specs(cfg)[group]])
