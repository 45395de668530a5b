"""`Model` — Python mirror of the reference's Lua class (src/model/model.lua), same method names,
argument meaning, return tuple and assert texts, over the C ABI of libaocr.so.

    model = Model(); model.create(opt)            # model.lua:36,83-112
    loss, stats = model.step(batch, forward_only) # model.lua:226 -> loss*batch_size, {num_nonzeros, accuracy}
    model.global_step, model.optim_state["learningRate"]   # fields the train loop reads/mutates (train.lua:83-89)
    model.params[i] / model.grad_params[i]        # the 5 flat vectors optim.sgd_list consumes (model.lua:161-168)
"""
import json
import os

import numpy as np

from .capi import AocrConfig, AocrError, Handle, GROUPS
from .optim import sgd_list

_DEFAULTS = dict(batch_size=400, max_encoder_l=80, max_decoder_l=50, encoder_num_hidden=512, encoder_num_layers=1,
                 decoder_num_layers=2, target_vocab_size=39, target_embedding_size=20, input_feed=False, dropout=0.0,
                 learning_rate=0.1, prealloc=False)   # src/train.lua:41-63


class _GroupProxy:
    """Stands in for the flat Torch tensor `params[i]` / `grad_params[i]`: the subset of tensor methods
    optim.sgd_list calls (optim_sgd.lua:49-51,90) forwards to the device."""

    def __init__(self, model, group, grads):
        self.model, self.group, self.grads = model, group, grads

    def size(self):
        return self.model.handle.group_sizes[self.group]

    def norm(self):
        pn, gn = self.model.handle.group_norms()
        return gn[self.group] if self.grads else pn[self.group]

    def mul(self, s):
        assert self.grads, "only gradients are scaled by sgd_list"
        self.model.handle.grad_scale(self.group, float(s))
        return self

    def add(self, a, other):
        assert not self.grads and other.grads and other.group == self.group
        self.model.handle.param_axpy(self.group, float(a))
        return self

    def numpy(self):
        h = self.model.handle
        return h.get_grads(self.group) if self.grads else h.get_params(self.group)


class Model:
    def __init__(self, log=print, device=0, gemm_mode=0, dp_rank=0, dp_world=1, global_batch=0):
        self.log = log
        self.device = device
        self.gemm_mode, self.dp_rank, self.dp_world, self.global_batch = gemm_mode, dp_rank, dp_world, global_batch
        self.handle = None
        self.visualize = False
        self.visualize_file = None

    # model:create(config) — model.lua:83-112: builds the modules, whose constructors draw the initial weights
    def create(self, config):
        c = dict(_DEFAULTS)
        c.update({k: v for k, v in dict(config).items() if k in _DEFAULTS})
        self.config = c
        self.global_step = 0
        self.optim_state = {"learningRate": c["learning_rate"]}
        self._build()
        self.handle.init_params(dict(config).get("seed", 910820))   # Torch7 reset() distributions, seeded (train.lua:60)
        return self

    def _build(self):
        c = self.config
        for k in ("dropout", "encoder_num_hidden", "encoder_num_layers", "decoder_num_layers", "target_vocab_size",
                  "target_embedding_size", "max_encoder_l", "max_decoder_l", "input_feed", "batch_size", "prealloc"):
            self.log("%s: %s" % (k, c[k]))                      # model.lua:116-128
        cfg = AocrConfig(batch_size=c["batch_size"], max_encoder_l=c["max_encoder_l"], max_decoder_l=c["max_decoder_l"],
                         encoder_num_hidden=c["encoder_num_hidden"], encoder_num_layers=c["encoder_num_layers"],
                         decoder_num_layers=c["decoder_num_layers"], target_vocab_size=c["target_vocab_size"],
                         target_embedding_size=c["target_embedding_size"], input_feed=1 if c["input_feed"] else 0,
                         dropout=c["dropout"], learning_rate=c["learning_rate"], dp_rank=self.dp_rank,
                         dp_world=self.dp_world, global_batch=self.global_batch, gemm_mode=self.gemm_mode)
        self.handle = Handle(cfg, self.device)
        self.params = [_GroupProxy(self, i, False) for i in range(5)]       # model.lua:161-168
        self.grad_params = [_GroupProxy(self, i, True) for i in range(5)]
        self.log("Number of parameters: %d" % sum(self.handle.group_sizes))

    # parameter import (parity runs load the oracle's weights; checkpoints)
    def set_parameters(self, params, bn_stats=None):
        for i, g in enumerate(GROUPS):
            self.handle.set_params(i, params[g] if isinstance(params, dict) else params[i])
        if bn_stats is not None:
            for i, k in enumerate(("bn3", "bn5", "bn7")):
                self.handle.set_bn_stats(i, *bn_stats[k])

    def get_parameters(self):
        return {g: self.handle.get_params(i) for i, g in enumerate(GROUPS)}

    def get_gradients(self):
        return {g: self.handle.get_grads(i) for i, g in enumerate(GROUPS)}

    # model:step(batch, forward_only, beam_size, trie) — model.lua:226-706
    def step(self, batch, forward_only, beam_size=1, trie=None, use_lua_optim=False):
        images, targets, targets_eval, num_nonzeros = batch[0], batch[1], batch[2], batch[3]
        if forward_only:
            beam_size = min(beam_size or 1, self.config["target_vocab_size"])        # model.lua:228-229
        try:
            if not forward_only:
                if use_lua_optim:   # unmodified optim.sgd_list semantics over the proxies (optim_sgd.lua)
                    def feval(_):
                        loss_sum = self.handle.forward_backward(images, targets, targets_eval)
                        return loss_sum / images.shape[0], self.grad_params, [num_nonzeros, 0.0]
                    _, loss, stats = sgd_list(feval, self.params, self.optim_state)
                    return loss[0] * images.shape[0], stats                      # model.lua:700-701
                loss_sum = self.handle.train_step(images, targets, targets_eval, self.optim_state["learningRate"])
                return loss_sum, [num_nonzeros, 0.0]
            if beam_size == 1 and trie is None:
                out = self.handle.decode_greedy(images, targets, targets_eval)
            else:   # beam search / dictionary-constrained decode (model.lua:380-387,405-445,460-514); trie: aocr.Trie
                out = self.handle.decode_beam(images, targets, targets_eval, beam_size, trie)
        except AocrError as e:
            if e.code == -1:
                raise AssertionError(e.msg) from e   # the reference raises Lua asserts (model.lua:264,287)
            raise
        if self.visualize and self.visualize_file:
            from .data import numlist2str
            paths = batch[4]
            L = self.config["max_decoder_l"]
            te = np.ones((images.shape[0], L), np.int64)
            te[:, :targets_eval.shape[1]] = targets_eval
            for i in range(len(paths)):                                          # model.lua:628-633
                self.visualize_file.write("%s\t%s\t%s\t%f\t%f\n" % (
                    paths[i], numlist2str(_cut(te[i])), numlist2str(_cut(out["labels"][i])),
                    out["pred_scores"][i], out["gold_scores"][i]))
            self.visualize_file.flush()
        self.last_decode = out
        return out["loss_sum"], [num_nonzeros, float(out["num_correct"])]         # model.lua:703-704

    # model:vis(output_dir) — model.lua:708-718
    def vis(self, output_dir):
        self.visualize = True
        os.makedirs(output_dir, exist_ok=True)
        self.visualize_path = os.path.join(output_dir, "results.txt")
        try:
            self.visualize_file = open(self.visualize_path, "w")
        except OSError:
            self.log("Error: visualize file %s cannot be created" % self.visualize_path)
            self.visualize, self.visualize_file = False, None

    # model:save(model_path) — model.lua:720-725.  A path ending in ".t7" is written in Torch7's container (the named
    # tensors of every group, aocr/checkpoint.py; lua/t7_convert.lua rebuilds the reference's module trees from it);
    # anything else goes into one .npz holding the same four items {layers, config, global_step, optim_state}.
    def save(self, model_path):
        p = self.get_parameters()
        bn = [self.handle.get_bn_stats(i) for i in range(3)]
        if model_path.endswith(".t7"):
            from .checkpoint import save_checkpoint
            return save_checkpoint(model_path, self.config, p, bn, self.global_step, self.optim_state)
        path = model_path if model_path.endswith(".npz") else model_path + ".npz"
        np.savez(path,
                 **{"param_" + g: p[g] for g in GROUPS},
                 **{"bn%d_mean" % i: bn[i][0] for i in range(3)}, **{"bn%d_var" % i: bn[i][1] for i in range(3)},
                 config=json.dumps(self.config), global_step=self.global_step,
                 optim_state=json.dumps(self.optim_state))
        return path

    # model:load(model_path, config) — model.lua:45-80.  ".t7": a reference checkpoint (its five module trees) or the
    # named-tensor table written by save().
    def load(self, model_path, config=None):
        if model_path.endswith(".t7"):
            assert os.path.isfile(model_path), "Model %s does not exist!" % model_path    # model.lua:50
            from .checkpoint import load_checkpoint
            ck = load_checkpoint(model_path)
            saved_cfg, params, bn = ck["config"], ck["params"], ck["bn_stats"]
            step, opt = ck["global_step"], ck["optim_state"]
        else:
            path = model_path if model_path.endswith(".npz") else model_path + ".npz"
            assert os.path.isfile(path), "Model %s does not exist!" % model_path      # model.lua:50
            z = np.load(path, allow_pickle=False)
            saved_cfg, step, opt = json.loads(str(z["config"])), int(z["global_step"]), json.loads(str(z["optim_state"]))
            params = {g: z["param_" + g] for g in GROUPS}
            bn = [(z["bn%d_mean" % i], z["bn%d_var" % i]) for i in range(3)]
        c = dict(_DEFAULTS)
        c.update(saved_cfg)
        for k in ("max_encoder_l", "max_decoder_l", "batch_size", "prealloc"):    # model.lua:71-74
            if config and k in config:
                c[k] = config[k]
        self.config = c
        self.global_step = int(step)
        self.optim_state = dict(opt) if opt else {"learningRate": c.get("learning_rate", 0.1)}
        self._build()
        for i, g in enumerate(GROUPS):
            self.handle.set_params(i, params[g])
        for i in range(3):
            self.handle.set_bn_stats(i, bn[i][0], bn[i][1])
        return self

    # model:shutdown() — model.lua:727-731
    def shutdown(self):
        if self.visualize_file:
            self.visualize_file.close()
            self.visualize_file = None
        if self.handle:
            self.handle.close()
            self.handle = None


def _cut(ids):
    out = []
    for v in ids:
        if int(v) == 3:
            break
        out.append(int(v))
    return out
