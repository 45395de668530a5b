--[[ model.lua — drop-in replacement of src/model/model.lua: same class name, methods, arguments and return values,
     with the Torch7 module graph replaced by libaocr.so (LuaJIT FFI).  `th src/train.lua` keeps working unchanged:
       model = Model(); model:create(opt) | model:load(path, opt)
       loss, stats = model:step(batch, forward_only, beam_size, trie)      -- model.lua:226
       model.global_step, model.optim_state.learningRate                   -- train.lua:83-89,115,163-166
       model.params[i], model.grad_params[i]  (norm/mul/add proxies)       -- consumed by optim.sgd_list
     Mirrors aocr/model.py line for line (that twin is the one exercised by the test-suite). ]]
local A = require 'aocr_ffi'
local ffi, lib = A.ffi, A.lib
local model = torch.class('Model')

function model:__init()
  if logging ~= nil then log = function(msg) logging:info(msg) end else log = print end   -- model.lua:36-42
end

local function proxy(self, group, is_grad)   -- the tensor methods optim.sgd_list calls (optim_sgd.lua:49-51,90)
  local p = {}
  function p:norm()
    local pn, gn = ffi.new('double[5]'), ffi.new('double[5]')
    A.check(self_h, lib.aocr_group_norms(self_h, pn, gn))
    return is_grad and gn[group] or pn[group]
  end
  function p:mul(s) A.check(self_h, lib.aocr_grad_scale(self_h, group, s)); return p end
  function p:add(a, other) A.check(self_h, lib.aocr_param_axpy(self_h, group, a)); return p end
  return p
end

function model:_build()
  local c = ffi.new('aocr_config')
  c.batch_size = self.batch_size; c.max_encoder_l = self.max_encoder_l; c.max_decoder_l = self.max_decoder_l
  c.encoder_num_hidden = self.encoder_num_hidden; c.encoder_num_layers = self.encoder_num_layers
  c.decoder_num_layers = self.decoder_num_layers; c.target_vocab_size = self.target_vocab_size
  c.target_embedding_size = self.target_embedding_size; c.input_feed = self.input_feed and 1 or 0
  c.dropout = self.dropout; c.learning_rate = self.optim_state.learningRate or 0.1
  c.dp_rank = 0; c.dp_world = 1; c.global_batch = 0; c.gemm_mode = 0
  local out = ffi.new('aocr_handle*[1]')
  local rc = lib.aocr_create(c, (gpu_id or 1) - 1, out)
  if rc ~= 0 then error(ffi.string(lib.aocr_last_error(nil))) end
  self.h = ffi.gc(out[0], lib.aocr_destroy)
  self_h = self.h
  self.params, self.grad_params = {}, {}
  for i = 1, 5 do self.params[i] = proxy(self, i - 1, false); self.grad_params[i] = proxy(self, i - 1, true) end
  self.visualize = false
end

function model:create(config)   -- model.lua:83-112
  self.dropout = config.dropout; self.encoder_num_hidden = config.encoder_num_hidden
  self.encoder_num_layers = config.encoder_num_layers; self.decoder_num_layers = config.decoder_num_layers
  self.target_vocab_size = config.target_vocab_size; self.target_embedding_size = config.target_embedding_size
  self.max_encoder_l = config.max_encoder_l; self.max_decoder_l = config.max_decoder_l
  self.input_feed = config.input_feed; self.batch_size = config.batch_size; self.prealloc = config.prealloc
  self.global_step = 0
  self.optim_state = { learningRate = config.learning_rate }
  self:_build()
  -- fresh parameters: drawn by the caller with Torch7's reset() distributions and pushed with aocr_set_params
end

function model:step(batch, forward_only, beam_size, trie)   -- model.lua:226-706
  local images, targets, targets_eval, num_nonzeros = batch[1]:float():contiguous(), batch[2]:int():contiguous(),
                                                      batch[3]:int():contiguous(), batch[4]
  local b, W, T = images:size(1), images:size(4), targets:size(2)
  local loss = ffi.new('double[1]')
  if not forward_only then
    A.check(self.h, lib.aocr_train_step(self.h, images:data(), b, W, targets:data(), targets_eval:data(), T,
                                        self.optim_state.learningRate, loss))
    return loss[0], {num_nonzeros, 0}
  end
  assert((beam_size or 1) == 1 and trie == nil, 'beam search / dictionary decode: use the reference path (out of scope)')
  local labels = torch.IntTensor(b, self.max_decoder_l)
  local pred, gold = torch.DoubleTensor(b), torch.DoubleTensor(b)
  local nc = ffi.new('int32_t[1]')
  A.check(self.h, lib.aocr_decode_greedy(self.h, images:data(), b, W, targets:data(), targets_eval:data(), T,
                                         labels:data(), pred:data(), gold:data(), loss, nc))
  return loss[0], {num_nonzeros, nc[0]}
end

function model:shutdown() self.h = nil; collectgarbage() end   -- model.lua:727-731
