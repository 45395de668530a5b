--[[ model.lua — drop-in replacement of src/model/model.lua: same class name, methods, arguments and return values,
     with the Torch7 module graph replaced by libaocr.so (LuaJIT FFI).  `th src/train.lua` keeps working unchanged:
       model = Model(); model:create(opt) | model:load(path, opt)          -- model.lua:45-112
       loss, stats = model:step(batch, forward_only, beam_size, trie)      -- model.lua:226
       model.global_step, model.optim_state.learningRate                   -- train.lua:83-89,115,163-166
       model.params[i], model.grad_params[i]  (norm/mul/add proxies)       -- consumed by optim.sgd_list
       model:vis(dir), model:save(path), model:shutdown()                  -- model.lua:708-731; train.lua:79,125,177,260
     Mirrors aocr/model.py line for line (that twin is the one exercised by the test-suite; LuaJIT/Torch7 are absent
     from the build image). ]]
local A = require 'aocr_ffi'
local K = require 'aocr_ckpt'
local ffi, lib = A.ffi, A.lib
local model = torch.class('Model')

-- src/train.lua:288 builds the dictionary with the global `loadDictionary` of src/utils/utils.lua (a nested tds.Hash);
-- the library takes the same trie as a flat child table.  utils.lua is loaded here first (train.lua has put src/utils on
-- package.path before it requires 'model'; the later `require 'utils'` of data_gen.lua is then a cache hit) and the
-- global is replaced, so an unmodified train.lua hands model:step the flat table.
require 'utils'
loadDictionary = function(dictionary_path, allow_digit_prefix) return A.loadDictionary(dictionary_path, allow_digit_prefix) end

local CONFIG_KEYS = {'dropout', 'encoder_num_hidden', 'encoder_num_layers', 'decoder_num_layers', 'target_vocab_size',
                     'target_embedding_size', 'max_encoder_l', 'max_decoder_l', 'input_feed', 'batch_size', 'prealloc'}
local BN_CHANNELS = {256, 512, 512}

function model:__init()
  if logging ~= nil then log = function(msg) logging:info(msg) end else log = print end   -- model.lua:36-42
end

-- params[i] / grad_params[i]: the tensor methods optim.sgd_list calls (optim_sgd.lua:49-51,90).  The handle is
-- captured per model instance (two models never alias).
local function proxy(h, group, is_grad)
  local p = {}
  function p:norm()
    local pn, gn = ffi.new('double[5]'), ffi.new('double[5]')
    A.check(h, lib.aocr_group_norms(h, pn, gn))
    return is_grad and gn[group] or pn[group]
  end
  function p:mul(s) A.check(h, lib.aocr_grad_scale(h, group, s)); return p end
  function p:add(a, other) A.check(h, lib.aocr_param_axpy(h, group, a)); return p end
  function p:float()          -- host copy (checkpoints, inspection)
    local sizes, n = ffi.new('int64_t[5]'), ffi.new('int32_t[1]')
    A.check(h, lib.aocr_param_groups(h, n, sizes))
    local t = torch.FloatTensor(tonumber(sizes[group]))
    if is_grad then A.check(h, lib.aocr_get_grads(h, group, t:data(), t:nElement()))
    else A.check(h, lib.aocr_get_params(h, group, t:data(), t:nElement())) end
    return t
  end
  return p
end

function model:_build()
  for _, k in ipairs(CONFIG_KEYS) do log(string.format('%s: %s', k, tostring(self[k]))) end   -- model.lua:116-128
  local c = ffi.new('aocr_config')
  c.batch_size = self.batch_size; c.max_encoder_l = self.max_encoder_l; c.max_decoder_l = self.max_decoder_l
  c.encoder_num_hidden = self.encoder_num_hidden; c.encoder_num_layers = self.encoder_num_layers
  c.decoder_num_layers = self.decoder_num_layers; c.target_vocab_size = self.target_vocab_size
  c.target_embedding_size = self.target_embedding_size; c.input_feed = self.input_feed and 1 or 0
  c.dropout = self.dropout or 0; c.learning_rate = self.optim_state.learningRate or 0.1
  c.dp_rank = 0; c.dp_world = 1; c.global_batch = 0; c.gemm_mode = 0
  local out = ffi.new('aocr_handle*[1]')
  local rc = lib.aocr_create(c, (gpu_id or 1) - 1, out)
  if rc ~= 0 then error(ffi.string(lib.aocr_last_error(nil))) end
  self.h = ffi.gc(out[0], lib.aocr_destroy)
  self.config = {}                                                                           -- saved by model:save
  for _, k in ipairs(CONFIG_KEYS) do self.config[k] = self[k] end
  self.params, self.grad_params = {}, {}
  for i = 1, 5 do self.params[i] = proxy(self.h, i - 1, false); self.grad_params[i] = proxy(self.h, i - 1, true) end
  self.visualize = false
end

function model:create(config)   -- model.lua:83-112
  for _, k in ipairs(CONFIG_KEYS) do self[k] = config[k] end
  self.global_step = 0
  self.optim_state = { learningRate = config.learning_rate }
  self:_build()
  -- fresh parameters with the distributions Torch7's module constructors draw (reset()); seeded like train.lua:60,223
  A.check(self.h, lib.aocr_init_params(self.h, (config.seed or 910820)))
end

-- Checkpoints.  model:save writes a Torch7-serialised table {params = {5 x FloatTensor}, bn = {3 x {mean, var}}, config,
-- global_step, optim_state}: the 5 flat vectors are what the reference's modules' getParameters() return
-- (model.lua:161-168).  model:load reads that, the named-tensor table of aocr/checkpoint.py, AND the reference's own
-- checkpoint {{cnn_model, encoder_fw, encoder_bw, decoder, output_projector}, config, global_step, optim_state}
-- (model.lua:45-80,720-725) — `-load_model` keeps working on a model directory trained by the reference.  To hand
-- weights back to the reference: lua/t7_convert.lua import.
function model:save(model_path)   -- model.lua:720-725
  local ck = { params = {}, bn = {}, config = self.config, global_step = self.global_step, optim_state = self.optim_state }
  for i = 1, 5 do ck.params[i] = self.params[i]:float() end
  for l = 1, 3 do
    local m, v = torch.FloatTensor(BN_CHANNELS[l]), torch.FloatTensor(BN_CHANNELS[l])
    A.check(self.h, lib.aocr_get_bn_stats(self.h, l - 1, m:data(), v:data(), BN_CHANNELS[l]))
    ck.bn[l] = { m, v }
  end
  torch.save(model_path, ck)
end

function model:load(model_path, config)   -- model.lua:45-80
  config = config or {}
  assert(paths.filep(model_path), string.format('Model %s does not exist!', model_path))
  -- classes a reference checkpoint deserialises into (train.lua:4-7 has loaded nn / nngraph / cudnn; nn.LinearNoBias
  -- lives in src/utils/model_utils.lua, CUDA tensors need cutorch)
  for _, name in ipairs({'nn', 'nngraph', 'cutorch', 'cunn', 'cudnn', 'model_utils'}) do pcall(require, name) end
  local ck = K.normalise(torch.load(model_path))
  for _, k in ipairs(CONFIG_KEYS) do self[k] = ck.config[k] end
  self.max_encoder_l = config.max_encoder_l or ck.config.max_encoder_l        -- model.lua:71-74
  self.max_decoder_l = config.max_decoder_l or ck.config.max_decoder_l
  self.batch_size = config.batch_size or ck.config.batch_size
  self.prealloc = config.prealloc
  self.global_step = ck.global_step
  self.optim_state = ck.optim_state
  self:_build()
  for i = 1, 5 do
    local t = ck.params[i]:float():contiguous()
    A.check(self.h, lib.aocr_set_params(self.h, i - 1, t:data(), t:nElement()))
  end
  for l = 1, 3 do
    local m, v = ck.bn[l][1]:float():contiguous(), ck.bn[l][2]:float():contiguous()
    A.check(self.h, lib.aocr_set_bn_stats(self.h, l - 1, m:data(), v:data(), BN_CHANNELS[l]))
  end
end

function model:vis(output_dir)   -- model.lua:708-718
  self.visualize = true
  self.visualize_path = paths.concat(output_dir, 'results.txt')
  local file, err = io.open(self.visualize_path, 'w')
  self.visualize_file = file
  if err then
    log(string.format('Error: visualize file %s cannot be created', self.visualize_path))
    self.visualize = false
    self.visualize_file = nil
  end
end

local function cut_at_eos(row, n)   -- labels up to (not including) the first EOS (3), utils.lua:145-171
  local t = {}
  for i = 1, n do
    if row[i] == 3 then break end
    t[#t + 1] = row[i]
  end
  return t
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
  beam_size = math.min(beam_size or 1, self.target_vocab_size)                                -- model.lua:228
  local labels = torch.IntTensor(b, self.max_decoder_l)
  local pred, gold = torch.DoubleTensor(b), torch.DoubleTensor(b)
  local nc = ffi.new('int32_t[1]')
  if beam_size == 1 and trie == nil then
    A.check(self.h, lib.aocr_decode_greedy(self.h, images:data(), b, W, targets:data(), targets_eval:data(), T,
                                           labels:data(), pred:data(), gold:data(), loss, nc))
  else
    -- beam search, optionally constrained to a dictionary trie (model.lua:380-387,405-445,460-514): `trie` is what
    -- aocr_ffi.loadDictionary returns (utils.lua:177-218 as a flat child table)
    A.check(self.h, lib.aocr_decode_beam(self.h, images:data(), b, W, targets:data(), targets_eval:data(), T,
                                         beam_size, trie and trie.table or nil, trie and trie.nodes or 0,
                                         labels:data(), pred:data(), gold:data(), loss, nc))
  end
  if self.visualize and self.visualize_file then                                               -- model.lua:628-633
    local img_paths = batch[5]
    local te = torch.IntTensor(b, self.max_decoder_l):fill(1)
    te[{{}, {1, T}}]:copy(targets_eval)
    for i = 1, b do
      self.visualize_file:write(string.format('%s\t%s\t%s\t%f\t%f\n', img_paths[i],
        numlist2str(cut_at_eos(te[i], self.max_decoder_l)), numlist2str(cut_at_eos(labels[i], self.max_decoder_l)),
        pred[i], gold[i]))
    end
    self.visualize_file:flush()
  end
  return loss[0], {num_nonzeros, nc[0]}
end

function model:shutdown()   -- model.lua:727-731
  if self.visualize_file then self.visualize_file:close(); self.visualize_file = nil end
  self.h = nil
  collectgarbage()
end
