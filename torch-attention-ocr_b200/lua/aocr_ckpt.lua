--[[ aocr_ckpt.lua — where the weights live in the reference's checkpoints, and the external layout of the five flat
     parameter groups (kept in step with aocr/layout.py and aocr/checkpoint.py, the twins the test-suite executes).

     A reference checkpoint is { {cnn_model, encoder_fw, encoder_bw, decoder, output_projector}, config, global_step,
     optim_state } (src/model/model.lua:724).  Modules are located structurally: nn.Sequential.modules in construction
     order for the CNN (cnn.lua:9-45) and the projector (output_projector.lua:4-6); in an LSTM graph, h2h of layer L is
     the Linear fed directly by the graph input prev_h[L] (LSTM.lua:47), i2h its partner in the same CAddTable
     (LSTM.lua:82-84); W_a / W_c are the (H,H) / (H,2H) LinearNoBias of the nested attention graph (LSTM.lua:130,155).
     The Linear modules carry no names unless -prealloc was given (memory.lua:55-61), hence structure, not names. ]]
local M = {}

M.GROUPS = { 'cnn', 'enc_fw', 'enc_bw', 'decoder', 'proj' }
M.BN_KEYS = { 'bn3', 'bn5', 'bn7' }

-- tensor order inside each flat group: module construction order, weight then bias (aocr/layout.py param_specs)
local ENC = { 'i2h.W', 'i2h.b', 'h2h.W', 'h2h.b' }
M.ORDER = {
  cnn = { 'conv1.W', 'conv1.b', 'conv2.W', 'conv2.b', 'conv3.W', 'conv3.b', 'bn3.gamma', 'bn3.beta', 'conv4.W', 'conv4.b',
          'conv5.W', 'conv5.b', 'bn5.gamma', 'bn5.beta', 'conv6.W', 'conv6.b', 'conv7.W', 'conv7.b', 'bn7.gamma', 'bn7.beta' },
  enc_fw = ENC, enc_bw = ENC,
  decoder = { 'emb', 'l1.i2h.W', 'l1.i2h.b', 'l1.h2h.W', 'l1.h2h.b', 'l2.i2h.W', 'l2.i2h.b', 'l2.h2h.W', 'l2.h2h.b',
              'attn.Wa', 'attn.Wc' },
  proj = { 'W', 'b' },
}

local CNN = { {'conv1', false}, {'conv2', false}, {'conv3', true}, {'conv4', false}, {'conv5', true}, {'conv6', false}, {'conv7', true} }

local function class_of(m) return torch.typename(m) or '' end
local function ends_with(s, suffix) return s:sub(-#suffix) == suffix end

-- name -> tensor of the CNN, and the running statistics of its three batch-norm layers
function M.cnn_tensors(cnn)
  local convs, bns = {}, {}
  for _, m in ipairs(cnn.modules) do
    local c = class_of(m)
    if ends_with(c, 'SpatialConvolution') or ends_with(c, 'SpatialConvolutionMM') then table.insert(convs, m) end
    if ends_with(c, 'SpatialBatchNormalization') then table.insert(bns, m) end
  end
  assert(#convs == 7 and #bns == 3, 'unexpected CNN structure (cnn.lua:12-42: 7 convolutions, 3 batch-norms)')
  local t, stats, bi = {}, {}, 0
  for i, spec in ipairs(CNN) do
    t[spec[1] .. '.W'] = convs[i].weight
    t[spec[1] .. '.b'] = convs[i].bias
    if spec[2] then
      bi = bi + 1
      local b = bns[bi]
      local key = 'bn' .. spec[1]:sub(-1)
      t[key .. '.gamma'] = b.weight
      t[key .. '.beta'] = b.bias
      local var = b.running_var
      if var == nil then      -- nn before 2016 kept running_std = 1 / sqrt(var + eps)
        local std = b.running_std:double()
        var = torch.cdiv(torch.DoubleTensor(std:size()):fill(1), torch.cmul(std, std)):add(-(b.eps or 1e-5))
      end
      stats[key] = { running_mean = b.running_mean, running_var = var }
    end
  end
  return t, stats
end

local function parents(data)
  local out = {}
  for i, p in ipairs(data.mapindex or {}) do out[i] = p end
  return out
end
local function is_linear(m) local c = class_of(m); return c == 'nn.Linear' or c == 'nn.LinearNoBias' end

function M.lstm_tensors(g, use_attention)
  local pairs_ = {}
  for _, node in ipairs(g.forwardnodes) do
    local d = node.data
    if d.module and class_of(d.module) == 'nn.CAddTable' then
      local ps = parents(d)
      if #ps == 2 and is_linear(ps[1].module) and is_linear(ps[2].module) then
        local feeds = {}
        for i = 1, 2 do
          local gp = parents(ps[i])
          feeds[i] = (#gp == 1) and gp[1].selectindex or nil
        end
        local h
        if feeds[1] and feeds[2] then h = (feeds[1] > feeds[2]) and 1 or 2
        elseif feeds[1] then h = 1 else h = 2 end
        assert(feeds[h], 'LSTM graph: no Linear of a gate sum is fed by a graph input (LSTM.lua:78-84)')
        table.insert(pairs_, { i2h = ps[3 - h].module, h2h = ps[h].module, idx = feeds[h] })
      end
    end
  end
  table.sort(pairs_, function(a, b) return a.idx < b.idx end)
  local t = {}
  if not use_attention then
    assert(#pairs_ == 1, 'encoder graph: expected one LSTM layer')
    t['i2h.W'], t['i2h.b'], t['h2h.W'], t['h2h.b'] = pairs_[1].i2h.weight, pairs_[1].i2h.bias, pairs_[1].h2h.weight, pairs_[1].h2h.bias
    return t
  end
  assert(#pairs_ == 2, 'decoder graph: expected two LSTM layers')
  for L = 1, 2 do
    local p = 'l' .. L .. '.'
    t[p .. 'i2h.W'], t[p .. 'i2h.b'], t[p .. 'h2h.W'], t[p .. 'h2h.b'] = pairs_[L].i2h.weight, pairs_[L].i2h.bias, pairs_[L].h2h.weight, pairs_[L].h2h.bias
  end
  for _, node in ipairs(g.forwardnodes) do
    local m = node.data.module
    if m and class_of(m) == 'nn.LookupTable' then t['emb'] = m.weight end
    if m and class_of(m) == 'nn.gModule' then
      for _, an in ipairs(m.forwardnodes) do
        local am = an.data.module
        if am and is_linear(am) then
          if am.weight:size(1) == am.weight:size(2) then t['attn.Wa'] = am.weight else t['attn.Wc'] = am.weight end
        end
      end
    end
  end
  assert(t['emb'] and t['attn.Wa'] and t['attn.Wc'], 'decoder graph: embedding / attention weights not found')
  return t
end

function M.proj_tensors(proj)
  for _, m in ipairs(proj.modules) do
    if class_of(m) == 'nn.Linear' then return { W = m.weight, b = m.bias } end
  end
  error('output projector: Linear not found (output_projector.lua:5)')
end

-- the five module trees -> { group = { name = tensor } }, { bnK = { running_mean, running_var } }
function M.all_tensors(mods)
  local cnn, stats = M.cnn_tensors(mods[1])
  return { cnn = cnn, enc_fw = M.lstm_tensors(mods[2], false), enc_bw = M.lstm_tensors(mods[3], false),
           decoder = M.lstm_tensors(mods[4], true), proj = M.proj_tensors(mods[5]) }, stats
end

-- { name = tensor } of one group -> one FloatTensor in the external layout (what aocr_set_params takes)
function M.flatten(group, ts)
  local n = 0
  for _, name in ipairs(M.ORDER[group]) do
    n = n + assert(ts[name], 'missing tensor ' .. group .. '.' .. name):nElement()
  end
  local flat, off = torch.FloatTensor(n), 1
  for _, name in ipairs(M.ORDER[group]) do
    local t = ts[name]:float():contiguous()
    local k = t:nElement()
    flat:narrow(1, off, k):copy(t:view(k))
    off = off + k
  end
  return flat
end

-- which of the three payloads a loaded checkpoint is: 'reference' | 'named' | 'flat'
function M.kind(ck)
  if type(ck) == 'table' and ck.format == 'aocr-params-v1' then return 'named' end
  if type(ck) == 'table' and type(ck[1]) == 'table' and #ck[1] == 5 and ck.params == nil then return 'reference' end
  if type(ck) == 'table' and type(ck.params) == 'table' and #ck.params == 5 then return 'flat' end
  error('not a torch-Attention-OCR checkpoint: expected {{5 modules}, config, global_step, optim_state} (model.lua:724), '
        .. 'a named-tensor table or the five flat groups')
end

-- any of the three -> { params = {5 x FloatTensor}, bn = {3 x {mean, var}}, config, global_step, optim_state }
function M.normalise(ck)
  local kind = M.kind(ck)
  if kind == 'flat' then return ck end
  local named, stats, out
  if kind == 'named' then
    named, stats = ck.params, ck.bn
    out = { config = ck.config or {}, global_step = ck.global_step or 0, optim_state = ck.optim_state or {} }
  else
    named, stats = M.all_tensors(ck[1])
    out = { config = ck[2] or {}, global_step = ck[3] or 0, optim_state = ck[4] or {} }
  end
  out.params, out.bn = {}, {}
  for i, g in ipairs(M.GROUPS) do out.params[i] = M.flatten(g, named[g]) end
  for l, k in ipairs(M.BN_KEYS) do
    out.bn[l] = { stats[k].running_mean:float():contiguous(), stats[k].running_var:float():contiguous() }
  end
  return out
end

return M
