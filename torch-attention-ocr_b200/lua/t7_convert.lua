--[[ t7_convert.lua — moves weights between the reference's checkpoints and libaocr (run under Torch7, from the
     reference's source root so that its `require`s resolve):

       th t7_convert.lua export <reference_checkpoint.t7> <named.t7>
           the five module trees of model.lua:724  ->  the named-tensor table that aocr/checkpoint.py (and
           lua/model.lua's model:load through the Python twin) read: {format='aocr-params-v1', params={group={name=tensor}},
           bn={bn3={running_mean=,running_var=},...}, config, global_step, optim_state}
       th t7_convert.lua import <named.t7> <reference_checkpoint.t7>
           the reverse: builds the reference model with model:create(config) (src/model/model.lua:83-112) and copies
           every named tensor into the module that owns it, then model:save

     Modules are located the way aocr/checkpoint.py locates them (the two are kept in step): nn.Sequential.modules in
     construction order for the CNN and the projector; in an LSTM graph, h2h of layer L is the Linear fed directly by the
     graph input prev_h[L] (LSTM.lua:47), i2h its partner in the same CAddTable (LSTM.lua:82-84); W_a / W_c are the
     (H,H) / (H,2H) LinearNoBias of the nested attention graph (LSTM.lua:130,155). ]]
require 'torch'
require 'nn'
require 'nngraph'

local CNN = { {'conv1', false}, {'conv2', false}, {'conv3', true}, {'conv4', false}, {'conv5', true}, {'conv6', false}, {'conv7', true} }

local function class_of(m) return torch.typename(m) or '' end
local function ends_with(s, suffix) return s:sub(-#suffix) == suffix end

-- name -> tensor accessors of the CNN (cnn.lua:9-45)
local function cnn_tensors(cnn)
  local convs, bns = {}, {}
  for _, m in ipairs(cnn.modules) do
    local c = class_of(m)
    if ends_with(c, 'SpatialConvolution') or ends_with(c, 'SpatialConvolutionMM') then table.insert(convs, m) end
    if ends_with(c, 'SpatialBatchNormalization') then table.insert(bns, m) end
  end
  assert(#convs == 7 and #bns == 3, 'unexpected CNN structure')
  local t, stats, bi = {}, {}, 0
  for i, spec in ipairs(CNN) do
    t[spec[1] .. '.W'] = convs[i].weight
    t[spec[1] .. '.b'] = convs[i].bias
    if spec[2] then
      bi = bi + 1
      local key = 'bn' .. spec[1]:sub(-1)
      t[key .. '.gamma'] = bns[bi].weight
      t[key .. '.beta'] = bns[bi].bias
      stats[key] = { running_mean = bns[bi].running_mean, running_var = bns[bi].running_var }
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

local function lstm_tensors(g, use_attention)
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
        assert(feeds[h], 'LSTM graph: no Linear of a gate sum is fed by a graph input')
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

local function proj_tensors(proj)
  for _, m in ipairs(proj.modules) do
    if class_of(m) == 'nn.Linear' then return { W = m.weight, b = m.bias } end
  end
  error('output projector: Linear not found')
end

local function all_tensors(mods)
  local cnn, stats = cnn_tensors(mods[1])
  return { cnn = cnn, enc_fw = lstm_tensors(mods[2], false), enc_bw = lstm_tensors(mods[3], false),
           decoder = lstm_tensors(mods[4], true), proj = proj_tensors(mods[5]) }, stats
end

local function export(src, dst)
  local ck = torch.load(src)
  local named, stats = all_tensors(ck[1])
  local out = { format = 'aocr-params-v1', params = {}, bn = {}, config = {}, global_step = ck[3] or 0, optim_state = ck[4] or {} }
  for g, ts in pairs(named) do
    out.params[g] = {}
    for n, t in pairs(ts) do out.params[g][n] = t:double():clone() end
  end
  for k, s in pairs(stats) do out.bn[k] = { running_mean = s.running_mean:double():clone(), running_var = s.running_var:double():clone() } end
  for k, v in pairs(ck[2] or {}) do
    if type(v) == 'number' or type(v) == 'string' or type(v) == 'boolean' then out.config[k] = v end
  end
  torch.save(dst, out)
  print('wrote ' .. dst)
end

local function import(src, dst)
  local ck = torch.load(src)
  assert(ck.format == 'aocr-params-v1', 'not a named-tensor checkpoint')
  require 'src/model/model'                                   -- the reference's Model class (run from its source root)
  local m = Model()
  m:create(ck.config)
  local named, stats = all_tensors({ m.cnn_model, m.encoder_fw, m.encoder_bw, m.decoder, m.output_projector })
  for g, ts in pairs(named) do
    for n, t in pairs(ts) do
      local s = assert(ck.params[g][n], 'missing tensor ' .. g .. '.' .. n)
      assert(s:nElement() == t:nElement(), 'size mismatch at ' .. g .. '.' .. n)
      t:copy(s:typeAs(t):viewAs(t))
    end
  end
  for k, s in pairs(stats) do
    s.running_mean:copy(ck.bn[k].running_mean:typeAs(s.running_mean))
    s.running_var:copy(ck.bn[k].running_var:typeAs(s.running_var))
  end
  m.global_step = ck.global_step or 0
  m.optim_state = ck.optim_state or m.optim_state
  m:save(dst)
  print('wrote ' .. dst)
end

local mode, src, dst = arg[1], arg[2], arg[3]
if mode == 'export' and src and dst then export(src, dst)
elseif mode == 'import' and src and dst then import(src, dst)
else print('usage: th t7_convert.lua export|import <in.t7> <out.t7>') end
