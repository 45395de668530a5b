--[[ t7_convert.lua — moves weights between the reference's checkpoints and libaocr (run under Torch7, from the
     reference's source root so that its `require`s resolve):

       th t7_convert.lua export <reference_checkpoint.t7> <named.t7>
           the five module trees of model.lua:724  ->  the named-tensor table that aocr/checkpoint.py and
           lua/model.lua's model:load read (both also read the reference checkpoint directly): {format='aocr-params-v1', params={group={name=tensor}},
           bn={bn3={running_mean=,running_var=},...}, config, global_step, optim_state}
       th t7_convert.lua import <named.t7> <reference_checkpoint.t7>
           the reverse: builds the reference model with model:create(config) (src/model/model.lua:83-112) and copies
           every named tensor into the module that owns it, then model:save

     Modules are located by lua/aocr_ckpt.lua, the way aocr/checkpoint.py locates them (the two are kept in step). ]]
require 'torch'
require 'nn'
require 'nngraph'

package.path = ((arg and arg[0] or ''):match('^(.*)/[^/]*$') or '.') .. '/?.lua;' .. package.path   -- this script's directory
local K = require 'aocr_ckpt'      -- where the weights live in the module trees (shared with lua/model.lua)
local all_tensors = K.all_tensors

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
