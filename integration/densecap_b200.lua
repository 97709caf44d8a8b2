--[[
LuaJIT FFI shim for libdensecap_b200.so -- the reference-side binding a maintainer would add to jcjohnson/densecap to run the
test-time path on the B200 library (INTEGRATION.md).  UNTESTED HERE: this image has no LuaJIT / Torch7; the Python ctypes
binding in densecap_b200/_lib.py exercises exactly the same symbols and is what the test-suite runs.

Usage (replaces run_model.lua:146-154 and the forward_test call at :77):
  local B200 = require 'integration.densecap_b200'
  local model = B200.from_torch_model(checkpoint.model, {max_proposals = 1000, image_h = 600, image_w = 720})
  model:setTestArgs{rpn_nms_thresh = 0.7, final_nms_thresh = 0.3, num_proposals = 1000}
  local boxes, scores, captions = model:forward_test(img_caffe)          -- same return values as DenseCapModel:forward_test
--]]
local ffi = require 'ffi'

ffi.cdef[[
typedef struct dc_ctx dc_ctx;
typedef struct dc_config { int32_t widths[5]; int32_t rpn_filters, num_anchors, fc_dim, pool_hw, rnn_size,
  encoding_size, vocab_size, seq_length, pool_ceil, max_batch, max_image_h, max_image_w, max_proposals,
  reserved0, fuse_pool; float anchors[64]; float field_centers[4]; } dc_config;
int dc_config_default(dc_config*);
int dc_create(const dc_config*, int device, dc_ctx**);
int dc_destroy(dc_ctx*);
const char* dc_last_error(const dc_ctx*);
int dc_load_weights(dc_ctx*, const char* name, const float* host, const int64_t* shape, int ndim);
int dc_set_test_args(dc_ctx*, float rpn_nms_thresh, float final_nms_thresh, int num_proposals, int clip_boxes);
int dc_forward_test_host(dc_ctx*, const float* img_host, int batch, int H, int W,
                         int32_t* num_final, float* final_boxes, float* final_scores, int64_t* final_seq);
int dc_forward_test_host_u8(dc_ctx*, const uint8_t* img_host_u8, int batch, int H, int W,
                            int32_t* num_final, float* final_boxes, float* final_scores, int64_t* final_seq);   -- decoded RGB bytes, run_model.lua:66-74 on the device
int dc_set_beam_size(dc_ctx*, int beam_size);      -- 0 = LanguageModel:sample, n > 0 = :beamsearch (LanguageModel.lua:129-132)
int dc_set_decode_all(dc_ctx*, int enable);       -- captions for every proposal, not only the final-NMS survivors
]]

local C = ffi.load('densecap_b200')
local M = {}
M.__index = M

local function check(ctx, rc)
  if rc ~= 0 then error(ffi.string(C.dc_last_error(ctx)), 3) end     -- Lua error() = the reference's assert behaviour
end

local function load(self, name, tensor)
  local t = tensor:float():contiguous()
  local shape = ffi.new('int64_t[?]', t:dim())
  for i = 1, t:dim() do shape[i - 1] = t:size(i) end
  check(self.ctx, C.dc_load_weights(self.ctx, name, t:data(), shape, t:dim()))
end

-- collect the weighted modules of a container in `modules` order (nn.Sequential / ConcatTable)
local function weighted(m, out)
  out = out or {}
  if m.weight and m.weight:nElement() > 0 then table.insert(out, m) end
  if m.modules then for _, c in ipairs(m.modules) do weighted(c, out) end end
  return out
end

function M.from_torch_model(model, opt)
  opt = opt or {}
  local self = setmetatable({}, M)
  local cfg = ffi.new('dc_config[1]')
  check(nil, C.dc_config_default(cfg))
  local lm = model.nets.language_model
  cfg[0].vocab_size = lm.vocab_size
  cfg[0].seq_length = lm.seq_length
  cfg[0].max_proposals = opt.max_proposals or 1000
  cfg[0].max_image_h = opt.image_h or 600
  cfg[0].max_image_w = opt.image_w or 720
  local h = ffi.new('dc_ctx*[1]')
  check(nil, C.dc_create(cfg, opt.gpu or 0, h))
  self.ctx = ffi.gc(h[0], C.dc_destroy)
  self.P, self.T = cfg[0].max_proposals, lm.seq_length
  self.language_model = lm                                            -- decodeSequence stays in Lua (host string join)
  local names = {'conv1_1', 'conv1_2', 'conv2_1', 'conv2_2', 'conv3_1', 'conv3_2', 'conv3_3', 'conv4_1', 'conv4_2', 'conv4_3',
                 'conv5_1', 'conv5_2', 'conv5_3'}
  local convs = weighted(model.nets.conv_net1); weighted(model.nets.conv_net2, convs)
  local cin = 3
  for i, n in ipairs(names) do
    local m = convs[i]
    load(self, n .. '.weight', m.weight:view(m.nOutputPlane, cin, 3, 3)); load(self, n .. '.bias', m.bias)
    cin = m.nOutputPlane
  end
  local rpn = weighted(model.nets.localization_layer.nets.rpn)          -- 3x3 conv, box conv, score conv
  load(self, 'rpn.conv.weight', rpn[1].weight:view(rpn[1].nOutputPlane, cin, 3, 3)); load(self, 'rpn.conv.bias', rpn[1].bias)
  load(self, 'rpn.box.weight', rpn[2].weight:view(rpn[2].nOutputPlane, -1)); load(self, 'rpn.box.bias', rpn[2].bias)
  load(self, 'rpn.score.weight', rpn[3].weight:view(rpn[3].nOutputPlane, -1)); load(self, 'rpn.score.bias', rpn[3].bias)
  local fcs = weighted(model.nets.recog_base)
  load(self, 'fc6.weight', fcs[1].weight); load(self, 'fc6.bias', fcs[1].bias)
  load(self, 'fc7.weight', fcs[2].weight); load(self, 'fc7.bias', fcs[2].bias)
  load(self, 'obj.weight', model.nets.objectness_branch.weight); load(self, 'obj.bias', model.nets.objectness_branch.bias)
  load(self, 'boxreg.weight', model.nets.box_reg_branch.weight); load(self, 'boxreg.bias', model.nets.box_reg_branch.bias)
  local enc = weighted(lm.image_encoder)[1]
  load(self, 'lm.encoder.weight', enc.weight); load(self, 'lm.encoder.bias', enc.bias)
  load(self, 'lm.lookup.weight', lm.lookup_table.weight)
  for _, m in ipairs(weighted(lm.rnn)) do
    if torch.isTypeOf(m, nn.LSTM) then load(self, 'lm.lstm.weight', m.weight); load(self, 'lm.lstm.bias', m.bias)
    else load(self, 'lm.vocab.weight', m.weight); load(self, 'lm.vocab.bias', m.bias) end
  end
  return self
end

function M:setTestArgs(kwargs)                                           -- DenseCapModel.lua:185-191
  kwargs = kwargs or {}
  check(self.ctx, C.dc_set_test_args(self.ctx, kwargs.rpn_nms_thresh or 0.7, kwargs.final_nms_thresh or 0.3,
                                     kwargs.num_proposals or 1000, 1))
end

function M:evaluate() end

-- run_model.lua's `-beam_size` option (LanguageModel.lua:129-132 dispatches on self.beam_size)
function M:setBeamSize(n) check(self.ctx, C.dc_set_beam_size(self.ctx, n or 0)) end

function M:forward_test(img)                                             -- DenseCapModel.lua:319-327
  assert(img:dim() == 4 and img:size(1) == 1 and img:size(2) == 3, 'input must be 1 x 3 x H x W')
  local x = img:float():contiguous()
  local H, W = x:size(3), x:size(4)
  local n = ffi.new('int32_t[1]')
  local boxes, scores = torch.FloatTensor(self.P, 4), torch.FloatTensor(self.P)
  local seq = torch.LongTensor(self.P, self.T)
  check(self.ctx, C.dc_forward_test_host(self.ctx, x:data(), 1, H, W, n, boxes:data(), scores:data(), seq:data()))
  local k = n[0]
  if k == 0 then return torch.FloatTensor(0, 4), torch.FloatTensor(0, 1), {} end
  seq = seq:narrow(1, 1, k)
  return boxes:narrow(1, 1, k), scores:narrow(1, 1, k):view(k, 1), self.language_model:decodeSequence(seq)
end

return M
