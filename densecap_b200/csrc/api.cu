// C ABI of libdensecap_b200.so: context, weights, and the orchestration of the DenseCap test-time
// forward pass (DenseCapModel.lua:242-282) over the kernels in the sibling .cu files.
#include <stdarg.h>

#include <algorithm>

#include "common.cuh"

// ------------------------------------------------------------------------------------------------
struct Weight {
  __half* planes = nullptr;     // split planes (hi then lo), K-major
  size_t plane_elems = 0;
  float inv_scale = 1.f;        // planes hold w * 2^s (keeps the lo plane out of fp16 subnormals); accumulators are multiplied by 2^-s
  float* f32 = nullptr;         // fp32 copy (biases, conv1_1, tiny heads)
  size_t n_f32 = 0;
  bool loaded = false;
};

struct SplitBuf {
  __half* base = nullptr;
  size_t cap_elems = 0;         // capacity per plane
  __half* hi() const { return base; }
  __half* lo(size_t n) const { return base + n; }   // planes are packed for the CURRENT tensor size n
};

struct PhaseTimer {
  std::vector<const char*> names;
  std::vector<cudaEvent_t> events;   // events[i] marks the START of phase i; last = end
  bool enabled = false;
  bool layers = false;               // dc_enable_timing(ctx, 2): one mark per conv layer of the trunk as well
  size_t used = 0;
  cudaEvent_t roi0 = nullptr, roi1 = nullptr;   // bracket the ROI-pooling kernel of image 0 (HBM roofline of that kernel)
  bool roi_valid = false;
};

struct dc_ctx {
  dc_config cfg;
  int device = 0;
  char err[512];
  int64_t launches = 0;
  LaunchState launch;                // per-context launcher state (tensor-map encoder, SM count, smem attributes)
  // test args (DenseCapModel.lua:185-191, LocalizationLayer.lua:233-238)
  float rpn_nms_thresh = 0.7f, final_nms_thresh = 0.3f;
  int num_proposals = 1000, clip_boxes = 1;
  // setImageSize protocol
  int image_h = 0, image_w = 0;
  bool size_consumed = true;
  std::map<std::string, Weight> w;
  std::vector<void*> allocs;
  // geometry capacity
  int fh_max = 0, fw_max = 0, A_max = 0, P = 0, B = 0;
  // buffers
  SplitBuf act[2], feats, rpn_hidden, roi, fc6o, fc7o, xh, enc, encx;
  float *rpn_head = nullptr, *codes = nullptr, *obj = nullptr, *trans = nullptr, *boxes_all = nullptr, *cstate = nullptr, *dec_gate_ws = nullptr;
  float *pp_boxes = nullptr, *pp_x1y1 = nullptr, *pp_prob = nullptr, *roi_boxes = nullptr, *fin_x1y1 = nullptr;
  int32_t *pp_src = nullptr, *pp_nvalid = nullptr, *pp_blocks = nullptr;
  int32_t *rpn_idx32 = nullptr, *rpn_count = nullptr, *fin_idx32 = nullptr, *fin_count = nullptr;
  // caption decoder (lm_decode.cu): dependency counters, row maps, number of rows decoded; decode_all = 1 decodes every proposal
  uint32_t* dec_counters = nullptr;
  int32_t *dec_rows = nullptr, *dec_dst_final = nullptr, *dec_dst_all = nullptr, *dec_count1 = nullptr, *dec_tokens = nullptr;
  float* lm_gate_table = nullptr;    // (V+2, 4H) fp32: b + lookup . Wx (gate-interleaved), rebuilt when an LM weight is loaded
  bool lm_gate_table_valid = false;
  int decode_all = 0;
  int beam_size = 0;                 // 0 = greedy LanguageModel:sample, > 0 = LanguageModel:beamsearch in the full forward
  struct BeamWs {                    // workspace of the beam search (allocated on first use)
    char* mem = nullptr;
    int rows_cap = 0;
    float *logits = nullptr, *c[2] = {nullptr, nullptr}, *cand_lp = nullptr, *beam_lp = nullptr;
    __half *h[2] = {nullptr, nullptr}, *x = nullptr;
    int32_t *cand_ix = nullptr, *beams[2] = {nullptr, nullptr}, *src_row = nullptr, *tokens = nullptr;
  } beam;
  int64_t *rpn_idx64 = nullptr, *fin_idx64 = nullptr, *seq64 = nullptr, *fin_seq = nullptr;
  float *fin_boxes = nullptr, *fin_scores = nullptr;
  unsigned long long* packed = nullptr;
  float* anchors_dev = nullptr;
  unsigned int* overflow_flag = nullptr;   // device word, see GemmEpilogue::overflow and dc_check_range
  float* splitk_ws = nullptr;        // partial-sum planes of the split-K convolutions (conv5_x / RPN 3x3 on small feature maps)
  size_t splitk_elems = 0;
  float* scratch_f32 = nullptr;      // generic fp32 scratch (HWC features for dc_bilinear_roi_pool, ...)
  size_t scratch_elems = 0;
  void* roi_taps = nullptr;          // per-(roi,position) bilinear tap records, one block per batch image
  size_t roi_taps_stride = 0;        // bytes between the per-image blocks
  int roi_taps_boxes = 0;
  std::vector<NmsWorkspace> nms;     // one NMS workspace per batch image (images run on forked streams)
  // per-image side streams: the latency-bound per-image stages (RPN post-processing, NMS, ROI pooling, final NMS)
  // of different images of a batch overlap instead of queueing behind each other
  std::vector<cudaStream_t> side;
  std::vector<cudaEvent_t> side_done;
  // device staging of the host-boundary entry points (dc_forward_test_host / _u8), grown on demand, freed with the context
  float* host_stage_f32 = nullptr;
  uint8_t* host_stage_u8 = nullptr;
  size_t host_stage_f32_elems = 0, host_stage_u8_elems = 0;
  float* resize_tmp = nullptr;       // (3, H, out_w) intermediate of the two-pass image.scale, grown on demand
  size_t resize_tmp_elems = 0;
  cudaStream_t fin_stream0 = nullptr;   // final NMS of image 0 (runs beside the caption decode, see run_final_nms)
  cudaEvent_t fin_done0 = nullptr;
  cudaEvent_t fork_ev = nullptr;
  PhaseTimer timer;
  // training-time forward of the LocalizationLayer (setGroundTruth protocol, sampler / loss options)
  float mid_box_reg_weight = 0.05f, mid_objectness_weight = 0.1f, box_reg_decay = 0.f;   // LocalizationLayer.lua:48-54
  float sampler_low = 0.5f, sampler_high = 0.7f;                                         // :79-82
  int sampler_batch = 256, remove_outbounds = 1;
  int sampler_mode = 1;              // 1 = uniform random minibatch like torch.multinomial (BoxSampler.lua:138-152), 0 = first P / first M in index order
  uint64_t sampler_seed = 0, sampler_calls = 0;
  void* train = nullptr;             // TrainBuffers (train_kernels.cu), allocated on the first training call
  char* train_mem = nullptr;
  int train_A = 0, train_B2max = 0, train_batch = 0;
  int gt_count = 0;
  bool gt_consumed = true;
  float* train_losses = nullptr;
  // backward of the training-time forward (LocalizationLayer.lua:529-605)
  bool gt_bwd_consumed = true, size_bwd_consumed = true, train_fwd_valid = false;
  int train_fh = 0, train_fw = 0;
  char* bwd_mem = nullptr;           // workspace, allocated on the first backward call
  struct Bwd {
    float *go_bhwc, *gfeat_roi, *ggrid, *gboxes_pool, *g_boxes_all, *g_trans_all, *g_scores_all, *head_grad, *gh, *hid_f32, *feats_f32;
    __half *head_grad_planes, *gh_planes, *grpn_planes;
    float *dW_conv, *db_conv, *dW_head, *db_head;      // gradWeight / gradBias accumulators (accGradParameters)
  } bwd{};
  int bwd_S = 0;
};

static thread_local char g_err[512] = "";   // message of the last ctx-less failure (dc_create, leaf operators), per thread

int dc_set_error(dc_ctx* ctx, int code, const char* fmt, ...) {
  char* dst = ctx ? ctx->err : g_err;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
  if (ctx) strncpy(g_err, ctx->err, sizeof(g_err) - 1);
  return code;
}
LaunchState* dc_launch_state(dc_ctx* ctx) { return &ctx->launch; }
float* dc_splitk_workspace(dc_ctx* ctx, size_t* elems) { *elems = ctx->splitk_elems; return ctx->splitk_ws; }
unsigned int* dc_overflow_flag(dc_ctx* ctx) { return ctx ? ctx->overflow_flag : nullptr; }
void dc_count_launch(dc_ctx* ctx, int n) {
  if (ctx) ctx->launches += n;
}

extern "C" const char* dc_last_error(const dc_ctx* ctx) { return ctx ? ctx->err : g_err; }
extern "C" int dc_version(void) { return 100; }
extern "C" int dc_debug_peek_cuda_error(char* msg, int capacity) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaGetLastError(); else cudaGetLastError();
  if (msg && capacity > 0) snprintf(msg, (size_t)capacity, "%s", e == cudaSuccess ? "" : cudaGetErrorString(e));
  return e == cudaSuccess ? 0 : (int)e;
}
extern "C" int64_t dc_kernel_launches(const dc_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int dc_config_default(dc_config* c) {
  if (!c) return dc_set_error(nullptr, DC_E_INVALID, "dc_config_default: null");
  memset(c, 0, sizeof(*c));
  const int widths[5] = {64, 128, 256, 512, 512};
  memcpy(c->widths, widths, sizeof(widths));
  c->rpn_filters = 256;
  c->num_anchors = 12;
  c->fc_dim = 4096;
  c->pool_hw = 7;
  c->rnn_size = 512;
  c->encoding_size = 512;
  c->vocab_size = 10497;
  c->seq_length = 15;
  c->pool_ceil = 1;
  c->max_batch = 1;
  c->max_image_h = 600;
  c->max_image_w = 720;
  c->max_proposals = 1000;
  c->reserved0 = 0;
  c->fuse_pool = 1;
  // LocalizationLayer.lua:613-618 (stored (2,k): widths then heights)
  const float aw[12] = {45, 90, 64, 90, 180, 128, 181, 362, 256, 362, 724, 512};
  const float ah[12] = {90, 45, 64, 180, 90, 128, 362, 181, 256, 724, 362, 512};
  for (int i = 0; i < 12; ++i) { c->anchors[i] = aw[i]; c->anchors[12 + i] = ah[i]; }
  c->field_centers[0] = 8.5f; c->field_centers[1] = 8.5f; c->field_centers[2] = 16.f; c->field_centers[3] = 16.f;
  return DC_OK;
}

static int pooled(int v, int ceil_mode) { return ceil_mode ? (v + 1) / 2 : v / 2; }
static void feat_hw(const dc_config& c, int h, int w, int& fh, int& fw) {
  fh = h; fw = w;
  for (int i = 0; i < 4; ++i) { fh = pooled(fh, c.pool_ceil); fw = pooled(fw, c.pool_ceil); }
}
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static const size_t kHeadGradLd = 128;     // 6k head-gradient columns padded to a multiple of the GEMM K block

template <typename T>
static int dev_alloc(dc_ctx* ctx, T** p, size_t n) {
  void* q = nullptr;
  const size_t bytes = align_up((n ? n : 1) * sizeof(T), 256);
  cudaError_t e = cudaMalloc(&q, bytes);
  if (e != cudaSuccess) return dc_set_error(ctx, DC_E_NOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
  cudaMemset(q, 0, bytes);
  ctx->allocs.push_back(q);
  *p = (T*)q;
  return DC_OK;
}
static int split_alloc(dc_ctx* ctx, SplitBuf& b, size_t elems) {
  b.cap_elems = align_up(elems, 64);
  return dev_alloc(ctx, &b.base, 2 * b.cap_elems);
}

static const char* kConvNames[13] = {"conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3",
                                     "conv4_1", "conv4_2", "conv4_3", "conv5_1", "conv5_2", "conv5_3"};
static const int kConvStage[13] = {0, 0, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4};
static const int kPoolAfter[13] = {0, 1, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 0};

static int weight_alloc_split(dc_ctx* ctx, const std::string& name, size_t elems) {
  Weight& w = ctx->w[name];
  w.plane_elems = align_up(elems, 64);
  return dev_alloc(ctx, &w.planes, 2 * w.plane_elems);
}
static int weight_alloc_f32(dc_ctx* ctx, const std::string& name, size_t n) {
  Weight& w = ctx->w[name];
  w.n_f32 = n;
  return dev_alloc(ctx, &w.f32, n);
}

extern "C" int dc_create(const dc_config* cfg, int device, dc_ctx** out) {
  if (!cfg || !out) return dc_set_error(nullptr, DC_E_INVALID, "dc_create: null argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return dc_set_error(nullptr, DC_E_CUDA, "dc_create: no CUDA device available (%s); this library has no CPU fallback",
                        cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return dc_set_error(nullptr, DC_E_INVALID, "dc_create: device %d out of range (%d devices)", device, ndev);
  const dc_config& c = *cfg;
  if (c.num_anchors <= 0 || c.num_anchors > DC_MAX_ANCHORS || c.max_batch <= 0 || c.max_proposals <= 0 || c.pool_hw < 2 ||
      c.pool_hw * c.pool_hw > 64 || c.seq_length <= 0 || c.vocab_size <= 0)
    return dc_set_error(nullptr, DC_E_INVALID, "dc_create: invalid configuration");
  if (c.reserved0 != 0) return dc_set_error(nullptr, DC_E_UNSUPPORTED, "dc_create: reserved0 must be 0 (there is one dense engine, tcgen05)");
  for (int i = 0; i < 5; ++i)
    if (c.widths[i] <= 0 || c.widths[i] % 64 != 0) return dc_set_error(nullptr, DC_E_UNSUPPORTED, "dc_create: stage widths must be multiples of 64");
  if (c.rpn_filters % 64 || c.fc_dim % 64 || c.rnn_size % 64 || c.encoding_size % 64)
    return dc_set_error(nullptr, DC_E_UNSUPPORTED, "dc_create: rpn_filters/fc_dim/rnn_size/encoding_size must be multiples of 64");
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return dc_set_error(nullptr, DC_E_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  dc_ctx* ctx = new dc_ctx();
  ctx->cfg = c;
  ctx->device = device;
  ctx->err[0] = 0;
  ctx->num_proposals = c.max_proposals;
  int rc = DC_OK;
#define TRY(x) do { rc = (x); if (rc != DC_OK) { strncpy(g_err, ctx->err, sizeof(g_err) - 1); dc_destroy(ctx); return rc; } } while (0)
  TRY(tc_init(ctx));
  ctx->B = c.max_batch;
  ctx->P = c.max_proposals;
  feat_hw(c, c.max_image_h, c.max_image_w, ctx->fh_max, ctx->fw_max);
  // the feature map of a smaller image with the other aspect ratio must fit too: bound by cells
  ctx->A_max = c.num_anchors * ctx->fh_max * ctx->fw_max;
  const size_t B = ctx->B, P = ctx->P;
  const size_t C5 = c.widths[4], F = c.rpn_filters, fc = c.fc_dim, E = c.encoding_size, H = c.rnn_size, T = c.seq_length;
  const size_t hw = (size_t)c.pool_hw * c.pool_hw;
  const size_t cells = (size_t)ctx->fh_max * ctx->fw_max;
  // largest activation: 64-channel full-resolution map (or wider stage at lower resolution)
  size_t act_elems = 0;
  {
    int h = c.max_image_h, w = c.max_image_w;
    for (int s = 0; s < 5; ++s) {
      act_elems = std::max(act_elems, B * (size_t)h * w * c.widths[s]);
      h = pooled(h, c.pool_ceil); w = pooled(w, c.pool_ceil);
    }
  }
  TRY(split_alloc(ctx, ctx->act[0], act_elems));
  TRY(split_alloc(ctx, ctx->act[1], act_elems));
  TRY(split_alloc(ctx, ctx->feats, B * cells * C5));
  TRY(split_alloc(ctx, ctx->rpn_hidden, B * cells * F));
  TRY(dev_alloc(ctx, &ctx->rpn_head, B * cells * 6 * c.num_anchors));
  TRY(split_alloc(ctx, ctx->roi, B * P * hw * C5));
  TRY(split_alloc(ctx, ctx->fc6o, B * P * fc));
  TRY(split_alloc(ctx, ctx->fc7o, B * P * fc));
  TRY(split_alloc(ctx, ctx->xh, 2 * B * P * H));            // two ping-pong h_t operands of the decoder (lm_decode.cu)
  TRY(split_alloc(ctx, ctx->enc, B * P * E));
  TRY(split_alloc(ctx, ctx->encx, B * P * E));
  TRY(dev_alloc(ctx, &ctx->dec_tokens, B * P));
  TRY(dev_alloc(ctx, &ctx->lm_gate_table, (size_t)(c.vocab_size + 2) * 4 * H));
  TRY(dev_alloc(ctx, &ctx->dec_counters, lm_decode_counter_words((int)T, (int)(B * P))));
  TRY(dev_alloc(ctx, &ctx->dec_rows, 1));
  TRY(dev_alloc(ctx, &ctx->dec_count1, 1));
  TRY(dev_alloc(ctx, &ctx->dec_dst_final, B * P));
  TRY(dev_alloc(ctx, &ctx->dec_dst_all, B * P));
  TRY(dev_alloc(ctx, &ctx->codes, B * P * fc));
  TRY(dev_alloc(ctx, &ctx->obj, B * P));
  TRY(dev_alloc(ctx, &ctx->trans, B * P * 4));
  TRY(dev_alloc(ctx, &ctx->boxes_all, B * P * 4));
  TRY(dev_alloc(ctx, &ctx->cstate, B * P * H));
  TRY(dev_alloc(ctx, &ctx->dec_gate_ws, B * P * 4 * H));
  const size_t A = ctx->A_max;
  TRY(dev_alloc(ctx, &ctx->pp_boxes, B * A * 4));
  TRY(dev_alloc(ctx, &ctx->pp_x1y1, B * A * 4));
  TRY(dev_alloc(ctx, &ctx->pp_prob, B * A));
  TRY(dev_alloc(ctx, &ctx->pp_src, B * A));
  TRY(dev_alloc(ctx, &ctx->pp_nvalid, B));
  TRY(dev_alloc(ctx, &ctx->pp_blocks, B * (A / 256 + 2)));
  TRY(dev_alloc(ctx, &ctx->roi_boxes, B * P * 4));
  TRY(dev_alloc(ctx, &ctx->fin_x1y1, B * P * 4));
  TRY(dev_alloc(ctx, &ctx->rpn_idx32, B * P));
  TRY(dev_alloc(ctx, &ctx->rpn_idx64, B * P));
  TRY(dev_alloc(ctx, &ctx->rpn_count, B));
  TRY(dev_alloc(ctx, &ctx->fin_idx32, B * P));
  TRY(dev_alloc(ctx, &ctx->fin_idx64, B * P));
  TRY(dev_alloc(ctx, &ctx->fin_count, B));
  TRY(dev_alloc(ctx, &ctx->seq64, B * P * T));
  TRY(dev_alloc(ctx, &ctx->fin_seq, B * P * T));
  TRY(dev_alloc(ctx, &ctx->fin_boxes, B * P * 4));
  TRY(dev_alloc(ctx, &ctx->fin_scores, B * P));
  TRY(dev_alloc(ctx, &ctx->packed, B * P));
  TRY(dev_alloc(ctx, &ctx->anchors_dev, (size_t)2 * c.num_anchors));
  DC_CHECK_CUDA(ctx, cudaMemcpy(ctx->anchors_dev, c.anchors, sizeof(float) * 2 * c.num_anchors, cudaMemcpyHostToDevice));
  TRY(dev_alloc(ctx, &ctx->overflow_flag, 1));
  ctx->splitk_elems = 8 * cells * std::max(C5, F);                 // up to 8 K-splits of a conv5-sized map (batch 1-2 only, see k_tc_conv3x3)
  ctx->splitk_elems = std::max<size_t>(ctx->splitk_elems, (size_t)(148 / 2) * 256 * 256);   // K-range partials of the GEMM tail tiles (k_tc_gemm): <= one item per SM pair
  TRY(dev_alloc(ctx, &ctx->splitk_ws, ctx->splitk_elems));
  ctx->scratch_elems = std::max(cells * C5, (size_t)1 << 20);
  TRY(dev_alloc(ctx, &ctx->scratch_f32, ctx->scratch_elems));
  const int nms_cap = (int)align_up(std::max(A, P), 64);
  ctx->nms.resize(B);
  for (size_t b = 0; b < B; ++b) {
    char* mem = nullptr;
    TRY(dev_alloc(ctx, &mem, nms_workspace_bytes(nms_cap)));
    nms_workspace_carve(ctx->nms[b], mem, nms_cap);
  }
  {
    char* mem = nullptr;
    ctx->roi_taps_boxes = (int)std::max((size_t)P, (size_t)1024);
    ctx->roi_taps_stride = align_up(roi_taps_bytes(ctx->roi_taps_boxes, c.pool_hw * c.pool_hw), 256);
    TRY(dev_alloc(ctx, &mem, ctx->roi_taps_stride * B));
    ctx->roi_taps = mem;
  }
  DC_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
  DC_CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->fin_stream0, cudaStreamNonBlocking));
  DC_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&ctx->fin_done0, cudaEventDisableTiming));
  for (size_t b = 1; b < B; ++b) {
    cudaStream_t sst;
    cudaEvent_t ev;
    DC_CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&sst, cudaStreamNonBlocking));
    DC_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    ctx->side.push_back(sst);
    ctx->side_done.push_back(ev);
  }
  // ---- weights ----
  {
    int cin = 3;
    for (int i = 0; i < 13; ++i) {
      const int cout = c.widths[kConvStage[i]];
      const std::string n = kConvNames[i];
      if (i == 0) {
        TRY(weight_alloc_f32(ctx, n + ".weight", (size_t)cout * 27));
        if (cout == 64) TRY(weight_alloc_split(ctx, n + ".weight.tc", (size_t)64 * 64));      // tcgen05 filter bank, K padded 27 -> 64
      } else TRY(weight_alloc_split(ctx, n + ".weight", (size_t)cout * 9 * cin));
      TRY(weight_alloc_f32(ctx, n + ".bias", cout));
      cin = cout;
    }
    const int k = c.num_anchors;
    TRY(weight_alloc_split(ctx, "rpn.conv.weight", F * 9 * C5));
    TRY(weight_alloc_f32(ctx, "rpn.conv.bias", F));
    TRY(weight_alloc_split(ctx, "rpn.head.weight", (size_t)6 * k * F));   // rows [0,4k) box, [4k,6k) score
    TRY(weight_alloc_f32(ctx, "rpn.head.bias", (size_t)6 * k));
    // layouts for the backward-data passes of the training path: transposed + tap-flipped 3x3 bank, transposed head (K padded to 128)
    TRY(weight_alloc_split(ctx, "rpn.conv.weight.bwd", C5 * 9 * F));
    TRY(weight_alloc_split(ctx, "rpn.head.weight.bwd", F * kHeadGradLd));
    TRY(weight_alloc_f32(ctx, "rpn.head.wT", F * kHeadGradLd));
    TRY(weight_alloc_split(ctx, "fc6.weight", fc * hw * C5));
    TRY(weight_alloc_f32(ctx, "fc6.bias", fc));
    TRY(weight_alloc_split(ctx, "fc7.weight", fc * fc));
    TRY(weight_alloc_f32(ctx, "fc7.bias", fc));
    TRY(weight_alloc_f32(ctx, "obj.weight", fc));
    TRY(weight_alloc_f32(ctx, "obj.bias", 1));
    TRY(weight_alloc_f32(ctx, "boxreg.weight", 4 * fc));
    TRY(weight_alloc_f32(ctx, "boxreg.bias", 4));
    const size_t V = c.vocab_size;
    TRY(weight_alloc_split(ctx, "lm.encoder.weight", E * fc));
    TRY(weight_alloc_f32(ctx, "lm.encoder.bias", E));
    TRY(weight_alloc_split(ctx, "lm.lookup.weight", (V + 2) * E));
    TRY(weight_alloc_split(ctx, "lm.lstm.weight", 4 * H * (E + H)));
    TRY(weight_alloc_f32(ctx, "lm.lstm.bias", 4 * H));
    TRY(weight_alloc_split(ctx, "lm.vocab.weight", (V + 1) * H));
    TRY(weight_alloc_f32(ctx, "lm.vocab.bias", V + 1));
  }
#undef TRY
  *out = ctx;
  return DC_OK;
}

extern "C" int dc_destroy(dc_ctx* ctx) {
  if (!ctx) return DC_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (void* p : ctx->allocs) cudaFree(p);
  for (cudaStream_t sst : ctx->side) cudaStreamDestroy(sst);
  for (cudaEvent_t ev : ctx->side_done) cudaEventDestroy(ev);
  if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
  if (ctx->fin_stream0) cudaStreamDestroy(ctx->fin_stream0);
  if (ctx->fin_done0) cudaEventDestroy(ctx->fin_done0);
  if (ctx->host_stage_f32) cudaFree(ctx->host_stage_f32);
  if (ctx->host_stage_u8) cudaFree(ctx->host_stage_u8);
  if (ctx->resize_tmp) cudaFree(ctx->resize_tmp);
  if (ctx->beam.mem) cudaFree(ctx->beam.mem);
  if (ctx->train_mem) cudaFree(ctx->train_mem);
  if (ctx->bwd_mem) cudaFree(ctx->bwd_mem);
  free(ctx->train);
  for (cudaEvent_t ev : ctx->timer.events) cudaEventDestroy(ev);
  if (ctx->timer.roi0) { cudaEventDestroy(ctx->timer.roi0); cudaEventDestroy(ctx->timer.roi1); }
  delete ctx;
  return DC_OK;
}

// ------------------------------------------------------------------------------------------------
static size_t shape_numel(const int64_t* shape, int ndim) {
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= (size_t)shape[i];
  return n;
}
static bool shape_is(const int64_t* s, int nd, std::initializer_list<int64_t> want) {
  if ((size_t)nd != want.size()) return false;
  int i = 0;
  for (int64_t v : want) if (s[i++] != v) return false;
  return true;
}

// Power-of-two pre-scale for a weight tensor.  NN weights are typically |w| < 0.25, for which the fp16 "lo" plane
// (~ w * 2^-12) would be SUBNORMAL and the split would keep only ~2^-17 relative precision.  Scaling the tensor so
// that max|w| lands in [2^14, 2^15) keeps hi and lo normal for every element within 2^-17 of the maximum; the exact
// inverse power of two is applied to the fp32 accumulator in the GEMM epilogue.
static int weight_scale(dc_ctx* ctx, const float* raw_dev, size_t n, float* scale_out, cudaStream_t st) {
  float* d = nullptr;
  DC_CHECK_CUDA(ctx, cudaMalloc((void**)&d, sizeof(float)));
  cudaMemsetAsync(d, 0, sizeof(float), st);
  int rc = k_absmax(ctx, raw_dev, n, d, st);
  float m = 0.f;
  cudaError_t e = cudaMemcpyAsync(&m, d, sizeof(float), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (rc) return rc;
  if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "weight_scale: %s", cudaGetErrorString(e));
  float sc = 1.f;
  if (m > 0.f && m < 3.0e38f) {
    int ex = 0;
    frexpf(m, &ex);                 // m = f * 2^ex, f in [0.5, 1)
    int shift = 15 - ex;            // max|w| * 2^shift in [2^14, 2^15)
    if (shift > 40) shift = 40;
    if (shift < -40) shift = -40;
    sc = ldexpf(1.f, shift);
  }
  *scale_out = sc;
  return DC_OK;
}

extern "C" int dc_load_weights(dc_ctx* ctx, const char* name, const float* data_host, const int64_t* shape, int ndim) {
  if (!ctx || !name || !data_host || !shape) return dc_set_error(ctx, DC_E_INVALID, "dc_load_weights: null argument");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  const dc_config& c = ctx->cfg;
  const std::string n = name;
  const size_t numel = shape_numel(shape, ndim);
  const int64_t C5 = c.widths[4], F = c.rpn_filters, fc = c.fc_dim, E = c.encoding_size, H = c.rnn_size, V = c.vocab_size;
  const int64_t k = c.num_anchors, hw = (int64_t)c.pool_hw * c.pool_hw;
  cudaStream_t st = 0;
  // staging copy of the raw fp32 tensor
  float* raw = nullptr;
  auto stage = [&]() -> int {
    cudaError_t e = cudaMalloc((void**)&raw, numel * sizeof(float));
    if (e != cudaSuccess) return dc_set_error(ctx, DC_E_NOMEM, "dc_load_weights(%s): staging cudaMalloc failed: %s", name, cudaGetErrorString(e));
    e = cudaMemcpy(raw, data_host, numel * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(raw); raw = nullptr; return dc_set_error(ctx, DC_E_CUDA, "dc_load_weights(%s): H2D failed: %s", name, cudaGetErrorString(e)); }
    return DC_OK;
  };
  auto finish = [&](int rc) -> int {
    cudaError_t e = cudaStreamSynchronize(st);
    if (raw) cudaFree(raw);
    if (rc != DC_OK) return rc;
    if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "dc_load_weights(%s): %s", name, cudaGetErrorString(e));
    return DC_OK;
  };
  auto bad_shape = [&]() { return dc_set_error(ctx, DC_E_INVALID, "dc_load_weights(%s): unexpected shape", name); };
  auto copy_f32 = [&](const std::string& key, size_t offset, size_t count) -> int {
    Weight& w = ctx->w[key];
    if (offset + count > w.n_f32) return bad_shape();
    DC_CHECK_CUDA(ctx, cudaMemcpy(w.f32 + offset, data_host, count * sizeof(float), cudaMemcpyHostToDevice));
    w.loaded = true;
    return DC_OK;
  };

  // ---- VGG convs ----
  {
    int cin = 3;
    for (int i = 0; i < 13; ++i) {
      const int cout = c.widths[kConvStage[i]];
      const std::string base = kConvNames[i];
      if (n == base + ".weight") {
        if (!shape_is(shape, ndim, {cout, cin, 3, 3})) return bad_shape();
        if (i == 0) {
          int rc = copy_f32(n, 0, numel);
          if (rc || cout != 64) return rc;
          Weight& wt = ctx->w[n + ".tc"];
          Weight& wf = ctx->w[n];
          float sc = 1.f;
          if ((rc = weight_scale(ctx, wf.f32, numel, &sc, st))) return rc;
          rc = k_conv_first_weight_to_split(ctx, wf.f32, wt.planes, wt.planes + wt.plane_elems, sc, st);
          wt.inv_scale = 1.f / sc;
          wt.loaded = true;
          return finish(rc);
        }
        int rc = stage();
        if (rc) return rc;
        Weight& w = ctx->w[n];
        float sc = 1.f;
        if ((rc = weight_scale(ctx, raw, numel, &sc, st))) return finish(rc);
        rc = k_conv_weight_to_split(ctx, raw, w.planes, w.planes + w.plane_elems, cout, cin, st, sc);
        w.inv_scale = 1.f / sc;
        w.loaded = true;
        return finish(rc);
      }
      if (n == base + ".bias") {
        if (!shape_is(shape, ndim, {cout})) return bad_shape();
        return copy_f32(n, 0, numel);
      }
      cin = cout;
    }
  }
  if (n == "rpn.conv.weight") {
    if (!shape_is(shape, ndim, {F, C5, 3, 3})) return bad_shape();
    int rc = stage();
    if (rc) return rc;
    Weight& w = ctx->w[n];
    float sc = 1.f;
    if ((rc = weight_scale(ctx, raw, numel, &sc, st))) return finish(rc);
    rc = k_conv_weight_to_split(ctx, raw, w.planes, w.planes + w.plane_elems, (int)F, (int)C5, st, sc);
    w.inv_scale = 1.f / sc;
    w.loaded = true;
    if (rc == DC_OK) {
      // backward-data bank: W'[c][f][r'][s'] = W[f][c][2-r'][2-s']
      float* tmp = nullptr;
      if (cudaMalloc((void**)&tmp, numel * sizeof(float)) != cudaSuccess) return finish(dc_set_error(ctx, DC_E_NOMEM, "rpn.conv.weight: staging"));
      Weight& wb = ctx->w["rpn.conv.weight.bwd"];
      rc = k_conv_weight_flip_transpose(ctx, raw, tmp, (int)F, (int)C5, st);
      if (rc == DC_OK) rc = k_conv_weight_to_split(ctx, tmp, wb.planes, wb.planes + wb.plane_elems, (int)C5, (int)F, st, sc);
      wb.inv_scale = 1.f / sc;
      wb.loaded = true;
      cudaStreamSynchronize(st);
      cudaFree(tmp);
    }
    return finish(rc);
  }
  if (n == "rpn.conv.bias") { if (!shape_is(shape, ndim, {F})) return bad_shape(); return copy_f32(n, 0, numel); }
  if (n == "rpn.box.weight" || n == "rpn.score.weight") {
    const bool box = (n == "rpn.box.weight");
    const int64_t rows = box ? 4 * k : 2 * k;
    if (!(shape_is(shape, ndim, {rows, F, 1, 1}) || shape_is(shape, ndim, {rows, F}))) return bad_shape();
    int rc = stage();
    if (rc) return rc;
    Weight& w = ctx->w["rpn.head.weight"];
    const size_t off = box ? 0 : (size_t)4 * k * F;
    // box and score rows share one GEMM, hence one (fixed) pre-scale: 2^12 suits the O(0.01..1) RPN head weights
    const float sc = 4096.f;
    rc = k_rowmajor_to_split(ctx, raw, w.planes + off, w.planes + w.plane_elems + off, numel, st, sc);
    w.inv_scale = 1.f / sc;
    w.loaded = true;
    if (rc == DC_OK) {
      Weight& wt = ctx->w["rpn.head.wT"];
      Weight& wb = ctx->w["rpn.head.weight.bwd"];
      rc = k_transpose_into(ctx, raw, (int)rows, (int)F, wt.f32, (int)kHeadGradLd, box ? 0 : 4 * (int)k, st);
      if (rc == DC_OK) rc = k_rowmajor_to_split(ctx, wt.f32, wb.planes, wb.planes + wb.plane_elems, (size_t)F * kHeadGradLd, st, sc);
      wb.inv_scale = 1.f / sc;
      wb.loaded = true;
    }
    return finish(rc);
  }
  if (n == "rpn.box.bias") { if (!shape_is(shape, ndim, {4 * k})) return bad_shape(); return copy_f32("rpn.head.bias", 0, numel); }
  if (n == "rpn.score.bias") { if (!shape_is(shape, ndim, {2 * k})) return bad_shape(); return copy_f32("rpn.head.bias", (size_t)4 * k, numel); }
  if (n == "fc6.weight") {
    if (!shape_is(shape, ndim, {fc, C5 * hw})) return bad_shape();
    int rc = stage();
    if (rc) return rc;
    Weight& w = ctx->w[n];
    float sc = 1.f;
    if ((rc = weight_scale(ctx, raw, numel, &sc, st))) return finish(rc);
    rc = k_fc6_weight_to_split(ctx, raw, w.planes, w.planes + w.plane_elems, (int)fc, (int)C5, (int)hw, st, sc);
    w.inv_scale = 1.f / sc;
    w.loaded = true;
    return finish(rc);
  }
  struct RowMajor { const char* name; int64_t rows, cols; };
  const RowMajor rm[] = {{"fc7.weight", fc, fc}, {"lm.encoder.weight", E, fc}, {"lm.lookup.weight", V + 2, E}, {"lm.vocab.weight", V + 1, H}};
  for (const RowMajor& r : rm) {
    if (n == r.name) {
      if (!shape_is(shape, ndim, {r.rows, r.cols})) return bad_shape();
      int rc = stage();
      if (rc) return rc;
      Weight& w = ctx->w[n];
      float sc = 1.f;
      // the lookup table is an ACTIVATION source (rows are copied into the x operand), so it stays unscaled
      if (n != "lm.lookup.weight" && (rc = weight_scale(ctx, raw, numel, &sc, st))) return finish(rc);
      rc = k_rowmajor_to_split(ctx, raw, w.planes, w.planes + w.plane_elems, numel, st, sc);
      w.inv_scale = 1.f / sc;
      w.loaded = true;
      if (n == "lm.lookup.weight") ctx->lm_gate_table_valid = false;
      return finish(rc);
    }
  }
  if (n == "lm.lstm.weight") {
    if (!shape_is(shape, ndim, {E + H, 4 * H})) return bad_shape();
    int rc = stage();
    if (rc) return rc;
    Weight& w = ctx->w[n];
    // rows of the B operand are gate-interleaved (n = 4*j + gate) so the LSTM cell fuses into the GEMM epilogue
    float sc = 1.f;
    if ((rc = weight_scale(ctx, raw, numel, &sc, st))) return finish(rc);
    rc = k_transpose_to_split(ctx, raw, w.planes, w.planes + w.plane_elems, (int)(E + H), (int)(4 * H), (int)H, st, sc);
    w.inv_scale = 1.f / sc;
    w.loaded = true;
    ctx->lm_gate_table_valid = false;
    return finish(rc);
  }
  if (n == "lm.lstm.bias") {
    if ((int64_t)numel != 4 * H) return bad_shape();
    std::vector<float> perm(numel);
    for (int64_t c = 0; c < 4 * H; ++c) perm[4 * (c % H) + c / H] = data_host[c];
    Weight& w = ctx->w[n];
    DC_CHECK_CUDA(ctx, cudaMemcpy(w.f32, perm.data(), numel * sizeof(float), cudaMemcpyHostToDevice));
    w.loaded = true;
    ctx->lm_gate_table_valid = false;
    return DC_OK;
  }
  struct F32 { const char* name; int64_t n; };
  const F32 fl[] = {{"fc6.bias", fc}, {"fc7.bias", fc}, {"obj.weight", fc}, {"obj.bias", 1}, {"boxreg.weight", 4 * fc}, {"boxreg.bias", 4},
                    {"lm.encoder.bias", E}, {"lm.vocab.bias", V + 1}};
  for (const F32& f : fl) {
    if (n == f.name) {
      if ((int64_t)numel != f.n) return bad_shape();
      return copy_f32(n, 0, numel);
    }
  }
  return dc_set_error(ctx, DC_E_NOTFOUND, "dc_load_weights: unknown weight name '%s'", name);
}

extern "C" int dc_set_test_args(dc_ctx* ctx, float rpn_nms_thresh, float final_nms_thresh, int num_proposals, int clip_boxes) {
  if (!ctx) return dc_set_error(nullptr, DC_E_INVALID, "dc_set_test_args: null ctx");
  DC_REQUIRE(ctx, num_proposals == -1 || (num_proposals > 0 && num_proposals <= ctx->P), DC_E_INVALID,
             "num_proposals=%d exceeds the configured max_proposals=%d", num_proposals, ctx->P);
  ctx->rpn_nms_thresh = rpn_nms_thresh;
  ctx->final_nms_thresh = final_nms_thresh;
  ctx->num_proposals = (num_proposals == -1) ? ctx->P : num_proposals;
  ctx->clip_boxes = clip_boxes ? 1 : 0;
  return DC_OK;
}

extern "C" int dc_set_decode_all(dc_ctx* ctx, int enable) {
  if (!ctx) return dc_set_error(nullptr, DC_E_INVALID, "dc_set_decode_all: null ctx");
  ctx->decode_all = enable ? 1 : 0;
  return DC_OK;
}

// Activations travel as split-fp16 planes (hi + lo), exact only for |x| < 65504; the kernels flag any value outside that range
// (He-init / trained VGG activations stay below ~1e4, but nothing in the math forbids larger ones).  Reads and clears the sticky
// flag after synchronising `stream`; returns DC_E_RANGE when a forward since the last check overflowed (its results are invalid).
extern "C" int dc_check_range(dc_ctx* ctx, void* stream) {
  if (!ctx) return dc_set_error(nullptr, DC_E_INVALID, "dc_check_range: null ctx");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  unsigned int flag = 0;
  DC_CHECK_CUDA(ctx, cudaMemcpyAsync(&flag, ctx->overflow_flag, sizeof(flag), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  DC_CHECK_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
  if (flag) {
    DC_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->overflow_flag, 0, sizeof(flag), (cudaStream_t)stream));
    return dc_set_error(ctx, DC_E_RANGE, "an activation left the fp16 range of the split hi+lo representation (|x| >= 65504 or NaN): "
                                         "results of the forward passes since the last check are invalid; rescale the input or the weights");
  }
  return DC_OK;
}

// Accuracy / speed knob of the tcgen05 kernels: k-blocks (64 K each) accumulated in the truncating TMEM accumulator before the partial sum
// is added in fp32 registers with round-to-nearest, per tile width (64 / 128 / 256 columns).  1 = closest to an fp32 sgemm.
extern "C" int dc_set_accumulation_groups(dc_ctx* ctx, int kb_tile64, int kb_tile128, int kb_tile256) {
  if (!ctx) return dc_set_error(nullptr, DC_E_INVALID, "dc_set_accumulation_groups: null ctx");
  DC_REQUIRE(ctx, kb_tile64 >= 1 && kb_tile128 >= 1 && kb_tile256 >= 1, DC_E_INVALID, "accumulation groups must be >= 1 k-block");
  ctx->launch.group_kb[0] = kb_tile64; ctx->launch.group_kb[1] = kb_tile128; ctx->launch.group_kb[2] = kb_tile256;
  return DC_OK;
}

extern "C" int dc_set_beam_size(dc_ctx* ctx, int beam_size) {
  if (!ctx) return dc_set_error(nullptr, DC_E_INVALID, "dc_set_beam_size: null ctx");
  DC_REQUIRE(ctx, beam_size >= 0 && beam_size <= 32, DC_E_INVALID, "beam_size must be in [0, 32] (0 = greedy sampling)");
  ctx->beam_size = beam_size;
  return DC_OK;
}

extern "C" int dc_set_image_size(dc_ctx* ctx, int image_h, int image_w) {
  if (!ctx) return dc_set_error(nullptr, DC_E_INVALID, "dc_set_image_size: null ctx");
  DC_REQUIRE(ctx, image_h > 0 && image_w > 0, DC_E_INVALID, "setImageSize: bad size %dx%d", image_h, image_w);
  ctx->image_h = image_h;
  ctx->image_w = image_w;
  ctx->size_consumed = false;
  ctx->size_bwd_consumed = false;
  return DC_OK;
}

// LocalizationLayer options used by the training-time forward (LocalizationLayer.lua:79-87; train_opts.lua:18-42)
extern "C" int dc_set_train_args(dc_ctx* ctx, float mid_box_reg_weight, float mid_objectness_weight, float box_reg_decay,
                                 float sampler_low_thresh, float sampler_high_thresh, int sampler_batch_size, int train_remove_outbounds_boxes) {
  if (!ctx) return dc_set_error(nullptr, DC_E_INVALID, "dc_set_train_args: null ctx");
  DC_REQUIRE(ctx, sampler_batch_size >= 2 && sampler_batch_size <= ctx->P * ctx->B && sampler_batch_size <= 1024, DC_E_INVALID,
             "sampler_batch_size=%d exceeds the ROI capacity %d", sampler_batch_size, std::min(ctx->P * ctx->B, 1024));
  ctx->mid_box_reg_weight = mid_box_reg_weight;
  ctx->mid_objectness_weight = mid_objectness_weight;
  ctx->box_reg_decay = box_reg_decay;
  ctx->sampler_low = sampler_low_thresh;
  ctx->sampler_high = sampler_high_thresh;
  ctx->sampler_batch = sampler_batch_size;
  ctx->remove_outbounds = train_remove_outbounds_boxes ? 1 : 0;
  return DC_OK;
}

// nn.BoxSampler minibatch selection: mode 1 = uniform random like the reference's torch.multinomial (default), seeded: the same seed
// replays the same sequence of minibatches; mode 0 = the first P positives / first M negatives in index order (the reference's
// debug_pos_sample_idx / debug_neg_sample_idx hook, BoxSampler.lua:154-159) -- what the parity tests inject into the oracle too.
extern "C" int dc_set_sampler(dc_ctx* ctx, int mode, uint64_t seed) {
  if (!ctx) return dc_set_error(nullptr, DC_E_INVALID, "dc_set_sampler: null ctx");
  DC_REQUIRE(ctx, mode == 0 || mode == 1, DC_E_INVALID, "dc_set_sampler: mode must be 0 (deterministic) or 1 (random)");
  ctx->sampler_mode = mode;
  ctx->sampler_seed = seed;
  ctx->sampler_calls = 0;
  return DC_OK;
}

size_t train_buffers_struct_bytes();
size_t train_buffers_carve(void* tb, char* base, int A, int B2max, int batch_size);
float* train_buffers_roi_boxes(void* tb);
float* train_buffers_gt(void* tb);
float* train_buffers_pos_target_boxes(void* tb);
int32_t* train_buffers_counts(void* tb);
int k_loc_train_sampler(dc_ctx* ctx, const float* head, int ld, int H, int W, int k, const float* anchors_dev, float x0, float y0, float sx,
                        float sy, const float* gt_boxes, int B2, float lo, float hi, int batch_size, int has_bounds, float img_w, float img_h,
                        void* tb, int64_t* pos_target_idx, int64_t* pos_input_idx, int64_t* neg_input_idx, int sampler_mode, uint32_t sampler_salt,
                        cudaStream_t st);
int k_loc_train_losses(dc_ctx* ctx, void* tb, int A, float w_obj, float w_box, float w_decay, float* losses, cudaStream_t st);

static const int kMaxGtBoxes = 512;
static int ensure_train_buffers(dc_ctx* ctx) {
  if (ctx->train && ctx->train_batch >= ctx->sampler_batch) return DC_OK;
  if (ctx->train_mem) cudaFree(ctx->train_mem);
  if (ctx->bwd_mem) cudaFree(ctx->bwd_mem);
  if (!ctx->train) ctx->train = calloc(1, train_buffers_struct_bytes());
  const size_t bytes = train_buffers_carve(nullptr, nullptr, ctx->A_max, kMaxGtBoxes, ctx->sampler_batch);
  cudaError_t e = cudaMalloc((void**)&ctx->train_mem, bytes);
  if (e != cudaSuccess) { ctx->train_mem = nullptr; return dc_set_error(ctx, DC_E_NOMEM, "training workspace cudaMalloc(%zu) failed", bytes); }
  train_buffers_carve(ctx->train, ctx->train_mem, ctx->A_max, kMaxGtBoxes, ctx->sampler_batch);
  ctx->train_batch = ctx->sampler_batch;
  if (!ctx->train_losses) { int rc = dev_alloc(ctx, &ctx->train_losses, 8); if (rc) return rc; }
  return DC_OK;
}

// LocalizationLayer:setGroundTruth (LocalizationLayer.lua:181-186).  Only the boxes take part in the layer's own
// computation; gt_labels are gathered by the caller with pos_target_idx (the layer just indexes them, :455-457).
extern "C" int dc_set_ground_truth(dc_ctx* ctx, const float* gt_boxes, int num_gt, void* stream) {
  if (!ctx || !gt_boxes) return dc_set_error(ctx, DC_E_INVALID, "dc_set_ground_truth: null argument");
  DC_REQUIRE(ctx, num_gt >= 1 && num_gt <= kMaxGtBoxes, DC_E_INVALID, "gt_boxes must have shape (1, B1, 4) with 1 <= B1 <= %d", kMaxGtBoxes);
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure_train_buffers(ctx);
  if (rc) return rc;
  DC_CHECK_CUDA(ctx, cudaMemcpyAsync(train_buffers_gt(ctx->train), gt_boxes, (size_t)num_gt * 4 * sizeof(float), cudaMemcpyDefault,
                                     (cudaStream_t)stream));
  ctx->gt_count = num_gt;
  ctx->gt_consumed = false;
  ctx->gt_bwd_consumed = false;
  return DC_OK;
}

// ------------------------------------------------------------------------------------------------
// phase timing (LocalizationLayer:timeit, LocalizationLayer.lua:219-230)
// ------------------------------------------------------------------------------------------------
static void phase_mark(dc_ctx* ctx, const char* name, cudaStream_t st) {
  PhaseTimer& t = ctx->timer;
  if (!t.enabled) return;
  if (t.used == t.events.size()) {
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    t.events.push_back(ev);
    t.names.push_back(name);
  }
  t.names[t.used] = name;
  cudaEventRecord(t.events[t.used], st);
  ++t.used;
}
extern "C" int dc_enable_timing(dc_ctx* ctx, int enable) {
  if (!ctx) return DC_E_INVALID;
  ctx->timer.enabled = enable != 0;
  ctx->timer.layers = enable == 2;
  ctx->timer.used = 0;
  return DC_OK;
}
extern "C" int dc_get_times(dc_ctx* ctx, const char** names_out, float* ms_out, int capacity, int* count_out) {
  if (!ctx || !count_out) return DC_E_INVALID;
  PhaseTimer& t = ctx->timer;
  int n = 0;
  if (t.used >= 2) {
    DC_CHECK_CUDA(ctx, cudaEventSynchronize(t.events[t.used - 1]));
    for (size_t i = 0; i + 1 < t.used && n < capacity; ++i, ++n) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, t.events[i], t.events[i + 1]);
      if (names_out) names_out[n] = t.names[i];
      if (ms_out) ms_out[n] = ms;
    }
  }
  if (t.roi_valid && n < capacity) {
    float ms = 0.f;
    if (cudaEventSynchronize(t.roi1) == cudaSuccess && cudaEventElapsedTime(&ms, t.roi0, t.roi1) == cudaSuccess) {
      if (names_out) names_out[n] = "kernel:roi_taps+roi_pool(image0)";
      if (ms_out) ms_out[n] = ms;
      ++n;
    }
  }
  *count_out = n;
  return DC_OK;
}

// ------------------------------------------------------------------------------------------------
// dense-layer helpers
// ------------------------------------------------------------------------------------------------
static int need_weight(dc_ctx* ctx, const char* name, Weight** out) {
  auto it = ctx->w.find(name);
  DC_REQUIRE(ctx, it != ctx->w.end(), DC_E_NOTFOUND, "internal: weight '%s' not registered", name);
  DC_REQUIRE(ctx, it->second.loaded, DC_E_PROTOCOL, "weight '%s' has not been loaded (dc_load_weights)", name);
  *out = &it->second;
  return DC_OK;
}

static int gemm(dc_ctx* ctx, const __half* a_hi, size_t a_plane, int lda, const Weight& w, int ldb, int M, int N, int K,
                const GemmEpilogue& ep_in, cudaStream_t st) {
  GemmEpilogue ep = ep_in;
  ep.acc_scale = w.inv_scale;
  return k_tc_gemm(ctx, a_hi, a_plane, lda, w.planes, w.plane_elems, ldb, M, N, K, ep, st);
}

// y = relu(conv3x3(x)+b) [+pool]; returns the spatial size of y through oh/ow
static int conv_layer(dc_ctx* ctx, const __half* x, size_t x_plane, __half* y, int N, int H, int W, int Cin, int Cout, const Weight& wt,
                      const Weight& bias, int pool_after, __half* pool_tmp, int* oh, int* ow, cudaStream_t st) {
  const int ceil_mode = ctx->cfg.pool_ceil;
  const int PH = pooled(H, ceil_mode), PW = pooled(W, ceil_mode);
  const bool fuse = pool_after && ctx->cfg.fuse_pool && ceil_mode;
  if (fuse) {
    const size_t yp = (size_t)N * PH * PW * Cout;
    int rc = k_tc_conv3x3(ctx, x, x_plane, wt.planes, wt.plane_elems, bias.f32, wt.inv_scale, y, y + yp, N, H, W, Cin, Cout, 1, st);
    *oh = PH; *ow = PW;
    return rc;
  }
  __half* dst = pool_after ? pool_tmp : y;
  const size_t yp = (size_t)N * H * W * Cout;
  int rc = k_tc_conv3x3(ctx, x, x_plane, wt.planes, wt.plane_elems, bias.f32, wt.inv_scale, dst, dst + yp, N, H, W, Cin, Cout, 0, st);
  if (rc) return rc;
  if (pool_after) {      // floor-mode pooling (pool_ceil = 0) or fuse_pool = 0: stand-alone pooling kernel
    const size_t pp = (size_t)N * PH * PW * Cout;
    rc = k_maxpool2x2_split(ctx, dst, dst + yp, y, y + pp, N, H, W, Cout, ceil_mode, st);
    *oh = PH; *ow = PW;
    return rc;
  }
  *oh = H; *ow = W;
  return DC_OK;
}

// conv_net1 + conv_net2 (DenseCapModel.lua:61-76): img NCHW fp32 -> ctx->feats (split NHWC)
static int run_trunk(dc_ctx* ctx, const float* img, int N, int H, int W, int* fh, int* fw, cudaStream_t st) {
  const dc_config& c = ctx->cfg;
  DC_REQUIRE(ctx, N >= 1 && N <= ctx->B, DC_E_INVALID, "batch %d exceeds max_batch %d", N, ctx->B);
  DC_REQUIRE(ctx, (size_t)H * W <= (size_t)c.max_image_h * c.max_image_w && H > 0 && W > 0, DC_E_INVALID,
             "image %dx%d exceeds the configured maximum %dx%d", H, W, c.max_image_h, c.max_image_w);
  int efh, efw;
  feat_hw(c, H, W, efh, efw);
  DC_REQUIRE(ctx, (size_t)efh * efw <= (size_t)ctx->fh_max * ctx->fw_max, DC_E_INVALID, "feature map too large");
  Weight *w0, *b0;
  int rc;
  if ((rc = need_weight(ctx, "conv1_1.weight", &w0))) return rc;
  if ((rc = need_weight(ctx, "conv1_1.bias", &b0))) return rc;
  int cur = 0;
  int h = H, w = W, cin = c.widths[0];
  {
    __half* y = ctx->act[0].base;
    const size_t yp = (size_t)N * H * W * cin;
#ifndef DC_DISABLE_TC_CONV1
    const bool tc_first = (cin == 64);
#else
    const bool tc_first = false;
#endif
    if (tc_first) {              // VGG-16 width: im2col + tcgen05 (tc_conv_first.cu); other widths: the generic CUDA-core kernel
      Weight* wt;
      if ((rc = need_weight(ctx, "conv1_1.weight.tc", &wt))) return rc;
      if ((rc = k_tc_conv_first(ctx, img, wt->planes, wt->plane_elems, b0->f32, wt->inv_scale, y, y + yp, N, H, W, st))) return rc;
    } else if ((rc = k_conv_first(ctx, img, w0->f32, b0->f32, y, y + yp, N, H, W, cin, st))) return rc;
  }
  for (int i = 1; i < 13; ++i) {
    if (ctx->timer.layers) phase_mark(ctx, kConvNames[i], st);
    const int cout = c.widths[kConvStage[i]];
    Weight *wt, *bs;
    if ((rc = need_weight(ctx, (std::string(kConvNames[i]) + ".weight").c_str(), &wt))) return rc;
    if ((rc = need_weight(ctx, (std::string(kConvNames[i]) + ".bias").c_str(), &bs))) return rc;
    const __half* x = ctx->act[cur].base;
    const size_t xp = (size_t)N * h * w * cin;
    int oh, ow;
    if (i == 12) {
      // last layer writes straight into the feature buffer (no pool after conv5_3; pool5 is excluded)
      if ((rc = conv_layer(ctx, x, xp, ctx->feats.base, N, h, w, cin, cout, *wt, *bs, 0, nullptr, &oh, &ow, st))) return rc;
    } else {
      const bool unfused_pool = kPoolAfter[i] && !(c.fuse_pool && c.pool_ceil);
      // unfused pool: conv x(cur) -> tmp(other), pool tmp(other) -> y(cur): x is dead by then
      __half* y = unfused_pool ? ctx->act[cur].base : ctx->act[cur ^ 1].base;
      __half* tmp = ctx->act[cur ^ 1].base;
      if ((rc = conv_layer(ctx, x, xp, y, N, h, w, cin, cout, *wt, *bs, kPoolAfter[i], tmp, &oh, &ow, st))) return rc;
      if (!unfused_pool) cur ^= 1;
    }
    h = oh; w = ow; cin = cout;
  }
  *fh = h; *fw = w;
  return DC_OK;
}

__global__ void identity_idx_kernel(const int32_t* __restrict__ count_in, int P, int32_t* idx32, int64_t* idx64, int32_t* count_out) {
  const int n = min(*count_in, P);
  for (int i = threadIdx.x; i < n; i += blockDim.x) { idx32[i] = i; idx64[i] = i + 1; }
  if (threadIdx.x == 0) *count_out = n;
}
__global__ void set_count_kernel(int32_t* p, int v) { *p = v; }

// nn.Sequential RPN (LocalizationLayer.lua:76-109): 3x3 conv + ReLU, then the 1x1 box and score heads as ONE GEMM
// whose 6k output columns are [4k box transforms | 2k scores] per feature cell -> ctx->rpn_head (N*cells, 6k)
static int run_rpn_head(dc_ctx* ctx, int N, int fh, int fw, cudaStream_t st) {
  const dc_config& c = ctx->cfg;
  const int C5 = c.widths[4], F = c.rpn_filters, k = c.num_anchors;
  const int cells = fh * fw;
  DC_REQUIRE(ctx, k * cells <= ctx->A_max, DC_E_INVALID, "feature map %dx%d exceeds capacity", fh, fw);
  Weight *wc, *bc, *wh, *bh;
  int rc;
  if ((rc = need_weight(ctx, "rpn.conv.weight", &wc))) return rc;
  if ((rc = need_weight(ctx, "rpn.conv.bias", &bc))) return rc;
  if ((rc = need_weight(ctx, "rpn.head.weight", &wh))) return rc;
  if ((rc = need_weight(ctx, "rpn.head.bias", &bh))) return rc;
  int oh, ow;
  const size_t fplane = (size_t)N * cells * C5;
  if ((rc = conv_layer(ctx, ctx->feats.base, fplane, ctx->rpn_hidden.base, N, fh, fw, C5, F, *wc, *bc, 0, nullptr, &oh, &ow, st))) return rc;
  GemmEpilogue ep;
  ep.bias = bh->f32;
  ep.out_f32 = ctx->rpn_head;
  ep.ldf = 6 * k;
  const size_t hplane = (size_t)N * cells * F;
  return gemm(ctx, ctx->rpn_hidden.base, hplane, F, *wh, F, N * cells, 6 * k, F, ep, st);
}

// LocalizationLayer.lua:262-338 for image b of the batch, from the RPN head outputs in ctx->rpn_head: anchors + reshape +
// ApplyBoxTransform + clip + objectness (fused), NMS, index gathers -> ctx->roi_boxes / rpn_idx / rpn_count rows of image b
static int run_proposals_image(dc_ctx* ctx, int b, int fh, int fw, int img_h, int img_w, cudaStream_t sb) {
  const dc_config& c = ctx->cfg;
  const int k = c.num_anchors, cells = fh * fw, A = k * cells, P = ctx->P;
  int rc;
  RpnPostOut po;
  po.boxes_xcycwh = ctx->pp_boxes + (size_t)b * ctx->A_max * 4;
  po.boxes_x1y1x2y2 = ctx->pp_x1y1 + (size_t)b * ctx->A_max * 4;
  po.prob = ctx->pp_prob + (size_t)b * ctx->A_max;
  po.src_row = ctx->pp_src + (size_t)b * ctx->A_max;
  po.n_valid = ctx->pp_nvalid + b;
  po.block_counts = ctx->pp_blocks + (size_t)b * (ctx->A_max / 256 + 2);
  if ((rc = k_rpn_postprocess(ctx, ctx->rpn_head + (size_t)b * cells * 6 * k, 6 * k, fh, fw, k, ctx->anchors_dev, c.field_centers[0],
                              c.field_centers[1], c.field_centers[2], c.field_centers[3], (float)img_w, (float)img_h, ctx->clip_boxes, po, sb)))
    return rc;
  if ((rc = k_nms(ctx, ctx->pp_x1y1 + (size_t)b * ctx->A_max * 4, 4, ctx->pp_prob + (size_t)b * ctx->A_max, 1, A, ctx->pp_nvalid + b,
                  ctx->rpn_nms_thresh, ctx->num_proposals, ctx->nms[b], ctx->rpn_idx64 + (size_t)b * P, ctx->rpn_idx32 + (size_t)b * P,
                  ctx->rpn_count + b, sb)))
    return rc;
  return k_gather_rows_f32(ctx, ctx->pp_boxes + (size_t)b * ctx->A_max * 4, 4, ctx->rpn_idx32 + (size_t)b * P, ctx->rpn_count + b, P,
                           ctx->roi_boxes + (size_t)b * P * 4, sb);
}

// RPN + post-processing + NMS + ROI pooling for every image of the batch (LocalizationLayer.lua:250-363)
static int run_localization(dc_ctx* ctx, int N, int fh, int fw, int img_h, int img_w, cudaStream_t st) {
  const dc_config& c = ctx->cfg;
  const int C5 = c.widths[4];
  const int cells = fh * fw;
  int rc;
  phase_mark(ctx, "rpn:forward_test", st);
  if ((rc = run_rpn_head(ctx, N, fh, fw, st))) return rc;
  const size_t fplane = (size_t)N * cells * C5;
  const int P = ctx->P;
  const int hw = c.pool_hw * c.pool_hw;
  const size_t rplane = (size_t)N * P * hw * C5;
  phase_mark(ctx, "nms+roi_pooling", st);
  // fork: image 0 stays on the caller's stream, image b > 0 runs on side stream b-1
  if (N > 1) DC_CHECK_CUDA(ctx, cudaEventRecord(ctx->fork_ev, st));
  for (int b = 0; b < N; ++b) {
    cudaStream_t sb = (b == 0) ? st : ctx->side[b - 1];
    if (b > 0) DC_CHECK_CUDA(ctx, cudaStreamWaitEvent(sb, ctx->fork_ev, 0));
    if ((rc = run_proposals_image(ctx, b, fh, fw, img_h, img_w, sb))) return rc;
    const size_t foff = (size_t)b * cells * C5;
    const size_t roff = (size_t)b * P * hw * C5;
    if (b == 0 && ctx->timer.enabled) {
      if (!ctx->timer.roi0) { cudaEventCreate(&ctx->timer.roi0); cudaEventCreate(&ctx->timer.roi1); }
      cudaEventRecord(ctx->timer.roi0, sb);
    }
    if ((rc = k_roi_pool(ctx, ctx->feats.base + foff, ctx->feats.base + fplane + foff, nullptr, C5, fh, fw, ctx->roi_boxes + (size_t)b * P * 4,
                         ctx->rpn_count + b, P, img_h, img_w, c.pool_hw, c.pool_hw, (char*)ctx->roi_taps + ctx->roi_taps_stride * b,
                         ctx->roi.base + roff, ctx->roi.base + rplane + roff, nullptr, 1, sb)))
      return rc;
    if (b == 0 && ctx->timer.enabled) { cudaEventRecord(ctx->timer.roi1, sb); ctx->timer.roi_valid = true; }
    if (b > 0) {
      DC_CHECK_CUDA(ctx, cudaEventRecord(ctx->side_done[b - 1], sb));
      DC_CHECK_CUDA(ctx, cudaStreamWaitEvent(st, ctx->side_done[b - 1], 0));     // join
    }
  }
  return DC_OK;
}

// recognition head on ctx->roi (rows = N*P): fc6, fc7, objectness, box regression, final boxes
static int run_recog(dc_ctx* ctx, int rows, cudaStream_t st) {
  const dc_config& c = ctx->cfg;
  const int C5 = c.widths[4], fc = c.fc_dim, hw = c.pool_hw * c.pool_hw;
  Weight *w6, *b6, *w7, *b7, *wo, *bo, *wr, *br;
  int rc;
  if ((rc = need_weight(ctx, "fc6.weight", &w6))) return rc;
  if ((rc = need_weight(ctx, "fc6.bias", &b6))) return rc;
  if ((rc = need_weight(ctx, "fc7.weight", &w7))) return rc;
  if ((rc = need_weight(ctx, "fc7.bias", &b7))) return rc;
  if ((rc = need_weight(ctx, "obj.weight", &wo))) return rc;
  if ((rc = need_weight(ctx, "obj.bias", &bo))) return rc;
  if ((rc = need_weight(ctx, "boxreg.weight", &wr))) return rc;
  if ((rc = need_weight(ctx, "boxreg.bias", &br))) return rc;
  phase_mark(ctx, "recog:fc6", st);
  const int K6 = hw * C5;
  const size_t rplane = (size_t)rows * K6;
  const size_t fplane = (size_t)rows * fc;
  {
    GemmEpilogue ep;
    ep.bias = b6->f32; ep.relu = 1;
    ep.out_hi = ctx->fc6o.base; ep.out_lo = ctx->fc6o.base + fplane; ep.ldo = fc;
    if ((rc = gemm(ctx, ctx->roi.base, rplane, K6, *w6, K6, rows, fc, K6, ep, st))) return rc;
  }
  phase_mark(ctx, "recog:fc7", st);
  {
    GemmEpilogue ep;
    ep.bias = b7->f32; ep.relu = 1;
    ep.out_hi = ctx->fc7o.base; ep.out_lo = ctx->fc7o.base + fplane; ep.ldo = fc;
    ep.out_f32 = ctx->codes; ep.ldf = fc;
    if ((rc = gemm(ctx, ctx->fc6o.base, fplane, fc, *w7, fc, rows, fc, fc, ep, st))) return rc;
  }
  phase_mark(ctx, "recog:heads", st);
  if ((rc = k_recog_heads(ctx, ctx->codes, fc, wo->f32, bo->f32, wr->f32, br->f32, ctx->obj, ctx->trans, rows, fc, st))) return rc;
  if ((rc = k_apply_box_transform(ctx, ctx->roi_boxes, ctx->trans, ctx->boxes_all, rows, st))) return rc;
  return DC_OK;
}

// LanguageModel image_encoder (Linear + ReLU, LanguageModel.lua:27-29) on ctx->fc7o (rows) -> ctx->enc (rows, E) split planes.
// Runs for EVERY proposal: in the full forward it executes on the main stream while the final NMS runs on the side streams,
// so it costs no wall time, and the decoder then only gathers the rows it keeps.
static int run_lm_encoder(dc_ctx* ctx, int rows, cudaStream_t st) {
  const dc_config& c = ctx->cfg;
  const int fc = c.fc_dim, E = c.encoding_size;
  Weight *we, *be;
  int rc;
  if ((rc = need_weight(ctx, "lm.encoder.weight", &we))) return rc;
  if ((rc = need_weight(ctx, "lm.encoder.bias", &be))) return rc;
  const size_t fplane = (size_t)rows * fc;
  GemmEpilogue ep;
  ep.bias = be->f32; ep.relu = 1;
  ep.out_hi = ctx->enc.base; ep.out_lo = ctx->enc.base + ctx->enc.cap_elems; ep.ldo = E; ep.out_col0 = 0;
  return gemm(ctx, ctx->fc7o.base, fplane, fc, *we, fc, rows, E, fc, ep, st);
}

struct LmWeights { Weight *wl, *wlstm, *blstm, *wv, *bv; };
static int lm_weights(dc_ctx* ctx, LmWeights& w, cudaStream_t st) {
  const dc_config& c = ctx->cfg;
  int rc;
  if ((rc = need_weight(ctx, "lm.lookup.weight", &w.wl))) return rc;
  if ((rc = need_weight(ctx, "lm.lstm.weight", &w.wlstm))) return rc;
  if ((rc = need_weight(ctx, "lm.lstm.bias", &w.blstm))) return rc;
  if ((rc = need_weight(ctx, "lm.vocab.weight", &w.wv))) return rc;
  if ((rc = need_weight(ctx, "lm.vocab.bias", &w.bv))) return rc;
  // gate table = b + lookup . Wx for every token (the input half of the LSTM pre-activations), built on first use after a weight load
  const int E = c.encoding_size, H = c.rnn_size, V = c.vocab_size;
  if (!ctx->lm_gate_table_valid) {
    GemmEpilogue ep;
    ep.bias = w.blstm->f32;
    ep.acc_scale = w.wlstm->inv_scale;
    ep.out_f32 = ctx->lm_gate_table;
    ep.ldf = 4 * H;
    if ((rc = k_tc_gemm(ctx, w.wl->planes, w.wl->plane_elems, E, w.wlstm->planes, w.wlstm->plane_elems, E + H, V + 2, 4 * H, E, ep, st))) return rc;
    ctx->lm_gate_table_valid = true;
  }
  return DC_OK;
}

// LanguageModel:sample, greedy (LanguageModel.lua:293-348), as one compaction launch + ONE persistent decoder launch (lm_decode.cu)
// over the rows selected by (counts, idx32) of each of the B images (row stride P).  compact_only: row maps only.
static int run_lm_decode(dc_ctx* ctx, int B, int P, const int32_t* counts, const int32_t* idx32, int survivors, int64_t* fin_seq, float* logits,
                         cudaStream_t st, bool compact_only = false) {
  const dc_config& c = ctx->cfg;
  LmWeights w;
  int rc;
  if ((rc = lm_weights(ctx, w, st))) return rc;
  LmDecodeArgs a{};
  a.B = B; a.P = P; a.E = c.encoding_size; a.H = c.rnn_size; a.V = c.vocab_size; a.T = c.seq_length;
  a.counts = counts; a.idx32 = idx32; a.survivors = survivors;
  a.enc_planes = ctx->enc.base; a.enc_plane_elems = ctx->enc.cap_elems;
  a.x_planes = ctx->encx.base; a.x_plane_elems = ctx->encx.cap_elems;
  a.h_planes = ctx->xh.base;
  a.cstate = ctx->cstate;
  a.gate_ws = ctx->dec_gate_ws;
  a.gate_table = ctx->lm_gate_table;
  a.tokens = ctx->dec_tokens;
  a.lstm_planes = w.wlstm->planes; a.lstm_plane_elems = w.wlstm->plane_elems; a.lstm_bias = w.blstm->f32; a.lstm_acc_scale = w.wlstm->inv_scale;
  a.vocab_planes = w.wv->planes; a.vocab_plane_elems = w.wv->plane_elems; a.vocab_bias = w.bv->f32; a.vocab_acc_scale = w.wv->inv_scale;
  a.packed = ctx->packed;
  a.counters = ctx->dec_counters;
  a.m_rows = ctx->dec_rows;
  a.dst_final = ctx->dec_dst_final; a.dst_all = ctx->dec_dst_all;
  a.fin_seq = fin_seq; a.all_seq = ctx->seq64;
  a.logits = logits;
  return k_lm_decode(ctx, a, st, compact_only);
}

// LanguageModel:beamsearch (LanguageModel.lua:170-290) over M rows of ctx->enc listed in rows_src (device; null = rows 0..M-1), in chunks of
// rows_cap / beam ROIs: row = roi * beam + b.  Tokens of the best beam of ROI r go to seq_a[dst_a[r]] and (optionally) seq_b[dst_b[r]].
static int run_lm_beamsearch(dc_ctx* ctx, int M, int beam, const int32_t* rows_src, const int32_t* dst_a, int64_t* seq_a, const int32_t* dst_b,
                             int64_t* seq_b, cudaStream_t st) {
  const dc_config& c = ctx->cfg;
  const int E = c.encoding_size, H = c.rnn_size, V = c.vocab_size, V1 = V + 1, T = c.seq_length;
  DC_REQUIRE(ctx, beam >= 1 && beam <= 32 && beam <= V1, DC_E_INVALID, "beam_size must be in [1, 32]");
  if (M <= 0) return DC_OK;
  LmWeights w;
  int rc;
  if ((rc = lm_weights(ctx, w, st))) return rc;
  dc_ctx::BeamWs& ws = ctx->beam;
  const int rows_cap = 4096;
  if (!ws.mem) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
    const size_t o_log = take((size_t)rows_cap * V1 * 4), o_h0 = take((size_t)2 * rows_cap * H * 2), o_h1 = take((size_t)2 * rows_cap * H * 2),
                 o_c0 = take((size_t)rows_cap * H * 4), o_c1 = take((size_t)rows_cap * H * 4), o_x = take((size_t)2 * rows_cap * E * 2),
                 o_clp = take((size_t)rows_cap * 32 * 4), o_cix = take((size_t)rows_cap * 32 * 4), o_blp = take((size_t)rows_cap * 4),
                 o_b0 = take((size_t)rows_cap * T * 4), o_b1 = take((size_t)rows_cap * T * 4), o_src = take((size_t)rows_cap * 4),
                 o_tok = take((size_t)rows_cap * 4);
    DC_CHECK_CUDA(ctx, cudaMalloc((void**)&ws.mem, off));
    ws.rows_cap = rows_cap;
    ws.logits = (float*)(ws.mem + o_log); ws.h[0] = (__half*)(ws.mem + o_h0); ws.h[1] = (__half*)(ws.mem + o_h1);
    ws.c[0] = (float*)(ws.mem + o_c0); ws.c[1] = (float*)(ws.mem + o_c1); ws.x = (__half*)(ws.mem + o_x);
    ws.cand_lp = (float*)(ws.mem + o_clp); ws.cand_ix = (int32_t*)(ws.mem + o_cix); ws.beam_lp = (float*)(ws.mem + o_blp);
    ws.beams[0] = (int32_t*)(ws.mem + o_b0); ws.beams[1] = (int32_t*)(ws.mem + o_b1); ws.src_row = (int32_t*)(ws.mem + o_src);
    ws.tokens = (int32_t*)(ws.mem + o_tok);
  }
  const size_t hplane = (size_t)rows_cap * H, xplane = (size_t)rows_cap * E;
  const int chunk = rows_cap / beam;
  for (int r0 = 0; r0 < M; r0 += chunk) {
    const int R = std::min(chunk, M - r0), rows = R * beam;
    // image time step (:198-201): encoded vector through the LSTM, h0 = c0 = 0
    if ((rc = k_bs_gather_split(ctx, ctx->enc.base, ctx->enc.base + ctx->enc.cap_elems, E, rows_src, r0, R, ws.x, ws.x + xplane, st))) return rc;
    {
      GemmEpilogue ep;
      ep.acc_scale = w.wlstm->inv_scale; ep.bias = w.blstm->f32;
      ep.lstm_c = ws.c[0]; ep.lstm_H = H; ep.lstm_first = 1;
      ep.out_hi = ws.h[0]; ep.out_lo = ws.h[0] + hplane; ep.ldo = H;
      if ((rc = k_tc_gemm(ctx, ws.x, xplane, E, w.wlstm->planes, w.wlstm->plane_elems, E + H, R, 4 * H, E, ep, st))) return rc;
    }
    // START token (:203-206)
    if ((rc = k_bs_fill_i32(ctx, ws.tokens, V + 1, R, st))) return rc;
    auto lstm_step = [&](int nrows, const __half* h_in, float* cbuf, __half* h_out) -> int {
      GemmEpilogue ep;
      ep.acc_scale = w.wlstm->inv_scale;
      ep.gather_tab = ctx->lm_gate_table; ep.gather_idx = ws.tokens; ep.gather_ld = 4 * H;
      ep.lstm_c = cbuf; ep.lstm_H = H; ep.lstm_first = 0;
      ep.out_hi = h_out; ep.out_lo = h_out + hplane; ep.ldo = H;
      return k_tc_gemm(ctx, h_in, hplane, H, w.wlstm->planes + E, w.wlstm->plane_elems, E + H, nrows, 4 * H, H, ep, st);
    };
    auto vocab = [&](int nrows, const __half* h_in) -> int {
      GemmEpilogue ep;
      ep.acc_scale = w.wv->inv_scale; ep.bias = w.bv->f32;
      ep.out_f32 = ws.logits; ep.ldf = V1;
      return k_tc_gemm(ctx, h_in, hplane, H, w.wv->planes, w.wv->plane_elems, H, nrows, V1, H, ep, st);
    };
    if ((rc = lstm_step(R, ws.h[0], ws.c[0], ws.h[1]))) return rc;
    if ((rc = vocab(R, ws.h[1]))) return rc;
    // beams = top-beam words; states replicated (:208-227): h -> ws.h[0], c -> ws.c[1]
    if ((rc = k_bs_first(ctx, ws.logits, R, V1, beam, T, H, ws.beam_lp, ws.beams[0], ws.tokens, ws.h[1], ws.h[1] + hplane, ws.c[0], ws.h[0],
                         ws.h[0] + hplane, ws.c[1], st)))
      return rc;
    int cur_c = 1, cur_b = 0;
    for (int t = 1; t < T; ++t) {
      if ((rc = lstm_step(rows, ws.h[0], ws.c[cur_c], ws.h[1]))) return rc;
      if ((rc = vocab(rows, ws.h[1]))) return rc;
      if ((rc = k_bs_row_topk(ctx, ws.logits, rows, V1, beam, T, V + 1, ws.beams[cur_b], ws.cand_lp, ws.cand_ix, st))) return rc;
      if ((rc = k_bs_merge(ctx, R, beam, T, t, ws.cand_lp, ws.cand_ix, ws.beam_lp, ws.beams[cur_b], ws.beams[cur_b ^ 1], ws.src_row, ws.tokens, st)))
        return rc;
      if ((rc = k_bs_reindex(ctx, rows, H, ws.src_row, ws.h[1], ws.h[1] + hplane, ws.c[cur_c], ws.h[0], ws.h[0] + hplane, ws.c[cur_c ^ 1], st)))
        return rc;
      cur_c ^= 1; cur_b ^= 1;
    }
    if (seq_a && (rc = k_bs_final(ctx, R, beam, T, ws.beam_lp, ws.beams[cur_b], dst_a ? dst_a + r0 : nullptr, seq_a + (dst_a ? 0 : (size_t)r0 * T), st)))
      return rc;
    if (seq_b && (rc = k_bs_final(ctx, R, beam, T, ws.beam_lp, ws.beams[cur_b], dst_b ? dst_b + r0 : nullptr, seq_b + (dst_b ? 0 : (size_t)r0 * T), st)))
      return rc;
  }
  return DC_OK;
}

// Final NMS (DenseCapModel.lua:261-271) needs only the final boxes and objectness scores of the recognition head.  It is forked
// onto per-image side streams right after the head and runs beside the LM image encoder; join_final_nms makes the caller's stream
// wait for the selection, which decides the rows the caption decoder processes (:272-274 keeps only those rows of the LM output).
static int run_final_nms(dc_ctx* ctx, int N, cudaStream_t st) {
  const int P = ctx->P;
  int rc;
  if ((rc = k_xcycwh_to_x1y1x2y2(ctx, ctx->boxes_all, ctx->fin_x1y1, N * P, st))) return rc;
  DC_CHECK_CUDA(ctx, cudaEventRecord(ctx->fork_ev, st));
  for (int b = 0; b < N; ++b) {
    cudaStream_t sb = (b == 0) ? ctx->fin_stream0 : ctx->side[b - 1];
    DC_CHECK_CUDA(ctx, cudaStreamWaitEvent(sb, ctx->fork_ev, 0));
    const size_t o = (size_t)b * P;
    if (ctx->final_nms_thresh > 0.f) {
      if ((rc = k_nms(ctx, ctx->fin_x1y1 + o * 4, 4, ctx->obj + o, 1, P, ctx->rpn_count + b, ctx->final_nms_thresh, 0, ctx->nms[b],
                      ctx->fin_idx64 + o, ctx->fin_idx32 + o, ctx->fin_count + b, sb)))
        return rc;
    } else {
      identity_idx_kernel<<<1, 256, 0, sb>>>(ctx->rpn_count + b, P, ctx->fin_idx32 + o, ctx->fin_idx64 + o, ctx->fin_count + b);
      DC_CHECK_LAUNCH(ctx);
    }
    DC_CHECK_CUDA(ctx, cudaEventRecord(b == 0 ? ctx->fin_done0 : ctx->side_done[b - 1], sb));
  }
  return DC_OK;
}
static int join_final_nms(dc_ctx* ctx, int N, cudaStream_t st) {
  for (int b = 0; b < N; ++b) DC_CHECK_CUDA(ctx, cudaStreamWaitEvent(st, b == 0 ? ctx->fin_done0 : ctx->side_done[b - 1], 0));
  return DC_OK;
}

// output[4], output[1] (and, when every proposal was decoded, output[5]) filtered by the final-NMS selection (DenseCapModel.lua:272-274)
static int run_final_gather(dc_ctx* ctx, int N, bool with_seq, cudaStream_t st) {
  const int P = ctx->P, T = ctx->cfg.seq_length;
  int rc;
  for (int b = 0; b < N; ++b) {
    const size_t o = (size_t)b * P;
    if ((rc = k_final_gather(ctx, ctx->fin_idx32 + o, ctx->fin_count + b, P, ctx->boxes_all + o * 4, ctx->obj + o,
                             with_seq ? ctx->seq64 + o * T : nullptr, T, ctx->fin_boxes + o * 4, ctx->fin_scores + o,
                             with_seq ? ctx->fin_seq + o * T : nullptr, st)))
      return rc;
  }
  return DC_OK;
}

static void fill_result(dc_ctx* ctx, int N, dc_result* out) {
  out->batch = N;
  out->max_proposals = ctx->P;
  out->num_proposals = ctx->rpn_count;
  out->num_final = ctx->fin_count;
  out->final_boxes = ctx->fin_boxes;
  out->final_scores = ctx->fin_scores;
  out->final_seq = ctx->fin_seq;
  out->final_idx = ctx->fin_idx64;
  out->roi_boxes = ctx->roi_boxes;
  out->box_trans = ctx->trans;
  out->all_boxes = ctx->boxes_all;
  out->all_scores = ctx->obj;
  out->all_seq = ctx->seq64;
  out->codes = ctx->codes;
  out->proposal_idx = ctx->rpn_idx64;
}

extern "C" int dc_forward_test(dc_ctx* ctx, const float* img, int batch, int image_h, int image_w, dc_result* out, void* stream) {
  if (!ctx || !img) return dc_set_error(ctx, DC_E_INVALID, "dc_forward_test: null argument");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  // DenseCapModel:updateOutput sets the image size itself (DenseCapModel.lua:245-246)
  if ((rc = dc_set_image_size(ctx, image_h, image_w))) return rc;
  ctx->timer.used = 0;
  phase_mark(ctx, "conv_trunk", st);
  int fh, fw;
  if ((rc = run_trunk(ctx, img, batch, image_h, image_w, &fh, &fw, st))) return rc;
  ctx->size_consumed = true;
  if ((rc = run_localization(ctx, batch, fh, fw, image_h, image_w, st))) return rc;
  const int rows = batch * ctx->P;
  if ((rc = run_recog(ctx, rows, st))) return rc;
  phase_mark(ctx, "final_nms|lm:encoder", st);
  if ((rc = run_final_nms(ctx, batch, st))) return rc;            // side streams
  if ((rc = run_lm_encoder(ctx, rows, st))) return rc;            // main stream, beside the NMS
  if ((rc = join_final_nms(ctx, batch, st))) return rc;
  phase_mark(ctx, "lm:sample", st);
  // decode only what the reference returns (the final-NMS survivors), or every proposal when all_seq was asked for
  const int32_t* dcounts = ctx->decode_all ? ctx->rpn_count : ctx->fin_count;
  const int32_t* didx = ctx->decode_all ? nullptr : ctx->fin_idx32;
  if (ctx->beam_size > 0) {
    // LanguageModel:beamsearch (self.beam_size ~= nil, LanguageModel.lua:129-132): row maps from the compaction kernel, then one host
    // read of the row count (the beam search is sized by it; this optional path costs ~beam x the greedy decode anyway)
    if ((rc = run_lm_decode(ctx, batch, ctx->P, dcounts, didx, ctx->decode_all ? 0 : 1, ctx->decode_all ? nullptr : ctx->fin_seq, nullptr, st, true)))
      return rc;
    int32_t m_rows = 0;
    DC_CHECK_CUDA(ctx, cudaMemcpyAsync(&m_rows, ctx->dec_rows, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    DC_CHECK_CUDA(ctx, cudaStreamSynchronize(st));
    rc = run_lm_beamsearch(ctx, m_rows, ctx->beam_size, ctx->dec_dst_all, ctx->dec_dst_all, ctx->seq64, ctx->decode_all ? nullptr : ctx->dec_dst_final,
                           ctx->decode_all ? nullptr : ctx->fin_seq, st);
  } else {
    rc = run_lm_decode(ctx, batch, ctx->P, dcounts, didx, ctx->decode_all ? 0 : 1, ctx->decode_all ? nullptr : ctx->fin_seq, nullptr, st);
  }
  if (rc) return rc;
  phase_mark(ctx, "final_gather", st);
  if ((rc = run_final_gather(ctx, batch, ctx->decode_all != 0, st))) return rc;
  phase_mark(ctx, "end", st);
  if (out) fill_result(ctx, batch, out);
  return DC_OK;
}

static int ensure_host_stage(dc_ctx* ctx, size_t f32_elems, size_t u8_elems) {
  if (ctx->host_stage_f32_elems < f32_elems) {
    if (ctx->host_stage_f32) cudaFree(ctx->host_stage_f32);
    ctx->host_stage_f32 = nullptr; ctx->host_stage_f32_elems = 0;
    DC_CHECK_CUDA(ctx, cudaMalloc((void**)&ctx->host_stage_f32, f32_elems * sizeof(float)));
    ctx->host_stage_f32_elems = f32_elems;
  }
  if (ctx->host_stage_u8_elems < u8_elems) {
    if (ctx->host_stage_u8) cudaFree(ctx->host_stage_u8);
    ctx->host_stage_u8 = nullptr; ctx->host_stage_u8_elems = 0;
    DC_CHECK_CUDA(ctx, cudaMalloc((void**)&ctx->host_stage_u8, u8_elems));
    ctx->host_stage_u8_elems = u8_elems;
  }
  return DC_OK;
}

extern "C" int dc_forward_test_host(dc_ctx* ctx, const float* img_host, int batch, int image_h, int image_w, int32_t* num_final_host,
                                    float* final_boxes_host, float* final_scores_host, int64_t* final_seq_host) {
  if (!ctx || !img_host) return dc_set_error(ctx, DC_E_INVALID, "dc_forward_test_host: null argument");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  DC_REQUIRE(ctx, batch >= 1 && batch <= ctx->B, DC_E_INVALID, "batch %d exceeds max_batch %d", batch, ctx->B);
  const size_t n = (size_t)batch * 3 * image_h * image_w;
  int rc0 = ensure_host_stage(ctx, n, 0);
  if (rc0) return rc0;
  float* staging = ctx->host_stage_f32;
  cudaStream_t st = 0;
  DC_CHECK_CUDA(ctx, cudaMemcpyAsync(staging, img_host, n * sizeof(float), cudaMemcpyHostToDevice, st));
  dc_result r;
  int rc = dc_forward_test(ctx, staging, batch, image_h, image_w, &r, st);
  if (rc) return rc;
  const size_t rows = (size_t)batch * ctx->P;
  const int T = ctx->cfg.seq_length;
  if (num_final_host) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(num_final_host, r.num_final, batch * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (final_boxes_host) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(final_boxes_host, r.final_boxes, rows * 4 * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (final_scores_host) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(final_scores_host, r.final_scores, rows * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (final_seq_host) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(final_seq_host, r.final_seq, rows * T * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  return dc_check_range(ctx, st);          // synchronises; DC_E_RANGE if an activation overflowed the split-fp16 planes
}

// run_model.lua:66-74 preprocessing (no resize) on the device: uint8 RGB HWC (batch,H,W,3) -> fp32 BGR mean-subtracted NCHW
extern "C" int dc_preprocess_u8(dc_ctx* ctx, const uint8_t* img_u8, int batch, int image_h, int image_w, float* img_nchw, void* stream) {
  if (!ctx || !img_u8 || !img_nchw) return dc_set_error(ctx, DC_E_INVALID, "dc_preprocess_u8: null argument");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  return k_preprocess_u8(ctx, img_u8, img_nchw, batch, image_h, image_w, (cudaStream_t)stream);
}

// dc_forward_test_host with the decoded uint8 image at the boundary: 4x fewer host->device bytes, preprocessing on the device
extern "C" int dc_forward_test_host_u8(dc_ctx* ctx, const uint8_t* img_host_u8, int batch, int image_h, int image_w, int32_t* num_final_host,
                                       float* final_boxes_host, float* final_scores_host, int64_t* final_seq_host) {
  if (!ctx || !img_host_u8) return dc_set_error(ctx, DC_E_INVALID, "dc_forward_test_host_u8: null argument");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  DC_REQUIRE(ctx, batch >= 1 && batch <= ctx->B, DC_E_INVALID, "batch %d exceeds max_batch %d", batch, ctx->B);
  const size_t n = (size_t)batch * 3 * image_h * image_w;
  int rc0 = ensure_host_stage(ctx, n, n);
  if (rc0) return rc0;
  uint8_t* staging_u8 = ctx->host_stage_u8;
  float* staging_f32 = ctx->host_stage_f32;
  cudaStream_t st = 0;
  DC_CHECK_CUDA(ctx, cudaMemcpyAsync(staging_u8, img_host_u8, n, cudaMemcpyHostToDevice, st));
  int rc = k_preprocess_u8(ctx, staging_u8, staging_f32, batch, image_h, image_w, st);
  if (rc) return rc;
  dc_result r;
  if ((rc = dc_forward_test(ctx, staging_f32, batch, image_h, image_w, &r, st))) return rc;
  const size_t rows = (size_t)batch * ctx->P;
  const int T = ctx->cfg.seq_length;
  if (num_final_host) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(num_final_host, r.num_final, batch * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (final_boxes_host) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(final_boxes_host, r.final_boxes, rows * 4 * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (final_scores_host) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(final_scores_host, r.final_scores, rows * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (final_seq_host) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(final_seq_host, r.final_seq, rows * T * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  return dc_check_range(ctx, st);          // synchronises; DC_E_RANGE if an activation overflowed the split-fp16 planes
}

extern "C" int dc_vgg_trunk_forward(dc_ctx* ctx, const float* img, int batch, int image_h, int image_w, float* feats_nchw, void* stream) {
  if (!ctx || !img || !feats_nchw) return dc_set_error(ctx, DC_E_INVALID, "dc_vgg_trunk_forward: null argument");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  int fh, fw, rc;
  if ((rc = run_trunk(ctx, img, batch, image_h, image_w, &fh, &fw, st))) return rc;
  const int C5 = ctx->cfg.widths[4];
  const size_t plane = (size_t)batch * fh * fw * C5;
  return k_split_nhwc_to_nchw(ctx, ctx->feats.base, ctx->feats.base + plane, feats_nchw, batch, C5, fh, fw, st);
}

extern "C" int dc_localization_forward_test(dc_ctx* ctx, const float* feats_nchw, int feat_h, int feat_w, float* roi_features,
                                            float* roi_boxes, int64_t* idx, int32_t* num_out, void* stream) {
  if (!ctx || !feats_nchw) return dc_set_error(ctx, DC_E_INVALID, "dc_localization_forward_test: null argument");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  // LocalizationLayer.lua:259-261
  DC_REQUIRE(ctx, ctx->image_h > 0 && ctx->image_w > 0 && !ctx->size_consumed, DC_E_PROTOCOL,
             "Must call setImageSize before each forward pass");
  ctx->size_consumed = true;
  cudaStream_t st = (cudaStream_t)stream;
  const dc_config& c = ctx->cfg;
  const int C5 = c.widths[4];
  DC_REQUIRE(ctx, (size_t)feat_h * feat_w <= (size_t)ctx->fh_max * ctx->fw_max, DC_E_INVALID, "feature map too large");
  const size_t plane = (size_t)feat_h * feat_w * C5;
  int rc;
  ctx->timer.used = 0;
  if ((rc = k_nchw_to_split_nhwc(ctx, feats_nchw, ctx->feats.base, ctx->feats.base + plane, 1, C5, feat_h, feat_w, st))) return rc;
  if ((rc = run_localization(ctx, 1, feat_h, feat_w, ctx->image_h, ctx->image_w, st))) return rc;
  phase_mark(ctx, "end", st);
  const int P = ctx->P, hw = c.pool_hw * c.pool_hw;
  if (roi_features) {
    const size_t rplane = (size_t)P * hw * C5;
    if ((rc = k_bhwc_split_to_bchw(ctx, ctx->roi.base, ctx->roi.base + rplane, roi_features, ctx->num_proposals, C5, hw, st))) return rc;
  }
  if (roi_boxes) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(roi_boxes, ctx->roi_boxes, (size_t)ctx->num_proposals * 4 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (idx) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(idx, ctx->rpn_idx64, (size_t)ctx->num_proposals * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  if (num_out) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(num_out, ctx->rpn_count, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  return DC_OK;
}

// (4k | 2k, H', W') NCHW head maps of the reference's RPN -> the (cells, 6k) row-major head matrix of this library
__global__ void rpn_maps_to_head_kernel(const float* __restrict__ trans, const float* __restrict__ scores, float* __restrict__ head, int k, int cells) {
  const int ld = 6 * k;
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < (size_t)cells * ld; o += (size_t)gridDim.x * blockDim.x) {
    const int col = o % ld;
    const size_t cell = o / ld;
    head[o] = col < 4 * k ? trans[(size_t)col * cells + cell] : scores[(size_t)(col - 4 * k) * cells + cell];
  }
}

// LocalizationLayer:_forward_test after self.rpn:forward (LocalizationLayer.lua:262-338), fed with the RPN's raw outputs
extern "C" int dc_rpn_proposals(dc_ctx* ctx, const float* rpn_trans_nchw, const float* rpn_scores_nchw, int feat_h, int feat_w, float* roi_boxes,
                                int64_t* idx, int32_t* num_out, void* stream) {
  if (!ctx || !rpn_trans_nchw || !rpn_scores_nchw) return dc_set_error(ctx, DC_E_INVALID, "dc_rpn_proposals: null argument");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  DC_REQUIRE(ctx, ctx->image_h > 0 && ctx->image_w > 0 && !ctx->size_consumed, DC_E_PROTOCOL, "Must call setImageSize before each forward pass");
  ctx->size_consumed = true;
  cudaStream_t st = (cudaStream_t)stream;
  const int k = ctx->cfg.num_anchors, cells = feat_h * feat_w;
  DC_REQUIRE(ctx, feat_h > 0 && feat_w > 0 && (size_t)cells <= (size_t)ctx->fh_max * ctx->fw_max, DC_E_INVALID, "feature map too large");
  rpn_maps_to_head_kernel<<<(unsigned)std::min<size_t>(1184, DC_CEIL_DIV((size_t)cells * 6 * k, (size_t)256)), 256, 0, st>>>(
      rpn_trans_nchw, rpn_scores_nchw, ctx->rpn_head, k, cells);
  DC_CHECK_LAUNCH(ctx);
  int rc;
  if ((rc = run_proposals_image(ctx, 0, feat_h, feat_w, ctx->image_h, ctx->image_w, st))) return rc;
  const size_t n = (size_t)ctx->num_proposals;
  if (roi_boxes) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(roi_boxes, ctx->roi_boxes, n * 4 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (idx) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(idx, ctx->rpn_idx64, n * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  if (num_out) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(num_out, ctx->rpn_count, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  return DC_OK;
}

// nn.LocalizationLayer:_forward_train (LocalizationLayer.lua:383-526)
extern "C" int dc_localization_forward_train(dc_ctx* ctx, const float* feats_nchw, int feat_h, int feat_w, float* roi_features,
                                             float* roi_boxes, float* pos_target_boxes, int64_t* pos_target_idx, int64_t* pos_input_idx,
                                             int64_t* neg_input_idx, int32_t* num_pos_neg, float* losses, void* stream) {
  if (!ctx || !feats_nchw) return dc_set_error(ctx, DC_E_INVALID, "dc_localization_forward_train: null argument");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  // same order as the reference's asserts (:385-393)
  DC_REQUIRE(ctx, ctx->gt_count > 0 && !ctx->gt_consumed, DC_E_PROTOCOL, "Must call setGroundTruth before training-time forward pass");
  ctx->gt_consumed = true;
  DC_REQUIRE(ctx, ctx->image_h > 0 && ctx->image_w > 0 && !ctx->size_consumed, DC_E_PROTOCOL,
             "Must call setImageSize before each forward pass");
  ctx->size_consumed = true;
  cudaStream_t st = (cudaStream_t)stream;
  const dc_config& c = ctx->cfg;
  const int C5 = c.widths[4], k = c.num_anchors, hw = c.pool_hw * c.pool_hw;
  DC_REQUIRE(ctx, (size_t)feat_h * feat_w <= (size_t)ctx->fh_max * ctx->fw_max, DC_E_INVALID, "feature map too large");
  int rc;
  if ((rc = ensure_train_buffers(ctx))) return rc;
  const int cells = feat_h * feat_w, A = k * cells, S = ctx->sampler_batch;
  const size_t plane = (size_t)cells * C5;
  ctx->timer.used = 0;
  if ((rc = k_nchw_to_split_nhwc(ctx, feats_nchw, ctx->feats.base, ctx->feats.base + plane, 1, C5, feat_h, feat_w, st))) return rc;
  phase_mark(ctx, "rpn:forward", st);
  if ((rc = run_rpn_head(ctx, 1, feat_h, feat_w, st))) return rc;
  phase_mark(ctx, "sampler:forward", st);
  // torch.multinomial draws fresh samples on every call: the salt mixes the seed with a per-context call counter
  uint32_t sampler_salt = (uint32_t)(ctx->sampler_seed ^ (ctx->sampler_seed >> 32)) * 0x9e3779b1u + (uint32_t)ctx->sampler_calls * 0x85ebca6bu + 1u;
  if (ctx->sampler_mode == 1) ++ctx->sampler_calls;
  if ((rc = k_loc_train_sampler(ctx, ctx->rpn_head, 6 * k, feat_h, feat_w, k, ctx->anchors_dev, c.field_centers[0], c.field_centers[1],
                                c.field_centers[2], c.field_centers[3], train_buffers_gt(ctx->train), ctx->gt_count, ctx->sampler_low,
                                ctx->sampler_high, S, ctx->remove_outbounds, (float)ctx->image_w, (float)ctx->image_h, ctx->train, pos_target_idx, pos_input_idx,
                                neg_input_idx, ctx->sampler_mode, sampler_salt, st)))
    return rc;
  phase_mark(ctx, "roi_pooling:forward", st);
  const size_t rplane = (size_t)S * hw * C5;
  if ((rc = k_roi_pool(ctx, ctx->feats.base, ctx->feats.base + plane, nullptr, C5, feat_h, feat_w, train_buffers_roi_boxes(ctx->train), nullptr, S,
                       ctx->image_h, ctx->image_w, c.pool_hw, c.pool_hw, ctx->roi_taps, ctx->roi.base, ctx->roi.base + rplane, nullptr, 1, st)))
    return rc;
  phase_mark(ctx, "losses", st);
  if ((rc = k_loc_train_losses(ctx, ctx->train, A, ctx->mid_objectness_weight, ctx->mid_box_reg_weight, ctx->box_reg_decay, ctx->train_losses,
                               st)))
    return rc;
  phase_mark(ctx, "end", st);
  if (roi_features && (rc = k_bhwc_split_to_bchw(ctx, ctx->roi.base, ctx->roi.base + rplane, roi_features, S, C5, hw, st))) return rc;
  if (roi_boxes) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(roi_boxes, train_buffers_roi_boxes(ctx->train), (size_t)S * 16, cudaMemcpyDeviceToDevice, st));
  if (pos_target_boxes)
    DC_CHECK_CUDA(ctx, cudaMemcpyAsync(pos_target_boxes, train_buffers_pos_target_boxes(ctx->train), (size_t)S * 16, cudaMemcpyDeviceToDevice, st));
  if (num_pos_neg) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(num_pos_neg, train_buffers_counts(ctx->train), 4 * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  if (losses) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(losses, ctx->train_losses, 5 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  ctx->train_fwd_valid = true;
  ctx->train_fh = feat_h;
  ctx->train_fw = feat_w;
  return DC_OK;
}

// ---- backward of the training-time forward ------------------------------------------------------------------
int k_loc_bwd_scatter(dc_ctx* ctx, void* tb, int A, int batch_size, const float* grad_roi_boxes_in, const float* grad_boxes_pool, float* g_boxes_all,
                      float* g_trans_all, float* g_scores_all, cudaStream_t st);
int k_loc_bwd_head_grad(dc_ctx* ctx, void* tb, int H, int W, int k, const float* g_boxes_all, const float* g_trans_all, const float* g_scores_all,
                        float w_decay, float* head_grad, int ldh, cudaStream_t st);
int k_bwd_relu_mask(dc_ctx* ctx, float* gh, const __half* hid_hi, const __half* hid_lo, size_t n, __half* out_hi, __half* out_lo, float* hid_f32,
                    cudaStream_t st);
int k_bwd_combine(dc_ctx* ctx, const float* g_roi_hwc, const __half* g_rpn_hi, const __half* g_rpn_lo, int C, int cells, float* out_chw, cudaStream_t st);
int k_bwd_wgrad(dc_ctx* ctx, const float* A, int Mdim, const float* B, int Ndim, int H, int W, int taps, float* out, const float* bias_src,
                float* bias_out, cudaStream_t st);
int k_bchw_to_bhwc_f32(dc_ctx* ctx, const float* in, float* out, int K, int C, int npos, cudaStream_t st);

static int ensure_bwd_buffers(dc_ctx* ctx) {
  const int S = ctx->sampler_batch;
  if (ctx->bwd_mem && ctx->bwd_S >= S) return DC_OK;
  const dc_config& c = ctx->cfg;
  const size_t C5 = c.widths[4], F = c.rpn_filters, hw = (size_t)c.pool_hw * c.pool_hw, cells = (size_t)ctx->fh_max * ctx->fw_max, A = ctx->A_max;
  // keep the parameter-gradient accumulators when only the sampler batch grew
  std::vector<float> keep;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  const size_t o_go = take(S * hw * C5 * 4), o_gfr = take(cells * C5 * 4), o_gg = take(S * hw * 8), o_gbp = take((size_t)S * 16),
               o_gba = take(A * 16), o_gta = take(A * 16), o_gsa = take(A * 8), o_hg = take(cells * kHeadGradLd * 4), o_gh = take(cells * F * 4),
               o_hf = take(cells * F * 4), o_ff = take(cells * C5 * 4), o_hgp = take(2 * cells * kHeadGradLd * 2), o_ghp = take(2 * cells * F * 2),
               o_grp = take(2 * cells * C5 * 2), o_dwc = take(F * C5 * 9 * 4), o_dbc = take(F * 4), o_dwh = take(kHeadGradLd * F * 4),
               o_dbh = take(kHeadGradLd * 4);
  char* mem = nullptr;
  cudaError_t e = cudaMalloc((void**)&mem, off);
  if (e != cudaSuccess) return dc_set_error(ctx, DC_E_NOMEM, "backward workspace cudaMalloc(%zu) failed", off);
  cudaMemset(mem, 0, off);
  if (ctx->bwd_mem) {
    cudaMemcpy(mem + o_dwc, ctx->bwd.dW_conv, F * C5 * 9 * 4, cudaMemcpyDeviceToDevice);
    cudaMemcpy(mem + o_dbc, ctx->bwd.db_conv, F * 4, cudaMemcpyDeviceToDevice);
    cudaMemcpy(mem + o_dwh, ctx->bwd.dW_head, kHeadGradLd * F * 4, cudaMemcpyDeviceToDevice);
    cudaMemcpy(mem + o_dbh, ctx->bwd.db_head, kHeadGradLd * 4, cudaMemcpyDeviceToDevice);
    cudaFree(ctx->bwd_mem);
  }
  ctx->bwd_mem = mem;
  ctx->bwd_S = S;
  dc_ctx::Bwd& b = ctx->bwd;
  b.go_bhwc = (float*)(mem + o_go); b.gfeat_roi = (float*)(mem + o_gfr); b.ggrid = (float*)(mem + o_gg); b.gboxes_pool = (float*)(mem + o_gbp);
  b.g_boxes_all = (float*)(mem + o_gba); b.g_trans_all = (float*)(mem + o_gta); b.g_scores_all = (float*)(mem + o_gsa);
  b.head_grad = (float*)(mem + o_hg); b.gh = (float*)(mem + o_gh); b.hid_f32 = (float*)(mem + o_hf); b.feats_f32 = (float*)(mem + o_ff);
  b.head_grad_planes = (__half*)(mem + o_hgp); b.gh_planes = (__half*)(mem + o_ghp); b.grpn_planes = (__half*)(mem + o_grp);
  b.dW_conv = (float*)(mem + o_dwc); b.db_conv = (float*)(mem + o_dbc); b.dW_head = (float*)(mem + o_dwh); b.db_head = (float*)(mem + o_dbh);
  return DC_OK;
}

// nn.LocalizationLayer:backward = updateGradInput (LocalizationLayer.lua:529-605) + accGradParameters of the RPN convolutions
extern "C" int dc_localization_backward_train(dc_ctx* ctx, const float* feats_nchw, int feat_h, int feat_w, const float* grad_roi_features,
                                              const float* grad_roi_boxes, float* grad_feats_nchw, void* stream) {
  if (!ctx || !feats_nchw || !grad_roi_features || !grad_roi_boxes || !grad_feats_nchw)
    return dc_set_error(ctx, DC_E_INVALID, "dc_localization_backward_train: null argument");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  // the reference's asserts (:531-537), with its messages
  DC_REQUIRE(ctx, ctx->gt_count > 0 && !ctx->gt_bwd_consumed, DC_E_PROTOCOL, "Must call setGroundTruth before each forward pass");
  ctx->gt_bwd_consumed = true;
  DC_REQUIRE(ctx, ctx->image_h > 0 && ctx->image_w > 0 && !ctx->size_bwd_consumed, DC_E_PROTOCOL, "Must call setImageSize before each forward pass");
  ctx->size_bwd_consumed = true;
  DC_REQUIRE(ctx, ctx->train_fwd_valid && ctx->train_fh == feat_h && ctx->train_fw == feat_w, DC_E_PROTOCOL,
             "backward needs the training-time forward of the same input to have run on this context");
  ctx->train_fwd_valid = false;
  cudaStream_t st = (cudaStream_t)stream;
  const dc_config& c = ctx->cfg;
  const int C5 = c.widths[4], F = c.rpn_filters, k = c.num_anchors, hw = c.pool_hw * c.pool_hw;
  const int cells = feat_h * feat_w, A = k * cells, S = ctx->sampler_batch, ldh = (int)kHeadGradLd;
  DC_REQUIRE(ctx, 6 * k <= ldh, DC_E_UNSUPPORTED, "backward supports at most %d anchors", ldh / 6);
  int rc;
  if ((rc = ensure_bwd_buffers(ctx))) return rc;
  dc_ctx::Bwd& b = ctx->bwd;
  Weight *wcb, *whb;
  if ((rc = need_weight(ctx, "rpn.conv.weight.bwd", &wcb))) return rc;
  if ((rc = need_weight(ctx, "rpn.head.weight.bwd", &whb))) return rc;
  const size_t plane = (size_t)cells * C5;
  ctx->timer.used = 0;
  // the input again (it is the caller's tensor in the reference too): split NHWC planes for the sampler backward, fp32 HWC for gradWeight
  if ((rc = k_nchw_to_split_nhwc(ctx, feats_nchw, ctx->feats.base, ctx->feats.base + plane, 1, C5, feat_h, feat_w, st))) return rc;
  if ((rc = k_split_to_rowmajor(ctx, ctx->feats.base, ctx->feats.base + plane, b.feats_f32, plane, st))) return rc;
  // ---- RoI pooling backward (:573-580)
  phase_mark(ctx, "roi_pooling:backward", st);
  if ((rc = k_bchw_to_bhwc_f32(ctx, grad_roi_features, b.go_bhwc, S, C5, hw, st))) return rc;
  if ((rc = k_roi_pool_backward(ctx, ctx->feats.base, ctx->feats.base + plane, C5, feat_h, feat_w, train_buffers_roi_boxes(ctx->train), S, ctx->image_h,
                                ctx->image_w, c.pool_hw, c.pool_hw, b.go_bhwc, b.gfeat_roi, b.ggrid, b.gboxes_pool, st)))
    return rc;
  // ---- sampler backward (:583-596) and the RPN's box-space modules (:599-602)
  phase_mark(ctx, "sampler:backward", st);
  if ((rc = k_loc_bwd_scatter(ctx, ctx->train, A, S, grad_roi_boxes, b.gboxes_pool, b.g_boxes_all, b.g_trans_all, b.g_scores_all, st))) return rc;
  phase_mark(ctx, "rpn:backward", st);
  if ((rc = k_loc_bwd_head_grad(ctx, ctx->train, feat_h, feat_w, k, b.g_boxes_all, b.g_trans_all, b.g_scores_all, ctx->box_reg_decay, b.head_grad, ldh,
                                st)))
    return rc;
  // 1x1 heads, backward data: grad_hidden (cells, F) = head_grad (cells, 6k) . W_head (6k, F)
  const size_t hplane = (size_t)cells * ldh;
  if ((rc = k_rowmajor_to_split(ctx, b.head_grad, b.head_grad_planes, b.head_grad_planes + hplane, hplane, st))) return rc;
  {
    GemmEpilogue ep;
    ep.out_f32 = b.gh;
    ep.ldf = F;
    if ((rc = gemm(ctx, b.head_grad_planes, hplane, ldh, *whb, ldh, cells, F, ldh, ep, st))) return rc;
  }
  // ReLU backward + split; the hidden activations of the forward pass are still in ctx->rpn_hidden
  const size_t gplane = (size_t)cells * F;
  if ((rc = k_bwd_relu_mask(ctx, b.gh, ctx->rpn_hidden.base, ctx->rpn_hidden.base + gplane, gplane, b.gh_planes, b.gh_planes + gplane, b.hid_f32, st)))
    return rc;
  // 3x3 conv, backward data = 3x3 conv of grad_hidden with the flipped / transposed bank (no bias, no ReLU)
  rc = k_tc_conv3x3(ctx, b.gh_planes, gplane, wcb->planes, wcb->plane_elems, nullptr, wcb->inv_scale, b.grpn_planes, b.grpn_planes + plane, 1, feat_h,
                    feat_w, F, C5, 0, st, 0);
  if (rc) return rc;
  if ((rc = k_bwd_combine(ctx, b.gfeat_roi, b.grpn_planes, b.grpn_planes + plane, C5, cells, grad_feats_nchw, st))) return rc;
  // ---- accGradParameters: gradWeight / gradBias of the RPN convolutions accumulate until dc_zero_grad_parameters
  phase_mark(ctx, "rpn:accGradParameters", st);
  if ((rc = k_bwd_wgrad(ctx, b.head_grad, ldh, b.hid_f32, F, feat_h, feat_w, 1, b.dW_head, b.head_grad, b.db_head, st))) return rc;
  if ((rc = k_bwd_wgrad(ctx, b.gh, F, b.feats_f32, C5, feat_h, feat_w, 9, b.dW_conv, b.gh, b.db_conv, st))) return rc;
  phase_mark(ctx, "end", st);
  return DC_OK;
}

// nn.Module:zeroGradParameters for the layer's parameters (the RPN)
extern "C" int dc_zero_grad_parameters(dc_ctx* ctx, void* stream) {
  if (!ctx) return dc_set_error(nullptr, DC_E_INVALID, "dc_zero_grad_parameters: null ctx");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure_bwd_buffers(ctx);
  if (rc) return rc;
  const dc_config& c = ctx->cfg;
  const size_t C5 = c.widths[4], F = c.rpn_filters;
  cudaStream_t st = (cudaStream_t)stream;
  DC_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->bwd.dW_conv, 0, F * C5 * 9 * 4, st));
  DC_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->bwd.db_conv, 0, F * 4, st));
  DC_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->bwd.dW_head, 0, kHeadGradLd * F * 4, st));
  DC_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->bwd.db_head, 0, kHeadGradLd * 4, st));
  return DC_OK;
}

// gradWeight / gradBias by the reference's parameter names: rpn.conv.weight (F,C,3,3), rpn.conv.bias (F), rpn.box.weight (4k,F),
// rpn.box.bias (4k), rpn.score.weight (2k,F), rpn.score.bias (2k).  out: device or host memory with room for numel floats.
extern "C" int dc_get_grad_parameters(dc_ctx* ctx, const char* name, float* out, int64_t numel, void* stream) {
  if (!ctx || !name || !out) return dc_set_error(ctx, DC_E_INVALID, "dc_get_grad_parameters: null argument");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  DC_REQUIRE(ctx, ctx->bwd_mem != nullptr, DC_E_PROTOCOL, "no gradients yet: run dc_localization_backward_train first");
  const dc_config& c = ctx->cfg;
  const int64_t C5 = c.widths[4], F = c.rpn_filters, k = c.num_anchors;
  const std::string n(name);
  const float* src = nullptr;
  int64_t want = 0;
  if (n == "rpn.conv.weight") { src = ctx->bwd.dW_conv; want = F * C5 * 9; }
  else if (n == "rpn.conv.bias") { src = ctx->bwd.db_conv; want = F; }
  else if (n == "rpn.box.weight") { src = ctx->bwd.dW_head; want = 4 * k * F; }
  else if (n == "rpn.box.bias") { src = ctx->bwd.db_head; want = 4 * k; }
  else if (n == "rpn.score.weight") { src = ctx->bwd.dW_head + 4 * k * F; want = 2 * k * F; }
  else if (n == "rpn.score.bias") { src = ctx->bwd.db_head + 4 * k; want = 2 * k; }
  else return dc_set_error(ctx, DC_E_INVALID, "dc_get_grad_parameters: unknown parameter '%s'", name);
  DC_REQUIRE(ctx, numel == want, DC_E_INVALID, "dc_get_grad_parameters(%s): expected %lld elements", name, (long long)want);
  DC_CHECK_CUDA(ctx, cudaMemcpyAsync(out, src, (size_t)want * sizeof(float), cudaMemcpyDefault, (cudaStream_t)stream));
  return DC_OK;
}

extern "C" int dc_recog_forward(dc_ctx* ctx, const float* roi_features, const float* roi_boxes, int K, float* codes, float* obj_scores,
                                float* box_trans, float* final_boxes, void* stream) {
  if (!ctx || !roi_features || !roi_boxes) return dc_set_error(ctx, DC_E_INVALID, "dc_recog_forward: null argument");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  DC_REQUIRE(ctx, K >= 1 && K <= ctx->P * ctx->B, DC_E_INVALID, "K=%d exceeds capacity %d", K, ctx->P * ctx->B);
  cudaStream_t st = (cudaStream_t)stream;
  const dc_config& c = ctx->cfg;
  const int C5 = c.widths[4], hw = c.pool_hw * c.pool_hw, fc = c.fc_dim;
  int rc;
  const size_t rplane = (size_t)K * hw * C5;
  if ((rc = k_bchw_to_split_rows(ctx, roi_features, ctx->roi.base, ctx->roi.base + rplane, K, C5, hw, st))) return rc;
  DC_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->roi_boxes, roi_boxes, (size_t)K * 4 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  ctx->timer.used = 0;
  if ((rc = run_recog(ctx, K, st))) return rc;
  if (codes) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(codes, ctx->codes, (size_t)K * fc * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (obj_scores) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(obj_scores, ctx->obj, (size_t)K * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (box_trans) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(box_trans, ctx->trans, (size_t)K * 4 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (final_boxes) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(final_boxes, ctx->boxes_all, (size_t)K * 4 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return DC_OK;
}

extern "C" int dc_lm_sample(dc_ctx* ctx, const float* codes, int K, int64_t* seq, float* logits, void* stream) {
  if (!ctx || !codes || !seq) return dc_set_error(ctx, DC_E_INVALID, "dc_lm_sample: null argument");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  DC_REQUIRE(ctx, K >= 1 && K <= ctx->P * ctx->B, DC_E_INVALID, "K=%d exceeds capacity %d", K, ctx->P * ctx->B);
  cudaStream_t st = (cudaStream_t)stream;
  const int fc = ctx->cfg.fc_dim, T = ctx->cfg.seq_length;
  const size_t fplane = (size_t)K * fc;
  int rc;
  if ((rc = k_rowmajor_to_split(ctx, codes, ctx->fc7o.base, ctx->fc7o.base + fplane, fplane, st))) return rc;
  ctx->timer.used = 0;
  if ((rc = run_lm_encoder(ctx, K, st))) return rc;
  set_count_kernel<<<1, 1, 0, st>>>(ctx->dec_count1, K);
  DC_CHECK_LAUNCH(ctx);
  if ((rc = run_lm_decode(ctx, 1, K, ctx->dec_count1, nullptr, 0, nullptr, logits, st))) return rc;
  DC_CHECK_CUDA(ctx, cudaMemcpyAsync(seq, ctx->seq64, (size_t)K * T * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  return DC_OK;
}

// LanguageModel:beamsearch(image_vectors, beam_size) (LanguageModel.lua:170-290)
extern "C" int dc_lm_beamsearch(dc_ctx* ctx, const float* codes, int K, int beam_size, int64_t* seq, void* stream) {
  if (!ctx || !codes || !seq) return dc_set_error(ctx, DC_E_INVALID, "dc_lm_beamsearch: null argument");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  DC_REQUIRE(ctx, K >= 1 && K <= ctx->P * ctx->B, DC_E_INVALID, "K=%d exceeds capacity %d", K, ctx->P * ctx->B);
  cudaStream_t st = (cudaStream_t)stream;
  const int fc = ctx->cfg.fc_dim;
  const size_t fplane = (size_t)K * fc;
  int rc;
  if ((rc = k_rowmajor_to_split(ctx, codes, ctx->fc7o.base, ctx->fc7o.base + fplane, fplane, st))) return rc;
  if ((rc = run_lm_encoder(ctx, K, st))) return rc;
  return run_lm_beamsearch(ctx, K, beam_size > 0 ? beam_size : 20, nullptr, nullptr, seq, nullptr, nullptr, st);
}

// ------------------------------------------------------------------------------------------------
// leaf operators
// ------------------------------------------------------------------------------------------------
extern "C" int dc_make_anchors(float* out, int N, int H, int W, int k, float x0, float y0, float sx, float sy, const float* anchors_host, void* stream) {
  if (!out || !anchors_host || k <= 0 || k > DC_MAX_ANCHORS) return dc_set_error(nullptr, DC_E_INVALID, "dc_make_anchors: bad argument");
  float* a = nullptr;
  if (cudaMalloc((void**)&a, sizeof(float) * 2 * k) != cudaSuccess) return dc_set_error(nullptr, DC_E_NOMEM, "dc_make_anchors: cudaMalloc");
  cudaMemcpyAsync(a, anchors_host, sizeof(float) * 2 * k, cudaMemcpyHostToDevice, (cudaStream_t)stream);
  int rc = k_make_anchors(nullptr, out, N, H, W, k, x0, y0, sx, sy, a, (cudaStream_t)stream);
  cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(a);
  return rc;
}
extern "C" int dc_reshape_box_features(const float* in, float* out, int N, int k, int D, int H, int W, void* stream) {
  if (!in || !out) return dc_set_error(nullptr, DC_E_INVALID, "dc_reshape_box_features: null");
  return k_reshape_box_features(nullptr, in, out, N, k, D, H, W, (cudaStream_t)stream);
}
extern "C" int dc_apply_box_transform(const float* boxes, const float* trans, float* out, int n, void* stream) {
  if (n > 0 && (!boxes || !trans || !out)) return dc_set_error(nullptr, DC_E_INVALID, "dc_apply_box_transform: null");
  return k_apply_box_transform(nullptr, boxes, trans, out, n, (cudaStream_t)stream);
}
extern "C" int dc_invert_box_transform(const float* anchors, const float* targets, float* out, int n, void* stream) {
  if (n > 0 && (!anchors || !targets || !out)) return dc_set_error(nullptr, DC_E_INVALID, "dc_invert_box_transform: null");
  return k_invert_box_transform(nullptr, anchors, targets, out, n, (cudaStream_t)stream);
}
extern "C" int dc_xcycwh_to_x1y1x2y2(const float* in, float* out, int n, void* stream) {
  if (n > 0 && (!in || !out)) return dc_set_error(nullptr, DC_E_INVALID, "dc_xcycwh_to_x1y1x2y2: null");
  return k_xcycwh_to_x1y1x2y2(nullptr, in, out, n, (cudaStream_t)stream);
}
extern "C" int dc_x1y1x2y2_to_xcycwh(const float* in, float* out, int n, void* stream) {
  if (n > 0 && (!in || !out)) return dc_set_error(nullptr, DC_E_INVALID, "dc_x1y1x2y2_to_xcycwh: null");
  return k_x1y1x2y2_to_xcycwh(nullptr, in, out, n, (cudaStream_t)stream);
}
extern "C" int dc_clip_boxes(const float* boxes, float* out, uint8_t* valid, int n, float x_min, float y_min, float x_max, float y_max, void* stream) {
  if (n > 0 && (!boxes || !out || !valid)) return dc_set_error(nullptr, DC_E_INVALID, "dc_clip_boxes: null");
  return k_clip_boxes(nullptr, boxes, out, valid, n, x_min, y_min, x_max, y_max, (cudaStream_t)stream);
}
extern "C" int dc_objectness_prob(const float* scores, float* prob, int n, void* stream) {
  if (n > 0 && (!scores || !prob)) return dc_set_error(nullptr, DC_E_INVALID, "dc_objectness_prob: null");
  return k_objectness_prob(nullptr, scores, prob, n, (cudaStream_t)stream);
}
extern "C" int dc_box_to_affine(const float* boxes, float* theta, int B, int image_h, int image_w, void* stream) {
  if (B > 0 && (!boxes || !theta)) return dc_set_error(nullptr, DC_E_INVALID, "dc_box_to_affine: null");
  return k_box_to_affine(nullptr, boxes, theta, B, image_h, image_w, (cudaStream_t)stream);
}
extern "C" int dc_box_iou(const float* box1, const float* box2, float* out, int B1, int B2, void* stream) {
  if (B1 > 0 && B2 > 0 && (!box1 || !box2 || !out)) return dc_set_error(nullptr, DC_E_INVALID, "dc_box_iou: null");
  return k_box_iou(nullptr, box1, box2, out, B1, B2, (cudaStream_t)stream);
}

extern "C" int dc_nms(dc_ctx* ctx, const float* boxes_scores, int n, float thresh, int max_boxes, int64_t* idx_out, int32_t* count_out, void* stream) {
  if (!ctx || !count_out || (n > 0 && (!boxes_scores || !idx_out))) return dc_set_error(ctx, DC_E_INVALID, "dc_nms: null argument");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  return k_nms(ctx, boxes_scores, 5, boxes_scores + 4, 5, n, nullptr, thresh, max_boxes, ctx->nms[0], idx_out, nullptr, count_out, (cudaStream_t)stream);
}

__global__ void chw_to_hwc_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int HW) {
  const size_t total = (size_t)C * HW;
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
    const int c = o % C;
    const size_t p = o / C;
    out[o] = in[(size_t)c * HW + p];
  }
}
extern "C" int dc_bilinear_roi_pool(dc_ctx* ctx, const float* feats_chw, int C, int feat_h, int feat_w, const float* boxes, int B, int image_h,
                                    int image_w, int HH, int WW, float* out, int out_layout, void* stream) {
  if (!ctx || !feats_chw || !boxes || !out) return dc_set_error(ctx, DC_E_INVALID, "dc_bilinear_roi_pool: null argument");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)C * feat_h * feat_w;
  DC_REQUIRE(ctx, n <= ctx->scratch_elems, DC_E_INVALID, "dc_bilinear_roi_pool: feature map larger than the ctx scratch (%zu > %zu)", n, ctx->scratch_elems);
  // nn.Transpose({1,2},{2,3}) (BilinearRoiPooling.lua:56): CHW -> HWC, done once
  chw_to_hwc_kernel<<<(unsigned)std::min<size_t>(2048, DC_CEIL_DIV(n, (size_t)256)), 256, 0, st>>>(feats_chw, ctx->scratch_f32, C, feat_h * feat_w);
  DC_CHECK_LAUNCH(ctx);
  // the tap workspace holds roi_taps_boxes * pool_hw^2 records; a larger output grid (the reference's own test uses 8 x 7) simply
  // processes fewer boxes per chunk
  DC_REQUIRE(ctx, HH >= 1 && WW >= 1, DC_E_INVALID, "dc_bilinear_roi_pool: bad output size %dx%d", HH, WW);
  const size_t records = (size_t)ctx->roi_taps_boxes * ctx->cfg.pool_hw * ctx->cfg.pool_hw;
  const int chunk = (int)std::min<size_t>(records / ((size_t)HH * WW), (size_t)ctx->roi_taps_boxes);
  DC_REQUIRE(ctx, chunk >= 1, DC_E_UNSUPPORTED, "dc_bilinear_roi_pool: output grid %dx%d exceeds the tap workspace", HH, WW);
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int nb = std::min(chunk, B - b0);
    int rc = k_roi_pool(ctx, nullptr, nullptr, ctx->scratch_f32, C, feat_h, feat_w, boxes + (size_t)b0 * 4, nullptr, nb, image_h, image_w, HH, WW,
                        ctx->roi_taps, nullptr, nullptr, out + (size_t)b0 * C * HH * WW, out_layout, st);
    if (rc) return rc;
  }
  return DC_OK;
}

// stnbhwd nn.AffineGridGeneratorBHWD(HH, WW):updateOutput (call site modules/BilinearRoiPooling.lua:52)
extern "C" int dc_affine_grid_generator(const float* theta, int B, int HH, int WW, float* grids, void* stream) {
  if (B > 0 && (!theta || !grids)) return dc_set_error(nullptr, DC_E_INVALID, "dc_affine_grid_generator: null");
  if (HH < 2 || WW < 2) return dc_set_error(nullptr, DC_E_INVALID, "dc_affine_grid_generator: output size must be at least 2x2");
  return k_affine_grid(nullptr, theta, B, HH, WW, grids, (cudaStream_t)stream);
}

// nn.BatchBilinearSamplerBHWD:updateOutput (modules/BatchBilinearSamplerBHWD.lua:104-122) on caller-supplied grids
extern "C" int dc_batch_bilinear_sampler(dc_ctx* ctx, const float* feats_hwc, int feat_h, int feat_w, int C, const float* grids, int B, int HH, int WW,
                                         float* out, void* stream) {
  if (!ctx || !feats_hwc || !grids || !out) return dc_set_error(ctx, DC_E_INVALID, "dc_batch_bilinear_sampler: null argument");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  DC_REQUIRE(ctx, HH >= 1 && WW >= 1 && B >= 0, DC_E_INVALID, "dc_batch_bilinear_sampler: bad output size %dx%d", HH, WW);
  const size_t records = (size_t)ctx->roi_taps_boxes * ctx->cfg.pool_hw * ctx->cfg.pool_hw;
  const int chunk = (int)std::min<size_t>(records / ((size_t)HH * WW), (size_t)ctx->roi_taps_boxes);
  DC_REQUIRE(ctx, chunk >= 1, DC_E_UNSUPPORTED, "dc_batch_bilinear_sampler: output grid %dx%d exceeds the tap workspace", HH, WW);
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int nb = std::min(chunk, B - b0);
    int rc = k_bilinear_sampler(ctx, feats_hwc, C, feat_h, feat_w, grids + (size_t)b0 * HH * WW * 2, nb, HH, WW, ctx->roi_taps,
                                out + (size_t)b0 * HH * WW * C, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return DC_OK;
}

// DenseCapModel:extractFeatures (DenseCapModel.lua:285-304): the network without the language model, final NMS on the final boxes,
// boxes and recog_base features (fc7 codes) of the kept rows in score order.
extern "C" int dc_extract_features(dc_ctx* ctx, const float* img, int image_h, int image_w, float* boxes_xcycwh, float* feats, int32_t* num_out,
                                   void* stream) {
  if (!ctx || !img) return dc_set_error(ctx, DC_E_INVALID, "dc_extract_features: null argument");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if ((rc = dc_set_image_size(ctx, image_h, image_w))) return rc;
  ctx->timer.used = 0;
  int fh, fw;
  if ((rc = run_trunk(ctx, img, 1, image_h, image_w, &fh, &fw, st))) return rc;
  ctx->size_consumed = true;
  if ((rc = run_localization(ctx, 1, fh, fw, image_h, image_w, st))) return rc;
  if ((rc = run_recog(ctx, ctx->P, st))) return rc;
  if ((rc = run_final_nms(ctx, 1, st))) return rc;
  if ((rc = join_final_nms(ctx, 1, st))) return rc;
  const int P = ctx->P, fc = ctx->cfg.fc_dim;
  if (boxes_xcycwh && (rc = k_gather_rows_f32(ctx, ctx->boxes_all, 4, ctx->fin_idx32, ctx->fin_count, P, boxes_xcycwh, st))) return rc;
  if (feats && (rc = k_gather_rows_f32(ctx, ctx->codes, fc, ctx->fin_idx32, ctx->fin_count, P, feats, st))) return rc;
  if (num_out) DC_CHECK_CUDA(ctx, cudaMemcpyAsync(num_out, ctx->fin_count, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  return DC_OK;
}

// run_model.lua:66-74 WITH the resize: image.scale(img, image_size) (larger side -> image_size, bilinear) on the device
extern "C" int dc_preprocess_u8_resize(dc_ctx* ctx, const uint8_t* img_u8, int image_h, int image_w, int image_size, float* img_nchw, int* out_h,
                                       int* out_w, void* stream) {
  if (!ctx || !img_u8 || !img_nchw || !out_h || !out_w) return dc_set_error(ctx, DC_E_INVALID, "dc_preprocess_u8_resize: null argument");
  DC_REQUIRE(ctx, image_h > 0 && image_w > 0 && image_size > 0, DC_E_INVALID, "dc_preprocess_u8_resize: bad size");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  const int imax = std::max(image_h, image_w);
  // image/init.lua: height = iheight*size/imax, width = iwidth*size/imax (Lua doubles), truncated by the tensor allocation
  const int oh = (int)((double)image_h * (double)image_size / (double)imax);
  const int ow = (int)((double)image_w * (double)image_size / (double)imax);
  DC_REQUIRE(ctx, oh >= 1 && ow >= 1, DC_E_INVALID, "dc_preprocess_u8_resize: degenerate output %dx%d", oh, ow);
  const size_t need = (size_t)3 * image_h * ow;
  if (ctx->resize_tmp_elems < need) {
    if (ctx->resize_tmp) cudaFree(ctx->resize_tmp);
    ctx->resize_tmp = nullptr; ctx->resize_tmp_elems = 0;
    DC_CHECK_CUDA(ctx, cudaMalloc((void**)&ctx->resize_tmp, need * sizeof(float)));
    ctx->resize_tmp_elems = need;
  }
  *out_h = oh; *out_w = ow;
  return k_image_scale_preprocess(ctx, img_u8, image_h, image_w, oh, ow, ctx->resize_tmp, img_nchw, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// dense test hooks
// ------------------------------------------------------------------------------------------------
struct TmpBufs {
  std::vector<void*> p;
  ~TmpBufs() { for (void* q : p) cudaFree(q); }
  template <typename T> T* get(size_t n) {
    void* q = nullptr;
    if (cudaMalloc(&q, (n ? n : 1) * sizeof(T)) != cudaSuccess) return nullptr;
    p.push_back(q);
    return (T*)q;
  }
};

extern "C" int dc_linear(dc_ctx* ctx, const float* x, const float* w, const float* b, float* y, int M, int N, int K, int relu, void* stream) {
  if (!ctx || !x || !w || !y) return dc_set_error(ctx, DC_E_INVALID, "dc_linear: null argument");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  TmpBufs tmp;
  const size_t ap = align_up((size_t)M * K, 64), bp = align_up((size_t)N * K, 64);
  __half* a = tmp.get<__half>(2 * ap);
  __half* bw = tmp.get<__half>(2 * bp);
  DC_REQUIRE(ctx, a && bw, DC_E_NOMEM, "dc_linear: temporary allocation failed");
  int rc;
  if ((rc = k_rowmajor_to_split(ctx, x, a, a + ap, (size_t)M * K, st))) return rc;
  float sc = 1.f;
  if ((rc = weight_scale(ctx, w, (size_t)N * K, &sc, st))) return rc;
  if ((rc = k_rowmajor_to_split(ctx, w, bw, bw + bp, (size_t)N * K, st, sc))) return rc;
  GemmEpilogue ep;
  ep.acc_scale = 1.f / sc;
  ep.bias = b; ep.relu = relu; ep.out_f32 = y; ep.ldf = N;
  rc = k_tc_gemm(ctx, a, ap, K, bw, bp, K, M, N, K, ep, st);
  cudaError_t e = cudaStreamSynchronize(st);
  if (rc) return rc;
  if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "dc_linear: %s", cudaGetErrorString(e));
  return DC_OK;
}

extern "C" int dc_conv3x3_relu(dc_ctx* ctx, const float* x, const float* w, const float* b, float* y, int N, int Cin, int Cout, int H, int W, int pool,
                               void* stream) {
  if (!ctx || !x || !w || !y) return dc_set_error(ctx, DC_E_INVALID, "dc_conv3x3_relu: null argument");
  DC_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  TmpBufs tmp;
  const size_t xp = align_up((size_t)N * H * W * Cin, 64), wp = align_up((size_t)Cout * 9 * Cin, 64), yp = align_up((size_t)N * H * W * Cout, 64);
  const int PH = (H + 1) / 2, PW = (W + 1) / 2;
  const size_t pp = align_up((size_t)N * PH * PW * Cout, 64);
  __half* xs = tmp.get<__half>(2 * xp);
  __half* ws = tmp.get<__half>(2 * wp);
  __half* ys = tmp.get<__half>(2 * yp);
  __half* ps = tmp.get<__half>(2 * pp);
  DC_REQUIRE(ctx, xs && ws && ys && ps, DC_E_NOMEM, "dc_conv3x3_relu: temporary allocation failed");
  int rc;
  const size_t xn = (size_t)N * H * W * Cin, yn = (size_t)N * H * W * Cout, pn = (size_t)N * PH * PW * Cout;
  if ((rc = k_nchw_to_split_nhwc(ctx, x, xs, xs + xn, N, Cin, H, W, st))) return rc;
  float sc = 1.f;
  if ((rc = weight_scale(ctx, w, (size_t)Cout * Cin * 9, &sc, st))) return rc;
  if ((rc = k_conv_weight_to_split(ctx, w, ws, ws + wp, Cout, Cin, st, sc))) return rc;
  const float isc = 1.f / sc;
  if (pool == 2) {        // fused pool epilogue
    rc = k_tc_conv3x3(ctx, xs, xn, ws, wp, b, isc, ps, ps + pn, N, H, W, Cin, Cout, 1, st);
  } else {
    rc = k_tc_conv3x3(ctx, xs, xn, ws, wp, b, isc, ys, ys + yn, N, H, W, Cin, Cout, 0, st);
    if (!rc && pool) rc = k_maxpool2x2_split(ctx, ys, ys + yn, ps, ps + pn, N, H, W, Cout, 1, st);
  }
  if (!rc) {
    if (pool) rc = k_split_nhwc_to_nchw(ctx, ps, ps + pn, y, N, Cout, PH, PW, st);
    else rc = k_split_nhwc_to_nchw(ctx, ys, ys + yn, y, N, Cout, H, W, st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (rc) return rc;
  if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "dc_conv3x3_relu: %s", cudaGetErrorString(e));
  return dc_check_range(ctx, st);
}
