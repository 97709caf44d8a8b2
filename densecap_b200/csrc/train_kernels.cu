// Training-time forward of nn.LocalizationLayer (densecap/LocalizationLayer.lua:383-526; BASELINE config 5):
// the pieces that the test-time path does not have -- nn.BoxSampler masks + selection
// (densecap/modules/BoxSampler.lua:64-167), BoxSamplerHelper gathers (modules/BoxSamplerHelper.lua:68-129),
// nn.InvertBoxTransform targets + the |t|>10 mask hack (LocalizationLayer.lua:474-489), the objectness
// cross-entropy (modules/OurCrossEntropyCriterion.lua:15-32), SmoothL1 box regression and the RegularizeLayer
// decay term (modules/RegularizeLayer.lua:11-16).
// Sampling: torch.multinomial is irreproducible; exactly like the reference's own debug hook
// (debug_pos_sample_idx / debug_neg_sample_idx, BoxSampler.lua:154-159) the selection is injected, and the
// injection implemented here is deterministic: the first P positives / first M negatives in index order.
#include "common.cuh"

__device__ __forceinline__ float4 t_xcycwh_to_x1y1x2y2(float4 b) {   // box_utils.lua:270-298 (same op order as box_kernels.cu)
  const float hw = __fdiv_rn(__fadd_rn(b.z, -1.f), 2.f);
  const float hh = __fdiv_rn(__fadd_rn(b.w, -1.f), 2.f);
  float4 r;
  r.x = __fadd_rn(__fmul_rn(hw, -1.f), b.x);
  r.z = __fadd_rn(hw, b.x);
  r.y = __fadd_rn(__fmul_rn(hh, -1.f), b.y);
  r.w = __fadd_rn(hh, b.y);
  return r;
}

// All A = k*H*W RPN rows, uncompacted (the training path feeds rpn_out to the sampler as is, LocalizationLayer.lua:406-412)
__global__ void rpn_rows_kernel(const float* __restrict__ head, int ld, int H, int W, int k, const float* __restrict__ anchors,
                                float x0, float y0, float sx, float sy, float4* __restrict__ boxes, float4* __restrict__ anc_out,
                                float4* __restrict__ trans_out, float2* __restrict__ scores_out) {
  const int A = k * H * W;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A) return;
  const int x = r % W, y = (r / W) % H, a = r / (W * H);
  const float* hp = head + (size_t)(y * W + x) * ld;
  float4 anc;
  anc.x = __fadd_rn(__fmul_rn((float)x, sx), x0);
  anc.y = __fadd_rn(__fmul_rn((float)y, sy), y0);
  anc.z = anchors[a];
  anc.w = anchors[k + a];
  const float4 t = make_float4(hp[4 * a + 0], hp[4 * a + 1], hp[4 * a + 2], hp[4 * a + 3]);
  float4 b;                                                  // nn.ApplyBoxTransform
  b.x = __fadd_rn(__fmul_rn(t.x, anc.z), anc.x);
  b.y = __fadd_rn(__fmul_rn(t.y, anc.w), anc.y);
  b.z = __fmul_rn((float)exp((double)t.z), anc.z);
  b.w = __fmul_rn((float)exp((double)t.w), anc.w);
  boxes[r] = b;
  anc_out[r] = anc;
  trans_out[r] = t;
  scores_out[r] = make_float2(hp[4 * k + 2 * a], hp[4 * k + 2 * a + 1]);
}

// ious:max(3): per input box the best IoU and the FIRST target achieving it (BoxSampler.lua:70); then pos/neg masks
// by threshold and the out-of-bounds exclusion (:77-98)
__global__ void sampler_rows_kernel(const float* __restrict__ iou, int A, int B2, const float4* __restrict__ boxes, float lo, float hi,
                                    int has_bounds, float x_min, float y_min, float x_max, float y_max, int32_t* __restrict__ argmax,
                                    uint8_t* __restrict__ pos, uint8_t* __restrict__ neg) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A) return;
  const float* row = iou + (size_t)r * B2;
  float best = row[0];
  int bi = 0;
  for (int j = 1; j < B2; ++j) {
    const float v = row[j];
    if (v > best) { best = v; bi = j; }
  }
  argmax[r] = bi;
  bool p = best > hi, n = best < lo;
  if (has_bounds) {
    const float4 c = t_xcycwh_to_x1y1x2y2(boxes[r]);
    const bool oob = (c.x < x_min) || (c.y < y_min) || (c.z > x_max) || (c.w > y_max);
    if (oob) { p = false; n = false; }
  }
  pos[r] = p ? 1 : 0;
  neg[r] = n ? 1 : 0;
}

// ious:max(2): per target the FIRST input box with the best IoU (BoxSampler.lua:73); one block per target
__global__ void sampler_cols_kernel(const float* __restrict__ iou, int A, int B2, int32_t* __restrict__ target_idx) {
  const int j = blockIdx.x;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int r = threadIdx.x; r < A; r += blockDim.x) {
    const float v = iou[(size_t)r * B2 + j];
    if (v > best || (v == best && r < bi)) { best = v; bi = r; }
  }
  __shared__ float sv[256];
  __shared__ int si[256];
  sv[threadIdx.x] = best;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      const float v = sv[threadIdx.x + o];
      const int i = si[threadIdx.x + o];
      if (v > sv[threadIdx.x] || (v == sv[threadIdx.x] && i < si[threadIdx.x])) { sv[threadIdx.x] = v; si[threadIdx.x] = i; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) target_idx[j] = si[0];
}

// pos_mask:scatter(2, target_idx, 1); neg_mask:scatter(2, target_idx, 0)   (BoxSampler.lua:104-105)
__global__ void sampler_force_kernel(const int32_t* __restrict__ target_idx, int B2, uint8_t* __restrict__ pos, uint8_t* __restrict__ neg) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= B2) return;
  pos[target_idx[j]] = 1;
  neg[target_idx[j]] = 0;
}

// Ordered compaction of the first `cap` set entries of a mask by ONE block of 1024 threads; returns the total count.
__device__ int block_first_nonzero(const uint8_t* __restrict__ mask, bool invert_of, const uint8_t* __restrict__ other, int A, int cap,
                                   int32_t* __restrict__ out) {
  __shared__ int warp_cnt[32];
  __shared__ int running;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  for (int base = 0; base < A; base += blockDim.x) {
    const int r = base + threadIdx.x;
    bool f = false;
    if (r < A) f = invert_of ? (other[r] == 0) : (mask[r] != 0);
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = running;
    for (int w = 0; w < warp; ++w) off += warp_cnt[w];
    off += __popc(bal & ((1u << lane) - 1u));
    if (f && off < cap) out[off] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
      int c = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) c += warp_cnt[w];
      running += c;
    }
    __syncthreads();
  }
  return running;
}

// counts: [0] total_pos, [1] total_neg, [2] P, [3] M
__global__ void __launch_bounds__(1024)
sampler_select_kernel(const uint8_t* __restrict__ pos, const uint8_t* __restrict__ neg, int A, int batch_size, int32_t* __restrict__ pos_idx,
                      int32_t* __restrict__ neg_idx, int32_t* __restrict__ counts) {
  const int total_pos = block_first_nonzero(pos, false, nullptr, A, batch_size, pos_idx);
  int total_neg = block_first_nonzero(neg, false, nullptr, A, batch_size, neg_idx);
  if (total_neg == 0) total_neg = block_first_nonzero(nullptr, true, pos, A, batch_size, neg_idx);   // neg = not pos (BoxSampler.lua:111-128)
  const int P = min(batch_size / 2, total_pos);
  const int M = batch_size - P;
  __syncthreads();
  // too few negatives: sampled with replacement -> deterministic cyclic reuse
  const int have = min(total_neg, batch_size);
  for (int j = have + threadIdx.x; j < M; j += blockDim.x) neg_idx[j] = have > 0 ? neg_idx[j % have] : 0;
  if (threadIdx.x == 0) { counts[0] = total_pos; counts[1] = total_neg; counts[2] = P; counts[3] = M; }
}

// BoxSamplerHelper gathers + roi_boxes = cat(pos_boxes, neg_boxes) (LocalizationLayer.lua:444-446)
__global__ void sampler_gather_kernel(const int32_t* __restrict__ counts, const int32_t* __restrict__ pos_idx, const int32_t* __restrict__ neg_idx,
                                      const int32_t* __restrict__ argmax, const float4* __restrict__ boxes, const float4* __restrict__ anc,
                                      const float4* __restrict__ trans, const float2* __restrict__ scores, const float4* __restrict__ gt,
                                      int batch_size, float4* __restrict__ roi_boxes, float4* __restrict__ pos_anchors,
                                      float4* __restrict__ pos_trans, float2* __restrict__ pos_scores, float2* __restrict__ neg_scores,
                                      float4* __restrict__ pos_target_boxes, int64_t* __restrict__ pos_target_idx,
                                      int64_t* __restrict__ pos_input_idx, int64_t* __restrict__ neg_input_idx) {
  const int P = counts[2], M = counts[3];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch_size) return;
  if (i < P) {
    const int r = pos_idx[i];
    roi_boxes[i] = boxes[r];
    pos_anchors[i] = anc[r];
    pos_trans[i] = trans[r];
    pos_scores[i] = scores[r];
    const int t = argmax[r];
    pos_target_boxes[i] = gt[t];
    if (pos_target_idx) pos_target_idx[i] = t + 1;
    if (pos_input_idx) pos_input_idx[i] = r + 1;
  } else if (i - P < M) {
    const int r = neg_idx[i - P];
    roi_boxes[i] = boxes[r];
    neg_scores[i - P] = scores[r];
    if (neg_input_idx) neg_input_idx[i - P] = r + 1;
  }
}

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double red[32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x < 32) {
    s = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  __syncthreads();
  return s;   // valid in warp 0
}

// losses: [0] obj_loss_pos, [1] obj_loss_neg, [2] box_reg_loss, [3] box_decay_loss, [4] total; grads (optional) of the three criterions
__global__ void __launch_bounds__(1024)
loc_losses_kernel(const int32_t* __restrict__ counts, const float2* __restrict__ pos_scores, const float2* __restrict__ neg_scores,
                  const float4* __restrict__ pos_anchors, const float4* __restrict__ pos_target_boxes, float4* __restrict__ pos_trans,
                  const float4* __restrict__ all_trans, int A, float w_obj, float w_box, float w_decay, float* __restrict__ losses,
                  float4* __restrict__ pos_trans_targets, int32_t* __restrict__ masked_out) {
  const int P = counts[2], M = counts[3];
  // OurCrossEntropyCriterion = LogSoftMax + ClassNLL (size-averaged); target 1 for positives, 2 for negatives
  double lp = 0.0, ln = 0.0;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const double a = pos_scores[i].x, b = pos_scores[i].y, m = fmax(a, b);
    lp += -(a - m - log(exp(a - m) + exp(b - m)));
  }
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const double a = neg_scores[i].x, b = neg_scores[i].y, m = fmax(a, b);
    ln += -(b - m - log(exp(a - m) + exp(b - m)));
  }
  // InvertBoxTransform targets, |t| > 10 mask hack (zeroes prediction AND target in place), SmoothL1 (size-averaged over P*4)
  double lb = 0.0;
  int masked = 0;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const float4 an = pos_anchors[i], tb = pos_target_boxes[i];
    float4 t;
    t.x = __fdiv_rn(__fsub_rn(tb.x, an.x), an.z);
    t.y = __fdiv_rn(__fsub_rn(tb.y, an.y), an.w);
    t.z = (float)log((double)__fdiv_rn(tb.z, an.z));
    t.w = (float)log((double)__fdiv_rn(tb.w, an.w));
    float4 p = pos_trans[i];
    const float mx = fmaxf(fmaxf(fabsf(t.x), fabsf(t.y)), fmaxf(fabsf(t.z), fabsf(t.w)));
    if (mx > 10.f) {
      t = make_float4(0.f, 0.f, 0.f, 0.f);
      p = make_float4(0.f, 0.f, 0.f, 0.f);
      pos_trans[i] = p;
      ++masked;
    }
    if (pos_trans_targets) pos_trans_targets[i] = t;
    const double d[4] = {(double)p.x - t.x, (double)p.y - t.y, (double)p.z - t.z, (double)p.w - t.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double a = fabs(d[c]);
      lb += (a < 1.0) ? 0.5 * d[c] * d[c] : a - 0.5;
    }
  }
  double dec = 0.0;
  if (w_decay != 0.f) {
    for (int i = threadIdx.x; i < A; i += blockDim.x) {
      const float4 t = all_trans[i];
      dec += (double)t.x * t.x + (double)t.y * t.y + (double)t.z * t.z + (double)t.w * t.w;
    }
  }
  const double slp = block_sum(lp), sln = block_sum(ln), slb = block_sum(lb), sdec = block_sum(dec), smask = block_sum((double)masked);
  if (threadIdx.x == 0) {
    const float l0 = w_obj * (float)(slp / (P > 0 ? P : 1));
    const float l1 = w_obj * (float)(sln / (M > 0 ? M : 1));
    const float l2 = w_box * (float)(slb / (P > 0 ? 4 * P : 1));
    const float l3 = (float)(0.5 * (double)w_decay * sdec);
    losses[0] = l0; losses[1] = l1; losses[2] = l2; losses[3] = l3;
    losses[4] = (float)((double)l0 + l1 + l2 + l3);
    if (masked_out) *masked_out = (int)smask;
  }
}

// ---------------------------------------------------------------------------------------------------------
struct TrainBuffers {
  float4 *boxes, *anc, *trans, *gt, *roi_boxes, *pos_anchors, *pos_trans, *pos_target_boxes, *pos_trans_targets;
  float2 *scores, *pos_scores, *neg_scores;
  float* iou;
  int32_t *argmax, *target_idx, *pos_idx, *neg_idx, *counts, *masked;
  uint8_t *pos, *neg;
};

int k_loc_train_sampler(dc_ctx* ctx, const float* head, int ld, int H, int W, int k, const float* anchors_dev, float x0, float y0, float sx,
                        float sy, const float* gt_boxes, int B2, float lo, float hi, int batch_size, int has_bounds, float img_w, float img_h,
                        void* tb_, int64_t* pos_target_idx, int64_t* pos_input_idx, int64_t* neg_input_idx, cudaStream_t st) {
  TrainBuffers& tb = *reinterpret_cast<TrainBuffers*>(tb_);
  const int A = k * H * W;
  rpn_rows_kernel<<<DC_CEIL_DIV(A, 256), 256, 0, st>>>(head, ld, H, W, k, anchors_dev, x0, y0, sx, sy, tb.boxes, tb.anc, tb.trans, tb.scores);
  DC_CHECK_LAUNCH(ctx);
  int rc = k_box_iou(ctx, (const float*)tb.boxes, gt_boxes, tb.iou, A, B2, st);
  if (rc) return rc;
  sampler_rows_kernel<<<DC_CEIL_DIV(A, 256), 256, 0, st>>>(tb.iou, A, B2, tb.boxes, lo, hi, has_bounds, 1.f, 1.f, img_w, img_h, tb.argmax, tb.pos,
                                                           tb.neg);
  DC_CHECK_LAUNCH(ctx);
  sampler_cols_kernel<<<B2, 256, 0, st>>>(tb.iou, A, B2, tb.target_idx);
  DC_CHECK_LAUNCH(ctx);
  sampler_force_kernel<<<DC_CEIL_DIV(B2, 128), 128, 0, st>>>(tb.target_idx, B2, tb.pos, tb.neg);
  DC_CHECK_LAUNCH(ctx);
  sampler_select_kernel<<<1, 1024, 0, st>>>(tb.pos, tb.neg, A, batch_size, tb.pos_idx, tb.neg_idx, tb.counts);
  DC_CHECK_LAUNCH(ctx);
  sampler_gather_kernel<<<DC_CEIL_DIV(batch_size, 128), 128, 0, st>>>(tb.counts, tb.pos_idx, tb.neg_idx, tb.argmax, tb.boxes, tb.anc, tb.trans,
                                                                    tb.scores, (const float4*)gt_boxes, batch_size, tb.roi_boxes, tb.pos_anchors,
                                                                    tb.pos_trans, tb.pos_scores, tb.neg_scores, tb.pos_target_boxes,
                                                                    pos_target_idx, pos_input_idx, neg_input_idx);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

int k_loc_train_losses(dc_ctx* ctx, void* tb_, int A, float w_obj, float w_box, float w_decay, float* losses, cudaStream_t st) {
  TrainBuffers& tb = *reinterpret_cast<TrainBuffers*>(tb_);
  loc_losses_kernel<<<1, 1024, 0, st>>>(tb.counts, tb.pos_scores, tb.neg_scores, tb.pos_anchors, tb.pos_target_boxes, tb.pos_trans, tb.trans, A, w_obj,
                                        w_box, w_decay, losses, tb.pos_trans_targets, tb.masked);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

size_t train_buffers_struct_bytes() { return sizeof(TrainBuffers); }
// carve the buffers out of one allocation; returns bytes needed when base == nullptr
size_t train_buffers_carve(void* tb_, char* base, int A, int B2max, int batch_size) {
  TrainBuffers dummy;
  TrainBuffers& tb = tb_ ? *reinterpret_cast<TrainBuffers*>(tb_) : dummy;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += (bytes + 255) / 256 * 256; return p; };
  tb.boxes = (float4*)take((size_t)A * 16); tb.anc = (float4*)take((size_t)A * 16); tb.trans = (float4*)take((size_t)A * 16);
  tb.scores = (float2*)take((size_t)A * 8);
  tb.iou = (float*)take((size_t)A * B2max * 4);
  tb.argmax = (int32_t*)take((size_t)A * 4); tb.target_idx = (int32_t*)take((size_t)B2max * 4);
  tb.pos = (uint8_t*)take(A); tb.neg = (uint8_t*)take(A);
  tb.pos_idx = (int32_t*)take((size_t)batch_size * 4); tb.neg_idx = (int32_t*)take((size_t)batch_size * 4);
  tb.counts = (int32_t*)take(16); tb.masked = (int32_t*)take(4);
  tb.gt = (float4*)take((size_t)B2max * 16);
  tb.roi_boxes = (float4*)take((size_t)batch_size * 16); tb.pos_anchors = (float4*)take((size_t)batch_size * 16);
  tb.pos_trans = (float4*)take((size_t)batch_size * 16); tb.pos_target_boxes = (float4*)take((size_t)batch_size * 16);
  tb.pos_trans_targets = (float4*)take((size_t)batch_size * 16);
  tb.pos_scores = (float2*)take((size_t)batch_size * 8); tb.neg_scores = (float2*)take((size_t)batch_size * 8);
  return off;
}
float* train_buffers_roi_boxes(void* tb_) { return (float*)reinterpret_cast<TrainBuffers*>(tb_)->roi_boxes; }
float* train_buffers_gt(void* tb_) { return (float*)reinterpret_cast<TrainBuffers*>(tb_)->gt; }
float* train_buffers_pos_target_boxes(void* tb_) { return (float*)reinterpret_cast<TrainBuffers*>(tb_)->pos_target_boxes; }
int32_t* train_buffers_counts(void* tb_) { return reinterpret_cast<TrainBuffers*>(tb_)->counts; }
