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
#include <algorithm>

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
  const int total = running;       // every thread reads the final value ...
  __syncthreads();                 // ... before the next call of this function can reset it
  return total;
}

// ---- uniform sampling (torch.multinomial over a constant probability vector, BoxSampler.lua:138-152) --------------------------
// A keyed bijection of the row index gives every candidate a UNIQUE pseudo-random 32-bit key; "n candidates without replacement"
// = the n candidates with the smallest keys.  The n-th smallest key is found by bisection on the key value (one block-wide count
// per bit), then the selected rows are compacted in index order.  With replacement (only when there are fewer negatives than
// requested, like the reference) draw j takes candidate hash(j) mod total.
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {       // murmur3 finaliser: a bijection on 32-bit words
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}
__device__ int block_count_le(const uint8_t* __restrict__ mask, bool invert_of, const uint8_t* __restrict__ other, int A, uint32_t salt, uint32_t thr) {
  __shared__ int warp_cnt[32];
  __shared__ int total;
  int c = 0;
  for (int r = threadIdx.x; r < A; r += blockDim.x) {
    const bool f = invert_of ? (other[r] == 0) : (mask[r] != 0);
    c += (f && fmix32((uint32_t)r ^ salt) <= thr) ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += warp_cnt[w];
    total = t;
  }
  __syncthreads();
  return total;
}
// ordered compaction of the candidates whose key is <= thr (at most cap of them); returns nothing
__device__ void block_compact_le(const uint8_t* __restrict__ mask, bool invert_of, const uint8_t* __restrict__ other, int A, uint32_t salt, uint32_t thr,
                                 int cap, int32_t* __restrict__ out) {
  __shared__ int warp_cnt[32];
  __shared__ int running;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  for (int base = 0; base < A; base += blockDim.x) {
    const int r = base + threadIdx.x;
    bool f = false;
    if (r < A) f = (invert_of ? (other[r] == 0) : (mask[r] != 0)) && fmix32((uint32_t)r ^ salt) <= thr;
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
}
// n <= total candidates, n >= 1: smallest key value thr with #(key <= thr) >= n  (keys are unique, so the count is exactly n)
__device__ uint32_t block_nth_key(const uint8_t* __restrict__ mask, bool invert_of, const uint8_t* __restrict__ other, int A, uint32_t salt, int n) {
  uint32_t lo = 0u, hi = 0xFFFFFFFFu;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (block_count_le(mask, invert_of, other, A, salt, mid) >= n) hi = mid; else lo = mid + 1u;
  }
  return lo;
}

// counts: [0] total_pos, [1] total_neg, [2] P, [3] M.  mode 0: the first P positives / first M negatives in index order (the
// reference's debug_pos_sample_idx / debug_neg_sample_idx injection, BoxSampler.lua:154-159); mode 1: uniform random (seeded).
__global__ void __launch_bounds__(1024)
sampler_select_kernel(const uint8_t* __restrict__ pos, const uint8_t* __restrict__ neg, int A, int batch_size, int mode, uint32_t salt,
                      int32_t* __restrict__ pos_idx, int32_t* __restrict__ neg_idx, int32_t* __restrict__ counts) {
  __shared__ int32_t s_list[1024];
  const int total_pos = block_first_nonzero(pos, false, nullptr, A, batch_size, pos_idx);
  int total_neg = block_first_nonzero(neg, false, nullptr, A, batch_size, neg_idx);
  const bool neg_is_not_pos = (total_neg == 0);
  if (neg_is_not_pos) total_neg = block_first_nonzero(nullptr, true, pos, A, batch_size, neg_idx);   // neg = not pos (BoxSampler.lua:111-128)
  const int P = min(batch_size / 2, total_pos);
  const int M = batch_size - P;
  __syncthreads();
  if (mode == 1) {
    const uint32_t salt_p = fmix32(salt ^ 0x9e3779b9u), salt_n = fmix32(salt ^ 0x7f4a7c15u);
    if (P > 0 && total_pos > P) {                                     // all positives are taken when total_pos <= P
      const uint32_t thr = block_nth_key(pos, false, nullptr, A, salt_p, P);
      block_compact_le(pos, false, nullptr, A, salt_p, thr, P, pos_idx);
    }
    if (total_neg > M) {                                              // without replacement
      const uint32_t thr = block_nth_key(neg, neg_is_not_pos, pos, A, salt_n, M);
      block_compact_le(neg, neg_is_not_pos, pos, A, salt_n, thr, M, neg_idx);
    } else if (total_neg > 0 && total_neg < M) {                      // with replacement: neg_idx[0, total_neg) holds every negative
      __syncthreads();
      for (int j = threadIdx.x; j < total_neg; j += blockDim.x) s_list[j] = neg_idx[j];
      __syncthreads();
      for (int j = threadIdx.x; j < M; j += blockDim.x) neg_idx[j] = s_list[fmix32((uint32_t)j * 0x9e3779b1u ^ salt_n) % (uint32_t)total_neg];
    }
  } else {
    // too few negatives: sampled with replacement -> deterministic cyclic reuse
    const int have = min(total_neg, batch_size);
    __syncthreads();
    for (int j = have + threadIdx.x; j < M; j += blockDim.x) neg_idx[j] = have > 0 ? neg_idx[j % have] : 0;
  }
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
                  float4* __restrict__ pos_trans_targets, int32_t* __restrict__ masked_out, float4* __restrict__ g_pos_trans,
                  float2* __restrict__ g_pos_scores, float2* __restrict__ g_neg_scores) {
  const int P = counts[2], M = counts[3];
  // OurCrossEntropyCriterion = LogSoftMax + ClassNLL (size-averaged); target 1 for positives, 2 for negatives
  double lp = 0.0, ln = 0.0;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const double a = pos_scores[i].x, b = pos_scores[i].y, m = fmax(a, b);
    const double lse = log(exp(a - m) + exp(b - m));
    lp += -(a - m - lse);
    // gradient of the size-averaged NLL: (softmax - onehot(class 1)) / P, then scaled by the loss weight in fp32
    g_pos_scores[i] = make_float2((float)((exp(a - m - lse) - 1.0) / P) * w_obj, (float)(exp(b - m - lse) / P) * w_obj);
  }
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const double a = neg_scores[i].x, b = neg_scores[i].y, m = fmax(a, b);
    const double lse = log(exp(a - m) + exp(b - m));
    ln += -(b - m - lse);
    g_neg_scores[i] = make_float2((float)(exp(a - m - lse) / M) * w_obj, (float)((exp(b - m - lse) - 1.0) / M) * w_obj);
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
    float gq[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double a = fabs(d[c]);
      lb += (a < 1.0) ? 0.5 * d[c] * d[c] : a - 0.5;
      gq[c] = (float)(fmin(fmax(d[c], -1.0), 1.0) / (4.0 * P)) * w_box;        // SmoothL1 gradient (size-averaged), weighted
    }
    g_pos_trans[i] = make_float4(gq[0], gq[1], gq[2], gq[3]);
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
  float4* g_pos_trans;                     // criterion gradients, already scaled by the loss weights (LocalizationLayer.lua:553-570)
  float2 *g_pos_scores, *g_neg_scores;
  float* iou;
  int32_t *argmax, *target_idx, *pos_idx, *neg_idx, *counts, *masked;
  uint8_t *pos, *neg;
};

int k_loc_train_sampler(dc_ctx* ctx, const float* head, int ld, int H, int W, int k, const float* anchors_dev, float x0, float y0, float sx,
                        float sy, const float* gt_boxes, int B2, float lo, float hi, int batch_size, int has_bounds, float img_w, float img_h,
                        void* tb_, int64_t* pos_target_idx, int64_t* pos_input_idx, int64_t* neg_input_idx, int sampler_mode, uint32_t sampler_salt,
                        cudaStream_t st) {
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
  sampler_select_kernel<<<1, 1024, 0, st>>>(tb.pos, tb.neg, A, batch_size, sampler_mode, sampler_salt, tb.pos_idx, tb.neg_idx, tb.counts);
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
                                        w_box, w_decay, losses, tb.pos_trans_targets, tb.masked, tb.g_pos_trans, tb.g_pos_scores, tb.g_neg_scores);
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
  tb.g_pos_trans = (float4*)take((size_t)batch_size * 16);
  tb.g_pos_scores = (float2*)take((size_t)batch_size * 8); tb.g_neg_scores = (float2*)take((size_t)batch_size * 8);
  return off;
}
float* train_buffers_roi_boxes(void* tb_) { return (float*)reinterpret_cast<TrainBuffers*>(tb_)->roi_boxes; }
float* train_buffers_gt(void* tb_) { return (float*)reinterpret_cast<TrainBuffers*>(tb_)->gt; }
float* train_buffers_pos_target_boxes(void* tb_) { return (float*)reinterpret_cast<TrainBuffers*>(tb_)->pos_target_boxes; }
int32_t* train_buffers_counts(void* tb_) { return reinterpret_cast<TrainBuffers*>(tb_)->counts; }

// =========================================================================================================
// Backward of the training-time LocalizationLayer (LocalizationLayer.lua:529-605)
// =========================================================================================================
// BoxSamplerHelper backward (modules/BoxSamplerHelper.lua:142-182): indexCopy of the sampled gradients into zeroed (A, .) tensors.
// Negatives drawn with replacement repeat an index; indexCopy keeps the LAST copy, so only the last occurrence writes.
// grad_roi_boxes_in: caller's gradOutput[2] (S,4); grad_boxes_pool: RoI-pooling gradient w.r.t. the boxes (S,4) -- summed (:577).
__global__ void bwd_scatter_kernel(const int32_t* __restrict__ counts, const int32_t* __restrict__ pos_idx, const int32_t* __restrict__ neg_idx,
                                   const float4* __restrict__ grad_roi_boxes_in, const float4* __restrict__ grad_boxes_pool,
                                   const float4* __restrict__ g_pos_trans, const float2* __restrict__ g_pos_scores,
                                   const float2* __restrict__ g_neg_scores, int batch_size, float4* __restrict__ g_boxes_all,
                                   float4* __restrict__ g_trans_all, float2* __restrict__ g_scores_all) {
  const int P = counts[2], M = counts[3];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch_size) return;
  auto total = [&](int row) {
    const float4 a = grad_roi_boxes_in[row], b = grad_boxes_pool[row];
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  };
  if (i < P) {
    const int r = pos_idx[i];
    g_boxes_all[r] = total(i);
    g_trans_all[r] = g_pos_trans[i];
    g_scores_all[r] = g_pos_scores[i];
  } else if (i - P < M) {
    const int j = i - P, r = neg_idx[j];
    for (int j2 = j + 1; j2 < M; ++j2)
      if (neg_idx[j2] == r) return;
    g_boxes_all[r] = total(i);
    g_scores_all[r] = g_neg_scores[j];
  }
}

// Gradient w.r.t. the RPN head outputs, one row (anchor a, cell y,x) per thread:
//   ApplyBoxTransform backward (modules/ApplyBoxTransform.lua:93-128) + the Identity branch of the ConcatTable,
//   RegularizeLayer backward (modules/RegularizeLayer.lua:18-22: w * input + gradOutput),
//   ReshapeBoxFeatures backward (modules/ReshapeBoxFeatures.lua:36-44) = channel 4a+d / 4k+2a+d of cell (y,x).
// head_grad: (cells, ldh) fp32, columns [0,4k) box transform, [4k,6k) scores, rest (padding up to ldh) pre-zeroed.
__global__ void bwd_head_grad_kernel(int A, int H, int W, int k, const float4* __restrict__ anc, const float4* __restrict__ trans,
                                     const float4* __restrict__ boxes, const float4* __restrict__ g_boxes_all,
                                     const float4* __restrict__ g_trans_all, const float2* __restrict__ g_scores_all, float w_decay,
                                     float* __restrict__ head_grad, int ldh) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A) return;
  const int x = r % W, y = (r / W) % H, a = r / (W * H);
  const float4 an = anc[r], t = trans[r], o = boxes[r], g = g_boxes_all[r], gt0 = g_trans_all[r];
  float4 gt;
  gt.x = __fadd_rn(gt0.x, __fmul_rn(g.x, an.z));
  gt.y = __fadd_rn(gt0.y, __fmul_rn(g.y, an.w));
  gt.z = __fadd_rn(gt0.z, __fmul_rn(o.z, g.z));
  gt.w = __fadd_rn(gt0.w, __fmul_rn(o.w, g.w));
  gt.x = __fadd_rn(__fmul_rn(t.x, w_decay), gt.x);
  gt.y = __fadd_rn(__fmul_rn(t.y, w_decay), gt.y);
  gt.z = __fadd_rn(__fmul_rn(t.z, w_decay), gt.z);
  gt.w = __fadd_rn(__fmul_rn(t.w, w_decay), gt.w);
  float* row = head_grad + (size_t)(y * W + x) * ldh;
  row[4 * a + 0] = gt.x; row[4 * a + 1] = gt.y; row[4 * a + 2] = gt.z; row[4 * a + 3] = gt.w;
  const float2 gs = g_scores_all[r];
  row[4 * k + 2 * a + 0] = gs.x; row[4 * k + 2 * a + 1] = gs.y;
}

// ReLU backward of the RPN hidden layer + re-split for the 3x3 backward-data convolution: gh = gh * (hidden > 0)
__global__ void bwd_relu_mask_kernel(float* __restrict__ gh, const __half* __restrict__ hid_hi, const __half* __restrict__ hid_lo, size_t n,
                                     __half* __restrict__ out_hi, __half* __restrict__ out_lo, float* __restrict__ hid_f32) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float h = __half2float(hid_hi[i]) + __half2float(hid_lo[i]);
    const float g = h > 0.f ? gh[i] : 0.f;
    gh[i] = g;
    if (hid_f32) hid_f32[i] = h;
    __half a, b;
    split_f32(g, a, b);
    out_hi[i] = a; out_lo[i] = b;
  }
}

// grad_cnn_features (1,C,H,W) = RoI-pooling gradient (HWC fp32) + RPN gradient (split NHWC planes)   (:578, :601)
__global__ void bwd_combine_kernel(const float* __restrict__ g_roi_hwc, const __half* __restrict__ g_rpn_hi, const __half* __restrict__ g_rpn_lo,
                                   int C, int cells, float* __restrict__ out_chw) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, p0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int p = p0 + j, c = c0 + threadIdx.x;
    float v = 0.f;
    if (p < cells && c < C) {
      const size_t o = (size_t)p * C + c;
      v = g_roi_hwc[o] + (__half2float(g_rpn_hi[o]) + __half2float(g_rpn_lo[o]));
    }
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, p = p0 + threadIdx.x;
    if (p < cells && c < C) out_chw[(size_t)c * cells + p] = tile[threadIdx.x][j];
  }
}

// (K, C, npos) -> (K, npos, C)
__global__ void bchw_to_bhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int npos) {
  const float* src = in + (size_t)blockIdx.x * C * npos;
  float* dst = out + (size_t)blockIdx.x * C * npos;
  for (int i = threadIdx.x; i < C * npos; i += blockDim.x) {
    const int p = i / C, c = i % C;
    dst[i] = src[(size_t)c * npos + p];
  }
}

// accGradParameters of the RPN convolutions: C[m][n] (+)= sum over cells of A[cell][m] * B[shift_tap(cell)][n]
//   A (cells, Mdim) and B (cells, Ndim) fp32 row-major over the H x W map; tap = blockIdx.z shifts B by (r-1, s-1) with zero
//   padding (taps == 1: no shift).  Output element (m, n, tap) at out[(m*Ndim + n)*taps + tap]  (= (Cout, Cin, 3, 3) layout).
constexpr int WG_T = 64, WG_K = 16;
__global__ void __launch_bounds__(256)
bwd_wgrad_kernel(const float* __restrict__ A, int Mdim, const float* __restrict__ B, int Ndim, int H, int W, int taps, float* __restrict__ out) {
  __shared__ float sa[WG_K][WG_T + 1], sb[WG_K][WG_T + 1];
  const int m0 = blockIdx.y * WG_T, n0 = blockIdx.x * WG_T, tap = blockIdx.z;
  const int dr = taps == 9 ? tap / 3 - 1 : 0, ds = taps == 9 ? tap % 3 - 1 : 0;
  const int cells = H * W;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;           // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4] = {};
  for (int k0 = 0; k0 < cells; k0 += WG_K) {
    for (int e = threadIdx.x; e < WG_K * WG_T; e += 256) {
      const int kk = e / WG_T, j = e % WG_T, cell = k0 + kk;
      float va = 0.f, vb = 0.f;
      if (cell < cells) {
        if (m0 + j < Mdim) va = A[(size_t)cell * Mdim + m0 + j];
        const int y = cell / W + dr, x = cell % W + ds;
        if (n0 + j < Ndim && y >= 0 && y < H && x >= 0 && x < W) vb = B[(size_t)(y * W + x) * Ndim + n0 + j];
      }
      sa[kk][j] = va; sb[kk][j] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < WG_K; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sa[kk][ty * 4 + i]; b[i] = sb[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < Mdim && n < Ndim) out[((size_t)m * Ndim + n) * taps + tap] += acc[i][j];
    }
}
// gradBias[m] += sum over cells of A[cell][m]
__global__ void bwd_colsum_kernel(const float* __restrict__ A, int Mdim, int cells, float* __restrict__ out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mdim) return;
  double s = 0.0;
  for (int c = 0; c < cells; ++c) s += A[(size_t)c * Mdim + m];
  out[m] += (float)s;
}

// weight re-layouts for the backward-data passes
// (F, C, 3, 3) -> (C, F, 3, 3) with both taps flipped: full-correlation kernel of the transposed convolution
__global__ void conv_weight_flip_transpose_kernel(const float* __restrict__ w, float* __restrict__ out, int F, int C) {
  const size_t n = (size_t)F * C * 9;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int tap = i % 9;
    const int f = (i / 9) % F;
    const int c = i / ((size_t)9 * F);
    out[i] = w[((size_t)f * C + c) * 9 + (8 - tap)];
  }
}
// rows (R, F) row-major -> columns [col0, col0+R) of (F, ld) row-major
__global__ void transpose_into_kernel(const float* __restrict__ w, int R, int F, float* __restrict__ out, int ld, int col0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * F) return;
  const int r = i / F, f = i % F;
  out[(size_t)f * ld + col0 + r] = w[i];
}

int k_conv_weight_flip_transpose(dc_ctx* ctx, const float* w, float* out, int F, int C, cudaStream_t st) {
  conv_weight_flip_transpose_kernel<<<256, 256, 0, st>>>(w, out, F, C);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
int k_transpose_into(dc_ctx* ctx, const float* w, int R, int F, float* out, int ld, int col0, cudaStream_t st) {
  transpose_into_kernel<<<DC_CEIL_DIV(R * F, 256), 256, 0, st>>>(w, R, F, out, ld, col0);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
int k_bchw_to_bhwc_f32(dc_ctx* ctx, const float* in, float* out, int K, int C, int npos, cudaStream_t st) {
  if (K <= 0) return DC_OK;
  bchw_to_bhwc_kernel<<<K, 256, 0, st>>>(in, out, C, npos);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

int k_loc_bwd_scatter(dc_ctx* ctx, void* tb_, int A, int batch_size, const float* grad_roi_boxes_in, const float* grad_boxes_pool,
                      float* g_boxes_all, float* g_trans_all, float* g_scores_all, cudaStream_t st) {
  TrainBuffers& tb = *reinterpret_cast<TrainBuffers*>(tb_);
  DC_CHECK_CUDA(ctx, cudaMemsetAsync(g_boxes_all, 0, (size_t)A * 16, st));
  DC_CHECK_CUDA(ctx, cudaMemsetAsync(g_trans_all, 0, (size_t)A * 16, st));
  DC_CHECK_CUDA(ctx, cudaMemsetAsync(g_scores_all, 0, (size_t)A * 8, st));
  bwd_scatter_kernel<<<DC_CEIL_DIV(batch_size, 128), 128, 0, st>>>(tb.counts, tb.pos_idx, tb.neg_idx, (const float4*)grad_roi_boxes_in,
                                                                 (const float4*)grad_boxes_pool, tb.g_pos_trans, tb.g_pos_scores, tb.g_neg_scores,
                                                                 batch_size, (float4*)g_boxes_all, (float4*)g_trans_all, (float2*)g_scores_all);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
int k_loc_bwd_head_grad(dc_ctx* ctx, void* tb_, int H, int W, int k, const float* g_boxes_all, const float* g_trans_all, const float* g_scores_all,
                        float w_decay, float* head_grad, int ldh, cudaStream_t st) {
  TrainBuffers& tb = *reinterpret_cast<TrainBuffers*>(tb_);
  const int A = k * H * W;
  DC_CHECK_CUDA(ctx, cudaMemsetAsync(head_grad, 0, (size_t)H * W * ldh * sizeof(float), st));
  bwd_head_grad_kernel<<<DC_CEIL_DIV(A, 256), 256, 0, st>>>(A, H, W, k, tb.anc, tb.trans, tb.boxes, (const float4*)g_boxes_all,
                                                          (const float4*)g_trans_all, (const float2*)g_scores_all, w_decay, head_grad, ldh);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
int k_bwd_relu_mask(dc_ctx* ctx, float* gh, const __half* hid_hi, const __half* hid_lo, size_t n, __half* out_hi, __half* out_lo, float* hid_f32,
                    cudaStream_t st) {
  bwd_relu_mask_kernel<<<(unsigned)std::min<size_t>(148 * 8, DC_CEIL_DIV(n, (size_t)256)), 256, 0, st>>>(gh, hid_hi, hid_lo, n, out_hi, out_lo, hid_f32);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
int k_bwd_combine(dc_ctx* ctx, const float* g_roi_hwc, const __half* g_rpn_hi, const __half* g_rpn_lo, int C, int cells, float* out_chw,
                  cudaStream_t st) {
  dim3 grid(DC_CEIL_DIV(C, 32), DC_CEIL_DIV(cells, 32));
  bwd_combine_kernel<<<grid, dim3(32, 8), 0, st>>>(g_roi_hwc, g_rpn_hi, g_rpn_lo, C, cells, out_chw);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
int k_bwd_wgrad(dc_ctx* ctx, const float* A, int Mdim, const float* B, int Ndim, int H, int W, int taps, float* out, const float* bias_src,
                float* bias_out, cudaStream_t st) {
  dim3 grid(DC_CEIL_DIV(Ndim, WG_T), DC_CEIL_DIV(Mdim, WG_T), taps);
  bwd_wgrad_kernel<<<grid, 256, 0, st>>>(A, Mdim, B, Ndim, H, W, taps, out);
  DC_CHECK_LAUNCH(ctx);
  if (bias_out) {
    bwd_colsum_kernel<<<DC_CEIL_DIV(Mdim, 128), 128, 0, st>>>(bias_src, Mdim, H * W, bias_out);
    DC_CHECK_LAUNCH(ctx);
  }
  return DC_OK;
}
