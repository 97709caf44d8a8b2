// Box-pipeline leaf operators of the DenseCap localization layer, one kernel per reference module,
// plus the fused RPN post-processing kernel used on the hot path.
//
// Bit-exactness: the reference evaluates these with one fp32 rounding per Torch tensor op and no
// FMA contraction, and integer outputs (NMS picks, valid masks) depend on those roundings.  Every
// arithmetic step below therefore uses the explicit round-to-nearest intrinsics (__fadd_rn,
// __fmul_rn, __fdiv_rn), which nvcc never contracts, in the reference's operation order.
// exp() is evaluated in double and rounded once, which equals glibc's correctly-rounded expf
// used by the reference CPU path (and the oracle's exp_f32).
#include "common.cuh"

__device__ __forceinline__ float exp_cr(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float log_cr(float x) { return (float)log((double)x); }

// ---- box_utils.xcycwh_to_x1y1x2y2 (box_utils.lua:270-298) ------------------------------------
__device__ __forceinline__ float4 d_xcycwh_to_x1y1x2y2(float4 b) {
  const float hw = __fdiv_rn(__fadd_rn(b.z, -1.f), 2.f);
  const float hh = __fdiv_rn(__fadd_rn(b.w, -1.f), 2.f);
  float4 r;
  r.x = __fadd_rn(__fmul_rn(hw, -1.f), b.x);
  r.z = __fadd_rn(hw, b.x);
  r.y = __fadd_rn(__fmul_rn(hh, -1.f), b.y);
  r.w = __fadd_rn(hh, b.y);
  return r;
}
// ---- box_utils.x1y1x2y2_to_xcycwh (box_utils.lua:382-410) ------------------------------------
__device__ __forceinline__ float4 d_x1y1x2y2_to_xcycwh(float4 c) {
  float4 r;
  r.x = __fdiv_rn(__fadd_rn(c.x, c.z), 2.f);
  r.y = __fdiv_rn(__fadd_rn(c.y, c.w), 2.f);
  r.z = __fsub_rn(c.z, c.x);
  r.w = __fsub_rn(c.w, c.y);
  return r;
}
// ---- nn.ApplyBoxTransform (modules/ApplyBoxTransform.lua:63-90) ------------------------------
__device__ __forceinline__ float4 d_apply_box_transform(float4 a, float4 t) {
  float4 r;
  r.x = __fadd_rn(__fmul_rn(t.x, a.z), a.x);
  r.y = __fadd_rn(__fmul_rn(t.y, a.w), a.y);
  r.z = __fmul_rn(exp_cr(t.z), a.z);
  r.w = __fmul_rn(exp_cr(t.w), a.w);
  return r;
}
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
// ---- box_utils.clip_boxes(..., 'xcycwh') (box_utils.lua:486-523) ------------------------------
__device__ __forceinline__ float4 d_clip_box(float4 b, float x_min, float y_min, float x_max, float y_max, bool& valid) {
  float4 c = d_xcycwh_to_x1y1x2y2(b);
  c.x = clampf(c.x, x_min, x_max - 1.f);
  c.y = clampf(c.y, y_min, y_max - 1.f);
  c.z = clampf(c.z, x_min + 1.f, x_max);
  c.w = clampf(c.w, y_min + 1.f, y_max);
  valid = (c.z > c.x) && (c.w > c.y);
  return d_x1y1x2y2_to_xcycwh(c);
}
// ---- objectness probability (LocalizationLayer.lua:305-308) ----------------------------------
__device__ __forceinline__ float d_objectness(float pos, float neg) {
  const float ep = exp_cr(pos), en = exp_cr(neg);
  return __fmul_rn(__fdiv_rn(1.f, __fadd_rn(ep, en)), ep);   // (pos+neg):pow(-1):cmul(pos)
}

// ================================= leaf kernels =============================================
__global__ void make_anchors_kernel(float* __restrict__ out, int N, int H, int W, int k, float x0, float y0, float sx,
                                    float sy, const float* __restrict__ anchors) {
  const size_t total = (size_t)N * 4 * k * H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = i % W;
    const int y = (i / W) % H;
    const int ch = (i / ((size_t)W * H)) % (4 * k);
    const int a = ch >> 2, d = ch & 3;
    float v;
    if (d == 0) v = __fadd_rn(__fmul_rn((float)x, sx), x0);
    else if (d == 1) v = __fadd_rn(__fmul_rn((float)y, sy), y0);
    else if (d == 2) v = anchors[a];
    else v = anchors[k + a];
    out[i] = v;
  }
}
int k_make_anchors(dc_ctx* ctx, float* out, int N, int H, int W, int k, float x0, float y0, float sx, float sy,
                   const float* anchors_dev, cudaStream_t st) {
  const size_t total = (size_t)N * 4 * k * H * W;
  if (total == 0) return DC_OK;
  make_anchors_kernel<<<(unsigned)min((size_t)2048, DC_CEIL_DIV(total, (size_t)256)), 256, 0, st>>>(out, N, H, W, k, x0, y0, sx, sy,
                                                                                                  anchors_dev);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

// (N, k*D, H, W) -> (N, k*H*W, D): out[n][(a*H+y)*W+x][d] = in[n][a*D+d][y][x]
__global__ void reshape_box_features_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int k, int D, int H, int W) {
  const size_t total = (size_t)N * k * D * H * W;
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
    const int d = o % D;
    size_t r = o / D;
    const int x = r % W; r /= W;
    const int y = r % H; r /= H;
    const int a = r % k;
    const int n = r / k;
    out[o] = in[(((size_t)n * k * D + (size_t)a * D + d) * H + y) * W + x];
  }
}
int k_reshape_box_features(dc_ctx* ctx, const float* in, float* out, int N, int k, int D, int H, int W, cudaStream_t st) {
  const size_t total = (size_t)N * k * D * H * W;
  if (total == 0) return DC_OK;
  reshape_box_features_kernel<<<(unsigned)min((size_t)4096, DC_CEIL_DIV(total, (size_t)256)), 256, 0, st>>>(in, out, N, k, D, H, W);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

template <int OP>
__global__ void box4_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (OP == 0) out[i] = d_apply_box_transform(a[i], b[i]);
  if (OP == 1) out[i] = d_xcycwh_to_x1y1x2y2(a[i]);
  if (OP == 2) out[i] = d_x1y1x2y2_to_xcycwh(a[i]);
  if (OP == 3) {  // nn.InvertBoxTransform (modules/InvertBoxTransform.lua:36-60)
    const float4 an = a[i], t = b[i];
    float4 r;
    r.x = __fdiv_rn(__fsub_rn(t.x, an.x), an.z);
    r.y = __fdiv_rn(__fsub_rn(t.y, an.y), an.w);
    r.z = log_cr(__fdiv_rn(t.z, an.z));
    r.w = log_cr(__fdiv_rn(t.w, an.w));
    out[i] = r;
  }
}
#define LAUNCH_BOX4(OP, a, b, out, n)                                                         \
  if ((n) > 0) {                                                                              \
    box4_kernel<OP><<<DC_CEIL_DIV((n), 256), 256, 0, st>>>((const float4*)(a), (const float4*)(b), (float4*)(out), (n)); \
    DC_CHECK_LAUNCH(ctx);                                                                     \
  }                                                                                           \
  return DC_OK;
int k_apply_box_transform(dc_ctx* ctx, const float* boxes, const float* trans, float* out, int n, cudaStream_t st) { LAUNCH_BOX4(0, boxes, trans, out, n) }
int k_xcycwh_to_x1y1x2y2(dc_ctx* ctx, const float* in, float* out, int n, cudaStream_t st) { LAUNCH_BOX4(1, in, in, out, n) }
int k_x1y1x2y2_to_xcycwh(dc_ctx* ctx, const float* in, float* out, int n, cudaStream_t st) { LAUNCH_BOX4(2, in, in, out, n) }
int k_invert_box_transform(dc_ctx* ctx, const float* anchors, const float* targets, float* out, int n, cudaStream_t st) { LAUNCH_BOX4(3, anchors, targets, out, n) }

__global__ void clip_boxes_kernel(const float4* __restrict__ boxes, float4* __restrict__ out, uint8_t* __restrict__ valid, int n,
                                  float x_min, float y_min, float x_max, float y_max) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool v;
  out[i] = d_clip_box(boxes[i], x_min, y_min, x_max, y_max, v);
  valid[i] = v ? 1 : 0;
}
int k_clip_boxes(dc_ctx* ctx, const float* boxes, float* out, uint8_t* valid, int n, float x_min, float y_min, float x_max,
                 float y_max, cudaStream_t st) {
  if (n <= 0) return DC_OK;
  clip_boxes_kernel<<<DC_CEIL_DIV(n, 256), 256, 0, st>>>((const float4*)boxes, (float4*)out, valid, n, x_min, y_min, x_max, y_max);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

__global__ void objectness_prob_kernel(const float2* __restrict__ scores, float* __restrict__ prob, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 s = scores[i];
  prob[i] = d_objectness(s.x, s.y);
}
int k_objectness_prob(dc_ctx* ctx, const float* scores, float* prob, int n, cudaStream_t st) {
  if (n <= 0) return DC_OK;
  objectness_prob_kernel<<<DC_CEIL_DIV(n, 256), 256, 0, st>>>((const float2*)scores, prob, n);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

// nn.BoxToAffine (modules/BoxToAffine.lua:69-93); shared with the ROI-pool kernel.
__global__ void box_to_affine_kernel(const float4* __restrict__ boxes, float* __restrict__ theta, int B, float H, float W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const float4 b = boxes[i];
  float* t = theta + (size_t)i * 6;
  t[0] = __fdiv_rn(b.w, H);
  t[1] = 0.f;
  t[2] = __fdiv_rn(__fadd_rn(__fmul_rn(b.y, 2.f), -1.f - H), H - 1.f);
  t[3] = 0.f;
  t[4] = __fdiv_rn(b.z, W);
  t[5] = __fdiv_rn(__fadd_rn(__fmul_rn(b.x, 2.f), -1.f - W), W - 1.f);
}
int k_box_to_affine(dc_ctx* ctx, const float* boxes, float* theta, int B, int H, int W, cudaStream_t st) {
  if (B <= 0) return DC_OK;
  box_to_affine_kernel<<<DC_CEIL_DIV(B, 256), 256, 0, st>>>((const float4*)boxes, theta, B, (float)H, (float)W);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

// stnbhwd nn.AffineGridGeneratorBHWD(HH, WW) (call site modules/BilinearRoiPooling.lua:52):
// grid[b][i][j] = baseGrid[i][j] . theta[b]^T with baseGrid = (y_i, x_j, 1), y_i = -1 + (i/(HH-1))*2 evaluated in double and stored as
// float; the 3-term dot product is summed left to right (sgemm order), no FMA.  theta (B,2,3) -> grids (B,HH,WW,2) = (gy, gx).
__global__ void affine_grid_kernel(const float* __restrict__ theta, int B, int HH, int WW, float2* __restrict__ grids) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * HH * WW) return;
  const int b = idx / (HH * WW), p = idx % (HH * WW), i = p / WW, j = p % WW;
  const float* t = theta + (size_t)b * 6;
  const float yb = (float)(-1.0 + ((double)i / (double)(HH - 1)) * 2.0);
  const float xb = (float)(-1.0 + ((double)j / (double)(WW - 1)) * 2.0);
  float2 g;
  g.x = __fadd_rn(__fadd_rn(__fmul_rn(t[0], yb), __fmul_rn(t[1], xb)), t[2]);
  g.y = __fadd_rn(__fadd_rn(__fmul_rn(t[3], yb), __fmul_rn(t[4], xb)), t[5]);
  grids[idx] = g;
}
int k_affine_grid(dc_ctx* ctx, const float* theta, int B, int HH, int WW, float* grids, cudaStream_t st) {
  if (B <= 0) return DC_OK;
  affine_grid_kernel<<<DC_CEIL_DIV(B * HH * WW, 256), 256, 0, st>>>(theta, B, HH, WW, (float2*)grids);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

// nn.BoxIoU (modules/BoxIoU.lua:40-73): area = w*h, corners from (w-1)/2, no +1.
__global__ void box_iou_kernel(const float4* __restrict__ b1, const float4* __restrict__ b2, float* __restrict__ out, int B1, int B2) {
  __shared__ float4 s_lohi[128];
  __shared__ float s_area[128];
  const int j0 = blockIdx.x * 128;
  for (int t = threadIdx.x; t < 128 && j0 + t < B2; t += blockDim.x) {
    const float4 b = b2[j0 + t];
    s_lohi[t] = d_xcycwh_to_x1y1x2y2(b);
    s_area[t] = __fmul_rn(b.z, b.w);
  }
  __syncthreads();
  const int i = blockIdx.y * blockDim.x + threadIdx.x;
  if (i >= B1) return;
  const float4 bi = b1[i];
  const float4 l1 = d_xcycwh_to_x1y1x2y2(bi);
  const float a1 = __fmul_rn(bi.z, bi.w);
  const int jn = min(128, B2 - j0);
  for (int t = 0; t < jn; ++t) {
    const float4 l2 = s_lohi[t];
    const float x0 = fmaxf(l1.x, l2.x), y0 = fmaxf(l1.y, l2.y);
    const float x1 = fminf(l1.z, l2.z), y1 = fminf(l1.w, l2.w);
    const float w = fmaxf(__fsub_rn(x1, x0), 0.f), h = fmaxf(__fsub_rn(y1, y0), 0.f);
    const float inter = __fmul_rn(w, h);
    const float den = __fadd_rn(__fsub_rn(a1, inter), s_area[t]);
    out[(size_t)i * B2 + j0 + t] = __fmul_rn(__fdiv_rn(1.f, den), inter);
  }
}
int k_box_iou(dc_ctx* ctx, const float* b1, const float* b2, float* out, int B1, int B2, cudaStream_t st) {
  if (B1 <= 0 || B2 <= 0) return DC_OK;
  dim3 grid(DC_CEIL_DIV(B2, 128), DC_CEIL_DIV(B1, 128));
  box_iou_kernel<<<grid, 128, 0, st>>>((const float4*)b1, (const float4*)b2, out, B1, B2);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

// ============================ fused RPN post-processing =====================================
// One thread per output row r = (a*H + y)*W + x (ReshapeBoxFeatures order).  Pass 0 counts the
// valid boxes of each block, pass 1 recomputes and scatters them compacted in row order
// (= clamp_data masked-select, LocalizationLayer.lua:285-298).
struct RpnRow {
  float4 box;   // clipped xcycwh
  float4 x1y1;  // xcycwh_to_x1y1x2y2(box)  (LocalizationLayer.lua:302)
  float prob;
  bool valid;
};
__device__ __forceinline__ RpnRow rpn_row(const float* __restrict__ head, int ld, int H, int W, int k,
                                          const float* __restrict__ anchors, float x0, float y0, float sx, float sy,
                                          float img_w, float img_h, int clip, int r) {
  const int x = r % W;
  const int y = (r / W) % H;
  const int a = r / (W * H);
  const float* hp = head + (size_t)(y * W + x) * ld;
  float4 anc;
  anc.x = __fadd_rn(__fmul_rn((float)x, sx), x0);   // MakeAnchors.lua:44-47
  anc.y = __fadd_rn(__fmul_rn((float)y, sy), y0);
  anc.z = anchors[a];
  anc.w = anchors[k + a];
  const float4 t = make_float4(hp[4 * a + 0], hp[4 * a + 1], hp[4 * a + 2], hp[4 * a + 3]);
  RpnRow o;
  o.box = d_apply_box_transform(anc, t);
  o.valid = true;
  if (clip) o.box = d_clip_box(o.box, 1.f, 1.f, img_w, img_h, o.valid);   // LocalizationLayer.lua:273-279
  o.x1y1 = d_xcycwh_to_x1y1x2y2(o.box);
  o.prob = d_objectness(hp[4 * k + 2 * a], hp[4 * k + 2 * a + 1]);
  return o;
}

template <int PASS>
__global__ void rpn_post_kernel(const float* __restrict__ head, int ld, int H, int W, int k, const float* __restrict__ anchors,
                                float x0, float y0, float sx, float sy, float img_w, float img_h, int clip, RpnPostOut out) {
  const int A = k * H * W;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  RpnRow row;
  row.valid = false;
  if (r < A) row = rpn_row(head, ld, H, W, k, anchors, x0, y0, sx, sy, img_w, img_h, clip, r);
  const unsigned ballot = __ballot_sync(0xffffffffu, row.valid);
  __shared__ int warp_counts[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_counts[warp] = __popc(ballot);
  __syncthreads();
  if (PASS == 0) {
    if (threadIdx.x == 0) {
      int c = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) c += warp_counts[w];
      out.block_counts[blockIdx.x] = c;
    }
    return;
  }
  // PASS 1: block offset = sum of the counts of all preceding blocks (a few dozen values)
  __shared__ int block_off;
  if (warp == 0) {
    int s = 0;
    for (int b = lane; b < (int)blockIdx.x; b += 32) s += out.block_counts[b];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) block_off = s;
  }
  __syncthreads();
  int off = block_off;
  for (int w = 0; w < warp; ++w) off += warp_counts[w];
  off += __popc(ballot & ((1u << lane) - 1u));
  if (row.valid) {
    ((float4*)out.boxes_xcycwh)[off] = row.box;
    ((float4*)out.boxes_x1y1x2y2)[off] = row.x1y1;
    out.prob[off] = row.prob;
    out.src_row[off] = r;
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == blockDim.x - 1) {
    // last thread of the last block: total = its exclusive offset + its own flag
    *out.n_valid = off + (row.valid ? 1 : 0);
  }
}

int k_rpn_postprocess(dc_ctx* ctx, const float* head, int ld, int H, int W, int k, const float* anchors_dev, float x0, float y0,
                      float sx, float sy, float img_w, float img_h, int clip, RpnPostOut out, cudaStream_t st) {
  const int A = k * H * W;
  if (A <= 0) return DC_OK;
  const int blocks = DC_CEIL_DIV(A, 256);
  rpn_post_kernel<0><<<blocks, 256, 0, st>>>(head, ld, H, W, k, anchors_dev, x0, y0, sx, sy, img_w, img_h, clip, out);
  DC_CHECK_LAUNCH(ctx);
  rpn_post_kernel<1><<<blocks, 256, 0, st>>>(head, ld, H, W, k, anchors_dev, x0, y0, sx, sy, img_w, img_h, clip, out);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
