// Layout / format conversion kernels between the reference's tensor layouts (fp32 NCHW, (out,in)
// Linear weights, ...) and the library's internal "split-fp16 planes" NHWC / K-major layouts.
#include "common.cuh"
#include <algorithm>

static inline unsigned grid_for(size_t n, int threads = 256, unsigned cap = 148 * 16) {
  size_t b = DC_CEIL_DIV(n, (size_t)threads);
  if (b < 1) b = 1;
  return (unsigned)(b > cap ? cap : b);
}

// (N,C,H,W) fp32 -> split (N,H,W,C).  32x32 smem transpose tile over (C, HW) per image.
__global__ void nchw_to_split_nhwc_kernel(const float* __restrict__ in, __half* __restrict__ hi, __half* __restrict__ lo, int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* src = in + (size_t)n * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? src[(size_t)c * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < HW && c < C) {
      __half h, l;
      split_f32(tile[threadIdx.x][i], h, l);
      const size_t o = ((size_t)n * HW + p) * C + c;
      hi[o] = h;
      lo[o] = l;
    }
  }
}
int k_nchw_to_split_nhwc(dc_ctx* ctx, const float* in, __half* hi, __half* lo, int N, int C, int H, int W, cudaStream_t st) {
  if ((size_t)N * C * H * W == 0) return DC_OK;
  dim3 grid(DC_CEIL_DIV(H * W, 32), DC_CEIL_DIV(C, 32), N), block(32, 8);
  nchw_to_split_nhwc_kernel<<<grid, block, 0, st>>>(in, hi, lo, C, H * W);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

__global__ void split_nhwc_to_nchw_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, float* __restrict__ out, int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (p < HW && c < C) {
      const size_t g = ((size_t)n * HW + p) * C + c;
      v = join_f32(hi[g], lo[g]);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  float* dst = out + (size_t)n * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) dst[(size_t)c * HW + p] = tile[threadIdx.x][i];
  }
}
int k_split_nhwc_to_nchw(dc_ctx* ctx, const __half* hi, const __half* lo, float* out, int N, int C, int H, int W, cudaStream_t st) {
  if ((size_t)N * C * H * W == 0) return DC_OK;
  dim3 grid(DC_CEIL_DIV(H * W, 32), DC_CEIL_DIV(C, 32), N), block(32, 8);
  split_nhwc_to_nchw_kernel<<<grid, block, 0, st>>>(hi, lo, out, C, H * W);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

__global__ void rowmajor_to_split_kernel(const float* __restrict__ in, __half* __restrict__ hi, __half* __restrict__ lo, size_t n, float scale) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    __half h, l;
    split_f32(in[i] * scale, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}
int k_rowmajor_to_split(dc_ctx* ctx, const float* in, __half* hi, __half* lo, size_t n, cudaStream_t st, float scale) {
  if (n == 0) return DC_OK;
  rowmajor_to_split_kernel<<<grid_for(n), 256, 0, st>>>(in, hi, lo, n, scale);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
__global__ void split_to_rowmajor_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, float* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = join_f32(hi[i], lo[i]);
}
int k_split_to_rowmajor(dc_ctx* ctx, const __half* hi, const __half* lo, float* out, size_t n, cudaStream_t st) {
  if (n == 0) return DC_OK;
  split_to_rowmajor_kernel<<<grid_for(n), 256, 0, st>>>(hi, lo, out, n);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

// nn.SpatialMaxPooling(2,2,2,2) (ceil mode as emitted by loadcaffe) on split NHWC.  The max of
// the reconstructed values is selected and its (hi,lo) pair copied, so pooling is exact.
__global__ void maxpool2x2_split_kernel(const __half* __restrict__ ihi, const __half* __restrict__ ilo, __half* __restrict__ ohi,
                                        __half* __restrict__ olo, int N, int H, int W, int C, int OH, int OW) {
  const size_t total = (size_t)N * OH * OW * C;
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
    const int c = o % C;
    size_t r = o / C;
    const int ox = r % OW; r /= OW;
    const int oy = r % OH;
    const int n = r / OH;
    float best = -INFINITY;
    __half bh = __float2half(0.f), bl = __float2half(0.f);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int y = oy * 2 + dy, x = ox * 2 + dx;
        if (y < H && x < W) {
          const size_t g = (((size_t)n * H + y) * W + x) * C + c;
          const __half h = ihi[g], l = ilo[g];
          const float v = join_f32(h, l);
          if (v > best) { best = v; bh = h; bl = l; }
        }
      }
    ohi[o] = bh;
    olo[o] = bl;
  }
}
int k_maxpool2x2_split(dc_ctx* ctx, const __half* ihi, const __half* ilo, __half* ohi, __half* olo, int N, int H, int W, int C,
                       int ceil_mode, cudaStream_t st) {
  const int OH = ceil_mode ? (H + 1) / 2 : H / 2, OW = ceil_mode ? (W + 1) / 2 : W / 2;
  const size_t total = (size_t)N * OH * OW * C;
  if (total == 0) return DC_OK;
  maxpool2x2_split_kernel<<<grid_for(total, 256, 148 * 32), 256, 0, st>>>(ihi, ilo, ohi, olo, N, H, W, C, OH, OW);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

// conv weight (Cout,Cin,3,3) -> split [Cout][(r*3+s)*Cin + ci]   (K-major B operand of the implicit GEMM)
__global__ void conv_weight_to_split_kernel(const float* __restrict__ w, __half* __restrict__ hi, __half* __restrict__ lo, int Cout, int Cin, float scale) {
  const size_t total = (size_t)Cout * 9 * Cin;
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
    const int ci = o % Cin;
    const int tap = (o / Cin) % 9;
    const int co = o / ((size_t)9 * Cin);
    __half h, l;
    split_f32(w[((size_t)co * Cin + ci) * 9 + tap] * scale, h, l);
    hi[o] = h;
    lo[o] = l;
  }
}
int k_conv_weight_to_split(dc_ctx* ctx, const float* w, __half* hi, __half* lo, int Cout, int Cin, cudaStream_t st, float scale) {
  const size_t total = (size_t)Cout * 9 * Cin;
  if (total == 0) return DC_OK;
  conv_weight_to_split_kernel<<<grid_for(total), 256, 0, st>>>(w, hi, lo, Cout, Cin, scale);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

// fc6 weight (out, C*hw) with input order (c, pos) -> split (out, hw*C) in the order of the pooled ROI rows (roi_feature_index),
// so that the output of the ROI pooling feeds fc6 without the reference's nn.Transpose/View.
__global__ void fc6_weight_to_split_kernel(const float* __restrict__ w, __half* __restrict__ hi, __half* __restrict__ lo, int out, int C, int hw, float scale) {
  const size_t K = (size_t)C * hw;
  const size_t total = (size_t)out * K;
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
    const int pos = o % hw;                      // o walks the reference order (row, c, pos): coalesced reads
    const int c = (o / hw) % C;
    const size_t row = o / K;
    __half h, l;
    split_f32(w[o] * scale, h, l);
    const size_t d = row * K + roi_feature_index(c, pos, C, hw);
    hi[d] = h;
    lo[d] = l;
  }
}
int k_fc6_weight_to_split(dc_ctx* ctx, const float* w, __half* hi, __half* lo, int out, int C, int hw, cudaStream_t st, float scale) {
  const size_t total = (size_t)out * C * hw;
  if (total == 0) return DC_OK;
  fc6_weight_to_split_kernel<<<grid_for(total, 256, 148 * 64), 256, 0, st>>>(w, hi, lo, out, C, hw, scale);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

// (rows, cols) fp32 -> split (cols, rows)   (torch-rnn LSTM weight ((D+H),4H) -> [4H][D+H])
__global__ void transpose_to_split_kernel(const float* __restrict__ w, __half* __restrict__ hi, __half* __restrict__ lo, int rows, int cols,
                                          int interleave_h, float scale) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? w[(size_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) {
      __half h, l;
      split_f32(tile[threadIdx.x][i] * scale, h, l);
      const int oc = interleave_h > 0 ? 4 * (c % interleave_h) + c / interleave_h : c;
      hi[(size_t)oc * rows + r] = h;
      lo[(size_t)oc * rows + r] = l;
    }
  }
}
int k_transpose_to_split(dc_ctx* ctx, const float* w, __half* hi, __half* lo, int rows, int cols, int interleave_h, cudaStream_t st, float scale) {
  if ((size_t)rows * cols == 0) return DC_OK;
  dim3 grid(DC_CEIL_DIV(cols, 32), DC_CEIL_DIV(rows, 32)), block(32, 8);
  transpose_to_split_kernel<<<grid, block, 0, st>>>(w, hi, lo, rows, cols, interleave_h, scale);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

// (K,C,hw) fp32 BCHW -> split rows (K, hw*C) in roi_feature_index order
__global__ void bchw_to_split_rows_kernel(const float* __restrict__ in, __half* __restrict__ hi, __half* __restrict__ lo, int K, int C, int hw) {
  const size_t total = (size_t)K * C * hw;
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
    const int pos = o % hw;
    const int c = (o / hw) % C;
    const size_t b = o / ((size_t)C * hw);
    __half h, l;
    split_f32(in[o], h, l);
    const size_t d = b * ((size_t)C * hw) + roi_feature_index(c, pos, C, hw);
    hi[d] = h;
    lo[d] = l;
  }
}
int k_bchw_to_split_rows(dc_ctx* ctx, const float* in, __half* hi, __half* lo, int K, int C, int hw, cudaStream_t st) {
  const size_t total = (size_t)K * C * hw;
  if (total == 0) return DC_OK;
  bchw_to_split_rows_kernel<<<grid_for(total, 256, 148 * 32), 256, 0, st>>>(in, hi, lo, K, C, hw);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
__global__ void bhwc_split_to_bchw_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, float* __restrict__ out, int K, int C, int hw) {
  const size_t total = (size_t)K * C * hw;
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
    const int pos = o % hw;
    const int c = (o / hw) % C;
    const size_t b = o / ((size_t)C * hw);
    const size_t g = b * ((size_t)C * hw) + roi_feature_index(c, pos, C, hw);
    out[o] = join_f32(hi[g], lo[g]);
  }
}
int k_bhwc_split_to_bchw(dc_ctx* ctx, const __half* hi, const __half* lo, float* out, int K, int C, int hw, cudaStream_t st) {
  const size_t total = (size_t)K * C * hw;
  if (total == 0) return DC_OK;
  bhwc_split_to_bchw_kernel<<<grid_for(total, 256, 148 * 32), 256, 0, st>>>(hi, lo, out, K, C, hw);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

// max |x| of a tensor (used to pick the power-of-two weight pre-scale); out must be zero-initialised
__global__ void absmax_kernel(const float* __restrict__ in, size_t n, float* __restrict__ out) {
  float m = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = fabsf(in[i]);
    if (v == v) m = fmaxf(m, v);
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(m));   // valid for non-negative floats
}
int k_absmax(dc_ctx* ctx, const float* in, size_t n, float* out_dev, cudaStream_t st) {
  if (n == 0) return DC_OK;
  absmax_kernel<<<grid_for(n, 256, 148 * 8), 256, 0, st>>>(in, n, out_dev);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

// Image preprocessing of run_model.lua:66-74 / webcam/daemon.lua:66-78 without the resize: decoded uint8 RGB HWC ->
// float [0,1] (image.load: byte / 255), RGB -> BGR, * 255, - VGG mean (add(-1, mean): one product, one sum), NCHW.
__global__ void preprocess_u8_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int N, int H, int W) {
  const size_t hw = (size_t)H * W, total = (size_t)N * hw;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / hw, p = i % hw;
    const uint8_t* px = in + i * 3;
    const float mean[3] = {103.939f, 116.779f, 123.68f};              // B, G, R
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __fmul_rn(__fdiv_rn((float)px[2 - c], 255.f), 255.f);
      out[(n * 3 + c) * hw + p] = __fadd_rn(v, __fmul_rn(-1.f, mean[c]));
    }
  }
}
int k_preprocess_u8(dc_ctx* ctx, const uint8_t* in, float* out, int N, int H, int W, cudaStream_t st) {
  const size_t total = (size_t)N * H * W;
  if (total == 0) return DC_OK;
  preprocess_u8_kernel<<<(unsigned)std::min<size_t>(148 * 8, DC_CEIL_DIV(total, (size_t)256)), 256, 0, st>>>(in, out, N, H, W);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// image.scale(img, size), bilinear mode (run_model.lua:69, webcam/daemon.lua:70; torch `image` rock, restated in
// oracle/densecap_oracle.py image_scale): two separable 1-D passes in fp32, width first.  One thread per output element;
// the arithmetic (float index products, (1-f)*a + f*b without contraction, area average acc / n) follows the C routine
// Main_scaleLinear_rowcol operation by operation.  Pass 1 reads the decoded uint8 RGB HWC image (image.load: byte / 255),
// pass 2 applies the rest of run_model.lua:70-74 (RGB -> BGR, * 255, - VGG mean) and writes NCHW.
// ---------------------------------------------------------------------------------------------------------------------
template <typename Load>
__device__ __forceinline__ float scale_1d(Load src, int src_len, int dst_len, int di) {
  if (dst_len > src_len) {
    if (src_len == 1) return src(0);
    if (di == dst_len - 1) return src(src_len - 1);
    const float scale = __fdiv_rn((float)(src_len - 1), (float)(dst_len - 1));
    float sf = __fmul_rn((float)di, scale);
    const int si = (int)sf;
    sf = __fsub_rn(sf, (float)si);
    return __fadd_rn(__fmul_rn(__fsub_rn(1.f, sf), src(si)), __fmul_rn(sf, src(si + 1)));
  }
  if (dst_len < src_len) {
    const float scale = __fdiv_rn((float)src_len, (float)dst_len);
    const float s0 = __fmul_rn((float)di, scale);
    const int si0 = (int)s0;
    const float f0 = (di == 0) ? 0.f : __fsub_rn(s0, (float)si0);
    const float s1 = __fmul_rn((float)(di + 1), scale);
    const int si1 = (int)s1;
    const float f1 = __fsub_rn(s1, (float)si1);
    float acc = __fmul_rn(__fsub_rn(1.f, f0), src(di == 0 ? 0 : si0));
    float n = __fsub_rn(1.f, f0);
    for (int si = (di == 0 ? 0 : si0) + 1; si < si1; ++si) { acc = __fadd_rn(acc, src(si)); n = __fadd_rn(n, 1.f); }
    if (si1 < src_len) { acc = __fadd_rn(acc, __fmul_rn(f1, src(si1))); n = __fadd_rn(n, f1); }
    return __fdiv_rn(acc, n);
  }
  return src(di);
}
__global__ void image_scale_w_kernel(const uint8_t* __restrict__ in_hwc, int H, int W, int ow, float* __restrict__ tmp) {
  const size_t total = (size_t)3 * H * ow;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = i % ow, y = (i / ow) % H, c = i / ((size_t)ow * H);
    const uint8_t* row = in_hwc + (size_t)y * W * 3 + c;
    tmp[i] = scale_1d([&](int sx) { return __fdiv_rn((float)row[(size_t)sx * 3], 255.f); }, W, ow, x);
  }
}
__global__ void image_scale_h_kernel(const float* __restrict__ tmp, int H, int oh, int ow, float* __restrict__ out_nchw) {
  const size_t total = (size_t)3 * oh * ow;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = i % ow, y = (i / ow) % oh, c = i / ((size_t)ow * oh);          // c = OUTPUT channel (B, G, R)
    const float* col = tmp + (size_t)(2 - c) * H * ow + x;
    const float v = scale_1d([&](int sy) { return col[(size_t)sy * ow]; }, H, oh, y);
    const float mean[3] = {103.939f, 116.779f, 123.68f};
    out_nchw[i] = __fadd_rn(__fmul_rn(v, 255.f), __fmul_rn(-1.f, mean[c]));
  }
}
int k_image_scale_preprocess(dc_ctx* ctx, const uint8_t* in_hwc, int H, int W, int oh, int ow, float* tmp, float* out_nchw, cudaStream_t st) {
  if ((size_t)H * W == 0 || (size_t)oh * ow == 0) return DC_OK;
  image_scale_w_kernel<<<(unsigned)std::min<size_t>(148 * 8, DC_CEIL_DIV((size_t)3 * H * ow, (size_t)256)), 256, 0, st>>>(in_hwc, H, W, ow, tmp);
  DC_CHECK_LAUNCH(ctx);
  image_scale_h_kernel<<<(unsigned)std::min<size_t>(148 * 8, DC_CEIL_DIV((size_t)3 * oh * ow, (size_t)256)), 256, 0, st>>>(tmp, H, oh, ow, out_nchw);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

// conv1_1 filter bank (64, 3, 3, 3) fp32 -> split [64][64] K-major rows: k = c*9 + r*3 + s (the natural flatten) for k < 27, zeros above
__global__ void conv_first_weight_kernel(const float* __restrict__ w, __half* __restrict__ hi, __half* __restrict__ lo, float scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 64) return;
  const int co = i / 64, k = i % 64;
  __half h = __float2half(0.f), l = __float2half(0.f);
  if (k < 27) split_f32(w[co * 27 + k] * scale, h, l);
  hi[i] = h; lo[i] = l;
}
int k_conv_first_weight_to_split(dc_ctx* ctx, const float* w, __half* hi, __half* lo, float scale, cudaStream_t st) {
  conv_first_weight_kernel<<<16, 256, 0, st>>>(w, hi, lo, scale);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
