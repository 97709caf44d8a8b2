// The small dense operators of the path that have no tensor-core shape and run on the CUDA cores:
// conv1_1 (K = 27) and the 4096->1 / 4096->4 objectness / box-regression heads.
#include "common.cuh"

// conv1_1 (VGG layer 1): Cin = 3, K = 27 -- no tensor-core shape; the layer is bound by the 110 MB split-fp16
// output write, so a direct fp32 kernel.  A block owns a 32 x 8 pixel tile: the 3 x 10 x 34 input halo and the
// (27, Cout) weights sit in shared memory; each thread keeps the 27 taps of ITS pixel in registers and sweeps the
// output channels in groups of 16, so every weight read is a warp-wide broadcast LDS.128 shared by 16 FMAs, the
// image is read from global memory exactly once per block, and each lane stores 2 x 32 B per plane per group.
// img NCHW fp32 (reference input layout, DenseCapModel.lua:244) -> split NHWC, bias + ReLU fused.
constexpr int CF_TW = 32, CF_TH = 8;
__global__ void __launch_bounds__(CF_TW * CF_TH)
conv_first_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                  __half* __restrict__ y_hi, __half* __restrict__ y_lo, int N, int H, int W, int Cout) {
  extern __shared__ __align__(16) float sw[];   // [27][Cout] weights, bias[Cout], halo [3][CF_TH+2][CF_TW+2]
  float* sb = sw + 27 * Cout;
  float* halo = sb + Cout;
  const int tid = threadIdx.x;
  for (int i = tid; i < 27 * Cout; i += blockDim.x) {
    const int co = i % Cout, k = i / Cout;       // k = ci*9 + r*3 + s  (reference layout (Cout,3,3,3))
    sw[i] = w[(size_t)co * 27 + k];
  }
  for (int i = tid; i < Cout; i += blockDim.x) sb[i] = bias ? bias[i] : 0.f;
  const int n = blockIdx.z;
  const int x0 = blockIdx.x * CF_TW, y0 = blockIdx.y * CF_TH;
  constexpr int HW_ = CF_TW + 2, HH_ = CF_TH + 2;
  for (int i = tid; i < 3 * HH_ * HW_; i += blockDim.x) {
    const int hx = i % HW_, hy = (i / HW_) % HH_, ci = i / (HW_ * HH_);
    const int yy = y0 + hy - 1, xx = x0 + hx - 1;
    halo[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + (((size_t)n * 3 + ci) * H + yy) * W + xx) : 0.f;
  }
  __syncthreads();
  const int tx = tid % CF_TW, ty = tid / CF_TW;
  const int x = x0 + tx, y = y0 + ty;
  if (x >= W || y >= H) return;
  float v[27];
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s2 = 0; s2 < 3; ++s2) v[ci * 9 + r * 3 + s2] = halo[(ci * HH_ + ty + r) * HW_ + tx + s2];
  const size_t pix = ((size_t)n * H + y) * W + x;
  for (int gq = 0; gq < Cout / 16; ++gq) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = sb[gq * 16 + j];
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      const float4* wp = reinterpret_cast<const float4*>(sw + k * Cout + gq * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 wv = wp[q];
        acc[4 * q + 0] = fmaf(v[k], wv.x, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(v[k], wv.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(v[k], wv.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(v[k], wv.w, acc[4 * q + 3]);
      }
    }
    uint32_t hh[8], ll[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      __half h0, l0, h1, l1;
      split_f32(fmaxf(acc[2 * j], 0.f), h0, l0);
      split_f32(fmaxf(acc[2 * j + 1], 0.f), h1, l1);
      hh[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      ll[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    const size_t o = pix * Cout + gq * 16;
    reinterpret_cast<uint4*>(y_hi + o)[0] = make_uint4(hh[0], hh[1], hh[2], hh[3]);
    reinterpret_cast<uint4*>(y_hi + o)[1] = make_uint4(hh[4], hh[5], hh[6], hh[7]);
    reinterpret_cast<uint4*>(y_lo + o)[0] = make_uint4(ll[0], ll[1], ll[2], ll[3]);
    reinterpret_cast<uint4*>(y_lo + o)[1] = make_uint4(ll[4], ll[5], ll[6], ll[7]);
  }
}
// Cout = 64 variant (the VGG-16 case), lanes = output-channel pairs.  The generic kernel above gives every lane one pixel,
// so each 16-byte store of a warp touches 32 different 128-byte lines and the kernel is bound by the store path (ncu r06:
// 92 us for 110 MB).  Here a warp owns one image row of the tile: lane l keeps the 2 x 27 weights of channels (2l, 2l+1) in
// registers, the 3 x 3 x 3 input window slides along the row (three broadcast LDS.128 per pixel; the window columns rotate
// through an unroll-by-3, no register moves), and the 64 channels of a pixel leave as ONE coalesced 128-byte store per plane.
constexpr int CF2_TW = 60, CF2_TH = 8;
__device__ __forceinline__ void cf2_pixel(const float4 (&L)[3], const float4 (&Cn)[3], const float4 (&R)[3], const float (&w0)[27],
                                          const float (&w1)[27], float b0, float b1, bool store, size_t pix, int lane,
                                          __half* __restrict__ y_hi, __half* __restrict__ y_lo) {
  float a0 = b0, a1 = b1;
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float vl = ci == 0 ? L[r].x : (ci == 1 ? L[r].y : L[r].z);
      const float vc = ci == 0 ? Cn[r].x : (ci == 1 ? Cn[r].y : Cn[r].z);
      const float vr = ci == 0 ? R[r].x : (ci == 1 ? R[r].y : R[r].z);
      const int k = ci * 9 + r * 3;                        // k = ci*9 + r*3 + s, same accumulation order as the generic kernel
      a0 = fmaf(vl, w0[k], a0); a1 = fmaf(vl, w1[k], a1);
      a0 = fmaf(vc, w0[k + 1], a0); a1 = fmaf(vc, w1[k + 1], a1);
      a0 = fmaf(vr, w0[k + 2], a0); a1 = fmaf(vr, w1[k + 2], a1);
    }
  if (store) {
    __half h0, l0, h1, l1;
    split_f32(fmaxf(a0, 0.f), h0, l0);
    split_f32(fmaxf(a1, 0.f), h1, l1);
    const size_t o = pix * 64 + 2 * lane;
    *reinterpret_cast<uint32_t*>(y_hi + o) = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
    *reinterpret_cast<uint32_t*>(y_lo + o) = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
  }
}
__global__ void __launch_bounds__(CF2_TH * 32, 2)
conv_first64_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                    __half* __restrict__ y_hi, __half* __restrict__ y_lo, int N, int H, int W) {
  __shared__ __align__(16) float4 halo[CF2_TH + 2][CF2_TW + 2];    // (ch0, ch1, ch2, -) per input pixel
  __shared__ float sw[27 * 64];                                     // [k][co]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.z, x0 = blockIdx.x * CF2_TW, y0 = blockIdx.y * CF2_TH;
  for (int i = tid; i < 27 * 64; i += blockDim.x) {
    const int co = i & 63, k = i >> 6;                              // reference layout (Cout, 3, 3, 3): w[co*27 + k]
    sw[i] = __ldg(w + (size_t)co * 27 + k);
  }
  for (int i = tid; i < (CF2_TH + 2) * (CF2_TW + 2); i += blockDim.x) {
    const int hx = i % (CF2_TW + 2), hy = i / (CF2_TW + 2);
    const int yy = y0 + hy - 1, xx = x0 + hx - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const size_t o = ((size_t)n * 3 * H + yy) * W + xx;
      v.x = __ldg(img + o); v.y = __ldg(img + o + (size_t)H * W); v.z = __ldg(img + o + 2 * (size_t)H * W);
    }
    halo[hy][hx] = v;
  }
  __syncthreads();
  const int y = y0 + warp;
  if (y >= H) return;
  float w0[27], w1[27];
#pragma unroll
  for (int k = 0; k < 27; ++k) {
    const float2 t = *reinterpret_cast<const float2*>(sw + k * 64 + 2 * lane);
    w0[k] = t.x; w1[k] = t.y;
  }
  const float b0 = bias ? __ldg(bias + 2 * lane) : 0.f, b1 = bias ? __ldg(bias + 2 * lane + 1) : 0.f;
  const size_t row = ((size_t)n * H + y) * W;
  float4 A[3], B[3], C[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) { A[r] = halo[warp + r][0]; B[r] = halo[warp + r][1]; }
#pragma unroll 1
  for (int x = 0; x < CF2_TW; x += 3) {
    if (x0 + x >= W) break;                                         // warp-uniform
#pragma unroll
    for (int r = 0; r < 3; ++r) C[r] = halo[warp + r][x + 2];
    cf2_pixel(A, B, C, w0, w1, b0, b1, x0 + x < W, row + x0 + x, lane, y_hi, y_lo);
#pragma unroll
    for (int r = 0; r < 3; ++r) A[r] = halo[warp + r][x + 3];
    cf2_pixel(B, C, A, w0, w1, b0, b1, x0 + x + 1 < W, row + x0 + x + 1, lane, y_hi, y_lo);
#pragma unroll
    for (int r = 0; r < 3; ++r) B[r] = halo[warp + r][x + 4];
    cf2_pixel(C, A, B, w0, w1, b0, b1, x0 + x + 2 < W, row + x0 + x + 2, lane, y_hi, y_lo);
  }
}

int k_conv_first(dc_ctx* ctx, const float* img, const float* w, const float* bias, __half* y_hi, __half* y_lo, int N, int H, int W,
                 int Cout, cudaStream_t st) {
  DC_REQUIRE(ctx, Cout % 16 == 0, DC_E_UNSUPPORTED, "conv1_1: Cout must be a multiple of 16");
  if ((size_t)N * H * W == 0) return DC_OK;
  if (Cout == 64) {
    dim3 grid(DC_CEIL_DIV(W, CF2_TW), DC_CEIL_DIV(H, CF2_TH), N);
    conv_first64_kernel<<<grid, CF2_TH * 32, 0, st>>>(img, w, bias, y_hi, y_lo, N, H, W);
    DC_CHECK_LAUNCH(ctx);
    return DC_OK;
  }
  const size_t smem = (size_t)(27 * Cout + Cout + 3 * (CF_TH + 2) * (CF_TW + 2)) * sizeof(float);
  dim3 grid(DC_CEIL_DIV(W, CF_TW), DC_CEIL_DIV(H, CF_TH), N);
  conv_first_kernel<<<grid, CF_TW * CF_TH, smem, st>>>(img, w, bias, y_hi, y_lo, N, H, W, Cout);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

// objectness (fc_dim -> 1) and box regression (fc_dim -> 4) heads (DenseCapModel.lua:90-99): one 128-thread
// block per ROI row computes all five dot products in a single pass over the codes.
__global__ void __launch_bounds__(128)
recog_heads_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w_obj, const float* __restrict__ b_obj,
                   const float* __restrict__ w_box, const float* __restrict__ b_box, float* __restrict__ obj,
                   float* __restrict__ trans, int K) {
  const int m = blockIdx.x;
  const float* xr = x + (size_t)m * ldx;
  float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k = threadIdx.x; k < K; k += 128) {
    const float v = xr[k];
    s[0] = fmaf(v, __ldg(w_obj + k), s[0]);
#pragma unroll
    for (int j = 0; j < 4; ++j) s[1 + j] = fmaf(v, __ldg(w_box + (size_t)j * K + k), s[1 + j]);
  }
  __shared__ float red[4][5];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    float v = s[j];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    const int j = threadIdx.x;
    const float v = red[0][j] + red[1][j] + red[2][j] + red[3][j];
    if (j == 0) obj[m] = v + (b_obj ? b_obj[0] : 0.f);
    else trans[(size_t)m * 4 + j - 1] = v + (b_box ? b_box[j - 1] : 0.f);
  }
}
int k_recog_heads(dc_ctx* ctx, const float* x, int ldx, const float* w_obj, const float* b_obj, const float* w_box, const float* b_box,
                  float* obj, float* trans, int M, int K, cudaStream_t st) {
  if (M <= 0) return DC_OK;
  recog_heads_kernel<<<M, 128, 0, st>>>(x, ldx, w_obj, b_obj, w_box, b_box, obj, trans, K);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
