// fp32 CUDA-core engine (DC_ENGINE_SIMT) operating on the same split-fp16 data layout as the
// tcgen05 engine: bring-up path and on-device cross-check for the tensor-core kernels, plus the
// three genuinely small dense ops of the path (conv1_1 with K=27, the 4096->1 / 4096->4 heads).
#include "common.cuh"

__device__ __forceinline__ unsigned long long pack_argmax(float v, int col) {
  uint32_t u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)col);
}

// Shared epilogue for one output element.
__device__ __forceinline__ void gemm_epilogue_store(const GemmEpilogue& ep, int m, int n, float acc) {
  float v = acc;
  if (ep.bias) v += ep.bias[n];
  if (ep.relu) v = fmaxf(v, 0.f);
  if (ep.out_hi) {
    __half h, l;
    split_f32(v, h, l);
    const size_t o = (size_t)m * ep.ldo + ep.out_col0 + n;
    ep.out_hi[o] = h;
    ep.out_lo[o] = l;
  }
  if (ep.out_f32) ep.out_f32[(size_t)m * ep.ldf + n] = v;
  if (ep.argmax) atomicMax(ep.argmax + m, pack_argmax(v, n));
}

constexpr int ST_BM = 64, ST_BN = 64, ST_BK = 16;

struct ConvGeom {
  int H, W, Cin;   // A row m -> (n, y, x) with m = (n*H + y)*W + x ; k = (r*3+s)*Cin + ci
};

template <bool CONV>
__global__ void __launch_bounds__(256)
simt_gemm_kernel(const __half* __restrict__ a_hi, const __half* __restrict__ a_lo, int lda, const __half* __restrict__ b_hi,
                 const __half* __restrict__ b_lo, int ldb, int M, int N, int K, ConvGeom g, GemmEpilogue ep) {
  __shared__ float As[ST_BK][ST_BM + 1];
  __shared__ float Bs[ST_BK][ST_BN + 1];
  const int m0 = blockIdx.y * ST_BM, n0 = blockIdx.x * ST_BN;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;   // 16x16 threads, each 4x4 outputs
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += ST_BK) {
    // load A tile: 64 rows x 16 k  (1024 elements, 4 per thread)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      const int kk = idx % ST_BK, mm = idx / ST_BK;
      const int m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < M && k < K) {
        if (CONV) {
          const int ci = k % g.Cin, tap = k / g.Cin;
          const int r = tap / 3, s = tap % 3;
          const int x = m % g.W, y = (m / g.W) % g.H, n = m / (g.W * g.H);
          const int yy = y + r - 1, xx = x + s - 1;
          if (yy >= 0 && yy < g.H && xx >= 0 && xx < g.W) {
            const size_t o = (((size_t)n * g.H + yy) * g.W + xx) * g.Cin + ci;
            v = join_f32(a_hi[o], a_lo[o]);
          }
        } else {
          const size_t o = (size_t)m * lda + k;
          v = join_f32(a_hi[o], a_lo[o]);
        }
      }
      As[kk][mm] = v;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      const int kk = idx % ST_BK, nn = idx / ST_BK;
      const int n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < N && k < K) {
        const size_t o = (size_t)n * ldb + k;
        v = join_f32(b_hi[o], b_lo[o]);
      }
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < ST_BK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) gemm_epilogue_store(ep, m, n, acc[i][j]);
    }
}

int k_simt_gemm(dc_ctx* ctx, const __half* a_hi, const __half* a_lo, int lda, const __half* b_hi, const __half* b_lo, int ldb,
                int M, int N, int K, GemmEpilogue ep, cudaStream_t st) {
  if (M <= 0 || N <= 0) return DC_OK;
  dim3 grid(DC_CEIL_DIV(N, ST_BN), DC_CEIL_DIV(M, ST_BM));
  ConvGeom g{0, 0, 0};
  simt_gemm_kernel<false><<<grid, 256, 0, st>>>(a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, K, g, ep);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

int k_simt_conv3x3(dc_ctx* ctx, const __half* x_hi, const __half* x_lo, const __half* w_hi, const __half* w_lo, const float* bias,
                   __half* y_hi, __half* y_lo, int N, int H, int W, int Cin, int Cout, cudaStream_t st) {
  const int M = N * H * W, K = 9 * Cin;
  if (M <= 0) return DC_OK;
  GemmEpilogue ep;
  ep.bias = bias;
  ep.relu = 1;
  ep.out_hi = y_hi;
  ep.out_lo = y_lo;
  ep.ldo = Cout;
  dim3 grid(DC_CEIL_DIV(Cout, ST_BN), DC_CEIL_DIV(M, ST_BM));
  ConvGeom g{H, W, Cin};
  simt_gemm_kernel<true><<<grid, 256, 0, st>>>(x_hi, x_lo, 0, w_hi, w_lo, K, M, Cout, K, g, ep);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

// conv1_1 (VGG layer 1): Cin = 3, K = 27: bandwidth bound on the 64-channel output write, so a direct
// fp32 kernel: one thread = one pixel x 8 output channels; weights + bias in shared memory.
// img NCHW fp32 (reference input layout, DenseCapModel.lua:244) -> split NHWC, bias + ReLU fused.
__global__ void __launch_bounds__(256)
conv_first_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                  __half* __restrict__ y_hi, __half* __restrict__ y_lo, int N, int H, int W, int Cout) {
  extern __shared__ float sw[];   // [27][Cout] + bias[Cout]
  float* sb = sw + 27 * Cout;
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) {
    const int co = i % Cout, k = i / Cout;       // k = ci*9 + r*3 + s  (reference layout (Cout,3,3,3))
    sw[i] = w[(size_t)co * 27 + k];
  }
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sb[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int groups = Cout / 8;
  const size_t total = (size_t)N * H * W * groups;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int gq = t % groups;
    const size_t pix = t / groups;
    const int x = pix % W, y = (pix / W) % H, n = pix / ((size_t)W * H);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = sb[gq * 8 + j];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int yy = y + r - 1, xx = x + s - 1;
          float v = 0.f;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(img + (((size_t)n * 3 + ci) * H + yy) * W + xx);
          const float* wp = sw + (ci * 9 + r * 3 + s) * Cout + gq * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wp[j], acc[j]);
        }
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split_f32(fmaxf(acc[j], 0.f), h[j], l[j]);
    const size_t o = pix * Cout + gq * 8;
    *reinterpret_cast<uint4*>(y_hi + o) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(y_lo + o) = *reinterpret_cast<const uint4*>(l);
  }
}
int k_conv_first(dc_ctx* ctx, const float* img, const float* w, const float* bias, __half* y_hi, __half* y_lo, int N, int H, int W,
                 int Cout, cudaStream_t st) {
  DC_REQUIRE(ctx, Cout % 8 == 0, DC_E_UNSUPPORTED, "conv1_1: Cout must be a multiple of 8");
  const size_t total = (size_t)N * H * W * (Cout / 8);
  if (total == 0) return DC_OK;
  const size_t smem = (size_t)(27 * Cout + Cout) * sizeof(float);
  unsigned blocks = (unsigned)min((size_t)148 * 8, DC_CEIL_DIV(total, (size_t)256));
  conv_first_kernel<<<blocks, 256, smem, st>>>(img, w, bias, y_hi, y_lo, N, H, W, Cout);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

// objectness (fc_dim -> 1) and box regression (fc_dim -> 4) heads (DenseCapModel.lua:90-99): one warp per row.
__global__ void small_heads_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w, const float* __restrict__ b,
                                   float* __restrict__ y, int M, int K, int J) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float* xr = x + (size_t)warp * ldx;
  for (int j = 0; j < J; ++j) {
    const float* wr = w + (size_t)j * K;
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s = fmaf(xr[k], wr[k], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[(size_t)warp * J + j] = s + (b ? b[j] : 0.f);
  }
}
int k_small_heads(dc_ctx* ctx, const float* x, int ldx, const float* w, const float* b, float* y, int M, int K, int J, cudaStream_t st) {
  if (M <= 0 || J <= 0) return DC_OK;
  small_heads_kernel<<<DC_CEIL_DIV(M * 32, 256), 256, 0, st>>>(x, ldx, w, b, y, M, K, J);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
