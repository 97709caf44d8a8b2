// Element-wise / gather kernels around the language-model GEMMs (densecap/LanguageModel.lua:293-348,
// torch-rnn nn.LSTM step) and the final result gathers (DenseCapModel.lua:272-274).
#include "common.cuh"

// One launch per decode step t in [0, T]:
//   t > 0 : finish step t-1 -- torch.max(scores, 2) (LanguageModel.lua:330): first maximal index, 1-based -> seq[:, t-1];
//   t < T : nn.LookupTable row of the input token (START = V+1 when t == 0, else the token just sampled;
//           LanguageModel.lua:318-326) -> x half of the [x | h] operand of the gate GEMM, and reset of the
//           packed arg-max cell for the coming vocabulary projection.
__global__ void __launch_bounds__(64)
lm_token_step_kernel(unsigned long long* __restrict__ packed, int32_t* __restrict__ tokens, int64_t* __restrict__ seq64, int T, int t,
                     int start_token, const __half* __restrict__ lut_hi, const __half* __restrict__ lut_lo, int E,
                     __half* __restrict__ xh_hi, __half* __restrict__ xh_lo, int ld) {
  // launched with programmatic stream serialisation between the vocabulary GEMM and the next gate GEMM: wait for the
  // arg-max cells of the former, then let the latter start its prologue while the rows are copied
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int m = blockIdx.x;
  __shared__ int s_tok;
  if (threadIdx.x == 0) {
    int tok = start_token;
    if (t > 0) {
      const unsigned long long pk = packed[m];
      const uint32_t col = 0xFFFFFFFFu - (uint32_t)(pk & 0xFFFFFFFFull);
      tok = (pk == 0ull) ? 1 : (int)min(col, (uint32_t)start_token - 1u) + 1;   // all-NaN row -> first index, like torch.max
      tokens[(size_t)m * T + (t - 1)] = tok;
      seq64[(size_t)m * T + (t - 1)] = tok;
    }
    packed[m] = 0ull;
    s_tok = tok;
  }
  __syncthreads();
  if (t >= T) return;
  const size_t src = (size_t)(s_tok - 1) * E;
  const size_t dst = (size_t)m * ld;
  for (int e = threadIdx.x * 8; e < E; e += 64 * 8) {
    *reinterpret_cast<uint4*>(xh_hi + dst + e) = __ldg(reinterpret_cast<const uint4*>(lut_hi + src + e));
    *reinterpret_cast<uint4*>(xh_lo + dst + e) = __ldg(reinterpret_cast<const uint4*>(lut_lo + src + e));
  }
}
int k_lm_token_step(dc_ctx* ctx, unsigned long long* packed, int32_t* tokens, int64_t* seq64, int T, int t, int start_token,
                    const __half* lut_hi, const __half* lut_lo, int E, __half* xh_hi, __half* xh_lo, int ld, int M, cudaStream_t st) {
  if (M <= 0) return DC_OK;
  DC_REQUIRE(ctx, E % 8 == 0 && ld % 8 == 0, DC_E_UNSUPPORTED, "lm: encoding size must be a multiple of 8");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(M);
  cfg.blockDim = dim3(64);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, lm_token_step_kernel, packed, tokens, seq64, T, t, start_token, lut_hi, lut_lo, E, xh_hi, xh_lo, ld);
  if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "lm_token_step launch: %s", cudaGetErrorString(e));
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// gates (M,4H) = bias + x Wx + h Wh (bias added by the GEMM epilogue), INTERLEAVED columns 4*j + {i,f,o,g}
// (the weight rows are permuted at load time so that the tcgen05 engine can fuse this cell into its epilogue;
// this stand-alone kernel serves the CUDA-core engine).
// c = f*c + i*g ; h = o*tanh(c).  h is written (split) into columns [col0, col0+H) of the [x|h] operand.
__global__ void lstm_pointwise_kernel(const float* __restrict__ gates, float* __restrict__ c, __half* __restrict__ xh_hi,
                                      __half* __restrict__ xh_lo, int ld, int col0, int M, int H, int first) {
  const size_t total = (size_t)M * H;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int j = idx % H;
    const size_t m = idx / H;
    const float* g = gates + m * 4 * (size_t)H;
    const float4 a = *reinterpret_cast<const float4*>(g + 4 * j);
    const float ig = sigmoidf_(a.x);
    const float fg = sigmoidf_(a.y);
    const float og = sigmoidf_(a.z);
    const float gg = tanhf(a.w);
    const float cp = first ? 0.f : c[idx];
    const float cn = fg * cp + ig * gg;
    c[idx] = cn;
    const float h = og * tanhf(cn);
    __half hh, hl;
    split_f32(h, hh, hl);
    xh_hi[m * ld + col0 + j] = hh;
    xh_lo[m * ld + col0 + j] = hl;
  }
}
int k_lstm_pointwise(dc_ctx* ctx, const float* gates, float* c, __half* xh_hi, __half* xh_lo, int ld, int col0, int M, int H,
                     int first, cudaStream_t st) {
  const size_t total = (size_t)M * H;
  if (total == 0) return DC_OK;
  lstm_pointwise_kernel<<<(unsigned)min((size_t)148 * 8, DC_CEIL_DIV(total, (size_t)256)), 256, 0, st>>>(gates, c, xh_hi, xh_lo, ld, col0,
                                                                                                      M, H, first);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

__global__ void fill_u32_kernel(uint32_t* p, uint32_t v, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
int k_fill_u32(dc_ctx* ctx, uint32_t* p, uint32_t v, size_t n, cudaStream_t st) {
  if (n == 0) return DC_OK;
  fill_u32_kernel<<<(unsigned)min((size_t)1024, DC_CEIL_DIV(n, (size_t)256)), 256, 0, st>>>(p, v, n);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
// Apply the final-NMS selection to output[4], output[1], output[5] (DenseCapModel.lua:272-274).
__global__ void final_gather_kernel(const int32_t* __restrict__ idx32, const int32_t* __restrict__ count, int P,
                                    const float4* __restrict__ boxes, const float* __restrict__ scores, const int64_t* __restrict__ seq,
                                    int T, float4* __restrict__ out_boxes, float* __restrict__ out_scores, int64_t* __restrict__ out_seq) {
  const int i = blockIdx.x;
  const int n = min(*count, P);
  if (i >= n) {   // clear the tail so stale rows never leak
    if (threadIdx.x == 0) { out_boxes[i] = make_float4(0, 0, 0, 0); out_scores[i] = 0.f; }
    for (int t = threadIdx.x; t < T; t += blockDim.x) out_seq[(size_t)i * T + t] = 0;
    return;
  }
  const int src = idx32[i];
  if (threadIdx.x == 0) { out_boxes[i] = boxes[src]; out_scores[i] = scores[src]; }
  for (int t = threadIdx.x; t < T; t += blockDim.x) out_seq[(size_t)i * T + t] = seq[(size_t)src * T + t];
}
int k_final_gather(dc_ctx* ctx, const int32_t* idx32, const int32_t* count, int P, const float* boxes, const float* scores,
                   const int64_t* seq, int T, float* out_boxes, float* out_scores, int64_t* out_seq, cudaStream_t st) {
  if (P <= 0) return DC_OK;
  final_gather_kernel<<<P, 32, 0, st>>>(idx32, count, P, (const float4*)boxes, scores, seq, T, (float4*)out_boxes, out_scores, out_seq);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

// dst[i,:] = src[idx[i],:] for i < count, zeros for the rest (index gathers, LocalizationLayer.lua:334-338)
__global__ void gather_rows_kernel(const float* __restrict__ src, int cols, const int32_t* __restrict__ idx32,
                                   const int32_t* __restrict__ count, int max_rows, float* __restrict__ dst) {
  const int i = blockIdx.x;
  const int n = count ? min(*count, max_rows) : max_rows;
  for (int c = threadIdx.x; c < cols; c += blockDim.x)
    dst[(size_t)i * cols + c] = (i < n) ? src[(size_t)idx32[i] * cols + c] : 0.f;
}
int k_gather_rows_f32(dc_ctx* ctx, const float* src, int cols, const int32_t* idx32, const int32_t* count, int max_rows, float* dst,
                      cudaStream_t st) {
  if (max_rows <= 0) return DC_OK;
  gather_rows_kernel<<<max_rows, 32, 0, st>>>(src, cols, idx32, count, max_rows, dst);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
