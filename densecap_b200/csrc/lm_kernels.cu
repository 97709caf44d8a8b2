// Element-wise / gather kernels around the language-model GEMMs (densecap/LanguageModel.lua:293-348,
// torch-rnn nn.LSTM step) and the final result gathers (DenseCapModel.lua:272-274).
#include "common.cuh"

__global__ void fill_u32_kernel(uint32_t* p, uint32_t v, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
int k_fill_u32(dc_ctx* ctx, uint32_t* p, uint32_t v, size_t n, cudaStream_t st) {
  if (n == 0) return DC_OK;
  fill_u32_kernel<<<(unsigned)min((size_t)1024, DC_CEIL_DIV(n, (size_t)256)), 256, 0, st>>>(p, v, n);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
// Apply the final-NMS selection to output[4], output[1] and -- when seq is given -- output[5] (DenseCapModel.lua:272-274).
__global__ void final_gather_kernel(const int32_t* __restrict__ idx32, const int32_t* __restrict__ count, int P,
                                    const float4* __restrict__ boxes, const float* __restrict__ scores, const int64_t* __restrict__ seq,
                                    int T, float4* __restrict__ out_boxes, float* __restrict__ out_scores, int64_t* __restrict__ out_seq) {
  const int i = blockIdx.x;
  const int n = min(*count, P);
  if (i >= n) {   // clear the tail so stale rows never leak
    if (threadIdx.x == 0) { out_boxes[i] = make_float4(0, 0, 0, 0); out_scores[i] = 0.f; }
    if (seq) for (int t = threadIdx.x; t < T; t += blockDim.x) out_seq[(size_t)i * T + t] = 0;
    return;
  }
  const int src = idx32[i];
  if (threadIdx.x == 0) { out_boxes[i] = boxes[src]; out_scores[i] = scores[src]; }
  if (seq) for (int t = threadIdx.x; t < T; t += blockDim.x) out_seq[(size_t)i * T + t] = seq[(size_t)src * T + t];
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
