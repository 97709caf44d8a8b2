// LanguageModel:beamsearch (densecap/LanguageModel.lua:170-290) -- the element-wise / selection kernels around the gate and
// vocabulary GEMMs.  The reference runs one image vector at a time with the beams in the minibatch dimension; here all ROIs of a
// chunk advance together: row = roi * beam + b.  Semantics kept (and restated by oracle.lm_beamsearch):
//   * step 0 (START token): top-beam words of log_softmax(scores) start the beams, the LSTM state is replicated (:216-227);
//   * step t: per beam the top-beam next words (:249-251); a beam that already contains END contributes zeros (:243-247, the
//     log-probabilities are MULTIPLIED by the mask, so its candidates are words 1..beam with value -0);
//     the beam*beam sums beam_logprob + word_logprob are ranked again and the beam best survive with their parent's state (:253-279);
//   * the best beam at the end is the output row (:283-284).
//   * torch.topk leaves the order of exactly equal values unspecified; like the oracle, equal values rank lower index first.
// log_softmax = (x - max) - float(log(sum exp(double(x - max)))): nn.LogSoftMax accumulates in double (THNN accreal).
#include "common.cuh"

namespace {

struct BestPair { float v; int i; };
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }
__device__ __forceinline__ BestPair warp_best(BestPair p) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, p.v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, p.i, o);
    if (better(ov, oi, p.v, p.i)) { p.v = ov; p.i = oi; }
  }
  return p;
}
// block-wide (value desc, index asc) best; result valid in every thread.  red: 2 x 32 words of shared memory
__device__ __forceinline__ BestPair block_best(BestPair p, float* red_v, int* red_i) {
  p = warp_best(p);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) { red_v[warp] = p.v; red_i[warp] = p.i; }
  __syncthreads();
  BestPair q;
  q.v = (lane < nw) ? red_v[lane] : -INFINITY;
  q.i = (lane < nw) ? red_i[lane] : 0x7fffffff;
  return warp_best(q);
}

constexpr int BS_THREADS = 256;

// log_softmax of one row into shared memory `lp` (V1 floats); returns nothing.  scratch: 32 doubles + 32 floats
__device__ void row_log_softmax(const float* __restrict__ logits, int V1, float* lp, float* red_f, double* red_d) {
  float m = -INFINITY;
  for (int j = threadIdx.x; j < V1; j += blockDim.x) { const float x = logits[j]; lp[j] = x; m = fmaxf(m, x); }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) red_f[warp] = m;
  __syncthreads();
  m = (lane < nw) ? red_f[lane] : -INFINITY;
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  double s = 0.0;
  for (int j = threadIdx.x; j < V1; j += blockDim.x) { const float d = __fsub_rn(lp[j], m); lp[j] = d; s += exp((double)d); }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red_d[warp] = s;
  __syncthreads();
  s = (lane < nw) ? red_d[lane] : 0.0;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float lse = (float)log(s);
  for (int j = threadIdx.x; j < V1; j += blockDim.x) lp[j] = __fsub_rn(lp[j], lse);
  __syncthreads();
}

// top-k of lp[0..n) (destroys the selected entries), (value desc, index asc); thread 0 writes out_v / out_i
__device__ void block_topk(float* lp, int n, int k, float* out_v, int* out_i, float* red_v, int* red_i) {
  for (int r = 0; r < k; ++r) {
    BestPair p{-INFINITY, 0x7fffffff};
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const float v = lp[j];
      if (better(v, j, p.v, p.i)) { p.v = v; p.i = j; }
    }
    p = block_best(p, red_v, red_i);
    if (threadIdx.x == 0) {
      out_v[r] = p.v; out_i[r] = p.i;
      if (p.i < n) lp[p.i] = -INFINITY;       // NaN-free rows always find an index; -inf never wins again against finite values
    }
    __syncthreads();
  }
}

// Step 0 for one ROI per block: beams from the START-token scores; LSTM state replicated for all beams.
__global__ void __launch_bounds__(BS_THREADS)
bs_first_kernel(const float* __restrict__ logits, int V1, int Bm, int T, int H, float* __restrict__ beam_lp, int32_t* __restrict__ beams,
                int32_t* __restrict__ tokens, const __half* __restrict__ h_hi, const __half* __restrict__ h_lo, const float* __restrict__ c,
                __half* __restrict__ hd_hi, __half* __restrict__ hd_lo, float* __restrict__ cd) {
  extern __shared__ float lp[];
  __shared__ float red_v[32]; __shared__ int red_i[32]; __shared__ double red_d[32];
  __shared__ float top_v[32]; __shared__ int top_i[32];
  const int r = blockIdx.x;
  row_log_softmax(logits + (size_t)r * V1, V1, lp, red_v, red_d);
  block_topk(lp, V1, Bm, top_v, top_i, red_v, red_i);
  for (int b = threadIdx.x; b < Bm; b += blockDim.x) {
    beam_lp[(size_t)r * Bm + b] = top_v[b];
    tokens[(size_t)r * Bm + b] = top_i[b] + 1;
  }
  for (int e = threadIdx.x; e < Bm * T; e += blockDim.x) {
    const int b = e / T, t = e % T;
    beams[((size_t)r * Bm + b) * T + t] = (t == 0) ? top_i[b] + 1 : 1;                       // beams filled with 1 (:213)
  }
  for (int e = threadIdx.x; e < Bm * H; e += blockDim.x) {
    const int b = e / H, j = e % H;
    const size_t d = ((size_t)r * Bm + b) * H + j, s = (size_t)r * H + j;
    hd_hi[d] = h_hi[s]; hd_lo[d] = h_lo[s]; cd[d] = c[s];
  }
}

// Step t, one beam row per block: top-beam next words of the masked log-probabilities.
__global__ void __launch_bounds__(BS_THREADS)
bs_row_topk_kernel(const float* __restrict__ logits, int V1, int Bm, int T, int end_token, const int32_t* __restrict__ beams,
                   float* __restrict__ cand_lp, int32_t* __restrict__ cand_ix) {
  extern __shared__ float lp[];
  __shared__ float red_v[32]; __shared__ int red_i[32]; __shared__ double red_d[32];
  __shared__ float top_v[32]; __shared__ int top_i[32];
  __shared__ int s_has_end;
  const int row = blockIdx.x;
  if (threadIdx.x == 0) {
    int e = 0;
    for (int t = 0; t < T; ++t) e |= (beams[(size_t)row * T + t] == end_token);
    s_has_end = e;
  }
  row_log_softmax(logits + (size_t)row * V1, V1, lp, red_v, red_d);
  if (s_has_end) {
    // next_word_logprobs:cmul(0): every entry becomes (-)0, topk of equal values = the first Bm indices
    for (int k = threadIdx.x; k < Bm; k += blockDim.x) { cand_lp[(size_t)row * Bm + k] = __fmul_rn(lp[k], 0.f); cand_ix[(size_t)row * Bm + k] = k; }
    return;
  }
  block_topk(lp, V1, Bm, top_v, top_i, red_v, red_i);
  for (int k = threadIdx.x; k < Bm; k += blockDim.x) { cand_lp[(size_t)row * Bm + k] = top_v[k]; cand_ix[(size_t)row * Bm + k] = top_i[k]; }
}

// Step t, one ROI per block: rank the Bm*Bm extensions, keep the Bm best with their parents.
__global__ void __launch_bounds__(BS_THREADS)
bs_merge_kernel(int Bm, int T, int t, const float* __restrict__ cand_lp, const int32_t* __restrict__ cand_ix, float* __restrict__ beam_lp,
                const int32_t* __restrict__ beams_in, int32_t* __restrict__ beams_out, int32_t* __restrict__ src_row, int32_t* __restrict__ tokens) {
  __shared__ float all_next[1024];
  __shared__ float red_v[32]; __shared__ int red_i[32];
  __shared__ float top_v[32]; __shared__ int top_i[32];
  const int r = blockIdx.x, n = Bm * Bm;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int b = e / Bm;
    all_next[e] = __fadd_rn(cand_lp[(size_t)r * n + e], beam_lp[(size_t)r * Bm + b]);
  }
  __syncthreads();
  block_topk(all_next, n, Bm, top_v, top_i, red_v, red_i);
  for (int b = threadIdx.x; b < Bm; b += blockDim.x) {
    const int sel = top_i[b], parent = sel / Bm;
    beam_lp[(size_t)r * Bm + b] = top_v[b];
    src_row[(size_t)r * Bm + b] = r * Bm + parent;
    tokens[(size_t)r * Bm + b] = cand_ix[(size_t)r * n + sel] + 1;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < Bm * T; e += blockDim.x) {
    const int b = e / T, tt = e % T;
    const int sel = top_i[b], parent = sel / Bm;
    beams_out[((size_t)r * Bm + b) * T + tt] = (tt == t) ? cand_ix[(size_t)r * n + sel] + 1 : beams_in[((size_t)r * Bm + parent) * T + tt];
  }
}

// state of every surviving beam = state of its parent (:263-278)
__global__ void bs_reindex_kernel(int rows, int H, const int32_t* __restrict__ src_row, const __half* __restrict__ hi_in, const __half* __restrict__ lo_in,
                                  const float* __restrict__ c_in, __half* __restrict__ hi_out, __half* __restrict__ lo_out, float* __restrict__ c_out) {
  const int row = blockIdx.x;
  const int src = src_row[row];
  for (int j = threadIdx.x * 8; j < H; j += blockDim.x * 8) {
    const size_t d = (size_t)row * H + j, s = (size_t)src * H + j;
    *reinterpret_cast<uint4*>(hi_out + d) = *reinterpret_cast<const uint4*>(hi_in + s);
    *reinterpret_cast<uint4*>(lo_out + d) = *reinterpret_cast<const uint4*>(lo_in + s);
    *reinterpret_cast<float4*>(c_out + d) = *reinterpret_cast<const float4*>(c_in + s);
    *reinterpret_cast<float4*>(c_out + d + 4) = *reinterpret_cast<const float4*>(c_in + s + 4);
  }
}

// best beam of every ROI -> output row (:283-284; max returns the first maximal index)
__global__ void bs_final_kernel(int R, int Bm, int T, const float* __restrict__ beam_lp, const int32_t* __restrict__ beams, const int32_t* __restrict__ dst_row,
                                int64_t* __restrict__ seq) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  int best = 0;
  float bv = beam_lp[(size_t)r * Bm];
  for (int b = 1; b < Bm; ++b) { const float v = beam_lp[(size_t)r * Bm + b]; if (v > bv) { bv = v; best = b; } }
  const size_t d = (size_t)(dst_row ? dst_row[r] : r) * T;
  for (int t = 0; t < T; ++t) seq[d + t] = beams[((size_t)r * Bm + best) * T + t];
}

__global__ void bs_fill_i32_kernel(int32_t* p, int32_t v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
// rows of a split (.., E) tensor selected by an index list -> compact planes
__global__ void bs_gather_split_kernel(const __half* __restrict__ hi_in, const __half* __restrict__ lo_in, int E, const int32_t* __restrict__ rows_src,
                                       int first, __half* __restrict__ hi_out, __half* __restrict__ lo_out) {
  const int r = blockIdx.x;
  const int src = rows_src ? rows_src[first + r] : first + r;
  for (int j = threadIdx.x * 8; j < E; j += blockDim.x * 8) {
    *reinterpret_cast<uint4*>(hi_out + (size_t)r * E + j) = *reinterpret_cast<const uint4*>(hi_in + (size_t)src * E + j);
    *reinterpret_cast<uint4*>(lo_out + (size_t)r * E + j) = *reinterpret_cast<const uint4*>(lo_in + (size_t)src * E + j);
  }
}

}  // namespace

int k_bs_first(dc_ctx* ctx, const float* logits, int R, int V1, int Bm, int T, int H, float* beam_lp, int32_t* beams, int32_t* tokens,
               const __half* h_hi, const __half* h_lo, const float* c, __half* hd_hi, __half* hd_lo, float* cd, cudaStream_t st) {
  if (R <= 0) return DC_OK;
  const size_t smem = (size_t)V1 * sizeof(float);
  cudaError_t e = dc_ensure_smem(ctx, bs_first_kernel, smem);
  if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "beam search smem attribute: %s", cudaGetErrorString(e));
  bs_first_kernel<<<R, BS_THREADS, smem, st>>>(logits, V1, Bm, T, H, beam_lp, beams, tokens, h_hi, h_lo, c, hd_hi, hd_lo, cd);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
int k_bs_row_topk(dc_ctx* ctx, const float* logits, int rows, int V1, int Bm, int T, int end_token, const int32_t* beams, float* cand_lp,
                  int32_t* cand_ix, cudaStream_t st) {
  if (rows <= 0) return DC_OK;
  const size_t smem = (size_t)V1 * sizeof(float);
  cudaError_t e = dc_ensure_smem(ctx, bs_row_topk_kernel, smem);
  if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "beam search smem attribute: %s", cudaGetErrorString(e));
  bs_row_topk_kernel<<<rows, BS_THREADS, smem, st>>>(logits, V1, Bm, T, end_token, beams, cand_lp, cand_ix);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
int k_bs_merge(dc_ctx* ctx, int R, int Bm, int T, int t, const float* cand_lp, const int32_t* cand_ix, float* beam_lp, const int32_t* beams_in,
               int32_t* beams_out, int32_t* src_row, int32_t* tokens, cudaStream_t st) {
  if (R <= 0) return DC_OK;
  bs_merge_kernel<<<R, BS_THREADS, 0, st>>>(Bm, T, t, cand_lp, cand_ix, beam_lp, beams_in, beams_out, src_row, tokens);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
int k_bs_reindex(dc_ctx* ctx, int rows, int H, const int32_t* src_row, const __half* hi_in, const __half* lo_in, const float* c_in, __half* hi_out,
                 __half* lo_out, float* c_out, cudaStream_t st) {
  if (rows <= 0) return DC_OK;
  bs_reindex_kernel<<<rows, 64, 0, st>>>(rows, H, src_row, hi_in, lo_in, c_in, hi_out, lo_out, c_out);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
int k_bs_final(dc_ctx* ctx, int R, int Bm, int T, const float* beam_lp, const int32_t* beams, const int32_t* dst_row, int64_t* seq, cudaStream_t st) {
  if (R <= 0) return DC_OK;
  bs_final_kernel<<<DC_CEIL_DIV(R, 128), 128, 0, st>>>(R, Bm, T, beam_lp, beams, dst_row, seq);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
int k_bs_fill_i32(dc_ctx* ctx, int32_t* p, int32_t v, int n, cudaStream_t st) {
  if (n <= 0) return DC_OK;
  bs_fill_i32_kernel<<<DC_CEIL_DIV(n, 256), 256, 0, st>>>(p, v, n);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
int k_bs_gather_split(dc_ctx* ctx, const __half* hi_in, const __half* lo_in, int E, const int32_t* rows_src, int first, int R, __half* hi_out,
                      __half* lo_out, cudaStream_t st) {
  if (R <= 0) return DC_OK;
  bs_gather_split_kernel<<<R, 64, 0, st>>>(hi_in, lo_in, E, rows_src, first, hi_out, lo_out);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
