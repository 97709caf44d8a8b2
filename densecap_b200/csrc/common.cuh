// Internal shared declarations for libdensecap_b200.so (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <set>
#include <string>
#include <vector>

#include "../../include/densecap_b200.h"

// ---------------------------------------------------------------------------------------
// Activation / operand format used by every dense kernel ("split-fp16 planes"):
// a fp32 value a is stored as two fp16 numbers hi = rn(a), lo = rn(a - hi); a tensor is two
// planes [hi plane][lo plane] of identical shape (same bytes as fp32).  a*b is evaluated on
// the tensor cores as hi*hi + hi*lo + lo*hi with fp32 accumulation (error ~2^-21 relative),
// which is what lets tcgen05 kind::f16 reproduce the reference's fp32 sgemm results.
// ---------------------------------------------------------------------------------------
struct SplitTensor {
  __half* hi = nullptr;   // plane 0
  __half* lo = nullptr;   // plane 1 (== hi + plane_elems)
  size_t plane_elems = 0;
};

// cvt.rn.satfinite saturates to +-65504 in the conversion itself (no inf planes, no extra min/max instructions)
__device__ __forceinline__ __half f32_to_f16_sat(float a) {
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(a));
  return __ushort_as_half(r);
}
__device__ __forceinline__ void split_f32(float a, __half& hi, __half& lo) {
  hi = f32_to_f16_sat(a);
  lo = f32_to_f16_sat(a - __half2float(hi));
}
__device__ __forceinline__ float join_f32(__half hi, __half lo) {
  return __half2float(hi) + __half2float(lo);
}

#define DC_CEIL_DIV(a, b) (((a) + (b) - 1) / (b))

struct dc_ctx;
int dc_set_error(dc_ctx* ctx, int code, const char* fmt, ...);
void dc_count_launch(dc_ctx* ctx, int n = 1);

// Launcher state owned by ONE context (a process may hold several contexts on different GPUs; nothing here is process-global):
// the driver entry point for tensor-map encoding, the SM count of the context's device, and which kernels already had their
// dynamic shared-memory limit raised on that device.
struct LaunchState {
  void* encode_tiled = nullptr;            // cuTensorMapEncodeTiled (resolved by tc_init)
  int num_sms = 148;
  // k-blocks (64 K each) a tcgen05 tile accumulates in TMEM before the epilogue warps take the partial sum over (hierarchical
  // accumulation, tc_common.cuh): index 0 / 1 / 2 = 64- / 128- / 256-column tiles.  Smaller = closer to fp32 sgemm, more TMEM sweeps.
  int group_kb[3] = {2, 2, 1};
  std::set<const void*> smem_attr_done;    // kernel function pointers
};
LaunchState* dc_launch_state(dc_ctx* ctx);
// fp32 workspace of the split-K convolutions (partial-sum planes); null when the context has none
float* dc_splitk_workspace(dc_ctx* ctx, size_t* elems);
// sticky device flag: some split-fp16 activation left the fp16 range (see GemmEpilogue::overflow); null without a context
unsigned int* dc_overflow_flag(dc_ctx* ctx);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (context, kernel)
template <typename Kernel>
static inline cudaError_t dc_ensure_smem(dc_ctx* ctx, Kernel kernel, size_t bytes) {
  LaunchState* ls = dc_launch_state(ctx);
  const void* key = reinterpret_cast<const void*>(kernel);
  if (ls->smem_attr_done.count(key)) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) ls->smem_attr_done.insert(key);
  return e;
}

#define DC_CHECK_CUDA(ctx, expr)                                                            \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return dc_set_error((ctx), DC_E_CUDA, "%s failed: %s (%s:%d)", #expr,                 \
                          cudaGetErrorString(_e), __FILE__, __LINE__);                      \
  } while (0)

#define DC_CHECK_LAUNCH(ctx)                                                                \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess)                                                                  \
      return dc_set_error((ctx), DC_E_CUDA, "kernel launch failed: %s (%s:%d)",             \
                          cudaGetErrorString(_e), __FILE__, __LINE__);                      \
    dc_count_launch(ctx);                                                                   \
  } while (0)

#define DC_REQUIRE(ctx, cond, code, ...)                                                    \
  do {                                                                                      \
    if (!(cond)) return dc_set_error((ctx), (code), __VA_ARGS__);                           \
  } while (0)

// ---------------------------------------------------------------------------------------
// Kernel launchers implemented across the .cu files (all asynchronous on `st`).
// ---------------------------------------------------------------------------------------
// box_kernels.cu
int k_make_anchors(dc_ctx*, float* out, int N, int H, int W, int k, float x0, float y0, float sx, float sy,
                   const float* anchors_dev, cudaStream_t st);
int k_reshape_box_features(dc_ctx*, const float* in, float* out, int N, int k, int D, int H, int W, cudaStream_t st);
int k_apply_box_transform(dc_ctx*, const float* boxes, const float* trans, float* out, int n, cudaStream_t st);
int k_invert_box_transform(dc_ctx*, const float* anchors, const float* targets, float* out, int n, cudaStream_t st);
int k_xcycwh_to_x1y1x2y2(dc_ctx*, const float* in, float* out, int n, cudaStream_t st);
int k_x1y1x2y2_to_xcycwh(dc_ctx*, const float* in, float* out, int n, cudaStream_t st);
int k_clip_boxes(dc_ctx*, const float* boxes, float* out, uint8_t* valid, int n, float x_min, float y_min,
                 float x_max, float y_max, cudaStream_t st);
int k_objectness_prob(dc_ctx*, const float* scores, float* prob, int n, cudaStream_t st);
int k_box_to_affine(dc_ctx*, const float* boxes, float* theta, int B, int H, int W, cudaStream_t st);
int k_box_iou(dc_ctx*, const float* b1, const float* b2, float* out, int B1, int B2, cudaStream_t st);
int k_affine_grid(dc_ctx*, const float* theta, int B, int HH, int WW, float* grids, cudaStream_t st);

// Fused RPN post-processing (MakeAnchors + ReshapeBoxFeatures + ApplyBoxTransform + clip_boxes +
// objectness prob + clamp_data compaction).  head: (HW, ld) fp32 rows = pixels, columns
// [0,4k) box deltas (4a+d), [4k,6k) scores (2a+d).  Outputs are compacted over valid boxes, in
// row order r=(a*H+y)*W+x.  n_valid (1) int32.
struct RpnPostOut {
  float* boxes_xcycwh;   // (A,4)
  float* boxes_x1y1x2y2; // (A,4)
  float* prob;           // (A)
  int32_t* src_row;      // (A) original row r of each compacted box (0-based)
  int32_t* n_valid;      // (1)
  int32_t* block_counts; // scratch (>= number of blocks)
};
int k_rpn_postprocess(dc_ctx*, const float* head, int ld, int H, int W, int k, const float* anchors_dev,
                      float x0, float y0, float sx, float sy, float img_w, float img_h, int clip,
                      RpnPostOut out, cudaStream_t st);

// nms_kernels.cu
struct NmsWorkspace {
  uint64_t* keys;        // (cap) sort keys
  int32_t* order;        // (cap) sorted position -> box index (valid below NmsState::sorted_n)
  float* sorted_boxes;   // (cap,4) x1y1x2y2 in sorted order
  float* kept_boxes;     // (cap,4) boxes kept so far, in pick order
  uint64_t* mask;        // (1024, 16) suppression bitmask of the current super-chunk
  uint64_t* mask_prev;   // (cap, 16)  kept-so-far x current super-chunk
  void* state;           // NmsState
  int cap;               // max boxes
};
size_t nms_workspace_bytes(int cap);
void nms_workspace_carve(NmsWorkspace& ws, void* base, int cap);
// boxes (n,4) x1y1x2y2 and scores (n) may come from a strided (n,5) array: pass strides in floats.
// n_dev (optional): device int32 holding the actual n (<= n_max) -- used after compaction.
int k_nms(dc_ctx*, const float* boxes, int box_stride, const float* scores, int score_stride, int n_max,
          const int32_t* n_dev, float thresh, int max_boxes, NmsWorkspace ws, int64_t* idx_out,
          int32_t* idx32_out, int32_t* count_out, cudaStream_t st);

// roi_pool.cu
// feats: split NHWC planes (hi, lo) of one image (H',W',C) or fp32 HWC (feats_f32 != nullptr).
// taps_ws: scratch of roi_taps_bytes(max_boxes, HH*WW) bytes.
size_t roi_taps_bytes(int max_boxes, int npos);
int k_roi_pool(dc_ctx*, const __half* feat_hi, const __half* feat_lo, const float* feats_f32, int C, int fh, int fw,
               const float* boxes, const int32_t* nboxes_dev, int max_boxes, int img_h, int img_w, int HH, int WW,
               void* taps_ws, __half* out_hi, __half* out_lo, float* out_f32, int out_layout, cudaStream_t st);

// nn.BatchBilinearSamplerBHWD on caller-supplied grids: feats (H',W',C) fp32, grids (B,HH,WW,2) -> out (B,HH,WW,C)
int k_bilinear_sampler(dc_ctx*, const float* feats_hwc, int C, int fh, int fw, const float* grids, int B, int HH, int WW, void* taps_ws, float* out,
                       cudaStream_t st);

// layout_kernels.cu
int k_image_scale_preprocess(dc_ctx*, const uint8_t* in_hwc, int H, int W, int oh, int ow, float* tmp, float* out_nchw, cudaStream_t st);
int k_nchw_to_split_nhwc(dc_ctx*, const float* in, __half* hi, __half* lo, int N, int C, int H, int W, cudaStream_t st);
int k_split_nhwc_to_nchw(dc_ctx*, const __half* hi, const __half* lo, float* out, int N, int C, int H, int W, cudaStream_t st);
// scale: values are multiplied by it (a power of two) before the split -- used for weights, see weight_scale()
int k_rowmajor_to_split(dc_ctx*, const float* in, __half* hi, __half* lo, size_t n, cudaStream_t st, float scale = 1.f);
int k_absmax(dc_ctx*, const float* in, size_t n, float* out_dev, cudaStream_t st);   // out_dev must be zeroed
int k_split_to_rowmajor(dc_ctx*, const __half* hi, const __half* lo, float* out, size_t n, cudaStream_t st);
int k_maxpool2x2_split(dc_ctx*, const __half* ihi, const __half* ilo, __half* ohi, __half* olo, int N, int H, int W, int C,
                       int ceil_mode, cudaStream_t st);
// weight re-layout: conv (Cout,Cin,3,3) -> [Cout][(r*3+s)*Cin + ci] split; fc6 (out, C*HH*WW) (c,y,x) -> (y,x,c);
// lstm ((D+H),4H) -> [4H][D+H]
int k_conv_weight_to_split(dc_ctx*, const float* w, __half* hi, __half* lo, int Cout, int Cin, cudaStream_t st, float scale = 1.f);
// Order of the (channel, position) pairs inside one row of the pooled ROI features = the K order of fc6's A operand.
// The pooling kernel gives every CTA one 32-channel slice of the feature map, so a row is laid out slice-major:
// [slice s][position p][32 channels] -- the 49 positions x 32 channels x 2 B a CTA produces for one ROI are then 3136 contiguous
// bytes per plane and a warp's store of 8 positions covers four full 128-byte lines (with plain (p, c) order every store was a
// 64-byte fragment of a 1 KB row).  fc6's weight columns are permuted to the same order once at load time.
__host__ __device__ inline size_t roi_feature_index(int c, int pos, int C, int hw) {
  const int cs = (C % 32 == 0) ? 32 : C;
  return ((size_t)(c / cs) * hw + pos) * cs + (size_t)(c % cs);
}
int k_fc6_weight_to_split(dc_ctx*, const float* w, __half* hi, __half* lo, int out, int C, int hw, cudaStream_t st, float scale = 1.f);
// interleave_h > 0: output row of input column c is 4*(c % H) + c / H (LSTM gate interleave, H = interleave_h)
int k_transpose_to_split(dc_ctx*, const float* w, __half* hi, __half* lo, int rows, int cols, int interleave_h, cudaStream_t st, float scale = 1.f);
// BCHW (K,C,HH,WW) fp32 -> split rows (K, HH*WW*C) in (y,x,c) order
int k_bchw_to_split_rows(dc_ctx*, const float* in, __half* hi, __half* lo, int K, int C, int hw, cudaStream_t st);
int k_bhwc_split_to_bchw(dc_ctx*, const __half* hi, const __half* lo, float* out, int K, int C, int hw, cudaStream_t st);

// epilogue description shared by the GEMM launchers
struct GemmEpilogue {
  float acc_scale = 1.f;          // accumulator multiplier (exact power of two undoing the weight pre-scaling)
  const float* bias = nullptr;    // (N) or null
  int relu = 0;
  __half* out_hi = nullptr;       // split output planes (M, ldo) or null
  __half* out_lo = nullptr;
  int ldo = 0;                    // row stride (elements) of split output
  int out_col0 = 0;               // column offset inside the split output row
  float* out_f32 = nullptr;       // fp32 output (M, ldf) or null
  int ldf = 0;
  unsigned long long* argmax = nullptr;  // (M) packed (sortable(value)<<32 | ~col) atomicMax target or null
  // Fused torch-rnn LSTM cell (tcgen05 engine): the N columns are the gate pre-activations in INTERLEAVED order
  // n = 4*j + {i,f,o,g}; the epilogue applies the cell, updates c (M,H) and writes h (split) to out_hi/out_lo.
  float* lstm_c = nullptr;
  int lstm_H = 0;
  int lstm_first = 0;             // 1: c_{t-1} = 0
  // LSTM epilogue only: instead of `bias`, add row gather_idx[m] (1-based token) of a precomputed fp32 table (gather_ld columns)
  // = b + lookup[token] . Wx  (the input half of the gate pre-activations, see lm_decode.cu)
  const float* gather_tab = nullptr;
  const int32_t* gather_idx = nullptr;
  int gather_ld = 0;
  // Range guard of the split-fp16 planes: an fp32 value with |x| >= 65504 (or NaN) cannot be represented as hi + lo fp16 (the
  // conversion saturates); the epilogue ORs 1 into this word when it writes such a value (null = unchecked)
  unsigned int* overflow = nullptr;
};
// small_dense_kernels.cu
// conv1_1: NCHW fp32 (N,3,H,W) -> split NHWC (N,H,W,Cout), bias+ReLU fused.  w (Cout,3,3,3) fp32 reference layout.
int k_conv_first(dc_ctx*, const float* img, const float* w, const float* bias, __half* y_hi, __half* y_lo,
                 int N, int H, int W, int Cout, cudaStream_t st);
// objectness (1 row) + box regression (4 rows) heads in one pass over the codes: obj (M), trans (M,4)
int k_recog_heads(dc_ctx*, const float* x, int ldx, const float* w_obj, const float* b_obj, const float* w_box, const float* b_box,
                  float* obj, float* trans, int M, int K, cudaStream_t st);

// tc_conv_first.cu : conv1_1 (K = 27) on tcgen05; w_planes = split [64][64] bank from k_conv_first_weight_to_split
int k_tc_conv_first(dc_ctx*, const float* img, const __half* w_planes, size_t w_plane_elems, const float* bias, float acc_scale, __half* y_hi, __half* y_lo,
                    int N, int H, int W, cudaStream_t st);
int k_conv_first_weight_to_split(dc_ctx*, const float* w, __half* hi, __half* lo, float scale, cudaStream_t st);

// tc_gemm.cu : tcgen05 engine
int tc_init(dc_ctx*);   // resolves cuTensorMapEncodeTiled, sets smem attributes
int k_tc_gemm(dc_ctx*, const __half* a_planes, size_t a_plane_elems, int lda, const __half* b_planes, size_t b_plane_elems, int ldb,
              int M, int N, int K, GemmEpilogue ep, cudaStream_t st);
// x: split NHWC (N,H,W,Cin) planes; w: split [Cout][9*Cin]; y: split NHWC (N,H,W,Cout) or pooled (N,ceil(H/2),ceil(W/2),Cout)
int k_tc_conv3x3(dc_ctx*, const __half* x_planes, size_t x_plane_elems, const __half* w_planes, size_t w_plane_elems,
                 const float* bias, float acc_scale, __half* y_hi, __half* y_lo, int N, int H, int W, int Cin, int Cout, int fuse_pool,
                 cudaStream_t st, int relu = 1);

// lm_decode.cu : persistent caption decoder (LanguageModel:sample over the rows to decode of all images of a batch)
struct LmDecodeArgs {
  int B, P, E, H, V, T;                     // images, row stride per image, encoding / rnn size, vocabulary, steps
  const int32_t* counts;                    // (B) device: rows to decode per image
  const int32_t* idx32;                     // (B*P) device: per-image source row of each row to decode, or null = identity
  int survivors;                            // 1: rows are final-NMS survivors, tokens go to fin_seq rows b*P+i (and all_seq[src])
  const __half* enc_planes; size_t enc_plane_elems;       // (B*P, E) split: encoded image vectors of every proposal
  __half* x_planes; size_t x_plane_elems;   // (B*P, E) split scratch: the compacted encoded vectors (input of the first LSTM step)
  __half* h_planes;                         // [hi: buf0, buf1][lo: buf0, buf1], each (B*P, H): h_t ping-pong
  float* cstate;                            // (B*P, H)
  float* gate_ws;                           // (B*P, 4H) fp32 scratch: h . Wh^T of the current step (weights-stationary decoder); null = tile decoder
  const float* gate_table;                  // (V+2, 4H) fp32: b + lookup . Wx, gate-interleaved columns
  int32_t* tokens;                          // (B*P) scratch: current input token of each row
  const __half* lstm_planes; size_t lstm_plane_elems; const float* lstm_bias; float lstm_acc_scale;
  const __half* vocab_planes; size_t vocab_plane_elems; const float* vocab_bias; float vocab_acc_scale;
  unsigned long long* packed;               // (B*P)
  uint32_t* counters;                       // lm_decode_counter_words(T, B*P) words
  int32_t* m_rows;                          // (1) device out: total rows decoded
  int32_t* dst_final; int32_t* dst_all;     // (B*P) scratch
  int64_t* fin_seq; int64_t* all_seq;       // (B*P, T) token outputs (all_seq may be null)
  float* logits;                            // optional (rows, T, V+1)
};
size_t lm_decode_counter_words(int T, int rows_cap);
// compact_only: build the row maps (m_rows, dst_final, dst_all, compacted x) without launching the greedy decoder
int k_lm_decode(dc_ctx*, const LmDecodeArgs& a, cudaStream_t st, bool compact_only = false);

// beam_search.cu : selection / bookkeeping kernels of LanguageModel:beamsearch (rows = roi * beam + b)
int k_bs_first(dc_ctx*, const float* logits, int R, int V1, int Bm, int T, int H, float* beam_lp, int32_t* beams, int32_t* tokens, const __half* h_hi,
               const __half* h_lo, const float* c, __half* hd_hi, __half* hd_lo, float* cd, cudaStream_t st);
int k_bs_row_topk(dc_ctx*, const float* logits, int rows, int V1, int Bm, int T, int end_token, const int32_t* beams, float* cand_lp, int32_t* cand_ix,
                  cudaStream_t st);
int k_bs_merge(dc_ctx*, int R, int Bm, int T, int t, const float* cand_lp, const int32_t* cand_ix, float* beam_lp, const int32_t* beams_in,
               int32_t* beams_out, int32_t* src_row, int32_t* tokens, cudaStream_t st);
int k_bs_reindex(dc_ctx*, int rows, int H, const int32_t* src_row, const __half* hi_in, const __half* lo_in, const float* c_in, __half* hi_out,
                 __half* lo_out, float* c_out, cudaStream_t st);
int k_bs_final(dc_ctx*, int R, int Bm, int T, const float* beam_lp, const int32_t* beams, const int32_t* dst_row, int64_t* seq, cudaStream_t st);
int k_bs_fill_i32(dc_ctx*, int32_t* p, int32_t v, int n, cudaStream_t st);
int k_bs_gather_split(dc_ctx*, const __half* hi_in, const __half* lo_in, int E, const int32_t* rows_src, int first, int R, __half* hi_out,
                      __half* lo_out, cudaStream_t st);

// lm_kernels.cu
int k_final_gather(dc_ctx*, const int32_t* idx32, const int32_t* count, int P, const float* boxes, const float* scores,
                   const int64_t* seq, int T, float* out_boxes, float* out_scores, int64_t* out_seq, cudaStream_t st);
int k_gather_rows_f32(dc_ctx*, const float* src, int cols, const int32_t* idx32, const int32_t* count, int max_rows, float* dst,
                      cudaStream_t st);
int k_fill_u32(dc_ctx*, uint32_t* p, uint32_t v, size_t n, cudaStream_t st);

// roi_pool.cu -- backward of BilinearRoiPooling (training path): grad_out BHWC (K, HH*WW, C) fp32 ->
// grad_feat HWC fp32 (zeroed here, atomically accumulated), grad_boxes (K,4); grad_grid_ws: K*HH*WW float2 scratch
int k_roi_pool_backward(dc_ctx*, const __half* feat_hi, const __half* feat_lo, int C, int fh, int fw, const float* boxes, int nboxes,
                        int img_h, int img_w, int HH, int WW, const float* grad_out_bhwc, float* grad_feat_hwc, float* grad_grid_ws,
                        float* grad_boxes, cudaStream_t st);
// train_kernels.cu -- weight re-layouts for the backward-data passes
int k_conv_weight_flip_transpose(dc_ctx*, const float* w, float* out, int F, int C, cudaStream_t st);
int k_transpose_into(dc_ctx*, const float* w, int R, int F, float* out, int ld, int col0, cudaStream_t st);
int k_preprocess_u8(dc_ctx*, const uint8_t* in_hwc_rgb, float* out_nchw_bgr, int N, int H, int W, cudaStream_t st);
