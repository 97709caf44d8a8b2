// Greedy NMS (box_utils.nms, densecap/box_utils.lua:154-256) as sort + bitmask + scan.
//
// Reference: sort scores ascending on the CPU, repeatedly pop the best remaining box i, drop every
// remaining j with iou(i,j) > thresh, stop after max_boxes picks.  Equivalent GPU formulation
// (SURVEY.md App. B5): sort descending; mask[p][q] = !(iou(p,q) <= thresh) for q > p in sorted
// order as 64-bit words; sequential scan "keep p iff not removed; removed |= mask[p]".
// The IoU expression reproduces the reference's fp32 operation order exactly (no FMA):
//   area = ((x2-x1)+1)*((y2-y1)+1); w = max((min(x2)-max(x1))+1, 0); inter = w*h;
//   union = (area + area_i) - inter; iou = inter/union; keep iff iou <= thresh (fp32 compare).
// Tie rule (the reference's quicksort is unstable; the oracle defines it): equal scores are
// visited higher-index first, i.e. descending order of the 64-bit key (sortable(score)<<32 | idx).
#include "common.cuh"

size_t nms_workspace_bytes(int cap) {
  const size_t words = DC_CEIL_DIV(cap, 64);
  size_t b = 0;
  b += (size_t)cap * sizeof(uint64_t);        // keys
  b += (size_t)cap * sizeof(int32_t);         // order
  b += (size_t)cap * 4 * sizeof(float);       // sorted boxes
  b += (size_t)cap * words * sizeof(uint64_t);  // mask
  return b + 1024;
}
void nms_workspace_carve(NmsWorkspace& ws, void* base, int cap) {
  char* p = (char*)base;
  const size_t words = DC_CEIL_DIV(cap, 64);
  ws.cap = cap;
  ws.keys = (uint64_t*)p;          p += (size_t)cap * sizeof(uint64_t);
  ws.sorted_boxes = (float*)p;     p += (size_t)cap * 4 * sizeof(float);
  ws.mask = (uint64_t*)p;          p += (size_t)cap * words * sizeof(uint64_t);
  ws.order = (int32_t*)p;
}

__device__ __forceinline__ uint32_t sortable_f32(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void nms_keys_kernel(const float* __restrict__ scores, int stride, int n_max, const int32_t* __restrict__ n_dev,
                                uint64_t* __restrict__ keys) {
  const int n = n_dev ? min(*n_dev, n_max) : n_max;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = ((uint64_t)sortable_f32(scores[(size_t)i * stride]) << 32) | (uint32_t)i;
}

// Rank sort: position of box i in descending key order = #{j : key_j > key_i}.  Keys are unique
// (index in the low bits), so ranks are a permutation.  O(n^2) compares, fully parallel; n ~ 2e4.
constexpr int RANK_TILE = 2048;
__global__ void nms_rank_kernel(const uint64_t* __restrict__ keys, const float* __restrict__ boxes, int box_stride, int n_max,
                                const int32_t* __restrict__ n_dev, int32_t* __restrict__ order, float4* __restrict__ sorted_boxes) {
  __shared__ uint64_t tile[RANK_TILE];
  const int n = n_dev ? min(*n_dev, n_max) : n_max;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x * blockDim.x >= n) return;
  const uint64_t mine = (i < n) ? keys[i] : 0;
  int rank = 0;
  for (int j0 = 0; j0 < n; j0 += RANK_TILE) {
    const int jn = min(RANK_TILE, n - j0);
    __syncthreads();
    for (int t = threadIdx.x; t < jn; t += blockDim.x) tile[t] = keys[j0 + t];
    __syncthreads();
    int t = 0;
    for (; t + 4 <= jn; t += 4) {
      rank += (tile[t] > mine) + (tile[t + 1] > mine) + (tile[t + 2] > mine) + (tile[t + 3] > mine);
    }
    for (; t < jn; ++t) rank += (tile[t] > mine);
  }
  if (i < n) {
    order[rank] = i;
    const float* b = boxes + (size_t)i * box_stride;
    sorted_boxes[rank] = make_float4(b[0], b[1], b[2], b[3]);
  }
}

__device__ __forceinline__ float nms_area(float4 b) {
  return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
}
// true iff box q is suppressed by kept box p, i.e. NOT (iou <= thresh)   (box_utils.lua:219-227,241)
__device__ __forceinline__ bool nms_suppresses(float4 p, float ap, float4 q, float aq, float thresh) {
  const float xx1 = fmaxf(q.x, p.x), yy1 = fmaxf(q.y, p.y);
  const float xx2 = fminf(q.z, p.z), yy2 = fminf(q.w, p.w);
  const float w = fmaxf(__fadd_rn(__fsub_rn(xx2, xx1), 1.f), 0.f);
  const float h = fmaxf(__fadd_rn(__fsub_rn(yy2, yy1), 1.f), 0.f);
  const float inter = __fmul_rn(w, h);
  const float uni = __fsub_rn(__fadd_rn(aq, ap), inter);
  const float iou = __fdiv_rn(inter, uni);
  return !(iou <= thresh);
}

// mask[p][cb] bit t = box (64*cb + t) suppressed by box p; only cb >= p/64 is produced, and only
// bits with q > p are set.  grid = (col blocks, row blocks), 64 threads.
__global__ void nms_mask_kernel(const float4* __restrict__ sb, int n_max, const int32_t* __restrict__ n_dev, float thresh,
                                uint64_t* __restrict__ mask, int words) {
  const int n = n_dev ? min(*n_dev, n_max) : n_max;
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (cb < rb || rb * 64 >= n || cb * 64 >= n) return;
  __shared__ float4 cbox[64];
  __shared__ float carea[64];
  const int cn = min(64, n - cb * 64);
  if ((int)threadIdx.x < cn) {
    const float4 b = sb[cb * 64 + threadIdx.x];
    cbox[threadIdx.x] = b;
    carea[threadIdx.x] = nms_area(b);
  }
  __syncthreads();
  const int p = rb * 64 + threadIdx.x;
  if (p >= n) return;
  const float4 pb = sb[p];
  const float pa = nms_area(pb);
  uint64_t bits = 0;
  const int t0 = (cb == rb) ? threadIdx.x + 1 : 0;
  for (int t = t0; t < cn; ++t)
    if (nms_suppresses(pb, pa, cbox[t], carea[t], thresh)) bits |= (1ull << t);
  mask[(size_t)p * words + cb] = bits;
}

// Sequential scan over 64-box chunks by one block.  removed[] lives in shared memory.
constexpr int SCAN_THREADS = 512;
__global__ void nms_scan_kernel(const uint64_t* __restrict__ mask, int words_ld, const int32_t* __restrict__ order, int n_max,
                                const int32_t* __restrict__ n_dev, int max_boxes, int64_t* __restrict__ idx_out,
                                int32_t* __restrict__ idx32_out, int32_t* __restrict__ count_out) {
  extern __shared__ uint64_t removed[];
  __shared__ uint64_t s_keep;
  __shared__ int s_count;
  const int n = n_dev ? min(*n_dev, n_max) : n_max;
  const int words = DC_CEIL_DIV(n, 64);
  const int cap = (max_boxes > 0) ? max_boxes : n;
  for (int w = threadIdx.x; w < words; w += blockDim.x) removed[w] = 0;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int c = 0; c < words; ++c) {
    if (threadIdx.x < 32) {
      const int base = c * 64;
      const int p0 = base + lane, p1 = base + 32 + lane;
      const uint64_t d0 = (p0 < n) ? mask[(size_t)p0 * words_ld + c] : 0ull;
      const uint64_t d1 = (p1 < n) ? mask[(size_t)p1 * words_ld + c] : 0ull;
      uint64_t cur = removed[c];
      uint64_t keep = 0;
      int count = s_count;
      const int lim = min(64, n - base);
      for (int i = 0; i < lim; ++i) {
        const uint64_t row = __shfl_sync(0xffffffffu, (i < 32) ? d0 : d1, i & 31);
        if (!((cur >> i) & 1ull) && count < cap) {
          keep |= (1ull << i);
          cur |= row;
          ++count;
        }
      }
      // emit kept indices in scan order (= decreasing score)
      const int before = s_count;
      if ((keep >> lane) & 1ull) {
        const int pos = before + __popcll(keep & ((1ull << lane) - 1ull));
        const int bi = order[p0];
        if (idx_out) idx_out[pos] = (int64_t)bi + 1;
        if (idx32_out) idx32_out[pos] = bi;
      }
      if ((keep >> (lane + 32)) & 1ull) {
        const int pos = before + __popcll(keep & ((1ull << (lane + 32)) - 1ull));
        const int bi = order[p1];
        if (idx_out) idx_out[pos] = (int64_t)bi + 1;
        if (idx32_out) idx32_out[pos] = bi;
      }
      if (lane == 0) {
        s_keep = keep;
        s_count = count;
      }
    }
    __syncthreads();
    const uint64_t keep = s_keep;
    const bool done = (s_count >= cap);
    if (!done && keep) {
      for (int w = c + 1 + threadIdx.x; w < words; w += blockDim.x) {
        uint64_t acc = 0;
        uint64_t k = keep;
        while (k) {
          const int i = __ffsll((long long)k) - 1;
          k &= k - 1;
          acc |= mask[(size_t)(c * 64 + i) * words_ld + w];
        }
        removed[w] |= acc;
      }
    }
    __syncthreads();
    if (done) break;
  }
  if (threadIdx.x == 0) *count_out = s_count;
}

__global__ void nms_empty_kernel(int32_t* count_out) { *count_out = 0; }

int k_nms(dc_ctx* ctx, const float* boxes, int box_stride, const float* scores, int score_stride, int n_max, const int32_t* n_dev,
          float thresh, int max_boxes, NmsWorkspace ws, int64_t* idx_out, int32_t* idx32_out, int32_t* count_out, cudaStream_t st) {
  if (n_max <= 0) {
    nms_empty_kernel<<<1, 1, 0, st>>>(count_out);
    DC_CHECK_LAUNCH(ctx);
    return DC_OK;
  }
  DC_REQUIRE(ctx, n_max <= ws.cap, DC_E_INVALID, "nms: %d boxes exceed workspace capacity %d", n_max, ws.cap);
  const int words = DC_CEIL_DIV(n_max, 64);
  nms_keys_kernel<<<DC_CEIL_DIV(n_max, 256), 256, 0, st>>>(scores, score_stride, n_max, n_dev, ws.keys);
  DC_CHECK_LAUNCH(ctx);
  nms_rank_kernel<<<DC_CEIL_DIV(n_max, 128), 128, 0, st>>>(ws.keys, boxes, box_stride, n_max, n_dev, ws.order, (float4*)ws.sorted_boxes);
  DC_CHECK_LAUNCH(ctx);
  dim3 grid(words, words);
  nms_mask_kernel<<<grid, 64, 0, st>>>((const float4*)ws.sorted_boxes, n_max, n_dev, thresh, ws.mask, words);
  DC_CHECK_LAUNCH(ctx);
  nms_scan_kernel<<<1, SCAN_THREADS, words * sizeof(uint64_t), st>>>(ws.mask, words, ws.order, n_max, n_dev, max_boxes, idx_out,
                                                                      idx32_out, count_out);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
