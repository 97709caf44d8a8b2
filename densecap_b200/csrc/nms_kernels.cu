// Greedy NMS (box_utils.nms, densecap/box_utils.lua:154-256) as rank sort + bitmask + scan.
//
// Reference: sort scores ascending on the CPU, repeatedly pop the best remaining box i, drop every
// remaining j with iou(i,j) > thresh, stop after max_boxes picks.  Equivalent GPU formulation
// (SURVEY.md App. B5): visit boxes in descending score order; keep p iff no previously KEPT box
// suppresses it.  The IoU expression reproduces the reference's fp32 operation order exactly (no FMA):
//   area = ((x2-x1)+1)*((y2-y1)+1); w = max((min(x2)-max(x1))+1, 0); inter = w*h;
//   union = (area + area_i) - inter; iou = inter/union; keep iff iou <= thresh (fp32 compare).
// Tie rule (the reference's quicksort is unstable; the oracle defines it): equal scores are visited
// higher-index first, i.e. descending order of the 64-bit key (sortable(score)<<32 | idx).
//
// Work-efficient schedule: the sorted candidates are processed in super-chunks of M = 1024.  Per
// super-chunk the kernel builds (a) the upper-triangular 64-bit suppression mask of the chunk against
// itself and (b) the mask of the boxes kept so far against the chunk, then scans the chunk entirely out
// of shared memory (mask 128 KB) and appends the survivors.  Only rows of KEPT boxes are ever needed, so
// the n^2/2 mask of the textbook formulation (210 M IoUs at n = 20520) shrinks to M^2/2 + kept*M per
// super-chunk, and the loop stops at the super-chunk that reaches max_boxes (with the reference's default
// initialisation 1000 picks are reached within the first ~1300 candidates).  One launch, no host sync.
#include "common.cuh"

constexpr int NMS_M = 1024;              // super-chunk size
constexpr int NMS_W = NMS_M / 64;        // 16 mask words per row
constexpr int NMS_THREADS = 256;

struct NmsState {
  int32_t count;      // boxes kept so far
  int32_t done;       // no further super-chunk needs to run
  uint32_t ticket;    // last-block-done counter
  int32_t pad;
};

size_t nms_workspace_bytes(int cap) {
  size_t b = 0;
  b += (size_t)cap * sizeof(uint64_t);              // keys
  b += (size_t)cap * 4 * sizeof(float);             // sorted boxes
  b += (size_t)cap * 4 * sizeof(float);             // kept boxes
  b += (size_t)NMS_M * NMS_W * sizeof(uint64_t);    // intra mask
  b += (size_t)cap * NMS_W * sizeof(uint64_t);      // prev mask (row = kept box)
  b += (size_t)cap * sizeof(int32_t) * 2;           // order, rank
  b += sizeof(NmsState);
  return b + 1024;
}
void nms_workspace_carve(NmsWorkspace& ws, void* base, int cap) {
  char* p = (char*)base;
  ws.cap = cap;
  ws.keys = (uint64_t*)p;          p += (size_t)cap * sizeof(uint64_t);
  ws.sorted_boxes = (float*)p;     p += (size_t)cap * 4 * sizeof(float);
  ws.kept_boxes = (float*)p;       p += (size_t)cap * 4 * sizeof(float);
  ws.mask = (uint64_t*)p;          p += (size_t)NMS_M * NMS_W * sizeof(uint64_t);
  ws.mask_prev = (uint64_t*)p;     p += (size_t)cap * NMS_W * sizeof(uint64_t);
  ws.order = (int32_t*)p;          p += (size_t)cap * sizeof(int32_t);
  ws.rank = (int32_t*)p;           p += (size_t)cap * sizeof(int32_t);
  ws.state = (void*)p;
}

__device__ __forceinline__ uint32_t sortable_f32(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// keys + reset of rank[] and of the scan state
__global__ void nms_keys_kernel(const float* __restrict__ scores, int stride, int n_max, const int32_t* __restrict__ n_dev,
                                uint64_t* __restrict__ keys, int32_t* __restrict__ rank, NmsState* __restrict__ state,
                                int32_t* __restrict__ count_out) {
  const int n = n_dev ? min(*n_dev, n_max) : n_max;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    keys[i] = ((uint64_t)sortable_f32(scores[(size_t)i * stride]) << 32) | (uint32_t)i;
    rank[i] = 0;
  }
  if (i == 0) {
    state->count = 0;
    state->done = (n <= 0) ? 1 : 0;
    state->ticket = 0;
    *count_out = 0;
  }
}

// Rank sort, 2-D decomposition: block (x = 512 keys i, y = slice of 2048 keys j) adds its partial count
// #{j in slice : key_j > key_i} to rank[i], where key = (sortable(score), index).  Because the index is the
// tie-break, a slice lying entirely ABOVE the i block satisfies key_j > key_i <=> score_j >= score_i and a slice
// entirely BELOW it <=> score_j > score_i: one 32-bit compare per pair; only the slice that overlaps the i block
// needs the explicit tie term.
constexpr int RANK_SLICE = 2048;
__global__ void __launch_bounds__(256)
nms_rank_kernel(const uint64_t* __restrict__ keys, int n_max, const int32_t* __restrict__ n_dev, int32_t* __restrict__ rank) {
  __shared__ __align__(16) uint32_t tile[RANK_SLICE];
  const int n = n_dev ? min(*n_dev, n_max) : n_max;
  const int i0 = blockIdx.x * 512;
  const int j0 = blockIdx.y * RANK_SLICE;
  if (i0 >= n || j0 >= n) return;
  const int jn = min(RANK_SLICE, n - j0);
  for (int t = threadIdx.x; t < RANK_SLICE; t += blockDim.x) tile[t] = (t < jn) ? (uint32_t)(keys[j0 + t] >> 32) : 0u;   // 0 < every real key
  __syncthreads();
  const int ia = i0 + threadIdx.x, ib = ia + 256;        // two keys per thread
  const uint32_t sa = (ia < n) ? (uint32_t)(keys[ia] >> 32) : 0xFFFFFFFFu;
  const uint32_t sb = (ib < n) ? (uint32_t)(keys[ib] >> 32) : 0xFFFFFFFFu;
  int ra = 0, rb = 0;
  // j < i0: strict compare; j in [i0, i0+512): explicit tie-break; j >= i0+512: ties count.
  const int t_lo = min(max(i0 - j0, 0), jn);              // first t of the overlapping segment
  const int t_hi = min(max(i0 + 512 - j0, 0), jn);        // first t after it
  {
    int t = 0;
    for (; t + 4 <= (t_lo & ~3); t += 4) {                  // vectorised part of the strict segment
      const uint4 k = *reinterpret_cast<const uint4*>(tile + t);
      ra += (k.x > sa) + (k.y > sa) + (k.z > sa) + (k.w > sa);
      rb += (k.x > sb) + (k.y > sb) + (k.z > sb) + (k.w > sb);
    }
    for (; t < t_lo; ++t) { const uint32_t k = tile[t]; ra += (k > sa); rb += (k > sb); }
    for (; t < t_hi; ++t) {                                 // overlap: index tie-break
      const uint32_t k = tile[t];
      const int j = j0 + t;
      ra += (k > sa) | ((k == sa) & (j > ia));
      rb += (k > sb) | ((k == sb) & (j > ib));
    }
    for (; t < jn && (t & 3); ++t) { const uint32_t k = tile[t]; ra += (k >= sa); rb += (k >= sb); }
    for (; t + 4 <= jn; t += 4) {                           // vectorised part of the >= segment
      const uint4 k = *reinterpret_cast<const uint4*>(tile + t);
      ra += (k.x >= sa) + (k.y >= sa) + (k.z >= sa) + (k.w >= sa);
      rb += (k.x >= sb) + (k.y >= sb) + (k.z >= sb) + (k.w >= sb);
    }
    for (; t < jn; ++t) { const uint32_t k = tile[t]; ra += (k >= sa); rb += (k >= sb); }
  }
  if (ia < n && ra) atomicAdd(rank + ia, ra);
  if (ib < n && rb) atomicAdd(rank + ib, rb);
}

__global__ void nms_scatter_kernel(const int32_t* __restrict__ rank, const float* __restrict__ boxes, int box_stride, int n_max,
                                   const int32_t* __restrict__ n_dev, int32_t* __restrict__ order, float4* __restrict__ sorted_boxes) {
  const int n = n_dev ? min(*n_dev, n_max) : n_max;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = rank[i];
  order[r] = i;
  const float* b = boxes + (size_t)i * box_stride;
  sorted_boxes[r] = make_float4(b[0], b[1], b[2], b[3]);
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

// ---------------------------------------------------------------------------------------------------------
// Super-chunk machinery.  Phase A = 64x64 suppression tiles (upper-triangular tiles of the chunk against itself +
// tiles of the boxes kept so far against the chunk), phase B = scan of the chunk out of shared memory.
// Launch schedule (k_nms): the first NMS_WIDE_CHUNKS super-chunks run phase A as a WIDE grid (one 64-thread block
// per tile, all SMs) followed by a one-block scan kernel; both exit immediately once max_boxes survivors exist, so in
// the common case (1000 picks within the first ~1300 candidates) the whole NMS is a handful of short launches.
// Candidates beyond those super-chunks (rare: NMS that really has to walk the whole list) are handled by ONE launch
// of a single thread-block cluster that loops phase A / phase B with barrier.cluster in between.
// ---------------------------------------------------------------------------------------------------------
constexpr int NMS_CTAS = 8;
constexpr int NMS_WIDE_CHUNKS = 2;

__device__ __forceinline__ void nms_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t nms_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// One 64x64 tile by 64 threads (lrow = row inside the tile).
__device__ __forceinline__ void nms_mask_tile(int tile, int lrow, const float4* __restrict__ sb, int base, int cn_total, int cwords,
                                              int kept0, float thresh, uint64_t* __restrict__ mask_intra,
                                              uint64_t* __restrict__ mask_prev, const float4* __restrict__ kept_boxes) {
  const int intra_tiles = cwords * (cwords + 1) / 2;       // (rb, cb) with cb >= rb, row-major
  int rb, cb;
  const bool intra = tile < intra_tiles;
  if (intra) {
    int t = tile;
    rb = 0;
    while (t >= cwords - rb) { t -= cwords - rb; ++rb; }
    cb = rb + t;
  } else {
    const int t = tile - intra_tiles;
    rb = t / cwords;
    cb = t % cwords;
  }
  const int col0 = cb * 64;
  const int cn = min(64, cn_total - col0);
  uint64_t bits = 0;
  if (intra) {
    const int p_local = rb * 64 + lrow;
    if (p_local < cn_total) {
      const float4 pb = sb[base + p_local];
      const float pa = nms_area(pb);
      // off-diagonal tiles: every q of the tile is later than p.  Diagonal tile: the FULL row (all q != p) -- the IoU test is
      // symmetric bit for bit, and the scan resolves a 64-box block through the bits of the EARLIER boxes of each row.
      const int self = p_local - col0;
      for (int t = 0; t < cn; ++t) {
        const float4 q = __ldg(sb + base + col0 + t);    // same address for the 64 threads of the tile: broadcast
        if (t != self && nms_suppresses(pb, pa, q, nms_area(q), thresh)) bits |= (1ull << t);
      }
    }
    mask_intra[(size_t)p_local * NMS_W + cb] = bits;
  } else {
    const int r = rb * 64 + lrow;
    if (r < kept0) {
      // kept_boxes is appended to between launches / cluster barriers: read it at L2 (another SM's L1 may hold a line
      // that was only partly written when it was cached for an earlier super-chunk)
      const float4 pb = __ldcg(kept_boxes + r);
      const float pa = nms_area(pb);
      for (int t = 0; t < cn; ++t) {
        const float4 q = __ldg(sb + base + col0 + t);
        if (nms_suppresses(pb, pa, q, nms_area(q), thresh)) bits |= (1ull << t);
      }
      mask_prev[(size_t)r * NMS_W + cb] = bits;
    }
  }
}

// Scan of one super-chunk by one CTA of NMS_THREADS threads.  smem_mask: [NMS_M][NMS_W] words.
__device__ __forceinline__ void nms_scan_chunk(uint64_t* smem_mask, uint64_t* s_removed, int* s_pos, int* s_cnt, const float4* __restrict__ sb,
                                               const int32_t* __restrict__ order, int base, int cn_total, int cwords, int kept0, int cap,
                                               const uint64_t* __restrict__ mask_intra, const uint64_t* __restrict__ mask_prev,
                                               float4* __restrict__ kept_boxes, NmsState* __restrict__ state, int64_t* __restrict__ idx_out,
                                               int32_t* __restrict__ idx32_out, int32_t* __restrict__ count_out) {
  const int tid = threadIdx.x;
  {
    // chunk mask -> shared memory, 8 independent loads in flight per thread (a plain copy loop serialises on L2 latency)
    const int total_words = cn_total * NMS_W;
    for (int i0 = tid; i0 < total_words; i0 += NMS_THREADS * 8) {
      uint64_t v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * NMS_THREADS;
        v[u] = (i < total_words) ? __ldcg(mask_intra + i) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * NMS_THREADS;
        if (i < total_words) smem_mask[i] = v[u];
      }
    }
  }
  if (tid < NMS_W) s_removed[tid] = 0;
  __syncthreads();
  {
    // suppression bits inherited from the boxes kept by earlier super-chunks
    const int w = tid & (NMS_W - 1);
    uint64_t acc = 0;
    constexpr int RSTEP = NMS_THREADS / NMS_W;
    if (w < cwords) {
      for (int r0 = tid >> 4; r0 < kept0; r0 += RSTEP * 8) {
        uint64_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = r0 + u * RSTEP;
          v[u] = (r < kept0) ? __ldcg(mask_prev + (size_t)r * NMS_W + w) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc |= v[u];
      }
    }
    if (acc) atomicOr((unsigned long long*)&s_removed[w], (unsigned long long)acc);
  }
  __syncthreads();
  if (tid < 32) {
    const int lane = tid;
    int count = kept0;
    uint64_t my_removed = (lane < NMS_W) ? s_removed[lane] : 0ull;    // lane w owns removed word w
    for (int c = 0; c < cwords && count < cap; ++c) {
      const int lim = min(64, cn_total - c * 64);
      const uint64_t d0 = (lane < lim) ? smem_mask[(size_t)(c * 64 + lane) * NMS_W + c] : 0ull;
      const uint64_t d1 = (lane + 32 < lim) ? smem_mask[(size_t)(c * 64 + 32 + lane) * NMS_W + c] : 0ull;
      const uint64_t cur = __shfl_sync(0xffffffffu, my_removed, c);
      const int before = count;
      // Greedy NMS inside the 64-box block as a fixpoint instead of a serial walk: box i survives iff it is alive and no
      // SURVIVING earlier box of the block suppresses it, K_i = alive_i & !(row_i & below_i & K).  The solution is unique
      // (induction on i) and iterating from K = alive fixes at least one more leading box per round; real chains are a few
      // boxes deep, so a block costs a handful of ballot rounds instead of one shuffle-latency per survivor.
      const uint64_t alive0 = ~cur & ((lim == 64) ? ~0ull : ((1ull << lim) - 1ull));
      const bool a0 = (alive0 >> lane) & 1ull, a1 = (alive0 >> (lane + 32)) & 1ull;
      const uint64_t m0 = d0 & ((1ull << lane) - 1ull), m1 = d1 & ((1ull << (lane + 32)) - 1ull);
      uint64_t keep = alive0;
      for (;;) {
        const bool k0 = a0 && (m0 & keep) == 0ull, k1 = a1 && (m1 & keep) == 0ull;
        const uint64_t next = (uint64_t)__ballot_sync(0xffffffffu, k0) | ((uint64_t)__ballot_sync(0xffffffffu, k1) << 32);
        if (next == keep) break;
        keep = next;
      }
      int nk = __popcll(keep);
      if (count + nk > cap) {                      // max_boxes reached inside this block: keep the first (cap - count) survivors
        uint64_t rest = keep;
        for (int z = cap - count; z > 0; --z) rest &= rest - 1;
        keep &= ~rest;
        nk = cap - count;
      }
      count += nk;
      // remember the sorted positions of the survivors (index translation + box copy happen after the loop, in parallel)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int i = lane + 32 * half;
        if ((keep >> i) & 1ull) s_pos[before - kept0 + __popcll(keep & ((1ull << i) - 1ull))] = base + c * 64 + i;
      }
      // propagate the suppression rows of the survivors to the later words of this super-chunk
      if (lane > c && lane < cwords) {
        uint64_t k = keep, acc = 0;
        while (k) {
          const int i = __ffsll((long long)k) - 1;
          k &= k - 1;
          acc |= smem_mask[(size_t)(c * 64 + i) * NMS_W + lane];
        }
        my_removed |= acc;
      }
    }
    if (lane == 0) *s_cnt = count;
  }
  __syncthreads();
  const int count = *s_cnt;
  // emit survivors in scan order (= decreasing score) and keep their boxes for later super-chunks
  for (int e = tid; e < count - kept0; e += NMS_THREADS) {
    const int sp = s_pos[e];
    const int bi = order[sp];
    const int pos = kept0 + e;
    if (idx_out) idx_out[pos] = (int64_t)bi + 1;
    if (idx32_out) idx32_out[pos] = bi;
    kept_boxes[pos] = sb[sp];
  }
  if (tid == 0) {
    state->count = count;
    *count_out = count;
  }
  __threadfence();
}

// wide phase A of super-chunk s: one 64-thread block per tile
__global__ void __launch_bounds__(64)
nms_mask_kernel(int s, const float4* __restrict__ sb, int n_max, const int32_t* __restrict__ n_dev, float thresh, int max_boxes,
                uint64_t* __restrict__ mask_intra, uint64_t* __restrict__ mask_prev, const float4* __restrict__ kept_boxes,
                const NmsState* __restrict__ state) {
  const int n = n_dev ? min(*n_dev, n_max) : n_max;
  const int cap = (max_boxes > 0) ? max_boxes : n;
  const int base = s * NMS_M;
  const int kept0 = state->count;
  if (base >= n || kept0 >= cap) return;
  const int cn_total = min(NMS_M, n - base);
  const int cwords = DC_CEIL_DIV(cn_total, 64);
  const int total_tiles = cwords * (cwords + 1) / 2 + DC_CEIL_DIV(kept0, 64) * cwords;
  if ((int)blockIdx.x >= total_tiles) return;
  nms_mask_tile(blockIdx.x, threadIdx.x, sb, base, cn_total, cwords, kept0, thresh, mask_intra, mask_prev, kept_boxes);
}

// phase B of super-chunk s by one block
__global__ void __launch_bounds__(NMS_THREADS, 1)
nms_scan_kernel(int s, const float4* __restrict__ sb, const int32_t* __restrict__ order, int n_max, const int32_t* __restrict__ n_dev,
                int max_boxes, const uint64_t* __restrict__ mask_intra, const uint64_t* __restrict__ mask_prev, float4* __restrict__ kept_boxes,
                NmsState* __restrict__ state, int64_t* __restrict__ idx_out, int32_t* __restrict__ idx32_out, int32_t* __restrict__ count_out) {
  extern __shared__ __align__(16) uint64_t smem_mask[];     // [NMS_M][NMS_W] then s_pos[NMS_M]
  __shared__ uint64_t s_removed[NMS_W];
  __shared__ int s_cnt;
  int* s_pos = reinterpret_cast<int*>(smem_mask + (size_t)NMS_M * NMS_W);
  const int n = n_dev ? min(*n_dev, n_max) : n_max;
  const int cap = (max_boxes > 0) ? max_boxes : n;
  const int base = s * NMS_M;
  const int kept0 = state->count;
  if (base >= n || kept0 >= cap) return;
  const int cn_total = min(NMS_M, n - base);
  nms_scan_chunk(smem_mask, s_removed, s_pos, &s_cnt, sb, order, base, cn_total, DC_CEIL_DIV(cn_total, 64), kept0, cap, mask_intra, mask_prev,
                 kept_boxes, state, idx_out, idx32_out, count_out);
}

// tail: all remaining super-chunks in one launch of a single cluster
__global__ void __launch_bounds__(NMS_THREADS, 1)
nms_cluster_kernel(int first_chunk, const float4* __restrict__ sb, const int32_t* __restrict__ order, int n_max,
                   const int32_t* __restrict__ n_dev, float thresh, int max_boxes, uint64_t* __restrict__ mask_intra,
                   uint64_t* __restrict__ mask_prev, float4* __restrict__ kept_boxes, NmsState* __restrict__ state,
                   int64_t* __restrict__ idx_out, int32_t* __restrict__ idx32_out, int32_t* __restrict__ count_out) {
  extern __shared__ __align__(16) uint64_t smem_mask[];
  __shared__ uint64_t s_removed[NMS_W];
  __shared__ int s_cnt;
  int* s_pos = reinterpret_cast<int*>(smem_mask + (size_t)NMS_M * NMS_W);
  const int n = n_dev ? min(*n_dev, n_max) : n_max;
  const int cap = (max_boxes > 0) ? max_boxes : n;
  const int tid = threadIdx.x;
  const uint32_t rank_in_cluster = nms_cluster_ctarank();
  const int num_workers = NMS_CTAS * (NMS_THREADS / 64);
  const int worker = (int)rank_in_cluster * (NMS_THREADS / 64) + (tid >> 6);
  for (int base = first_chunk * NMS_M; base < n; base += NMS_M) {
    const int kept0 = __ldcg(&state->count);                 // written by CTA 0 before the previous barrier / launch (read at L2)
    if (kept0 >= cap) break;                                 // identical in every CTA of the cluster
    const int cn_total = min(NMS_M, n - base);
    const int cwords = DC_CEIL_DIV(cn_total, 64);
    const int total_tiles = cwords * (cwords + 1) / 2 + DC_CEIL_DIV(kept0, 64) * cwords;
    for (int tile = worker; tile < total_tiles; tile += num_workers)
      nms_mask_tile(tile, tid & 63, sb, base, cn_total, cwords, kept0, thresh, mask_intra, mask_prev, kept_boxes);
    __threadfence();
    nms_cluster_sync();
    if (rank_in_cluster == 0)
      nms_scan_chunk(smem_mask, s_removed, s_pos, &s_cnt, sb, order, base, cn_total, cwords, kept0, cap, mask_intra, mask_prev, kept_boxes, state,
                     idx_out, idx32_out, count_out);
    nms_cluster_sync();
  }
}

int k_nms(dc_ctx* ctx, const float* boxes, int box_stride, const float* scores, int score_stride, int n_max, const int32_t* n_dev,
          float thresh, int max_boxes, NmsWorkspace ws, int64_t* idx_out, int32_t* idx32_out, int32_t* count_out, cudaStream_t st) {
  NmsState* state = (NmsState*)ws.state;
  if (n_max <= 0) {
    nms_keys_kernel<<<1, 32, 0, st>>>(scores, score_stride, 0, nullptr, ws.keys, ws.rank, state, count_out);
    DC_CHECK_LAUNCH(ctx);
    return DC_OK;
  }
  DC_REQUIRE(ctx, n_max <= ws.cap, DC_E_INVALID, "nms: %d boxes exceed workspace capacity %d", n_max, ws.cap);
  const size_t smem = (size_t)NMS_M * NMS_W * sizeof(uint64_t) + (size_t)NMS_M * sizeof(int);
  {
    cudaError_t e = dc_ensure_smem(ctx, nms_cluster_kernel, smem);
    if (e == cudaSuccess) e = dc_ensure_smem(ctx, nms_scan_kernel, smem);
    if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "nms smem attribute: %s", cudaGetErrorString(e));
  }
  nms_keys_kernel<<<DC_CEIL_DIV(n_max, 256), 256, 0, st>>>(scores, score_stride, n_max, n_dev, ws.keys, ws.rank, state, count_out);
  DC_CHECK_LAUNCH(ctx);
  dim3 rgrid(DC_CEIL_DIV(n_max, 512), DC_CEIL_DIV(n_max, RANK_SLICE));
  nms_rank_kernel<<<rgrid, 256, 0, st>>>(ws.keys, n_max, n_dev, ws.rank);
  DC_CHECK_LAUNCH(ctx);
  nms_scatter_kernel<<<DC_CEIL_DIV(n_max, 256), 256, 0, st>>>(ws.rank, boxes, box_stride, n_max, n_dev, ws.order, (float4*)ws.sorted_boxes);
  DC_CHECK_LAUNCH(ctx);
  const int chunks = DC_CEIL_DIV(n_max, NMS_M);
  const int max_keep = (max_boxes > 0) ? (max_boxes < n_max ? max_boxes : n_max) : n_max;
  const int wide = chunks < NMS_WIDE_CHUNKS ? chunks : NMS_WIDE_CHUNKS;
  for (int s = 0; s < wide; ++s) {
    const int prev_rows = (s == 0) ? 0 : (max_keep < s * NMS_M ? max_keep : s * NMS_M);
    const int tiles = NMS_W * (NMS_W + 1) / 2 + DC_CEIL_DIV(prev_rows, 64) * NMS_W;
    nms_mask_kernel<<<tiles, 64, 0, st>>>(s, (const float4*)ws.sorted_boxes, n_max, n_dev, thresh, max_boxes, ws.mask, ws.mask_prev,
                                          (const float4*)ws.kept_boxes, state);
    DC_CHECK_LAUNCH(ctx);
    nms_scan_kernel<<<1, NMS_THREADS, smem, st>>>(s, (const float4*)ws.sorted_boxes, ws.order, n_max, n_dev, max_boxes, ws.mask, ws.mask_prev,
                                                  (float4*)ws.kept_boxes, state, idx_out, idx32_out, count_out);
    DC_CHECK_LAUNCH(ctx);
  }
  if (chunks > wide) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(NMS_CTAS);
    cfg.blockDim = dim3(NMS_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NMS_CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, nms_cluster_kernel, wide, (const float4*)ws.sorted_boxes, (const int32_t*)ws.order, n_max, n_dev,
                                       thresh, max_boxes, ws.mask, ws.mask_prev, (float4*)ws.kept_boxes, state, idx_out, idx32_out, count_out);
    if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "nms cluster launch: %s", cudaGetErrorString(e));
    DC_CHECK_LAUNCH(ctx);
  }
  return DC_OK;
}
