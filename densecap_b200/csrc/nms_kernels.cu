// Greedy NMS (box_utils.nms, densecap/box_utils.lua:154-256) as partial sort + bitmask + scan.
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
// Front end: the scan stops after max_boxes picks, which lie within the first few thousand candidates, so the boxes are
// never fully sorted.  nms_select_sort produces the NEXT 2048 positions of the descending key order on demand: a 12-bit-
// per-level radix select (histogram in shared memory) finds the 2048-th largest key below the last one emitted, the
// selected keys are compacted and bitonic-sorted in shared memory, and order[] / sorted_boxes[] are extended.  Round 0 is
// its own 1024-thread launch; later rounds (deep NMS only) run inside the tail kernel.
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

constexpr int NMS_SEL = 2048;            // sorted positions produced per select/sort round
struct NmsState {
  int32_t count;      // boxes kept so far
  int32_t done;       // no further super-chunk needs to run
  int32_t sorted_n;   // sorted positions produced so far (order[], sorted_boxes[] valid below it)
  int32_t pad;
  unsigned long long last_key;   // smallest key emitted so far (exclusive upper bound of the next round)
};

size_t nms_workspace_bytes(int cap) {
  size_t b = 0;
  b += (size_t)cap * sizeof(uint64_t);              // keys
  b += (size_t)cap * 4 * sizeof(float);             // sorted boxes
  b += (size_t)cap * 4 * sizeof(float);             // kept boxes
  b += (size_t)NMS_M * NMS_W * sizeof(uint64_t);    // intra mask
  b += (size_t)cap * NMS_W * sizeof(uint64_t);      // prev mask (row = kept box)
  b += (size_t)cap * sizeof(int32_t);               // order
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
  ws.state = (void*)p;
}

__device__ __forceinline__ uint32_t sortable_f32(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// keys + reset of the scan state
__global__ void nms_keys_kernel(const float* __restrict__ scores, int stride, int n_max, const int32_t* __restrict__ n_dev,
                                uint64_t* __restrict__ keys, NmsState* __restrict__ state, int32_t* __restrict__ count_out) {
  const int n = n_dev ? min(*n_dev, n_max) : n_max;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = ((uint64_t)sortable_f32(scores[(size_t)i * stride]) << 32) | (uint32_t)i;
  if (i == 0) {
    state->count = 0;
    state->done = (n <= 0) ? 1 : 0;
    state->sorted_n = 0;
    state->last_key = ~0ull;
    *count_out = 0;
  }
}

// ---------------------------------------------------------------------------------------------------------
// nms_select_sort: extend the descending order by the next m = min(NMS_SEL, n - base) positions, base = state->sorted_n.
// Candidates are the keys strictly below state->last_key (keys are unique 64-bit values, so "position >= base" is
// exactly "key < key at position base-1").  Radix select, most significant digit first, 12 bits per level:
//   hist[d] = #candidates matching the digits fixed so far whose next digit is d;  the boundary digit is the largest
//   d with #(digit > d) < need <= #(digit >= d);  everything above it is selected, the search continues inside it.
// It stops as soon as the boundary bin is needed entirely (typical: 3 levels = the 32 score bits; exact score ties
// continue into the index bits, which realises the higher-index-first tie rule).  Then: compaction of the selected
// keys into shared memory, bitonic sort (descending), gather of the boxes.  One CTA of NT threads; smem: NmsSelSmem.
// ---------------------------------------------------------------------------------------------------------
struct NmsSelSmem {
  uint32_t hist[4096];
  unsigned long long sel[NMS_SEL];
  uint32_t warp_sums[32];
  uint32_t n_sel;
  uint32_t bnd_digit, bnd_above, found;
};

template <int NT>
__device__ void nms_select_sort(NmsSelSmem& sm, const uint64_t* __restrict__ keys, int n, const float* __restrict__ boxes, int box_stride,
                                NmsState* __restrict__ state, int32_t* __restrict__ order, float4* __restrict__ sorted_boxes) {
  const int tid = threadIdx.x;
  const int base = __ldcg(&state->sorted_n);
  const unsigned long long last_key = __ldcg(&state->last_key);
  const int remaining = n - base;
  if (remaining <= 0) return;
  const int m = min(NMS_SEL, remaining);
  // ---- radix select of the m-th largest candidate key ----
  unsigned long long prefix = 0ull, pmask = 0ull;      // digits fixed so far
  uint32_t need = (uint32_t)m;
  bool take_all = (remaining <= NMS_SEL);              // everything left is selected: no search needed
  for (int shift = 52; !take_all; shift -= 12) {
    const int sh = shift < 0 ? 0 : shift;
    const uint32_t dmask = shift < 0 ? 0xFu : 0xFFFu;          // last level: the remaining 4 bits
    for (int i = tid; i < 4096; i += NT) sm.hist[i] = 0u;
    if (tid == 0) sm.found = 0u;
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
      const unsigned long long k = __ldg(keys + i);
      if (k < last_key && (k & pmask) == prefix) atomicAdd(&sm.hist[(uint32_t)(k >> sh) & dmask], 1u);
    }
    __syncthreads();
    // suffix sums over the 4096 bins: thread t owns bins [t*PB, t*PB + PB)
    constexpr int PB = 4096 / NT;
    uint32_t loc[PB];
    uint32_t tsum = 0;
#pragma unroll
    for (int j = 0; j < PB; ++j) { loc[j] = sm.hist[tid * PB + j]; tsum += loc[j]; }
    // inclusive suffix scan of tsum across threads (warp shuffle + warp totals)
    uint32_t suf = tsum;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_down_sync(0xffffffffu, suf, o);
      if (lane + o < 32) suf += v;
    }
    if (lane == 0) sm.warp_sums[warp] = suf;
    __syncthreads();
    uint32_t above_warps = 0;
    for (int w = warp + 1; w < NT / 32; ++w) above_warps += sm.warp_sums[w];
    uint32_t above = above_warps + suf - tsum;             // candidates in bins owned by later threads
    // walk my bins from the top: above = #(digit > current bin)
#pragma unroll
    for (int j = PB - 1; j >= 0; --j) {
      if (above < need && need <= above + loc[j]) { sm.bnd_digit = (uint32_t)(tid * PB + j); sm.bnd_above = above; sm.found = 1u; }
      above += loc[j];
    }
    __syncthreads();
    const uint32_t d = sm.bnd_digit, ab = sm.bnd_above;
    const uint32_t in_bin = sm.hist[d];
    need -= ab;
    prefix |= (unsigned long long)d << sh;
    pmask |= (unsigned long long)dmask << sh;
    __syncthreads();
    if (in_bin == need || shift < 0) break;             // whole boundary bin selected (always true once all 64 bits are fixed)
  }
  // selected = candidates with key >= threshold, threshold = prefix with the undecided low bits cleared
  const unsigned long long thr = take_all ? 0ull : prefix;
  if (tid == 0) sm.n_sel = 0u;
  __syncthreads();
  for (int i = tid; i < n; i += NT) {
    const unsigned long long k = __ldg(keys + i);
    if (k < last_key && k >= thr) {
      const uint32_t pos = atomicAdd(&sm.n_sel, 1u);
      if (pos < (uint32_t)NMS_SEL) sm.sel[pos] = k;
    }
  }
  __syncthreads();
  const int cnt = (int)min(sm.n_sel, (uint32_t)NMS_SEL);      // == m by construction
  for (int i = cnt + tid; i < NMS_SEL; i += NT) sm.sel[i] = 0ull;      // padding sorts to the end (real keys are > 0: sortable(score) != 0)
  __syncthreads();
  // ---- bitonic sort, descending ----
  for (int k2 = 2; k2 <= NMS_SEL; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < NMS_SEL; i += NT) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = sm.sel[i], b = sm.sel[l];
          const bool desc = (i & k2) == 0;
          if ((a < b) == desc) { sm.sel[i] = b; sm.sel[l] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < cnt; i += NT) {
    const int bi = (int)(uint32_t)(sm.sel[i] & 0xFFFFFFFFull);
    order[base + i] = bi;
    const float* b = boxes + (size_t)bi * box_stride;
    sorted_boxes[base + i] = make_float4(b[0], b[1], b[2], b[3]);
  }
  __syncthreads();
  if (tid == 0 && cnt > 0) {
    state->last_key = sm.sel[cnt - 1];
    state->sorted_n = base + cnt;
  }
  __threadfence();
}

constexpr int NMS_SEL_THREADS = 1024;
__global__ void __launch_bounds__(NMS_SEL_THREADS, 1)
nms_select_kernel(const uint64_t* __restrict__ keys, int n_max, const int32_t* __restrict__ n_dev, const float* __restrict__ boxes, int box_stride,
                  NmsState* __restrict__ state, int32_t* __restrict__ order, float4* __restrict__ sorted_boxes) {
  extern __shared__ __align__(16) unsigned char sel_smem_raw[];
  NmsSelSmem& sm = *reinterpret_cast<NmsSelSmem*>(sel_smem_raw);
  const int n = n_dev ? min(*n_dev, n_max) : n_max;
  nms_select_sort<NMS_SEL_THREADS>(sm, keys, n, boxes, box_stride, state, order, sorted_boxes);
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
static_assert(NMS_WIDE_CHUNKS * NMS_M == NMS_SEL, "the wide super-chunks cover exactly the first select/sort round");

__device__ __forceinline__ void nms_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t nms_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// One 64x64 tile by 256 threads: thread (lrow, qt) evaluates row lrow against the 16 columns of quarter qt; the four partial
// masks of a row are combined with shuffles (the 4 threads of a row are adjacent lanes) and written by qt == 0.
constexpr int NMS_TILE_THREADS = 256;
__device__ __forceinline__ void nms_mask_tile(int tile, int tt, const float4* __restrict__ sb, int base, int cn_total, int cwords,
                                              int kept0, float thresh, uint64_t* __restrict__ mask_intra,
                                              uint64_t* __restrict__ mask_prev, const float4* __restrict__ kept_boxes) {
  const int lrow = tt >> 2, qt = tt & 3;
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
  const int t0 = qt * 16, t1 = min(cn, t0 + 16);
  uint64_t bits = 0;
  const int r = rb * 64 + lrow;
  const bool row_ok = intra ? (r < cn_total) : (r < kept0);
  if (row_ok) {
    // kept_boxes is appended to between launches / cluster barriers: read it at L2 (another SM's L1 may hold a line
    // that was only partly written when it was cached for an earlier super-chunk)
    const float4 pb = intra ? sb[base + r] : __ldcg(kept_boxes + r);
    const float pa = nms_area(pb);
    // off-diagonal intra tiles: every q of the tile is later than p.  Diagonal tile: the FULL row (all q != p) -- the IoU test is
    // symmetric bit for bit, and the scan resolves a 64-box block through the bits of the EARLIER boxes of each row.
    const int self = intra ? (r - col0) : -1;
#pragma unroll 4
    for (int t = t0; t < t1; ++t) {
      const float4 q = __ldg(sb + base + col0 + t);
      if (t != self && nms_suppresses(pb, pa, q, nms_area(q), thresh)) bits |= (1ull << t);
    }
  }
  bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
  bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
  if (qt == 0) {
    if (intra) mask_intra[(size_t)r * NMS_W + cb] = bits;          // rows >= cn_total of the last block are zero (never read)
    else if (row_ok) mask_prev[(size_t)r * NMS_W + cb] = bits;
  }
}

// Scan of one super-chunk by one CTA (any multiple of 32 threads >= 64).  smem_mask: [NMS_M][NMS_W] words.
// Block c of 64 sorted boxes is resolved by warp 0 (fixpoint below) as soon as its "removed" word is final.  The removed word of
// the NEXT block gets the rows of block c's survivors from warp 0 itself (critical path: two LDS + a warp OR-reduction); the
// words of all later blocks are accumulated meanwhile by the other warps, one word per warp, so the serial chain per block is
// fixpoint + reduction + one barrier instead of a dependent shared-memory walk over every survivor.
__device__ __forceinline__ uint64_t nms_warp_or64(uint64_t v) {
  const uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)v), hi = __reduce_or_sync(0xffffffffu, (uint32_t)(v >> 32));
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void nms_scan_chunk(uint64_t* smem_mask, uint64_t* s_removed, uint64_t* s_keep, int* s_pos, int* s_cnt,
                                               const float4* __restrict__ sb, const int32_t* __restrict__ order, int base, int cn_total, int cwords,
                                               int kept0, int cap, const uint64_t* __restrict__ mask_intra, const uint64_t* __restrict__ mask_prev,
                                               float4* __restrict__ kept_boxes, NmsState* __restrict__ state, int64_t* __restrict__ idx_out,
                                               int32_t* __restrict__ idx32_out, int32_t* __restrict__ count_out) {
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
  {
    // chunk mask -> shared memory, 8 independent loads in flight per thread (a plain copy loop serialises on L2 latency)
    const int total_words = cn_total * NMS_W;
    for (int i0 = tid; i0 < total_words; i0 += nthreads * 8) {
      uint64_t v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * nthreads;
        v[u] = (i < total_words) ? __ldcg(mask_intra + i) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * nthreads;
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
    const int rstep = nthreads / NMS_W;
    if (w < cwords) {
      for (int r0 = tid >> 4; r0 < kept0; r0 += rstep * 8) {
        uint64_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = r0 + u * rstep;
          v[u] = (r < kept0) ? __ldcg(mask_prev + (size_t)r * NMS_W + w) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc |= v[u];
      }
    }
    if (acc) atomicOr((unsigned long long*)&s_removed[w], (unsigned long long)acc);
  }
  if (tid == 0) *s_cnt = kept0;
  __syncthreads();
  // every thread carries its own copy of the running count, refreshed BETWEEN the two barriers of an iteration: reading *s_cnt at the top
  // of the loop raced with warp 0's update of the same iteration (compute-sanitizer racecheck), and a late reader could have left the loop
  // ahead of the barriers its siblings still wait at
  int cnt = kept0;
  for (int c = 0; c < cwords; ++c) {
    if (cnt >= cap) break;                          // uniform
    const int lim = min(64, cn_total - c * 64);
    if (warp == 0) {
      const int before = cnt;
      const uint64_t d0 = (lane < lim) ? smem_mask[(size_t)(c * 64 + lane) * NMS_W + c] : 0ull;
      const uint64_t d1 = (lane + 32 < lim) ? smem_mask[(size_t)(c * 64 + 32 + lane) * NMS_W + c] : 0ull;
      const uint64_t cur = s_removed[c];
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
      if (before + nk > cap) {                      // max_boxes reached inside this block: keep the first (cap - count) survivors
        uint64_t rest = keep;
        for (int z = cap - before; z > 0; --z) rest &= rest - 1;
        keep &= ~rest;
        nk = cap - before;
      }
      // remember the sorted positions of the survivors (index translation + box copy happen after the loop, in parallel)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int i = lane + 32 * half;
        if ((keep >> i) & 1ull) s_pos[before - kept0 + __popcll(keep & ((1ull << i) - 1ull))] = base + c * 64 + i;
      }
      // rows of this block's survivors -> removed word of the NEXT block (the only word on the critical path)
      if (c + 1 < cwords) {
        uint64_t acc = 0;
        if ((keep >> lane) & 1ull) acc |= smem_mask[(size_t)(c * 64 + lane) * NMS_W + c + 1];
        if ((keep >> (lane + 32)) & 1ull) acc |= smem_mask[(size_t)(c * 64 + 32 + lane) * NMS_W + c + 1];
        acc = nms_warp_or64(acc);
        if (lane == 0) s_removed[c + 1] |= acc;
      }
      if (lane == 0) { s_keep[0] = keep; *s_cnt = before + nk; }
    }
    __syncthreads();
    cnt = *s_cnt;                                   // settled: written before the barrier above, next written after the barrier below
    // the other warps: rows of block c's survivors -> removed words c+2 .. cwords-1, one word per warp (warp 0 joins late)
    {
      const uint64_t keep = s_keep[0];
      for (int w = c + 2 + warp; w < cwords; w += nwarps) {
        uint64_t acc = 0;
        if ((keep >> lane) & 1ull) acc |= smem_mask[(size_t)(c * 64 + lane) * NMS_W + w];
        if ((keep >> (lane + 32)) & 1ull) acc |= smem_mask[(size_t)(c * 64 + 32 + lane) * NMS_W + w];
        acc = nms_warp_or64(acc);
        if (lane == 0 && acc) s_removed[w] |= acc;
      }
    }
    __syncthreads();                                // s_keep / s_removed[w] settled before the next block is resolved
  }
  const int count = *s_cnt;
  // emit survivors in scan order (= decreasing score) and keep their boxes for later super-chunks
  for (int e = tid; e < count - kept0; e += nthreads) {
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

// wide phase A of super-chunk s: one 256-thread block per 64x64 tile
__global__ void __launch_bounds__(NMS_TILE_THREADS)
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
constexpr int NMS_SCAN_THREADS = 1024;
__global__ void __launch_bounds__(NMS_SCAN_THREADS, 1)
nms_scan_kernel(int s, const float4* __restrict__ sb, const int32_t* __restrict__ order, int n_max, const int32_t* __restrict__ n_dev,
                int max_boxes, const uint64_t* __restrict__ mask_intra, const uint64_t* __restrict__ mask_prev, float4* __restrict__ kept_boxes,
                NmsState* __restrict__ state, int64_t* __restrict__ idx_out, int32_t* __restrict__ idx32_out, int32_t* __restrict__ count_out) {
  extern __shared__ __align__(16) uint64_t smem_mask[];     // [NMS_M][NMS_W] then s_pos[NMS_M]
  __shared__ uint64_t s_removed[NMS_W];
  __shared__ uint64_t s_keep[1];
  __shared__ int s_cnt;
  int* s_pos = reinterpret_cast<int*>(smem_mask + (size_t)NMS_M * NMS_W);
  const int n = n_dev ? min(*n_dev, n_max) : n_max;
  const int cap = (max_boxes > 0) ? max_boxes : n;
  const int base = s * NMS_M;
  const int kept0 = state->count;
  if (base >= n || kept0 >= cap) return;
  const int cn_total = min(NMS_M, n - base);
  nms_scan_chunk(smem_mask, s_removed, s_keep, s_pos, &s_cnt, sb, order, base, cn_total, DC_CEIL_DIV(cn_total, 64), kept0, cap, mask_intra, mask_prev,
                 kept_boxes, state, idx_out, idx32_out, count_out);
}

// tail: all remaining super-chunks in one launch of a single cluster (deep NMS only).  CTA 0 extends the sorted order by
// another NMS_SEL positions whenever the scan has consumed what is there.
__global__ void __launch_bounds__(NMS_THREADS, 1)
nms_cluster_kernel(int first_chunk, const uint64_t* __restrict__ keys, const float* __restrict__ boxes, int box_stride,
                   float4* __restrict__ sb, int32_t* __restrict__ order, int n_max, const int32_t* __restrict__ n_dev, float thresh, int max_boxes,
                   uint64_t* __restrict__ mask_intra, uint64_t* __restrict__ mask_prev, float4* __restrict__ kept_boxes, NmsState* __restrict__ state,
                   int64_t* __restrict__ idx_out, int32_t* __restrict__ idx32_out, int32_t* __restrict__ count_out) {
  extern __shared__ __align__(16) uint64_t smem_mask[];     // [NMS_M][NMS_W], s_pos[NMS_M], NmsSelSmem
  __shared__ uint64_t s_removed[NMS_W];
  __shared__ uint64_t s_keep[1];
  __shared__ int s_cnt;
  int* s_pos = reinterpret_cast<int*>(smem_mask + (size_t)NMS_M * NMS_W);
  NmsSelSmem& sel = *reinterpret_cast<NmsSelSmem*>(s_pos + NMS_M);
  const int n = n_dev ? min(*n_dev, n_max) : n_max;
  const int cap = (max_boxes > 0) ? max_boxes : n;
  const int tid = threadIdx.x;
  const uint32_t rank_in_cluster = nms_cluster_ctarank();
  const int num_workers = NMS_CTAS * (NMS_THREADS / NMS_TILE_THREADS);
  const int worker = (int)rank_in_cluster * (NMS_THREADS / NMS_TILE_THREADS) + (tid / NMS_TILE_THREADS);
  for (int base = first_chunk * NMS_M; base < n; base += NMS_M) {
    const int kept0 = __ldcg(&state->count);                 // written by CTA 0 before the previous barrier / launch (read at L2)
    if (kept0 >= cap) break;                                 // identical in every CTA of the cluster
    if (__ldcg(&state->sorted_n) <= base) {                  // identical in every CTA: the order must be extended first
      if (rank_in_cluster == 0) nms_select_sort<NMS_THREADS>(sel, keys, n, boxes, box_stride, state, order, sb);
      __threadfence();
      nms_cluster_sync();
    }
    const int cn_total = min(NMS_M, n - base);
    const int cwords = DC_CEIL_DIV(cn_total, 64);
    const int total_tiles = cwords * (cwords + 1) / 2 + DC_CEIL_DIV(kept0, 64) * cwords;
    for (int tile = worker; tile < total_tiles; tile += num_workers)
      nms_mask_tile(tile, tid % NMS_TILE_THREADS, sb, base, cn_total, cwords, kept0, thresh, mask_intra, mask_prev, kept_boxes);
    __threadfence();
    nms_cluster_sync();
    if (rank_in_cluster == 0)
      nms_scan_chunk(smem_mask, s_removed, s_keep, s_pos, &s_cnt, sb, order, base, cn_total, cwords, kept0, cap, mask_intra, mask_prev, kept_boxes, state,
                     idx_out, idx32_out, count_out);
    nms_cluster_sync();
  }
}

int k_nms(dc_ctx* ctx, const float* boxes, int box_stride, const float* scores, int score_stride, int n_max, const int32_t* n_dev,
          float thresh, int max_boxes, NmsWorkspace ws, int64_t* idx_out, int32_t* idx32_out, int32_t* count_out, cudaStream_t st) {
  NmsState* state = (NmsState*)ws.state;
  if (n_max <= 0) {
    nms_keys_kernel<<<1, 32, 0, st>>>(scores, score_stride, 0, nullptr, ws.keys, state, count_out);
    DC_CHECK_LAUNCH(ctx);
    return DC_OK;
  }
  DC_REQUIRE(ctx, n_max <= ws.cap, DC_E_INVALID, "nms: %d boxes exceed workspace capacity %d", n_max, ws.cap);
  const size_t smem_scan = (size_t)NMS_M * NMS_W * sizeof(uint64_t) + (size_t)NMS_M * sizeof(int);
  const size_t smem_cluster = smem_scan + sizeof(NmsSelSmem);
  const size_t smem_sel = sizeof(NmsSelSmem);
  {
    cudaError_t e = dc_ensure_smem(ctx, nms_cluster_kernel, smem_cluster);
    if (e == cudaSuccess) e = dc_ensure_smem(ctx, nms_scan_kernel, smem_scan);
    if (e == cudaSuccess) e = dc_ensure_smem(ctx, nms_select_kernel, smem_sel);
    if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "nms smem attribute: %s", cudaGetErrorString(e));
  }
  nms_keys_kernel<<<DC_CEIL_DIV(n_max, 256), 256, 0, st>>>(scores, score_stride, n_max, n_dev, ws.keys, state, count_out);
  DC_CHECK_LAUNCH(ctx);
  // sorted positions [0, NMS_SEL) = the first NMS_WIDE_CHUNKS super-chunks
  nms_select_kernel<<<1, NMS_SEL_THREADS, smem_sel, st>>>(ws.keys, n_max, n_dev, boxes, box_stride, state, ws.order, (float4*)ws.sorted_boxes);
  DC_CHECK_LAUNCH(ctx);
  const int chunks = DC_CEIL_DIV(n_max, NMS_M);
  const int max_keep = (max_boxes > 0) ? (max_boxes < n_max ? max_boxes : n_max) : n_max;
  const int wide = chunks < NMS_WIDE_CHUNKS ? chunks : NMS_WIDE_CHUNKS;
  for (int s = 0; s < wide; ++s) {
    const int prev_rows = (s == 0) ? 0 : (max_keep < s * NMS_M ? max_keep : s * NMS_M);
    const int tiles = NMS_W * (NMS_W + 1) / 2 + DC_CEIL_DIV(prev_rows, 64) * NMS_W;
    nms_mask_kernel<<<tiles, NMS_TILE_THREADS, 0, st>>>(s, (const float4*)ws.sorted_boxes, n_max, n_dev, thresh, max_boxes, ws.mask, ws.mask_prev,
                                                        (const float4*)ws.kept_boxes, state);
    DC_CHECK_LAUNCH(ctx);
    nms_scan_kernel<<<1, NMS_SCAN_THREADS, smem_scan, st>>>(s, (const float4*)ws.sorted_boxes, ws.order, n_max, n_dev, max_boxes, ws.mask, ws.mask_prev,
                                                            (float4*)ws.kept_boxes, state, idx_out, idx32_out, count_out);
    DC_CHECK_LAUNCH(ctx);
  }
  if (chunks > wide) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(NMS_CTAS);
    cfg.blockDim = dim3(NMS_THREADS);
    cfg.dynamicSmemBytes = smem_cluster;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NMS_CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, nms_cluster_kernel, wide, (const uint64_t*)ws.keys, boxes, box_stride, (float4*)ws.sorted_boxes, ws.order,
                                       n_max, n_dev, thresh, max_boxes, ws.mask, ws.mask_prev, (float4*)ws.kept_boxes, state, idx_out, idx32_out, count_out);
    if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "nms cluster launch: %s", cudaGetErrorString(e));
    DC_CHECK_LAUNCH(ctx);
  }
  return DC_OK;
}
