// tcgen05 engine: C[M,N] = A[M,K] . B[N,K]^T on 5th-gen tensor cores with error-compensated
// split-fp16 operands, used for the VGG-16 3x3 convolutions (implicit GEMM), the RPN convs, fc6/fc7,
// the LSTM gate GEMM and the vocabulary projection of the DenseCap path.
//
// Numerics: the reference computes these layers with fp32 sgemm.  Every fp32 operand x is held as two
// fp16 planes (hi = rn(x), lo = rn(x - hi)); the kernel issues three kind::f16 MMAs per K step,
//   D += A_hi*B_hi ; D += A_hi*B_lo ; D += A_lo*B_hi      (fp32 accumulation in TMEM)
// which drops only the lo*lo term (<= 2^-22 relative) -- fp32-class accuracy at one third of the
// dense fp16 tensor rate, instead of the 2^-11 error of a single fp16/tf32 pass.
//
// Kernel structure (one persistent CTA per SM, 192 threads, no cluster):
//   warp 0   : TMA producer -- cp.async.bulk.tensor loads of the A and B stage tiles (both planes in one
//              box) into 128B-swizzled shared memory, signalled through full[] mbarriers;
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit releases the smem stage
//              (empty[]) and publishes the finished accumulator (tmem_full[]);
//   warps 2-5: epilogue -- tcgen05.ld of the 128xBLOCK_N fp32 accumulator (one TMEM lane quarter per
//              warp), bias + ReLU (+ fused 2x2 ceil-mode max-pool through warp shuffles for the VGG
//              layers that precede a pool), re-split to fp16 planes / fp32 / row arg-max, vectorised stores.
//   The accumulator is double buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile i overlaps
//   the MMAs of tile i+1.
//
// Implicit GEMM for the 3x3 convolution: activations are NHWC planes; an M tile is an 8x16 pixel patch
// of one image, so that the A tile of filter tap (r,s) and channel block c0 is exactly the 5-D TMA box
// {64 ch, 16 w, 8 h, 1 n, 2 planes} at (c0, w0+s-1, h0+r-1, n, 0); the zero padding of the convolution is
// TMA's out-of-bounds zero fill.  K runs over 9 taps x Cin/64 channel blocks.
#include "common.cuh"

namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;     // fp16 elements: 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int TILE_H = 8, TILE_W = 16;   // conv M tile = 8x16 output pixels
constexpr uint32_t A_PLANE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;
static int g_num_sms = 148;

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (sticky CUDA error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {   // ~2 s at 2 GHz
      printf("densecap_b200: mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, kind::f16 (fp16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <- lane base+i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B, one 128-byte swizzle atom along K:
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4 (unused here)
//   bits [32,46) stride byte offset >> 4 = 1024 B between 8-row groups
//   bits [46,48) descriptor version = 1 (sm_100)      bits [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (kind::f16): c_format=F32 (bits 4-5 = 1), a/b format = F16 (0), K-major A and B,
// n_dim = N>>3 at bits [17,23), m_dim = M>>4 at bits [24,29).
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ unsigned long long pack_argmax_tc(float v, int col) {
  uint32_t u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)col);
}

// ------------------------------------------------------------------------------------------------
struct Problem {
  int M, N, K;                 // GEMM view (conv: M = tiles*128 is implicit, N = Cout, K = 9*Cin)
  int num_k_blocks;
  int m_tiles, n_tiles;
  int m_fastest;               // tile order
  // conv geometry
  int H, W, Cin, kb_per_tap;   // kb_per_tap = Cin/64
  int tiles_h, tiles_w;
  int OH, OW;                  // pooled output size (POOL) or H, W
};

template <int BLOCK_N>
struct Cfg {
  static constexpr uint32_t B_PLANE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr uint32_t STAGE_BYTES = 2 * A_PLANE_BYTES + 2 * B_PLANE_BYTES;
  static constexpr int STAGES = (BLOCK_N <= 64) ? 4 : (BLOCK_N <= 128 ? 3 : 2);
  static constexpr uint32_t TMEM_COLS = (2 * BLOCK_N < 32) ? 32 : 2 * BLOCK_N;   // power of two for 64/128/256
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 2 * BLOCK_N * 4 /*bias*/;
};

template <int BLOCK_N, bool CONV, bool POOL>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, Problem pb, GemmEpilogue ep) {
  using C = Cfg<BLOCK_N>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte alignment required by SWIZZLE_128B
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* smem_aligned = smem_dyn + (smem_base - smem_u32(smem_dyn));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_aligned + (size_t)STAGES * C::STAGE_BYTES);
  // bars: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then tmem base address holder
  const uint32_t bar_base = smem_base + STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* bias_smem = reinterpret_cast<float*>(smem_aligned + (size_t)STAGES * C::STAGE_BYTES + 256);   // [2][BLOCK_N]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_holder), C::TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int num_tiles = pb.m_tiles * pb.n_tiles;

  auto tile_coords = [&](int tile, int& mt, int& nt) {
    if (pb.m_fastest) { mt = tile % pb.m_tiles; nt = tile / pb.m_tiles; }
    else { nt = tile % pb.n_tiles; mt = tile / pb.n_tiles; }
  };

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int mt, nt;
        tile_coords(tile, mt, nt);
        const int n0 = nt * BLOCK_N;
        int img = 0, h0 = 0, w0 = 0;
        if (CONV) {
          const int per_img = pb.tiles_h * pb.tiles_w;
          img = mt / per_img;
          const int r = mt % per_img;
          h0 = (r / pb.tiles_w) * TILE_H;
          w0 = (r % pb.tiles_w) * TILE_W;
        }
        for (int kb = 0; kb < pb.num_k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t a_dst = smem_base + stage * C::STAGE_BYTES;
          const uint32_t b_dst = a_dst + 2 * A_PLANE_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
          if (CONV) {
            const int tap = kb / pb.kb_per_tap;
            const int c0 = (kb % pb.kb_per_tap) * BLOCK_K;
            const int r = tap / 3, s = tap % 3;
            tma_load_5d(a_dst, &tmap_a, full_bar(stage), c0, w0 + s - 1, h0 + r - 1, img, 0);
          } else {
            tma_load_3d(a_dst, &tmap_a, full_bar(stage), kb * BLOCK_K, mt * BLOCK_M, 0);
          }
          tma_load_3d(b_dst, &tmap_b, full_bar(stage), kb * BLOCK_K, n0, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BLOCK_M, BLOCK_N);
      int stage = 0; uint32_t phase = 0;
      int iter = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < pb.num_k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t a_hi = smem_base + stage * C::STAGE_BYTES;
          const uint32_t a_lo = a_hi + A_PLANE_BYTES;
          const uint32_t b_hi = a_hi + 2 * A_PLANE_BYTES;
          const uint32_t b_lo = b_hi + C::B_PLANE_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint32_t koff = k * UMMA_K * 2;   // bytes along K inside the swizzle atom
            const uint64_t da_hi = make_kmajor_sw128_desc(a_hi + koff);
            const uint64_t da_lo = make_kmajor_sw128_desc(a_lo + koff);
            const uint64_t db_hi = make_kmajor_sw128_desc(b_hi + koff);
            const uint64_t db_lo = make_kmajor_sw128_desc(b_lo + koff);
            umma_f16(d_tmem, da_lo, db_hi, idesc, (kb | k) != 0);   // small terms first
            umma_f16(d_tmem, da_hi, db_lo, idesc, 1);
            umma_f16(d_tmem, da_hi, db_hi, idesc, 1);
          }
          umma_commit(empty_bar(stage));                       // smem stage free once these MMAs retire
          if (kb == pb.num_k_blocks - 1) umma_commit(tfull_bar(acc));   // accumulator complete
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // =============================== epilogue warps ===============================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row_in_tile = q * 32 + lane;
    int iter = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
      int mt, nt;
      tile_coords(tile, mt, nt);
      const int n0 = nt * BLOCK_N;
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      // ---- which output row does this thread own? ----
      long long out_row = -1;     // row index into the (rows, ld) output matrices, -1 = masked
      bool in_bounds = false;     // accumulator row corresponds to a real element
      if (CONV) {
        const int per_img = pb.tiles_h * pb.tiles_w;
        const int img = mt / per_img;
        const int r = mt % per_img;
        const int h0 = (r / pb.tiles_w) * TILE_H, w0 = (r % pb.tiles_w) * TILE_W;
        const int h = h0 + row_in_tile / TILE_W, w = w0 + row_in_tile % TILE_W;
        in_bounds = (h < pb.H) && (w < pb.W);
        if (!POOL) {
          if (in_bounds) out_row = ((long long)img * pb.H + h) * pb.W + w;
        } else {
          // after the shuffles lane (even w, even h) holds the 2x2 maximum
          const int ph = h >> 1, pw = w >> 1;
          if (((row_in_tile % TILE_W) & 1) == 0 && ((row_in_tile / TILE_W) & 1) == 0 && ph < pb.OH && pw < pb.OW)
            out_row = ((long long)img * pb.OH + ph) * pb.OW + pw;
        }
      } else {
        const int m = mt * BLOCK_M + row_in_tile;
        in_bounds = m < pb.M;
        if (in_bounds) out_row = m;
      }
      float best_v = -INFINITY;
      int best_c = 0x7fffffff;

      // Stage this tile's bias slice in shared memory (one coalesced load per thread, overlapped with the MMAs of
      // this tile) -- per-element global loads in the column loop would serialise on L2 latency and make every
      // short-K tile epilogue-bound.  Double-buffered by accumulator stage, so one named barrier per tile suffices.
      float* sbias = bias_smem + acc * BLOCK_N;
      if (ep.bias) {
        for (int j = threadIdx.x - 64; j < BLOCK_N; j += 128) sbias[j] = (n0 + j < pb.N) ? __ldg(ep.bias + n0 + j) : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }

      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t raw[32];
        tmem_ld_32x32(t_row + c * 32, raw);
        tmem_ld_wait();
        const int nc = n0 + c * 32;
        if (nc >= pb.N) continue;           // whole chunk outside N (uniform across the CTA)
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(raw[j]);
          if (ep.bias) x += sbias[c * 32 + j];
          if (ep.relu) x = fmaxf(x, 0.f);
          v[j] = x;
        }
        if (POOL) {
          // ReLU output is >= 0, so 0 is the identity of max for out-of-image pixels (ceil-mode windows)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = in_bounds ? v[j] : 0.f;
            x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 1));
            x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 16));
            v[j] = x;
          }
        }
        if (ep.lstm_c) {
          // torch-rnn nn.LSTM cell on interleaved gate columns: 8 hidden units per 32-column chunk
          if (out_row >= 0) {
            const int j0 = nc >> 2;
            float* cp = ep.lstm_c + (size_t)out_row * ep.lstm_H + j0;
            float cprev[8];
            if (ep.lstm_first) {
#pragma unroll
              for (int u = 0; u < 8; ++u) cprev[u] = 0.f;
            } else {
              const float4 c0 = reinterpret_cast<const float4*>(cp)[0], c1 = reinterpret_cast<const float4*>(cp)[1];
              cprev[0] = c0.x; cprev[1] = c0.y; cprev[2] = c0.z; cprev[3] = c0.w;
              cprev[4] = c1.x; cprev[5] = c1.y; cprev[6] = c1.z; cprev[7] = c1.w;
            }
            float cn[8];
            uint32_t hh[4], ll[4];
            __half hhv[8], hlv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const float ig = 1.f / (1.f + expf(-v[4 * u + 0]));
              const float fg = 1.f / (1.f + expf(-v[4 * u + 1]));
              const float og = 1.f / (1.f + expf(-v[4 * u + 2]));
              const float gg = tanhf(v[4 * u + 3]);
              cn[u] = fg * cprev[u] + ig * gg;
              split_f32(og * tanhf(cn[u]), hhv[u], hlv[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              hh[u] = (uint32_t)__half_as_ushort(hhv[2 * u]) | ((uint32_t)__half_as_ushort(hhv[2 * u + 1]) << 16);
              ll[u] = (uint32_t)__half_as_ushort(hlv[2 * u]) | ((uint32_t)__half_as_ushort(hlv[2 * u + 1]) << 16);
            }
            reinterpret_cast<float4*>(cp)[0] = make_float4(cn[0], cn[1], cn[2], cn[3]);
            reinterpret_cast<float4*>(cp)[1] = make_float4(cn[4], cn[5], cn[6], cn[7]);
            const size_t ho = (size_t)out_row * ep.ldo + ep.out_col0 + j0;
            *reinterpret_cast<uint4*>(ep.out_hi + ho) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
            *reinterpret_cast<uint4*>(ep.out_lo + ho) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
          }
        } else if (out_row >= 0) {
          const bool full = (nc + 32 <= pb.N);
          if (ep.out_hi) {
            __half* ph = ep.out_hi + (size_t)out_row * ep.ldo + ep.out_col0 + nc;
            __half* pl = ep.out_lo + (size_t)out_row * ep.ldo + ep.out_col0 + nc;
            if (full) {
              uint32_t hh[16], ll[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                __half h0, l0, h1, l1;
                split_f32(v[2 * j], h0, l0);
                split_f32(v[2 * j + 1], h1, l1);
                hh[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                ll[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                reinterpret_cast<uint4*>(ph)[j] = make_uint4(hh[4 * j], hh[4 * j + 1], hh[4 * j + 2], hh[4 * j + 3]);
                reinterpret_cast<uint4*>(pl)[j] = make_uint4(ll[4 * j], ll[4 * j + 1], ll[4 * j + 2], ll[4 * j + 3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {      // fully unrolled + predicated: v[] must stay in registers
                if (nc + j < pb.N) {
                  __half h0, l0;
                  split_f32(v[j], h0, l0);
                  ph[j] = h0;
                  pl[j] = l0;
                }
              }
            }
          }
          if (ep.out_f32) {
            float* pf = ep.out_f32 + (size_t)out_row * ep.ldf + nc;
            if (full && (ep.ldf & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                reinterpret_cast<float4*>(pf)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nc + j < pb.N) pf[j] = v[j];
            }
          }
          if (ep.argmax) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nc + j < pb.N && v[j] > best_v) { best_v = v[j]; best_c = nc + j; }
          }
        }
      }
      // TMEM reads of this accumulator are complete: hand it back to the MMA warp
      tcgen05_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (ep.argmax && out_row >= 0 && best_c != 0x7fffffff)
        atomicMax(ep.argmax + out_row, pack_argmax_tc(best_v, best_c));
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int encode_map(dc_ctx* ctx, CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box) {
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  DC_REQUIRE(ctx, g_encode != nullptr, DC_E_CUDA, "cuTensorMapEncodeTiled unavailable (tc_init not run)");
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DC_REQUIRE(ctx, r == CUDA_SUCCESS, DC_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
  return DC_OK;
}

template <int BLOCK_N, bool CONV, bool POOL>
static int launch(dc_ctx* ctx, const CUtensorMap& ma, const CUtensorMap& mb, const Problem& pb, const GemmEpilogue& ep, cudaStream_t st) {
  using C = Cfg<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<BLOCK_N, CONV, POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
    if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "tc_gemm smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = pb.m_tiles * pb.n_tiles;
  const int grid = tiles < g_num_sms ? tiles : g_num_sms;
  tc_gemm_kernel<BLOCK_N, CONV, POOL><<<grid, NUM_THREADS, C::SMEM_BYTES, st>>>(ma, mb, pb, ep);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

// Tile width by a bandwidth model: the kernel is bound by the per-SM operand fill rate (~45 B/clk measured), so the
// cost of a launch is  rounds x bytes-per-k-block  with rounds = ceil(tiles / SMs); wider tiles amortise the A operand.
static int pick_block_n(int N, int m_tiles) {
  const int cand[3] = {64, 128, 256};
  int best = 64;
  double best_cost = 1e30;
  for (int i = 0; i < 3; ++i) {
    const int bn = cand[i];
    if (bn > 64 && N <= bn / 2) continue;                    // mostly-empty tile
    const long tiles = (long)m_tiles * DC_CEIL_DIV(N, bn);
    const long rounds = DC_CEIL_DIV(tiles, (long)g_num_sms);
    const double cost = (double)rounds * (2.0 * A_PLANE_BYTES + 2.0 * bn * BLOCK_K * 2);
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace tc

int tc_init(dc_ctx* ctx) {
  using namespace tc;
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
      return dc_set_error(ctx, DC_E_CUDA, "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
    g_encode = (PFN_encodeTiled)fn;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp prop;
  DC_CHECK_CUDA(ctx, cudaGetDeviceProperties(&prop, dev));
  DC_REQUIRE(ctx, prop.major == 10, DC_E_UNSUPPORTED, "tcgen05 engine needs an sm_100 device, found sm_%d%d", prop.major, prop.minor);
  g_num_sms = prop.multiProcessorCount;
  return DC_OK;
}

int k_tc_gemm(dc_ctx* ctx, const __half* a_planes, size_t a_plane_elems, int lda, const __half* b_planes, size_t b_plane_elems,
              int ldb, int M, int N, int K, GemmEpilogue ep, cudaStream_t st) {
  using namespace tc;
  if (M <= 0 || N <= 0) return DC_OK;
  DC_REQUIRE(ctx, K > 0 && K % BLOCK_K == 0, DC_E_UNSUPPORTED, "tc_gemm: K=%d must be a positive multiple of 64", K);
  DC_REQUIRE(ctx, lda % 8 == 0 && ldb % 8 == 0 && a_plane_elems % 8 == 0 && b_plane_elems % 8 == 0, DC_E_UNSUPPORTED,
             "tc_gemm: leading dimensions / plane sizes must be multiples of 8 elements");
  DC_REQUIRE(ctx, !ep.out_hi || (ep.ldo % 8 == 0 && ep.out_col0 % 8 == 0), DC_E_UNSUPPORTED, "tc_gemm: split output alignment");
  DC_REQUIRE(ctx, !ep.lstm_c || (ep.out_hi && N == 4 * ep.lstm_H && ep.lstm_H % 8 == 0), DC_E_UNSUPPORTED, "tc_gemm: LSTM epilogue shape");
  Problem pb{};
  pb.M = M; pb.N = N; pb.K = K;
  pb.num_k_blocks = K / BLOCK_K;
  pb.m_tiles = DC_CEIL_DIV(M, BLOCK_M);
  const int bn = pick_block_n(N, pb.m_tiles);
  pb.n_tiles = DC_CEIL_DIV(N, bn);
  pb.m_fastest = pb.m_tiles <= pb.n_tiles;
  CUtensorMap ma, mb;
  {
    const uint64_t dims[3] = {(uint64_t)K, (uint64_t)M, 2};
    const uint64_t str[2] = {(uint64_t)lda * 2, (uint64_t)a_plane_elems * 2};
    const uint32_t box[3] = {BLOCK_K, BLOCK_M, 2};
    int rc = encode_map(ctx, &ma, a_planes, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)K, (uint64_t)N, 2};
    const uint64_t str[2] = {(uint64_t)ldb * 2, (uint64_t)b_plane_elems * 2};
    const uint32_t box[3] = {BLOCK_K, (uint32_t)bn, 2};
    int rc = encode_map(ctx, &mb, b_planes, 3, dims, str, box);
    if (rc) return rc;
  }
  if (bn == 64) return launch<64, false, false>(ctx, ma, mb, pb, ep, st);
  if (bn == 256) return launch<256, false, false>(ctx, ma, mb, pb, ep, st);
  return launch<128, false, false>(ctx, ma, mb, pb, ep, st);
}

int k_tc_conv3x3(dc_ctx* ctx, const __half* x_planes, size_t x_plane_elems, const __half* w_planes, size_t w_plane_elems,
                 const float* bias, __half* y_hi, __half* y_lo, int N, int H, int W, int Cin, int Cout, int fuse_pool, cudaStream_t st) {
  using namespace tc;
  if (N <= 0 || H <= 0 || W <= 0) return DC_OK;
  DC_REQUIRE(ctx, Cin % BLOCK_K == 0, DC_E_UNSUPPORTED, "tc_conv3x3: Cin=%d must be a multiple of 64", Cin);
  DC_REQUIRE(ctx, Cout % 8 == 0 && x_plane_elems % 8 == 0 && w_plane_elems % 8 == 0, DC_E_UNSUPPORTED, "tc_conv3x3: alignment");
  Problem pb{};
  pb.H = H; pb.W = W; pb.Cin = Cin;
  pb.kb_per_tap = Cin / BLOCK_K;
  pb.K = 9 * Cin;
  pb.num_k_blocks = 9 * pb.kb_per_tap;
  pb.N = Cout;
  pb.tiles_h = DC_CEIL_DIV(H, TILE_H);
  pb.tiles_w = DC_CEIL_DIV(W, TILE_W);
  pb.m_tiles = N * pb.tiles_h * pb.tiles_w;
  pb.M = pb.m_tiles * BLOCK_M;
  pb.OH = fuse_pool ? (H + 1) / 2 : H;
  pb.OW = fuse_pool ? (W + 1) / 2 : W;
  const int bn = pick_block_n(Cout, pb.m_tiles);
  pb.n_tiles = DC_CEIL_DIV(Cout, bn);
  pb.m_fastest = 0;   // n fastest: CTAs sharing an activation patch run together, the patch stays in L2
  GemmEpilogue ep;
  ep.bias = bias;
  ep.relu = 1;
  ep.out_hi = y_hi;
  ep.out_lo = y_lo;
  ep.ldo = Cout;
  CUtensorMap ma, mb;
  {
    const uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N, 2};
    const uint64_t str[4] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2, (uint64_t)x_plane_elems * 2};
    const uint32_t box[5] = {BLOCK_K, TILE_W, TILE_H, 1, 2};
    int rc = encode_map(ctx, &ma, x_planes, 5, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)pb.K, (uint64_t)Cout, 2};
    const uint64_t str[2] = {(uint64_t)pb.K * 2, (uint64_t)w_plane_elems * 2};
    const uint32_t box[3] = {BLOCK_K, (uint32_t)bn, 2};
    int rc = encode_map(ctx, &mb, w_planes, 3, dims, str, box);
    if (rc) return rc;
  }
  if (fuse_pool) {
    if (bn == 64) return launch<64, true, true>(ctx, ma, mb, pb, ep, st);
    if (bn == 256) return launch<256, true, true>(ctx, ma, mb, pb, ep, st);
    return launch<128, true, true>(ctx, ma, mb, pb, ep, st);
  }
  if (bn == 64) return launch<64, true, false>(ctx, ma, mb, pb, ep, st);
  if (bn == 256) return launch<256, true, false>(ctx, ma, mb, pb, ep, st);
  return launch<128, true, false>(ctx, ma, mb, pb, ep, st);
}
