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
// Kernel structure (one persistent CTA per SM, 320 threads):
//   warp 0   : TMA producer -- cp.async.bulk.tensor loads of the A and B stage tiles (both planes in one
//              box) into 128B-swizzled shared memory, signalled through full[] mbarriers;
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit releases the smem stage
//              (empty[]) and publishes the finished accumulator (tmem_full[]);
//   warps 2-9: epilogue -- tcgen05.ld of the 128xBLOCK_N fp32 accumulator (two warps per TMEM lane quarter, half of the columns each;
//              warp), bias + ReLU (+ fused 2x2 ceil-mode max-pool through warp shuffles for the VGG
//              layers that precede a pool), re-split to fp16 planes / fp32 / row arg-max, vectorised stores.
//   The accumulator is double buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile i overlaps
//   the MMAs of tile i+1.
//
// CG = 2 variant (cta_group::2): two CTAs of a cluster (an SM pair) compute one 256 x 256 tile.  Each CTA
// loads its own 128 rows of A and HALF of the B tile (128 of the 256 rows); the leader CTA's single thread
// issues tcgen05.mma.cta_group::2 (M = 256) which reads A/B from both CTAs' shared memory and writes each
// CTA's own 128 x 256 accumulator.  Per CTA and K block that is 64 KB of operands for 1536 MMA cycles
// (42 B/clk) instead of 64 KB per 768 cycles for the 1-CTA 128 x 128 tile -- the kernel is bound by the
// ~45 B/clk/SM operand fill rate measured with ncu, so halving the bytes per MMA is what raises tensor
// utilisation.  Barrier protocol (same smem offsets in both CTAs): TMA of both CTAs signals the LEADER's
// full[] barrier (cta_group::2 TMA + peer-bit mask); tcgen05.commit multicasts to both CTAs' empty[] /
// tmem_full[]; the epilogue threads of both CTAs arrive on the leader's tmem_empty[] (remote arrive).
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
constexpr int EPI_WARPS = 8;                       // two epilogue warps per TMEM lane quarter, each draining half of the columns
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;   // warp 0 = TMA producer, warp 1 = MMA issuer, warps 2.. = epilogue
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
// cta_group::2 TMA: data lands in the executing CTA's smem, complete_tx is signalled on the LEADER CTA's barrier
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma2_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// arrive (no tx) on the barrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 remAddr32;\n\t"
      "mapa.shared::cluster.u32 remAddr32, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [remAddr32];\n\t}"
      ::"r"(bar), "r"(cta)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit of the leader's MMAs, arriving on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
// One lane of a fully converged warp.  The single-thread roles (TMA producer, MMA issuer) must be entered through
// elect.sync: with a plain `lane == 0` test the compiler cannot prove the region single-threaded and wraps every
// uniform-datapath instruction (UTCHMMA, UTCBAR, UBLKCP) in an ELECT/BRA.U.ANY waterfall loop, ~100 issue cycles per MMA.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
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

__device__ __forceinline__ float lstm_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float lstm_tanh(float x) {
  const float xc = fminf(fmaxf(x, -15.f), 15.f);          // tanh saturates to +-1 in fp32 beyond |x| ~ 9
  return 1.f - __fdividef(2.f, __expf(2.f * xc) + 1.f);
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

template <int BLOCK_N, int CG>
struct Cfg {
  static constexpr int B_ROWS = BLOCK_N / CG;                         // rows of the B tile held by one CTA
  static constexpr uint32_t B_PLANE_BYTES = B_ROWS * BLOCK_K * 2;
  static constexpr uint32_t STAGE_BYTES = 2 * A_PLANE_BYTES + 2 * B_PLANE_BYTES;
  static constexpr int STAGES = (STAGE_BYTES <= 49152) ? 4 : (STAGE_BYTES <= 65536 ? 3 : 2);
  // WIDE: one MMA multiplies A_hi with the hi AND lo plane of B at once (the two planes are adjacent in the stage, so they
  // form one 2*BLOCK_N-row K-major operand) -- see the MMA issuer.  The accumulator is then 2*BLOCK_N columns wide.
  static constexpr bool WIDE = (CG == 1 && BLOCK_N <= 128);
  static constexpr int ACC_COLS = WIDE ? 2 * BLOCK_N : BLOCK_N;
  static constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;                            // power of two: 256 or 512
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 2 * BLOCK_N * 4 /*bias*/;
};

// Epilogue warps (2-5) of both kernels: drain the finished accumulator of every tile this CTA owns.
// tfull0 / tempty0: shared addresses of the two tmem_full / tmem_empty barriers (8 bytes apart).
template <int BLOCK_N, bool CONV, bool POOL, int CG, bool WIDE>
__device__ __forceinline__ void epilogue_warps(const Problem& pb, const GemmEpilogue& ep, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0,
                                               float* bias_smem, int worker, int num_workers, uint32_t cta_rank, bool is_leader) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = pb.m_tiles * pb.n_tiles;
  auto tile_coords = [&](int tile, int& mt, int& nt) {
    if (pb.m_fastest) { mt = tile % pb.m_tiles; nt = tile / pb.m_tiles; }
    else { nt = tile % pb.n_tiles; mt = tile / pb.n_tiles; }
  };
  auto tfull_bar = [&](int s) { return tfull0 + 8u * s; };
  auto tempty_bar = [&](int s) { return tempty0 + 8u * s; };
  {
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int chunk_half = (warp - 2) >> 2; // warps 2-5 drain the first half of the accumulator columns, warps 6-9 the second
    const int row_in_tile = q * 32 + lane;
    int iter = 0;
    for (int tile = worker; tile < num_tiles; tile += num_workers, ++iter) {
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
        const int h0 = ((r / pb.tiles_w) * CG + (int)cta_rank) * TILE_H, w0 = (r % pb.tiles_w) * TILE_W;
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
        const int m = (mt * CG + (int)cta_rank) * BLOCK_M + row_in_tile;
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
        for (int j = threadIdx.x - 64; j < BLOCK_N; j += 32 * EPI_WARPS) sbias[j] = (n0 + j < pb.N) ? __ldg(ep.bias + n0 + j) : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
      }

      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
      constexpr int ACC_COLS = WIDE ? 2 * BLOCK_N : BLOCK_N;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS;
      constexpr int CHUNKS_PER_WARP = (BLOCK_N / 32) / (EPI_WARPS / 4);
#pragma unroll 1
      for (int c = chunk_half * CHUNKS_PER_WARP; c < (chunk_half + 1) * CHUNKS_PER_WARP; ++c) {
        uint32_t raw[32];
        tmem_ld_32x32(t_row + c * 32, raw);
        tmem_ld_wait();
        const int nc = n0 + c * 32;
        if (nc >= pb.N) continue;           // whole chunk outside N (uniform across the CTA)
        if (WIDE) {
          // columns [BLOCK_N, 2*BLOCK_N) hold the A_hi . B_lo partial sums of the same outputs
          uint32_t raw2[32];
          tmem_ld_32x32(t_row + BLOCK_N + c * 32, raw2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) raw[j] = __float_as_uint(__uint_as_float(raw[j]) + __uint_as_float(raw2[j]));
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(raw[j]) * ep.acc_scale;
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
              // sigmoid / tanh through ex2-based __expf (a few ulp for |x| <= 15, far inside the 1e-4 logit budget):
              // the accurate expf/tanhf sequences made this epilogue the critical path of the 16-k-block gate GEMM
              const float ig = lstm_sigmoid(v[4 * u + 0]);
              const float fg = lstm_sigmoid(v[4 * u + 1]);
              const float og = lstm_sigmoid(v[4 * u + 2]);
              const float gg = lstm_tanh(v[4 * u + 3]);
              cn[u] = fg * cprev[u] + ig * gg;
              split_f32(og * lstm_tanh(cn[u]), hhv[u], hlv[u]);
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
      if (CG == 2 && !is_leader) mbar_arrive_remote(tempty_bar(acc), 0);
      else mbar_arrive(tempty_bar(acc));
      if (ep.argmax && out_row >= 0 && best_c != 0x7fffffff)
        atomicMax(ep.argmax + out_row, pack_argmax_tc(best_v, best_c));
    }
  }
}

template <int BLOCK_N, bool CONV, bool POOL, int CG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, Problem pb, GemmEpilogue ep) {
  using C = Cfg<BLOCK_N, CG>;
  // CG == 2: the two CTAs of a cluster form one 256-row tile; rank 0 (leader) issues the MMAs
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = (cta_rank == 0);
  const int worker = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;        // tile-scheduler slot
  const int num_workers = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  if (CG == 2) cluster_sync_all();
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
      mbar_init(full_bar(s), CG);          // one producer arrival per CTA of the pair (leader's barrier is the live one)
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 32 * EPI_WARPS * CG);  // all epilogue threads of the pair arrive on the leader's barrier
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc2(smem_u32(tmem_holder), C::TMEM_COLS);
    else tmem_alloc(smem_u32(tmem_holder), C::TMEM_COLS);
  }
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the tail
  // of the previous kernel in the stream; from here on this kernel touches that kernel's outputs, so wait for it to finish.
  // Our own dependents may start their prologue as soon as SMs free up.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int num_tiles = pb.m_tiles * pb.n_tiles;

  auto tile_coords = [&](int tile, int& mt, int& nt) {
    if (pb.m_fastest) { mt = tile % pb.m_tiles; nt = tile / pb.m_tiles; }
    else { nt = tile % pb.n_tiles; mt = tile / pb.n_tiles; }
  };

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers) {
        int mt, nt;
        tile_coords(tile, mt, nt);
        const int n0 = nt * BLOCK_N + (int)cta_rank * C::B_ROWS;     // this CTA's slice of the B tile
        int img = 0, h0 = 0, w0 = 0;
        if (CONV) {
          const int per_img = pb.tiles_h * pb.tiles_w;               // CG == 2: tiles_h counts 16-row pair tiles
          img = mt / per_img;
          const int r = mt % per_img;
          h0 = ((r / pb.tiles_w) * CG + (int)cta_rank) * TILE_H;
          w0 = (r % pb.tiles_w) * TILE_W;
        }
        const int m0 = (mt * CG + (int)cta_rank) * BLOCK_M;
        for (int kb = 0; kb < pb.num_k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t a_dst = smem_base + stage * C::STAGE_BYTES;
          const uint32_t b_dst = a_dst + 2 * A_PLANE_BYTES;
          if (CG == 2) {
            if (is_leader) mbar_arrive_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);   // both CTAs' bytes land here
            else mbar_arrive_remote(full_bar(stage), 0);
          } else {
            mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
          }
          if (CONV) {
            const int tap = kb / pb.kb_per_tap;
            const int c0 = (kb % pb.kb_per_tap) * BLOCK_K;
            const int r = tap / 3, s = tap % 3;
            if (CG == 2) tma2_load_5d(a_dst, &tmap_a, full_bar(stage), c0, w0 + s - 1, h0 + r - 1, img, 0);
            else tma_load_5d(a_dst, &tmap_a, full_bar(stage), c0, w0 + s - 1, h0 + r - 1, img, 0);
          } else {
            if (CG == 2) tma2_load_3d(a_dst, &tmap_a, full_bar(stage), kb * BLOCK_K, m0, 0);
            else tma_load_3d(a_dst, &tmap_a, full_bar(stage), kb * BLOCK_K, m0, 0);
          }
          if (CG == 2) tma2_load_3d(b_dst, &tmap_b, full_bar(stage), kb * BLOCK_K, n0, 0);
          else tma_load_3d(b_dst, &tmap_b, full_bar(stage), kb * BLOCK_K, n0, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (is_leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(BLOCK_M * CG, BLOCK_N);
      constexpr uint32_t idesc_wide = make_idesc_f16(BLOCK_M, C::WIDE ? 2 * BLOCK_N : BLOCK_N);
      int stage = 0; uint32_t phase = 0;
      int iter = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers, ++iter) {
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * C::ACC_COLS;
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
            if (C::WIDE) {
              // D[:, 0:BN] += A_hi.B_hi and D[:, BN:2BN] += A_hi.B_lo in ONE N = 2*BN instruction (B planes are adjacent), then
              // D[:, 0:BN] += A_lo.B_hi.  Every MMA reads its 128 x 32 B A slice from shared memory (~90 clk, measured); two
              // instructions instead of three per K step is what lifts the narrow tiles.
              umma_f16(d_tmem, da_hi, db_hi, idesc_wide, (kb | k) != 0);
              umma_f16(d_tmem, da_lo, db_hi, idesc, 1);
            } else if (CG == 2) {
              umma2_f16(d_tmem, da_lo, db_hi, idesc, (kb | k) != 0);
              umma2_f16(d_tmem, da_hi, db_lo, idesc, 1);
              umma2_f16(d_tmem, da_hi, db_hi, idesc, 1);
            } else {
              umma_f16(d_tmem, da_lo, db_hi, idesc, (kb | k) != 0);   // small terms first
              umma_f16(d_tmem, da_hi, db_lo, idesc, 1);
              umma_f16(d_tmem, da_hi, db_hi, idesc, 1);
            }
          }
          if (CG == 2) {
            umma2_commit_mc(empty_bar(stage));                              // frees this stage in BOTH CTAs
            if (kb == pb.num_k_blocks - 1) umma2_commit_mc(tfull_bar(acc)); // accumulators of both CTAs complete
          } else {
            umma_commit(empty_bar(stage));                       // smem stage free once these MMAs retire
            if (kb == pb.num_k_blocks - 1) umma_commit(tfull_bar(acc));   // accumulator complete
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // =============================== epilogue warps ===============================
    epilogue_warps<BLOCK_N, CONV, POOL, CG, C::WIDE>(pb, ep, tmem_base, tfull_bar(0), tempty_bar(0), bias_smem, worker, num_workers, cta_rank, is_leader);
  }

  __syncwarp();     // lanes of the single-thread roles rejoin their warp before the aligned barriers below
  tcgen05_fence_before();
  // CG == 2: neither CTA may exit (or free TMEM) while the leader can still read its smem / signal its barriers
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    if (CG == 2) tmem_dealloc2(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// 3x3 convolution with vertical tap sharing ("halo" kernel).
//
// The plain implicit GEMM above re-fetches the activation patch for each of the 9 filter taps (9 x 32 KB per
// 64-channel block), and ncu shows the kernel bound by the per-SM operand fill rate.  Here the producer
// loads ONE {64 ch, 16 w, 10 h} halo box per (horizontal tap s, channel block): its rows are ordered (h, w), so
// the M = 128 rows of vertical tap r are simply the 128 rows starting 16*r rows (2048 B, a multiple of the
// 1024-byte swizzle atom) into the box -- three taps read the same shared-memory tile through descriptors
// that differ only in their start address.  A bytes per tile drop 2.4x (3 x 40 KB instead of 9 x 32 KB).
// A and B therefore run in two independent mbarrier rings: A ring = 2 stages x 40 KB (one fill per (s, cb)
// group), B ring = SB stages of one tap's weights each.  When a layer's whole filter bank fits the B ring
// (Cin = Cout = 64: 9 taps x 16 KB, conv1_2) it is loaded once per CTA and stays resident.
// ------------------------------------------------------------------------------------------------
constexpr int HALO_H = TILE_H + 2;
constexpr uint32_t AH_PLANE_BYTES = HALO_H * TILE_W * BLOCK_K * 2;      // 20480
constexpr uint32_t AH_STAGE_BYTES = 2 * AH_PLANE_BYTES;                 // 40960
constexpr int AH_STAGES = 2;

template <int BLOCK_N, int CG>
struct HaloCfg {
  static constexpr int B_ROWS = BLOCK_N / CG;
  static constexpr uint32_t B_PLANE_BYTES = B_ROWS * BLOCK_K * 2;
  static constexpr uint32_t B_STAGE_BYTES = 2 * B_PLANE_BYTES;
  static constexpr int SB = (B_STAGE_BYTES <= 16384) ? 9 : (B_STAGE_BYTES <= 32768 ? 4 : 2);
  static constexpr bool WIDE = (CG == 1 && BLOCK_N <= 128);
  static constexpr int ACC_COLS = WIDE ? 2 * BLOCK_N : BLOCK_N;
  static constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
  static constexpr uint32_t RING_BYTES = AH_STAGES * AH_STAGE_BYTES + SB * B_STAGE_BYTES;
  static constexpr size_t SMEM_BYTES = (size_t)RING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 2 * BLOCK_N * 4 /*bias*/;
};

template <int BLOCK_N, bool POOL, int CG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_conv_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, Problem pb, GemmEpilogue ep) {
  using C = HaloCfg<BLOCK_N, CG>;
  constexpr int SB = C::SB;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = (cta_rank == 0);
  const int worker = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int num_workers = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  if (CG == 2) cluster_sync_all();
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* smem_aligned = smem_dyn + (smem_base - smem_u32(smem_dyn));
  const uint32_t a_ring = smem_base;
  const uint32_t b_ring = smem_base + AH_STAGES * AH_STAGE_BYTES;
  // bars: a_full[2], a_empty[2], b_full[SB], b_empty[SB], tmem_full[2], tmem_empty[2], tmem base holder
  const uint32_t bar_base = smem_base + C::RING_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_aligned + C::RING_BYTES);
  auto afull_bar = [&](int s) { return bar_base + 8u * s; };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (AH_STAGES + s); };
  auto bfull_bar = [&](int s) { return bar_base + 8u * (2 * AH_STAGES + s); };
  auto bempty_bar = [&](int s) { return bar_base + 8u * (2 * AH_STAGES + SB + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * AH_STAGES + 2 * SB + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * AH_STAGES + 2 * SB + 2 + s); };
  static_assert((2 * AH_STAGES + 2 * SB + 4) * 8 + 8 <= 256, "barrier block overflow");
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * AH_STAGES + 2 * SB + 4);
  float* bias_smem = reinterpret_cast<float*>(smem_aligned + C::RING_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < AH_STAGES; ++s) { mbar_init(afull_bar(s), CG); mbar_init(aempty_bar(s), 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(bfull_bar(s), CG); mbar_init(bempty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 32 * EPI_WARPS * CG); }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc2(smem_u32(tmem_holder), C::TMEM_COLS);
    else tmem_alloc(smem_u32(tmem_holder), C::TMEM_COLS);
  }
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the tail
  // of the previous kernel in the stream; from here on this kernel touches that kernel's outputs, so wait for it to finish.
  // Our own dependents may start their prologue as soon as SMs free up.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int num_tiles = pb.m_tiles * pb.n_tiles;
  const int groups = 3 * pb.kb_per_tap;                                 // (s, channel block) groups per tile
  // whole filter bank of the (single) N tile fits the B ring: load it once, keep it
  const bool b_resident = (pb.n_tiles == 1) && (pb.num_k_blocks == SB);

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (elect_one()) {
      int sa = 0; uint32_t pa = 0;
      int sb = 0; uint32_t pbp = 0;
      bool first = true;
      for (int tile = worker; tile < num_tiles; tile += num_workers, first = false) {
        const int nt = tile % pb.n_tiles, mt = tile / pb.n_tiles;
        const int n0 = nt * BLOCK_N + (int)cta_rank * C::B_ROWS;
        const int per_img = pb.tiles_h * pb.tiles_w;
        const int img = mt / per_img;
        const int rr = mt % per_img;
        const int h0 = ((rr / pb.tiles_w) * CG + (int)cta_rank) * TILE_H;
        const int w0 = (rr % pb.tiles_w) * TILE_W;
        for (int g = 0; g < groups; ++g) {
          const int s = g / pb.kb_per_tap, cb = g % pb.kb_per_tap;
          mbar_wait(aempty_bar(sa), pa ^ 1);
          const uint32_t a_dst = a_ring + sa * AH_STAGE_BYTES;
          if (CG == 2) {
            if (is_leader) mbar_arrive_expect_tx(afull_bar(sa), 2 * AH_STAGE_BYTES);
            else mbar_arrive_remote(afull_bar(sa), 0);
            tma2_load_5d(a_dst, &tmap_a, afull_bar(sa), cb * BLOCK_K, w0 + s - 1, h0 - 1, img, 0);
          } else {
            mbar_arrive_expect_tx(afull_bar(sa), AH_STAGE_BYTES);
            tma_load_5d(a_dst, &tmap_a, afull_bar(sa), cb * BLOCK_K, w0 + s - 1, h0 - 1, img, 0);
          }
          if (++sa == AH_STAGES) { sa = 0; pa ^= 1; }
          if (b_resident && !first) continue;
          for (int r = 0; r < 3; ++r) {
            mbar_wait(bempty_bar(sb), pbp ^ 1);
            const uint32_t b_dst = b_ring + sb * C::B_STAGE_BYTES;
            const int k0 = ((r * 3 + s) * pb.kb_per_tap + cb) * BLOCK_K;
            if (CG == 2) {
              if (is_leader) mbar_arrive_expect_tx(bfull_bar(sb), 2 * C::B_STAGE_BYTES);
              else mbar_arrive_remote(bfull_bar(sb), 0);
              tma2_load_3d(b_dst, &tmap_b, bfull_bar(sb), k0, n0, 0);
            } else {
              mbar_arrive_expect_tx(bfull_bar(sb), C::B_STAGE_BYTES);
              tma_load_3d(b_dst, &tmap_b, bfull_bar(sb), k0, n0, 0);
            }
            if (++sb == SB) { sb = 0; pbp ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (is_leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(BLOCK_M * CG, BLOCK_N);
      constexpr uint32_t idesc_wide = make_idesc_f16(BLOCK_M, C::WIDE ? 2 * BLOCK_N : BLOCK_N);
      int sa = 0; uint32_t pa = 0;
      int sb = 0; uint32_t pbp = 0;
      int iter = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers, ++iter) {
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * C::ACC_COLS;
        for (int g = 0; g < groups; ++g) {
          mbar_wait(afull_bar(sa), pa);
          const uint32_t a_base = a_ring + sa * AH_STAGE_BYTES;
#pragma unroll 1
          for (int r = 0; r < 3; ++r) {
            if (!(b_resident && iter > 0)) mbar_wait(bfull_bar(sb), pbp);
            tcgen05_fence_after();
            const uint32_t a_hi = a_base + r * (TILE_W * BLOCK_K * 2);      // 16 rows = 2048 B down the halo box
            const uint32_t a_lo = a_hi + AH_PLANE_BYTES;
            const uint32_t b_hi = b_ring + sb * C::B_STAGE_BYTES;
            const uint32_t b_lo = b_hi + C::B_PLANE_BYTES;
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint32_t koff = k * UMMA_K * 2;
              const uint64_t da_hi = make_kmajor_sw128_desc(a_hi + koff);
              const uint64_t da_lo = make_kmajor_sw128_desc(a_lo + koff);
              const uint64_t db_hi = make_kmajor_sw128_desc(b_hi + koff);
              const uint64_t db_lo = make_kmajor_sw128_desc(b_lo + koff);
              const uint32_t accum = (g | r | k) != 0;
              if (C::WIDE) {
                umma_f16(d_tmem, da_hi, db_hi, idesc_wide, accum);
                umma_f16(d_tmem, da_lo, db_hi, idesc, 1);
              } else if (CG == 2) {
                umma2_f16(d_tmem, da_lo, db_hi, idesc, accum);
                umma2_f16(d_tmem, da_hi, db_lo, idesc, 1);
                umma2_f16(d_tmem, da_hi, db_hi, idesc, 1);
              } else {
                umma_f16(d_tmem, da_lo, db_hi, idesc, accum);
                umma_f16(d_tmem, da_hi, db_lo, idesc, 1);
                umma_f16(d_tmem, da_hi, db_hi, idesc, 1);
              }
            }
            const bool last = (g == groups - 1) && (r == 2);
            if (CG == 2) {
              if (!b_resident) umma2_commit_mc(bempty_bar(sb));
              if (r == 2) umma2_commit_mc(aempty_bar(sa));
              if (last) umma2_commit_mc(tfull_bar(acc));
            } else {
              if (!b_resident) umma_commit(bempty_bar(sb));
              if (r == 2) umma_commit(aempty_bar(sa));
              if (last) umma_commit(tfull_bar(acc));
            }
            if (++sb == SB) { sb = 0; pbp ^= 1; }
          }
          if (++sa == AH_STAGES) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else {
    // =============================== epilogue warps ===============================
    epilogue_warps<BLOCK_N, true, POOL, CG, C::WIDE>(pb, ep, tmem_base, tfull_bar(0), tempty_bar(0), bias_smem, worker, num_workers, cta_rank, is_leader);
  }

  __syncwarp();
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    if (CG == 2) tmem_dealloc2(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
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

template <typename Kernel>
static int launch_kernel(dc_ctx* ctx, Kernel kernel, bool& attr_set, size_t smem_bytes, int cg, const CUtensorMap& ma, const CUtensorMap& mb,
                         const Problem& pb, const GemmEpilogue& ep, cudaStream_t st) {
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "tc_gemm smem attribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = pb.m_tiles * pb.n_tiles;
  const int slots = g_num_sms / cg;                        // CTAs (CG=1) or CTA pairs (CG=2) resident at once
  const int workers = tiles < slots ? tiles : slots;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(workers * cg);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // see griddepcontrol.wait in the kernels
  attr[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
  if (cg == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, ma, mb, pb, ep);
  if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "tc_gemm launch: %s", cudaGetErrorString(e));
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

template <int BLOCK_N, bool CONV, bool POOL, int CG>
static int launch(dc_ctx* ctx, const CUtensorMap& ma, const CUtensorMap& mb, const Problem& pb, const GemmEpilogue& ep, cudaStream_t st) {
  static bool attr_set = false;
  return launch_kernel(ctx, tc_gemm_kernel<BLOCK_N, CONV, POOL, CG>, attr_set, Cfg<BLOCK_N, CG>::SMEM_BYTES, CG, ma, mb, pb, ep, st);
}

template <int BLOCK_N, bool POOL, int CG>
static int launch_halo(dc_ctx* ctx, const CUtensorMap& ma, const CUtensorMap& mb, const Problem& pb, const GemmEpilogue& ep, cudaStream_t st) {
  static bool attr_set = false;
  return launch_kernel(ctx, tc_conv_halo_kernel<BLOCK_N, POOL, CG>, attr_set, HaloCfg<BLOCK_N, CG>::SMEM_BYTES, CG, ma, mb, pb, ep, st);
}

// Tile shape by an MMA-issue model (ncu, r06): a tcgen05.mma of this kernel cannot retire faster than ~90 clk because it
// streams its 128 x 32 B A slice from shared memory, and it takes N/2 clk on the tensor pipe (M = 128 per CTA).  Per K = 16
// step a CTA spends
//   BLOCK_N =  64 (wide): max(64,90) + max(32,90)  = 180 clk for 128 x  64 outputs
//   BLOCK_N = 128 (wide): max(128,90) + max(64,90) = 218 clk for 128 x 128 outputs
//   BLOCK_N = 256       : 3 x 128                  = 384 clk for 128 x 256 outputs   (cta_group::2: same per CTA, half the B fill)
// and a launch costs rounds x that, rounds = ceil(tiles / resident CTAs).
struct TileCfg { int bn; int cg; };
static int g_allow_cg2 = -1;
static int g_allow_halo = -1;
static TileCfg pick_tile(int N, int m_tiles128, bool lstm) {
  if (g_allow_cg2 < 0) {
    const char* e = getenv("DC_TC_CG2");
    g_allow_cg2 = (e && e[0] == '0') ? 0 : 1;
  }
  TileCfg best{64, 1};
  double best_cost = 1e30;
  auto consider = [&](int bn, int cg, double kcyc) {
    if (bn > 64 && N <= bn / 2) return;                    // mostly-empty tile
    if (cg == 2 && (lstm || N <= 128 || m_tiles128 < 2)) return;
    const long tiles = (long)DC_CEIL_DIV(m_tiles128, cg) * DC_CEIL_DIV(N, bn);
    const long rounds = DC_CEIL_DIV(tiles, (long)(g_num_sms / cg));
    const double cost = (double)rounds * kcyc;
    if (cost < best_cost) { best_cost = cost; best = TileCfg{bn, cg}; }
  };
  consider(64, 1, 180.0);
  consider(128, 1, 218.0);
  consider(256, 1, 384.0);
  if (g_allow_cg2) consider(256, 2, 384.0 * 0.97);
  return best;
}
static TileCfg pick_tile_halo(int N, int m_tiles128) { return pick_tile(N, m_tiles128, false); }

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
  const int m_tiles128 = DC_CEIL_DIV(M, BLOCK_M);
  const TileCfg tc_ = pick_tile(N, m_tiles128, ep.lstm_c != nullptr);
  const int bn = tc_.bn, cg = tc_.cg;
  pb.m_tiles = DC_CEIL_DIV(m_tiles128, cg);
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
    const uint32_t box[3] = {BLOCK_K, (uint32_t)(bn / cg), 2};
    int rc = encode_map(ctx, &mb, b_planes, 3, dims, str, box);
    if (rc) return rc;
  }
  if (cg == 2) return launch<256, false, false, 2>(ctx, ma, mb, pb, ep, st);
  if (bn == 64) return launch<64, false, false, 1>(ctx, ma, mb, pb, ep, st);
  if (bn == 256) return launch<256, false, false, 1>(ctx, ma, mb, pb, ep, st);
  return launch<128, false, false, 1>(ctx, ma, mb, pb, ep, st);
}

int k_tc_conv3x3(dc_ctx* ctx, const __half* x_planes, size_t x_plane_elems, const __half* w_planes, size_t w_plane_elems,
                 const float* bias, float acc_scale, __half* y_hi, __half* y_lo, int N, int H, int W, int Cin, int Cout, int fuse_pool,
                 cudaStream_t st, int relu) {
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
  pb.tiles_w = DC_CEIL_DIV(W, TILE_W);
  const int m_tiles128 = N * DC_CEIL_DIV(H, TILE_H) * pb.tiles_w;
  if (g_allow_halo < 0) {
    const char* e = getenv("DC_TC_HALO");
    g_allow_halo = (e && e[0] == '0') ? 0 : 1;
    pick_tile(64, 1, false);                           // initialises g_allow_cg2
  }
  const bool halo = g_allow_halo != 0;
  const TileCfg tc_ = halo ? pick_tile_halo(Cout, m_tiles128) : pick_tile(Cout, m_tiles128, false);
  const int bn = tc_.bn, cg = tc_.cg;
  pb.tiles_h = DC_CEIL_DIV(H, TILE_H * cg);          // cg == 2: a pair tile covers 16 image rows (two stacked 8x16 patches)
  pb.m_tiles = N * pb.tiles_h * pb.tiles_w;
  pb.M = pb.m_tiles * BLOCK_M * cg;
  pb.OH = fuse_pool ? (H + 1) / 2 : H;
  pb.OW = fuse_pool ? (W + 1) / 2 : W;
  pb.n_tiles = DC_CEIL_DIV(Cout, bn);
  pb.m_fastest = 0;   // n fastest: CTAs sharing an activation patch run together, the patch stays in L2
  GemmEpilogue ep;
  ep.bias = bias;
  ep.acc_scale = acc_scale;
  ep.relu = relu;
  ep.out_hi = y_hi;
  ep.out_lo = y_lo;
  ep.ldo = Cout;
  CUtensorMap ma, mb;
  {
    const uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N, 2};
    const uint64_t str[4] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2, (uint64_t)x_plane_elems * 2};
    const uint32_t box[5] = {BLOCK_K, TILE_W, (uint32_t)(halo ? HALO_H : TILE_H), 1, 2};
    int rc = encode_map(ctx, &ma, x_planes, 5, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)pb.K, (uint64_t)Cout, 2};
    const uint64_t str[2] = {(uint64_t)pb.K * 2, (uint64_t)w_plane_elems * 2};
    const uint32_t box[3] = {BLOCK_K, (uint32_t)(bn / cg), 2};
    int rc = encode_map(ctx, &mb, w_planes, 3, dims, str, box);
    if (rc) return rc;
  }
  if (halo) {
    if (fuse_pool) {
      if (cg == 2) return launch_halo<256, true, 2>(ctx, ma, mb, pb, ep, st);
      if (bn == 64) return launch_halo<64, true, 1>(ctx, ma, mb, pb, ep, st);
      if (bn == 256) return launch_halo<256, true, 1>(ctx, ma, mb, pb, ep, st);
      return launch_halo<128, true, 1>(ctx, ma, mb, pb, ep, st);
    }
    if (cg == 2) return launch_halo<256, false, 2>(ctx, ma, mb, pb, ep, st);
    if (bn == 64) return launch_halo<64, false, 1>(ctx, ma, mb, pb, ep, st);
    if (bn == 256) return launch_halo<256, false, 1>(ctx, ma, mb, pb, ep, st);
    return launch_halo<128, false, 1>(ctx, ma, mb, pb, ep, st);
  }
  if (fuse_pool) {
    if (cg == 2) return launch<256, true, true, 2>(ctx, ma, mb, pb, ep, st);
    if (bn == 64) return launch<64, true, true, 1>(ctx, ma, mb, pb, ep, st);
    if (bn == 256) return launch<256, true, true, 1>(ctx, ma, mb, pb, ep, st);
    return launch<128, true, true, 1>(ctx, ma, mb, pb, ep, st);
  }
  if (cg == 2) return launch<256, true, false, 2>(ctx, ma, mb, pb, ep, st);
  if (bn == 64) return launch<64, true, false, 1>(ctx, ma, mb, pb, ep, st);
  if (bn == 256) return launch<256, true, false, 1>(ctx, ma, mb, pb, ep, st);
  return launch<128, true, false, 1>(ctx, ma, mb, pb, ep, st);
}
