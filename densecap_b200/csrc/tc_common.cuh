// Shared device / host helpers of the tcgen05 kernels (tc_gemm.cu: GEMM + implicit-GEMM convolution; lm_decode.cu: the
// persistent caption decoder): PTX wrappers, descriptors, tile configuration and the accumulator epilogue.
#pragma once
#include "common.cuh"

namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;     // fp16 elements: 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;                       // two epilogue warps per TMEM lane quarter, each draining half of the columns
constexpr int EPI_WARP0 = 4;                       // first epilogue warp
constexpr int EPI_THREAD0 = 32 * EPI_WARP0;
// Warp roles by WARPGROUP (setmaxnreg works on groups of 4 warps): warpgroup 0 = {warp 0: TMA producer, warp 1: MMA issuer + TMEM
// allocator, warps 2-3: idle}, warpgroups 1-2 = the 8 epilogue warps.  The kernels are compiled for 384 threads (168 registers); at
// start warpgroup 0 shrinks and the epilogue warpgroups grow (setmaxnreg), which is what lets an epilogue thread keep the 128 running sums of a
// 256-column tile (hierarchical accumulation, see epilogue_tile) in registers.
constexpr int NUM_THREADS = EPI_THREAD0 + 32 * EPI_WARPS;
// (128 control threads x C + 256 epilogue threads x E <= 384 x 168 = 64512 registers.)  ptxas allocates each role's code against the value its
// setmaxnreg announces.  Measured (A/B on B200): 56 / 224 removes the epilogue spills of the 256-column kernels (240 -> 24 bytes) at no cost to the
// control warps; 40 / 232 is another 1-5 % faster on the long-K pair-tile GEMM (fc6 / fc7) but makes the producer / issuer of the narrow convolution
// tiles spill (conv1_2 +6 %), so only tc_gemm_kernel<256, cta_group::2> uses it.
template <int REGS>
__device__ __forceinline__ void regs_shrink_control() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS>
__device__ __forceinline__ void regs_grow_epilogue() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }
constexpr int REGS_CONTROL = 56, REGS_EPILOGUE = 224;
constexpr int REGS_CONTROL_LONGK = 40, REGS_EPILOGUE_LONGK = 232;
constexpr int TILE_H = 8, TILE_W = 16;   // conv M tile = 8x16 output pixels
constexpr uint32_t A_PLANE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

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
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 256-bit global store (STG.256): p must be 32-byte aligned
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]),
               "r"(v[6]), "r"(v[7])
               : "memory");
}
// TMA tensor stores (global <- shared::cta) of the epilogue staging tiles, bulk-group completion
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(src) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(src) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
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
// Same for SWIZZLE_64B (64-byte rows: 32 fp16 along K per row, 8-row groups 512 B apart, layout type 4)
__device__ __forceinline__ uint64_t make_kmajor_sw64_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// Same for SWIZZLE_32B (32-byte rows: 16 fp16 along K per row, 8-row groups 256 B apart, layout type 6)
__device__ __forceinline__ uint64_t make_kmajor_sw32_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;
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
  int kb_per_group;            // k-blocks accumulated in TMEM before the partial sum is taken over by the epilogue (>= 1)
  int kb_first;                // convolution kernel: size of the first two groups of every tile (>= kb_per_group)
  int split_k;                 // convolution only: K splits per output tile (>= 1), partial sums to out_f32 + split * split_stride
  size_t split_stride;
  // GEMM kernel: this launch covers the tiles [tile_begin, tile_end) of the m_tiles x n_tiles grid (whole grid by default).  The TAIL
  // instantiation cuts each of its tiles into tail_split K-ranges and writes raw partial sums, see k_tc_gemm.
  int tile_begin, tile_end, tail_split;
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
  static constexpr bool TMAS = true;                    // split-plane outputs through TMA stores (8 staging slots behind the operand ring)
  static constexpr uint32_t STAGING_BYTES = 8 * 4096;
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ +
                                       (BLOCK_N == 256 ? 1 : 2) * BLOCK_N * 4 /*bias*/;
};

// Compile-time feature mask of the epilogue.  The output phase of a tile is straight-line code executed once per tile; with every feature
// compiled into every kernel the 256-column convolution kernel was 10,216 SASS instructions (163 KB, more than the instruction cache) and an ncu
// source profile showed its output phase stalled on instruction fetch (~13 us per tile).  Each kernel is therefore instantiated only with
// the epilogue features its callers use.
enum : int { F_SPLIT = 1, F_F32 = 2, F_LSTM = 4, F_ARGMAX = 8, F_GATHER = 16 };
constexpr int FEAT_CONV = F_SPLIT;                                   // bias + ReLU (+ pool) -> split planes
constexpr int FEAT_PARTIAL = F_F32;                                  // raw fp32 partial sums (split-K)
constexpr int FEAT_DENSE = F_SPLIT | F_F32;                          // linear layers: split planes and / or fp32
constexpr int FEAT_LSTM = F_LSTM | F_GATHER;                         // LSTM gate GEMM with the fused cell (beam search)
constexpr int FEAT_DECODE = F_LSTM | F_GATHER | F_ARGMAX | F_F32;    // tile decoder (lm_decode_kernel)

// Epilogue of ONE output tile (mt, nt) by the 8 epilogue warps (threads 128..383) of a CTA.
//
// Hierarchical accumulation.  The tensor core adds every MMA result into the fp32 TMEM accumulator with TRUNCATION (round toward
// zero), so a tile accumulated in one go drifts by ~0.5 ulp per MMA: measured 1.2e-4 relative on fc6 (K = 25088, 4704 MMAs per
// output) -- two orders of magnitude worse than an fp32 sgemm.  The MMA warp therefore closes an accumulator every
// pb.kb_per_group k-blocks (8-12 MMAs), alternating between the two TMEM stages, and the epilogue warps add the finished
// partial sums into registers with round-to-nearest (run[][] below) while the next group is being multiplied.  The drift is
// then bounded by the MMAs of ONE group (<= 4e-7), i.e. fp32-sgemm class, at the cost of one tcgen05.ld sweep per group.
// gi = running count of accumulator groups of this CTA (selects the TMEM stage and its barrier parity); advanced here.
// tfull0 / tempty0: shared addresses of the two tmem_full / tmem_empty barriers.
// Split-plane outputs leave the SM through the TMA store engine.  An epilogue thread owns one ROW of the tile, so a direct store is 16 bytes
// per lane with the row pitch between lanes: 32 half-filled sectors per instruction.  Measured with the stores predicated off (same math):
// conv trunk 2.93 -> 2.36 ms, fc6 1.68 -> 1.42 ms per 4-image step -- the stores, not the arithmetic or the accumulator sweeps, were the cost
// of the output phase.  With TMAS each epilogue warp writes its 32 rows x 32 columns x {hi, lo} into a private 4 KB shared-memory slot
// (SWIZZLE_64B rows, conflict-free STS.128) and one lane hands the slot to the TMA unit (box {32 cols, 32 rows | 16 w x 2 h, planes});
// image borders, ragged N and row tails are clipped by the tensor map.  slot0 = shared address of the 8 slots, omap = output map.
constexpr uint32_t EPI_SLOT_BYTES = 4096;
constexpr uint32_t EPI_STAGING_BYTES = EPI_WARPS * EPI_SLOT_BYTES;
struct OutStage {
  const CUtensorMap* map;
  uint32_t slot0;
};

template <int BLOCK_N, bool CONV, bool POOL, int CG, bool WIDE, int FEAT, bool TMAS = false>
__device__ __forceinline__ void epilogue_tile(const Problem& pb, const GemmEpilogue& ep, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0,
                                              float* bias_smem, int mt, int nt, int num_groups, int& gi, int tile_parity, uint32_t cta_rank, bool is_leader,
                                              OutStage ost = OutStage{nullptr, 0u}) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto tfull_bar = [&](int s) { return tfull0 + 8u * s; };
  auto tempty_bar = [&](int s) { return tempty0 + 8u * s; };
  const int q = warp & 3;                 // TMEM lane quarter this warp may access
  const int chunk_half = (warp - EPI_WARP0) >> 2; // epilogue warps 0-3 drain the first half of the accumulator columns, 4-7 the second
  const int row_in_tile = q * 32 + lane;
  const int n0 = nt * BLOCK_N;
  constexpr int ACC_COLS = WIDE ? 2 * BLOCK_N : BLOCK_N;
  constexpr int CHUNKS_PER_WARP = (BLOCK_N / 32) / (EPI_WARPS / 4);
  const int c_first = chunk_half * CHUNKS_PER_WARP;

  // ---- partial sums of every accumulator group of this tile -> registers (RN adds) ----
  float run[CHUNKS_PER_WARP][32];
  // decoder gate tiles: the input half of the pre-activations comes from a table row per output row (token); it is requested
  // when the first group arrives (the token is final by then) and added after the last one
  float4 gath[CHUNKS_PER_WARP][8];
  const bool use_gather = (FEAT & F_GATHER) && !CONV && WIDE && ep.gather_tab != nullptr;      // only the wide (<= 128-column) tiles carry the table gather
  // Stage this tile's bias slice in shared memory (one coalesced load per thread, issued BEFORE the wait for the first accumulator
  // group so that its L2 latency overlaps the MMAs of this tile) -- per-element global loads in the column loop would serialise on L2
  // latency.  Two buffers used alternately by consecutive tiles, so one named barrier per tile suffices.
  // (the 256-column GEMM configuration has room for one buffer only: an extra barrier keeps the previous tile's readers ahead of the refill)
  constexpr int BIAS_BUFS = (BLOCK_N == 256 && !CONV) ? 1 : 2;
  float* sbias = bias_smem + (BIAS_BUFS == 2 ? (tile_parity & 1) * BLOCK_N : 0);
  if (BIAS_BUFS == 1 && ep.bias && !use_gather) asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
  if (ep.bias && !use_gather) {
    for (int j = threadIdx.x - EPI_THREAD0; j < BLOCK_N; j += 32 * EPI_WARPS) sbias[j] = (n0 + j < pb.N) ? __ldg(ep.bias + n0 + j) : 0.f;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
  }
#pragma unroll 1
  for (int g = 0; g < num_groups; ++g, ++gi) {
    const int acc = gi & 1;
    mbar_wait(tfull_bar(acc), (uint32_t)((gi >> 1) & 1));
    tcgen05_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS;
    // two TMEM loads in flight per wait (the drain latency, not its bandwidth, is what can stall the MMA warp)
    constexpr int STEP = WIDE ? 1 : (CHUNKS_PER_WARP >= 2 ? 2 : 1);
#pragma unroll
    for (int cc = 0; cc < CHUNKS_PER_WARP; cc += STEP) {
      uint32_t raw[32], raw2[32];
      tmem_ld_32x32(t_row + (c_first + cc) * 32, raw);
      // WIDE: columns [BLOCK_N, 2*BLOCK_N) hold the small terms (A_hi.B_lo + A_lo.B_hi) of the same outputs; else: the next chunk
      if (WIDE) tmem_ld_32x32(t_row + BLOCK_N + (c_first + cc) * 32, raw2);
      else if (STEP == 2) tmem_ld_32x32(t_row + (c_first + cc + 1) * 32, raw2);
      tmem_ld_wait();
      if (WIDE) {
#pragma unroll
        for (int j = 0; j < 32; ++j) raw[j] = __float_as_uint(__uint_as_float(raw[j]) + __uint_as_float(raw2[j]));
      }
      if (g == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) run[cc][j] = __uint_as_float(raw[j]);
        if (!WIDE && STEP == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) run[cc + STEP - 1][j] = __uint_as_float(raw2[j]);
        }
        if (use_gather) {
          const int c = c_first + cc;
          const int m = (mt * CG + (int)cta_rank) * BLOCK_M + row_in_tile;
          const int tok = (m < pb.M) ? __ldcg(ep.gather_idx + m) : 1;           // written by another SM during this kernel: read at L2
          const float4* gp = reinterpret_cast<const float4*>(ep.gather_tab + (size_t)(tok - 1) * ep.gather_ld + n0 + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) gath[cc][j] = __ldg(gp + j);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) run[cc][j] += __uint_as_float(raw[j]);
        if (!WIDE && STEP == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) run[cc + STEP - 1][j] += __uint_as_float(raw2[j]);
        }
      }
    }
    // TMEM reads of this accumulator stage are complete: hand it back to the MMA warp
    tcgen05_fence_before();
    if (CG == 2 && !is_leader) mbar_arrive_remote(tempty_bar(acc), 0);
    else mbar_arrive(tempty_bar(acc));
  }

  // ---- which output row does this thread own? ----
  long long out_row = -1;     // row index into the (rows, ld) output matrices, -1 = masked
  bool in_bounds = false;     // accumulator row corresponds to a real element
  int img = 0, h0 = 0, w0 = 0;          // convolution: image and top-left pixel of this CTA's 8 x 16 patch
  if (CONV) {
    const int per_img = pb.tiles_h * pb.tiles_w;
    img = mt / per_img;
    const int r = mt % per_img;
    h0 = ((r / pb.tiles_w) * CG + (int)cta_rank) * TILE_H; w0 = (r % pb.tiles_w) * TILE_W;
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
  uint32_t absmax_bits = 0;           // range guard of the split-fp16 planes: largest |x| written (as IEEE bits)

#pragma unroll
  for (int cc = 0; cc < CHUNKS_PER_WARP; ++cc) {
    const int c = c_first + cc;
    const int nc = n0 + c * 32;
    if (nc >= pb.N) continue;           // whole chunk outside N (uniform across the CTA)
    float v[32];
    // acc_scale is a power of two, so run * acc_scale is exact and the fused form rounds exactly like (run * acc_scale) + bias
  #pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (use_gather) b4 = gath[cc][j4];
      else if (ep.bias) b4 = *reinterpret_cast<const float4*>(sbias + c * 32 + 4 * j4);
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
  #pragma unroll
      for (int u = 0; u < 4; ++u) {
        float x = fmaf(run[cc][4 * j4 + u], ep.acc_scale, bb[u]);
        if (ep.relu) x = fmaxf(x, 0.f);
        v[4 * j4 + u] = x;
        absmax_bits = max(absmax_bits, __float_as_uint(x) & 0x7fffffffu);      // |x| as ordered bits: inf and NaN sort above every finite value
      }
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
    if (TMAS && (FEAT & F_SPLIT) && ep.out_hi) {
      // ---- split planes through the TMA store engine (warp-uniform: every lane takes part, the tensor map clips what is outside) ----
      uint32_t hh[16], ll[16];
  #pragma unroll
      for (int j = 0; j < 16; ++j) {
        __half h0, l0, h1, l1;
        split_f32(v[2 * j], h0, l0);
        split_f32(v[2 * j + 1], h1, l1);
        hh[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        ll[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
      }
      const uint32_t slot = ost.slot0 + (uint32_t)(warp - EPI_WARP0) * EPI_SLOT_BYTES;
      // box rows: 32 accumulator rows (GEMM: 32 matrix rows; conv: 2 image rows x 16 pixels) or, pooled, the 8 pixels of one pooled row
      constexpr int BOX_ROWS = POOL ? 8 : 32;
      const bool writer = POOL ? ((lane & 1) == 0 && lane < 16) : true;       // pooled maxima live in the lanes of even w, even h
      const int brow = POOL ? (lane >> 1) : lane;
      if (lane == 0) bulk_wait_read0();                 // the previous store of this warp has finished READING the slot
      __syncwarp();
      if (writer) {
  #pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t a = slot + (uint32_t)brow * 64u + (uint32_t)((j ^ ((brow >> 1) & 3)) << 4);    // SWIZZLE_64B: 16-byte chunk ^ (row/2 mod 4)
          st_shared_v4(a, hh[4 * j], hh[4 * j + 1], hh[4 * j + 2], hh[4 * j + 3]);
          st_shared_v4(a + BOX_ROWS * 64u, ll[4 * j], ll[4 * j + 1], ll[4 * j + 2], ll[4 * j + 3]);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (CONV) {
          if (POOL) tma_store_5d(ost.map, slot, ep.out_col0 + nc, w0 >> 1, (h0 >> 1) + q, img, 0);
          else tma_store_5d(ost.map, slot, ep.out_col0 + nc, w0, h0 + 2 * q, img, 0);
        } else {
          tma_store_3d(ost.map, slot, ep.out_col0 + nc, (mt * CG + (int)cta_rank) * BLOCK_M + q * 32, 0);
        }
        bulk_commit();
      }
    }
    if ((FEAT & F_LSTM) && ep.lstm_c) {
      // torch-rnn nn.LSTM cell on interleaved gate columns: 8 hidden units per 32-column chunk
      if (out_row >= 0) {
        const int j0 = nc >> 2;
        float* cp = ep.lstm_c + (size_t)out_row * ep.lstm_H + j0;
        float cprev[8];
        if (ep.lstm_first) {
  #pragma unroll
          for (int u = 0; u < 8; ++u) cprev[u] = 0.f;
        } else {
          // read at L2: in the persistent decoder the previous step's cell of this tile was written by another SM
        const float4 c0 = __ldcg(reinterpret_cast<const float4*>(cp)), c1 = __ldcg(reinterpret_cast<const float4*>(cp) + 1);
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
      if (!TMAS && (FEAT & F_SPLIT) && ep.out_hi) {
        // N is a multiple of 8 for split outputs (checked on the host): 16-byte stores, predicated per 8-column group
        __half* ph = ep.out_hi + (size_t)out_row * ep.ldo + ep.out_col0 + nc;
        __half* pl = ep.out_lo + (size_t)out_row * ep.ldo + ep.out_col0 + nc;
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
          if (full || nc + 8 * j + 8 <= pb.N) {
            reinterpret_cast<uint4*>(ph)[j] = make_uint4(hh[4 * j], hh[4 * j + 1], hh[4 * j + 2], hh[4 * j + 3]);
            reinterpret_cast<uint4*>(pl)[j] = make_uint4(ll[4 * j], ll[4 * j + 1], ll[4 * j + 2], ll[4 * j + 3]);
          }
        }
      }
      if ((FEAT & F_F32) && ep.out_f32) {
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
      if ((FEAT & F_ARGMAX) && ep.argmax) {
  #pragma unroll
        for (int j = 0; j < 32; ++j)
          if (nc + j < pb.N && v[j] > best_v) { best_v = v[j]; best_c = nc + j; }
      }
    }
  }
  if ((FEAT & F_ARGMAX) && ep.argmax && out_row >= 0 && best_c != 0x7fffffff)
    atomicMax(ep.argmax + out_row, pack_argmax_tc(best_v, best_c));
  const bool out_of_range = absmax_bits >= 0x477fe000u;      // |x| >= 65504 (fp16 max), inf or NaN
  if ((FEAT & F_SPLIT) && ep.overflow && ep.out_hi && out_of_range && (POOL ? in_bounds : out_row >= 0)) atomicOr(ep.overflow, 1u);
}

// Epilogue warps of the GEMM / convolution kernels: drain the accumulator groups of every tile this CTA owns.
template <int BLOCK_N, bool CONV, bool POOL, int CG, bool WIDE, int FEAT, bool TMAS = false>
__device__ __forceinline__ void epilogue_warps(const Problem& pb, const GemmEpilogue& ep, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0,
                                               float* bias_smem, int worker, int num_workers, uint32_t cta_rank, bool is_leader,
                                               OutStage ost = OutStage{nullptr, 0u}) {
  const int num_groups = (pb.num_k_blocks + pb.kb_per_group - 1) / pb.kb_per_group;
  int gi = 0, iter = 0;
  for (int tile = pb.tile_begin + worker; tile < pb.tile_end; tile += num_workers, ++iter) {
    int mt, nt;
    if (pb.m_fastest) { mt = tile % pb.m_tiles; nt = tile / pb.m_tiles; }
    else { nt = tile % pb.n_tiles; mt = tile / pb.n_tiles; }
    epilogue_tile<BLOCK_N, CONV, POOL, CG, WIDE, FEAT, TMAS>(pb, ep, tmem_base, tfull0, tempty0, bias_smem, mt, nt, num_groups, gi, iter, cta_rank, is_leader, ost);
  }
  if (TMAS && (threadIdx.x & 31) == 0) bulk_wait0();        // the staging slots must outlive the last bulk stores
}

static int encode_map(dc_ctx* ctx, CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int swizzle_bytes = 128) {
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  const PFN_encodeTiled encode = (PFN_encodeTiled)dc_launch_state(ctx)->encode_tiled;
  DC_REQUIRE(ctx, encode != nullptr, DC_E_CUDA, "cuTensorMapEncodeTiled unavailable (tc_init not run)");
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                             : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : (swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE)),
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DC_REQUIRE(ctx, r == CUDA_SUCCESS, DC_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
  return DC_OK;
}


}  // namespace tc
