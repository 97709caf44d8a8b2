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
// Kernel structure (one persistent CTA per SM, 384 threads = 3 warpgroups, see tc_common.cuh):
//   warp 0   : TMA producer -- cp.async.bulk.tensor loads of the A and B stage tiles (both planes in one
//              box) into 128B-swizzled shared memory, signalled through full[] mbarriers;
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit releases the smem stage
//              (empty[]) and publishes a finished accumulator GROUP (tmem_full[]) every kb_per_group k-blocks
//              (hierarchical accumulation: the tensor core accumulates with truncation, tc_common.cuh);
//   warps 4-11: epilogue -- tcgen05.ld sweeps of the accumulator groups (two warps per TMEM lane quarter, half of
//              the columns each) summed in registers, then bias + ReLU (+ fused 2x2 ceil-mode max-pool through warp
//              shuffles), re-split to fp16 planes, staged per warp in shared memory and written with TMA tensor
//              stores (fp32 / partial-sum outputs: direct stores).
//   The accumulator is double buffered in TMEM so the sweeps of group i overlap the MMAs of group i+1.
//   Instantiations: epilogue feature mask FEAT (which output paths are compiled in) and TAIL (K-split items of the
//   leftover tiles of a mostly idle last round, see k_tc_gemm / k_tc_conv3x3).
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
#include "tc_common.cuh"

namespace tc {

template <int BLOCK_N, bool CONV, bool POOL, int CG, int FEAT, bool TAIL = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_out,
               Problem pb, GemmEpilogue ep) {
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
  // [operand ring][8 epilogue staging slots of 4 KB][barriers][bias]
  const uint32_t staging = smem_base + STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_aligned + (size_t)STAGES * C::STAGE_BYTES + C::STAGING_BYTES);
  // bars: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then tmem base address holder
  const uint32_t bar_base = staging + C::STAGING_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* bias_smem = reinterpret_cast<float*>(smem_aligned + (size_t)STAGES * C::STAGE_BYTES + C::STAGING_BYTES + 256);   // [2][BLOCK_N] ([1] at 256 columns)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_out);
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

  // Work items.  Normal launch: the tiles [tile_begin, tile_end), whole.  TAIL instantiation (k_tc_gemm: the leftover tiles of a mostly idle last
  // round): every tile in pb.tail_split K-ranges, item = (tile - tile_begin) * tail_split + part; the epilogue writes raw partial sums to the item's
  // 256 x 256 block of ep.out_f32 and gemm_tail_finalize_kernel adds the ranges.
  const int num_items = TAIL ? (pb.tile_end - pb.tile_begin) * pb.tail_split : (pb.tile_end - pb.tile_begin);
  auto item_decode = [&](int item, int& tile, int& kb0, int& kb1) {
    if (TAIL) {
      const int part = item % pb.tail_split;
      tile = pb.tile_begin + item / pb.tail_split;
      kb0 = (pb.num_k_blocks * part) / pb.tail_split;
      kb1 = (pb.num_k_blocks * (part + 1)) / pb.tail_split;
    } else {
      tile = pb.tile_begin + item; kb0 = 0; kb1 = pb.num_k_blocks;
    }
  };

  auto tile_coords = [&](int tile, int& mt, int& nt) {
    if (pb.m_fastest) { mt = tile % pb.m_tiles; nt = tile / pb.m_tiles; }
    else { nt = tile % pb.n_tiles; mt = tile / pb.n_tiles; }
  };

  if (warp < EPI_WARP0) {
    regs_shrink_control<(BLOCK_N == 256 && CG == 2) ? REGS_CONTROL_LONGK : REGS_CONTROL>();
  }
  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int item = worker; item < num_items; item += num_workers) {
        int tile, kb_begin, kb_end;
        item_decode(item, tile, kb_begin, kb_end);
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
        for (int kb = kb_begin; kb < kb_end; ++kb) {
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
      int gi = 0;                                     // accumulator groups issued so far (TMEM stage = gi & 1)
      const int G = pb.kb_per_group;
      // This loop is executed by ONE thread and paces the tensor pipe: no integer division, no modulo (an ncu source profile showed
      // the issuer ~80 % busy on address arithmetic when the group position was computed as kb % G).
      for (int item = worker; item < num_items; item += num_workers) {
        int tile, kb_begin, kb_end;
        item_decode(item, tile, kb_begin, kb_end);
        uint32_t d_tmem = 0;
        int acc = 0;
        int kb_in_group = 0;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          const bool group_start = kb_in_group == 0;
          if (group_start) {
            // a fresh accumulator every G k-blocks: the epilogue warps sum the groups in registers (see epilogue_tile)
            acc = gi & 1;
            mbar_wait(tempty_bar(acc), (uint32_t)(((gi >> 1) & 1) ^ 1));
            tcgen05_fence_after();
            d_tmem = tmem_base + acc * C::ACC_COLS;
          }
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t a_hi = smem_base + stage * C::STAGE_BYTES;
          const uint32_t a_lo = a_hi + A_PLANE_BYTES;
          const uint32_t b_hi = a_hi + 2 * A_PLANE_BYTES;
          const uint32_t b_lo = b_hi + C::B_PLANE_BYTES;
#pragma unroll
          for (int pass = 0; pass < (C::WIDE ? 1 : 3); ++pass)
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint32_t koff = k * UMMA_K * 2;   // bytes along K inside the swizzle atom
            const uint64_t da_hi = make_kmajor_sw128_desc(a_hi + koff);
            const uint64_t da_lo = make_kmajor_sw128_desc(a_lo + koff);
            const uint64_t db_hi = make_kmajor_sw128_desc(b_hi + koff);
            const uint64_t db_lo = make_kmajor_sw128_desc(b_lo + koff);
            const uint32_t accum = !(group_start && k == 0);
            if (C::WIDE) {
              // D[:, 0:BN] += A_hi.B_hi and D[:, BN:2BN] += A_hi.B_lo in ONE N = 2*BN instruction (B planes are adjacent), then
              // D[:, 0:BN] += A_lo.B_hi.  Every MMA reads its 128 x 32 B A slice from shared memory (~90 clk, measured); two
              // instructions instead of three per K step is what lifts the narrow tiles.
              // The accumulator's columns [BN, 2BN) collect ALL small terms (A_hi.B_lo and A_lo.B_hi), the columns [0, BN) only A_hi.B_hi:
              // every accumulation truncates at the ulp of the running sum, so the small terms must not be added onto the big one.
              umma_f16(d_tmem, da_hi, db_hi, idesc_wide, accum);
              umma_f16(d_tmem + BLOCK_N, da_lo, db_hi, idesc, 1);
            } else {
              // three passes over the four K steps of the block: both small terms first (while a fresh accumulator is still small their
              // truncation error is negligible), the A_hi.B_hi products last
              const uint64_t da = pass == 0 ? da_lo : da_hi, db = pass == 1 ? db_lo : db_hi;
              const uint32_t acc_flag = !(group_start && pass == 0 && k == 0);
              if (CG == 2) umma2_f16(d_tmem, da, db, idesc, acc_flag);
              else umma_f16(d_tmem, da, db, idesc, acc_flag);
            }
          }
          const bool group_end = kb_in_group == G - 1 || kb == kb_end - 1;
          if (CG == 2) {
            umma2_commit_mc(empty_bar(stage));                              // frees this stage in BOTH CTAs
            if (group_end) umma2_commit_mc(tfull_bar(acc));                 // partial accumulators of both CTAs complete
          } else {
            umma_commit(empty_bar(stage));                       // smem stage free once these MMAs retire
            if (group_end) umma_commit(tfull_bar(acc));          // partial accumulator complete
          }
          if (group_end) { ++gi; kb_in_group = 0; } else ++kb_in_group;
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // =============================== epilogue warps ===============================
    regs_grow_epilogue<(BLOCK_N == 256 && CG == 2) ? REGS_EPILOGUE_LONGK : REGS_EPILOGUE>();
    if (!TAIL) {
      epilogue_warps<BLOCK_N, CONV, POOL, CG, C::WIDE, FEAT, C::TMAS>(pb, ep, tmem_base, tfull_bar(0), tempty_bar(0), bias_smem, worker, num_workers, cta_rank,
                                                                      is_leader, OutStage{&tmap_out, staging});
    } else {
      const int G = pb.kb_per_group;
      int gi = 0, iter = 0;
      for (int item = worker; item < num_items; item += num_workers, ++iter) {
        int tile, kb_begin, kb_end;
        item_decode(item, tile, kb_begin, kb_end);
        GemmEpilogue e2 = ep;                                   // raw scaled partial sums into this item's 256 x 256 block (rows 0..255 of a 256-wide matrix)
        e2.out_f32 = ep.out_f32 + (size_t)item * (size_t)(CG * BLOCK_M) * BLOCK_N;
        epilogue_tile<BLOCK_N, CONV, POOL, CG, C::WIDE, FEAT, false>(pb, e2, tmem_base, tfull_bar(0), tempty_bar(0), bias_smem, 0, 0, (kb_end - kb_begin + G - 1) / G,
                                                                     gi, iter, cta_rank, is_leader);
      }
    }
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
  // split-plane outputs through TMA stores where 32 KB of staging fit beside the rings (a B stage less); the 64-column configuration keeps the
  // whole conv1_2 filter bank resident and the 1-CTA 256-column one needs 2 x 64 KB of B: both keep direct stores
  static constexpr bool TMAS = (B_STAGE_BYTES == 32768);
  static constexpr int SB = (B_STAGE_BYTES <= 16384) ? 9 : (B_STAGE_BYTES <= 32768 ? 3 : 2);
  static constexpr bool WIDE = (CG == 1 && BLOCK_N <= 128);
  static constexpr int ACC_COLS = WIDE ? 2 * BLOCK_N : BLOCK_N;
  static constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
  static constexpr uint32_t RING_BYTES = AH_STAGES * AH_STAGE_BYTES + SB * B_STAGE_BYTES;
  static constexpr uint32_t STAGING_BYTES = TMAS ? 8 * 4096 : 0;
  static constexpr size_t SMEM_BYTES = (size_t)RING_BYTES + STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 2 * BLOCK_N * 4 /*bias*/;
};

template <int BLOCK_N, bool POOL, int CG, int FEAT, bool TAIL = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_conv_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_out,
                    Problem pb, GemmEpilogue ep) {
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
  // [A ring][B ring][epilogue staging slots (TMAS)][barriers][bias]
  // bars: a_full[2], a_empty[2], b_full[SB], b_empty[SB], tmem_full[2], tmem_empty[2], tmem base holder
  const uint32_t staging = smem_base + C::RING_BYTES;
  const uint32_t bar_base = staging + C::STAGING_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_aligned + C::RING_BYTES + C::STAGING_BYTES);
  auto afull_bar = [&](int s) { return bar_base + 8u * s; };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (AH_STAGES + s); };
  auto bfull_bar = [&](int s) { return bar_base + 8u * (2 * AH_STAGES + s); };
  auto bempty_bar = [&](int s) { return bar_base + 8u * (2 * AH_STAGES + SB + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * AH_STAGES + 2 * SB + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * AH_STAGES + 2 * SB + 2 + s); };
  static_assert((2 * AH_STAGES + 2 * SB + 4) * 8 + 8 <= 256, "barrier block overflow");
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * AH_STAGES + 2 * SB + 4);
  float* bias_smem = reinterpret_cast<float*>(smem_aligned + C::RING_BYTES + C::STAGING_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_out);
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

  // Work item = (output tile, K split).  split_k > 1 (layers whose tile count cannot fill the GPU: conv5_x, the RPN 3x3): the
  // (tap column, channel block) groups of a tile are divided among split_k items; each writes its raw fp32 partial sums to its own
  // plane of a workspace, and a small finalize kernel adds the planes in a fixed order (+ bias, ReLU, fp16 split).
  // The launch covers the tiles [tile_begin, tile_end).  TAIL instantiation (k_tc_conv3x3: the leftover tiles of a mostly idle last round, each in
  // split_k K-ranges): the items write raw partial sums to compact 256 x 256 blocks, conv_tail_finalize_kernel maps them back to pixels.
  const int S = pb.split_k;
  const int num_tiles = (pb.tile_end - pb.tile_begin) * S;
  const int groups_all = 3 * pb.kb_per_tap;                             // (s, channel block) groups per output tile
  auto group_begin = [&](int sp) { return S == 1 ? sp * groups_all : (groups_all * sp) / S; };       // 32-bit; S == 1 needs no division
  // whole filter bank of the (single) N tile fits the B ring: load it once, keep it
  const bool b_resident = (pb.n_tiles == 1) && (pb.num_k_blocks == SB) && S == 1;

  if (warp < EPI_WARP0) {
    regs_shrink_control<REGS_CONTROL>();
  }
  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (elect_one()) {
      int sa = 0; uint32_t pa = 0;
      int sb = 0; uint32_t pbp = 0;
      bool first = true;
      for (int item = worker; item < num_tiles; item += num_workers, first = false) {
        const int sp = S == 1 ? 0 : item % S, tile = pb.tile_begin + (S == 1 ? item : item / S);
        const int nt = tile % pb.n_tiles, mt = tile / pb.n_tiles;
        const int n0 = nt * BLOCK_N + (int)cta_rank * C::B_ROWS;
        const int per_img = pb.tiles_h * pb.tiles_w;
        const int img = mt / per_img;
        const int rr = mt % per_img;
        const int h0 = ((rr / pb.tiles_w) * CG + (int)cta_rank) * TILE_H;
        const int w0 = (rr % pb.tiles_w) * TILE_W;
        const int g_last = group_begin(sp + 1);
        for (int g = group_begin(sp); g < g_last; ++g) {
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
      int gi = 0;                                     // accumulator groups issued so far (TMEM stage = gi & 1)
      const int G = pb.kb_per_group;
      bool first_tile = true;
      for (int item = worker; item < num_tiles; item += num_workers, first_tile = false) {
        const int sp = S == 1 ? 0 : item % S;
        const int g_begin = group_begin(sp), g_end = group_begin(sp + 1);
        const int nkb = (g_end - g_begin) * 3;        // k-blocks of this item
        uint32_t d_tmem = 0;
        int acc = 0;
        int kb = 0;                                   // (group, vertical tap) counter = k-block index inside the item
        int kb_in_group = 0;                          // position inside the accumulator group (no division in this single-thread loop)
        int gsz = pb.kb_first, grp = 0;               // the first two groups of a tile may be longer (pb.kb_first), see k_tc_conv3x3
        for (int g = g_begin; g < g_end; ++g) {
          mbar_wait(afull_bar(sa), pa);
          const uint32_t a_base = a_ring + sa * AH_STAGE_BYTES;
#pragma unroll 1
          for (int r = 0; r < 3; ++r, ++kb) {
            const bool group_start = kb_in_group == 0;
            if (group_start) {
              acc = gi & 1;
              mbar_wait(tempty_bar(acc), (uint32_t)(((gi >> 1) & 1) ^ 1));
              d_tmem = tmem_base + acc * C::ACC_COLS;
            }
            if (!(b_resident && !first_tile)) mbar_wait(bfull_bar(sb), pbp);
            tcgen05_fence_after();
            const uint32_t a_hi = a_base + r * (TILE_W * BLOCK_K * 2);      // 16 rows = 2048 B down the halo box
            const uint32_t a_lo = a_hi + AH_PLANE_BYTES;
            const uint32_t b_hi = b_ring + sb * C::B_STAGE_BYTES;
            const uint32_t b_lo = b_hi + C::B_PLANE_BYTES;
#pragma unroll
            for (int pass = 0; pass < (C::WIDE ? 1 : 3); ++pass)
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint32_t koff = k * UMMA_K * 2;
              const uint64_t da_hi = make_kmajor_sw128_desc(a_hi + koff);
              const uint64_t da_lo = make_kmajor_sw128_desc(a_lo + koff);
              const uint64_t db_hi = make_kmajor_sw128_desc(b_hi + koff);
              const uint64_t db_lo = make_kmajor_sw128_desc(b_lo + koff);
              const uint32_t accum = !(group_start && k == 0);
              if (C::WIDE) {
                umma_f16(d_tmem, da_hi, db_hi, idesc_wide, accum);          // big terms -> columns [0, BN), A_hi.B_lo -> [BN, 2BN)
                umma_f16(d_tmem + BLOCK_N, da_lo, db_hi, idesc, 1);         // A_lo.B_hi -> the small half as well (see tc_gemm_kernel)
              } else {
                const uint64_t da = pass == 0 ? da_lo : da_hi, db = pass == 1 ? db_lo : db_hi;
                const uint32_t acc_flag = !(group_start && pass == 0 && k == 0);
                if (CG == 2) umma2_f16(d_tmem, da, db, idesc, acc_flag);
                else umma_f16(d_tmem, da, db, idesc, acc_flag);
              }
            }
            const bool group_end = kb_in_group == gsz - 1 || kb == nkb - 1;
            if (CG == 2) {
              if (!b_resident) umma2_commit_mc(bempty_bar(sb));
              if (r == 2) umma2_commit_mc(aempty_bar(sa));
              if (group_end) umma2_commit_mc(tfull_bar(acc));
            } else {
              if (!b_resident) umma_commit(bempty_bar(sb));
              if (r == 2) umma_commit(aempty_bar(sa));
              if (group_end) umma_commit(tfull_bar(acc));
            }
            if (group_end) { ++gi; kb_in_group = 0; if (++grp == 2) gsz = G; } else ++kb_in_group;
            if (++sb == SB) { sb = 0; pbp ^= 1; }
          }
          if (++sa == AH_STAGES) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // =============================== epilogue warps ===============================
    regs_grow_epilogue<REGS_EPILOGUE>();
    const int G = pb.kb_per_group;
    int gi = 0, iter = 0;
    for (int item = worker; item < num_tiles; item += num_workers, ++iter) {
      const int sp = S == 1 ? 0 : item % S, tile = pb.tile_begin + (S == 1 ? item : item / S);
      const int nt = tile % pb.n_tiles, mt = tile / pb.n_tiles;
      const int nkb = (group_begin(sp + 1) - group_begin(sp)) * 3;
      GemmEpilogue e2 = ep;
      if (TAIL) e2.out_f32 = ep.out_f32 + (size_t)item * (size_t)(CG * BLOCK_M) * BLOCK_N;     // this item's 256 x 256 block
      else if (S > 1) e2.out_f32 = ep.out_f32 + (size_t)sp * pb.split_stride;                   // this split's partial-sum plane
      int ngroups = 0, rem = nkb;
      for (int z = 0; z < 2 && rem > 0; ++z) { rem -= min(pb.kb_first, rem); ++ngroups; }
      ngroups += (rem + G - 1) / G;
      if (TAIL)      // rows 0..255 of a 256-wide fp32 matrix: the GEMM form of the epilogue (no pixel mapping, no pooling)
        epilogue_tile<BLOCK_N, false, false, CG, C::WIDE, FEAT, false>(pb, e2, tmem_base, tfull_bar(0), tempty_bar(0), bias_smem, 0, 0, ngroups, gi, iter, cta_rank,
                                                                       is_leader);
      else
        epilogue_tile<BLOCK_N, true, POOL, CG, C::WIDE, FEAT, C::TMAS>(pb, e2, tmem_base, tfull_bar(0), tempty_bar(0), bias_smem, mt, nt, ngroups, gi, iter,
                                                                 cta_rank, is_leader, OutStage{&tmap_out, staging});
    }
    if (C::TMAS && (threadIdx.x & 31) == 0) bulk_wait0();        // the staging slots must outlive the last bulk stores
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
template <typename Kernel>
static int launch_kernel(dc_ctx* ctx, Kernel kernel, size_t smem_bytes, int cg, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo,
                         const Problem& pb, const GemmEpilogue& ep, cudaStream_t st) {
  {
    cudaError_t e = dc_ensure_smem(ctx, kernel, smem_bytes);
    if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "tc_gemm smem attribute: %s", cudaGetErrorString(e));
  }
  const int tiles = (pb.tile_end - pb.tile_begin) * (pb.split_k > 1 ? pb.split_k : 1) * (pb.tail_split > 1 ? pb.tail_split : 1);
  const int slots = dc_launch_state(ctx)->num_sms / cg;                        // CTAs (CG=1) or CTA pairs (CG=2) resident at once
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, ma, mb, mo, pb, ep);
  if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "tc_gemm launch: %s", cudaGetErrorString(e));
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

template <int BLOCK_N, int CG>
static int launch(dc_ctx* ctx, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const Problem& pb, const GemmEpilogue& ep, cudaStream_t st) {
  // linear layers write split planes and / or fp32; the LSTM gate GEMM with the fused cell is its own instantiation (epilogue feature mask)
  if (ep.lstm_c || ep.gather_tab)
    return launch_kernel(ctx, tc_gemm_kernel<BLOCK_N, false, false, CG, FEAT_LSTM>, Cfg<BLOCK_N, CG>::SMEM_BYTES, CG, ma, mb, mo, pb, ep, st);
  return launch_kernel(ctx, tc_gemm_kernel<BLOCK_N, false, false, CG, FEAT_DENSE>, Cfg<BLOCK_N, CG>::SMEM_BYTES, CG, ma, mb, mo, pb, ep, st);
}

template <int BLOCK_N, bool POOL, int CG, int FEAT = FEAT_CONV>
static int launch_halo(dc_ctx* ctx, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const Problem& pb, const GemmEpilogue& ep,
                       cudaStream_t st) {
  return launch_kernel(ctx, tc_conv_halo_kernel<BLOCK_N, POOL, CG, FEAT>, HaloCfg<BLOCK_N, CG>::SMEM_BYTES, CG, ma, mb, mo, pb, ep, st);
}

// Tile shape by an MMA-issue model (ncu, r06): a tcgen05.mma of this kernel cannot retire faster than ~90 clk because it
// streams its 128 x 32 B A slice from shared memory, and it takes N/2 clk on the tensor pipe (M = 128 per CTA).  Per K = 16
// step a CTA spends
//   BLOCK_N =  64 (wide): max(64,90) + max(32,90)  = 180 clk for 128 x  64 outputs
//   BLOCK_N = 128 (wide): max(128,90) + max(64,90) = 218 clk for 128 x 128 outputs
//   BLOCK_N = 256       : 3 x 128                  = 384 clk for 128 x 256 outputs   (cta_group::2: same per CTA, half the B fill)
// and a launch costs rounds x that, rounds = ceil(tiles / resident CTAs).
// k-blocks (of 64) per TMEM accumulator group for a tile width, see epilogue_tile and LaunchState::group_kb
static int group_kblocks(dc_ctx* ctx, int bn) {
  const LaunchState* ls = dc_launch_state(ctx);
  const int g = ls->group_kb[bn == 64 ? 0 : (bn == 128 ? 1 : 2)];
  return g < 1 ? 1 : g;
}
struct TileCfg { int bn; int cg; };
static TileCfg pick_tile(int num_sms, int N, int m_tiles128, bool lstm) {
  TileCfg best{64, 1};
  double best_cost = 1e30;
  auto consider = [&](int bn, int cg, double kcyc) {
    if (bn > 64 && N <= bn / 2) return;                    // mostly-empty tile
    if (cg == 2 && (lstm || N <= 128 || m_tiles128 < 2)) return;
    const long tiles = (long)DC_CEIL_DIV(m_tiles128, cg) * DC_CEIL_DIV(N, bn);
    const long rounds = DC_CEIL_DIV(tiles, (long)(num_sms / cg));
    const double cost = (double)rounds * kcyc;
    if (cost < best_cost) { best_cost = cost; best = TileCfg{bn, cg}; }
  };
  consider(64, 1, 180.0);
  consider(128, 1, 218.0);
  consider(256, 1, 384.0);
  consider(256, 2, 384.0 * 0.97);
  return best;
}

// out = relu(sum_s ws[s] + bias) as split-fp16 planes; the S partial planes are added in index order (deterministic)
__global__ void splitk_finalize_kernel(const float* __restrict__ ws, int S, size_t stride, size_t rows, int N, const float* __restrict__ bias, int relu,
                                       __half* __restrict__ out_hi, __half* __restrict__ out_lo, unsigned int* __restrict__ overflow) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const size_t total8 = rows * (size_t)N / 8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total8; i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = i * 8;
    const int n = (int)(e % N);
    float v[8];
    {
      const float4 a = __ldcg(reinterpret_cast<const float4*>(ws + e)), b = __ldcg(reinterpret_cast<const float4*>(ws + e) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    for (int s2 = 1; s2 < S; ++s2) {
      const float4 a = __ldcg(reinterpret_cast<const float4*>(ws + s2 * stride + e)), b = __ldcg(reinterpret_cast<const float4*>(ws + s2 * stride + e) + 1);
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    uint32_t hh[4], ll[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x0 = v[2 * j], x1 = v[2 * j + 1];
      if (bias) { x0 += __ldg(bias + n + 2 * j); x1 += __ldg(bias + n + 2 * j + 1); }
      if (relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
      if (overflow && (!(fabsf(x0) < 65504.f) || !(fabsf(x1) < 65504.f))) atomicOr(overflow, 1u);
      __half h0, l0, h1, l1;
      split_f32(x0, h0, l0);
      split_f32(x1, h1, l1);
      hh[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      ll[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    *reinterpret_cast<uint4*>(out_hi + e) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
    *reinterpret_cast<uint4*>(out_lo + e) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
  }
}

// Tail tiles of the pair-tile GEMM: out = epilogue(sum of the K-range partial sums); the ranges are added in index order (deterministic)
__global__ void __launch_bounds__(256)
gemm_tail_finalize_kernel(const float* __restrict__ ws, int tile_begin, int n_tail, int S, int m_tiles, int n_tiles, int m_fastest, int M, int N,
                          GemmEpilogue ep) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  constexpr int TR = 2 * BLOCK_M, TC = 256;
  const size_t per_tile = (size_t)TR * TC;
  const size_t total8 = (size_t)n_tail * per_tile / 8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total8; i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = i * 8;
    const int j = (int)(e / per_tile);
    const int r = (int)((e % per_tile) / TC), c = (int)(e % TC);
    const int tile = tile_begin + j;
    const int mt = m_fastest ? tile % m_tiles : tile / n_tiles, nt = m_fastest ? tile / m_tiles : tile % n_tiles;
    const int m = mt * TR + r, n = nt * TC + c;
    if (m >= M || n >= N) continue;
    float v[8];
    const float* p0 = ws + ((size_t)j * S) * per_tile + (size_t)r * TC + c;
    {
      const float4 a = __ldcg(reinterpret_cast<const float4*>(p0)), b = __ldcg(reinterpret_cast<const float4*>(p0) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    for (int s2 = 1; s2 < S; ++s2) {
      const float4 a = __ldcg(reinterpret_cast<const float4*>(p0 + s2 * per_tile)), b = __ldcg(reinterpret_cast<const float4*>(p0 + s2 * per_tile) + 1);
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    bool oor = false;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float x = v[q];
      if (ep.bias && n + q < N) x += __ldg(ep.bias + n + q);
      if (ep.relu) x = fmaxf(x, 0.f);
      v[q] = x;
      oor |= !(fabsf(x) < 65504.f);
    }
    if (ep.out_f32) {
      float* pf = ep.out_f32 + (size_t)m * ep.ldf + n;
#pragma unroll
      for (int q = 0; q < 8; ++q) if (n + q < N) pf[q] = v[q];
    }
    if (ep.out_hi) {            // N is a multiple of 8 for split outputs
      uint32_t hh[4], ll[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        __half h0, l0, h1, l1;
        split_f32(v[2 * q], h0, l0);
        split_f32(v[2 * q + 1], h1, l1);
        hh[q] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        ll[q] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
      }
      const size_t o = (size_t)m * ep.ldo + ep.out_col0 + n;
      *reinterpret_cast<uint4*>(ep.out_hi + o) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
      *reinterpret_cast<uint4*>(ep.out_lo + o) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
      if (ep.overflow && oor) atomicOr(ep.overflow, 1u);
    }
  }
}

// Tail tiles of a pair-tile convolution: partial blocks [(tile - tile_begin) * S + part][256 rows = 2 x (8 x 16 pixels)][256 channels] -> bias, ReLU,
// optional 2x2 ceil-mode max-pool, split NHWC planes.  One thread = one output pixel (pooled: one pooled pixel) x 8 channels.
__global__ void __launch_bounds__(256)
conv_tail_finalize_kernel(const float* __restrict__ ws, int tile_begin, int n_tail, int S, int n_tiles, int tiles_h, int tiles_w, int H, int W, int Cout,
                          int pool, const float* __restrict__ bias, int relu, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                          unsigned int* __restrict__ overflow) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  constexpr int TC = 256;
  const size_t per_tile = (size_t)(2 * BLOCK_M) * TC;
  const int OH = pool ? (H + 1) / 2 : H, OW = pool ? (W + 1) / 2 : W;
  const int px_per_tile = pool ? 64 : 256;                     // 16 x 16 pixels, pooled 8 x 8
  const size_t total = (size_t)n_tail * px_per_tile * (TC / 8);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % (TC / 8)) * 8;
    const int px = (int)((i / (TC / 8)) % px_per_tile);
    const int j = (int)(i / ((size_t)(TC / 8) * px_per_tile));
    const int tile = tile_begin + j;
    const int nt = tile % n_tiles, mt = tile / n_tiles;
    const int per_img = tiles_h * tiles_w;
    const int img = mt / per_img, rem = mt % per_img;
    const int h0 = (rem / tiles_w) * 2 * TILE_H, w0 = (rem % tiles_w) * TILE_W;       // top-left pixel of the 16 x 16 pair tile
    const int n = nt * TC + c8;
    if (n >= Cout) continue;
    float best[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) best[q] = 0.f;                 // pooled outputs follow a ReLU (>= 0): 0 is the identity of max
    const int oy = pool ? (px / 8) : (px / 16), ox = pool ? (px % 8) : (px % 16);
    const int oh = (pool ? h0 / 2 : h0) + oy, ow = (pool ? w0 / 2 : w0) + ox;
    if (oh >= OH || ow >= OW) continue;
    const int taps = pool ? 4 : 1;
    for (int t = 0; t < taps; ++t) {
      const int ly = pool ? 2 * oy + (t >> 1) : oy, lx = pool ? 2 * ox + (t & 1) : ox;      // pixel inside the 16 x 16 tile
      if (h0 + ly >= H || w0 + lx >= W) continue;
      const int r = (ly / TILE_H) * BLOCK_M + (ly % TILE_H) * TILE_W + lx;                  // row of the partial block: CTA rank * 128 + patch row
      const float* p0 = ws + ((size_t)j * S) * per_tile + (size_t)r * TC + c8;
      float v[8];
      {
        const float4 a = __ldcg(reinterpret_cast<const float4*>(p0)), b = __ldcg(reinterpret_cast<const float4*>(p0) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      }
      for (int s2 = 1; s2 < S; ++s2) {
        const float4 a = __ldcg(reinterpret_cast<const float4*>(p0 + s2 * per_tile)), b = __ldcg(reinterpret_cast<const float4*>(p0 + s2 * per_tile) + 1);
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float x = v[q] + (bias ? __ldg(bias + n + q) : 0.f);
        if (relu) x = fmaxf(x, 0.f);
        best[q] = (pool && t > 0) ? fmaxf(best[q], x) : ((pool && t == 0) ? fmaxf(x, 0.f) : x);
      }
    }
    uint32_t hh[4], ll[4];
    bool oor = false;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      __half a0, b0, a1, b1;
      split_f32(best[2 * q], a0, b0);
      split_f32(best[2 * q + 1], a1, b1);
      oor |= !(fabsf(best[2 * q]) < 65504.f) || !(fabsf(best[2 * q + 1]) < 65504.f);
      hh[q] = (uint32_t)__half_as_ushort(a0) | ((uint32_t)__half_as_ushort(a1) << 16);
      ll[q] = (uint32_t)__half_as_ushort(b0) | ((uint32_t)__half_as_ushort(b1) << 16);
    }
    const size_t o = (((size_t)img * OH + oh) * OW + ow) * Cout + n;
    *reinterpret_cast<uint4*>(out_hi + o) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
    *reinterpret_cast<uint4*>(out_lo + o) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
    if (overflow && oor) atomicOr(overflow, 1u);
  }
}

}  // namespace tc

int tc_init(dc_ctx* ctx) {
  LaunchState* ls = dc_launch_state(ctx);
  if (!ls->encode_tiled) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
      return dc_set_error(ctx, DC_E_CUDA, "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
    ls->encode_tiled = fn;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp prop;
  DC_CHECK_CUDA(ctx, cudaGetDeviceProperties(&prop, dev));
  DC_REQUIRE(ctx, prop.major == 10, DC_E_UNSUPPORTED, "the tcgen05 kernels need an sm_100 device, found sm_%d%d", prop.major, prop.minor);
  ls->num_sms = prop.multiProcessorCount;
  return DC_OK;
}

int k_tc_gemm(dc_ctx* ctx, const __half* a_planes, size_t a_plane_elems, int lda, const __half* b_planes, size_t b_plane_elems,
              int ldb, int M, int N, int K, GemmEpilogue ep, cudaStream_t st) {
  using namespace tc;
  if (M <= 0 || N <= 0) return DC_OK;
  DC_REQUIRE(ctx, K > 0 && K % BLOCK_K == 0, DC_E_UNSUPPORTED, "tc_gemm: K=%d must be a positive multiple of 64", K);
  DC_REQUIRE(ctx, lda % 8 == 0 && ldb % 8 == 0 && a_plane_elems % 8 == 0 && b_plane_elems % 8 == 0, DC_E_UNSUPPORTED,
             "tc_gemm: leading dimensions / plane sizes must be multiples of 8 elements");
  DC_REQUIRE(ctx, !ep.out_hi || (ep.ldo % 8 == 0 && ep.out_col0 % 8 == 0 && (N % 8 == 0 || ep.lstm_c)), DC_E_UNSUPPORTED,
             "tc_gemm: split outputs need N, ld and the column offset to be multiples of 8");
  DC_REQUIRE(ctx, !ep.argmax, DC_E_UNSUPPORTED, "tc_gemm: the arg-max epilogue belongs to the decoder kernels");
  DC_REQUIRE(ctx, !ep.lstm_c || (ep.out_hi && N == 4 * ep.lstm_H && ep.lstm_H % 8 == 0), DC_E_UNSUPPORTED, "tc_gemm: LSTM epilogue shape");
  Problem pb{};
  pb.M = M; pb.N = N; pb.K = K;
  pb.num_k_blocks = K / BLOCK_K;
  pb.split_k = 1;
  const int m_tiles128 = DC_CEIL_DIV(M, BLOCK_M);
  TileCfg tc_ = pick_tile(dc_launch_state(ctx)->num_sms, N, m_tiles128, ep.lstm_c != nullptr);
  if (ep.gather_tab && tc_.bn == 256) tc_ = TileCfg{128, 1};          // the table-gather epilogue exists for the wide tiles only
  const int bn = tc_.bn, cg = tc_.cg;
  pb.kb_per_group = group_kblocks(ctx, bn);
  // short reductions (the RPN 1x1 heads, K = 256; the LM encoder / gate table, K = 512): an accumulator sweep per k-block costs nothing
  // measurable there and these outputs feed discrete decisions (NMS order, arg-max), so take the most accurate setting
  if (pb.num_k_blocks <= 8) pb.kb_per_group = 1;
  pb.m_tiles = DC_CEIL_DIV(m_tiles128, cg);
  pb.n_tiles = DC_CEIL_DIV(N, bn);
  pb.m_fastest = pb.m_tiles <= pb.n_tiles;
  pb.tile_begin = 0; pb.tile_end = pb.m_tiles * pb.n_tiles; pb.tail_split = 1;
  if (ep.out_hi && !ep.overflow) ep.overflow = dc_overflow_flag(ctx);
  // Tail split (fc6 / fc7 at batch 4: 256 pair tiles on 74 SM pairs = three full rounds + a round of 34 tiles).  The full rounds run as one
  // launch; the leftover tiles run as a second launch of the TAIL instantiation, each cut into S K-ranges so that this last round fills the
  // machine and lasts 1/S of a round, followed by a small kernel that adds the ranges and applies the real epilogue.
  int n_tail = 0, tail_S = 1;
  float* tail_ws = nullptr;
#ifndef DC_DISABLE_TAIL_SPLIT
  if (cg == 2 && !ep.lstm_c && !ep.gather_tab && !ep.argmax && pb.num_k_blocks >= 128) {   // long K only: at K = 4096 (fc7) the two extra launches cost more than the idle half round
    const int slots = dc_launch_state(ctx)->num_sms / 2;
    const int tiles = pb.m_tiles * pb.n_tiles;
    const int left = tiles % slots;
    size_t ws_elems = 0;
    float* ws = dc_splitk_workspace(ctx, &ws_elems);
    if (tiles > slots && left > 0 && 2 * left <= slots && ws) {
      int S = std::min(4, slots / left);
      while (S >= 2 && (size_t)left * S * (2 * BLOCK_M) * 256 > ws_elems) --S;
      if (S >= 2) { n_tail = left; tail_S = S; tail_ws = ws; pb.tile_end = tiles - left; }
    }
  }
#endif
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
  // output map of the TMA-store epilogue: (planes, M rows, out_col0 + N columns) fp16, box {32 columns, 32 rows, 2 planes}, SWIZZLE_64B
  CUtensorMap mo = ma;
  if (ep.out_hi && !ep.lstm_c) {
    DC_REQUIRE(ctx, ep.out_lo > ep.out_hi && ((size_t)(ep.out_lo - ep.out_hi) * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(ep.out_hi) & 15) == 0,
               DC_E_UNSUPPORTED, "tc_gemm: split output planes must be 16-byte aligned, lo plane after the hi plane");
    const uint64_t dims[3] = {(uint64_t)(ep.out_col0 + N), (uint64_t)M, 2};
    const uint64_t str[2] = {(uint64_t)ep.ldo * 2, (uint64_t)(ep.out_lo - ep.out_hi) * 2};
    const uint32_t box[3] = {32, 32, 2};
    int rc = encode_map(ctx, &mo, ep.out_hi, 3, dims, str, box, 64);
    if (rc) return rc;
  }
  if (cg == 2) {
    int rc = launch<256, 2>(ctx, ma, mb, mo, pb, ep, st);
    if (rc || n_tail == 0) return rc;
    Problem pt = pb;
    pt.tile_begin = pb.tile_end; pt.tile_end = pb.tile_end + n_tail; pt.tail_split = tail_S;
    GemmEpilogue et;                                        // raw partial sums, scaled
    et.acc_scale = ep.acc_scale; et.out_f32 = tail_ws; et.ldf = 256;
    if ((rc = launch_kernel(ctx, tc_gemm_kernel<256, false, false, 2, FEAT_PARTIAL, true>, Cfg<256, 2>::SMEM_BYTES, 2, ma, mb, mo, pt, et, st))) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)std::min<size_t>((size_t)dc_launch_state(ctx)->num_sms * 4, DC_CEIL_DIV((size_t)n_tail * 2 * BLOCK_M * 256 / 8, (size_t)256)));
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tail_finalize_kernel, (const float*)tail_ws, pt.tile_begin, n_tail, tail_S, pb.m_tiles, pb.n_tiles, pb.m_fastest,
                                       M, N, ep);
    if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "gemm_tail_finalize launch: %s", cudaGetErrorString(e));
    DC_CHECK_LAUNCH(ctx);
    return DC_OK;
  }
  if (bn == 64) return launch<64, 1>(ctx, ma, mb, mo, pb, ep, st);
  if (bn == 256) return launch<256, 1>(ctx, ma, mb, mo, pb, ep, st);
  return launch<128, 1>(ctx, ma, mb, mo, pb, ep, st);
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
  TileCfg tc_ = pick_tile(dc_launch_state(ctx)->num_sms, Cout, m_tiles128, false);
  // Small feature maps (conv5_x, RPN 3x3 at batch 1-2) have too few 256 x 256 pair tiles to fill the GPU and were run as 128 x 64 tiles at
  // ~51 % tensor-pipe activity.  Split K instead: every pair tile is divided into S items over the (tap column, channel block) groups.
  int split = 1;
  size_t ws_elems = 0;
  float* ws = dc_splitk_workspace(ctx, &ws_elems);
#ifdef DC_DISABLE_SPLITK
  ws = nullptr;
#endif
  if (!fuse_pool && Cout >= 256 && Cout % 8 == 0 && ws) {
    const int slots = dc_launch_state(ctx)->num_sms / 2;
    const long pair_tiles = (long)N * DC_CEIL_DIV(H, 2 * TILE_H) * pb.tiles_w * DC_CEIL_DIV(Cout, 256);
    const int groups_all = 3 * pb.kb_per_tap;
    if (pair_tiles * 2 <= slots) {
      int S = (int)std::min<long>(std::min<long>(slots / pair_tiles, groups_all / 2), 8);
      while (S >= 2 && (size_t)S * N * H * W * Cout > ws_elems) --S;
      if (S >= 2) { split = S; tc_ = TileCfg{256, 2}; }
    }
  }
  const int bn = tc_.bn, cg = tc_.cg;
  pb.kb_per_group = group_kblocks(ctx, bn);
  // The output phase of tile i keeps the epilogue warps away from the accumulator sweeps of tile i+1, and the issuer can run only two groups
  // ahead (two TMEM stages): with one-k-block groups it stalls at every tile boundary.  The first two groups of a tile are therefore longer; their
  // partial sums are the smallest of the tile, so the extra truncating accumulations cost the least accuracy there.  Measured (A/B, batch 4):
  // 2 k-blocks: conv3_x / conv4_x -3..5 %, trunk features 1.57e-6 -> 1.73e-6 from the fp64 oracle (fp32 oracle: 8.5e-7); 3 k-blocks: no
  // further speed-up and 2.67e-6 (over the 2x-noise budget of tests/test_gpu_fullsize.py).
#ifndef DC_FIRST_GROUP_KB
#define DC_FIRST_GROUP_KB 2
#endif
  pb.kb_first = std::max(pb.kb_per_group, (bn == 256 && pb.kb_per_group < DC_FIRST_GROUP_KB) ? DC_FIRST_GROUP_KB : pb.kb_per_group);
  pb.split_k = split;
  pb.split_stride = (size_t)N * H * W * Cout;
  pb.tiles_h = DC_CEIL_DIV(H, TILE_H * cg);          // cg == 2: a pair tile covers 16 image rows (two stacked 8x16 patches)
  pb.m_tiles = N * pb.tiles_h * pb.tiles_w;
  pb.M = pb.m_tiles * BLOCK_M * cg;
  pb.OH = fuse_pool ? (H + 1) / 2 : H;
  pb.OW = fuse_pool ? (W + 1) / 2 : W;
  pb.n_tiles = DC_CEIL_DIV(Cout, bn);
  pb.m_fastest = 0;   // n fastest: CTAs sharing an activation patch run together, the patch stays in L2
  pb.tile_begin = 0; pb.tile_end = pb.m_tiles * pb.n_tiles; pb.tail_split = 1;
  GemmEpilogue ep;
  ep.acc_scale = acc_scale;
  if (split > 1) {             // raw partial sums; bias / ReLU / fp16 split happen in splitk_finalize_kernel
    ep.out_f32 = ws;
    ep.ldf = Cout;
  } else {
    ep.bias = bias;
    ep.relu = relu;
    ep.out_hi = y_hi;
    ep.out_lo = y_lo;
    ep.ldo = Cout;
    ep.overflow = dc_overflow_flag(ctx);
  }
  CUtensorMap ma, mb;
  {
    const uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N, 2};
    const uint64_t str[4] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2, (uint64_t)x_plane_elems * 2};
    const uint32_t box[5] = {BLOCK_K, TILE_W, (uint32_t)HALO_H, 1, 2};
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
  // output map of the TMA-store epilogue: (planes, N, OH, OW, Cout) fp16 NHWC; box = the 2 image rows x 16 pixels (pooled: 1 x 8) x 32 channels
  // x 2 planes one epilogue warp produces per 32-column chunk, SWIZZLE_64B
  CUtensorMap mo = ma;
  if (split == 1) {
    DC_REQUIRE(ctx, y_lo > y_hi && ((size_t)(y_lo - y_hi) * 2) % 16 == 0, DC_E_UNSUPPORTED, "tc_conv3x3: lo plane must follow the hi plane (16-byte aligned)");
    const uint64_t dims[5] = {(uint64_t)Cout, (uint64_t)pb.OW, (uint64_t)pb.OH, (uint64_t)N, 2};
    const uint64_t str[4] = {(uint64_t)Cout * 2, (uint64_t)pb.OW * Cout * 2, (uint64_t)pb.OH * pb.OW * Cout * 2, (uint64_t)(y_lo - y_hi) * 2};
    const uint32_t box[5] = {32, fuse_pool ? 8u : 16u, fuse_pool ? 1u : 2u, 1, 2};
    int rc = encode_map(ctx, &mo, y_hi, 5, dims, str, box, 64);
    if (rc) return rc;
  }
  // Tail split (conv4_x at batch 4: 240 pair tiles on 74 SM pairs = three full rounds + a round of 18): the full rounds run as one launch, the
  // leftover tiles as a second launch of the TAIL instantiation, each cut into S ranges of its (tap column, channel block) groups, so that the last
  // round fills the machine and lasts 1/S of a round; conv_tail_finalize_kernel adds the ranges (+ bias, ReLU, pool, fp16 split).
#ifndef DC_DISABLE_TAIL_SPLIT
  if (cg == 2 && split == 1) {
    const int slots = dc_launch_state(ctx)->num_sms / 2;
    const int tiles = pb.m_tiles * pb.n_tiles;
    const int left = tiles % slots;
    const int groups_all = 3 * pb.kb_per_tap;
    // only when the tail is at most a quarter round (S = 4): with half a round left (conv3_x at batch 4, S = 2) the two extra launches and the
    // finalize pass cost what the split saves (A/B: conv4_x -8 %, conv3_x +-0)
    if (tiles > slots && left > 0 && 4 * left <= slots && ws && groups_all >= 12) {
      int S = std::min(std::min(4, slots / left), groups_all / 3);
      while (S >= 2 && (size_t)left * S * (2 * BLOCK_M) * 256 > ws_elems) --S;
      if (S >= 2) {
        pb.tile_end = tiles - left;
        int rc = fuse_pool ? launch_halo<256, true, 2>(ctx, ma, mb, mo, pb, ep, st) : launch_halo<256, false, 2>(ctx, ma, mb, mo, pb, ep, st);
        if (rc) return rc;
        Problem pt = pb;
        pt.tile_begin = pb.tile_end; pt.tile_end = tiles; pt.split_k = S;
        GemmEpilogue et;
        et.acc_scale = acc_scale; et.out_f32 = ws; et.ldf = 256;
        if ((rc = launch_kernel(ctx, tc_conv_halo_kernel<256, false, 2, FEAT_PARTIAL, true>, HaloCfg<256, 2>::SMEM_BYTES, 2, ma, mb, mo, pt, et, st))) return rc;
        const size_t work = (size_t)left * (fuse_pool ? 64 : 256) * 32;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)std::min<size_t>((size_t)dc_launch_state(ctx)->num_sms * 4, DC_CEIL_DIV(work, (size_t)256)));
        cfg.blockDim = dim3(256);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tail_finalize_kernel, (const float*)ws, pt.tile_begin, left, S, pb.n_tiles, pb.tiles_h, pb.tiles_w, H, W, Cout,
                                           fuse_pool ? 1 : 0, bias, relu, y_hi, y_lo, dc_overflow_flag(ctx));
        if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "conv_tail_finalize launch: %s", cudaGetErrorString(e));
        DC_CHECK_LAUNCH(ctx);
        return DC_OK;
      }
    }
  }
#endif
  if (fuse_pool) {
    if (cg == 2) return launch_halo<256, true, 2>(ctx, ma, mb, mo, pb, ep, st);
    if (bn == 64) return launch_halo<64, true, 1>(ctx, ma, mb, mo, pb, ep, st);
    if (bn == 256) return launch_halo<256, true, 1>(ctx, ma, mb, mo, pb, ep, st);
    return launch_halo<128, true, 1>(ctx, ma, mb, mo, pb, ep, st);
  }
  if (split > 1) {
    int rc = launch_halo<256, false, 2, FEAT_PARTIAL>(ctx, ma, mb, mo, pb, ep, st);
    if (rc) return rc;
    const size_t rows = (size_t)N * H * W;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)std::min<size_t>((size_t)dc_launch_state(ctx)->num_sms * 4, DC_CEIL_DIV(rows * Cout / 8, (size_t)256)));
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, splitk_finalize_kernel, (const float*)ws, split, pb.split_stride, rows, Cout, bias, relu, y_hi, y_lo, dc_overflow_flag(ctx));
    if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "splitk_finalize launch: %s", cudaGetErrorString(e));
    DC_CHECK_LAUNCH(ctx);
    return DC_OK;
  }
  if (cg == 2) return launch_halo<256, false, 2>(ctx, ma, mb, mo, pb, ep, st);
  if (bn == 64) return launch_halo<64, false, 1>(ctx, ma, mb, mo, pb, ep, st);
  if (bn == 256) return launch_halo<256, false, 1>(ctx, ma, mb, mo, pb, ep, st);
  return launch_halo<128, false, 1>(ctx, ma, mb, mo, pb, ep, st);
}
