// LanguageModel:sample (densecap/LanguageModel.lua:293-348), greedy decode, as ONE persistent tcgen05 kernel over the
// rows that survive the final NMS (DenseCapModel.lua:261-275 keeps only those rows of the LM output).
//
// Per token the reference runs: LookupTable -> nn.LSTM step (gates = b + x Wx + h Wh, cell) -> vocabulary Linear -> max.
// The input half b + x Wx depends only on the token, so it is a table (V+2 rows x 4H, built once when the weights are loaded:
// "gate table"); a step is then  gates = table[token] + h Wh  (K = H instead of E + H, and no embedding copy).
// The whole T-step loop is one launch of 148 CTAs that walk a static list of work items
//     Gf(mt,nt)                      first LSTM step on the encoded image, h0 = c0 = 0 (x = encoder output, K = E)
//     for t: { G_t(mt,nt) x16 , V_t(mt,nt) x83 }  per 128-row block mt   (gate GEMM + fused cell ; vocabulary GEMM + arg-max)
// round-robin over the CTAs.  Rows are independent, so dependencies exist only INSIDE a 128-row block:
//     G_t(mt,*) needs token_t of block mt   (flag_x[t][mt], set by whoever finishes V_{t-1}(mt,*) last; implies h_{t-1})
//     V_t(mt,*) needs h_t of block mt       (cnt_gate[t+1][mt] == 16)
// They are tracked with per-block counters in global memory (release/acquire at gpu scope + a generic->async proxy fence,
// because the consumers read the operands with TMA), not with grid-wide barriers: block 0 can be in its vocabulary
// projection while block 3 still runs its gate GEMM.  The CTA whose epilogue completes the LAST vocabulary tile of a block
// turns the packed arg-max cells into tokens (torch.max: first maximal index) and writes them to the outputs -- the
// "lm_token_step" launches of round 1 are gone.
// Every CTA processes its items in list order and the list is a topological order of the dependency graph, so the
// smallest unfinished item can always run: no deadlock as long as all CTAs are resident (grid <= #SMs, 1 CTA/SM).
// All waits are bounded (trap after ~2 s) so that a protocol bug surfaces as a CUDA error, never as a hung GPU.
//
// The MMA / TMA machinery (128x128 tiles, "wide" hi|lo B operand: 2 MMAs per K step, 3-stage smem ring, two TMEM accumulator
// stages, 8 epilogue warps with hierarchical accumulation) is the one of tc_gemm_kernel<128, gemm, 1>; ring and accumulator
// parities simply continue across items.
#include "tc_common.cuh"

namespace tc {

constexpr int DEC_BN = 128;

struct DecodeParams {
  int E, H, V1, T;                     // V1 = V + 1 (logit width)
  const int32_t* m_rows;               // device: number of rows to decode (M')
  __half* h_hi[2];                     // h_t ping-pong buffers, hi / lo planes (rows, H)
  __half* h_lo[2];
  float* cstate;                       // (rows, H)
  const float* lstm_bias;              // (4H) gate-interleaved (first step only; later steps take it from the gate table)
  float lstm_acc_scale;
  const float* gate_table;             // (V+2, 4H) fp32 = b + lookup . Wx, gate-interleaved columns
  int32_t* tokens;                     // (rows) input token of the current step
  const float* vocab_bias;             // (V1)
  float vocab_acc_scale;
  unsigned long long* packed;          // (rows) arg-max cells, zero on entry
  int start_token;                     // V + 1
  uint32_t* cnt_gate;                  // [(T+1)][mt_stride]  index 0 = first step
  uint32_t* cnt_vocab;                 // [T][mt_stride]
  uint32_t* flag_x;                    // [T][mt_stride]
  int mt_stride;
  const int32_t* dst_final;            // (rows) row of fin_seq for compact row r, or -1
  const int32_t* dst_all;              // (rows) row of all_seq, or -1
  int64_t* fin_seq;                    // (., T)
  int64_t* all_seq;                    // (., T)
  float* logits;                       // optional (rows, T, V1)
};

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t atom_add_release_u32(uint32_t* p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
// generic-proxy global writes of this thread -> visible to later async-proxy (TMA) reads, and vice versa
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__device__ __forceinline__ void wait_counter(const uint32_t* p, uint32_t want) {
  if (ld_acquire_u32(p) >= want) return;
  const long long t0 = clock64();
  while (ld_acquire_u32(p) < want) {
    __nanosleep(64);
    if (clock64() - t0 > 4000000000LL) {
      printf("densecap_b200: lm_decode dependency timeout (block %d, counter %p = %u, want %u)\n", blockIdx.x, (const void*)p, ld_acquire_u32(p), want);
      __trap();
    }
  }
}

struct DecItem { int vocab; int t; int mt; int nt; };   // t = -1: first LSTM step

__global__ void __launch_bounds__(NUM_THREADS, 1)
lm_decode_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_h0, const __grid_constant__ CUtensorMap map_h1,
                 const __grid_constant__ CUtensorMap map_lstm, const __grid_constant__ CUtensorMap map_vocab, DecodeParams p) {
  using C = Cfg<DEC_BN, 1>;
  static_assert(C::WIDE, "decoder uses the wide hi|lo B operand");
  constexpr int STAGES = C::STAGES;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* smem_aligned = smem_dyn + (smem_base - smem_u32(smem_dyn));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_aligned + (size_t)STAGES * C::STAGE_BYTES);
  const uint32_t bar_base = smem_base + STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  int* s_flag = reinterpret_cast<int*>(tmem_holder + 1);
  float* bias_smem = reinterpret_cast<float*>(smem_aligned + (size_t)STAGES * C::STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    prefetch_tmap(&map_x); prefetch_tmap(&map_h0); prefetch_tmap(&map_h1); prefetch_tmap(&map_lstm); prefetch_tmap(&map_vocab);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 32 * EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_holder), C::TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int M = __ldcg(p.m_rows);
  const int mT = (M + BLOCK_M - 1) / BLOCK_M;
  const int GN = (4 * p.H) / DEC_BN;
  const int VN = (p.V1 + DEC_BN - 1) / DEC_BN;
  const int per_mt = GN + VN;
  const int first_items = mT * GN;
  const int total_items = first_items + p.T * mT * per_mt;
  auto decode = [&](int g) {
    DecItem it;
    if (g < first_items) { it.vocab = 0; it.t = -1; it.mt = g / GN; it.nt = g % GN; return it; }
    const int g2 = g - first_items;
    it.t = g2 / (mT * per_mt);
    const int r = g2 % (mT * per_mt);
    it.mt = r / per_mt;
    const int q = r % per_mt;
    it.vocab = q >= GN;
    it.nt = it.vocab ? q - GN : q;
    return it;
  };
  auto kblocks = [&](const DecItem& it) { return (!it.vocab && it.t < 0) ? p.E / BLOCK_K : p.H / BLOCK_K; };
  const int worker = blockIdx.x, num_workers = gridDim.x;

  if (warp < EPI_WARP0) {
    regs_shrink_control<REGS_CONTROL>();
  }
  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int g = worker; g < total_items; g += num_workers) {
        const DecItem it = decode(g);
        // h buffers: the first step writes h into buffer 1; step t reads buffer (t odd ? 0 : 1) and writes the other one
        const CUtensorMap* ma;
        const CUtensorMap* mb;
        int kb0;                                   // K offset (in elements) into the B operand
        if (it.vocab) {
          wait_counter(p.cnt_gate + (size_t)(it.t + 1) * p.mt_stride + it.mt, (uint32_t)GN);      // h_t of this row block complete
          ma = (it.t & 1) ? &map_h1 : &map_h0;
          mb = &map_vocab;
          kb0 = 0;
        } else if (it.t < 0) {
          ma = &map_x;                              // encoded image vectors, K = E, Wx rows of the LSTM weight
          mb = &map_lstm;
          kb0 = 0;
        } else {
          if (it.t == 0) wait_counter(p.cnt_gate + it.mt, (uint32_t)GN);                          // first step complete (h in buffer 1)
          else wait_counter(p.flag_x + (size_t)it.t * p.mt_stride + it.mt, 1u);                   // token_t known (implies h_{t-1})
          ma = (it.t & 1) ? &map_h0 : &map_h1;
          mb = &map_lstm;
          kb0 = p.E;                                // Wh rows
        }
        fence_proxy_async_global();
        const int nkb = kblocks(it);
        const int m0 = it.mt * BLOCK_M, n0 = it.nt * DEC_BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t a_dst = smem_base + stage * C::STAGE_BYTES;
          const uint32_t b_dst = a_dst + 2 * A_PLANE_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
          tma_load_3d(a_dst, ma, full_bar(stage), kb * BLOCK_K, m0, 0);
          tma_load_3d(b_dst, mb, full_bar(stage), kb0 + kb * BLOCK_K, n0, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(BLOCK_M, DEC_BN);
      constexpr uint32_t idesc_wide = make_idesc_f16(BLOCK_M, 2 * DEC_BN);
      int stage = 0; uint32_t phase = 0;
      int gi = 0;                                    // accumulator groups issued so far (one per k-block, see epilogue_tile)
      for (int g = worker; g < total_items; g += num_workers) {
        const DecItem it = decode(g);
        const int nkb = kblocks(it);
        for (int kb = 0; kb < nkb; ++kb, ++gi) {
          const int acc = gi & 1;
          mbar_wait(tempty_bar(acc), (uint32_t)(((gi >> 1) & 1) ^ 1));
          const uint32_t d_tmem = tmem_base + acc * C::ACC_COLS;
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t a_hi = smem_base + stage * C::STAGE_BYTES;
          const uint32_t a_lo = a_hi + A_PLANE_BYTES;
          const uint32_t b_hi = a_hi + 2 * A_PLANE_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint32_t koff = k * UMMA_K * 2;
            const uint64_t da_hi = make_kmajor_sw128_desc(a_hi + koff);
            const uint64_t da_lo = make_kmajor_sw128_desc(a_lo + koff);
            const uint64_t db_hi = make_kmajor_sw128_desc(b_hi + koff);
            umma_f16(d_tmem, da_hi, db_hi, idesc_wide, k != 0);            // A_hi.[B_hi | B_lo]
            umma_f16(d_tmem + DEC_BN, da_lo, db_hi, idesc, 1);             // A_lo.B_hi -> the small-term half of the accumulator
          }
          umma_commit(empty_bar(stage));
          umma_commit(tfull_bar(acc));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // =============================== epilogue warps ===============================
    regs_grow_epilogue<REGS_EPILOGUE>();
    const int et = threadIdx.x - EPI_THREAD0;                // 0 .. 255
    int iter = 0, gi = 0;
    for (int g = worker; g < total_items; g += num_workers, ++iter) {
      const DecItem it = decode(g);
      Problem pb{};
      pb.M = M;
      GemmEpilogue ep;
      if (it.vocab) {
        pb.N = p.V1;
        ep.bias = p.vocab_bias;
        ep.acc_scale = p.vocab_acc_scale;
        ep.argmax = p.packed;
        if (p.logits) { ep.out_f32 = p.logits + (size_t)it.t * p.V1; ep.ldf = p.T * p.V1; }
      } else {
        pb.N = 4 * p.H;
        ep.acc_scale = p.lstm_acc_scale;
        if (it.t < 0) ep.bias = p.lstm_bias;
        else { ep.gather_tab = p.gate_table; ep.gather_idx = p.tokens; ep.gather_ld = 4 * p.H; }
        ep.lstm_c = p.cstate;
        ep.lstm_H = p.H;
        ep.lstm_first = it.t < 0;
        const int dst = (it.t < 0) ? 1 : (it.t & 1);
        ep.out_hi = p.h_hi[dst];
        ep.out_lo = p.h_lo[dst];
        ep.ldo = p.H;
        ep.out_col0 = 0;
      }
      epilogue_tile<DEC_BN, false, false, 1, true, FEAT_DECODE>(pb, ep, tmem_base, tfull_bar(0), tempty_bar(0), bias_smem, it.mt, it.nt, kblocks(it), gi, iter,
                                                   0u, true);
      // publish: this thread's global writes (h / c / arg-max cells / logits) -> gpu scope and the async proxy
      fence_proxy_async_global();
      __threadfence();
      asm volatile("bar.sync 2, %0;" ::"n"(32 * EPI_WARPS) : "memory");
      if (!it.vocab) {
        if (et == 0) atom_add_release_u32(p.cnt_gate + (size_t)(it.t + 1) * p.mt_stride + it.mt, 1u);
        continue;
      }
      if (et == 0) {
        const uint32_t old = atom_add_release_u32(p.cnt_vocab + (size_t)it.t * p.mt_stride + it.mt, 1u);
        *s_flag = (old == (uint32_t)(VN - 1));
        __threadfence();
      }
      asm volatile("bar.sync 2, %0;" ::"n"(32 * EPI_WARPS) : "memory");
      if (!*s_flag) { asm volatile("bar.sync 2, %0;" ::"n"(32 * EPI_WARPS) : "memory"); continue; }   // keeps s_flag stable until all read it
      // ---- last vocabulary tile of row block mt at step t: the tokens of step t (= inputs of step t+1) ----
      {
        const bool more = it.t + 1 < p.T;
        const int r = it.mt * BLOCK_M + et;
        if (et < BLOCK_M && r < M) {
          const unsigned long long pk = __ldcg(p.packed + r);
          const uint32_t col = 0xFFFFFFFFu - (uint32_t)(pk & 0xFFFFFFFFull);
          const int tok = (pk == 0ull) ? 1 : (int)min(col, (uint32_t)p.start_token - 1u) + 1;     // all-NaN row -> first index, like torch.max
          const int df = p.dst_final ? __ldg(p.dst_final + r) : -1;
          const int da = p.dst_all ? __ldg(p.dst_all + r) : -1;
          if (df >= 0) p.fin_seq[(size_t)df * p.T + it.t] = tok;
          if (da >= 0) p.all_seq[(size_t)da * p.T + it.t] = tok;
          p.tokens[r] = tok;
          p.packed[r] = 0ull;
        }
        __threadfence();
        asm volatile("bar.sync 2, %0;" ::"n"(32 * EPI_WARPS) : "memory");
        if (et == 0 && more) st_release_u32(p.flag_x + (size_t)(it.t + 1) * p.mt_stride + it.mt, 1u);
      }
    }
  }

  __syncwarp();
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ==================================================================================================================
// Weights-stationary persistent decoder (the default path; the tile decoder above remains for models whose LM weights do not
// fit the shared memory of one GPU -- see k_lm_decode).
//
// The tile decoder re-reads the vocabulary matrix (21 MB as split planes) from L2 for every 128-row block and step: at 740
// rows that is ~300 MB of L2->SM operand traffic per step, and ncu showed the kernel at 12.7 % tensor-pipe activity, bound
// by that traffic.  Here every CTA PAIR (cta_group::2) owns 176 output columns for the whole launch and keeps its slice of the
// weights (88 rows x H x {hi, lo} per CTA = 176 KB at H = 512) resident in shared memory; only the activations h_t
// (rows x H, ~1.5 MB) stream through a small ring every step.  Column owners: pairs [0, nv) = vocabulary rows, pairs
// [nv, nv + ng) = the recurrent LSTM weights Wh (gate-interleaved rows).  Because logits_t and the recurrent half of the next
// pre-activations both read h_t, ONE pass over h per step feeds both:
//     step s = 0      : A = x (encoded image),  gate pairs only (B = Wx)  -> cell with c = 0, + bias            -> h
//     step s = 1      : A = h,                  gate pairs only (B = Wh)  -> cell, + table[START]               -> h
//     step s = 2..T+1 : A = h,  all pairs: logits_{s-2} (+ arg-max) | h.Wh -> token = arg-max -> seq[s-2]; cell, + table[token] -> h
// (the last step has no cell).  The LSTM cell is an element-wise phase between the passes, executed by the epilogue warps of all
// CTAs (one row per CTA and round): gates = table[token] + h.Wh -- the gate table (b + lookup.Wx) removes the x half of the GEMM.
// MMA: M = 256 (two CTAs x 128 rows), N = 176, three kind::f16 instructions per K = 16 step; the A_hi.B_hi products go to one
// TMEM accumulator (columns [0,176)), both small terms to a second one ([176,352)), so the 32 truncating accumulations of a
// K = 512 product never see a small term added onto a large sum and no mid-tile accumulator sweeps are needed.
// A ring: 3 stages of {32 K x 128 rows x 2 planes} = 16 KB (SWIZZLE_64B rows) -- what is left beside the resident weights.
// Dependencies are per 256-row block (release/acquire counters at gpu scope + generic->async proxy fences, as above):
//     pass(s, mb) needs cell(s, mb) of ALL CTAs     (cnt_cell[s][mb] == #CTAs;  s >= 1)
//     cell(s+1, mb) needs pass(s, mb) of all owners (cnt_pass[s][mb] == #CTAs that ran the pass)
// Every role walks (s, mb) in lexicographic order, which is a topological order -> no deadlock while all CTAs are resident
// (grid = 2 (nv + ng) <= #SMs, one CTA per SM).  All waits are bounded.
// ==================================================================================================================
constexpr int WS_NP = 176;                                   // output columns per CTA pair
constexpr int WS_NC = WS_NP / 2;                             // weight rows resident per CTA
#ifndef DC_WS_KA
#define DC_WS_KA 32
#endif
constexpr int WS_KA = DC_WS_KA;                              // K extent of one A stage: 32 (64-byte rows, SWIZZLE_64B) or 16 (32-byte rows, SWIZZLE_32B)
constexpr uint32_t WS_A_PLANE = BLOCK_M * WS_KA * 2;         // 8 KB
constexpr uint32_t WS_A_STAGE = 2 * WS_A_PLANE;              // 16 KB (hi + lo)
// 48 KB of ring beside the resident weights: 3 x 16 KB.  (ncu at batch 4: L2 at 20 % of its peak, tensor pipe 37 % -- the h stream looks bound by
// the bytes in flight against the L2 latency; a finer ring of 6 x 8 KB SWIZZLE_32B stages, -DDC_WS_KA=16, keeps 40 KB instead of 32 KB on the way
// but measured no faster: 0.614 vs 0.600 ms per 4-image step.)
constexpr int WS_STAGES = 49152 / WS_A_STAGE;
constexpr uint32_t WS_B_PLANE = WS_NC * BLOCK_K * 2;         // 11264 B: 88 rows x 128 B (one 64-wide K block of one plane)
constexpr uint32_t WS_B_KB = 2 * WS_B_PLANE;                 // hi + lo
constexpr int WS_MAX_KB = 8;                                 // H <= 512
constexpr uint32_t WS_TMEM_COLS = 512;
constexpr int WS_ROWS = 2 * BLOCK_M;                         // rows per pass tile (pair)
static size_t ws_smem_bytes(int H) { return (size_t)(H / BLOCK_K) * WS_B_KB + (size_t)WS_STAGES * WS_A_STAGE + 1024 + 256 + WS_NP * 4; }

struct WsParams {
  int E, H, V1, T;
  int nv, ng;                          // vocabulary pairs, gate pairs
  const int32_t* m_rows;
  float* gH;                           // (rows, 4H) fp32: recurrent half of the gate pre-activations (acc_scale applied)
  float* cstate;                       // (rows, H)
  __half* h_hi; __half* h_lo;          // (rows, H) split planes: h_t, the A operand of the next pass
  const float* gate_table;             // (V+2, 4H) = b + lookup.Wx, gate-interleaved
  const float* lstm_bias;              // (4H) gate-interleaved (image step only)
  float lstm_acc_scale, vocab_acc_scale;
  const float* vocab_bias;
  unsigned long long* packed;          // (rows) arg-max cells, zero on entry
  int start_token;
  uint32_t* cnt_pass;                  // [T+2][mt_stride]
  uint32_t* cnt_cell;                  // [T+2][mt_stride]
  int mt_stride;
  const int32_t* dst_final; const int32_t* dst_all;
  int64_t* fin_seq; int64_t* all_seq;
  float* logits;                       // optional (rows, T, V1)
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
lm_decode_ws_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_vocab,
                    const __grid_constant__ CUtensorMap map_lstm, WsParams p) {
  const uint32_t cta_rank = cluster_ctarank();
  const bool is_leader = cta_rank == 0;
  const int pair = (int)(blockIdx.x >> 1);
  const bool gate_pair = pair >= p.nv;
  const int col0 = (gate_pair ? pair - p.nv : pair) * WS_NP;          // first output column of this pair (in its own matrix)
  cluster_sync_all();
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* smem_aligned = smem_dyn + (smem_base - smem_u32(smem_dyn));
  const int KB = p.H / BLOCK_K;                  // 64-wide K blocks of the resident weights
  const int KS = p.H / WS_KA;                    // A stages per pass tile (E == H is required)
  const uint32_t b_base = smem_base;
  const uint32_t a_ring = smem_base + (uint32_t)KB * WS_B_KB;
  const uint32_t bar_base = a_ring + WS_STAGES * WS_A_STAGE;
  unsigned char* bar_ptr = smem_aligned + (size_t)KB * WS_B_KB + (size_t)WS_STAGES * WS_A_STAGE;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (WS_STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * WS_STAGES), tempty_bar = tfull_bar + 8u, bfull_bar = tfull_bar + 16u, bready_bar = tfull_bar + 24u,
                 bfree_bar = tfull_bar + 32u;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar_ptr + 8 * (2 * WS_STAGES + 5));
  float* bias_smem = reinterpret_cast<float*>(bar_ptr + 256);          // [WS_NP] vocabulary bias / LSTM bias of this pair's columns

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    prefetch_tmap(&map_x); prefetch_tmap(&map_h); prefetch_tmap(&map_vocab); prefetch_tmap(&map_lstm);
    for (int s = 0; s < WS_STAGES; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 2 * 32 * EPI_WARPS);
    mbar_init(bfull_bar, 1);
    mbar_init(bready_bar, 1);
    mbar_init(bfree_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(smem_u32(tmem_holder), WS_TMEM_COLS);
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int M = __ldcg(p.m_rows);
  const int mT = (M + WS_ROWS - 1) / WS_ROWS;
  const int S_LAST = p.T + 1;                                // last step index
  const int first_step = gate_pair ? 0 : 2;                  // vocabulary pairs have nothing to do for the two state-only steps
  const uint32_t n_ctas = gridDim.x;

  if (warp < EPI_WARP0) regs_shrink_control<REGS_CONTROL>();
  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (elect_one()) {
      const CUtensorMap* mw = gate_pair ? &map_lstm : &map_vocab;
      const int n0 = col0 + (int)cta_rank * WS_NC;
      auto load_weights = [&](int k0) {
        mbar_arrive_expect_tx(bfull_bar, (uint32_t)KB * WS_B_KB);
        for (int kb = 0; kb < KB; ++kb) tma_load_3d(b_base + (uint32_t)kb * WS_B_KB, mw, bfull_bar, k0 + kb * BLOCK_K, n0, 0);
      };
      // resident weight slice: Wx (K offset 0) for the image step of the gate pairs, then Wh (K offset E); vocabulary rows have K offset 0
      load_weights(0);
      mbar_wait(bfull_bar, 0);
      if (!is_leader) mbar_arrive_remote(bready_bar, 0);
      int stage = 0; uint32_t phase = 0;
      for (int s = first_step; s <= S_LAST; ++s) {
        if (gate_pair && s == 1) {
          // swap Wx for Wh once every MMA of step 0 has retired in both CTAs (bfree is committed after the last one)
          mbar_wait(bfree_bar, 0);
          load_weights(p.E);
          mbar_wait(bfull_bar, 1);
          if (!is_leader) mbar_arrive_remote(bready_bar, 0);
        }
        const CUtensorMap* ma = (s == 0) ? &map_x : &map_h;
        for (int mb = 0; mb < mT; ++mb) {
          if (s >= 1) {
            wait_counter(p.cnt_cell + (size_t)s * p.mt_stride + mb, n_ctas);      // h rows of this block written by every CTA
            fence_proxy_async_global();
          }
          const int m0 = mb * WS_ROWS + (int)cta_rank * BLOCK_M;
          for (int ks = 0; ks < KS; ++ks) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            if (is_leader) mbar_arrive_expect_tx(full_bar(stage), 2 * WS_A_STAGE);
            else mbar_arrive_remote(full_bar(stage), 0);
            tma2_load_3d(a_ring + stage * WS_A_STAGE, ma, full_bar(stage), ks * WS_KA, m0, 0);
            if (++stage == WS_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA) ===============================
    if (is_leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_f16(2 * BLOCK_M, WS_NP);
      const uint32_t d_big = tmem_base, d_small = tmem_base + WS_NP;
      mbar_wait(bfull_bar, 0);
      mbar_wait(bready_bar, 0);
      int stage = 0; uint32_t phase = 0, tpar = 0;
      for (int s = first_step; s <= S_LAST; ++s) {
        if (gate_pair && s == 1) { mbar_wait(bfull_bar, 1); mbar_wait(bready_bar, 1); }
        for (int mb = 0; mb < mT; ++mb) {
          mbar_wait(tempty_bar, tpar ^ 1);
          tcgen05_fence_after();
          for (int ks = 0; ks < KS; ++ks) {
            mbar_wait(full_bar(stage), phase);
            tcgen05_fence_after();
            const uint32_t a_hi = a_ring + stage * WS_A_STAGE, a_lo = a_hi + WS_A_PLANE;
            constexpr int SPB = BLOCK_K / WS_KA;               // A stages per 64-wide K block of the resident weights
            const uint32_t b_hi = b_base + (uint32_t)(ks / SPB) * WS_B_KB + (uint32_t)(ks % SPB) * (uint32_t)(WS_KA * 2), b_lo = b_hi + WS_B_PLANE;
#pragma unroll
            for (int kk = 0; kk < WS_KA / UMMA_K; ++kk) {
              const uint32_t koff = kk * UMMA_K * 2;
              const uint64_t da_hi = WS_KA == 32 ? make_kmajor_sw64_desc(a_hi + koff) : make_kmajor_sw32_desc(a_hi + koff);
              const uint64_t da_lo = WS_KA == 32 ? make_kmajor_sw64_desc(a_lo + koff) : make_kmajor_sw32_desc(a_lo + koff);
              const uint64_t db_hi = make_kmajor_sw128_desc(b_hi + koff), db_lo = make_kmajor_sw128_desc(b_lo + koff);
              const uint32_t acc = (ks | kk) != 0;
              umma2_f16(d_small, da_lo, db_hi, idesc, acc);
              umma2_f16(d_small, da_hi, db_lo, idesc, 1);
              umma2_f16(d_big, da_hi, db_hi, idesc, acc);
            }
            umma2_commit_mc(empty_bar(stage));
            if (++stage == WS_STAGES) { stage = 0; phase ^= 1; }
          }
          umma2_commit_mc(tfull_bar);
          tpar ^= 1;
        }
        if (gate_pair && s == 0) umma2_commit_mc(bfree_bar);          // every MMA that reads Wx has retired when this arrives
      }
    }
  } else if (warp >= EPI_WARP0) {
    // =============================== epilogue warps: pass epilogue + LSTM cell / token phase ===============================
    regs_grow_epilogue<REGS_EPILOGUE>();
    const int et = threadIdx.x - EPI_THREAD0;                // 0 .. 255
    const int q = warp & 3, chalf = (warp - EPI_WARP0) >> 2;
    const int H4 = 4 * p.H;
    // bias of this pair's columns (constant for the whole launch)
    for (int j = et; j < WS_NP; j += 32 * EPI_WARPS) {
      const int n = col0 + j;
      bias_smem[j] = gate_pair ? (n < H4 ? __ldg(p.lstm_bias + n) : 0.f) : (n < p.V1 ? __ldg(p.vocab_bias + n) : 0.f);
    }
    asm volatile("bar.sync 2, %0;" ::"n"(32 * EPI_WARPS) : "memory");
    const float scale = gate_pair ? p.lstm_acc_scale : p.vocab_acc_scale;
    const int c_begin = chalf ? 6 : 0, c_end = chalf ? WS_NP / 16 : 6;        // 16-column chunks: warps 4-7 take 96 columns, 8-11 take 80
    uint32_t tpar = 0;
    for (int s = 0; s <= S_LAST; ++s) {
      const bool active = s >= first_step;
      const uint32_t owners = (s < 2) ? (uint32_t)(2 * p.ng) : n_ctas;         // CTAs that run pass s
      for (int mb = 0; mb < mT; ++mb) {
        if (active) {
          mbar_wait(tfull_bar, tpar);
          tcgen05_fence_after();
          const int row = mb * WS_ROWS + (int)cta_rank * BLOCK_M + q * 32 + lane;
          const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
          float best_v = -INFINITY;
          int best_c = 0x7fffffff;
#pragma unroll 1
          for (int c = c_begin; c < c_end; ++c) {
            uint32_t rb[16], rs[16];
            tmem_ld_32x16(t_row + c * 16, rb);
            tmem_ld_32x16(t_row + WS_NP + c * 16, rs);
            tmem_ld_wait();
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (__uint_as_float(rb[j]) + __uint_as_float(rs[j])) * scale;
            const int n = col0 + c * 16;
            if (row < M) {
              if (gate_pair) {
                if (n < H4) {                    // 4H is a multiple of 16
                  float* g = p.gH + (size_t)row * H4 + n;
#pragma unroll
                  for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(g)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
              } else {
                float* lg = p.logits ? p.logits + ((size_t)row * p.T + (s - 2)) * p.V1 + n : nullptr;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  if (n + j < p.V1) {
                    const float x = v[j] + bias_smem[c * 16 + j];
                    if (lg) lg[j] = x;
                    if (x > best_v) { best_v = x; best_c = n + j; }
                  }
                }
              }
            }
          }
          tcgen05_fence_before();
          if (!is_leader) mbar_arrive_remote(tempty_bar, 0);
          else mbar_arrive(tempty_bar);
          tpar ^= 1;
          if (!gate_pair && row < M && best_c != 0x7fffffff) atomicMax(p.packed + row, pack_argmax_tc(best_v, best_c));
          __threadfence();
          asm volatile("bar.sync 2, %0;" ::"n"(32 * EPI_WARPS) : "memory");
          if (et == 0) atom_add_release_u32(p.cnt_pass + (size_t)s * p.mt_stride + mb, 1u);
        }
        // ---- token + LSTM cell of step s for the rows of block mb this CTA owns (one row per round, 256 threads per row) ----
        if (et == 0) wait_counter(p.cnt_pass + (size_t)s * p.mt_stride + mb, owners);
        asm volatile("bar.sync 2, %0;" ::"n"(32 * EPI_WARPS) : "memory");
        const int r_end = min(M, (mb + 1) * WS_ROWS);
        for (int r = mb * WS_ROWS + (int)blockIdx.x; r < r_end; r += (int)n_ctas) {
          int tok = p.start_token;
          if (s >= 2) {
            const unsigned long long pk = __ldcg(p.packed + r);
            const uint32_t col = 0xFFFFFFFFu - (uint32_t)(pk & 0xFFFFFFFFull);
            tok = (pk == 0ull) ? 1 : (int)min(col, (uint32_t)p.start_token - 1u) + 1;     // all-NaN row -> first index, like torch.max
          }
          if (s < S_LAST) {
            for (int u2 = et; u2 < p.H / 2; u2 += 32 * EPI_WARPS) {      // two hidden units (8 interleaved gate columns) per thread
              const float4* gp = reinterpret_cast<const float4*>(p.gH + (size_t)r * H4 + 8 * u2);
              const float4 g0 = __ldcg(gp), g1 = __ldcg(gp + 1);
              float4 t0, t1;
              if (s == 0) {
                t0 = __ldg(reinterpret_cast<const float4*>(p.lstm_bias + 8 * u2));
                t1 = __ldg(reinterpret_cast<const float4*>(p.lstm_bias + 8 * u2) + 1);
              } else {
                const float4* tp = reinterpret_cast<const float4*>(p.gate_table + (size_t)(tok - 1) * H4 + 8 * u2);
                t0 = __ldg(tp); t1 = __ldg(tp + 1);
              }
              float2 cprev = make_float2(0.f, 0.f);
              float2* cp = reinterpret_cast<float2*>(p.cstate + (size_t)r * p.H + 2 * u2);
              if (s > 0) cprev = __ldcg(cp);
              // torch-rnn nn.LSTM cell on interleaved columns (i, f, o, g), same arithmetic as the fused GEMM epilogue (tc_common.cuh)
              const float i0 = lstm_sigmoid(g0.x + t0.x), f0 = lstm_sigmoid(g0.y + t0.y), o0 = lstm_sigmoid(g0.z + t0.z), gg0 = lstm_tanh(g0.w + t0.w);
              const float i1 = lstm_sigmoid(g1.x + t1.x), f1 = lstm_sigmoid(g1.y + t1.y), o1 = lstm_sigmoid(g1.z + t1.z), gg1 = lstm_tanh(g1.w + t1.w);
              const float c0 = f0 * cprev.x + i0 * gg0, c1 = f1 * cprev.y + i1 * gg1;
              *cp = make_float2(c0, c1);
              __half h0h, h0l, h1h, h1l;
              split_f32(o0 * lstm_tanh(c0), h0h, h0l);
              split_f32(o1 * lstm_tanh(c1), h1h, h1l);
              const size_t ho = (size_t)r * p.H + 2 * u2;
              *reinterpret_cast<__half2*>(p.h_hi + ho) = __halves2half2(h0h, h1h);
              *reinterpret_cast<__half2*>(p.h_lo + ho) = __halves2half2(h0l, h1l);
            }
          }
          if (s >= 2 && et == 0) {
            const int df = p.dst_final ? __ldg(p.dst_final + r) : -1;
            const int da = p.dst_all ? __ldg(p.dst_all + r) : -1;
            if (df >= 0) p.fin_seq[(size_t)df * p.T + (s - 2)] = tok;
            if (da >= 0) p.all_seq[(size_t)da * p.T + (s - 2)] = tok;
          }
        }
        if (s < S_LAST) {
          // publish h / c of this CTA's rows: gpu scope + async proxy (the next pass reads h with TMA)
          fence_proxy_async_global();
          __threadfence();
          asm volatile("bar.sync 2, %0;" ::"n"(32 * EPI_WARPS) : "memory");
          if (et == 0) {
            // every thread of the CTA has read its rows' arg-max cells (barrier above): clear them for the next vocabulary pass
            if (s >= 2)
              for (int r = mb * WS_ROWS + (int)blockIdx.x; r < r_end; r += (int)n_ctas) p.packed[r] = 0ull;
            atom_add_release_u32(p.cnt_cell + (size_t)(s + 1) * p.mt_stride + mb, 1u);
          }
        }
      }
    }
  }

  __syncwarp();
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc2(tmem_base, WS_TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Row compaction in front of the decoder: the rows to decode (survivors of the final NMS, or every proposal) of all
// images of the batch become rows [0, M') of the operand buffers.  One block per 8 rows (+ grid-stride), 256 threads.
//   counts (B) rows per image; idx32 (B*P) per-image source row of each kept row (NULL = identity).
//   enc (B*P, E) split planes = ReLU(W_e codes + b) of every proposal (LanguageModel.lua:27-29) -> x half of buffer 0;
//   lookup row of START -> x half of buffer 1 (input of step t = 0).
// Also zeroes the decoder's dependency counters, the arg-max cells and the token outputs.
// ------------------------------------------------------------------------------------------------------------------
struct CompactParams {
  int B, P, E, T;
  const int32_t* counts;
  const int32_t* idx32;
  int survivors;                      // 1: dst_final = b*P + i (fin_seq written by the decoder), 0: dst_final = -1
  const __half* enc_hi; const __half* enc_lo;
  __half* x_hi; __half* x_lo;         // (rows, E) compacted encoder output
  int start_token;
  int32_t* m_rows; int32_t* dst_final; int32_t* dst_all; int32_t* tokens;
  unsigned long long* packed;
  uint32_t* counters; int n_counters;
  int64_t* fin_seq; int64_t* all_seq;   // zeroed when non-null (B*P*T each)
};

__global__ void __launch_bounds__(256)
lm_compact_kernel(CompactParams c) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __shared__ int s_off[65];
  if (threadIdx.x == 0) {
    int o = 0;
    for (int b = 0; b < c.B; ++b) { s_off[b] = o; o += max(0, min(__ldcg(c.counts + b), c.P)); }
    s_off[c.B] = o;
    if (blockIdx.x == 0) *c.m_rows = o;
  }
  __syncthreads();
  const int M = s_off[c.B];
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gthreads = (size_t)gridDim.x * blockDim.x;
  for (size_t i = gtid; i < (size_t)c.n_counters; i += gthreads) c.counters[i] = 0u;
  const size_t nseq = (size_t)c.B * c.P * c.T;
  if (c.fin_seq) for (size_t i = gtid; i < nseq; i += gthreads) c.fin_seq[i] = 0;
  if (c.all_seq) for (size_t i = gtid; i < nseq; i += gthreads) c.all_seq[i] = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + warp; r < M; r += gridDim.x * 8) {
    int b = 0;
    while (b + 1 < c.B && r >= s_off[b + 1]) ++b;
    const int i = r - s_off[b];
    const int local = c.idx32 ? __ldcg(c.idx32 + (size_t)b * c.P + i) : i;
    const int src = b * c.P + local;
    if (lane == 0) {
      c.dst_final[r] = c.survivors ? b * c.P + i : -1;
      c.dst_all[r] = src;
      c.tokens[r] = c.start_token;                        // input of step t = 0 (LanguageModel.lua:318-321)
      c.packed[r] = 0ull;
    }
    for (int e = lane; e < c.E / 8; e += 32) {
      const size_t so = (size_t)src * c.E + (size_t)e * 8, d = (size_t)r * c.E + (size_t)e * 8;
      *reinterpret_cast<uint4*>(c.x_hi + d) = __ldcg(reinterpret_cast<const uint4*>(c.enc_hi + so));
      *reinterpret_cast<uint4*>(c.x_lo + d) = __ldcg(reinterpret_cast<const uint4*>(c.enc_lo + so));
    }
  }
}

}  // namespace tc

size_t lm_decode_counter_words(int T, int rows_cap) {
  const int mt_stride = DC_CEIL_DIV(rows_cap, tc::BLOCK_M) + 1;
  const size_t tile_words = (size_t)(3 * T + 1) * mt_stride;                       // tile decoder
  const size_t ws_words = (size_t)2 * (T + 2) * (DC_CEIL_DIV(rows_cap, tc::WS_ROWS) + 1);   // weights-stationary decoder
  return tile_words > ws_words ? tile_words : ws_words;
}

// Weights-stationary decoder launch; returns DC_OK with *launched = false when the model does not fit (caller falls back to the tile decoder)
static int launch_lm_decode_ws(dc_ctx* ctx, const LmDecodeArgs& a, cudaStream_t st, bool* launched) {
  using namespace tc;
  *launched = false;
  const int V1 = a.V + 1, H4 = 4 * a.H;
  const int nv = DC_CEIL_DIV(V1, WS_NP), ng = DC_CEIL_DIV(H4, WS_NP);
  const int num_sms = dc_launch_state(ctx)->num_sms;
  if (a.E != a.H || a.H % BLOCK_K != 0 || a.H > WS_MAX_KB * BLOCK_K || 2 * (nv + ng) > num_sms) return DC_OK;
  const size_t smem = ws_smem_bytes(a.H);
  const int rows_cap = a.B * a.P;
  const int mt_stride = DC_CEIL_DIV(rows_cap, WS_ROWS) + 1;
  const size_t hbuf = (size_t)rows_cap * a.H;
  __half* h_hi = a.h_planes;
  __half* h_lo = a.h_planes + 2 * hbuf;
  CUtensorMap mx, mh, mv, ml;
  int rc;
  {
    const uint64_t dims[3] = {(uint64_t)a.E, (uint64_t)rows_cap, 2};
    const uint64_t str[2] = {(uint64_t)a.E * 2, (uint64_t)a.x_plane_elems * 2};
    const uint32_t box[3] = {WS_KA, BLOCK_M, 2};
    if ((rc = encode_map(ctx, &mx, a.x_planes, 3, dims, str, box, 2 * WS_KA))) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a.H, (uint64_t)rows_cap, 2};
    const uint64_t str[2] = {(uint64_t)a.H * 2, (uint64_t)(2 * hbuf) * 2};
    const uint32_t box[3] = {WS_KA, BLOCK_M, 2};
    if ((rc = encode_map(ctx, &mh, h_hi, 3, dims, str, box, 2 * WS_KA))) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a.H, (uint64_t)V1, 2};
    const uint64_t str[2] = {(uint64_t)a.H * 2, (uint64_t)a.vocab_plane_elems * 2};
    const uint32_t box[3] = {BLOCK_K, WS_NC, 2};
    if ((rc = encode_map(ctx, &mv, a.vocab_planes, 3, dims, str, box))) return rc;
  }
  {
    const int ld = a.E + a.H;
    const uint64_t dims[3] = {(uint64_t)ld, (uint64_t)H4, 2};
    const uint64_t str[2] = {(uint64_t)ld * 2, (uint64_t)a.lstm_plane_elems * 2};
    const uint32_t box[3] = {BLOCK_K, WS_NC, 2};
    if ((rc = encode_map(ctx, &ml, a.lstm_planes, 3, dims, str, box))) return rc;
  }
  WsParams p{};
  p.E = a.E; p.H = a.H; p.V1 = V1; p.T = a.T;
  p.nv = nv; p.ng = ng;
  p.m_rows = a.m_rows;
  p.gH = a.gate_ws;
  p.cstate = a.cstate;
  p.h_hi = h_hi; p.h_lo = h_lo;
  p.gate_table = a.gate_table; p.lstm_bias = a.lstm_bias;
  p.lstm_acc_scale = a.lstm_acc_scale; p.vocab_acc_scale = a.vocab_acc_scale; p.vocab_bias = a.vocab_bias;
  p.packed = a.packed;
  p.start_token = a.V + 1;
  p.cnt_pass = a.counters;
  p.cnt_cell = a.counters + (size_t)(a.T + 2) * mt_stride;
  p.mt_stride = mt_stride;
  p.dst_final = a.dst_final; p.dst_all = a.dst_all; p.fin_seq = a.fin_seq; p.all_seq = a.all_seq;
  p.logits = a.logits;
  {
    cudaError_t e = dc_ensure_smem(ctx, lm_decode_ws_kernel, smem);
    if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "lm_decode_ws smem attribute: %s", cudaGetErrorString(e));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (nv + ng));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, lm_decode_ws_kernel, mx, mh, mv, ml, p);
  if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "lm_decode_ws launch: %s", cudaGetErrorString(e));
  DC_CHECK_LAUNCH(ctx);
  *launched = true;
  return DC_OK;
}

int k_lm_decode(dc_ctx* ctx, const LmDecodeArgs& a, cudaStream_t st, bool compact_only) {
  using namespace tc;
  DC_REQUIRE(ctx, a.E % BLOCK_K == 0 && a.H % BLOCK_K == 0 && (4 * a.H) % DEC_BN == 0 && a.T >= 1 && a.B >= 1 && a.B <= 64, DC_E_UNSUPPORTED,
             "lm_decode: encoding / rnn size must be multiples of 64 (4H of 128), batch <= 64");
  const int ld = a.E + a.H;
  const int rows_cap = a.B * a.P;
  const int mt_stride = DC_CEIL_DIV(rows_cap, BLOCK_M) + 1;
  const size_t hbuf = (size_t)rows_cap * a.H;                // one plane of one h buffer
  __half* h_hi0 = a.h_planes;
  __half* h_hi1 = a.h_planes + hbuf;
  __half* h_lo0 = a.h_planes + 2 * hbuf;
  __half* h_lo1 = a.h_planes + 3 * hbuf;
  uint32_t* cnt_gate = a.counters;
  uint32_t* cnt_vocab = cnt_gate + (size_t)(a.T + 1) * mt_stride;
  uint32_t* flag_x = cnt_vocab + (size_t)a.T * mt_stride;
  // ---- compaction ----
  {
    CompactParams c{};
    c.B = a.B; c.P = a.P; c.E = a.E; c.T = a.T;
    c.counts = a.counts; c.idx32 = a.idx32; c.survivors = a.survivors;
    c.enc_hi = a.enc_planes; c.enc_lo = a.enc_planes + a.enc_plane_elems;
    c.x_hi = a.x_planes; c.x_lo = a.x_planes + a.x_plane_elems;
    c.start_token = a.V + 1;
    c.m_rows = a.m_rows; c.dst_final = a.dst_final; c.dst_all = a.dst_all; c.tokens = a.tokens;
    c.packed = a.packed;
    c.counters = a.counters; c.n_counters = (int)lm_decode_counter_words(a.T, rows_cap);
    c.fin_seq = a.survivors ? a.fin_seq : nullptr;
    c.all_seq = a.all_seq;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)std::min(DC_CEIL_DIV(rows_cap, 8), 4 * dc_launch_state(ctx)->num_sms));
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, lm_compact_kernel, c);
    if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "lm_compact launch: %s", cudaGetErrorString(e));
    DC_CHECK_LAUNCH(ctx);
  }
  if (compact_only) return DC_OK;
  // ---- persistent decoder: weights-stationary kernel when the LM fits one GPU's shared memory (DenseCap's does), tile kernel otherwise ----
#ifndef DC_DISABLE_LM_WS
  if (a.gate_ws) {
    bool launched = false;
    int rc = launch_lm_decode_ws(ctx, a, st, &launched);
    if (rc || launched) return rc;
  }
#endif
  CUtensorMap mx, mh0, mh1, ml, mv;
  {
    const uint64_t dims[3] = {(uint64_t)a.E, (uint64_t)rows_cap, 2};
    const uint64_t str[2] = {(uint64_t)a.E * 2, (uint64_t)a.x_plane_elems * 2};
    const uint32_t box[3] = {BLOCK_K, BLOCK_M, 2};
    int rc = encode_map(ctx, &mx, a.x_planes, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a.H, (uint64_t)rows_cap, 2};
    const uint64_t str[2] = {(uint64_t)a.H * 2, (uint64_t)(2 * hbuf) * 2};
    const uint32_t box[3] = {BLOCK_K, BLOCK_M, 2};
    int rc = encode_map(ctx, &mh0, h_hi0, 3, dims, str, box);
    if (rc) return rc;
    if ((rc = encode_map(ctx, &mh1, h_hi1, 3, dims, str, box))) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)ld, (uint64_t)(4 * a.H), 2};
    const uint64_t str[2] = {(uint64_t)ld * 2, (uint64_t)a.lstm_plane_elems * 2};
    const uint32_t box[3] = {BLOCK_K, DEC_BN, 2};
    int rc = encode_map(ctx, &ml, a.lstm_planes, 3, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)a.H, (uint64_t)(a.V + 1), 2};
    const uint64_t str[2] = {(uint64_t)a.H * 2, (uint64_t)a.vocab_plane_elems * 2};
    const uint32_t box[3] = {BLOCK_K, DEC_BN, 2};
    int rc = encode_map(ctx, &mv, a.vocab_planes, 3, dims, str, box);
    if (rc) return rc;
  }
  DecodeParams p{};
  p.E = a.E; p.H = a.H; p.V1 = a.V + 1; p.T = a.T;
  p.m_rows = a.m_rows;
  p.h_hi[0] = h_hi0; p.h_hi[1] = h_hi1; p.h_lo[0] = h_lo0; p.h_lo[1] = h_lo1;
  p.cstate = a.cstate;
  p.lstm_bias = a.lstm_bias; p.lstm_acc_scale = a.lstm_acc_scale;
  p.gate_table = a.gate_table; p.tokens = a.tokens;
  p.vocab_bias = a.vocab_bias; p.vocab_acc_scale = a.vocab_acc_scale;
  p.packed = a.packed;
  p.start_token = a.V + 1;
  p.cnt_gate = cnt_gate; p.cnt_vocab = cnt_vocab; p.flag_x = flag_x; p.mt_stride = mt_stride;
  p.dst_final = a.dst_final; p.dst_all = a.dst_all; p.fin_seq = a.fin_seq; p.all_seq = a.all_seq;
  p.logits = a.logits;
  const size_t smem = Cfg<DEC_BN, 1>::SMEM_BYTES;
  {
    cudaError_t e = dc_ensure_smem(ctx, lm_decode_kernel, smem);
    if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "lm_decode smem attribute: %s", cudaGetErrorString(e));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(dc_launch_state(ctx)->num_sms);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, lm_decode_kernel, mx, mh0, mh1, ml, mv, p);
  if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "lm_decode launch: %s", cudaGetErrorString(e));
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
