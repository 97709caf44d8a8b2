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
    regs_shrink_control();
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
    regs_grow_epilogue();
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
      epilogue_tile<DEC_BN, false, false, 1, true>(pb, ep, tmem_base, tfull_bar(0), tempty_bar(0), bias_smem, it.mt, it.nt, kblocks(it), gi, iter,
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
  return (size_t)(3 * T + 1) * mt_stride;
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
  // ---- persistent decoder ----
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
