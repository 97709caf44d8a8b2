// LanguageModel:sample (densecap/LanguageModel.lua:293-348), greedy decode, as ONE persistent tcgen05 kernel over the
// rows that survive the final NMS (DenseCapModel.lua:261-275 keeps only those rows of the LM output).
//
// Per token the reference runs: LookupTable -> nn.LSTM step (gates = b + x Wx + h Wh, cell) -> vocabulary Linear -> max.
// Here the whole T-step loop is one launch of 148 CTAs that walk a static list of work items
//     Gf(mt,nt)                      first LSTM step on the encoded image, h0 = c0 = 0 (K = E)
//     for t: { G_t(mt,nt) x16 , V_t(mt,nt) x83 }  per 128-row block mt   (gate GEMM + fused cell ; vocabulary GEMM + arg-max)
// round-robin over the CTAs.  Rows are independent, so dependencies exist only INSIDE a 128-row block:
//     G_t(mt,*) needs x_t  = lookup(token_{t-1}) of block mt   (flag_x[t][mt], set by whoever finishes V_{t-1}(mt,*) last)
//               and h_{t-1} of block mt                       (implied: V_{t-1} waited for every G_{t-1}(mt,*))
//     V_t(mt,*) needs h_t of block mt                         (cnt_gate[t+1][mt] == 16)
// They are tracked with per-block counters in global memory (release/acquire at gpu scope + a generic->async proxy fence,
// because the consumers read the operands with TMA), not with grid-wide barriers: block 0 can be in its vocabulary
// projection while block 3 still runs its gate GEMM.  The CTA whose epilogue completes the LAST vocabulary tile of a block
// turns the packed arg-max cells into tokens (torch.max: first maximal index), writes them to the outputs and copies the
// embedding rows of the next step into the [x | h] operand -- the "lm_token_step" launches of round 1 are gone.
// Every CTA processes its items in list order and the list is a topological order of the dependency graph, so the
// smallest unfinished item can always run: no deadlock as long as all CTAs are resident (grid <= #SMs, 1 CTA/SM).
// All waits are bounded (trap after ~2 s) so that a protocol bug surfaces as a CUDA error, never as a hung GPU.
//
// The MMA / TMA machinery (128x128 tiles, "wide" hi|lo B operand: 2 MMAs per K step, 3-stage smem ring, double-buffered
// TMEM accumulator, 8 epilogue warps) is the one of tc_gemm_kernel<128, gemm, 1>; ring and accumulator parities simply
// continue across items.
#include "tc_common.cuh"

namespace tc {

constexpr int DEC_BN = 128;

struct DecodeParams {
  int E, H, V1, T, ld;                 // V1 = V + 1 (logit width), ld = E + H
  const int32_t* m_rows;               // device: number of rows to decode (M')
  __half* xh_hi[2];                    // [x | h] operand, buffers 0 / 1, hi plane (rows, ld)
  __half* xh_lo[2];
  float* cstate;                       // (rows, H)
  const float* lstm_bias;              // (4H) gate-interleaved
  float lstm_acc_scale;
  const float* vocab_bias;             // (V1)
  float vocab_acc_scale;
  const __half* lut_hi;                // lookup table planes (V+2, E)
  const __half* lut_lo;
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
lm_decode_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
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
    prefetch_tmap(&map_a0); prefetch_tmap(&map_a1); prefetch_tmap(&map_lstm); prefetch_tmap(&map_vocab);
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
  auto kblocks = [&](const DecItem& it) { return it.vocab ? p.H / BLOCK_K : (it.t < 0 ? p.E / BLOCK_K : p.ld / BLOCK_K); };
  const int worker = blockIdx.x, num_workers = gridDim.x;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int g = worker; g < total_items; g += num_workers) {
        const DecItem it = decode(g);
        const CUtensorMap* ma;
        const CUtensorMap* mb;
        int ka0;
        if (it.vocab) {
          wait_counter(p.cnt_gate + (size_t)(it.t + 1) * p.mt_stride + it.mt, (uint32_t)GN);      // h_t of this row block complete
          ma = (it.t & 1) ? &map_a1 : &map_a0;
          mb = &map_vocab;
          ka0 = p.E;
        } else {
          if (it.t == 0) wait_counter(p.cnt_gate + it.mt, (uint32_t)GN);                          // first step complete (h in buffer 1)
          else if (it.t > 0) wait_counter(p.flag_x + (size_t)it.t * p.mt_stride + it.mt, 1u);      // x_t looked up (implies h_{t-1})
          ma = (it.t < 0 || (it.t & 1)) ? &map_a0 : &map_a1;                                       // src buffer: first step 0, then 1,0,1,...
          mb = &map_lstm;
          ka0 = 0;
        }
        fence_proxy_async_global();
        const int nkb = kblocks(it);
        const int m0 = it.mt * BLOCK_M, n0 = it.nt * DEC_BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t a_dst = smem_base + stage * C::STAGE_BYTES;
          const uint32_t b_dst = a_dst + 2 * A_PLANE_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
          tma_load_3d(a_dst, ma, full_bar(stage), ka0 + kb * BLOCK_K, m0, 0);
          tma_load_3d(b_dst, mb, full_bar(stage), kb * BLOCK_K, n0, 0);
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
      int iter = 0;
      for (int g = worker; g < total_items; g += num_workers, ++iter) {
        const DecItem it = decode(g);
        const int nkb = kblocks(it);
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * C::ACC_COLS;
        for (int kb = 0; kb < nkb; ++kb) {
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
            umma_f16(d_tmem, da_hi, db_hi, idesc_wide, (kb | k) != 0);     // A_hi.[B_hi | B_lo]
            umma_f16(d_tmem, da_lo, db_hi, idesc, 1);                      // A_lo.B_hi
          }
          umma_commit(empty_bar(stage));
          if (kb == nkb - 1) umma_commit(tfull_bar(acc));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // =============================== epilogue warps ===============================
    const int et = threadIdx.x - 64;                         // 0 .. 255
    int iter = 0;
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
        ep.bias = p.lstm_bias;
        ep.acc_scale = p.lstm_acc_scale;
        ep.lstm_c = p.cstate;
        ep.lstm_H = p.H;
        ep.lstm_first = it.t < 0;
        const int dst = (it.t < 0) ? 1 : (it.t & 1);
        ep.out_hi = p.xh_hi[dst];
        ep.out_lo = p.xh_lo[dst];
        ep.ldo = p.ld;
        ep.out_col0 = p.E;
      }
      epilogue_tile<DEC_BN, false, false, 1, true>(pb, ep, tmem_base, tfull_bar(0), tempty_bar(0), bias_smem, it.mt, it.nt, iter & 1,
                                                   (uint32_t)((iter >> 1) & 1), 0u, true);
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
      // ---- last vocabulary tile of row block mt at step t: tokens of step t, embedding rows of step t+1 ----
      {
        const int lane = threadIdx.x & 31, ew = et >> 5;     // 8 warps x 16 rows
        const int row0 = it.mt * BLOCK_M;
        const bool more = it.t + 1 < p.T;
        const int dstb = it.t & 1;                           // buffer whose x half feeds G_{t+1}
        constexpr int RPW = BLOCK_M / EPI_WARPS;             // 16 rows per warp
#pragma unroll 1
        for (int r0 = 0; r0 < RPW; r0 += 4) {
          int tok[4];
          int rows[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = row0 + ew * RPW + r0 + u;
            rows[u] = r;
            tok[u] = 1;
            if (r < M) {
              const unsigned long long pk = __ldcg(p.packed + r);
              const uint32_t col = 0xFFFFFFFFu - (uint32_t)(pk & 0xFFFFFFFFull);
              tok[u] = (pk == 0ull) ? 1 : (int)min(col, (uint32_t)p.start_token - 1u) + 1;     // all-NaN row -> first index, like torch.max
              if (lane == 0) {
                const int df = p.dst_final ? __ldg(p.dst_final + r) : -1;
                const int da = p.dst_all ? __ldg(p.dst_all + r) : -1;
                if (df >= 0) p.fin_seq[(size_t)df * p.T + it.t] = tok[u];
                if (da >= 0) p.all_seq[(size_t)da * p.T + it.t] = tok[u];
                p.packed[r] = 0ull;
              }
            }
          }
          if (more) {
            // E halves per plane and row = E/8 uint4; lanes stride over them (E = 512: 2 per lane and plane)
            for (int e = lane; e < p.E / 8; e += 32) {
              uint4 vh[4], vl[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const size_t src = (size_t)(tok[u] - 1) * p.E + (size_t)e * 8;
                vh[u] = __ldg(reinterpret_cast<const uint4*>(p.lut_hi + src));
                vl[u] = __ldg(reinterpret_cast<const uint4*>(p.lut_lo + src));
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                if (rows[u] < M) {
                  const size_t dst = (size_t)rows[u] * p.ld + (size_t)e * 8;
                  *reinterpret_cast<uint4*>(p.xh_hi[dstb] + dst) = vh[u];
                  *reinterpret_cast<uint4*>(p.xh_lo[dstb] + dst) = vl[u];
                }
              }
            }
          }
        }
        fence_proxy_async_global();
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
  int B, P, E, ld, T;
  const int32_t* counts;
  const int32_t* idx32;
  int survivors;                      // 1: dst_final = b*P + i (fin_seq written by the decoder), 0: dst_final = -1
  const __half* enc_hi; const __half* enc_lo;
  __half* x0_hi; __half* x0_lo; __half* x1_hi; __half* x1_lo;
  const __half* lut_hi; const __half* lut_lo;
  int start_token;
  int32_t* m_rows; int32_t* dst_final; int32_t* dst_all;
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
  const size_t start_src = (size_t)(c.start_token - 1) * c.E;
  for (int r = blockIdx.x * 8 + warp; r < M; r += gridDim.x * 8) {
    int b = 0;
    while (b + 1 < c.B && r >= s_off[b + 1]) ++b;
    const int i = r - s_off[b];
    const int local = c.idx32 ? __ldcg(c.idx32 + (size_t)b * c.P + i) : i;
    const int src = b * c.P + local;
    if (lane == 0) {
      c.dst_final[r] = c.survivors ? b * c.P + i : -1;
      c.dst_all[r] = src;
      c.packed[r] = 0ull;
    }
    for (int e = lane; e < c.E / 8; e += 32) {
      const size_t so = (size_t)src * c.E + (size_t)e * 8, d = (size_t)r * c.ld + (size_t)e * 8;
      *reinterpret_cast<uint4*>(c.x0_hi + d) = __ldcg(reinterpret_cast<const uint4*>(c.enc_hi + so));
      *reinterpret_cast<uint4*>(c.x0_lo + d) = __ldcg(reinterpret_cast<const uint4*>(c.enc_lo + so));
      *reinterpret_cast<uint4*>(c.x1_hi + d) = __ldg(reinterpret_cast<const uint4*>(c.lut_hi + start_src + (size_t)e * 8));
      *reinterpret_cast<uint4*>(c.x1_lo + d) = __ldg(reinterpret_cast<const uint4*>(c.lut_lo + start_src + (size_t)e * 8));
    }
  }
}

}  // namespace tc

size_t lm_decode_counter_words(int T, int rows_cap) {
  const int mt_stride = DC_CEIL_DIV(rows_cap, tc::BLOCK_M) + 1;
  return (size_t)(3 * T + 1) * mt_stride;
}

int k_lm_decode(dc_ctx* ctx, const LmDecodeArgs& a, cudaStream_t st) {
  using namespace tc;
  DC_REQUIRE(ctx, a.E % BLOCK_K == 0 && a.H % BLOCK_K == 0 && (4 * a.H) % DEC_BN == 0 && a.T >= 1 && a.B >= 1 && a.B <= 64, DC_E_UNSUPPORTED,
             "lm_decode: encoding / rnn size must be multiples of 64 (4H of 128), batch <= 64");
  const int ld = a.E + a.H;
  const int rows_cap = a.B * a.P;
  const int mt_stride = DC_CEIL_DIV(rows_cap, BLOCK_M) + 1;
  const size_t bufelems = (size_t)rows_cap * ld;             // one plane of one [x|h] buffer
  __half* xh_hi0 = a.xh_planes;
  __half* xh_hi1 = a.xh_planes + bufelems;
  __half* xh_lo0 = a.xh_planes + 2 * bufelems;
  __half* xh_lo1 = a.xh_planes + 3 * bufelems;
  uint32_t* cnt_gate = a.counters;
  uint32_t* cnt_vocab = cnt_gate + (size_t)(a.T + 1) * mt_stride;
  uint32_t* flag_x = cnt_vocab + (size_t)a.T * mt_stride;
  // ---- compaction ----
  {
    CompactParams c{};
    c.B = a.B; c.P = a.P; c.E = a.E; c.ld = ld; c.T = a.T;
    c.counts = a.counts; c.idx32 = a.idx32; c.survivors = a.survivors;
    c.enc_hi = a.enc_planes; c.enc_lo = a.enc_planes + a.enc_plane_elems;
    c.x0_hi = xh_hi0; c.x0_lo = xh_lo0; c.x1_hi = xh_hi1; c.x1_lo = xh_lo1;
    c.lut_hi = a.lut_planes; c.lut_lo = a.lut_planes + a.lut_plane_elems;
    c.start_token = a.V + 1;
    c.m_rows = a.m_rows; c.dst_final = a.dst_final; c.dst_all = a.dst_all;
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
  // ---- persistent decoder ----
  CUtensorMap ma0, ma1, ml, mv;
  {
    const uint64_t dims[3] = {(uint64_t)ld, (uint64_t)rows_cap, 2};
    const uint64_t str[2] = {(uint64_t)ld * 2, (uint64_t)(2 * bufelems) * 2};
    const uint32_t box[3] = {BLOCK_K, BLOCK_M, 2};
    int rc = encode_map(ctx, &ma0, xh_hi0, 3, dims, str, box);
    if (rc) return rc;
    if ((rc = encode_map(ctx, &ma1, xh_hi1, 3, dims, str, box))) return rc;
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
  p.E = a.E; p.H = a.H; p.V1 = a.V + 1; p.T = a.T; p.ld = ld;
  p.m_rows = a.m_rows;
  p.xh_hi[0] = xh_hi0; p.xh_hi[1] = xh_hi1; p.xh_lo[0] = xh_lo0; p.xh_lo[1] = xh_lo1;
  p.cstate = a.cstate;
  p.lstm_bias = a.lstm_bias; p.lstm_acc_scale = a.lstm_acc_scale;
  p.vocab_bias = a.vocab_bias; p.vocab_acc_scale = a.vocab_acc_scale;
  p.lut_hi = a.lut_planes; p.lut_lo = a.lut_planes + a.lut_plane_elems;
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, lm_decode_kernel, ma0, ma1, ml, mv, p);
  if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "lm_decode launch: %s", cudaGetErrorString(e));
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
