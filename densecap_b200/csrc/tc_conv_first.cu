// conv1_1 (VGG layer 1: 3 -> 64 channels, K = 27) on the tcgen05 tensor cores.
//
// The layer has no TMA-able operand: its input is the caller's fp32 NCHW image (DenseCapModel.lua:244) and K = 27 is not a
// multiple of anything.  Four "builder" warps therefore assemble the im2col A tile themselves: thread = one pixel of the 8 x 16
// output patch, 27 coalesced image loads (zero padding = predication), split into hi / lo fp16, written as the first four 16-byte
// chunks of a SWIZZLE_128B K-major row (k = c*9 + r*3 + s, padded to 32 with zeros).  The MMA warp multiplies the tile with the
// resident 64 x 32 filter bank in two K = 16 steps (wide hi|lo B operand: 4 MMAs per 128 x 64 outputs) and the usual eight epilogue
// warps apply bias + ReLU and write the split NHWC activation.  The layer becomes what it should be -- bound by its 110 MB output
// write -- instead of by fp32 FMA issue (round 1: 56 us on the CUDA cores).
//
// Warpgroups (512 threads): 0 = {filter-bank loader, MMA issuer + TMEM, 2 idle}, 1-2 = epilogue, 3 = builders.
#include "tc_common.cuh"

namespace tc {

constexpr int CF_THREADS = 512;
constexpr int CF_STAGES = 4;
constexpr uint32_t CF_A_STAGE_BYTES = 2 * A_PLANE_BYTES;                 // hi + lo planes, 128 rows x 128 B (first 64 B of a row used)
constexpr uint32_t CF_B_BYTES = 2 * 64 * BLOCK_K * 2;                    // 64 filters x 64 k (32 used) x {hi, lo}
constexpr int CF_BUILD_WARP0 = 12;
constexpr size_t CF_SMEM_USED = (size_t)CF_STAGES * CF_A_STAGE_BYTES + CF_B_BYTES + EPI_STAGING_BYTES + 256 + 2 * 64 * 4;   // before the input patches
constexpr size_t CF_SMEM_BYTES = CF_SMEM_USED + 2 * 3 * (TILE_H + 2) * (TILE_W + 2) * 4 + 1024;

__global__ void __launch_bounds__(CF_THREADS, 1)
tc_conv_first_kernel(const float* __restrict__ img, const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_out, Problem pb,
                     GemmEpilogue ep) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* smem_aligned = smem_dyn + (smem_base - smem_u32(smem_dyn));
  const uint32_t a_ring = smem_base;
  const uint32_t b_base = smem_base + CF_STAGES * CF_A_STAGE_BYTES;
  const uint32_t staging = b_base + CF_B_BYTES;                            // 8 epilogue staging slots (TMA-store epilogue, tc_common.cuh)
  const uint32_t bar_base = staging + EPI_STAGING_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_aligned + (size_t)CF_STAGES * CF_A_STAGE_BYTES + CF_B_BYTES + EPI_STAGING_BYTES);
  // bars: a_full[4], a_empty[4], b_full, tmem_full[2], tmem_empty[2], tmem holder
  auto afull_bar = [&](int s) { return bar_base + 8u * s; };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (CF_STAGES + s); };
  const uint32_t bfull_bar = bar_base + 8u * (2 * CF_STAGES);
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * CF_STAGES + 1 + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * CF_STAGES + 3 + s); };
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * CF_STAGES + 5);
  float* bias_smem = reinterpret_cast<float*>(smem_aligned + (size_t)CF_STAGES * CF_A_STAGE_BYTES + CF_B_BYTES + EPI_STAGING_BYTES + 256);
  using C = Cfg<64, 1>;                                   // wide 64-column tile: accumulator 128 columns, 2 stages = 256 TMEM columns

  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_out);
    for (int s = 0; s < CF_STAGES; ++s) { mbar_init(afull_bar(s), 128); mbar_init(aempty_bar(s), 1); }
    mbar_init(bfull_bar, 1);
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

  const int num_tiles = pb.m_tiles;
  const int worker = blockIdx.x, num_workers = gridDim.x;
  const int H = pb.H, W = pb.W;

  if (warp < EPI_WARP0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 0) {
      // ---- filter bank: loaded once, resident ----
      if (elect_one()) {
        mbar_arrive_expect_tx(bfull_bar, CF_B_BYTES);
        tma_load_3d(b_base, &tmap_w, bfull_bar, 0, 0, 0);
      }
    } else if (warp == 1) {
      // ---- MMA issuer ----
      if (elect_one()) {
        constexpr uint32_t idesc = make_idesc_f16(BLOCK_M, 64);
        constexpr uint32_t idesc_wide = make_idesc_f16(BLOCK_M, 128);
        mbar_wait(bfull_bar, 0);
        const uint32_t b_hi = b_base;
        int stage = 0; uint32_t phase = 0;
        int gi = 0;
        for (int tile = worker; tile < num_tiles; tile += num_workers, ++gi) {
          const int acc = gi & 1;
          mbar_wait(tempty_bar(acc), (uint32_t)(((gi >> 1) & 1) ^ 1));
          const uint32_t d_tmem = tmem_base + acc * C::ACC_COLS;
          mbar_wait(afull_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t a_hi = a_ring + stage * CF_A_STAGE_BYTES, a_lo = a_hi + A_PLANE_BYTES;
#pragma unroll
          for (int k = 0; k < 2; ++k) {                     // K = 32 (27 taps + zero padding)
            const uint32_t koff = k * UMMA_K * 2;
            umma_f16(d_tmem, make_kmajor_sw128_desc(a_hi + koff), make_kmajor_sw128_desc(b_hi + koff), idesc_wide, k != 0);   // A_hi.[W_hi | W_lo]
            umma_f16(d_tmem + 64, make_kmajor_sw128_desc(a_lo + koff), make_kmajor_sw128_desc(b_hi + koff), idesc, 1);        // A_lo.W_hi -> small half
          }
          umma_commit(aempty_bar(stage));
          umma_commit(tfull_bar(acc));
          if (++stage == CF_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp < CF_BUILD_WARP0) {
    // ---- epilogue warps: bias + ReLU + fp16 split, NHWC store ----
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    int gi = 0, iter = 0;
    for (int tile = worker; tile < num_tiles; tile += num_workers, ++iter)
      epilogue_tile<64, true, false, 1, true, FEAT_CONV, true>(pb, ep, tmem_base, tfull_bar(0), tempty_bar(0), bias_smem, tile, 0, 1, gi, iter, 0u, true,
                                                              OutStage{&tmap_out, staging});
    if ((threadIdx.x & 31) == 0) bulk_wait0();               // the staging slots must outlive the last bulk stores
  } else {
    // ---- builders: im2col rows of the 8 x 16 patch ----
    asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    const int t = threadIdx.x - 32 * CF_BUILD_WARP0;         // 0 .. 127 = row of the A tile
    const int py = t / TILE_W, px = t % TILE_W;
    const int per_img = pb.tiles_h * pb.tiles_w;
    const size_t plane = (size_t)H * W;
    // The 128 im2col rows of a tile overlap almost completely (3456 taps, 540 distinct input values): the builders stage the
    // 3 x 10 x 18 input halo of the tile in shared memory (coalesced row loads, zero padding by predication) and every thread reads its
    // 27 taps from there.  Two patch buffers: the halo of the NEXT tile is requested (registers) before the current tile is assembled.
    constexpr int PATCH = 3 * (TILE_H + 2) * (TILE_W + 2);   // 540 floats
    constexpr int PLOADS = (PATCH + 127) / 128;              // 5 loads per thread
    float* patch = reinterpret_cast<float*>(smem_aligned + CF_SMEM_USED);      // [2][PATCH]
    auto fetch = [&](int tile, float (&r)[PLOADS]) {
      const int n = tile / per_img, rr = tile % per_img;
      const int h0 = (rr / pb.tiles_w) * TILE_H - 1, w0 = (rr % pb.tiles_w) * TILE_W - 1;
      const float* base = img + (size_t)n * 3 * plane;
#pragma unroll
      for (int i = 0; i < PLOADS; ++i) {
        const int e = t + 128 * i;
        const int c = e / ((TILE_H + 2) * (TILE_W + 2)), rem = e % ((TILE_H + 2) * (TILE_W + 2));
        const int yy = h0 + rem / (TILE_W + 2), xx = w0 + rem % (TILE_W + 2);
        r[i] = (e < PATCH && tile < num_tiles && yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(base + (size_t)c * plane + (size_t)yy * W + xx) : 0.f;
      }
    };
    auto deposit = [&](int buf, const float (&r)[PLOADS]) {
#pragma unroll
      for (int i = 0; i < PLOADS; ++i) {
        const int e = t + 128 * i;
        if (e < PATCH) patch[buf * PATCH + e] = r[i];
      }
    };
    int stage = 0; uint32_t phase = 0;
    float nxt[PLOADS];
    fetch(worker, nxt);
    deposit(0, nxt);
    asm volatile("bar.sync 3, 128;" ::: "memory");
    int buf = 0;
    for (int tile = worker; tile < num_tiles; tile += num_workers, buf ^= 1) {
      fetch(tile + num_workers, nxt);                        // in flight while this tile is assembled
      const float* pt = patch + buf * PATCH + py * (TILE_W + 2) + px;
      float v[32];
#pragma unroll
      for (int k = 0; k < 27; ++k) {
        const int c = k / 9, r = (k % 9) / 3, s = k % 3;
        v[k] = pt[c * ((TILE_H + 2) * (TILE_W + 2)) + r * (TILE_W + 2) + s];
      }
#pragma unroll
      for (int k = 27; k < 32; ++k) v[k] = 0.f;
      uint32_t hh[16], ll[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        __half h0, l0, h1, l1;
        split_f32(v[2 * j], h0, l0);
        split_f32(v[2 * j + 1], h1, l1);
        hh[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        ll[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
      }
      mbar_wait(aempty_bar(stage), phase ^ 1);
      unsigned char* row_hi = smem_aligned + (size_t)stage * CF_A_STAGE_BYTES + (size_t)t * 128;
      unsigned char* row_lo = row_hi + A_PLANE_BYTES;
#pragma unroll
      for (int j = 0; j < 4; ++j) {                          // logical 16-byte chunk j of the row lives at chunk j ^ (row % 8)
        const int pj = (j ^ (t & 7)) * 16;
        *reinterpret_cast<uint4*>(row_hi + pj) = make_uint4(hh[4 * j], hh[4 * j + 1], hh[4 * j + 2], hh[4 * j + 3]);
        *reinterpret_cast<uint4*>(row_lo + pj) = make_uint4(ll[4 * j], ll[4 * j + 1], ll[4 * j + 2], ll[4 * j + 3]);
      }
      fence_proxy_async();                                   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      mbar_arrive(afull_bar(stage));
      if (++stage == CF_STAGES) { stage = 0; phase ^= 1; }
      deposit(buf ^ 1, nxt);                                 // the other buffer was last read one tile ago (barrier below, previous iteration)
      asm volatile("bar.sync 3, 128;" ::: "memory");
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

}  // namespace tc

// w_planes: split [64][64] K-major filter bank (k = c*9 + r*3 + s for k < 27, zeros above), see k_conv_first_weight_to_split
int k_tc_conv_first(dc_ctx* ctx, const float* img, const __half* w_planes, size_t w_plane_elems, const float* bias, float acc_scale, __half* y_hi,
                    __half* y_lo, int N, int H, int W, cudaStream_t st) {
  using namespace tc;
  if ((size_t)N * H * W == 0) return DC_OK;
  Problem pb{};
  pb.H = H; pb.W = W; pb.Cin = 3;
  pb.N = 64; pb.K = 32; pb.num_k_blocks = 1; pb.kb_per_group = 1; pb.split_k = 1;
  pb.tiles_w = DC_CEIL_DIV(W, TILE_W);
  pb.tiles_h = DC_CEIL_DIV(H, TILE_H);
  pb.m_tiles = N * pb.tiles_h * pb.tiles_w;
  pb.n_tiles = 1;
  pb.M = pb.m_tiles * BLOCK_M;
  pb.OH = H; pb.OW = W;
  GemmEpilogue ep;
  ep.bias = bias; ep.acc_scale = acc_scale; ep.relu = 1;
  ep.out_hi = y_hi; ep.out_lo = y_lo; ep.ldo = 64;
  ep.overflow = dc_overflow_flag(ctx);
  CUtensorMap mw;
  {
    const uint64_t dims[3] = {64, 64, 2};
    const uint64_t str[2] = {64 * 2, (uint64_t)w_plane_elems * 2};
    const uint32_t box[3] = {BLOCK_K, 64, 2};
    int rc = encode_map(ctx, &mw, w_planes, 3, dims, str, box);
    if (rc) return rc;
  }
  CUtensorMap mo;
  {
    DC_REQUIRE(ctx, y_lo > y_hi && ((size_t)(y_lo - y_hi) * 2) % 16 == 0, DC_E_UNSUPPORTED, "tc_conv_first: lo plane must follow the hi plane (16-byte aligned)");
    const uint64_t dims[5] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)N, 2};
    const uint64_t str[4] = {64 * 2, (uint64_t)W * 64 * 2, (uint64_t)H * W * 64 * 2, (uint64_t)(y_lo - y_hi) * 2};
    const uint32_t box[5] = {32, 16, 2, 1, 2};
    int rc = encode_map(ctx, &mo, y_hi, 5, dims, str, box, 64);
    if (rc) return rc;
  }
  cudaError_t e = dc_ensure_smem(ctx, tc_conv_first_kernel, CF_SMEM_BYTES);
  if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "tc_conv_first smem attribute: %s", cudaGetErrorString(e));
  const int sms = dc_launch_state(ctx)->num_sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pb.m_tiles < sms ? pb.m_tiles : sms);
  cfg.blockDim = dim3(CF_THREADS);
  cfg.dynamicSmemBytes = CF_SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, tc_conv_first_kernel, img, mw, mo, pb, ep);
  if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "tc_conv_first launch: %s", cudaGetErrorString(e));
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
