// nn.BilinearRoiPooling (densecap/modules/BilinearRoiPooling.lua:35-67,84-91) as two kernels:
//   roi_taps_kernel : BoxToAffine (modules/BoxToAffine.lua:69-93) -> stnbhwd AffineGridGeneratorBHWD(HH,WW)
//                     -> per output position the 4 bilinear corner offsets + weights of stnbhwd
//                     BilinearSamplerBHWD_updateOutput; one thread per (roi, pos).
//   roi_pool_kernel : the gather itself (modules/BatchBilinearSamplerBHWD.lua:104-122), with the two
//                     nn.Transpose copies removed by layout choice (features NHWC in, BHWC out).
//
// B200 design of the gather: the op is HBM-write bound (K*HH*WW*C*4 B out = 100 MB at K=1000) but every
// output element reads 4 feature values, i.e. 4x read amplification if served from L2.  Each CTA
// therefore owns one 32-channel slice of the whole feature map, staged ONCE into shared memory
// (38*45*32*4 B = 219 KB of the 227 KB, plus one all-zero row that out-of-map corners point at), and serves
// all of its gathers from there.  Four lanes cover the 32 channels of one output position (8 channels per
// lane: 2 LDS.128 per corner; odd positions read their two halves in the opposite order so that one LDS
// instruction of the warp touches every bank exactly 4 times = the 512-byte minimum), so the 32-byte tap
// record and its address arithmetic are amortised over 8 channels; the hi/lo split is done on packed
// half2 pairs (cvt.rn.f16x2.f32).  grid = (C/32 slices) x (148/(C/32) ROI groups), 1024 threads.
// The slice staging does not depend on the taps, so the kernel is launched with programmatic stream
// serialisation and overlaps roi_taps_kernel; it waits (griddepcontrol.wait) only before the first tap read.
// Output rows are slice-major (roi_feature_index, common.cuh): the 8 positions a warp emits per step are 512 contiguous
// bytes per plane, so the plain 16-byte stores are fully coalesced.  ncu (profiles/kernels_r08_notes.md): shared-memory
// wavefronts = ideal, LSU data pipe 72 % busy, issue slots 57 % -- the gather is bound by shared-memory bandwidth, not HBM.
//
// The interpolation arithmetic follows the reference CPU kernel's operation order without FMA
// contraction (mul/add intrinsics), so fp32 outputs are bit-identical to the oracle.
#include "tc_common.cuh"

constexpr int ROI_THREADS = 1024;
constexpr int ROI_MAX_POS = 64;  // HH*WW <= 64

struct __align__(16) RoiTap {
  int off[4];     // float index (cell*ch) of TL,TR,BL,BR inside a ch-channel slice; outside -> the zero row (fh*fw*ch)
  float w[4];     // (wx*wy), ((1-wx)*wy), (wx*(1-wy)), ((1-wx)*(1-wy))
};

size_t roi_taps_bytes(int max_boxes, int npos) { return (size_t)max_boxes * npos * sizeof(RoiTap) + 256; }

__global__ void roi_taps_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ nboxes_dev, int max_boxes,
                                float img_h, float img_w, int fh, int fw, int HH, int WW, int ch, RoiTap* __restrict__ taps) {
  // let the dependent roi_pool_kernel start staging its feature slice right away (it waits before reading taps); this kernel is itself
  // launched with programmatic stream serialisation, so wait for the producer of `boxes` / `nboxes_dev` before reading them
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int npos = HH * WW;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= max_boxes * npos) return;
  const int b = idx / npos, p = idx % npos;
  const int nboxes = nboxes_dev ? min(*nboxes_dev, max_boxes) : max_boxes;
  RoiTap t;
  const int zero_row = fh * fw * ch;
  if (b >= nboxes) {
    t.off[0] = t.off[1] = t.off[2] = t.off[3] = zero_row;
    t.w[0] = t.w[1] = t.w[2] = t.w[3] = 0.f;
    taps[idx] = t;
    return;
  }
  const float4 box = boxes[b];
  // BoxToAffine.lua:87-90
  const float th11 = __fdiv_rn(box.w, img_h);
  const float th13 = __fdiv_rn(__fadd_rn(__fmul_rn(box.y, 2.f), -1.f - img_h), img_h - 1.f);
  const float th22 = __fdiv_rn(box.z, img_w);
  const float th23 = __fdiv_rn(__fadd_rn(__fmul_rn(box.x, 2.f), -1.f - img_w), img_w - 1.f);
  const int i = p / WW, j = p % WW;
  // AffineGridGeneratorBHWD base grid: -1 + i/(HH-1)*2 evaluated in double, stored as float
  const float yb = (float)(-1.0 + ((double)i / (double)(HH - 1)) * 2.0);
  const float xb = (float)(-1.0 + ((double)j / (double)(WW - 1)) * 2.0);
  // grid = theta . (y, x, 1); theta12 = theta21 = 0 (their products are exact zeros)
  const float gy = __fadd_rn(__fadd_rn(__fmul_rn(th11, yb), __fmul_rn(0.f, xb)), th13);
  const float gx = __fadd_rn(__fadd_rn(__fmul_rn(0.f, yb), __fmul_rn(th22, xb)), th23);
  // BilinearSamplerBHWD: coordinates on the FEATURE map
  const float xcoord = __fdiv_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)(fw - 1)), 2.f);
  const float ycoord = __fdiv_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)(fh - 1)), 2.f);
  const float xf = floorf(xcoord), yf = floorf(ycoord);
  const float wx = __fsub_rn(1.f, __fsub_rn(xcoord, xf));
  const float wy = __fsub_rn(1.f, __fsub_rn(ycoord, yf));
  // clamp before the int conversion so that huge/NaN coordinates simply fall outside
  const int xl = (int)fminf(fmaxf(xf, -2.f), (float)fw + 1.f);
  const int yt = (int)fminf(fmaxf(yf, -2.f), (float)fh + 1.f);
  const float omx = __fsub_rn(1.f, wx), omy = __fsub_rn(1.f, wy);
  t.w[0] = __fmul_rn(wx, wy);
  t.w[1] = __fmul_rn(omx, wy);
  t.w[2] = __fmul_rn(wx, omy);
  t.w[3] = __fmul_rn(omx, omy);
  const bool xin0 = xl >= 0 && xl <= fw - 1, xin1 = xl + 1 >= 0 && xl + 1 <= fw - 1;
  const bool yin0 = yt >= 0 && yt <= fh - 1, yin1 = yt + 1 >= 0 && yt + 1 <= fh - 1;
  const bool finite = (xcoord == xcoord) && (ycoord == ycoord);
  t.off[0] = (finite && xin0 && yin0) ? (yt * fw + xl) * ch : zero_row;
  t.off[1] = (finite && xin1 && yin0) ? (yt * fw + xl + 1) * ch : zero_row;
  t.off[2] = (finite && xin0 && yin1) ? ((yt + 1) * fw + xl) * ch : zero_row;
  t.off[3] = (finite && xin1 && yin1) ? ((yt + 1) * fw + xl + 1) * ch : zero_row;
  taps[idx] = t;
}

// Tap records from CALLER-SUPPLIED sampling grids (stand-alone nn.BatchBilinearSamplerBHWD, modules/BatchBilinearSamplerBHWD.lua:104-122):
// grids (B, HH, WW, 2) = (gy, gx) in [-1, 1]; same coordinate arithmetic as roi_taps_kernel from `xcoord` on.
__global__ void roi_taps_from_grid_kernel(const float2* __restrict__ grids, int total, int fh, int fw, int ch, RoiTap* __restrict__ taps) {
  asm volatile("griddepcontrol.launch_dependents;");
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const float2 g = grids[idx];
  const float gy = g.x, gx = g.y;
  RoiTap t;
  const int zero_row = fh * fw * ch;
  const float xcoord = __fdiv_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)(fw - 1)), 2.f);
  const float ycoord = __fdiv_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)(fh - 1)), 2.f);
  const float xf = floorf(xcoord), yf = floorf(ycoord);
  const float wx = __fsub_rn(1.f, __fsub_rn(xcoord, xf));
  const float wy = __fsub_rn(1.f, __fsub_rn(ycoord, yf));
  const int xl = (int)fminf(fmaxf(xf, -2.f), (float)fw + 1.f);
  const int yt = (int)fminf(fmaxf(yf, -2.f), (float)fh + 1.f);
  const float omx = __fsub_rn(1.f, wx), omy = __fsub_rn(1.f, wy);
  t.w[0] = __fmul_rn(wx, wy);
  t.w[1] = __fmul_rn(omx, wy);
  t.w[2] = __fmul_rn(wx, omy);
  t.w[3] = __fmul_rn(omx, omy);
  const bool xin0 = xl >= 0 && xl <= fw - 1, xin1 = xl + 1 >= 0 && xl + 1 <= fw - 1;
  const bool yin0 = yt >= 0 && yt <= fh - 1, yin1 = yt + 1 >= 0 && yt + 1 <= fh - 1;
  const bool finite = (xcoord == xcoord) && (ycoord == ycoord);
  t.off[0] = (finite && xin0 && yin0) ? (yt * fw + xl) * ch : zero_row;
  t.off[1] = (finite && xin1 && yin0) ? (yt * fw + xl + 1) * ch : zero_row;
  t.off[2] = (finite && xin0 && yin1) ? ((yt + 1) * fw + xl) * ch : zero_row;
  t.off[3] = (finite && xin1 && yin1) ? ((yt + 1) * fw + xl + 1) * ch : zero_row;
  taps[idx] = t;
}

__device__ __forceinline__ float roi_blend(const RoiTap& t, float tl, float tr, float bl, float br) {
  float v = __fmul_rn(t.w[0], tl);
  v = __fadd_rn(v, __fmul_rn(t.w[1], tr));
  v = __fadd_rn(v, __fmul_rn(t.w[2], bl));
  v = __fadd_rn(v, __fmul_rn(t.w[3], br));
  return v;
}

// two fp32 values -> packed (hi, hi) and (lo, lo) half2 words of the split-fp16 representation (common.cuh split_f32)
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));        // low half <- a
  const __half2 h = *reinterpret_cast<const __half2*>(&hi);
  const float ra = __fsub_rn(a, __low2float(h)), rb = __fsub_rn(b, __high2float(h));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}

// Packed fp32x2 arithmetic (sm_100 FMUL2 / FADD2 / FFMA2): two IEEE single-precision operations per issue slot.  The blend must keep
// the reference's unfused operation order, and ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (also when the product is spelled
// fma(a, b, -0.0) with a literal -0), so the product is fma(a, b, nz) with nz = (-0.0, -0.0) passed in as a KERNEL ARGUMENT: adding -0
// leaves every product (and the sign of a zero product) unchanged, and an opaque addend leaves nothing to contract.
__device__ __forceinline__ unsigned long long f2_pack(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b, unsigned long long nz) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz));
  return r;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

template <int CH>
__global__ void __launch_bounds__(ROI_THREADS, 1)
roi_pool_kernel(const __half* __restrict__ feat_hi, const __half* __restrict__ feat_lo, const float* __restrict__ feats_f32,
                int C, int fh, int fw, const RoiTap* __restrict__ taps, int total_items, int npos,
                __half* __restrict__ out_hi, __half* __restrict__ out_lo, float* __restrict__ out_f32, int out_layout, unsigned long long nz2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* slice = reinterpret_cast<float*>(smem_raw);                               // [fh*fw + 1][CH], last row = zeros
  constexpr int LPP = CH / 8;                   // lanes per output position (8 channels per lane)
  constexpr int POS_PER_WARP = 32 / LPP;        // positions handled concurrently by one warp
  const int c0 = blockIdx.x * CH;
  const int cells = fh * fw;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NWARPS = ROI_THREADS / 32;
  // ---- stage the channel slice (4 channels per thread and step) ---------------------------------
  {
    constexpr int QUADS = CH / 4;
#pragma unroll 4                                   // several independent L2 round trips in flight per thread
    for (int e = threadIdx.x; e < cells * QUADS; e += ROI_THREADS) {
      const int cell = e / QUADS, cq4 = (e % QUADS) * 4;
      const size_t g = (size_t)cell * C + c0 + cq4;
      float4 v;
      if (feats_f32) {
        v = *reinterpret_cast<const float4*>(feats_f32 + g);
      } else {
        const uint2 h = *reinterpret_cast<const uint2*>(feat_hi + g);
        const uint2 l = *reinterpret_cast<const uint2*>(feat_lo + g);
        const __half2 h01 = *reinterpret_cast<const __half2*>(&h.x), h23 = *reinterpret_cast<const __half2*>(&h.y);
        const __half2 l01 = *reinterpret_cast<const __half2*>(&l.x), l23 = *reinterpret_cast<const __half2*>(&l.y);
        v.x = __low2float(h01) + __low2float(l01);
        v.y = __high2float(h01) + __high2float(l01);
        v.z = __low2float(h23) + __low2float(l23);
        v.w = __high2float(h23) + __high2float(l23);
      }
      *reinterpret_cast<float4*>(slice + cell * CH + cq4) = v;
    }
    if (threadIdx.x < CH) slice[cells * CH + threadIdx.x] = 0.f;
  }
  __syncthreads();
  // the taps come from roi_taps_kernel, the launch just before this one (programmatic dependent launch)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int sub = lane / LPP;
  const int half = sub & 1;                               // odd positions read their two 4-channel halves in swapped order
  const int cA = (lane % LPP) * 8 + 4 * half, cB = (lane % LPP) * 8 + 4 * (half ^ 1);
  // items (roi*npos + pos) are dealt in contiguous runs of POS_PER_WARP to (blockIdx.y, warp)
  const int stride = gridDim.y * NWARPS * POS_PER_WARP;
  // 4 channels of one output position: ((w0 tl + w1 tr) + w2 bl) + w3 br per channel, every product and sum rounded separately
  // (reference operation order), evaluated two channels per instruction
  auto blend4 = [&](const uint4& ta, const uint4& tb, int cc, float (&v)[4]) {
    const float w0 = __uint_as_float(tb.x), w1 = __uint_as_float(tb.y), w2 = __uint_as_float(tb.z), w3 = __uint_as_float(tb.w);
    const unsigned long long p0 = f2_pack(w0, w0), p1 = f2_pack(w1, w1), p2 = f2_pack(w2, w2), p3 = f2_pack(w3, w3);
    const ulonglong2 tl = *reinterpret_cast<const ulonglong2*>(slice + ta.x + cc);
    const ulonglong2 tr = *reinterpret_cast<const ulonglong2*>(slice + ta.y + cc);
    const ulonglong2 bl = *reinterpret_cast<const ulonglong2*>(slice + ta.z + cc);
    const ulonglong2 br = *reinterpret_cast<const ulonglong2*>(slice + ta.w + cc);
    const unsigned long long a01 = f2_add(f2_add(f2_add(f2_mul(p0, tl.x, nz2), f2_mul(p1, tr.x, nz2)), f2_mul(p2, bl.x, nz2)), f2_mul(p3, br.x, nz2));
    const unsigned long long a23 = f2_add(f2_add(f2_add(f2_mul(p0, tl.y, nz2), f2_mul(p1, tr.y, nz2)), f2_mul(p2, bl.y, nz2)), f2_mul(p3, br.y, nz2));
    f2_unpack(a01, v[0], v[1]);
    f2_unpack(a23, v[2], v[3]);
  };
  const int c8 = (lane % LPP) * 8;
  // split output rows are laid out slice-major (roi_feature_index): the 8 positions a warp emits per step are 8 x 64 contiguous bytes per plane
  const size_t row_elems = (size_t)npos * C;
  const size_t col_base = roi_feature_index(c0 + c8, 0, C, npos);
  const int cs = (C % 32 == 0) ? 32 : C;
  auto emit = [&](int it, int roi, int pos, const uint4& ta, const uint4& tb) {
    float va[4], vb[4];
    blend4(ta, tb, cA, va);
    blend4(ta, tb, cB, vb);
    if (out_hi) {
      // one 16-byte store per plane: the lane's 8 channels in memory order (va holds the upper half on odd positions)
      uint32_t ha0, la0, ha1, la1, hb0, lb0, hb1, lb1;
      split_pair(va[0], va[1], ha0, la0);
      split_pair(va[2], va[3], ha1, la1);
      split_pair(vb[0], vb[1], hb0, lb0);
      split_pair(vb[2], vb[3], hb1, lb1);
      const uint4 ph = half ? make_uint4(hb0, hb1, ha0, ha1) : make_uint4(ha0, ha1, hb0, hb1);
      const uint4 pl = half ? make_uint4(lb0, lb1, la0, la1) : make_uint4(la0, la1, lb0, lb1);
      const size_t o = (size_t)roi * row_elems + col_base + (size_t)pos * cs;
      *reinterpret_cast<uint4*>(out_hi + o) = ph;
      *reinterpret_cast<uint4*>(out_lo + o) = pl;
    }
    if (out_f32) {
      if (out_layout == 1) {
        *reinterpret_cast<float4*>(out_f32 + (size_t)it * C + c0 + cA) = make_float4(va[0], va[1], va[2], va[3]);
        *reinterpret_cast<float4*>(out_f32 + (size_t)it * C + c0 + cB) = make_float4(vb[0], vb[1], vb[2], vb[3]);
      } else {
        float* oa = out_f32 + ((size_t)roi * C + c0 + cA) * npos + pos;
        float* ob = out_f32 + ((size_t)roi * C + c0 + cB) * npos + pos;
#pragma unroll
        for (int j = 0; j < 4; ++j) { oa[(size_t)j * npos] = va[j]; ob[(size_t)j * npos] = vb[j]; }
      }
    }
  };
  // (plain L2 loads, not __ldg: the taps are written by the still-running previous launch, so the read-only path
  // contract does not hold and the compiler must not hoist these above griddepcontrol.wait)
  // software pipeline of depth one: the tap record of the next item is requested before the current item is gathered,
  // so the L2 latency of the tap load is covered by this warp's own work
  int it = (blockIdx.y * NWARPS + warp) * POS_PER_WARP + sub;
  int roi = it / npos, pos = it - roi * npos;                   // advanced incrementally (no division in the loop)
  const int step_roi = stride / npos, step_pos = stride - step_roi * npos;
  auto load_tap = [&](int i, uint4& a, uint4& b) {
    a = make_uint4(0, 0, 0, 0); b = a;
    if (i < total_items) { const uint4* tp = reinterpret_cast<const uint4*>(taps + i); a = __ldcg(tp); b = __ldcg(tp + 1); }
  };
  // software pipeline of depth two: tap records are requested two steps ahead of their use (L2 latency ~ one step of this warp)
  uint4 t0, t1, u0, u1;
  load_tap(it, t0, t1);
  load_tap(it + stride, u0, u1);
  while (it < total_items) {
    uint4 n0, n1;
    load_tap(it + 2 * stride, n0, n1);
    emit(it, roi, pos, t0, t1);
    t0 = u0; t1 = u1; u0 = n0; u1 = n1;
    it += stride;
    roi += step_roi; pos += step_pos;
    if (pos >= npos) { pos -= npos; ++roi; }
  }
}

// grids != nullptr: stand-alone sampler on caller-supplied (max_boxes, HH, WW, 2) grids (boxes / image size unused)
static int roi_pool_impl(dc_ctx* ctx, const __half* feat_hi, const __half* feat_lo, const float* feats_f32, int C, int fh, int fw,
                         const float* boxes, const float* grids, const int32_t* nboxes_dev, int max_boxes, int img_h, int img_w, int HH, int WW,
                         void* taps_ws, __half* out_hi, __half* out_lo, float* out_f32, int out_layout, cudaStream_t st) {
  if (max_boxes <= 0) return DC_OK;
  DC_REQUIRE(ctx, HH * WW <= ROI_MAX_POS && HH > 1 && WW > 1, DC_E_UNSUPPORTED, "roi_pool: HH*WW must be in (1,64]");
  DC_REQUIRE(ctx, C % 8 == 0, DC_E_UNSUPPORTED, "roi_pool: C must be a multiple of 8");
  const int npos = HH * WW;
  const int total = max_boxes * npos;
  RoiTap* taps = reinterpret_cast<RoiTap*>(taps_ws);
  const size_t limit = 227 * 1024;
  int ch = 32;
  while (ch > 8 && ((size_t)(fh * fw + 1) * ch * 4 > limit || C % ch != 0)) ch >>= 1;
  const size_t smem = (size_t)(fh * fw + 1) * ch * 4;
  if (grids) roi_taps_from_grid_kernel<<<DC_CEIL_DIV(total, 256), 256, 0, st>>>((const float2*)grids, total, fh, fw, ch, taps);
  else {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(DC_CEIL_DIV(total, 256)); lc.blockDim = dim3(256); lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&lc, roi_taps_kernel, (const float4*)boxes, nboxes_dev, max_boxes, (float)img_h, (float)img_w, fh, fw, HH, WW, ch, taps);
    if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "roi_taps launch: %s", cudaGetErrorString(e));
  }
  DC_CHECK_LAUNCH(ctx);
  DC_REQUIRE(ctx, smem <= limit && C % ch == 0, DC_E_UNSUPPORTED, "roi_pool: feature map %dx%d too large for the shared-memory slice", fh, fw);
  const int slices = C / ch;
  int groups = dc_launch_state(ctx)->num_sms / slices;
  if (groups < 1) groups = 1;
  dim3 grid(slices, groups);
#define ROI_LAUNCH(CHV)                                                                                              \
  do {                                                                                                               \
    cudaError_t e = dc_ensure_smem(ctx, roi_pool_kernel<CHV>, limit);                                                  \
    if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "roi_pool smem attribute: %s", cudaGetErrorString(e));    \
    cudaLaunchConfig_t lc = {};                                                                                      \
    lc.gridDim = grid; lc.blockDim = dim3(ROI_THREADS); lc.dynamicSmemBytes = smem; lc.stream = st;                   \
    cudaLaunchAttribute at[1];                                                                                       \
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                   \
    at[0].val.programmaticStreamSerializationAllowed = 1;                                                            \
    lc.attrs = at; lc.numAttrs = 1;                                                                                  \
    e = cudaLaunchKernelEx(&lc, roi_pool_kernel<CHV>, feat_hi, feat_lo, feats_f32, C, fh, fw, (const RoiTap*)taps, total, npos, \
                           out_hi, out_lo, out_f32, out_layout, 0x8000000080000000ULL);                              \
    if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "roi_pool launch: %s", cudaGetErrorString(e));          \
  } while (0)
  if (ch == 32) ROI_LAUNCH(32);
  else if (ch == 16) ROI_LAUNCH(16);
  else ROI_LAUNCH(8);
#undef ROI_LAUNCH
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}

int k_roi_pool(dc_ctx* ctx, const __half* feat_hi, const __half* feat_lo, const float* feats_f32, int C, int fh, int fw,
               const float* boxes, const int32_t* nboxes_dev, int max_boxes, int img_h, int img_w, int HH, int WW,
               void* taps_ws, __half* out_hi, __half* out_lo, float* out_f32, int out_layout, cudaStream_t st) {
  return roi_pool_impl(ctx, feat_hi, feat_lo, feats_f32, C, fh, fw, boxes, nullptr, nboxes_dev, max_boxes, img_h, img_w, HH, WW, taps_ws, out_hi, out_lo,
                       out_f32, out_layout, st);
}
// nn.BatchBilinearSamplerBHWD:updateOutput: feats (H', W', C) fp32 HWC, grids (B, HH, WW, 2) -> out (B, HH, WW, C) fp32
int k_bilinear_sampler(dc_ctx* ctx, const float* feats_hwc, int C, int fh, int fw, const float* grids, int B, int HH, int WW, void* taps_ws, float* out,
                       cudaStream_t st) {
  return roi_pool_impl(ctx, nullptr, nullptr, feats_hwc, C, fh, fw, nullptr, grids, nullptr, B, 0, 0, HH, WW, taps_ws, nullptr, nullptr, out, 1, st);
}

// ---------------------------------------------------------------------------------------------------------
// Backward of nn.BilinearRoiPooling (modules/BilinearRoiPooling.lua:94-104 -> self.net:backward):
//   roi_pool_bwd_kernel  : stnbhwd BilinearSamplerBHWD_updateGradInput summed over the ROIs as in
//                          modules/BatchBilinearSamplerBHWD.lua:125-157 -- scatter-add of w * gradOutput into the feature
//                          gradient (fp32 atomics, HWC) and the gradient w.r.t. the sampling grid of each output position;
//   roi_grid_bwd_kernel  : AffineGridGeneratorBHWD backward (gradTheta = gradGrid^T . baseGrid, App. B1) followed by
//                          nn.BoxToAffine backward (modules/BoxToAffine.lua:96-118) -> gradient w.r.t. the boxes.
// One block per (roi, position); threads sweep the channels.  grad_out is BHWC (K, HH*WW, C) fp32.
// ---------------------------------------------------------------------------------------------------------
struct RoiCoord { int xl, yt; float wx, wy; bool ok; };
__device__ __forceinline__ RoiCoord roi_coord(float4 box, int i, int j, float img_h, float img_w, int fh, int fw, int HH, int WW) {
  // identical arithmetic to roi_taps_kernel (forward)
  const float th11 = __fdiv_rn(box.w, img_h);
  const float th13 = __fdiv_rn(__fadd_rn(__fmul_rn(box.y, 2.f), -1.f - img_h), img_h - 1.f);
  const float th22 = __fdiv_rn(box.z, img_w);
  const float th23 = __fdiv_rn(__fadd_rn(__fmul_rn(box.x, 2.f), -1.f - img_w), img_w - 1.f);
  const float yb = (float)(-1.0 + ((double)i / (double)(HH - 1)) * 2.0);
  const float xb = (float)(-1.0 + ((double)j / (double)(WW - 1)) * 2.0);
  const float gy = __fadd_rn(__fadd_rn(__fmul_rn(th11, yb), __fmul_rn(0.f, xb)), th13);
  const float gx = __fadd_rn(__fadd_rn(__fmul_rn(0.f, yb), __fmul_rn(th22, xb)), th23);
  const float xcoord = __fdiv_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)(fw - 1)), 2.f);
  const float ycoord = __fdiv_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)(fh - 1)), 2.f);
  const float xf = floorf(xcoord), yf = floorf(ycoord);
  RoiCoord c;
  c.wx = __fsub_rn(1.f, __fsub_rn(xcoord, xf));
  c.wy = __fsub_rn(1.f, __fsub_rn(ycoord, yf));
  c.xl = (int)fminf(fmaxf(xf, -2.f), (float)fw + 1.f);
  c.yt = (int)fminf(fmaxf(yf, -2.f), (float)fh + 1.f);
  c.ok = (xcoord == xcoord) && (ycoord == ycoord);
  return c;
}

__global__ void __launch_bounds__(128)
roi_pool_bwd_kernel(const __half* __restrict__ feat_hi, const __half* __restrict__ feat_lo, int C, int fh, int fw,
                    const float4* __restrict__ boxes, int nboxes, float img_h, float img_w, int HH, int WW,
                    const float* __restrict__ grad_out_bhwc, float* __restrict__ grad_feat_hwc, float2* __restrict__ grad_grid) {
  const int npos = HH * WW;
  const int item = blockIdx.x;
  const int b = item / npos, p = item % npos;
  if (b >= nboxes) return;
  const RoiCoord rc = roi_coord(boxes[b], p / WW, p % WW, img_h, img_w, fh, fw, HH, WW);
  const bool xin0 = rc.xl >= 0 && rc.xl <= fw - 1, xin1 = rc.xl + 1 >= 0 && rc.xl + 1 <= fw - 1;
  const bool yin0 = rc.yt >= 0 && rc.yt <= fh - 1, yin1 = rc.yt + 1 >= 0 && rc.yt + 1 <= fh - 1;
  const bool v[4] = {rc.ok && xin0 && yin0, rc.ok && xin1 && yin0, rc.ok && xin0 && yin1, rc.ok && xin1 && yin1};
  const int cell[4] = {rc.yt * fw + rc.xl, rc.yt * fw + rc.xl + 1, (rc.yt + 1) * fw + rc.xl, (rc.yt + 1) * fw + rc.xl + 1};
  const float omx = 1.f - rc.wx, omy = 1.f - rc.wy;
  const float w[4] = {rc.wx * rc.wy, omx * rc.wy, rc.wx * omy, omx * omy};
  const float* go = grad_out_bhwc + (size_t)item * C;
  float dot[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float g = go[c];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (v[t]) {
        const size_t o = (size_t)cell[t] * C + c;
        atomicAdd(grad_feat_hwc + o, w[t] * g);
        dot[t] += (__half2float(feat_hi[o]) + __half2float(feat_lo[o])) * g;
      }
    }
  }
  __shared__ float red[4][4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    float x = dot[t];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) red[t][threadIdx.x >> 5] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float d[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) d[t] = red[t][0] + red[t][1] + red[t][2] + red[t][3];
    // d out / d ycoord = -wx*TL - (1-wx)*TR + wx*BL + (1-wx)*BR ; d out / d xcoord = -wy*TL + wy*TR - (1-wy)*BL + (1-wy)*BR
    const float gy = (-rc.wx * d[0] + rc.wx * d[2] - omx * d[1] + omx * d[3]) * (float)(fh - 1) * 0.5f;
    const float gx = (-rc.wy * d[0] + rc.wy * d[1] - omy * d[2] + omy * d[3]) * (float)(fw - 1) * 0.5f;
    grad_grid[item] = make_float2(gy, gx);
  }
}

__global__ void roi_grid_bwd_kernel(const float2* __restrict__ grad_grid, int nboxes, int HH, int WW, float img_h, float img_w,
                                    float4* __restrict__ grad_boxes) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= nboxes) return;
  const int npos = HH * WW;
  float t00 = 0.f, t02 = 0.f, t11 = 0.f, t12 = 0.f;      // the four entries of gradTheta that BoxToAffine backward reads
  for (int p = lane; p < npos; p += 32) {
    const float2 g = grad_grid[(size_t)b * npos + p];
    const float yb = (float)(-1.0 + ((double)(p / WW) / (double)(HH - 1)) * 2.0);
    const float xb = (float)(-1.0 + ((double)(p % WW) / (double)(WW - 1)) * 2.0);
    t00 += g.x * yb; t02 += g.x;
    t11 += g.y * xb; t12 += g.y;
  }
  for (int o = 16; o > 0; o >>= 1) {
    t00 += __shfl_xor_sync(0xffffffffu, t00, o); t02 += __shfl_xor_sync(0xffffffffu, t02, o);
    t11 += __shfl_xor_sync(0xffffffffu, t11, o); t12 += __shfl_xor_sync(0xffffffffu, t12, o);
  }
  if (lane == 0) grad_boxes[b] = make_float4(t12 * (2.f / (img_w - 1.f)), t02 * (2.f / (img_h - 1.f)), t11 * (1.f / img_w), t00 * (1.f / img_h));
}

int k_roi_pool_backward(dc_ctx* ctx, const __half* feat_hi, const __half* feat_lo, int C, int fh, int fw, const float* boxes, int nboxes,
                        int img_h, int img_w, int HH, int WW, const float* grad_out_bhwc, float* grad_feat_hwc, float* grad_grid_ws,
                        float* grad_boxes, cudaStream_t st) {
  if (nboxes <= 0) return DC_OK;
  DC_REQUIRE(ctx, HH > 1 && WW > 1, DC_E_UNSUPPORTED, "roi_pool_backward: HH, WW must be > 1");
  DC_CHECK_CUDA(ctx, cudaMemsetAsync(grad_feat_hwc, 0, (size_t)fh * fw * C * sizeof(float), st));
  roi_pool_bwd_kernel<<<nboxes * HH * WW, 128, 0, st>>>(feat_hi, feat_lo, C, fh, fw, (const float4*)boxes, nboxes, (float)img_h, (float)img_w, HH, WW,
                                                       grad_out_bhwc, grad_feat_hwc, (float2*)grad_grid_ws);
  DC_CHECK_LAUNCH(ctx);
  roi_grid_bwd_kernel<<<DC_CEIL_DIV(nboxes, 4), 128, 0, st>>>((const float2*)grad_grid_ws, nboxes, HH, WW, (float)img_h, (float)img_w,
                                                              (float4*)grad_boxes);
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
