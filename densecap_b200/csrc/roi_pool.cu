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
// (38*45*32*4 B = 219 KB of the 227 KB), and serves all of its gathers from there.  A half-warp covers
// (8 lanes cover the 32 channels of one output position with 4 channels per lane: LDS.128 reads, 8-byte
// split / 16-byte fp32 stores, full 64/128-byte segments per position); the per-position tap record (32 B)
// arrives through a broadcast global load.  grid = (C/32 slices) x (148/(C/32) ROI groups), 1024 threads.
//
// The interpolation arithmetic follows the reference CPU kernel's operation order without FMA
// contraction (mul/add intrinsics), so fp32 outputs are bit-identical to the oracle.
#include "common.cuh"

constexpr int ROI_THREADS = 1024;
constexpr int ROI_MAX_POS = 64;  // HH*WW <= 64

struct __align__(16) RoiTap {
  int off[4];     // cell index (y*fw+x) of TL,TR,BL,BR or -1 when outside
  float w[4];     // (wx*wy), ((1-wx)*wy), (wx*(1-wy)), ((1-wx)*(1-wy))
};

size_t roi_taps_bytes(int max_boxes, int npos) { return (size_t)max_boxes * npos * sizeof(RoiTap) + 256; }

__global__ void roi_taps_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ nboxes_dev, int max_boxes,
                                float img_h, float img_w, int fh, int fw, int HH, int WW, RoiTap* __restrict__ taps) {
  const int npos = HH * WW;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= max_boxes * npos) return;
  const int b = idx / npos, p = idx % npos;
  const int nboxes = nboxes_dev ? min(*nboxes_dev, max_boxes) : max_boxes;
  RoiTap t;
  if (b >= nboxes) {
    t.off[0] = t.off[1] = t.off[2] = t.off[3] = -1;
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
  t.off[0] = (finite && xin0 && yin0) ? yt * fw + xl : -1;
  t.off[1] = (finite && xin1 && yin0) ? yt * fw + xl + 1 : -1;
  t.off[2] = (finite && xin0 && yin1) ? (yt + 1) * fw + xl : -1;
  t.off[3] = (finite && xin1 && yin1) ? (yt + 1) * fw + xl + 1 : -1;
  taps[idx] = t;
}

__device__ __forceinline__ float roi_blend(const RoiTap& t, float tl, float tr, float bl, float br) {
  float v = __fmul_rn(t.w[0], tl);
  v = __fadd_rn(v, __fmul_rn(t.w[1], tr));
  v = __fadd_rn(v, __fmul_rn(t.w[2], bl));
  v = __fadd_rn(v, __fmul_rn(t.w[3], br));
  return v;
}

template <int CH>
__global__ void __launch_bounds__(ROI_THREADS, 1)
roi_pool_kernel(const __half* __restrict__ feat_hi, const __half* __restrict__ feat_lo, const float* __restrict__ feats_f32,
                int C, int fh, int fw, const RoiTap* __restrict__ taps, int total_items, int npos,
                __half* __restrict__ out_hi, __half* __restrict__ out_lo, float* __restrict__ out_f32, int out_layout) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* slice = reinterpret_cast<float*>(smem_raw);                               // [fh*fw][CH]
  constexpr int QUADS = CH / 4;                 // lanes per output position (4 channels per lane)
  constexpr int POS_PER_WARP = 32 / QUADS;      // positions handled concurrently by one warp
  const int c0 = blockIdx.x * CH;
  const int cells = fh * fw;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NWARPS = ROI_THREADS / 32;
  // ---- stage the channel slice (4 channels per lane) ------------------------------------------
  for (int cell = warp * POS_PER_WARP + lane / QUADS; cell < cells; cell += NWARPS * POS_PER_WARP) {
    const int cq = (lane % QUADS) * 4;
    const size_t g = (size_t)cell * C + c0 + cq;
    float4 v;
    if (feats_f32) {
      v = *reinterpret_cast<const float4*>(feats_f32 + g);
    } else {
      const uint2 h = *reinterpret_cast<const uint2*>(feat_hi + g);
      const uint2 l = *reinterpret_cast<const uint2*>(feat_lo + g);
      const __half2 h01 = *reinterpret_cast<const __half2*>(&h.x), h23 = *reinterpret_cast<const __half2*>(&h.y);
      const __half2 l01 = *reinterpret_cast<const __half2*>(&l.x), l23 = *reinterpret_cast<const __half2*>(&l.y);
      v.x = __half2float(__low2half(h01)) + __half2float(__low2half(l01));
      v.y = __half2float(__high2half(h01)) + __half2float(__high2half(l01));
      v.z = __half2float(__low2half(h23)) + __half2float(__low2half(l23));
      v.w = __half2float(__high2half(h23)) + __half2float(__high2half(l23));
    }
    *reinterpret_cast<float4*>(slice + cell * CH + cq) = v;
  }
  __syncthreads();
  const int sub = lane / QUADS;
  const int cq = (lane % QUADS) * 4;
  // items (roi*npos + pos) are dealt in contiguous runs of POS_PER_WARP to (blockIdx.y, warp); two runs per
  // iteration so that four 16-byte tap loads are in flight per lane before the first gather is issued.
  const int stride = gridDim.y * NWARPS * POS_PER_WARP;
  auto emit = [&](int it, const uint4& ta, const uint4& tb) {
    RoiTap t;
    t.off[0] = (int)ta.x; t.off[1] = (int)ta.y; t.off[2] = (int)ta.z; t.off[3] = (int)ta.w;
    t.w[0] = __uint_as_float(tb.x); t.w[1] = __uint_as_float(tb.y); t.w[2] = __uint_as_float(tb.z); t.w[3] = __uint_as_float(tb.w);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 tl = t.off[0] >= 0 ? *reinterpret_cast<const float4*>(slice + t.off[0] * CH + cq) : z;
    const float4 tr = t.off[1] >= 0 ? *reinterpret_cast<const float4*>(slice + t.off[1] * CH + cq) : z;
    const float4 bl = t.off[2] >= 0 ? *reinterpret_cast<const float4*>(slice + t.off[2] * CH + cq) : z;
    const float4 br = t.off[3] >= 0 ? *reinterpret_cast<const float4*>(slice + t.off[3] * CH + cq) : z;
    const float v0 = roi_blend(t, tl.x, tr.x, bl.x, br.x);
    const float v1 = roi_blend(t, tl.y, tr.y, bl.y, br.y);
    const float v2 = roi_blend(t, tl.z, tr.z, bl.z, br.z);
    const float v3 = roi_blend(t, tl.w, tr.w, bl.w, br.w);
    if (out_hi) {
      __half h0, l0, h1, l1, h2, l2, h3, l3;
      split_f32(v0, h0, l0);
      split_f32(v1, h1, l1);
      split_f32(v2, h2, l2);
      split_f32(v3, h3, l3);
      const size_t o = (size_t)it * C + c0 + cq;
      uint2 ph, pl;
      ph.x = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
      ph.y = (uint32_t)__half_as_ushort(h2) | ((uint32_t)__half_as_ushort(h3) << 16);
      pl.x = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
      pl.y = (uint32_t)__half_as_ushort(l2) | ((uint32_t)__half_as_ushort(l3) << 16);
      *reinterpret_cast<uint2*>(out_hi + o) = ph;
      *reinterpret_cast<uint2*>(out_lo + o) = pl;
    }
    if (out_f32) {
      if (out_layout == 1) {
        *reinterpret_cast<float4*>(out_f32 + (size_t)it * C + c0 + cq) = make_float4(v0, v1, v2, v3);
      } else {
        const int b = it / npos, p = it % npos;
        float* o = out_f32 + ((size_t)b * C + c0 + cq) * npos + p;
        o[0] = v0; o[npos] = v1; o[2 * (size_t)npos] = v2; o[3 * (size_t)npos] = v3;
      }
    }
  };
  for (int it0 = (blockIdx.y * NWARPS + warp) * POS_PER_WARP; it0 < total_items; it0 += 2 * stride) {
    const int itA = it0 + sub, itB = itA + stride;
    const bool okA = itA < total_items, okB = itB < total_items;
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0, b0 = a0, b1 = a0;
    if (okA) { const uint4* tp = reinterpret_cast<const uint4*>(taps + itA); a0 = __ldg(tp); a1 = __ldg(tp + 1); }
    if (okB) { const uint4* tp = reinterpret_cast<const uint4*>(taps + itB); b0 = __ldg(tp); b1 = __ldg(tp + 1); }
    if (okA) emit(itA, a0, a1);
    if (okB) emit(itB, b0, b1);
  }
}

int k_roi_pool(dc_ctx* ctx, const __half* feat_hi, const __half* feat_lo, const float* feats_f32, int C, int fh, int fw,
               const float* boxes, const int32_t* nboxes_dev, int max_boxes, int img_h, int img_w, int HH, int WW,
               void* taps_ws, __half* out_hi, __half* out_lo, float* out_f32, int out_layout, cudaStream_t st) {
  if (max_boxes <= 0) return DC_OK;
  DC_REQUIRE(ctx, HH * WW <= ROI_MAX_POS && HH > 1 && WW > 1, DC_E_UNSUPPORTED, "roi_pool: HH*WW must be in (1,64]");
  DC_REQUIRE(ctx, C % 8 == 0, DC_E_UNSUPPORTED, "roi_pool: C must be a multiple of 8");
  const int npos = HH * WW;
  const int total = max_boxes * npos;
  RoiTap* taps = reinterpret_cast<RoiTap*>(taps_ws);
  roi_taps_kernel<<<DC_CEIL_DIV(total, 256), 256, 0, st>>>((const float4*)boxes, nboxes_dev, max_boxes, (float)img_h, (float)img_w, fh, fw,
                                                          HH, WW, taps);
  DC_CHECK_LAUNCH(ctx);
  const size_t limit = 227 * 1024;
  int ch = 32;
  while (ch > 8 && ((size_t)fh * fw * ch * 4 > limit || C % ch != 0)) ch >>= 1;
  const size_t smem = (size_t)fh * fw * ch * 4;
  DC_REQUIRE(ctx, smem <= limit && C % ch == 0, DC_E_UNSUPPORTED, "roi_pool: feature map %dx%d too large for the shared-memory slice", fh, fw);
  const int slices = C / ch;
  int groups = 148 / slices;
  if (groups < 1) groups = 1;
  dim3 grid(slices, groups);
#define ROI_LAUNCH(CHV)                                                                                              \
  do {                                                                                                               \
    cudaError_t e = cudaFuncSetAttribute(roi_pool_kernel<CHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "roi_pool smem attribute: %s", cudaGetErrorString(e));    \
    roi_pool_kernel<CHV><<<grid, ROI_THREADS, smem, st>>>(feat_hi, feat_lo, feats_f32, C, fh, fw, taps, total, npos,   \
                                                          out_hi, out_lo, out_f32, out_layout);                          \
  } while (0)
  if (ch == 32) ROI_LAUNCH(32);
  else if (ch == 16) ROI_LAUNCH(16);
  else ROI_LAUNCH(8);
#undef ROI_LAUNCH
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
