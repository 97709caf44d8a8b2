// nn.BilinearRoiPooling (densecap/modules/BilinearRoiPooling.lua:35-67,84-91) fused into one kernel:
//   BoxToAffine (modules/BoxToAffine.lua:69-93)  ->  stnbhwd AffineGridGeneratorBHWD(HH,WW)
//   -> BatchBilinearSamplerBHWD (modules/BatchBilinearSamplerBHWD.lua:104-122; stnbhwd
//   BilinearSamplerBHWD_updateOutput) with the two nn.Transpose copies
//   removed by layout choice (features NHWC in, BHWC out).
//
// B200 design: the op is HBM-write bound (K*HH*WW*C*4 B out, 100 MB at K=1000) but every output
// element gathers 4 feature values, i.e. 4x read amplification if served from L2.  Each CTA
// therefore owns one 32-channel slice of the whole feature map, staged ONCE into shared memory
// (38*45*32*4 B = 219 KB of the 227 KB), and serves all of its gathers from there; a warp
// processes one ROI at a time with lane = channel, so shared reads are conflict-free and every
// global store is a full 64/128-byte segment.  grid = (C/32 slices, ~148/(C/32) ROI groups).
//
// The interpolation arithmetic follows the reference CPU kernel's operation order without FMA
// contraction (mul/add intrinsics), so fp32 outputs are bit-identical to the oracle.
#include "common.cuh"

constexpr int ROI_THREADS = 256;
constexpr int ROI_MAX_POS = 64;  // HH*WW <= 64

struct RoiTap {
  int off[4];     // cell index (y*fw+x) of TL,TR,BL,BR or -1 when outside
  float w[4];     // (wx*wy), ((1-wx)*wy), (wx*(1-wy)), ((1-wx)*(1-wy))
};

template <int CH>
__global__ void __launch_bounds__(ROI_THREADS, 1)
roi_pool_kernel(const __half* __restrict__ feat_hi, const __half* __restrict__ feat_lo, const float* __restrict__ feats_f32,
                int C, int fh, int fw, const float4* __restrict__ boxes, const int32_t* __restrict__ nboxes_dev, int max_boxes,
                float img_h, float img_w, int HH, int WW, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                float* __restrict__ out_f32, int out_layout) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* slice = reinterpret_cast<float*>(smem_raw);                               // [fh*fw][CH]
  RoiTap* taps = reinterpret_cast<RoiTap*>(slice + (size_t)fh * fw * CH);          // [warps][HH*WW]
  constexpr int LANES_PER_CELL = CH;            // CH in {32,16,8}
  constexpr int CELLS_PER_WARP = 32 / CH;
  const int c0 = blockIdx.x * CH;
  const int cells = fh * fw;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = ROI_THREADS / 32;
  // ---- stage the channel slice --------------------------------------------------------------
  for (int cell = warp * CELLS_PER_WARP + lane / LANES_PER_CELL; cell < cells; cell += nwarps * CELLS_PER_WARP) {
    const int ch = lane % LANES_PER_CELL;
    const size_t g = (size_t)cell * C + c0 + ch;
    const float v = feats_f32 ? feats_f32[g] : join_f32(feat_hi[g], feat_lo[g]);
    slice[cell * CH + ch] = v;
  }
  __syncthreads();
  const int nboxes = nboxes_dev ? min(*nboxes_dev, max_boxes) : max_boxes;
  const int npos = HH * WW;
  RoiTap* my = taps + warp * npos;
  const int sub = lane / LANES_PER_CELL;        // which of the CELLS_PER_WARP concurrent positions
  const int ch = lane % LANES_PER_CELL;
  // ROIs are dealt round-robin to (blockIdx.y, warp)
  for (int b = blockIdx.y * nwarps + warp; b < max_boxes; b += gridDim.y * nwarps) {
    const bool live = b < nboxes;
    if (live) {
      const float4 box = boxes[b];
      // BoxToAffine.lua:87-90
      const float th11 = __fdiv_rn(box.w, img_h);
      const float th13 = __fdiv_rn(__fadd_rn(__fmul_rn(box.y, 2.f), -1.f - img_h), img_h - 1.f);
      const float th22 = __fdiv_rn(box.z, img_w);
      const float th23 = __fdiv_rn(__fadd_rn(__fmul_rn(box.x, 2.f), -1.f - img_w), img_w - 1.f);
      for (int p = lane; p < npos; p += 32) {
        const int i = p / WW, j = p % WW;
        // AffineGridGeneratorBHWD base grid: -1 + (i)/(HH-1)*2 evaluated in double, stored as float
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
        RoiTap t;
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
        my[p] = t;
      }
    }
    __syncwarp();
    for (int p0 = 0; p0 < npos; p0 += CELLS_PER_WARP) {
      const int p = p0 + sub;
      if (p < npos) {
        float v = 0.f;
        if (live) {
          const RoiTap t = my[p];
          const float tl = t.off[0] >= 0 ? slice[t.off[0] * CH + ch] : 0.f;
          const float tr = t.off[1] >= 0 ? slice[t.off[1] * CH + ch] : 0.f;
          const float bl = t.off[2] >= 0 ? slice[t.off[2] * CH + ch] : 0.f;
          const float br = t.off[3] >= 0 ? slice[t.off[3] * CH + ch] : 0.f;
          v = __fmul_rn(t.w[0], tl);
          v = __fadd_rn(v, __fmul_rn(t.w[1], tr));
          v = __fadd_rn(v, __fmul_rn(t.w[2], bl));
          v = __fadd_rn(v, __fmul_rn(t.w[3], br));
        }
        if (out_hi) {
          __half h, l;
          split_f32(v, h, l);
          const size_t o = ((size_t)b * npos + p) * C + c0 + ch;
          out_hi[o] = h;
          out_lo[o] = l;
        }
        if (out_f32) {
          const size_t o = (out_layout == 1) ? ((size_t)b * npos + p) * C + c0 + ch
                                             : ((size_t)b * C + c0 + ch) * npos + p;
          out_f32[o] = v;
        }
      }
    }
    __syncwarp();
  }
}

int k_roi_pool(dc_ctx* ctx, const __half* feat_hi, const __half* feat_lo, const float* feats_f32, int C, int fh, int fw,
               const float* boxes, const int32_t* nboxes_dev, int max_boxes, int img_h, int img_w, int HH, int WW,
               __half* out_hi, __half* out_lo, float* out_f32, int out_layout, cudaStream_t st) {
  if (max_boxes <= 0) return DC_OK;
  DC_REQUIRE(ctx, HH * WW <= ROI_MAX_POS && HH > 1 && WW > 1, DC_E_UNSUPPORTED, "roi_pool: HH*WW must be in (1,64]");
  DC_REQUIRE(ctx, C % 8 == 0, DC_E_UNSUPPORTED, "roi_pool: C must be a multiple of 8");
  const size_t tap_bytes = (size_t)(ROI_THREADS / 32) * HH * WW * sizeof(RoiTap);
  const size_t limit = 227 * 1024;
  int ch = 32;
  while (ch > 8 && ((size_t)fh * fw * ch * 4 + tap_bytes > limit || C % ch != 0)) ch >>= 1;
  const size_t smem = (size_t)fh * fw * ch * 4 + tap_bytes;
  DC_REQUIRE(ctx, smem <= limit && C % ch == 0, DC_E_UNSUPPORTED, "roi_pool: feature map %dx%d too large for the shared-memory slice", fh, fw);
  const int slices = C / ch;
  int groups = 148 / slices;
  if (groups < 1) groups = 1;
  groups = min(groups, DC_CEIL_DIV(max_boxes, ROI_THREADS / 32));
  dim3 grid(slices, groups);
#define ROI_LAUNCH(CHV)                                                                                              \
  do {                                                                                                               \
    cudaError_t e = cudaFuncSetAttribute(roi_pool_kernel<CHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return dc_set_error(ctx, DC_E_CUDA, "roi_pool smem attribute: %s", cudaGetErrorString(e));    \
    roi_pool_kernel<CHV><<<grid, ROI_THREADS, smem, st>>>(feat_hi, feat_lo, feats_f32, C, fh, fw, (const float4*)boxes, \
                                                          nboxes_dev, max_boxes, (float)img_h, (float)img_w, HH, WW,     \
                                                          out_hi, out_lo, out_f32, out_layout);                          \
  } while (0)
  if (ch == 32) ROI_LAUNCH(32);
  else if (ch == 16) ROI_LAUNCH(16);
  else ROI_LAUNCH(8);
#undef ROI_LAUNCH
  DC_CHECK_LAUNCH(ctx);
  return DC_OK;
}
