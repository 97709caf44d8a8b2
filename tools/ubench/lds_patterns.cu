// Micro-benchmark: shared-memory wavefronts per LDS.128 / LDS.64 for the address patterns the ROI-pooling gather can use.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_patterns lds_patterns.cu ; prints cycles per warp-level load.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ROWS = 1700;          // 128-byte rows
constexpr int ITERS = 2048;
constexpr int NW = 16;              // warps

__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int VEC>   // 16 or 8 bytes per lane
__global__ void __launch_bounds__(NW * 32, 1) bench(int pattern, int misalign, long long* cycles_out, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* base = reinterpret_cast<float*>(smem + misalign);
  for (int i = threadIdx.x; i < ROWS * 32; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // 4 address sets per lane (byte offsets), fixed over the loop
  uint32_t off[4];
  for (int k = 0; k < 4; ++k) {
    auto row = [&](int unit) { return hash(warp * 977 + k * 131 + unit * 7919 + 12345) % (ROWS - 8); };
    uint32_t o = 0;
    switch (pattern) {
      case 0: o = row(0) * 128 + lane * 16; break;                                              // 512 B contiguous
      case 1: o = row(lane / 8) * 128 + (lane % 8) * 16; break;                                 // quarter-warp = one random row
      case 2: o = row(lane / 4) * 128 + (lane % 4) * 32 + 16 * ((lane / 4) & 1); break;         // 8 rows, parity-swizzled halves
      case 3: o = row(lane / 4) * 128 + (lane % 4) * 32; break;                                 // 8 rows, same 16 banks
      case 4: o = row(0) * 128 + (lane % 4) * 32 + 16 * ((lane / 4) & 1); break;                // one row, broadcast
      case 5: o = (row(0) + lane / 8) * 128 + (lane % 8) * 16 + 16; break;                      // contiguous but shifted by 16 B
      case 6: o = row(lane / 4) * 128 + (lane % 4) * 16 + 64 * ((lane / 4) & 1); break;         // 8 rows, each 64 B contiguous; odd positions upper half
      case 7: o = row(lane / 8) * 128 + ((lane % 8) ^ (lane / 8)) * 16; break;                  // quarter rows, xor-permuted chunks
      case 8: o = row(lane / 16) * 128 + (lane % 16) * 8; break;                                // LDS.64: half-warp = one row
      case 9: o = row(lane / 8) * 128 + (lane % 8) * 8 + 64 * ((lane / 8) & 1); break;          // LDS.64: 4 rows x 64 B
      case 10: o = row(lane / 2) * 128 + (lane % 2) * 16 + 32 * ((lane / 2) & 3); break;        // 16 rows x 32 B
      case 11: o = row(lane / 8) * 128 + (lane % 8) * 16; if ((lane / 8) & 1) o = row(lane / 8 - 1) * 128 + 128 + (lane % 8) * 16; break;   // rows adjacent in pairs
    }
    off[k] = o;
  }
  float acc = 0.f;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(base);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t a = sbase + off[k] + ((uint32_t)(it & 7) << 7);
      if (VEC == 16) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
        acc += v.x + v.y + v.z + v.w;
      } else {
        float2 v;
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
        acc += v.x + v.y;
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) *cycles_out = t1 - t0;
  sink[threadIdx.x] = acc;
}

int main() {
  long long* d_c; float* d_s;
  cudaMalloc(&d_c, 8); cudaMalloc(&d_s, 4096 * 4);
  const size_t smem = ROWS * 128 + 256;
  cudaFuncSetAttribute(bench<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(bench<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const char* names[] = {"LDS.128 512B contiguous", "LDS.128 quarter=row (old ROI layout)", "LDS.128 8 rows x 4 lanes, parity swizzle (new ROI layout)",
                         "LDS.128 8 rows x 4 lanes, same 16 banks", "LDS.128 one row broadcast", "LDS.128 contiguous +16B shift",
                         "LDS.128 8 rows x 64B contiguous halves", "LDS.128 quarter=row xor-permuted", "LDS.64 half=row", "LDS.64 4 rows x 64B",
                         "LDS.128 16 rows x 32B", "LDS.128 quarter=row, rows adjacent in pairs"};
  for (int p = 0; p < 12; ++p) {
    for (int rep = 0; rep < 2; ++rep) {
      const bool v8 = (p == 8 || p == 9);
      if (v8) bench<8><<<1, NW * 32, smem>>>(p, 0, d_c, d_s); else bench<16><<<1, NW * 32, smem>>>(p, 0, d_c, d_s);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("pattern %d: %s\n", p, cudaGetErrorString(e)); return 1; }
      long long c; cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
      if (rep == 1) printf("pattern %2d  %-62s %.2f cycles per warp-load\n", p, names[p], (double)c / ((double)ITERS * 4 * NW));
    }
  }
  return 0;
}
