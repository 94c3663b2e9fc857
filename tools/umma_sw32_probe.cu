// Hardware probe (development tool, not product code; NOT YET RUN -- written at the end of round 1 for
// the round-2 stem kernels).  Questions it answers on a B200:
//   K : K-major operand with 32-byte rows (16 bf16 of K), SWIZZLE_32B, SBO = 256 B: does a descriptor
//       whose start address is shifted by whole 32-byte rows still address the swizzled tile?  (the
//       128-byte-row / SWIZZLE_128B case is proven by umma_shift_probe.cu)
//   MN: MN-major A operand with 32-byte rows (16 channels per row, one row per K index), SWIZZLE_32B,
//       M = 128 as EIGHT 16-element atoms that are LBO bytes apart: LBO = 32 B puts atom i one row
//       further down, i.e. the eight atoms are eight horizontally neighbouring filter taps of a
//       space-to-depth stem (wgrad).  B is an ordinary MN-major SWIZZLE_128B tile.
// Each case prints the number of mismatching outputs (0 = the layout assumption holds).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/sw32_probe tools/umma_sw32_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../rigl_b200/csrc/tc_ptx.cuh"

using namespace rigl::ptx;

constexpr int kRows = 512;          // 32-byte rows available in the A region (16 KB)
constexpr int kCases = 12;
__constant__ int c_shift[kCases];

__host__ __device__ inline float a_val(int r, int k) { return (float)(((r * 7 + k * 3) % 17) - 8); }
__host__ __device__ inline float b_val(int n, int k) { return (float)(((n * 5 + k * 11) % 13) - 6); }

// descriptor with an explicit swizzle mode: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t swz) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)swz << 61;
  return d;
}

// element (row, col<16) of a 32-byte-row tile under SWIZZLE_32B (16-byte chunk ^= address bit 7)
__device__ inline void st_sw32(unsigned char* base, int row, int col, float v) {
  const int chunk = col >> 3;
  const size_t off = (size_t)row * 32 + (size_t)((chunk ^ ((row >> 2) & 1)) << 4) + (size_t)(col & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(base + off) = __float2bfloat16(v);
}
// element (row, col<64) of a 128-byte-row tile under SWIZZLE_128B
__device__ inline void st_sw128(unsigned char* base, int row, int col, float v) {
  const int chunk = col >> 3;
  const size_t off = (size_t)row * 128 + (size_t)((chunk ^ (row & 7)) << 4) + (size_t)(col & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(base + off) = __float2bfloat16(v);
}

// mode 0: K-major.  A[row][16] (SW32), B[64 n][16] (SW32).           D[m][n] = sum_k A[s+m][k] * B[n][k]
// mode 1: MN-major. A row p holds 16 channels (SW32); M index m = atom*16 + e reads row (s + p + atom), element e.
//                   B row p holds 64 n (SW128).                       D[m][n] = sum_{p<16} A[s+p+m/16][m%16] * Bt[p][n]
// variant (mode 1 only): 0 = LBO 32 B / SBO 256 B as derived in DESIGN.md, 1 = the two swapped.
__global__ void __launch_bounds__(128) k_probe(int mode, float* out /*[kCases][2][128][64]*/) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* a0 = smem;                    // kRows * 32 B = 16 KB
  unsigned char* b0 = smem + kRows * 32;       // mode 0: 64 rows x 32 B; mode 1: 16 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < kRows * 16; i += 128) st_sw32(a0, i / 16, i % 16, a_val(i / 16, i % 16));
  if (mode == 0) {
    for (int i = tid; i < 64 * 16; i += 128) st_sw32(b0, i / 16, i % 16, b_val(i / 16, i % 16));
  } else {
    for (int i = tid; i < 16 * 64; i += 128) st_sw128(b0, i / 64, i % 64, b_val(i % 64, i / 64));   // Bt[p][n] = b_val(n, p)
  }
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  uint32_t parity = 0;
  for (int ci = 0; ci < kCases; ++ci) {
    for (int variant = 0; variant < 2; ++variant) {
      const int s = c_shift[ci];
      if (warp == 1) {
        if (elect_one()) {
          const uint32_t sa = smem_u32(a0) + (uint32_t)s * 32u;
          if (mode == 0) {
            const uint64_t da = desc(sa, 16, variant ? 512 : 256, 6);      // variant 1: wrong SBO on purpose (must fail)
            const uint64_t db = desc(smem_u32(b0), 16, 256, 6);
            umma_bf16(tmem, da, db, make_idesc_bf16(128, 64, 0, 0), 0u);
          } else {
            const uint64_t da = variant ? desc(sa, 256, 32, 6) : desc(sa, 32, 256, 6);
            const uint64_t db = desc(smem_u32(b0), 8192, 1024, 2);
            umma_bf16(tmem, da, db, make_idesc_bf16(128, 64, 1, 1), 0u);
          }
          umma_commit(smem_u32(&bar));
        }
        __syncwarp();
      }
      mbar_wait(smem_u32(&bar), parity);
      parity ^= 1;
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), r0);
      tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + 32, r1);
      tmem_ld_wait();
      float* o = out + (((size_t)ci * 2 + variant) * 128 + (warp * 32 + lane)) * 64;
#pragma unroll
      for (int j = 0; j < 32; ++j) { o[j] = __uint_as_float(r0[j]); o[32 + j] = __uint_as_float(r1[j]); }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  const int shifts[kCases] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 127, 128, 131};
  cudaMemcpyToSymbol(c_shift, shifts, sizeof(shifts));
  const size_t n_out = (size_t)kCases * 2 * 128 * 64;
  float* d_out;
  cudaMalloc(&d_out, n_out * 4);
  std::vector<float> h(n_out);
  const int smem = kRows * 32 + 4096 + 2048;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(d_out, 0xFF, n_out * 4);
    k_probe<<<1, 128, smem>>>(mode, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h.data(), d_out, n_out * 4, cudaMemcpyDeviceToHost);
    for (int ci = 0; ci < kCases; ++ci)
      for (int v = 0; v < 2; ++v) {
        const int s = shifts[ci];
        int bad = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            float ref = 0.f;
            if (mode == 0) for (int k = 0; k < 16; ++k) ref += a_val(s + m, k) * b_val(n, k);
            else for (int p = 0; p < 16; ++p) ref += a_val(s + p + m / 16, m % 16) * b_val(n, p);
            if (h[(((size_t)ci * 2 + v) * 128 + m) * 64 + n] != ref) ++bad;
          }
        printf("SW32PROBE mode=%s shift=%3d %s mismatches=%d/8192\n", mode ? "MN(8 atoms)" : "K ", s,
               mode ? (v ? "LBO=256,SBO=32 " : "LBO=32,SBO=256 ") : (v ? "SBO=512(control, must fail)" : "SBO=256"), bad);
      }
  }
  return 0;
}
