// Hardware probe (development tool, not product code): does a tcgen05.mma shared-memory
// descriptor whose start address is shifted by whole 128-byte rows (NOT 1024-byte aligned)
// still address a SWIZZLE_128B tile correctly?  If it does, the 9 taps of a 3x3 filter can be
// fed from ONE halo tile in shared memory instead of 9 shifted TMA boxes (L2 -> SM traffic / 9).
//
//   K-major  A: rows = pixels (M), 128 B = 64 channels (K).   shift = +s rows along M.
//   MN-major A: rows = pixels (K), 128 B = 64 channels (M).   shift = +s rows along K.
// Variant 0 leaves the descriptor's base-offset field 0, variant 1 sets it to (start>>7)&7.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/umma_shift_probe tools/umma_shift_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../rigl_b200/csrc/tc_ptx.cuh"

using namespace rigl::ptx;

constexpr int kRows = 320;          // A rows available (pixels)
constexpr int kShifts = 12;
__constant__ int c_shifts[kShifts];

__host__ __device__ inline float a_val(int r, int k) { return (float)(((r * 7 + k * 3) % 17) - 8); }
__host__ __device__ inline float b_val(int n, int k) { return (float)(((n * 5 + k * 11) % 13) - 6); }

__device__ inline void st_swz(unsigned char* base, int row, int col, float v) {   // col in [0,64)
  const int chunk = col >> 3;
  const size_t off = (size_t)row * 128 + (size_t)((chunk ^ (row & 7)) << 4) + (size_t)(col & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(base + off) = __float2bfloat16(v);
}

// mode 0: K-major A [pixels=M rows][64 ch=K], B K-major [64 n][64 k]        D[m][n] = sum_k A[s+m][k] B[n][k]
// mode 1: MN-major A: rows = K index (pixels), 128 B = 64 M channels, 2 atoms (M=128): atom a at a*kRows*128
//         B MN-major: rows = K index (pixels), 128 B = 64 n.                 D[m][n] = sum_p A[s+p][m] B[s+p][n]
__global__ void __launch_bounds__(128) k_probe(int mode, float* out /*[kShifts][2][128][64]*/) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* a0 = smem;                               // kRows*128 (x2 atoms in mode 1)
  unsigned char* b0 = smem + 2 * kRows * 128;             // mode 0: 64 rows; mode 1: kRows rows
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < kRows * 64; i += 128) {
    const int r = i / 64, c = i % 64;
    if (mode == 0) {
      st_swz(a0, r, c, a_val(r, c));
      if (r < 64) st_swz(b0, r, c, b_val(r, c));
    } else {
      st_swz(a0, r, c, a_val(r, c));                          // channels 0..63
      st_swz(a0 + kRows * 128, r, c, a_val(r, c + 64));       // channels 64..127
      st_swz(b0, r, c, b_val(c, r));                          // B[n=c][k=r]
    }
  }
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  uint32_t parity = 0;

  for (int si = 0; si < kShifts; ++si) {
    for (int variant = 0; variant < 2; ++variant) {
      const int s = c_shifts[si];
      if (tid == 0) {
        if (mode == 0) {
          const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t sa = smem_u32(a0) + (uint32_t)s * 128u + (uint32_t)kk * 32u;
            uint64_t da = make_smem_desc(sa, 16, 1024);
            if (variant) da |= (uint64_t)((sa >> 7) & 7u) << 49;
            const uint64_t db = make_smem_desc(smem_u32(b0) + (uint32_t)kk * 32u, 16, 1024);
            umma_bf16(tmem, da, db, idesc, kk > 0);
          }
        } else {
          const uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
          for (int kk = 0; kk < 4; ++kk) {          // K = 64 pixels = 4 steps of 16 rows
            const uint32_t sa = smem_u32(a0) + (uint32_t)(s + kk * 16) * 128u;
            const uint32_t sb = smem_u32(b0) + (uint32_t)(s + kk * 16) * 128u;
            uint64_t da = make_smem_desc(sa, kRows * 128, 1024);
            uint64_t db = make_smem_desc(sb, kRows * 128, 1024);
            if (variant) { da |= (uint64_t)((sa >> 7) & 7u) << 49; db |= (uint64_t)((sb >> 7) & 7u) << 49; }
            umma_bf16(tmem, da, db, idesc, kk > 0);
          }
        }
        umma_commit(smem_u32(&bar));
      }
      mbar_wait(smem_u32(&bar), parity);
      parity ^= 1;
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), r0);
      tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + 32, r1);
      tmem_ld_wait();
      float* o = out + (((size_t)si * 2 + variant) * 128 + (warp * 32 + lane)) * 64;
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
  const int shifts[kShifts] = {0, 1, 2, 3, 5, 7, 8, 9, 29, 30, 58, 59};
  cudaMemcpyToSymbol(c_shifts, shifts, sizeof(shifts));
  const size_t n_out = (size_t)kShifts * 2 * 128 * 64;
  float* d_out;
  cudaMalloc(&d_out, n_out * 4);
  std::vector<float> h(n_out);
  const int smem = 3 * kRows * 128 + 2048;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(d_out, 0xFF, n_out * 4);
    k_probe<<<1, 128, smem>>>(mode, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h.data(), d_out, n_out * 4, cudaMemcpyDeviceToHost);
    for (int si = 0; si < kShifts; ++si)
      for (int v = 0; v < 2; ++v) {
        const int s = shifts[si];
        int bad = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            float ref = 0.f;
            if (mode == 0) for (int k = 0; k < 64; ++k) ref += a_val(s + m, k) * b_val(n, k);
            else for (int p = 0; p < 64; ++p) ref += a_val(s + p, m) * b_val(n, s + p);
            if (h[(((size_t)si * 2 + v) * 128 + m) * 64 + n] != ref) ++bad;
          }
        printf("PROBE mode=%s shift=%2d base_offset=%s mismatches=%d/8192\n", mode ? "MN" : "K ", s,
               v ? "(addr>>7)&7" : "0", bad);
      }
  }
  return 0;
}
