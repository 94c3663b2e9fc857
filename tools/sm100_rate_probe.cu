// Hardware rate probe (development tool, not product code).  Measures on every SM at once:
//   mma : cycles per tcgen05.mma (M=128, K=16 bf16, N = 64/128/256) issued back to back from
//         shared-memory operands that are already resident (no loads): the tensor-pipe + smem
//         operand-fetch floor for each tile shape, K-major and MN-major.
//   tma : bytes/clk/SM a lone TMA producer sustains into an 8-stage ring of 16 KB boxes, with the
//         source (a) L2 resident and shared by all SMs, (b) L2 resident, distinct per SM,
//         (c) streaming from HBM.
// These are the two denominators of the implicit-GEMM kernels' main loop.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/rate_probe tools/sm100_rate_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_bf16.h>
#include "../rigl_b200/csrc/tc_ptx.cuh"

using namespace rigl::ptx;

static CUtensorMap g_map;
static long long* g_fill = nullptr;
static int g_fill_boxes = 0;

static void report_fill(int grid, int boxes) {
  if (!boxes) return;
  std::vector<long long> f(grid);
  cudaMemcpy(f.data(), g_fill, grid * 8, cudaMemcpyDeviceToHost);
  std::sort(f.begin(), f.end());
  printf("        concurrent fills: %d boxes in median %lld clk -> %.2f B/clk/SM\n", boxes, f[grid / 2],
         (double)boxes * 16384.0 / (double)f[grid / 2]);
}


// Background TMA traffic: one elected lane streams `boxes` 16 KB boxes into a private 4-slot ring
// (it waits for a slot's previous load itself), so the MMA operand fetch shares the shared-memory
// port and the TMA unit with fills exactly as in a GEMM main loop.  Returns its own duration.
__device__ __forceinline__ long long fill_stream(const CUtensorMap* map, uint32_t ring, uint64_t* bars, int boxes) {
  const long long t0 = clock64();
  if (elect_one()) {
    const int row0 = (int)blockIdx.x * 4096;
    for (int i = 0; i < boxes; ++i) {
      const int s = i & 3;
      if (i >= 4) mbar_wait(smem_u32(&bars[s]), ((i >> 2) - 1) & 1);
      mbar_arrive_expect_tx(smem_u32(&bars[s]), 16384);
      tma_load_3d(ring + s * 16384, map, smem_u32(&bars[s]), 0, row0 + (i & 31) * 128, 0);
    }
    for (int i = boxes; i < boxes + 4; ++i) mbar_wait(smem_u32(&bars[i & 3]), ((i >> 2) - 1) & 1);
  }
  __syncwarp();
  return clock64() - t0;
}

// ---------------------------------------------------------------- MMA rate
template <int N, int MN_MAJOR>
__global__ void __launch_bounds__(128, 1) k_mma_rate(int iters, long long* cycles, const __grid_constant__ CUtensorMap fmap,
                                                     int fill_boxes, long long* fill_cycles) {
  __shared__ uint64_t fbar[4];
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  // two stages of A (16 KB) + B (N*128 B), zero-filled (values do not matter for timing)
  constexpr uint32_t kA = 128 * 64 * 2, kB = N * 64 * 2, kStage = kA + kB;
  for (uint32_t i = tid * 16u; i < 2 * kStage; i += 128 * 16u)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + i), "r"(0u) : "memory");
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&fbar[i]), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 2 && fill_boxes > 0) {
    const long long d = fill_stream(&fmap, base + 2 * kStage, fbar, fill_boxes);
    if (elect_one()) fill_cycles[blockIdx.x] = d;
    __syncwarp();
  }
  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, MN_MAJOR, MN_MAJOR);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint32_t st = base + (it & 1) * kStage;
        const uint64_t da = MN_MAJOR ? make_smem_desc(st, 8192, 1024) : make_smem_desc(st, 16, 1024);
        const uint64_t db = MN_MAJOR ? make_smem_desc(st + kA, 8192, 1024) : make_smem_desc(st + kA, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + (uint32_t)((it & 1) * N), da + (MN_MAJOR ? 128 * k : 2 * k), db + (MN_MAJOR ? 128 * k : 2 * k),
                    idesc, (it < 2 && k == 0) ? 0u : 1u);
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    t1 = clock64();
    if (elect_one()) cycles[blockIdx.x] = t1 - t0;
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---------------------------------------------------------------- CTA-pair MMA rate (cta_group::2, M = 256)
__device__ __forceinline__ void umma_bf16_2cta_plain(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                     uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int N, int MASKED>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_mma2_rate(int iters, long long* cycles,
                                                                                 const __grid_constant__ CUtensorMap fmap,
                                                                                 int fill_boxes, long long* fill_cycles) {
  __shared__ uint64_t fbar[4];
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool leader = cluster_ctarank() == 0;
  constexpr uint32_t kA = 128 * 64 * 2, kB = (N / 2) * 64 * 2, kStage = kA + kB;   // each CTA: its A rows + half of B
  for (uint32_t i = tid * 16u; i < 2 * kStage; i += 128 * 16u)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + i), "r"(0u) : "memory");
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&fbar[i]), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc_2cta(smem_u32(&tmem_slot), 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 2 && fill_boxes > 0) {
    const long long d = fill_stream(&fmap, base + 2 * kStage, fbar, fill_boxes);
    if (elect_one()) fill_cycles[blockIdx.x] = d;
    __syncwarp();
  }
  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(256, N, 0, 0);
    long long t0 = clock64();
    if (leader) {
      if (elect_one()) {
        for (int it = 0; it < iters; ++it) {
          const uint32_t st = base + (it & 1) * kStage;
          const uint64_t da = make_smem_desc(st, 16, 1024), db = make_smem_desc(st + kA, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (MASKED) umma_bf16_2cta(tmem + (uint32_t)((it & 1) * N), da + 2 * k, db + 2 * k, idesc, (it < 2 && k == 0) ? 0u : 1u);
            else umma_bf16_2cta_plain(tmem + (uint32_t)((it & 1) * N), da + 2 * k, db + 2 * k, idesc, (it < 2 && k == 0) ? 0u : 1u);
          }
        }
        umma_commit_2cta_mc(smem_u32(&bar), 0x3);
      }
      __syncwarp();
    }
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (elect_one()) cycles[blockIdx.x] = t1 - t0;
    __syncwarp();
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) { tc_fence_after(); tmem_dealloc_2cta(tmem, 512); }
}

template <int N, int MASKED>
static void run_mma2(int sms, long long* d_cycles, const char* name) {
  const int iters = 2000;
  const int boxes = g_fill_boxes ? (int)(iters * 4.0 * (N / 2) * 1.6 / 300.0) : 0;
  const int smem = 2 * (128 * 64 * 2 + (N / 2) * 64 * 2) + 4 * 16384 + 2048;
  cudaFuncSetAttribute(k_mma2_rate<N, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int grid = sms / 2 * 2;
  std::vector<long long> h(grid);
  for (int rep = 0; rep < 2; ++rep) {
    k_mma2_rate<N, MASKED><<<grid, 128, smem>>>(iters, d_cycles, g_map, boxes, g_fill);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
  }
  cudaMemcpy(h.data(), d_cycles, grid * 8, cudaMemcpyDeviceToHost);
  std::sort(h.begin(), h.end());
  printf("RATE mma2 %-39s clk/MMA median %7.2f  max %7.2f   (floor %d per 256xN)\n", name,
         (double)h[grid / 2] / (iters * 4.0), (double)h.back() / (iters * 4.0), N / 2);
  report_fill(grid, boxes);
}

// ---------------------------------------------------------------- TMA rate
constexpr int kMaxStages = 13;
constexpr uint32_t kBoxBytes = 128 * 64 * 2;      // 128 rows x 64 bf16

// src viewed as [rows][64] bf16; CTA b reads boxes (b * rows_per_cta + i * 128) % rows ...
__global__ void __launch_bounds__(64, 1) k_tma_rate(const __grid_constant__ CUtensorMap map, int boxes, int kStages,
                                                     int cta_stride_rows, int window_rows, long long* cycles) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t full[kMaxStages], empty[kMaxStages];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      const int row0 = (int)blockIdx.x * cta_stride_rows;
      int off = 0;
      for (int i = 0; i < boxes; ++i) {
        mbar_wait(smem_u32(&empty[stage]), phase ^ 1u);
        mbar_arrive_expect_tx(smem_u32(&full[stage]), kBoxBytes);
        tma_load_3d(base + stage * kBoxBytes, &map, smem_u32(&full[stage]), 0, row0 + off, 0);
        off += 128;
        if (off >= window_rows) off = 0;
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < boxes; ++i) {
      mbar_wait(smem_u32(&full[stage]), phase);
      __syncwarp();
      if (elect_one()) mbar_arrive(smem_u32(&empty[stage]));
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
    }
    if (elect_one()) cycles[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void report(const char* what, std::vector<long long>& h, double units_per_cta, const char* unit) {
  std::sort(h.begin(), h.end());
  const double med = (double)h[h.size() / 2], mx = (double)h.back();
  printf("RATE %-44s median %10.0f clk  max %10.0f clk  -> %8.2f %s (median)  %8.2f (max)\n", what, med, mx,
         units_per_cta / med, unit, units_per_cta / mx);
}

template <int N, int MN>
static void run_mma(int sms, long long* d_cycles, const char* name) {
  const int iters = 2000;
  const int boxes = g_fill_boxes ? (int)(iters * 4.0 * (N / 2 > 48 ? N / 2 : 48) * 1.6 / 300.0) : 0;   // ~ the MMA duration at 54 B/clk
  const int smem = 2 * (128 * 64 * 2 + N * 64 * 2) + 4 * 16384 + 2048;
  cudaFuncSetAttribute(k_mma_rate<N, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<long long> h(sms);
  for (int rep = 0; rep < 2; ++rep) {
    k_mma_rate<N, MN><<<sms, 128, smem>>>(iters, d_cycles, g_map, boxes, g_fill);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
  }
  cudaMemcpy(h.data(), d_cycles, sms * 8, cudaMemcpyDeviceToHost);
  // report clk per MMA: invert
  std::sort(h.begin(), h.end());
  printf("RATE mma %-40s clk/MMA median %7.2f  max %7.2f   (floor %d)\n", name, (double)h[sms / 2] / (iters * 4.0),
         (double)h.back() / (iters * 4.0), N / 2);
  report_fill(sms, boxes);
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long* d_cycles;
  cudaMalloc(&d_cycles, 1024 * 8);
  long long* d_fill;
  cudaMalloc(&d_fill, 1024 * 8);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(fn);
  const size_t rows = (size_t)sms * 131072;                   // 16 MB per SM
  void* src;
  cudaMalloc(&src, rows * 128);
  cudaMemset(src, 0, rows * 128);
  CUtensorMap map;
  const cuuint64_t gdim[3] = {64, rows, 1};
  const cuuint64_t gstr[2] = {128, rows * 128};
  const cuuint32_t box[3] = {64, 128, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, src, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  g_map = map; g_fill = d_fill;
  run_mma<64, 0>(sms, d_cycles, "M128 N64  K-major");
  run_mma<128, 0>(sms, d_cycles, "M128 N128 K-major");
  run_mma<256, 0>(sms, d_cycles, "M128 N256 K-major");
  run_mma<64, 1>(sms, d_cycles, "M128 N64  MN-major");
  run_mma<128, 1>(sms, d_cycles, "M128 N128 MN-major");
  run_mma<256, 1>(sms, d_cycles, "M128 N256 MN-major");
  run_mma2<256, 1>(sms, d_cycles, "pair M256 N256 (lane-mask form)");
  run_mma2<256, 0>(sms, d_cycles, "pair M256 N256 (plain form)");
  run_mma2<128, 1>(sms, d_cycles, "pair M256 N128 (lane-mask form)");
  run_mma2<128, 0>(sms, d_cycles, "pair M256 N128 (plain form)");
  run_mma2<64, 0>(sms, d_cycles, "pair M256 N64 (plain form)");
  // the same streams with concurrent TMA fills (4 x 16 KB ring per SM)
  g_fill_boxes = 1;   // sized inside the runners
  run_mma<256, 0>(sms, d_cycles, "M128 N256 K-major + fills");
  run_mma<128, 0>(sms, d_cycles, "M128 N128 K-major + fills");
  run_mma<64, 0>(sms, d_cycles, "M128 N64  K-major + fills");
  run_mma2<256, 0>(sms, d_cycles, "pair M256 N256 + fills");
  run_mma2<128, 0>(sms, d_cycles, "pair M256 N128 + fills");
  g_fill_boxes = 0;

  const int smem = kMaxStages * kBoxBytes + 2048;
  cudaFuncSetAttribute(k_tma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct Cfg { const char* name; int boxes, stride, window; } cfgs[] = {
      {"L2 same 16MB", 2048, 0, 131072},
      {"L2 own 512KB", 2048, 131072, 4096},
      {"HBM own 16MB", 1024, 131072, 131072},
  };
  const int grids[1] = {sms};
  const int depths[1] = {8};
  for (auto& c : cfgs)
    for (int gi = 0; gi < 1; ++gi)
      for (int di = 0; di < 1; ++di) {
        const int grid = grids[gi], st = depths[di];
        std::vector<long long> h(grid);
        for (int rep = 0; rep < 2; ++rep) {
          k_tma_rate<<<grid, 64, smem>>>(map, c.boxes, st, c.stride, c.window, d_cycles);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h.data(), d_cycles, grid * 8, cudaMemcpyDeviceToHost);
        char name[128];
        snprintf(name, sizeof(name), "tma %s grid=%d inflight=%dKB", c.name, grid, st * 16);
        report(name, h, (double)c.boxes * kBoxBytes, "B/clk/SM");
      }
  return 0;
}
