// Masked-weight operand packing: fp32 HWIO master weights + 1-bit mask ->
// bf16 GEMM operands in both K-major layouts, plus the per-tile survivor count
// that gates weight-tile loads.  This is where `mask * weights` of the contrib
// masked layers happens -- fused into the fp32->bf16 load of the weights, once
// per step, instead of materialising a dense fp32 masked_weights tensor.
//
// One CTA packs one 64(ci) x 64(co) tile of one tap -- exactly one entry of the survivor table, so the
// count is a plain store (no memset, no atomics).  `rigl_pack_plan_*` batches ALL layers of a model into
// ONE launch over a device-resident tile table (54 launches + 54 memsets per ResNet-50 step before).
#include <vector>

#include <cuda_bf16.h>

#include "common.cuh"
#include "conv_common.cuh"

namespace rigl {

struct PackLayerDev {
  const float* w;
  const uint32_t* bits;
  __nv_bfloat16* wf;
  __nv_bfloat16* wd;
  uint32_t* nnz;
  int cin, cout, cin_pad, cout_pad, n_tiles, k_tiles;
};

struct PackTask {
  uint32_t layer;
  uint16_t tap, nt, kt, pad;
};

// block (32, 8): four 32 x 32 sub-tiles, each transposed through shared memory
__device__ __forceinline__ void pack_tile(const PackLayerDev& L, int tap, int nt, int kt, float (*tile)[33],
                                          uint32_t* s_cnt) {
  if (threadIdx.x == 0 && threadIdx.y == 0) *s_cnt = 0;
  uint32_t cnt = 0;
#pragma unroll 1
  for (int sub = 0; sub < 4; ++sub) {
    const int co0 = nt * 64 + (sub & 1) * 32, ci0 = kt * 64 + (sub >> 1) * 32;
    __syncthreads();                 // previous sub-tile fully read (and s_cnt initialised)
    for (int r = threadIdx.y; r < 32; r += 8) {
      const int ci = ci0 + r, co = co0 + threadIdx.x;
      float v = 0.f;
      if (ci < L.cin && co < L.cout) {
        const int64_t e = ((int64_t)tap * L.cin + ci) * L.cout + co;
        const uint32_t bit = (__ldg(L.bits + (e >> 5)) >> (e & 31)) & 1u;
        v = bit ? __ldg(L.w + e) : 0.f;
        cnt += bit;
      }
      tile[r][threadIdx.x] = v;
      if (ci < L.cin && co < L.cout_pad) L.wd[((int64_t)tap * L.cin + ci) * L.cout_pad + co] = __float2bfloat16(v);
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
      const int co = co0 + r, ci = ci0 + threadIdx.x;
      if (co < L.cout && ci < L.cin_pad)
        L.wf[((int64_t)tap * L.cout + co) * L.cin_pad + ci] = __float2bfloat16(tile[threadIdx.x][r]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (threadIdx.x == 0 && cnt) atomicAdd(s_cnt, cnt);
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) L.nnz[((int64_t)tap * L.n_tiles + nt) * L.k_tiles + kt] = *s_cnt;
}

__global__ void __launch_bounds__(256) k_pack_weights(PackLayerDev L) {
  __shared__ float tile[32][33];
  __shared__ uint32_t s_cnt;
  pack_tile(L, blockIdx.z, blockIdx.x, blockIdx.y, tile, &s_cnt);
}

__global__ void __launch_bounds__(256)
k_pack_weights_batched(const PackLayerDev* __restrict__ layers, const PackTask* __restrict__ tasks) {
  __shared__ float tile[32][33];
  __shared__ uint32_t s_cnt;
  const PackTask t = tasks[blockIdx.x];
  pack_tile(layers[t.layer], t.tap, t.nt, t.kt, tile, &s_cnt);
}

static PackLayerDev make_layer(const float* w, const uint32_t* bits, int taps, int cin, int cout, void* packed) {
  const PackedLayout P = packed_layout(taps, cin, cout);
  uint8_t* base = static_cast<uint8_t*>(packed);
  PackLayerDev L;
  L.w = w; L.bits = bits;
  L.wf = reinterpret_cast<__nv_bfloat16*>(base + P.off_fprop);
  L.wd = reinterpret_cast<__nv_bfloat16*>(base + P.off_dgrad);
  L.nnz = reinterpret_cast<uint32_t*>(base + P.off_nnz);
  L.cin = cin; L.cout = cout; L.cin_pad = P.cin_pad; L.cout_pad = P.cout_pad;
  L.n_tiles = P.n_tiles; L.k_tiles = P.k_tiles;
  return L;
}

}  // namespace rigl

struct rigl_pack_plan {
  int n_layers = 0;
  int n_tasks = 0;
  rigl::PackLayerDev* d_layers = nullptr;
  rigl::PackTask* d_tasks = nullptr;
};

using namespace rigl;

extern "C" size_t rigl_packed_weights_bytes(int taps, int cin, int cout) {
  if (taps <= 0 || cin <= 0 || cout <= 0) return 0;
  return packed_layout(taps, cin, cout).total;
}

extern "C" int rigl_pack_masked_weights(const float* w_hwio, const uint32_t* mask_bits, int taps, int cin,
                                        int cout, void* packed, void* stream_) {
  RIGL_REQUIRE(w_hwio && mask_bits && packed && taps > 0 && cin > 0 && cout > 0,
               "rigl_pack_masked_weights: bad arguments");
  RIGL_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 255) == 0, "packed blob must be 256B aligned");
  cudaStream_t stream = (cudaStream_t)stream_;
  const PackLayerDev L = make_layer(w_hwio, mask_bits, taps, cin, cout, packed);
  RIGL_REQUIRE(L.k_tiles <= 65535 && taps <= 65535, "rigl_pack_masked_weights: tensor too large");
  dim3 grid(L.n_tiles, L.k_tiles, taps), block(32, 8);
  k_pack_weights<<<grid, block, 0, stream>>>(L);
  RIGL_LAUNCH_CHECK("k_pack_weights");
  return RIGL_OK;
}

extern "C" int rigl_pack_plan_create(const rigl_pack_desc* layers, int n_layers, rigl_pack_plan** out) {
  RIGL_REQUIRE(layers && out && n_layers > 0, "rigl_pack_plan_create: bad arguments");
  std::vector<PackLayerDev> host(n_layers);
  std::vector<PackTask> tasks;
  for (int l = 0; l < n_layers; ++l) {
    const rigl_pack_desc& d = layers[l];
    RIGL_REQUIRE(d.weights && d.mask_bits && d.packed && d.taps > 0 && d.cin > 0 && d.cout > 0,
                 "rigl_pack_plan_create: layer %d: bad arguments", l);
    RIGL_REQUIRE((reinterpret_cast<uintptr_t>(d.packed) & 255) == 0, "layer %d: packed blob must be 256B aligned", l);
    host[l] = make_layer(d.weights, d.mask_bits, d.taps, d.cin, d.cout, d.packed);
    RIGL_REQUIRE(host[l].n_tiles <= 65535 && host[l].k_tiles <= 65535 && d.taps <= 65535,
                 "rigl_pack_plan_create: layer %d too large", l);
    for (int t = 0; t < d.taps; ++t)
      for (int kt = 0; kt < host[l].k_tiles; ++kt)
        for (int nt = 0; nt < host[l].n_tiles; ++nt)
          tasks.push_back({(uint32_t)l, (uint16_t)t, (uint16_t)nt, (uint16_t)kt, 0});
  }
  rigl_pack_plan* p = new rigl_pack_plan();
  p->n_layers = n_layers;
  p->n_tasks = (int)tasks.size();
  cudaError_t e = cudaMalloc(&p->d_layers, sizeof(PackLayerDev) * n_layers);
  if (e == cudaSuccess) e = cudaMalloc(&p->d_tasks, sizeof(PackTask) * tasks.size());
  if (e == cudaSuccess) e = cudaMemcpy(p->d_layers, host.data(), sizeof(PackLayerDev) * n_layers, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(p->d_tasks, tasks.data(), sizeof(PackTask) * tasks.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(p->d_layers); cudaFree(p->d_tasks); delete p;
    return cuda_fail(e, "rigl_pack_plan_create");
  }
  *out = p;
  return RIGL_OK;
}

extern "C" int rigl_pack_plan_destroy(rigl_pack_plan* plan) {
  if (!plan) return RIGL_OK;
  cudaFree(plan->d_layers);
  cudaFree(plan->d_tasks);
  delete plan;
  return RIGL_OK;
}

extern "C" int rigl_pack_plan_run(rigl_pack_plan* plan, void* stream_) {
  RIGL_REQUIRE(plan != nullptr, "rigl_pack_plan_run: null plan");
  k_pack_weights_batched<<<plan->n_tasks, dim3(32, 8), 0, (cudaStream_t)stream_>>>(plan->d_layers, plan->d_tasks);
  RIGL_LAUNCH_CHECK("k_pack_weights_batched");
  return RIGL_OK;
}
