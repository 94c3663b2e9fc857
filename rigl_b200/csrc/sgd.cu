// Batched Nesterov-momentum step for ALL parameters of a model in one launch, with `mask * dense_grad` fused
// into the gradient load of the masked layers.
//
// Reference call site: imagenet_train_eval.py:355-365 -- tf.train.MomentumOptimizer(lr, momentum,
// use_nesterov=True) under the sparse wrapper, l2 regularisation on the raw weights;
// sparse_optimizers_base.py:478-485 hands it dL/dweights = mask * dL/d(mask*weights).  Per element
//   g     = (bit ? dense_grad * grad_scale : 0) + weight_decay * w          (grad_scale = 1/world under DP)
//   accum = momentum * accum + g
//   w    -= lr * (nesterov ? g + momentum * accum : accum)
// The learning rate is read from DEVICE memory, so a captured CUDA graph follows a schedule without re-capture.
// Before: one mask*grad kernel per layer (54) + five multi-tensor kernels over every parameter; now one pass that
// reads w, accum, grad (+ 1 bit) and writes w, accum: 20.125 bytes per masked weight.
#include <vector>

#include "common.cuh"

namespace rigl {

constexpr int kSgdThreads = 256;
constexpr int kSgdChunk = 8192;          // elements per block

struct SgdLayerDev {
  float* p;
  float* m;
  const float* g;
  const uint32_t* bits;
  uint32_t n;
  float wd, gscale;
  uint32_t vec_ok;       // every pointer 16-byte aligned
};

struct SgdTask { uint32_t layer, start; };

__device__ __forceinline__ void sgd_one(float& p, float& m, float g, bool on, float wd, float gscale, float lr,
                                        float mom, int nesterov) {
  const float ge = fmaf(wd, p, on ? g * gscale : 0.f);
  m = fmaf(mom, m, ge);
  p = fmaf(-lr, nesterov ? fmaf(mom, m, ge) : m, p);
}

__global__ void __launch_bounds__(kSgdThreads)
k_sgd_nesterov_batched(const SgdLayerDev* __restrict__ layers, const SgdTask* __restrict__ tasks,
                       const float* __restrict__ lr_dev, float mom, int nesterov) {
  const SgdTask t = tasks[blockIdx.x];
  const SgdLayerDev L = layers[t.layer];
  const float lr = __ldg(lr_dev);
  const uint32_t end = min(L.n, t.start + (uint32_t)kSgdChunk);
  if (L.vec_ok) {
#pragma unroll 2
    for (uint32_t e = t.start + 4 * threadIdx.x; e < end; e += 4 * kSgdThreads) {
      if (e + 3 < L.n) {
        float4 p = *reinterpret_cast<const float4*>(L.p + e);
        float4 m = *reinterpret_cast<const float4*>(L.m + e);
        const float4 g = __ldg(reinterpret_cast<const float4*>(L.g + e));
        const uint32_t nib = L.bits ? (__ldg(L.bits + (e >> 5)) >> (e & 31)) & 0xFu : 0xFu;
        sgd_one(p.x, m.x, g.x, nib & 1u, L.wd, L.gscale, lr, mom, nesterov);
        sgd_one(p.y, m.y, g.y, nib & 2u, L.wd, L.gscale, lr, mom, nesterov);
        sgd_one(p.z, m.z, g.z, nib & 4u, L.wd, L.gscale, lr, mom, nesterov);
        sgd_one(p.w, m.w, g.w, nib & 8u, L.wd, L.gscale, lr, mom, nesterov);
        *reinterpret_cast<float4*>(L.p + e) = p;
        *reinterpret_cast<float4*>(L.m + e) = m;
      } else {
        for (uint32_t j = e; j < L.n; ++j) {
          const bool on = L.bits ? (__ldg(L.bits + (j >> 5)) >> (j & 31)) & 1u : true;
          sgd_one(L.p[j], L.m[j], __ldg(L.g + j), on, L.wd, L.gscale, lr, mom, nesterov);
        }
      }
    }
  } else {
    for (uint32_t j = t.start + threadIdx.x; j < end; j += kSgdThreads) {
      const bool on = L.bits ? (__ldg(L.bits + (j >> 5)) >> (j & 31)) & 1u : true;
      sgd_one(L.p[j], L.m[j], __ldg(L.g + j), on, L.wd, L.gscale, lr, mom, nesterov);
    }
  }
}

}  // namespace rigl

struct rigl_sgd_plan {
  int n_tasks = 0;
  rigl::SgdLayerDev* d_layers = nullptr;
  rigl::SgdTask* d_tasks = nullptr;
};

using namespace rigl;

extern "C" int rigl_sgd_plan_create(const rigl_sgd_desc* params, int n_params, rigl_sgd_plan** out) {
  RIGL_REQUIRE(params && out && n_params > 0, "rigl_sgd_plan_create: bad arguments");
  std::vector<SgdLayerDev> host(n_params);
  std::vector<SgdTask> tasks;
  for (int i = 0; i < n_params; ++i) {
    const rigl_sgd_desc& d = params[i];
    RIGL_REQUIRE(d.param && d.momentum && d.grad && d.n >= 1 && d.n < (1ll << 31),
                 "rigl_sgd_plan_create: parameter %d: null tensor or bad size", i);
    RIGL_REQUIRE((reinterpret_cast<uintptr_t>(d.param) & 3) == 0 && (reinterpret_cast<uintptr_t>(d.momentum) & 3) == 0 &&
                     (reinterpret_cast<uintptr_t>(d.grad) & 3) == 0, "parameter %d: tensors must be float-aligned", i);
    SgdLayerDev& L = host[i];
    L.p = d.param; L.m = d.momentum; L.g = d.grad; L.bits = d.mask_bits; L.n = (uint32_t)d.n;
    L.wd = d.weight_decay; L.gscale = d.grad_scale;
    L.vec_ok = (aligned16(d.param) && aligned16(d.momentum) && aligned16(d.grad)) ? 1u : 0u;
    for (int64_t s = 0; s < d.n; s += kSgdChunk) tasks.push_back({(uint32_t)i, (uint32_t)s});
  }
  rigl_sgd_plan* p = new rigl_sgd_plan();
  p->n_tasks = (int)tasks.size();
  cudaError_t e = cudaMalloc(&p->d_layers, sizeof(SgdLayerDev) * n_params);
  if (e == cudaSuccess) e = cudaMalloc(&p->d_tasks, sizeof(SgdTask) * tasks.size());
  if (e == cudaSuccess) e = cudaMemcpy(p->d_layers, host.data(), sizeof(SgdLayerDev) * n_params, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(p->d_tasks, tasks.data(), sizeof(SgdTask) * tasks.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(p->d_layers); cudaFree(p->d_tasks); delete p;
    return cuda_fail(e, "rigl_sgd_plan_create");
  }
  *out = p;
  return RIGL_OK;
}

extern "C" int rigl_sgd_plan_destroy(rigl_sgd_plan* plan) {
  if (!plan) return RIGL_OK;
  cudaFree(plan->d_layers);
  cudaFree(plan->d_tasks);
  delete plan;
  return RIGL_OK;
}

extern "C" int rigl_sgd_plan_run(rigl_sgd_plan* plan, const float* lr_dev, float momentum, int nesterov,
                                 void* stream_) {
  RIGL_REQUIRE(plan && lr_dev, "rigl_sgd_plan_run: null argument");
  k_sgd_nesterov_batched<<<plan->n_tasks, kSgdThreads, 0, (cudaStream_t)stream_>>>(plan->d_layers, plan->d_tasks, lr_dev,
                                                                                 momentum, nesterov ? 1 : 0);
  RIGL_LAUNCH_CHECK("k_sgd_nesterov_batched");
  return RIGL_OK;
}
