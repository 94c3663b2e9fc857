// Fused batch-norm (+ReLU, +residual add) for NHWC bf16 activations -- HBM-bound
// streaming kernels around the masked convs.
//
// Replaces batch_norm_relu (rigl/imagenet_resnet/resnet_model.py:41-80:
// tf.layers.batch_normalization(fused=True, momentum=0.9, epsilon=1e-5) + relu) and the
// `relu(inputs + shortcut)` tail of the bottleneck block (:501).
//
//   forward : stats  : one read of y           -> per-channel mean / rstd (fp32 partials, fp64 combine)
//             apply  : read y (+residual)      -> a = relu(y*scale + shift (+ r)), bf16
//   backward: reduce : read da, y (, a)        -> dbeta = sum g, dgamma = sum g*xhat  (g = da * relu')
//                      (residual form also writes g, which IS the gradient of the shortcut)
//             apply  : read g|da, y            -> dy = scale * (g - dbeta/M - xhat*dgamma/M)
// Every kernel moves 16-byte vectors (8 channels) per thread with the channel dimension
// innermost, so global traffic is fully coalesced; reductions go registers -> smem ->
// per-block partials -> a tiny finalize kernel (fixed order: deterministic).
// Tensors that fit in L2 take the single-launch variants (k_bn_fwd_fused / k_bn_bwd_fused: the
// same three phases behind two grid barriers); rigl_bn_backward2 also sums the two gradients of a
// forked block output inside the reduce pass.
#include <cuda_bf16.h>

#include "common.cuh"

namespace rigl {

constexpr int kBnThreads = 256;

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

// Column sums of up to two row-wise quantities over a [rows, C] bf16 matrix.
// MODE 0: (y, y^2)                                   -> forward statistics
// MODE 1: (g, g*xhat), g = da * [fma(y,scale,shift) > 0 if relu]      (plain BN / BN+ReLU)
// MODE 2: (g, g*xhat), g = da * [act > 0], g written to gout          (residual form)
// partial[block][2][C] fp32.
template <int MODE, int THREADS>
__device__ __forceinline__ void colsum_rows(
    const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ da,
    const __nv_bfloat16* __restrict__ da2 /* MODE 2: optional second addend of the output gradient */,
    const __nv_bfloat16* __restrict__ act, __nv_bfloat16* __restrict__ gout, const float* __restrict__ mean,
    const float* __restrict__ rstd, const float* __restrict__ scale, const float* __restrict__ shift, int relu,
    long long row0, long long row1, int C, float* __restrict__ partial_row /* [2][C] */,
    float* red /* smem [rpi][vl][16] */, const uint8_t* __restrict__ relu_bits = nullptr) {
  const int V = C >> 3;                         // 16-byte vectors per row
  const int vl = V < THREADS ? V : THREADS;
  const int rpi = THREADS / vl;              // rows handled per block iteration
  const int r_in = threadIdx.x / vl, v0 = threadIdx.x % vl;
  for (int v = v0; v < V; v += vl) {            // (V > 256 only for C > 2048)
    float s0[8], s1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s0[i] = 0.f; s1[i] = 0.f; }
    float mu[8], rs[8], sc[8], sh[8];
    if (MODE != 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        mu[i] = mean[8 * v + i]; rs[i] = rstd[8 * v + i];
        sc[i] = scale[8 * v + i]; sh[i] = shift[8 * v + i];
      }
    }
    if (r_in < rpi) {
      // 4 rows per trip: all loads are issued before any arithmetic (memory-level parallelism)
      for (long long rb = row0 + r_in; rb < row1; rb += 4ll * rpi) {
        uint4 qy[4], qd[4], qa[4], qe[4];
        uint32_t qb[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long long r = rb + (long long)u * rpi;
          ok[u] = r < row1;
          if (ok[u]) {
            const long long off = r * C + 8 * v;
            qy[u] = *reinterpret_cast<const uint4*>(y + off);
            if (MODE != 0) qd[u] = *reinterpret_cast<const uint4*>(da + off);
            if (MODE == 2) {
              // ReLU mask of the block output: one BIT per element saved by the forward pass (1/16 of re-reading
              // the bf16 output), or the output itself
              if (relu_bits) qb[u] = relu_bits[off >> 3];
              else qa[u] = *reinterpret_cast<const uint4*>(act + off);
              if (da2) qe[u] = *reinterpret_cast<const uint4*>(da2 + off);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (!ok[u]) continue;
          const long long off = (rb + (long long)u * rpi) * C + 8 * v;
          float fy[8];
          unpack8(qy[u], fy);
          if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { s0[i] += fy[i]; s1[i] = fmaf(fy[i], fy[i], s1[i]); }
          } else {
            float g[8];
            unpack8(qd[u], g);
            if (MODE == 2) {
              if (da2) {       // the block output feeds two consumers: their gradients are summed here
                float g2[8];   // (rounded to bf16 like the separate elementwise add it replaces)
                unpack8(qe[u], g2);
#pragma unroll
                for (int i = 0; i < 8; ++i) g[i] = __bfloat162float(__float2bfloat16(g[i] + g2[i]));
              }
              if (relu_bits) {
#pragma unroll
                for (int i = 0; i < 8; ++i) g[i] = (!relu || ((qb[u] >> i) & 1u)) ? g[i] : 0.f;
              } else {
                float fa[8];
                unpack8(qa[u], fa);
#pragma unroll
                for (int i = 0; i < 8; ++i) g[i] = (!relu || fa[i] > 0.f) ? g[i] : 0.f;
              }
              *reinterpret_cast<uint4*>(gout + off) = pack8(g);
            } else if (relu) {
#pragma unroll
              for (int i = 0; i < 8; ++i) g[i] = fmaf(fy[i], sc[i], sh[i]) > 0.f ? g[i] : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              s0[i] += g[i];
              s1[i] = fmaf(g[i], (fy[i] - mu[i]) * rs[i], s1[i]);
            }
          }
        }
      }
    }
    // reduce over the rpi row-threads that share this vector lane
    if (r_in < rpi) {
      float* dst = red + ((size_t)r_in * vl + v0) * 16;
#pragma unroll
      for (int i = 0; i < 8; ++i) { dst[i] = s0[i]; dst[8 + i] = s1[i]; }
    }
    __syncthreads();
    if (r_in == 0) {
      float a0[8], a1[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { a0[i] = 0.f; a1[i] = 0.f; }
      for (int rr = 0; rr < rpi; ++rr) {
        const float* src = red + ((size_t)rr * vl + v0) * 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) { a0[i] += src[i]; a1[i] += src[8 + i]; }
      }
      float* p0 = partial_row + 8 * v;
#pragma unroll
      for (int i = 0; i < 8; ++i) { p0[i] = a0[i]; p0[C + i] = a1[i]; }
    }
    __syncthreads();
  }
}


template <int MODE>
__global__ void __launch_bounds__(kBnThreads)
k_bn_colsum(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ da,
            const __nv_bfloat16* __restrict__ da2, const __nv_bfloat16* __restrict__ act,
            __nv_bfloat16* __restrict__ gout, const float* __restrict__ mean, const float* __restrict__ rstd,
            const float* __restrict__ scale, const float* __restrict__ shift, int relu, long long rows, int C,
            long long rows_per_block, float* __restrict__ partial, const uint8_t* __restrict__ relu_bits) {
  extern __shared__ float red[];               // [rpi][V][16]
  const long long row0 = (long long)blockIdx.x * rows_per_block;
  const long long row1 = min(row0 + rows_per_block, rows);
  colsum_rows<MODE, kBnThreads>(y, da, da2, act, gout, mean, rstd, scale, shift, relu, row0, row1, C,
                                partial + (size_t)blockIdx.x * 2 * C, red, relu_bits);
}

// Sums partial[b][which][c] over b for an 8-channel slab with 1024 threads: lane = (channel c = lane & 7,
// sub-slice lane >> 3), slice = 4 * warp + sub (128 slices), so a C-channel layer runs C/8 CTAs (the reduction is
// L2-latency bound: what matters is how many loads are in flight, not bytes) and every load instruction fetches
// full 32-byte sectors.  fp64 accumulation in a fixed order: deterministic.
constexpr int kFinCh = 8, kFinSlices = 128;
__device__ __forceinline__ bool slab_sums(const float* __restrict__ partial, int nblocks, int C, int* c_out,
                                          double* s_out, double* q_out) {
  __shared__ double sm_s[32][kFinCh], sm_q[32][kFinCh];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * kFinCh + (lane & 7);
  const int slice = 4 * warp + (lane >> 3);
  double s = 0.0, q = 0.0;
  if (c < C) {
    int b = slice;
    for (; b + 3 * kFinSlices < nblocks; b += 4 * kFinSlices) {
      float vs[4], vq[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        vs[u] = __ldcg(partial + (size_t)(b + u * kFinSlices) * 2 * C + c);
        vq[u] = __ldcg(partial + (size_t)(b + u * kFinSlices) * 2 * C + C + c);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { s += (double)vs[u]; q += (double)vq[u]; }
    }
    for (; b < nblocks; b += kFinSlices) {
      s += (double)__ldcg(partial + (size_t)b * 2 * C + c);
      q += (double)__ldcg(partial + (size_t)b * 2 * C + C + c);
    }
  }
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane < 8) { sm_s[warp][lane] = s; sm_q[warp][lane] = q; }
  __syncthreads();
  if (threadIdx.x >= kFinCh) return false;
  s = 0.0; q = 0.0;
  for (int w = 0; w < 32; ++w) { s += sm_s[w][threadIdx.x]; q += sm_q[w][threadIdx.x]; }
  *s_out = s; *q_out = q; *c_out = c;
  return c < C;
}

// Forward finalize: mean, rstd, scale = gamma*rstd, shift = beta - mean*scale, running stats.
__global__ void __launch_bounds__(1024)
k_bn_finalize_fwd(const float* __restrict__ partial, int nblocks, int C, long long rows, float eps,
                  const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ mean,
                  float* __restrict__ rstd, float* __restrict__ scale, float* __restrict__ shift,
                  float* __restrict__ running_mean, float* __restrict__ running_var, float momentum) {
  int c;
  double s, q;
  if (!slab_sums(partial, nblocks, C, &c, &s, &q)) return;
  const double m = s / (double)rows;
  double var = q / (double)rows - m * m;
  if (var < 0.0) var = 0.0;
  const float r = (float)(1.0 / sqrt(var + (double)eps));
  mean[c] = (float)m;
  rstd[c] = r;
  const float sc = gamma[c] * r;
  scale[c] = sc;
  shift[c] = beta[c] - (float)m * sc;
  if (running_mean) {
    const double unbiased = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// Backward finalize: dbeta, dgamma and the per-channel affine form of the input gradient
//   dy = scale*g + P*y + Q,  P = -scale*rstd*dgamma/M,  Q = scale*(rstd*mean*dgamma/M - dbeta/M).
__global__ void __launch_bounds__(1024)
k_bn_finalize_bwd(const float* __restrict__ partial, int nblocks, int C, long long rows,
                  const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ scale,
                  float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef /*[2][C]: P, Q*/) {
  int c;
  double s, q;
  if (!slab_sums(partial, nblocks, C, &c, &s, &q)) return;
  dbeta[c] = (float)s;
  dgamma[c] = (float)q;
  const double c0 = s / (double)rows, c1 = q / (double)rows;
  const double sc = (double)scale[c], r = (double)rstd[c], m = (double)mean[c];
  coef[c] = (float)(-sc * r * c1);
  coef[C + c] = (float)(sc * (r * m * c1 - c0));
}

__device__ __forceinline__ void load8f(const float* __restrict__ p, int v, float (&o)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[2 * v], b = reinterpret_cast<const float4*>(p)[2 * v + 1];
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

// a = [relu](y*scale + shift (+ residual)).  The grid stride is a multiple of V whenever V
// divides the thread count, so each thread keeps ONE channel vector: its coefficients are
// loaded once and the loop only streams activations.
__global__ void __launch_bounds__(kBnThreads)
k_bn_apply(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ residual,
           const float* __restrict__ scale, const float* __restrict__ shift, int relu, long long nvec, int V,
           __nv_bfloat16* __restrict__ out, uint8_t* __restrict__ relu_bits) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool fixed_v = (stride % V) == 0;
  float sc[8], sh[8];
  if (fixed_v) { load8f(scale, (int)(i0 % V), sc); load8f(shift, (int)(i0 % V), sh); }
  for (long long i = i0; i < nvec; i += stride) {
    if (!fixed_v) { load8f(scale, (int)(i % V), sc); load8f(shift, (int)(i % V), sh); }
    float f[8];
    unpack8(reinterpret_cast<const uint4*>(y)[i], f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], sc[k], sh[k]);
    if (residual) {
      float r[8];
      unpack8(reinterpret_cast<const uint4*>(residual)[i], r);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] += r[k];
    }
    if (relu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
    }
    reinterpret_cast<uint4*>(out)[i] = pack8(f);
    if (relu_bits) {                    // which outputs are positive: all the backward needs of this tensor
      uint32_t m = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) m |= (f[k] > 0.f ? 1u : 0u) << k;
      relu_bits[i] = (uint8_t)m;
    }
  }
}

// dy = scale*g + P*y + Q;  g = da * relu' (recomputed from y) or the stored g.
__global__ void __launch_bounds__(kBnThreads)
k_bn_bwd_apply(const __nv_bfloat16* __restrict__ g_or_da, const __nv_bfloat16* __restrict__ y,
               const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ coef,
               int relu_recompute, long long nvec, int V, int C, __nv_bfloat16* __restrict__ dy) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool fixed_v = (stride % V) == 0;
  float sc[8], sh[8], P[8], Q[8];
  if (fixed_v) {
    const int v = (int)(i0 % V);
    load8f(scale, v, sc); load8f(shift, v, sh); load8f(coef, v, P); load8f(coef + C, v, Q);
  }
  for (long long i = i0; i < nvec; i += stride) {
    if (!fixed_v) {
      const int v = (int)(i % V);
      load8f(scale, v, sc); load8f(shift, v, sh); load8f(coef, v, P); load8f(coef + C, v, Q);
    }
    float g[8], fy[8], o[8];
    unpack8(reinterpret_cast<const uint4*>(g_or_da)[i], g);
    unpack8(reinterpret_cast<const uint4*>(y)[i], fy);
    if (relu_recompute) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (!(fmaf(fy[k], sc[k], sh[k]) > 0.f)) g[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = fmaf(sc[k], g[k], fmaf(P[k], fy[k], Q[k]));
    reinterpret_cast<uint4*>(dy)[i] = pack8(o);
  }
}


// ----------------------------------------------------------------------------
// Single-launch variants for tensors that fit in L2 (most ResNet-50 layers: 40 of its 53 BNs
// move <= 51 MB).  The three passes of a direction (column sums -> finalize -> apply) become
// three phases of ONE persistent kernel separated by grid barriers: two dependent launch
// boundaries (tail + ramp of every kernel, ~10 us of a ~35 us layer) disappear and the second pass
// re-reads the rows this CTA just summed while they are still in L2.  Same arithmetic, same
// summation order (per-CTA partials in CTA order): results are bit-identical to the 3-kernel path
// run with the same grid.  All CTAs must be co-resident (grid <= occupancy x SMs, checked by the host).
// ----------------------------------------------------------------------------
constexpr int kFusedThreads = 512;

struct BnSync { unsigned int arrived; unsigned int done; };

__device__ __forceinline__ void grid_barrier(BnSync* sync, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&sync->arrived, 1u);
    unsigned int v, spins = 0;
    long long t0 = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(&sync->arrived) : "memory");
      if (v < target && (++spins & 0x3FFu) == 0) {        // a CTA that never arrives (grid not co-resident) must
        const long long now = clock64();                   // not hang the GPU: trap after ~2 s
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000ll) __trap();
      }
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}
// The last CTA to finish re-arms the counters for the next launch (every CTA has left both barriers by then).
__device__ __forceinline__ void grid_barrier_release(BnSync* sync) {
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(&sync->done, 1u) == gridDim.x - 1) {
      sync->arrived = 0u;
      sync->done = 0u;
      __threadfence();
    }
  }
}

// Sums partial[b][which][c] over b for a 32-channel slab with 16 slices (512 threads).
__device__ __forceinline__ bool slab_sums16(const float* __restrict__ partial, int nblocks, int C, int c, int slice,
                                            double* s_out, double* q_out, double (*sm_s)[33], double (*sm_q)[33]) {
  double s = 0.0, q = 0.0;
  const int lane = threadIdx.x & 31;
  if (c < C) {
    int b = slice;
    for (; b + 3 * 16 < nblocks; b += 4 * 16) {
      float vs[4], vq[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        vs[u] = __ldcg(partial + (size_t)(b + u * 16) * 2 * C + c);
        vq[u] = __ldcg(partial + (size_t)(b + u * 16) * 2 * C + C + c);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { s += (double)vs[u]; q += (double)vq[u]; }
    }
    for (; b < nblocks; b += 16) {
      s += (double)__ldcg(partial + (size_t)b * 2 * C + c);
      q += (double)__ldcg(partial + (size_t)b * 2 * C + C + c);
    }
  }
  sm_s[slice][lane] = s;
  sm_q[slice][lane] = q;
  __syncthreads();
  if (slice == 0) {
    for (int j = 1; j < 16; ++j) { s += sm_s[j][lane]; q += sm_q[j][lane]; }
    *s_out = s;
    *q_out = q;
  }
  return slice == 0 && c < C;
}

__global__ void __launch_bounds__(kFusedThreads, 1)
k_bn_fwd_fused(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ residual,
               const float* __restrict__ gamma, const float* __restrict__ beta, long long rows, int C,
               long long rows_per_block, float eps, float momentum, int relu, float* __restrict__ running_mean,
               float* __restrict__ running_var, float* mean, float* rstd, float* scale, float* shift,
               __nv_bfloat16* __restrict__ out, float* partial, BnSync* sync, uint8_t* __restrict__ relu_bits) {
  extern __shared__ float red[];
  __shared__ double sm_s[16][33], sm_q[16][33];
  const long long row0 = min((long long)blockIdx.x * rows_per_block, rows);
  const long long row1 = min(row0 + rows_per_block, rows);
  // ---- phase 1: column sums of this CTA's rows
  colsum_rows<0, kFusedThreads>(y, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, row0, row1,
                                C, partial + (size_t)blockIdx.x * 2 * C, red);
  grid_barrier(sync, gridDim.x);
  // ---- phase 2: mean / rstd / scale / shift, 32 channels per CTA
  for (int slab = blockIdx.x; slab * 32 < C; slab += gridDim.x) {
    const int c = slab * 32 + (threadIdx.x & 31);
    double s, q;
    if (slab_sums16(partial, gridDim.x, C, c, threadIdx.x >> 5, &s, &q, sm_s, sm_q)) {
      const double m = s / (double)rows;
      double var = q / (double)rows - m * m;
      if (var < 0.0) var = 0.0;
      const float r = (float)(1.0 / sqrt(var + (double)eps));
      mean[c] = (float)m;
      rstd[c] = r;
      const float sc = gamma[c] * r;
      scale[c] = sc;
      shift[c] = beta[c] - (float)m * sc;
      if (running_mean) {
        const double unbiased = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
      }
    }
    __syncthreads();
  }
  grid_barrier(sync, 2 * gridDim.x);
  // ---- phase 3: apply to the same rows (L2-resident)
  {
    const int V = C >> 3;
    const long long i_end = row1 * V;
    const bool fixed_v = (kFusedThreads % V) == 0;
    float sc[8], sh[8];
    long long i = row0 * V + threadIdx.x;
    if (fixed_v && i < i_end) {
      const int v = (int)(i % V);
#pragma unroll
      for (int k = 0; k < 8; ++k) { sc[k] = __ldcg(scale + 8 * v + k); sh[k] = __ldcg(shift + 8 * v + k); }
    }
    // 4 vectors in flight per thread (each thread walks ~20 vectors: one at a time is latency bound)
    for (; i < i_end; i += 4ll * kFusedThreads) {
      uint4 qy[4], qr[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long idx = i + (long long)u * kFusedThreads;
        ok[u] = idx < i_end;
        if (ok[u]) {
          qy[u] = reinterpret_cast<const uint4*>(y)[idx];
          if (residual) qr[u] = reinterpret_cast<const uint4*>(residual)[idx];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (!ok[u]) continue;
        const long long idx = i + (long long)u * kFusedThreads;
        if (!fixed_v) {
          const int v = (int)(idx % V);
#pragma unroll
          for (int k = 0; k < 8; ++k) { sc[k] = __ldcg(scale + 8 * v + k); sh[k] = __ldcg(shift + 8 * v + k); }
        }
        float f[8];
        unpack8(qy[u], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], sc[k], sh[k]);
        if (residual) {
          float r[8];
          unpack8(qr[u], r);
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] += r[k];
        }
        if (relu) {
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
        }
        reinterpret_cast<uint4*>(out)[idx] = pack8(f);
        if (relu_bits) {
          uint32_t m = 0;
#pragma unroll
          for (int k = 0; k < 8; ++k) m |= (f[k] > 0.f ? 1u : 0u) << k;
          relu_bits[idx] = (uint8_t)m;
        }
      }
    }
  }
  grid_barrier_release(sync);
}

template <int MODE>
__global__ void __launch_bounds__(kFusedThreads, 1)
k_bn_bwd_fused(const __nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ da2,
               const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ act,
               const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ scale,
               const float* __restrict__ shift, long long rows, int C, long long rows_per_block, int relu,
               __nv_bfloat16* __restrict__ dy, __nv_bfloat16* gout, float* __restrict__ dgamma,
               float* __restrict__ dbeta, float* partial, float* coef, BnSync* sync,
               const uint8_t* __restrict__ relu_bits) {
  extern __shared__ float red[];
  __shared__ double sm_s[16][33], sm_q[16][33];
  const long long row0 = min((long long)blockIdx.x * rows_per_block, rows);
  const long long row1 = min(row0 + rows_per_block, rows);
  colsum_rows<MODE, kFusedThreads>(y, da, da2, act, gout, mean, rstd, scale, shift, relu, row0, row1, C,
                                   partial + (size_t)blockIdx.x * 2 * C, red, relu_bits);
  grid_barrier(sync, gridDim.x);
  for (int slab = blockIdx.x; slab * 32 < C; slab += gridDim.x) {
    const int c = slab * 32 + (threadIdx.x & 31);
    double s, q;
    if (slab_sums16(partial, gridDim.x, C, c, threadIdx.x >> 5, &s, &q, sm_s, sm_q)) {
      dbeta[c] = (float)s;
      dgamma[c] = (float)q;
      const double c0 = s / (double)rows, c1 = q / (double)rows;
      const double sc = (double)scale[c], r = (double)rstd[c], m = (double)mean[c];
      coef[c] = (float)(-sc * r * c1);
      coef[C + c] = (float)(sc * (r * m * c1 - c0));
    }
    __syncthreads();
  }
  grid_barrier(sync, 2 * gridDim.x);
  {
    const int V = C >> 3;
    const long long i_end = row1 * V;
    const bool fixed_v = (kFusedThreads % V) == 0;
    const __nv_bfloat16* g_src = (MODE == 2) ? gout : da;      // (gout was written by THIS CTA for these rows)
    float sc[8], sh[8], P[8], Q[8];
    long long i = row0 * V + threadIdx.x;
    auto load_coef = [&](int v) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        sc[k] = scale[8 * v + k]; sh[k] = shift[8 * v + k];
        P[k] = __ldcg(coef + 8 * v + k); Q[k] = __ldcg(coef + C + 8 * v + k);
      }
    };
    if (fixed_v && i < i_end) load_coef((int)(i % V));
    for (; i < i_end; i += 4ll * kFusedThreads) {
      uint4 qg[4], qy[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long idx = i + (long long)u * kFusedThreads;
        ok[u] = idx < i_end;
        if (ok[u]) {
          qg[u] = reinterpret_cast<const uint4*>(g_src)[idx];
          qy[u] = reinterpret_cast<const uint4*>(y)[idx];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (!ok[u]) continue;
        const long long idx = i + (long long)u * kFusedThreads;
        if (!fixed_v) load_coef((int)(idx % V));
        float g[8], fy[8], o[8];
        unpack8(qg[u], g);
        unpack8(qy[u], fy);
        if (MODE == 1 && relu) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (!(fmaf(fy[k], sc[k], sh[k]) > 0.f)) g[k] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = fmaf(sc[k], g[k], fmaf(P[k], fy[k], Q[k]));
        reinterpret_cast<uint4*>(dy)[idx] = pack8(o);
      }
    }
  }
  grid_barrier_release(sync);
}

static bool g_bn_fused = true;            // RIGL_BN_FUSED=0: always the 3-kernel path
static size_t g_bn_fused_max_bytes = (size_t)64 << 20;
static BnSync* g_bn_sync[16] = {};
static int g_bn_fused_grid[3] = {0, 0, 0};   // co-resident CTAs: fwd, bwd<1>, bwd<2> (0 = not probed)

static size_t fused_smem(int C) {
  const int V = C >> 3;
  const int vl = V < kFusedThreads ? V : kFusedThreads;
  const int rpi = kFusedThreads / vl;
  return (size_t)rpi * vl * 16 * sizeof(float);
}

// Grid (<= co-resident capacity) and rows per CTA of the fused kernels; 0 = use the 3-kernel path.
static int fused_plan(int which, long long rows, int C, long long* rows_per_block, BnSync** sync) {
  static bool env_read = false;
  if (!env_read) {
    if (const char* e = getenv("RIGL_BN_FUSED")) g_bn_fused = !(e[0] == '0');
    env_read = true;
  }
  if (!g_bn_fused || C % 8 || C > 4096 || (size_t)rows * C * 2 > g_bn_fused_max_bytes) return 0;
  // (measured on B200, tools/bench_bn_layer.py: the backward wins at every size <= 64 MB, the forward only
  //  for wide layers -- with few vectors per row the 148 x 512-thread grid hides less latency than 3 big grids)
  if (which == 0 && C < 512) return 0;
  const size_t smem = fused_smem(C);
  if (smem > 32 * 1024) return 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return 0;
  if (g_bn_fused_grid[which] == 0) {
    int sms = 0, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaSuccess;
    if (which == 0) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bn_fwd_fused, kFusedThreads, 32 * 1024);
    else if (which == 1) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bn_bwd_fused<1>, kFusedThreads, 32 * 1024);
    else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bn_bwd_fused<2>, kFusedThreads, 32 * 1024);
    if (e != cudaSuccess || per_sm < 1) { g_bn_fused_grid[which] = -1; return 0; }
    if (per_sm > 2) per_sm = 2;
    g_bn_fused_grid[which] = sms * per_sm;
  }
  if (g_bn_fused_grid[which] < 0) return 0;
  if (g_bn_sync[dev] == nullptr) {
    if (cudaMalloc(&g_bn_sync[dev], 3 * sizeof(BnSync)) != cudaSuccess) return 0;
    cudaMemset(g_bn_sync[dev], 0, 3 * sizeof(BnSync));
  }
  const int V = C >> 3;
  const int vl = V < kFusedThreads ? V : kFusedThreads;
  const int rpi = kFusedThreads / vl;
  long long rpb = (rows + g_bn_fused_grid[which] - 1) / g_bn_fused_grid[which];
  rpb = (rpb + rpi - 1) / rpi * rpi;
  if (rpb < rpi) rpb = rpi;
  int grid = (int)((rows + rpb - 1) / rpb);
  *rows_per_block = rpb;
  *sync = g_bn_sync[dev] + which;
  return grid;
}

static int colsum_blocks(long long rows, int C, long long* rows_per_block) {
  const int V = C >> 3;
  const int vl = V < kBnThreads ? V : kBnThreads;
  const int rpi = kBnThreads / vl;
  long long target = 148 * 6;                         // ~6 resident blocks per SM
  long long rpb = (rows + target - 1) / target;
  rpb = (rpb + rpi - 1) / rpi * rpi;
  if (rpb < rpi) rpb = rpi;
  *rows_per_block = rpb;
  return (int)((rows + rpb - 1) / rpb);
}

static size_t colsum_smem(int C) {
  const int V = C >> 3;
  const int vl = V < kBnThreads ? V : kBnThreads;
  const int rpi = kBnThreads / vl;
  return (size_t)rpi * vl * 16 * sizeof(float);
}

}  // namespace rigl

using namespace rigl;

extern "C" size_t rigl_bn_workspace_bytes(int64_t rows, int channels) {
  if (rows <= 0 || channels <= 0) return 0;
  long long rpb;
  const int nb = colsum_blocks(rows, channels, &rpb);
  return (size_t)nb * 2 * channels * sizeof(float) + 256;
}

extern "C" int rigl_bn_forward_train(const void* y, const void* residual, const float* gamma, const float* beta,
                                     int64_t rows, int channels, float eps, float momentum, int relu,
                                     float* running_mean, float* running_var, float* save_mean, float* save_rstd,
                                     float* save_scale, float* save_shift, void* out, void* ws, size_t ws_bytes,
                                     void* relu_bits, void* stream_) {
  RIGL_REQUIRE(y && gamma && beta && save_mean && save_rstd && save_scale && save_shift && out && ws,
               "rigl_bn_forward_train: null argument");
  RIGL_REQUIRE(rows > 0 && channels > 0 && channels % 8 == 0, "rigl_bn_forward_train: channels must be a multiple of 8");
  RIGL_REQUIRE(aligned16(y) && aligned16(out) && aligned16(residual) && aligned16(save_scale) && aligned16(save_shift),
               "rigl_bn_forward_train: tensors must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream_;
  long long rpb;
  {
    BnSync* sync = nullptr;
    long long frpb;
    const int grid = fused_plan(0, rows, channels, &frpb, &sync);
    if (grid > 0 && ws_bytes >= (size_t)grid * 2 * channels * sizeof(float)) {
      k_bn_fwd_fused<<<grid, kFusedThreads, fused_smem(channels), s>>>(
          (const __nv_bfloat16*)y, (const __nv_bfloat16*)residual, gamma, beta, rows, channels, frpb, eps, momentum, relu,
          running_mean, running_var, save_mean, save_rstd, save_scale, save_shift, (__nv_bfloat16*)out,
          static_cast<float*>(ws), sync, static_cast<uint8_t*>(relu_bits));
      RIGL_LAUNCH_CHECK("k_bn_fwd_fused");
      return RIGL_OK;
    }
  }
  const int nb = colsum_blocks(rows, channels, &rpb);
  if (ws_bytes < (size_t)nb * 2 * channels * sizeof(float)) {
    set_error("rigl_bn_forward_train: workspace too small");
    return RIGL_ERR_WORKSPACE;
  }
  float* partial = static_cast<float*>(ws);
  k_bn_colsum<0><<<nb, kBnThreads, colsum_smem(channels), s>>>(
      (const __nv_bfloat16*)y, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, rows, channels,
      rpb, partial, nullptr);
  RIGL_LAUNCH_CHECK("k_bn_colsum<0>");
  k_bn_finalize_fwd<<<(channels + kFinCh - 1) / kFinCh, 1024, 0, s>>>(partial, nb, channels, rows, eps, gamma, beta,
                                                                  save_mean, save_rstd, save_scale, save_shift,
                                                                  running_mean, running_var, momentum);
  RIGL_LAUNCH_CHECK("k_bn_finalize_fwd");
  const long long nvec = rows * (channels / 8);
  long long blocks = (nvec + kBnThreads - 1) / kBnThreads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_bn_apply<<<(unsigned)blocks, kBnThreads, 0, s>>>((const __nv_bfloat16*)y, (const __nv_bfloat16*)residual,
                                                     save_scale, save_shift, relu, nvec, channels / 8,
                                                     (__nv_bfloat16*)out, static_cast<uint8_t*>(relu_bits));
  RIGL_LAUNCH_CHECK("k_bn_apply");
  return RIGL_OK;
}

// Same as rigl_bn_forward_train but the column sums come from the producing conv's epilogue
// (partial[rows][2][channels], as written by rigl_masked_conv2d_fprop_bnstats): no stats pass.
extern "C" int rigl_bn_forward_train_partials(const void* y, const void* residual, const float* gamma,
                                              const float* beta, const float* partial, int partial_rows,
                                              int64_t rows, int channels, float eps, float momentum, int relu,
                                              float* running_mean, float* running_var, float* save_mean,
                                              float* save_rstd, float* save_scale, float* save_shift, void* out,
                                              void* relu_bits, void* stream_) {
  RIGL_REQUIRE(y && gamma && beta && partial && save_mean && save_rstd && save_scale && save_shift && out,
               "rigl_bn_forward_train_partials: null argument");
  RIGL_REQUIRE(rows > 0 && channels > 0 && channels % 8 == 0 && partial_rows > 0,
               "rigl_bn_forward_train_partials: bad sizes");
  cudaStream_t s = (cudaStream_t)stream_;
  k_bn_finalize_fwd<<<(channels + kFinCh - 1) / kFinCh, 1024, 0, s>>>(partial, partial_rows, channels, rows, eps, gamma,
                                                                  beta, save_mean, save_rstd, save_scale, save_shift,
                                                                  running_mean, running_var, momentum);
  RIGL_LAUNCH_CHECK("k_bn_finalize_fwd");
  const long long nvec = rows * (channels / 8);
  long long blocks = (nvec + kBnThreads - 1) / kBnThreads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_bn_apply<<<(unsigned)blocks, kBnThreads, 0, s>>>((const __nv_bfloat16*)y, (const __nv_bfloat16*)residual,
                                                     save_scale, save_shift, relu, nvec, channels / 8,
                                                     (__nv_bfloat16*)out, static_cast<uint8_t*>(relu_bits));
  RIGL_LAUNCH_CHECK("k_bn_apply");
  return RIGL_OK;
}

extern "C" int rigl_bn_apply(const void* y, const void* residual, const float* scale, const float* shift,
                             int64_t rows, int channels, int relu, void* out, void* stream_) {
  RIGL_REQUIRE(y && scale && shift && out && rows > 0 && channels > 0 && channels % 8 == 0,
               "rigl_bn_apply: bad arguments");
  const long long nvec = rows * (channels / 8);
  long long blocks = (nvec + kBnThreads - 1) / kBnThreads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_bn_apply<<<(unsigned)blocks, kBnThreads, 0, (cudaStream_t)stream_>>>(
      (const __nv_bfloat16*)y, (const __nv_bfloat16*)residual, scale, shift, relu, nvec, channels / 8,
      (__nv_bfloat16*)out, nullptr);
  RIGL_LAUNCH_CHECK("k_bn_apply");
  return RIGL_OK;
}

extern "C" int rigl_bn_backward(const void* da, const void* y, const void* act, const float* save_mean,
                                const float* save_rstd, const float* save_scale, const float* save_shift,
                                int64_t rows, int channels, int relu, void* dy, void* dresidual, float* dgamma,
                                float* dbeta, void* ws, size_t ws_bytes, void* stream_) {
  return rigl_bn_backward2(da, nullptr, y, act, save_mean, save_rstd, save_scale, save_shift, rows, channels, relu, dy,
                           dresidual, dgamma, dbeta, ws, ws_bytes, nullptr, stream_);
}

extern "C" int rigl_bn_backward2(const void* da, const void* da2, const void* y, const void* act,
                                 const float* save_mean, const float* save_rstd, const float* save_scale,
                                 const float* save_shift, int64_t rows, int channels, int relu, void* dy,
                                 void* dresidual, float* dgamma, float* dbeta, void* ws, size_t ws_bytes,
                                 const void* relu_bits, void* stream_) {
  RIGL_REQUIRE(da2 == nullptr || (dresidual != nullptr && aligned16(da2)),
               "rigl_bn_backward2: a second output gradient needs the residual form (dresidual != NULL)");
  RIGL_REQUIRE(da && y && save_mean && save_rstd && save_scale && save_shift && dy && dgamma && dbeta && ws,
               "rigl_bn_backward: null argument");
  RIGL_REQUIRE(rows > 0 && channels > 0 && channels % 8 == 0, "rigl_bn_backward: channels must be a multiple of 8");
  RIGL_REQUIRE((dresidual == nullptr) || (act != nullptr) || (relu_bits != nullptr),
               "rigl_bn_backward: the residual form needs the saved output or its ReLU bitmap");
  cudaStream_t s = (cudaStream_t)stream_;
  long long rpb;
  {
    BnSync* sync = nullptr;
    long long frpb;
    const bool res_form = dresidual != nullptr;
    const int grid = fused_plan(res_form ? 2 : 1, rows, channels, &frpb, &sync);
    if (grid > 0 && ws_bytes >= ((size_t)grid * 2 * channels + 2 * channels) * sizeof(float)) {
      float* partial = static_cast<float*>(ws);
      float* coef = partial + (size_t)grid * 2 * channels;
      if (res_form) {
        k_bn_bwd_fused<2><<<grid, kFusedThreads, fused_smem(channels), s>>>(
            (const __nv_bfloat16*)da, (const __nv_bfloat16*)da2, (const __nv_bfloat16*)y, (const __nv_bfloat16*)act,
            save_mean, save_rstd, save_scale, save_shift, rows, channels, frpb, relu, (__nv_bfloat16*)dy,
            (__nv_bfloat16*)dresidual, dgamma, dbeta, partial, coef, sync, static_cast<const uint8_t*>(relu_bits));
      } else {
        k_bn_bwd_fused<1><<<grid, kFusedThreads, fused_smem(channels), s>>>(
            (const __nv_bfloat16*)da, nullptr, (const __nv_bfloat16*)y, nullptr, save_mean, save_rstd, save_scale,
            save_shift, rows, channels, frpb, relu, (__nv_bfloat16*)dy, nullptr, dgamma, dbeta, partial, coef, sync,
            nullptr);
      }
      RIGL_LAUNCH_CHECK("k_bn_bwd_fused");
      return RIGL_OK;
    }
  }
  const int nb = colsum_blocks(rows, channels, &rpb);
  const size_t need = (size_t)nb * 2 * channels * sizeof(float) + 2 * channels * sizeof(float);
  if (ws_bytes < need) {
    set_error("rigl_bn_backward: workspace too small");
    return RIGL_ERR_WORKSPACE;
  }
  float* partial = static_cast<float*>(ws);
  float* coef = partial + (size_t)nb * 2 * channels;
  const bool residual_form = dresidual != nullptr;
  if (residual_form) {
    k_bn_colsum<2><<<nb, kBnThreads, colsum_smem(channels), s>>>(
        (const __nv_bfloat16*)y, (const __nv_bfloat16*)da, (const __nv_bfloat16*)da2, (const __nv_bfloat16*)act,
        (__nv_bfloat16*)dresidual, save_mean, save_rstd, save_scale, save_shift, relu, rows, channels, rpb, partial,
        static_cast<const uint8_t*>(relu_bits));
    RIGL_LAUNCH_CHECK("k_bn_colsum<2>");
  } else {
    k_bn_colsum<1><<<nb, kBnThreads, colsum_smem(channels), s>>>(
        (const __nv_bfloat16*)y, (const __nv_bfloat16*)da, nullptr, nullptr, nullptr, save_mean, save_rstd, save_scale,
        save_shift, relu, rows, channels, rpb, partial, nullptr);
    RIGL_LAUNCH_CHECK("k_bn_colsum<1>");
  }
  k_bn_finalize_bwd<<<(channels + kFinCh - 1) / kFinCh, 1024, 0, s>>>(partial, nb, channels, rows, save_mean, save_rstd,
                                                                  save_scale, dgamma, dbeta, coef);
  RIGL_LAUNCH_CHECK("k_bn_finalize_bwd");
  const long long nvec = rows * (channels / 8);
  long long blocks = (nvec + kBnThreads - 1) / kBnThreads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_bn_bwd_apply<<<(unsigned)blocks, kBnThreads, 0, s>>>(
      (const __nv_bfloat16*)(residual_form ? dresidual : da), (const __nv_bfloat16*)y, save_scale, save_shift, coef,
      (!residual_form && relu) ? 1 : 0, nvec, channels / 8, channels, (__nv_bfloat16*)dy);
  RIGL_LAUNCH_CHECK("k_bn_bwd_apply");
  return RIGL_OK;
}
