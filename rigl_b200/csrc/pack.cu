// Masked-weight operand packing: fp32 HWIO master weights + 1-bit mask ->
// bf16 GEMM operands in both K-major layouts, plus the per-tile survivor count
// that gates weight-tile loads.  This is where `mask * weights` of the contrib
// masked layers happens -- fused into the fp32->bf16 load of the weights, once
// per step, instead of materialising a dense fp32 masked_weights tensor.
#include <cuda_bf16.h>

#include "common.cuh"
#include "conv_common.cuh"

namespace rigl {

// block (32, 8); blockIdx = (co tile of 32, ci tile of 32, tap)
__global__ void k_pack_weights(const float* __restrict__ w, const uint32_t* __restrict__ bits, int cin,
                               int cout, int cin_pad, int cout_pad, __nv_bfloat16* __restrict__ wf,
                               __nv_bfloat16* __restrict__ wd, uint32_t* __restrict__ nnz, int n_tiles,
                               int k_tiles) {
  __shared__ float tile[32][33];
  __shared__ uint32_t s_cnt;
  const int tap = blockIdx.z;
  const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * 32;
  if (threadIdx.x == 0 && threadIdx.y == 0) s_cnt = 0;
  __syncthreads();
  uint32_t cnt = 0;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int ci = ci0 + r, co = co0 + threadIdx.x;
    float v = 0.f;
    if (ci < cin && co < cout) {
      const int64_t e = ((int64_t)tap * cin + ci) * cout + co;
      const uint32_t bit = (bits[e >> 5] >> (e & 31)) & 1u;
      v = bit ? w[e] : 0.f;
      cnt += bit;
    }
    tile[r][threadIdx.x] = v;
    if (ci < cin && co < cout_pad) wd[((int64_t)tap * cin + ci) * cout_pad + co] = __float2bfloat16(v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (threadIdx.x == 0 && cnt) atomicAdd(&s_cnt, cnt);
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int co = co0 + r, ci = ci0 + threadIdx.x;
    if (co < cout && ci < cin_pad) wf[((int64_t)tap * cout + co) * cin_pad + ci] = __float2bfloat16(tile[threadIdx.x][r]);
  }
  if (threadIdx.x == 0 && threadIdx.y == 0 && s_cnt)
    atomicAdd(&nnz[((int64_t)tap * n_tiles + (co0 >> 6)) * k_tiles + (ci0 >> 6)], s_cnt);
}

}  // namespace rigl

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
  const PackedLayout L = packed_layout(taps, cin, cout);
  uint8_t* base = static_cast<uint8_t*>(packed);
  RIGL_CUDA(cudaMemsetAsync(base + L.off_nnz, 0, L.total - L.off_nnz, stream));
  dim3 grid((L.cout_pad + 31) / 32, (L.cin_pad + 31) / 32, taps), block(32, 8);
  k_pack_weights<<<grid, block, 0, stream>>>(w_hwio, mask_bits, cin, cout, L.cin_pad, L.cout_pad,
                                             (__nv_bfloat16*)(base + L.off_fprop),
                                             (__nv_bfloat16*)(base + L.off_dgrad),
                                             (uint32_t*)(base + L.off_nnz), L.n_tiles, L.k_tiles);
  RIGL_LAUNCH_CHECK("k_pack_weights");
  return RIGL_OK;
}
