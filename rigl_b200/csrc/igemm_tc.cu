// tcgen05 implicit-GEMM path -- placeholder until the kernels land (next commit).
#include "common.cuh"
#include "conv_common.cuh"

namespace rigl {
bool tc_supported(const ConvGeom&, int) { return false; }
size_t tc_workspace_bytes(const ConvGeom&) { return 0; }
int tc_fprop(const ConvGeom&, const void*, const void*, void*, float*, const float*, void*, size_t, cudaStream_t) { return RIGL_ERR_UNSUPPORTED; }
int tc_dgrad(const ConvGeom&, const void*, const void*, void*, void*, size_t, cudaStream_t) { return RIGL_ERR_UNSUPPORTED; }
int tc_wgrad(const ConvGeom&, const void*, const void*, float*, float, void*, size_t, cudaStream_t) { return RIGL_ERR_UNSUPPORTED; }
}  // namespace rigl
