// Translation unit: the K1 variants - K1P (persistent, warp-specialised) and K1T (depthwise on the tensor core) - for both
// 16-bit storage types (see inst_k1_bf16.cu).
#include "kernels_fused_tc.cuh"
#include "kernels_k1p.cuh"

namespace whenet {
namespace fused {
template int launch_k1p<__nv_bfloat16>(cudaStream_t, K1PParams, int, int, int, size_t, int, int);
template int launch_k1p<__half>(cudaStream_t, K1PParams, int, int, int, size_t, int, int);
template int launch_k1t<__nv_bfloat16>(cudaStream_t, const K1TParams&, int, int, size_t, int);
template int launch_k1t<__half>(cudaStream_t, const K1TParams&, int, int, size_t, int);
}  // namespace fused
int tu_timeout_k1x() { return tc::read_and_clear_timeout_flag(); }
}  // namespace whenet
