// Translation unit: K1P (persistent, warp-specialised K1) for both 16-bit storage types (see inst_k1_bf16.cu).
#include "kernels_k1p.cuh"

namespace whenet {
namespace fused {
template int launch_k1p<__nv_bfloat16>(cudaStream_t, K1PParams, int, int, int, size_t, int, int);
template int launch_k1p<__half>(cudaStream_t, K1PParams, int, int, int, size_t, int, int);
}  // namespace fused
}  // namespace whenet
