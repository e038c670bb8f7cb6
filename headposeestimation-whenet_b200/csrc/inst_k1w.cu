// Translation unit: K1W (weight-stationary persistent K1, TMA input tiles) for both 16-bit storage types (see inst_k1_bf16.cu).
#include "kernels_k1w.cuh"

namespace whenet {
namespace fused {
template int launch_k1w<__nv_bfloat16>(cudaStream_t, K1WParams, int, int, int, int, size_t, int, int);
template int launch_k1w<__half>(cudaStream_t, K1WParams, int, int, int, int, size_t, int, int);
}  // namespace fused
}  // namespace whenet
