// Translation unit: K1 (kernels_fused.cuh) for fp16 storage (see inst_k1_bf16.cu).
#include "kernels_fused.cuh"

namespace whenet {
namespace fused {
template int launch_k1<__half>(cudaStream_t, K1Params, int, int, int, int, size_t, int);
template int launch_dw_only<__half>(cudaStream_t, K1Params, size_t, int);
}  // namespace fused
}  // namespace whenet
