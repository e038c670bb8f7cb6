// placeholder, replaced below
#pragma once
#include <cuda_runtime.h>
namespace whenet { namespace tc {
template <typename T>
int launch_pw_tc(cudaStream_t, const T*, const void*, const float*, const float*, const T*, T*, long long, int, int, int, bool) { return 1; }
}}
