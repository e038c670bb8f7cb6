// kernels_k0.cuh - K0: stem + block-1 depthwise in one kernel (16-bit storage modes).
//
//   uint8 RGB -> LUT normalise -> conv3x3 s2 SAME (+BN, swish) -> 32 channels IN SHARED MEMORY
//   -> depthwise 3x3 s1 SAME (+BN, swish) -> D1 [N][112][112][32] + squeeze partials -> (last CTA of the crop) SE gate
//
// The stem output (112*112*32 values per crop, the largest tensor the unfused path writes AND reads back) never
// reaches HBM.  One CTA = one 14x14 tile of block-1's depthwise output; its 16x16 stem halo is exactly one stem
// pixel per thread (256 threads), recomputed at the tile borders (1.31x) - FMA is cheap, HBM is not.
// Stem weights and the BN shift arrive as a __grid_constant__ parameter (constant bank), pre-multiplied by 1/2
// so that swish(x) = h + h*tanh(h) needs no extra multiply; same for the depthwise constants.
#pragma once
#include "kernels_fused.cuh"

namespace whenet {
namespace fused {

struct K0Params {
    const uint8_t* in;       // [N][224][224][3] RGB
    const float* lut;        // [3][256] normalisation table (whenet.py:23-26 in float64, cast to float32)
    const float* w_dw;       // [9][32]  0.5 * BN-folded depthwise weights
    const float* b_dw;       // [32]     0.5 * BN shift
    void* out;               // T [N][112][112][32]
    float* partial;          // [N][64][32]
    const float *w_se1t, *b_se1, *w_se2, *b_se2;
    float* gate;             // [N][32]
    int* se_counter;         // [N] tickets (nullptr: SE left to se_gate_kernel)
    int Cse;
};

template <typename T>
__global__ void __launch_bounds__(256, 2) k0_stem_dw_kernel(const __grid_constant__ StemParams sp, const K0Params p) {
    constexpr int IN_W = 33 * 3 + 1;                  // floats per staged input row (33 pixels + pad)
    constexpr int PITCH_E = 32 * 2 + 16;              // bytes per stem pixel in smem (odd multiple of 16: conflict-free)
    __shared__ float s_in[33 * IN_W];
    __shared__ float s_lut[768];
    __shared__ __align__(16) uint8_t s_e[256 * PITCH_E];
    __shared__ float s_red[28 * 32];
    __shared__ float s_se[32 + 8];
    __shared__ int s_last;

    const int tid = threadIdx.x;
    const int n = blockIdx.y, tile = blockIdx.x;
    const int ty0 = (tile >> 3) * 14, tx0 = (tile & 7) * 14;       // origin of the 14x14 output tile (8 x 8 tiles)
    const int sy0 = ty0 - 1, sx0 = tx0 - 1;                        // origin of the 16x16 stem halo
    const int iy0 = 2 * sy0, ix0 = 2 * sx0;                        // origin of the 33x33 input patch

    for (int i = tid; i < 768; i += 256) s_lut[i] = p.lut[i];
    __syncthreads();
    // ---- input patch -> normalised fp32 (zero outside the image: only index 224 is ever used by a valid stem pixel)
    {
        const uint8_t* src = p.in + (long long)n * 224 * 224 * 3;
        for (int i = tid; i < 33 * 99; i += 256) {
            const int r = i / 99, q = i - r * 99;                 // q = col*3 + channel
            const int iy = iy0 + r, ix = ix0 + q / 3;
            float v = 0.f;
            if (iy >= 0 && iy < 224 && ix >= 0 && ix < 224) v = s_lut[(q % 3) * 256 + src[((long long)iy * 224 + ix) * 3 + q % 3]];
            s_in[r * IN_W + q] = v;
        }
    }
    __syncthreads();
    // ---- stem: one halo pixel per thread, 32 channels; weights straight from the constant bank
    {
        const int ly = tid >> 4, lx = tid & 15;
        const int sy = sy0 + ly, sx = sx0 + lx;
        uint4 o[4];
        if (sy >= 0 && sy < 112 && sx >= 0 && sx < 112) {
            float acc[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[c] = sp.b[c];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const float* row = &s_in[(2 * ly + ky) * IN_W + 6 * lx];
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const float x = row[t];
#pragma unroll
                    for (int c = 0; c < 32; ++c) acc[c] = fmaf(x, sp.w[(ky * 9 + t) * 32 + c], acc[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < 32; ++c) acc[c] = swish_from_half(acc[c]);
#pragma unroll
            for (int g = 0; g < 4; ++g)
                o[g] = make_uint4(pack2<T>(acc[g * 8], acc[g * 8 + 1]), pack2<T>(acc[g * 8 + 2], acc[g * 8 + 3]),
                                  pack2<T>(acc[g * 8 + 4], acc[g * 8 + 5]), pack2<T>(acc[g * 8 + 6], acc[g * 8 + 7]));
        } else {
#pragma unroll
            for (int g = 0; g < 4; ++g) o[g] = make_uint4(0u, 0u, 0u, 0u);   // depthwise SAME padding of the stem OUTPUT
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(&s_e[tid * PITCH_E + g * 16]) = o[g];
    }
    __syncthreads();
    // ---- depthwise 3x3 s1: thread = (4-channel vector cv, strip of 7 outputs); 28 strips x 8 vectors = 224 threads
    const int cv = tid & 7, sidx = tid >> 3;
    float sum[4] = {0.f, 0.f, 0.f, 0.f};
    if (sidx < 28) {
        const int oyl = sidx >> 1, oxl0 = (sidx & 1) * 7;
        const float4 bq = *reinterpret_cast<const float4*>(p.b_dw + cv * 4);
        float acc[7][4];
#pragma unroll
        for (int r = 0; r < 7; ++r) { acc[r][0] = bq.x; acc[r][1] = bq.y; acc[r][2] = bq.z; acc[r][3] = bq.w; }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            float4 wr[3];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) wr[kx] = *reinterpret_cast<const float4*>(p.w_dw + (ky * 3 + kx) * 32 + cv * 4);
            const uint8_t* e = &s_e[((oyl + ky) * 16 + oxl0) * PITCH_E + cv * 8];
#pragma unroll
            for (int col = 0; col < 9; ++col) {
                const uint2 raw = *reinterpret_cast<const uint2*>(e + col * PITCH_E);
                float x0, x1, x2, x3;
                unpack2<T>(raw.x, x0, x1);
                unpack2<T>(raw.y, x2, x3);
#pragma unroll
                for (int r = 0; r < 7; ++r) {
                    const int kx = col - r;
                    if (kx >= 0 && kx < 3) {
                        acc[r][0] = fmaf(x0, wr[kx].x, acc[r][0]);
                        acc[r][1] = fmaf(x1, wr[kx].y, acc[r][1]);
                        acc[r][2] = fmaf(x2, wr[kx].z, acc[r][2]);
                        acc[r][3] = fmaf(x3, wr[kx].w, acc[r][3]);
                    }
                }
            }
        }
        T* dst = reinterpret_cast<T*>(p.out) + (((long long)n * 112 + ty0 + oyl) * 112 + tx0 + oxl0) * 32 + cv * 4;
#pragma unroll
        for (int r = 0; r < 7; ++r) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { acc[r][i] = swish_from_half(acc[r][i]); sum[i] += acc[r][i]; }
            uint2 o;
            o.x = pack2<T>(acc[r][0], acc[r][1]);
            o.y = pack2<T>(acc[r][2], acc[r][3]);
            *reinterpret_cast<uint2*>(dst + r * 32) = o;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) s_red[sidx * 32 + cv * 4 + i] = sum[i];
    }
    __syncthreads();
    if (tid < 32) {
        float tot = 0.f;
        for (int y = 0; y < 28; ++y) tot += s_red[y * 32 + tid];
        p.partial[((long long)n * 64 + tile) * 32 + tid] = tot;
    }
    // ---- SE excite by the last CTA of the crop
    if (p.se_counter) {
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const int ticket = atomicAdd(p.se_counter + n, 1);
            s_last = ticket == 63;
            if (s_last) p.se_counter[n] = 0;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            se_gate_crop<true>(p.partial + (long long)n * 64 * 32, 64, 1.0f / (112.0f * 112.0f), p.w_se1t, p.b_se1, p.w_se2, p.b_se2,
                         p.gate + (long long)n * 32, 32, p.Cse, s_se);
        }
    }
}

}  // namespace fused
}  // namespace whenet
