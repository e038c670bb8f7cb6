// microbench.cu - pipe-throughput probes for the design decisions of the fused expand+depthwise kernel (B200, sm_100a):
//   MUFU.TANH / EX2 / RCP rate, FFMA vs FFMA2 (fma.rn.f32x2) rate, FFMA2 + ALU (bf16 unpack) + LDS mixes, a polynomial swish.
// Every probe: 2 CTAs x 512 threads per SM, 8 independent chains per thread, cycles from clock64 (max over CTAs).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <algorithm>
#include <vector>

#define ITERS 2048
typedef unsigned long long u64;

__device__ __forceinline__ float tanh_approx(float x) { float t; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x)); return t; }
__device__ __forceinline__ float ex2_approx(float x) { float t; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x)); return t; }
__device__ __forceinline__ float rcp_approx(float x) { float t; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x)); return t; }
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(reinterpret_cast<u64&>(d)) : "l"(reinterpret_cast<const u64&>(a)), "l"(reinterpret_cast<const u64&>(b)));
}

template <int MODE>
__global__ void __launch_bounds__(512) probe(float* out, u64* cyc, float seed) {
    __shared__ __align__(16) float sm[8192];     // 32 KB: sbase (<= 4 KB) + 3 * 4096 + 16 stays inside
    for (int i = threadIdx.x; i < 8192; i += 512) sm[i] = seed * (float)i;
    __syncthreads();
    float v[8];
    float2 w2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = seed * (float)(threadIdx.x + i); w2[i] = make_float2(v[i], -v[i]); }
    const float2 ka = make_float2(seed, seed * 0.5f), kb = make_float2(0.999f, 1.001f);
    uint32_t u = __float_as_uint(seed) + threadIdx.x;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 255) * 16;
    const u64 t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = tanh_approx(v[i]);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = ex2_approx(v[i]);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = rcp_approx(v[i]);
        } else if (MODE == 3) {          // FFMA, 8 chains
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], 0.999f, seed);
        } else if (MODE == 4) {          // FFMA2, 8 chains (16 FMAs)
#pragma unroll
            for (int i = 0; i < 8; ++i) ffma2(w2[i], ka, kb);
        } else if (MODE == 5) {          // swish via tanh: MUFU + FFMA
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float t = tanh_approx(v[i]); v[i] = fmaf(v[i], t, v[i]); }
        } else if (MODE == 6) {          // FFMA2 x8 + 8 ALU ops (bf16 unpack: shl + lop)
#pragma unroll
            for (int i = 0; i < 8; ++i) ffma2(w2[i], ka, kb);
#pragma unroll
            for (int i = 0; i < 4; ++i) { const uint32_t a = u << 16, b = u & 0xffff0000u; u = (a ^ b) + it; v[i] = __uint_as_float(a); v[i + 4] = __uint_as_float(b); }
        } else if (MODE == 7) {          // depthwise-like: 1 LDS.64 + 4 ALU + 8 FFMA2 (16 FMA)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                uint32_t a, b;
                asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(sbase + ((it + r) & 3) * 4096));
                const float2 x01 = make_float2(__uint_as_float(a << 16), __uint_as_float(a & 0xffff0000u));
                const float2 x23 = make_float2(__uint_as_float(b << 16), __uint_as_float(b & 0xffff0000u));
#pragma unroll
                for (int i = 0; i < 4; ++i) { ffma2(w2[2 * i], x01, kb); ffma2(w2[2 * i + 1], x23, kb); }
            }
        } else if (MODE == 8) {          // same with scalar FFMA (16 FFMA per LDS.64)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                uint32_t a, b;
                asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(sbase + ((it + r) & 3) * 4096));
                const float x0 = __uint_as_float(a << 16), x1 = __uint_as_float(a & 0xffff0000u);
                const float x2 = __uint_as_float(b << 16), x3 = __uint_as_float(b & 0xffff0000u);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    w2[2 * i].x = fmaf(x0, 0.999f, w2[2 * i].x); w2[2 * i].y = fmaf(x1, 1.001f, w2[2 * i].y);
                    w2[2 * i + 1].x = fmaf(x2, 0.999f, w2[2 * i + 1].x); w2[2 * i + 1].y = fmaf(x3, 1.001f, w2[2 * i + 1].y);
                }
            }
        } else if (MODE == 9) {          // polynomial sigmoid-free swish: clamp + 6 FFMA (odd minimax of tanh) + FFMA, per element
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float h = fminf(fmaxf(v[i], -4.97f), 4.97f);
                const float s = h * h;
                float p = fmaf(s, -5.0e-8f, 3.0e-6f);
                p = fmaf(p, s, -1.0e-4f); p = fmaf(p, s, 2.1e-3f); p = fmaf(p, s, -2.2e-2f); p = fmaf(p, s, 1.33e-1f); p = fmaf(p, s, -3.33e-1f);
                const float t = fmaf(p * s, h, h);
                v[i] = fmaf(v[i], t, v[i]);
            }
        } else if (MODE == 10) {         // MUFU tanh x8 interleaved with FFMA2 x8: do the pipes overlap?
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[i] = tanh_approx(v[i]); ffma2(w2[i], ka, kb); }
        } else if (MODE == 11) {         // LDS.128 stream
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 q;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(sbase + ((it + i) & 3) * 4096));
                v[i] += q.x + q.w;
            }
        } else if (MODE == 12) {         // cvt.rn.bf16x2 pack x8
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t r;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v[i]), "f"(v[(i + 1) & 7]));
                v[i] = __uint_as_float(r);
            }
        }
    }
    const u64 t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += v[i] + w2[i].x + w2[i].y;
    out[blockIdx.x * 512 + threadIdx.x] = acc + __uint_as_float(u);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, double ops_per_thread_iter, float* d_out, u64* d_cyc, int ctas) {
    probe<MODE><<<ctas, 512>>>(d_out, d_cyc, 1e-3f);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    probe<MODE><<<ctas, 512>>>(d_out, d_cyc, 1e-3f);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    std::vector<u64> h(ctas);
    cudaMemcpy(h.data(), d_cyc, ctas * sizeof(u64), cudaMemcpyDeviceToHost);
    const u64 mx = *std::max_element(h.begin(), h.end());
    const double per_sm = ops_per_thread_iter * ITERS * 512.0 * 2.0;    // two CTAs per SM
    printf("%-44s  %8.1f cycles/iter/warp-set  %7.2f ops/clk/SM   (%.3f ms, %s)\n", name, (double)mx / ITERS, per_sm / (double)mx, ms,
           cudaGetErrorString(cudaGetLastError()));
}

// single-warp dependent chains: latency of one op (cycles per link)
template <int MODE>
__global__ void latency_probe(float* out, u64* cyc, float seed) {
    __shared__ float sm[64];
    sm[threadIdx.x] = seed * threadIdx.x;
    __syncthreads();
    float v = seed * (float)(threadIdx.x + 1);
    float2 w = make_float2(v, -v);
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sm);
    const u64 t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) v = tanh_approx(v);
            else if (MODE == 1) v = fmaf(v, 0.999f, seed);
            else if (MODE == 2) ffma2(w, w, make_float2(0.999f, 1.001f));
            else if (MODE == 3) { uint32_t a; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(a) : "r"(sa + (__float_as_uint(v) & 0x7c))); v = __uint_as_float(a); }
            else if (MODE == 4) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v), "f"(v)); v = __uint_as_float(r << 16); }
            else if (MODE == 5) v = ex2_approx(v);
        }
    }
    const u64 t1 = clock64();
    out[threadIdx.x] = v + w.x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE>
static void run_lat(const char* name, float* d_out, u64* d_cyc) {
    latency_probe<MODE><<<1, 32>>>(d_out, d_cyc, 1e-3f);
    cudaDeviceSynchronize();
    latency_probe<MODE><<<1, 32>>>(d_out, d_cyc, 1e-3f);
    cudaDeviceSynchronize();
    u64 c = 0; cudaMemcpy(&c, d_cyc, sizeof(u64), cudaMemcpyDeviceToHost);
    printf("latency %-40s %6.1f cycles per dependent op\n", name, (double)c / (256.0 * 8.0));
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int ctas = sms * 2;
    float* d_out; u64* d_cyc;
    cudaMalloc(&d_out, ctas * 512 * sizeof(float)); cudaMalloc(&d_cyc, ctas * sizeof(u64));
    printf("SMs %d, %d CTAs x 512 threads (32 warps/SM)\n", sms, ctas);
    run<0>("MUFU tanh.approx", 8, d_out, d_cyc, ctas);
    run<1>("MUFU ex2.approx", 8, d_out, d_cyc, ctas);
    run<2>("MUFU rcp.approx", 8, d_out, d_cyc, ctas);
    run<3>("FFMA (ops = FMAs)", 8, d_out, d_cyc, ctas);
    run<4>("FFMA2 (ops = FMAs)", 16, d_out, d_cyc, ctas);
    run<5>("swish = tanh + FFMA (ops = elements)", 8, d_out, d_cyc, ctas);
    run<6>("FFMA2 x8 + 8 ALU unpack (ops = FMAs)", 16, d_out, d_cyc, ctas);
    run<7>("dw-like LDS64 + unpack + FFMA2 (ops = FMAs)", 32, d_out, d_cyc, ctas);
    run<8>("dw-like LDS64 + unpack + FFMA  (ops = FMAs)", 32, d_out, d_cyc, ctas);
    run<9>("polynomial swish, FMA pipe only (elements)", 8, d_out, d_cyc, ctas);
    run<10>("tanh x8 + FFMA2 x8 interleaved (ops = tanh)", 8, d_out, d_cyc, ctas);
    run<11>("LDS.128 (ops = loads)", 8, d_out, d_cyc, ctas);
    run<12>("cvt.rn.bf16x2 (ops = cvts)", 8, d_out, d_cyc, ctas);
    run_lat<0>("MUFU.TANH", d_out, d_cyc);
    run_lat<5>("MUFU.EX2", d_out, d_cyc);
    run_lat<1>("FFMA", d_out, d_cyc);
    run_lat<2>("FFMA2", d_out, d_cyc);
    run_lat<3>("LDS.32 (address dependent)", d_out, d_cyc);
    run_lat<4>("F2FP.BF16 pack + shift", d_out, d_cyc);
    return 0;
}
