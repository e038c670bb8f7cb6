// kernels_fused.cuh - K1: the fused front half of an MBConv block.
//
//   expand 1x1 (tcgen05, accumulators in TMEM) -> BN shift + swish -> shared memory (never HBM)
//   -> depthwise KSxKS stride S, TF-SAME (CUDA-core FMA on the smem tile) -> BN shift + swish
//   -> D (global, 16-bit) + deterministic SE squeeze partial sums
//
// One CTA owns a TH x TW tile of the depthwise OUTPUT of one crop.  The matching input halo tile
// (IH x IW = (TH-1)*S+KS square, raster order = GEMM rows) is staged once in the UMMA K-major
// SWIZZLE_128B layout; the expanded channels are then produced and consumed CC at a time:
//
//   for each chunk of CC expanded channels:
//       W chunk -> smem ; tcgen05.mma  D[mt][128 x CC] = A[mt] (128 x Cin) * Wc^T   for every 128-row tile mt
//       TMEM -> registers -> +shift, swish, ZERO for halo pixels outside the image (the depthwise pads the
//               EXPANDED tensor with zeros, not the block input) -> 16-bit -> E[pixel][CC] in smem
//       depthwise strips straight out of E (LDS.128), weights of one kernel row in registers
//       -> store D, accumulate the squeeze sums
//
// The expanded tensor (the largest activation of the network: 112*112*96 values per crop in block 2)
// therefore never leaves the SM.
#pragma once
#include "kernels_tc.cuh"

namespace whenet {
namespace fused {

using tc::BK;
using tc::BM;

struct K1Params {
    const void* in;        // T [N][Hin][Hin][Cin]
    const void* wt;        // T [Cexp][Cin]   (BN-folded, K-major)
    const float* b_exp;    // [Cexp]
    const float* w_dw;     // [KS*KS][Cexp]   (BN-folded)
    const float* b_dw;     // [Cexp]
    void* out;             // T [N][Ho][Ho][Cexp]
    float* partial;        // [N][tiles][Cexp]
    int Hin, Ho, Cin, Cexp, pad;
    int TH, TW, IH, IW;    // output tile, input halo tile
    int tiles_x, tiles_y;
    int CC, n_chunks;      // expanded channels per chunk (multiple of 16), number of chunks
    int mtiles;            // ceil(IH*IW / 128)
    int nkb;               // ceil(Cin / 64)
    int tmem_cols;         // power of two >= mtiles*CC
    int pitchE;            // bytes per E row = CC*2 + 16
    int PY;                // strip lanes in the depthwise phase = 256 / (CC/4)
    uint32_t idesc;
    int smem_A, smem_W, smem_E;   // byte sizes of the three regions (A and W multiples of 1024)
};

template <typename T> __device__ __forceinline__ void unpack8(const uint4& raw, float (&v)[8]);
template <> __device__ __forceinline__ void unpack8<__nv_bfloat16>(const uint4& raw, float (&v)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void unpack8<__half>(const uint4& raw, float (&v)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <typename T> __device__ __forceinline__ uint4 pack8(const float (&v)[8]);
template <> __device__ __forceinline__ uint4 pack8<__nv_bfloat16>(const float (&v)[8]) {
    uint4 t; __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    return t;
}
template <> __device__ __forceinline__ uint4 pack8<__half>(const float (&v)[8]) {
    uint4 t; __half2* h = reinterpret_cast<__half2*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    return t;
}

template <typename T, int KS, int S, int R>
__global__ void __launch_bounds__(256) k1_expand_dw_kernel(const K1Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t s_tmem_base;
    __shared__ int s_abort;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                       // [nkb][mtiles*128 rows][128 B]   swizzled
    uint8_t* sW = sA + p.smem_A;              // [nkb][CC rows][128 B]           swizzled
    uint8_t* sE = sW + p.smem_W;              // [IH*IW rows (+slack)][pitchE]
    float* s_red = reinterpret_cast<float*>(sE + p.smem_E);   // [PY][CC]

    const T* in = reinterpret_cast<const T*>(p.in);
    const T* wt = reinterpret_cast<const T*>(p.wt);
    T* out = reinterpret_cast<T*>(p.out);

    const int n = blockIdx.y, tile = blockIdx.x;
    const int ty0 = (tile / p.tiles_x) * p.TH, tx0 = (tile % p.tiles_x) * p.TW;     // output-tile origin
    const int iy0 = ty0 * S - p.pad, ix0 = tx0 * S - p.pad;                         // input-tile origin (may be < 0)
    const int npix = p.IH * p.IW;
    const int rows_total = p.mtiles * BM;
    const int kchunks = p.Cin >> 3;

    if (tid == 0) {
        tc::mbar_init(&mbar, 1);
        s_abort = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&s_tmem_base)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }

    // ---- A: the input halo tile, rows in raster order, zero outside the image, zero pad chunk when Cin/8 is odd
    {
        const T* in_n = in + (long long)n * p.Hin * p.Hin * p.Cin;
        const int cpr = (kchunks + 1) & ~1;                  // chunks written per row
        for (int idx = tid; idx < npix * cpr; idx += 256) {
            const int r = idx / cpr, c = idx - r * cpr;
            const int ty = r / p.IW, tx = r - ty * p.IW;
            const int iy = iy0 + ty, ix = ix0 + tx;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (c < kchunks && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Hin)
                v = *reinterpret_cast<const uint4*>(in_n + ((long long)iy * p.Hin + ix) * p.Cin + c * 8);
            const int kb = c >> 3, cc = c & 7;
            *reinterpret_cast<uint4*>(sA + (size_t)kb * rows_total * 128 + (r >> 3) * 1024 + (r & 7) * 128 + ((cc ^ (r & 7)) << 4)) = v;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = s_tmem_base;

    // depthwise-phase thread coordinates
    const int CVc = p.CC >> 2;                             // 4-channel vectors per chunk
    const int cv = tid % CVc, py = tid / CVc;
    const bool dw_active = py < p.PY;
    const int spr = (p.TW + R - 1) / R;
    const int nstrips = p.TH * spr;
    constexpr int NCOL = (R - 1) * S + KS;

    for (int ch = 0; ch < p.n_chunks; ++ch) {
        const int cbase = ch * p.CC;                        // first expanded channel of this chunk
        // ---- W chunk: CC rows (output channels) x Cin, swizzled K-major
        {
            const int cpr = (kchunks + 1) & ~1;
            for (int idx = tid; idx < p.CC * cpr; idx += 256) {
                const int r = idx / cpr, c = idx - r * cpr;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (c < kchunks && cbase + r < p.Cexp)
                    v = *reinterpret_cast<const uint4*>(wt + (long long)(cbase + r) * p.Cin + c * 8);
                const int kb = c >> 3, cc = c & 7;
                *reinterpret_cast<uint4*>(sW + (size_t)kb * p.CC * 128 + (r >> 3) * 1024 + (r & 7) * 128 + ((cc ^ (r & 7)) << 4)) = v;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int mt = 0; mt < p.mtiles; ++mt) {
                for (int kb = 0; kb < p.nkb; ++kb) {
                    const int krem = min(BK, p.Cin - kb * BK);
                    const int ksteps = (krem + 15) >> 4;
                    const uint64_t ad = tc::make_desc(tc::smem_u32(sA + (size_t)kb * rows_total * 128 + (size_t)mt * BM * 128));
                    const uint64_t bd = tc::make_desc(tc::smem_u32(sW + (size_t)kb * p.CC * 128));
                    for (int k = 0; k < ksteps; ++k)
                        tc::umma_f16(tmem_d + (uint32_t)(mt * p.CC), ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), p.idesc, (kb | k) ? 1u : 0u);
                }
            }
            tc::umma_commit(&mbar);
        }
        if (!tc::mbar_wait(&mbar, ch & 1)) s_abort = 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        __syncthreads();
        const bool ok = !s_abort;

        // ---- epilogue 1: TMEM -> E.  warp w reads lane quadrant (w & 3); warps 0-3 take the low column units, 4-7 the high
        if (ok) {
            const int q = warp & 3;
            const int units = p.CC >> 4;                    // 16-column units in the chunk
            const int u0 = (warp >> 2) ? (units + 1) / 2 : 0;
            const int u1 = (warp >> 2) ? units : (units + 1) / 2;
            for (int mt = 0; mt < p.mtiles; ++mt) {
                const int r = mt * BM + q * 32 + lane;
                const int ty = r / p.IW, tx = r - ty * p.IW;
                const int iy = iy0 + ty, ix = ix0 + tx;
                const bool inside = r < npix && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Hin;
                for (int u = u0; u < u1; ++u) {
                    float v[16];
                    tc::tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * p.CC + u * 16), v);
                    if (r < npix) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float o[8];
                            const float* bp = p.b_exp + cbase + u * 16 + h * 8;
                            const float4 b0 = *reinterpret_cast<const float4*>(bp), b1 = *reinterpret_cast<const float4*>(bp + 4);
                            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                            for (int j = 0; j < 8; ++j) o[j] = inside ? swish_fast(v[h * 8 + j] + bb[j]) : 0.f;
                            *reinterpret_cast<uint4*>(sE + (size_t)r * p.pitchE + (u * 16 + h * 8) * 2) = pack8<T>(o);
                        }
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();

        // ---- depthwise on E: thread = (4-channel vector, strip lane); 8-byte LDS, fp32 FMA
        float sum[4] = {0.f, 0.f, 0.f, 0.f};
        if (ok && dw_active) {
            const int c0 = cbase + cv * 4;
            const float4 bq = *reinterpret_cast<const float4*>(p.b_dw + c0);
            const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
            T* out_n = out + (long long)n * p.Ho * p.Ho * p.Cexp;
            const uint8_t* e_cv = sE + cv * 8;
            for (int sidx = py; sidx < nstrips; sidx += p.PY) {
                const int oyl = sidx / spr, oxl0 = (sidx - oyl * spr) * R;
                float acc[R][4];
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[r][i] = bb[i];
#pragma unroll
                for (int ky = 0; ky < KS; ++ky) {
                    float wr[KS][4];
#pragma unroll
                    for (int kx = 0; kx < KS; ++kx) {
                        const float4 w0 = *reinterpret_cast<const float4*>(p.w_dw + (ky * KS + kx) * p.Cexp + c0);
                        wr[kx][0] = w0.x; wr[kx][1] = w0.y; wr[kx][2] = w0.z; wr[kx][3] = w0.w;
                    }
                    const uint8_t* erow = e_cv + (size_t)((oyl * S + ky) * p.IW + oxl0 * S) * p.pitchE;
#pragma unroll
                    for (int col = 0; col < NCOL; ++col) {
                        float x[4];
                        ld4(reinterpret_cast<const T*>(erow + (size_t)col * p.pitchE), x);
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const int kx = col - r * S;
                            if (kx >= 0 && kx < KS) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) acc[r][i] = fmaf(x[i], wr[kx][i], acc[r][i]);
                            }
                        }
                    }
                }
                const int oy = ty0 + oyl;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int oxl = oxl0 + r, ox = tx0 + oxl;
                    if (oxl < p.TW && oy < p.Ho && ox < p.Ho) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[r][i] = swish_fast(acc[r][i]);
                        st4(out_n + ((long long)oy * p.Ho + ox) * p.Cexp + c0, acc[r]);
#pragma unroll
                        for (int i = 0; i < 4; ++i) sum[i] += Store<T>::rnd(acc[r][i]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) s_red[py * p.CC + cv * 4 + i] = sum[i];
        }
        __syncthreads();
        if (ok && dw_active && py == 0) {
            float tot[4] = {0.f, 0.f, 0.f, 0.f};
            for (int y = 0; y < p.PY; ++y)
#pragma unroll
                for (int i = 0; i < 4; ++i) tot[i] += s_red[y * p.CC + cv * 4 + i];
            float* dst = p.partial + ((long long)n * gridDim.x + tile) * p.Cexp + cbase + cv * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = tot[i];
        }
        // E, s_red and TMEM are free again after the barrier at the top of the next chunk (W fill + sync)
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)p.tmem_cols) : "memory");
}

// Tile plan for one block; returns false when K1 does not cover the configuration.
inline bool plan_k1(int Hin, int Ho, int Cin, int Cexp, int k, int s, int pad, bool is_bf16, K1Params* p, int* R_out, size_t* smem_out) {
    int TH, R;
    if (Ho % 8 == 0 && s == 2 && k == 3) { TH = 8; R = 2; }          // 112->56: 17x17 halo, 3 GEMM tiles
    else if (Ho % 14 == 0 && s == 1) { TH = 14; R = 7; }             // 56/28/14 maps, stride 1: 16x16 / 18x18 halo
    else if (Ho % 7 == 0) { TH = 7; R = 7; }                         // stride-2 5x5 / 3x3 onto 28 / 14 / 7
    else return false;
    p->Hin = Hin; p->Ho = Ho; p->Cin = Cin; p->Cexp = Cexp; p->pad = pad;
    p->TH = TH; p->TW = TH;
    p->IH = (TH - 1) * s + k; p->IW = p->IH;
    p->tiles_x = Ho / TH; p->tiles_y = Ho / TH;
    p->mtiles = (p->IH * p->IW + BM - 1) / BM;
    p->nkb = (Cin + BK - 1) / BK;
    // chunk: largest divisor of Cexp that is a multiple of 16, <= 128, with mtiles*CC <= 256 TMEM columns
    int CC = 0;
    for (int c = 16; c <= 128; c += 16)
        if (Cexp % c == 0 && p->mtiles * c <= 256) CC = c;
    if (!CC) return false;
    p->CC = CC; p->n_chunks = Cexp / CC;
    int cols = 32;
    while (cols < p->mtiles * CC) cols <<= 1;
    p->tmem_cols = cols;
    p->pitchE = CC * 2 + 16;
    p->PY = 256 / (CC / 4);
    p->idesc = tc::make_idesc(is_bf16, CC);
    p->smem_A = p->nkb * p->mtiles * BM * 128;
    p->smem_W = ((p->nkb * CC * 128) + 1023) & ~1023;
    // slack rows: a strip whose tail lies beyond TW still LOADS (its results are discarded)
    p->smem_E = ((p->IH * p->IW + 2 * p->IW + 16) * p->pitchE + 15) & ~15;
    *R_out = R;
    *smem_out = (size_t)p->smem_A + p->smem_W + p->smem_E + (size_t)p->PY * CC * 4 + 1024;
    return *smem_out <= 200 * 1024;
}

template <typename T>
int launch_k1(cudaStream_t stream, const K1Params& p, int k, int s, int R, size_t smem, int n_crops) {
    dim3 grid(p.tiles_x * p.tiles_y, n_crops);
#define K1(KS, S, RR)                                                                                            \
    do {                                                                                                         \
        auto kfn = k1_expand_dw_kernel<T, KS, S, RR>;                                                            \
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -1; \
        kfn<<<grid, 256, smem, stream>>>(p);                                                                     \
        return 0;                                                                                                \
    } while (0)
    if (k == 3 && s == 2 && R == 2) K1(3, 2, 2);
    if (k == 3 && s == 1 && R == 7) K1(3, 1, 7);
    if (k == 5 && s == 1 && R == 7) K1(5, 1, 7);
    if (k == 5 && s == 2 && R == 7) K1(5, 2, 7);
    if (k == 3 && s == 2 && R == 7) K1(3, 2, 7);
#undef K1
    return 1;
}

}  // namespace fused
}  // namespace whenet
