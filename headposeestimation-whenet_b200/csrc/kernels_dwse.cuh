// kernels_dwse.cuh - KD: depthwise KSxKS + BN shift + swish + squeeze + excite (+ gating of its own output) for the blocks
// whose feature map is small enough that ONE CTA holds a whole crop (14x14 and 7x7, blocks 7-16).
//
// Why a second route next to K1 for these blocks: K1 keeps the expanded tensor on the SM, which saves HBM traffic the late
// blocks do not have (their expanded tensors - 29..135 MB per 256..512 crops - live in the 126 MB L2), and pays for it with
// a chain of CTA-wide phases (MMA wait -> TMEM epilogue -> barrier -> depthwise -> barrier -> reduce) at one CTA per SM;
// measured 13.9 k cycles per channel chunk at 7x7 for ~6.5 k warp instructions (profiles/README.md, round 2).  Here the
// expand conv runs as a plain tcgen05 GEMM whose epilogue writes E as fp16 (pw_tc2, OUT_H) and this kernel does the rest
// with every thread busy on identical work:
//
//   CTA = one crop; for each chunk of CC channels (cp.async double buffer: chunk i+1 lands while chunk i computes):
//       E[crop][all pixels][CC] -> zero-bordered smem tile (TF-SAME padding = the border, no bounds tests in the loop)
//       thread = (strip of 7 output pixels of one row, 4 channels): HFMA2 running sums over the fp16 tile, fp16 weights / 4
//       -> fp32: sum * 4 + shift, swish, squeeze partial sums (fixed order), 16-bit store of D
//   tail: channel means -> FC + swish -> FC + sigmoid -> gate (same device function as se_gate_kernel, same bits);
//         the CTA then rescales its own D (still in L2) so that the project conv runs ungated.
//
// The arithmetic of one output is exactly K1's (fp16 E, HFMA2 taps in the same order, fp32 epilogue), so both routes agree to
// the rounding of the expand accumulators (K1 carries the BN shift through the tensor core as a bf16 hi/lo pair, the GEMM
// route adds it in fp32).
#pragma once
#include "kernels_fused.cuh"

namespace whenet {
namespace fused {

struct DwSeParams {
    const __half* E;        // [N][HIN][HIN][C] fp16 (expand conv + BN + swish)
    const __half* w16;      // [KS*KS][C]  0.5 * BN-folded depthwise weights / kDwScale
    const float* b_dw;      // [C]         0.5 * BN shift
    void* out;              // T [N][Ho][Ho][C]
    float* partial;         // [N][1][C]   squeeze sums (tiles = 1)
    const float *w_se1t, *b_se1, *w_se2, *b_se2;
    float* gate;            // [N][C]
    int Cse;
    float inv_hw;
    int se_tail;            // 1: this CTA sees every channel of its crop -> computes the gate itself
    int scale_out;          // se_tail only: D *= gate in place
    int C, pad;             // channels, TF-SAME pad_before
    int n_chunks, chunks_per_cta;
    int N;
};

template <int KS, int S, int HIN>
struct DwSeGeom {
    static constexpr int HO = (HIN + S - 1) / S;
    static constexpr int R = 7;                                // outputs per strip (HO is 14 or 7)
    static constexpr int SPR = HO / R;                         // strips per output row
    static constexpr int NSTRIPS = HO * SPR;
    static constexpr int PW = (HO - 1) * S + KS;               // padded tile width
    static constexpr int NCOL = (R - 1) * S + KS;
};

template <int KS, int S, int HIN, int CC>
struct DwSeThreads { static constexpr int value = ((DwSeGeom<KS, S, HIN>::NSTRIPS * (CC / 4)) + 31) / 32 * 32; };

template <int KS, int S, int HIN, int CC>
constexpr size_t dwse_smem(int C, int Cse) {
    using G = DwSeGeom<KS, S, HIN>;
    return (size_t)2 * G::PW * G::PW * CC * 2                  // two tiles
           + (size_t)2 * (CC * 4 + KS * KS * CC * 2)           // two constant sets
           + (size_t)2 * G::NSTRIPS * CC * 4                   // two squeeze scratch sets
           + (size_t)(C + Cse + 32) * 4 + 128;
}

template <typename T, int KS, int S, int HIN, int CC>
__global__ void __launch_bounds__((DwSeThreads<KS, S, HIN, CC>::value)) dwse_kernel(const DwSeParams p) {
    using G = DwSeGeom<KS, S, HIN>;
    constexpr int NT = DwSeThreads<KS, S, HIN, CC>::value;
    constexpr int CV = CC / 4;                                 // 4-channel vectors per pixel
    constexpr int PITCH = CC * 2;                              // bytes per tile pixel
    constexpr int TILE_BYTES = G::PW * G::PW * PITCH;
    constexpr int CST_BYTES = CC * 4 + KS * KS * CC * 2;
    constexpr int RED_BYTES = G::NSTRIPS * CC * 4;
    extern __shared__ __align__(128) uint8_t smem_dw[];
    const uint32_t s0 = (tc::smem_u32(smem_dw) + 127u) & ~127u;
    const uint32_t sT = s0, sC = sT + 2 * TILE_BYTES, sR = sC + 2 * CST_BYTES;
    float* const sM = reinterpret_cast<float*>(smem_dw + (sR + 2 * RED_BYTES - tc::smem_u32(smem_dw)));     // [C] means | [Cse] hidden

    const int tid = threadIdx.x;
    const int n = blockIdx.x;
    const int C = p.C;
    const int ch_begin = blockIdx.y * p.chunks_per_cta;
    const int ch_end = min(p.n_chunks, ch_begin + p.chunks_per_cta);
    const __half* E_n = p.E + (long long)n * HIN * HIN * C;
    T* const out_n = reinterpret_cast<T*>(p.out) + (long long)n * G::HO * G::HO * C;

    // zero both tiles once: the cp.async fills below only ever touch the interior, the border IS the SAME padding
    {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int i = tid; i < 2 * TILE_BYTES / 16; i += NT) sts128(sT + (uint32_t)i * 16u, z);
    }
    __syncthreads();

    auto prefetch = [&](int ch, int buf) {
        const int cbase = ch * CC;
        constexpr int CPP = CC / 8;                            // 16-byte pieces per pixel
        const uint32_t t_dst = sT + buf * TILE_BYTES;
        for (int idx = tid; idx < HIN * HIN * CPP; idx += NT) {
            const int pix = idx / CPP, c = idx - pix * CPP;
            const int y = pix / HIN, x = pix - y * HIN;
            cp_async16(t_dst + (uint32_t)(((y + p.pad) * G::PW + x + p.pad) * PITCH + c * 16), E_n + (long long)pix * C + cbase + c * 8, true);
        }
        const uint32_t c_dst = sC + buf * CST_BYTES;
        constexpr int QB = CC / 4, QW = CC / 8;                // 16-byte pieces of the shift row / of one weight row
        for (int idx = tid; idx < QB + KS * KS * QW; idx += NT) {
            if (idx < QB) cp_async16(c_dst + (uint32_t)idx * 16, p.b_dw + cbase + idx * 4, true);
            else {
                const int t = idx - QB, row = t / QW, j = t - row * QW;
                cp_async16(c_dst + (uint32_t)(CC * 4 + row * CC * 2 + j * 16), p.w16 + (long long)row * C + cbase + j * 8, true);
            }
        }
    };
    // squeeze sums of a finished chunk: fixed order over the strips (four chains, as K1) -> reproducible bits
    auto finish_sums = [&](int ch) {
        if (tid < CC) {
            const uint32_t r0 = sR + (uint32_t)((ch & 1) * RED_BYTES + tid * 4);
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int y = 0; y < G::NSTRIPS; ++y) {
                float t;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(r0 + (uint32_t)(y * CC * 4)));
                s4[y & 3] += t;
            }
            const float tot = (s4[0] + s4[1]) + (s4[2] + s4[3]);
            p.partial[(long long)n * C + ch * CC + tid] = tot;
            if (p.se_tail) sM[ch * CC + tid] = tot * p.inv_hw;
        }
    };

    const int strip = tid / CV, cv = tid - strip * CV;
    const bool active = strip < G::NSTRIPS;
    const int oy = strip / G::SPR, ox0 = (strip - oy * G::SPR) * G::R;
    const uint32_t win = (uint32_t)(((oy * S) * G::PW + ox0 * S) * PITCH + cv * 8);       // top-left of this strip's input window

    prefetch(ch_begin, ch_begin & 1);
    asm volatile("cp.async.commit_group;" ::: "memory");

    for (int ch = ch_begin; ch < ch_end; ++ch) {
        const int buf = ch & 1;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                       // chunk ch has landed for everyone; everyone is done with chunk ch-1
        if (ch + 1 < ch_end) prefetch(ch + 1, buf ^ 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (ch > ch_begin) finish_sums(ch - 1);
        if (active) {
            const uint32_t cst = sC + buf * CST_BYTES;
            const float4 bq = lds_f4(cst + (uint32_t)cv * 16);
            const uint32_t cst_h = cst + (uint32_t)(CC * 4 + cv * 8);
            uint32_t erow = sT + buf * TILE_BYTES + win;
            __half2 hacc[G::R][2];
#pragma unroll
            for (int r = 0; r < G::R; ++r) { hacc[r][0] = __float2half2_rn(0.f); hacc[r][1] = __float2half2_rn(0.f); }
#pragma unroll
            for (int ky = 0; ky < KS; ++ky) {
                __half2 wr[KS][2];
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) {
                    uint32_t w0, w1;
                    lds64(cst_h + (uint32_t)((ky * KS + kx) * CC) * 2, w0, w1);
                    wr[kx][0] = *reinterpret_cast<__half2*>(&w0); wr[kx][1] = *reinterpret_cast<__half2*>(&w1);
                }
#pragma unroll
                for (int col = 0; col < G::NCOL; ++col) {
                    uint32_t a, b;
                    lds64(erow + (uint32_t)(col * PITCH), a, b);
                    const __half2 x01 = *reinterpret_cast<__half2*>(&a), x23 = *reinterpret_cast<__half2*>(&b);
#pragma unroll
                    for (int r = 0; r < G::R; ++r) {
                        const int kx = col - r * S;          // compile-time after unrolling
                        if (kx >= 0 && kx < KS) {
                            hacc[r][0] = __hfma2(x01, wr[kx][0], hacc[r][0]);
                            hacc[r][1] = __hfma2(x23, wr[kx][1], hacc[r][1]);
                        }
                    }
                }
                erow += G::PW * PITCH;
            }
            const float2 sc = make_float2(kDwScale, kDwScale);
            float sum[4] = {0.f, 0.f, 0.f, 0.f};
            T* dst = out_n + ((long long)oy * G::HO + ox0) * C + ch * CC + cv * 4;
#pragma unroll
            for (int r = 0; r < G::R; ++r) {
                float2 a0 = make_float2(bq.x, bq.y), a1 = make_float2(bq.z, bq.w);
                ffma2(a0, __half22float2(hacc[r][0]), sc);          // sum * kDwScale + shift, in fp32
                ffma2(a1, __half22float2(hacc[r][1]), sc);
                a0.x = swish_from_half(a0.x); a0.y = swish_from_half(a0.y);
                a1.x = swish_from_half(a1.x); a1.y = swish_from_half(a1.y);
                sum[0] += a0.x; sum[1] += a0.y; sum[2] += a1.x; sum[3] += a1.y;
                uint2 o;
                o.x = pack2<T>(a0.x, a0.y);
                o.y = pack2<T>(a1.x, a1.y);
                *reinterpret_cast<uint2*>(dst + (long long)r * C) = o;
            }
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(sR + (uint32_t)(buf * RED_BYTES + (strip * CC + cv * 4) * 4)),
                         "f"(sum[0]), "f"(sum[1]), "f"(sum[2]), "f"(sum[3]) : "memory");
        }
    }
    __syncthreads();
    finish_sums(ch_end - 1);
    __syncthreads();

    // ---- SE excite + gating of this crop's depthwise output (the CTA wrote all of it; it is still in L2)
    if (p.se_tail) {
        se_gate_fc<NT>(sM, sM + C, p.w_se1t, p.b_se1, p.w_se2, p.b_se2, p.gate + (long long)n * C, C, p.Cse, sM);
        __syncthreads();
        if (p.scale_out) {
            const int cv8 = C >> 3, total = G::HO * G::HO * cv8;
            const float inv_cv8 = 1.0f / (float)cv8;
            for (int v = tid; v < total; v += NT) {
                const int c8 = (v - div_small(v, inv_cv8) * cv8) * 8;
                uint4* ptr = reinterpret_cast<uint4*>(out_n) + v;
                *ptr = tc::scale8s<T>(__ldcg(ptr), tc::smem_u32(sM + c8));
            }
        }
    }
}

// which (kernel, stride, map size, channels) combinations have an instance, and with which chunk width
inline int dwse_chunk(int k, int s, int hin, int C) {
    if (hin == 14 && s == 1 && (k == 3 || k == 5) && C % 32 == 0) return 32;
    if (hin == 14 && s == 2 && k == 5 && C % 96 == 0) return 96;
    if (hin == 7 && s == 1 && (k == 3 || k == 5) && C % 128 == 0) return 128;
    return 0;
}

template <typename T>
int launch_dwse(cudaStream_t stream, DwSeParams p, int k, int s, int hin, int n_crops, int split) {
    const int CCr = dwse_chunk(k, s, hin, p.C);
    if (!CCr) return 1;
    p.N = n_crops;
    p.n_chunks = p.C / CCr;
    if (split < 1) split = 1;
    if (split > p.n_chunks) split = p.n_chunks;
    p.chunks_per_cta = (p.n_chunks + split - 1) / split;
    const int gy = (p.n_chunks + p.chunks_per_cta - 1) / p.chunks_per_cta;
    if (gy > 1) { p.se_tail = 0; p.scale_out = 0; }
#define DWSE(KS, S, HIN, CC)                                                                                              \
    do {                                                                                                                  \
        auto kfn = dwse_kernel<T, KS, S, HIN, CC>;                                                                        \
        const size_t smem = dwse_smem<KS, S, HIN, CC>(p.C, p.Cse);                                                        \
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;  \
        kfn<<<dim3(n_crops, gy), DwSeThreads<KS, S, HIN, CC>::value, smem, stream>>>(p);                                    \
        return 0;                                                                                                         \
    } while (0)
    if (hin == 14 && s == 1 && k == 3) DWSE(3, 1, 14, 32);
    if (hin == 14 && s == 1 && k == 5) DWSE(5, 1, 14, 32);
    if (hin == 14 && s == 2 && k == 5) DWSE(5, 2, 14, 96);
    if (hin == 7 && s == 1 && k == 5) DWSE(5, 1, 7, 128);
    if (hin == 7 && s == 1 && k == 3) DWSE(3, 1, 7, 128);
#undef DWSE
    return 1;
}

}  // namespace fused
}  // namespace whenet
