// kernels_fused_tc.cuh - K1T: the fused front half of an MBConv block with BOTH convolutions on the tensor core.
//
//   expand 1x1   : tcgen05.mma  E_acc[128 x CC] = [A | 1 1] * [Wc | shift]^T            (as in K1)
//   epilogue 1   : TMEM -> swish -> 16-bit -> shared memory, stored as PLANES of 8 channels: plane g = [pixel][8 ch]
//   depthwise kxk: tcgen05.mma again.  For tap (ky,kx) the A operand is the SAME plane pair viewed from a start
//                  address shifted by (ky*IW + kx) pixels (16 bytes per pixel, SWIZZLE_NONE K-major canonical layout
//                  with SBO = 128 B makes a plane one linear array of rows), the B operand is a 16x16 DIAGONAL
//                  matrix holding that tap's 16 per-channel weights.  k*k accumulating MMAs (M=128 pixels, N=16, K=16)
//                  produce the stencil for 128 raster positions x 16 channels; 15/16 of the multiplies hit zeros,
//                  which costs nothing next to the CUDA-core alternative (the tensor pipe is otherwise idle).
//   epilogue 2   : TMEM -> +shift, swish -> rows that are real outputs (stride / tile bounds) -> D (global, 16-bit)
//                  + squeeze sums (transposed warp-shuffle reduction, fixed order)
//
// Compared with K1 the depthwise costs no FMA/LDS/unpack instructions at all: the CUDA cores only run the two
// activation epilogues.  Depthwise weights are rounded to the 16-bit storage type here (K1 keeps them fp32).
#pragma once
#include "kernels_fused.cuh"

namespace whenet {
namespace fused {

struct K1TParams {
    const void* in;        // T [N][Hin][Hin][Cin]
    const void* wt_aug;    // T [Cexp][Cin+8]   0.5 * (weights | shift hi | shift lo | 0...)
    const float* w_dw;     // [KS*KS][Cexp]     0.5 * BN-folded depthwise weights
    const float* b_dw;     // [Cexp]            0.5 * BN shift
    void* out;             // T [N][Ho][Ho][Cexp]
    float* partial;        // [N][tiles][Cexp]
    int Hin, Ho, Cin, Cexp, pad;
    int TH, TW, IH, IW;
    int tiles_x, tiles_y;
    int CC, n_chunks;      // expanded channels per chunk (multiple of 16)
    int mtiles;            // ceil(IH*IW / 128) (<= 3)
    int cpr, nkb;          // expand operand row: 16-byte chunks (even), K blocks of 64
    int P;                 // pixel slots per E plane (rows_total + largest tap offset + 8)
    int tmem_cols;         // power of two >= 2*mtiles*CC (expand accumulators | depthwise accumulators)
    uint32_t idesc_exp, idesc_dw;
    int smem_A, smem_W, smem_E, smem_B;   // bytes: A, one W buffer, all E planes, all diagonal matrices
};

// SWIZZLE_NONE (INTERLEAVE) K-major descriptor: 8x16-byte core matrices; LBO = bytes between the two K chunks,
// SBO = bytes between 8-row groups (cute::UMMA canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units).
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

template <typename T> __device__ __forceinline__ uint16_t to_bits16(float v);
template <> __device__ __forceinline__ uint16_t to_bits16<__nv_bfloat16>(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }
template <> __device__ __forceinline__ uint16_t to_bits16<__half>(float v) { return __half_as_ushort(__float2half_rn(v)); }

template <typename T, int KS, int S>
__global__ void __launch_bounds__(256, 2) k1t_kernel(const K1TParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t mbar_exp, mbar_dw;
    __shared__ uint32_t s_tmem_base;
    __shared__ int s_abort;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = smem0;
    const uint32_t sW = sA + p.smem_A;                         // 2 buffers
    const uint32_t sE = sW + 2 * p.smem_W;                     // CC/8 planes of P x 16 B
    const uint32_t sB = sE + p.smem_E;                         // [tap][CC/16] x 512 B diagonal matrices
    const uint32_t sBias = sB + p.smem_B;                      // CC floats
    const uint32_t sR = sBias + (uint32_t)p.CC * 4;            // [8 warps][CC] squeeze partials

    const T* in = reinterpret_cast<const T*>(p.in);
    const T* wt = reinterpret_cast<const T*>(p.wt_aug);
    T* out = reinterpret_cast<T*>(p.out);

    const int n = blockIdx.y, tile = blockIdx.x;
    const int tyi = tile / p.tiles_x;
    const int ty0 = tyi * p.TH, tx0 = (tile - tyi * p.tiles_x) * p.TW;
    const int iy0 = ty0 * S - p.pad, ix0 = tx0 * S - p.pad;
    const int npix = p.IH * p.IW;
    const int rows_total = p.mtiles * BM;
    const int kchunks = p.Cin >> 3;
    const int Kaug = p.Cin + 8;
    const uint32_t plane_bytes = (uint32_t)p.P * 16;
    const int units = p.CC >> 4;
    constexpr int TAPS = KS * KS;

    if (tid == 0) {
        tc::mbar_init(&mbar_exp, 1);
        tc::mbar_init(&mbar_dw, 1);
        s_abort = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&s_tmem_base)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- A: input halo tile (cp.async, zero outside the image) + ones chunk + even-count pad chunk
    {
        const T* in_n = in + (long long)n * p.Hin * p.Hin * p.Cin;
        const int items = npix * kchunks;
        int r = tid / kchunks, c = tid - r * kchunks;
        int ty = r / p.IW, tx = r - ty * p.IW;
        const int dr = 256 / kchunks, dc = 256 - dr * kchunks;
        const int dty = dr / p.IW, dtx = dr - dty * p.IW;
        for (int idx = tid; idx < items; idx += 256) {
            const int iy = iy0 + ty, ix = ix0 + tx;
            const bool valid = iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Hin;
            const T* src = valid ? in_n + ((long long)iy * p.Hin + ix) * p.Cin + c * 8 : in_n;
            cp_async16(sA + (uint32_t)(c >> 3) * rows_total * 128 + sw128(r, c & 7), src, valid);
            c += dc; r += dr; ty += dty; tx += dtx;
            if (c >= kchunks) { c -= kchunks; ++r; ++tx; }
            if (tx >= p.IW) { tx -= p.IW; ++ty; }
            if (tx >= p.IW) { tx -= p.IW; ++ty; }
        }
        const uint4 ones = make_uint4(ones2<T>(), 0u, 0u, 0u), zero = make_uint4(0u, 0u, 0u, 0u);
        for (int rr = tid; rr < npix; rr += 256) {
            sts128(sA + (uint32_t)(kchunks >> 3) * rows_total * 128 + sw128(rr, kchunks & 7), ones);
            if (p.cpr > kchunks + 1)
                sts128(sA + (uint32_t)((kchunks + 1) >> 3) * rows_total * 128 + sw128(rr, (kchunks + 1) & 7), zero);
        }
        // the diagonal matrices: everything off the diagonal stays zero for the whole kernel
        for (int i = tid; i < p.smem_B / 16; i += 256) sts128(sB + i * 16, zero);
        // E plane slack rows are read by discarded output rows only; zero them once so no NaN pattern ever enters the pipe
        for (int i = tid; i < p.smem_E / 16; i += 256) sts128(sE + i * 16, zero);
    }
    auto prefetch_w = [&](int ch, int buf) {
        const int cbase = ch * p.CC;
        const uint32_t w_dst = sW + buf * p.smem_W;
        for (int idx = tid; idx < p.CC * p.cpr; idx += 256) {
            const int r = idx / p.cpr, c = idx - r * p.cpr;
            const bool valid = c <= kchunks;
            cp_async16(w_dst + (uint32_t)(c >> 3) * p.CC * 128 + sw128(r, c & 7),
                       valid ? wt + (long long)(cbase + r) * Kaug + c * 8 : wt, valid);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch_w(0, 0);

    // ---- per-thread row bookkeeping (TMEM lane = halo raster position)
    const int q4 = warp & 3, half = warp >> 2;
    uint32_t e_off[3];            // byte offset of this row inside a plane
    bool e_inside[3], e_valid[3];
    long long o_off[3];           // element offset of this row's output pixel in `out` (channel 0), or -1
#pragma unroll
    for (int mt = 0; mt < 3; ++mt) {
        const int r = mt * BM + q4 * 32 + lane;
        const int ty = r / p.IW, tx = r - ty * p.IW;
        const int iy = iy0 + ty, ix = ix0 + tx;
        e_valid[mt] = mt < p.mtiles && r < npix;
        e_inside[mt] = e_valid[mt] && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Hin;
        e_off[mt] = (uint32_t)r * 16;
        // row r is the top-left tap of output (ty/S, tx/S) when both are multiples of S and inside the tile
        const int oyl = ty / S, oxl = tx / S;
        const bool is_out = mt < p.mtiles && r < npix && (ty - oyl * S) == 0 && (tx - oxl * S) == 0 && oyl < p.TH && oxl < p.TW &&
                            ty0 + oyl < p.Ho && tx0 + oxl < p.Ho;
        o_off[mt] = is_out ? (((long long)n * p.Ho + ty0 + oyl) * p.Ho + tx0 + oxl) * p.Cexp : -1;
    }
    const uint32_t dw_col0 = (uint32_t)(p.mtiles * p.CC);      // TMEM columns of the depthwise accumulators

    auto issue_expand = [&](int buf) {
        const int ksteps_total = p.cpr >> 1;
        for (int mt = 0; mt < p.mtiles; ++mt)
            for (int ks = 0; ks < ksteps_total; ++ks) {
                const int kb = ks >> 2, k = ks & 3;
                const uint64_t ad = tc::make_desc(sA + (uint32_t)kb * rows_total * 128 + (uint32_t)mt * BM * 128);
                const uint64_t bd = tc::make_desc(sW + buf * p.smem_W + (uint32_t)kb * p.CC * 128);
                tc::umma_f16(s_tmem_base + (uint32_t)(mt * p.CC), ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), p.idesc_exp, ks ? 1u : 0u);
            }
        tc::umma_commit(&mbar_exp);
    };
    auto issue_depthwise = [&]() {
        for (int mt = 0; mt < p.mtiles; ++mt)
            for (int j = 0; j < units; ++j) {
                const uint32_t a_base = sE + (uint32_t)(2 * j) * plane_bytes + (uint32_t)(mt * BM) * 16;
#pragma unroll
                for (int t = 0; t < TAPS; ++t) {
                    const int ky = t / KS, kx = t - ky * KS;
                    const uint64_t ad = make_desc_nosw(a_base + (uint32_t)(ky * p.IW + kx) * 16, plane_bytes, 128);
                    const uint64_t bd = make_desc_nosw(sB + (uint32_t)(t * units + j) * 512, 256, 128);
                    tc::umma_f16(s_tmem_base + dw_col0 + (uint32_t)(mt * p.CC + j * 16), ad, bd, p.idesc_dw, t ? 1u : 0u);
                }
            }
        tc::umma_commit(&mbar_dw);
    };

    asm volatile("cp.async.wait_all;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = s_tmem_base;
    if (tid == 0) issue_expand(0);

    for (int ch = 0; ch < p.n_chunks; ++ch) {
        const int buf = ch & 1;
        const int cbase = ch * p.CC;
        if (ch + 1 < p.n_chunks) prefetch_w(ch + 1, buf ^ 1);
        // ---- this chunk's diagonal weights + shift (previous chunk's depthwise MMAs have completed: epilogue 2 waited)
        for (int idx = tid; idx < TAPS * p.CC; idx += 256) {
            const int t = idx / p.CC, c = idx - t * p.CC;
            const int j = c >> 4, nn = c & 15;
            const uint16_t bits = to_bits16<T>(p.w_dw[(long long)t * p.Cexp + cbase + c]);
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(sB + (uint32_t)(t * units + j) * 512 + (uint32_t)(nn >> 3) * 256 + (uint32_t)nn * 16 + (uint32_t)(nn & 7) * 2),
                         "h"(bits) : "memory");
        }
        if (tid < p.CC) {
            const float b = p.b_dw[cbase + tid];
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(sBias + (uint32_t)tid * 4), "f"(b) : "memory");
        }
        if (!tc::mbar_wait(&mbar_exp, ch & 1)) s_abort = 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        bool ok = !s_abort;

        // ---- epilogue 1: expand accumulators -> swish -> E planes
        if (ok) {
#pragma unroll
            for (int mt = 0; mt < 3; ++mt) {
                if (mt < p.mtiles) {
                    for (int u = (mt * units + half) & 1; u < units; u += 2) {
                        float v[16];
                        tc::tmem_ld16(tmem_d + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(mt * p.CC + u * 16), v);
                        if (e_valid[mt]) {
                            uint4 lo, hi;
                            if (e_inside[mt]) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) v[j] = swish_from_half(v[j]);
                                lo = make_uint4(pack2<T>(v[0], v[1]), pack2<T>(v[2], v[3]), pack2<T>(v[4], v[5]), pack2<T>(v[6], v[7]));
                                hi = make_uint4(pack2<T>(v[8], v[9]), pack2<T>(v[10], v[11]), pack2<T>(v[12], v[13]), pack2<T>(v[14], v[15]));
                            } else {
                                lo = make_uint4(0u, 0u, 0u, 0u);
                                hi = lo;
                            }
                            sts128(sE + (uint32_t)(2 * u) * plane_bytes + e_off[mt], lo);
                            sts128(sE + (uint32_t)(2 * u + 1) * plane_bytes + e_off[mt], hi);
                        }
                    }
                }
            }
        }
        // E planes, diagonal matrices and W(ch+1) are in place -> depthwise MMAs, then the next chunk's expand MMAs
        asm volatile("cp.async.wait_all;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0 && !s_abort) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue_depthwise();
            if (ch + 1 < p.n_chunks) issue_expand(buf ^ 1);
        }
        // squeeze scratch of this chunk
        for (int i = tid; i < 8 * p.CC; i += 256) asm volatile("st.shared.f32 [%0], %1;" ::"r"(sR + (uint32_t)i * 4), "f"(0.f) : "memory");
        if (!tc::mbar_wait(&mbar_dw, ch & 1)) s_abort = 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        __syncthreads();
        ok = !s_abort;

        // ---- epilogue 2: depthwise accumulators -> +shift, swish -> store real output rows, squeeze sums
        if (ok) {
#pragma unroll
            for (int mt = 0; mt < 3; ++mt) {
                if (mt < p.mtiles) {
                    for (int u = (mt * units + half) & 1; u < units; u += 2) {
                        float v[16];
                        tc::tmem_ld16(tmem_d + ((uint32_t)(q4 * 32) << 16) + dw_col0 + (uint32_t)(mt * p.CC + u * 16), v);
                        const bool row_out = o_off[mt] >= 0;
                        if (row_out) {
                            const float4 b0 = lds_f4(sBias + (uint32_t)(u * 16) * 4), b1 = lds_f4(sBias + (uint32_t)(u * 16 + 4) * 4);
                            const float4 b2 = lds_f4(sBias + (uint32_t)(u * 16 + 8) * 4), b3 = lds_f4(sBias + (uint32_t)(u * 16 + 12) * 4);
                            const float bb[16] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w, b3.x, b3.y, b3.z, b3.w};
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = swish_from_half(v[j] + bb[j]);
                            T* dst = out + o_off[mt] + cbase + u * 16;
                            *reinterpret_cast<uint4*>(dst) = make_uint4(pack2<T>(v[0], v[1]), pack2<T>(v[2], v[3]), pack2<T>(v[4], v[5]), pack2<T>(v[6], v[7]));
                            *reinterpret_cast<uint4*>(dst + 8) = make_uint4(pack2<T>(v[8], v[9]), pack2<T>(v[10], v[11]), pack2<T>(v[12], v[13]), pack2<T>(v[14], v[15]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = 0.f;
                        }
                        // transposed butterfly: 16 values x 32 lanes -> one channel total per lane (16 shuffles)
                        float w8[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float mine = (lane & 16) ? v[8 + i] : v[i], theirs = (lane & 16) ? v[i] : v[8 + i];
                            w8[i] = mine + __shfl_xor_sync(0xffffffffu, theirs, 16);
                        }
                        float w4[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float mine = (lane & 8) ? w8[4 + i] : w8[i], theirs = (lane & 8) ? w8[i] : w8[4 + i];
                            w4[i] = mine + __shfl_xor_sync(0xffffffffu, theirs, 8);
                        }
                        float w2[2];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const float mine = (lane & 4) ? w4[2 + i] : w4[i], theirs = (lane & 4) ? w4[i] : w4[2 + i];
                            w2[i] = mine + __shfl_xor_sync(0xffffffffu, theirs, 4);
                        }
                        float w1;
                        {
                            const float mine = (lane & 2) ? w2[1] : w2[0], theirs = (lane & 2) ? w2[0] : w2[1];
                            w1 = mine + __shfl_xor_sync(0xffffffffu, theirs, 2);
                        }
                        w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
                        if ((lane & 1) == 0) {
                            const int cch = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                            const uint32_t a = sR + (uint32_t)(warp * p.CC + u * 16 + cch) * 4;
                            float cur;
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(cur) : "r"(a));
                            cur += w1;
                            asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(cur) : "memory");
                        }
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (ok && tid < p.CC) {
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                float t;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(sR + (uint32_t)(w * p.CC + tid) * 4));
                tot += t;
            }
            p.partial[((long long)n * gridDim.x + tile) * p.Cexp + cbase + tid] = tot;
        }
        __syncthreads();     // the squeeze scratch and the diagonal matrices are rewritten at the top of the next chunk
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)p.tmem_cols) : "memory");
}

inline bool plan_k1t_candidate(int Hin, int Ho, int Cin, int Cexp, int k, int s, int pad, bool is_bf16, int TH, int TW, int CC,
                               K1TParams* p, size_t* smem_out) {
    if (Ho % TH || Ho % TW || Cexp % CC || (CC & 15)) return false;
    p->Hin = Hin; p->Ho = Ho; p->Cin = Cin; p->Cexp = Cexp; p->pad = pad;
    p->TH = TH; p->TW = TW;
    p->IH = (TH - 1) * s + k; p->IW = (TW - 1) * s + k;
    p->tiles_x = Ho / TW; p->tiles_y = Ho / TH;
    p->mtiles = (p->IH * p->IW + BM - 1) / BM;
    if (p->mtiles > 3 || 2 * p->mtiles * CC > 512) return false;
    p->cpr = ((Cin >> 3) + 1 + 1) & ~1;
    p->nkb = (p->cpr + 7) / 8;
    p->CC = CC; p->n_chunks = Cexp / CC;
    int cols = 32;
    while (cols < 2 * p->mtiles * CC) cols <<= 1;
    p->tmem_cols = cols;
    p->P = p->mtiles * BM + (k - 1) * p->IW + (k - 1) + 8;
    p->idesc_exp = tc::make_idesc(is_bf16, CC);
    p->idesc_dw = tc::make_idesc(is_bf16, 16);
    p->smem_A = p->nkb * p->mtiles * BM * 128;
    p->smem_W = ((p->nkb * CC * 128) + 1023) & ~1023;
    p->smem_E = (((CC / 8) * p->P * 16) + 1023) & ~1023;
    p->smem_B = ((k * k * (CC / 16) * 512) + 1023) & ~1023;
    *smem_out = (size_t)p->smem_A + 2 * p->smem_W + p->smem_E + p->smem_B + (size_t)CC * 4 + 8 * (size_t)CC * 4 + 1024;
    return *smem_out <= 200 * 1024;
}

inline bool plan_k1t(int Hin, int Ho, int Cin, int Cexp, int k, int s, int pad, bool is_bf16, K1TParams* p, size_t* smem_out) {
    struct Cand { int th, tw; };
    const Cand s1[] = {{14, 14}, {7, 14}, {7, 7}};
    const Cand s2[] = {{8, 8}, {7, 7}};
    const Cand* cands = s == 1 ? s1 : s2;
    const int ncand = s == 1 ? 3 : 2;
    bool found = false;
    double best = 0;
    for (int i = 0; i < ncand; ++i)
        for (int cc = 128; cc >= 16; cc -= 16) {
            K1TParams q{};
            size_t smem = 0;
            if (!plan_k1t_candidate(Hin, Ho, Cin, Cexp, k, s, pad, is_bf16, cands[i].th, cands[i].tw, cc, &q, &smem)) continue;
            const bool two = smem <= 110 * 1024 && q.tmem_cols <= 256;
            // cost per output element: both epilogues run over every halo row, fixed per-chunk overhead
            double cost = ((double)q.mtiles * BM * cc * 9.0 + 256.0 * 500.0) * q.n_chunks / ((double)cands[i].th * cands[i].tw * Cexp);
            if (!two) cost *= 1.5;
            if (!found || cost < best) { best = cost; *p = q; *smem_out = smem; found = true; }
        }
    return found;
}

template <typename T>
int launch_k1t(cudaStream_t stream, const K1TParams& p, int k, int s, size_t smem, int n_crops) {
    dim3 grid(p.tiles_x * p.tiles_y, n_crops);
#define K1T(KS, S)                                                                                               \
    do {                                                                                                         \
        auto kfn = k1t_kernel<T, KS, S>;                                                                         \
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -1; \
        kfn<<<grid, 256, smem, stream>>>(p);                                                                     \
        return 0;                                                                                                \
    } while (0)
    if (k == 3 && s == 1) K1T(3, 1);
    if (k == 3 && s == 2) K1T(3, 2);
    if (k == 5 && s == 1) K1T(5, 1);
    if (k == 5 && s == 2) K1T(5, 2);
#undef K1T
    return 1;
}

}  // namespace fused
}  // namespace whenet
