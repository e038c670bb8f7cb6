// kernels_tc.cuh - tcgen05 (5th-gen tensor core) kernels for the 1x1 convolutions + the helpers every tcgen05 kernel shares.
//
//   out[m, n] = act( bias[n] + sum_k (A[m,k] * gate[m/hw, k]) * Wt[n,k] ) (+ resid[m,n])
//
// A  : activations, NHWC == row-major [M = crops*H*W][K = Cin], 16-bit (bf16 / fp16)   -> "K-major"
// Wt : BN-folded weights transposed to [N = Cout][K], 16-bit                           -> "K-major"
// D  : fp32 accumulator in tensor memory (TMEM), 128 lanes (pixels) x n_tile columns
//
// Operands sit in shared memory in the canonical K-major SWIZZLE_128B UMMA layout (16-byte chunk c of row r lands at
// chunk c ^ (r & 7) of its 128-byte row; 8-row atoms of 1024 B); one elected thread issues
// tcgen05.mma.cta_group::1.kind::f16 (M=128, N=n_tile, K=16), tcgen05.commit -> mbarrier hands stages back, and the
// epilogue reads TMEM lanes with tcgen05.ld (thread == pixel row).
//
// Every mbarrier wait is bounded: a wait that exceeds its budget raises the context's timeout flag (mapped pinned host
// memory) and the CTA bails out instead of hanging the GPU.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels_simt.cuh"

namespace whenet {
namespace tc {

// Timeout flag: ONE int per context in mapped pinned host memory (whenet_api.cu); every tcgen05 kernel gets its device
// address as a parameter and raises it when a bounded mbarrier wait expires.  The host reads it straight from the pinned
// page after any stream synchronisation - no per-translation-unit device symbols, no extra copies.

constexpr int BM = 128;          // pixels per CTA == TMEM lanes == UMMA M
constexpr int BK = 64;           // channels per stage (one 128-byte swizzle row of 16-bit elements)
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// bounded parity wait; returns false on timeout
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* tflag) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t it = 0; it < (1u << 22); ++it) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return true;
    }
    *reinterpret_cast<volatile int*>(tflag) = 1;
    return false;
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 = 1 | [32,46) SBO >> 4 = 64 (8 rows x 128 B)
//   [46,48) version = 1 (Blackwell) | [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)64 << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor): D=f32, A/B = bf16 or f16, both K-major.
__host__ __device__ inline uint32_t make_idesc(bool is_bf16, int umma_n, int b_fmt = -1) {
    uint32_t d = 0;
    d |= 1u << 4;                           // c_format = F32
    d |= (is_bf16 ? 1u : 0u) << 7;          // a_format
    d |= ((b_fmt < 0 ? is_bf16 : b_fmt != 0) ? 1u : 0u) << 10;         // b_format (b_fmt: -1 = as A, 0 = f16, 1 = bf16)
    d |= (uint32_t)(umma_n >> 3) << 17;     // n_dim
    d |= (uint32_t)(BM >> 4) << 24;         // m_dim
    return d;
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

template <typename T> __device__ __forceinline__ uint4 scale8(uint4 raw, const float* g);
template <> __device__ __forceinline__ uint4 scale8<__nv_bfloat16>(uint4 raw, const float* g) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
    const float4 g0 = *reinterpret_cast<const float4*>(g), g1 = *reinterpret_cast<const float4*>(g + 4);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        h[i] = __floats2bfloat162_rn(f.x * gg[2 * i], f.y * gg[2 * i + 1]);
    }
    return raw;
}
template <> __device__ __forceinline__ uint4 scale8<__half>(uint4 raw, const float* g) {
    __half2* h = reinterpret_cast<__half2*>(&raw);
    const float4 g0 = *reinterpret_cast<const float4*>(g), g1 = *reinterpret_cast<const float4*>(g + 4);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __half22float2(h[i]);
        h[i] = __floats2half2_rn(f.x * gg[2 * i], f.y * gg[2 * i + 1]);
    }
    return raw;
}

// ----------------------------------------------------------------------------- pw_tc2: cp.async ring
// pw_tc2_kernel: one 128-pixel x n_tile(<=256) output tile per CTA, K in blocks of 64 channels:
//   * operand K blocks travel global -> shared with cp.async (16 B, zero-fill for tails) through a ring of
//     n_stages stages, several blocks in flight, no register staging;
//   * the SE gate is applied IN shared memory by the thread that copied the chunk (so no extra barrier):
//       GATE == 1: on the A rows (tiles may span up to four crops; their gate rows sit in smem)
//       GATE == 2: on the W rows (per-crop tiling: a tile never leaves its crop, used while H*W >= 784)
//   * stage reuse is gated by the mbarrier of the previous block's tcgen05.commit.
// exact floor(x / d) for small non-negative ints (x < 2^17, d < 2^8) with inv = 1.0f / d: (x + 0.5) / d is at least 0.5 / d
// away from every integer, far more than the float rounding error - replaces the ~20-instruction integer division
__device__ __forceinline__ int fdiv_small(int x, float inv) { return __float2int_rz(((float)x + 0.5f) * inv); }
__device__ __forceinline__ void cp_async16_z(uint32_t dst, const void* src, bool valid) {
    const uint32_t sz = valid ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128_(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <typename T> __device__ __forceinline__ uint4 scale8s(uint4 raw, uint32_t gaddr);   // gate values from smem
template <> __device__ __forceinline__ uint4 scale8s<__nv_bfloat16>(uint4 raw, uint32_t gaddr) {
    float4 g0, g1;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(g0.x), "=f"(g0.y), "=f"(g0.z), "=f"(g0.w) : "r"(gaddr));
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(g1.x), "=f"(g1.y), "=f"(g1.z), "=f"(g1.w) : "r"(gaddr + 16));
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        h[i] = __floats2bfloat162_rn(f.x * gg[2 * i], f.y * gg[2 * i + 1]);
    }
    return raw;
}
template <> __device__ __forceinline__ uint4 scale8s<__half>(uint4 raw, uint32_t gaddr) {
    float4 g0, g1;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(g0.x), "=f"(g0.y), "=f"(g0.z), "=f"(g0.w) : "r"(gaddr));
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(g1.x), "=f"(g1.y), "=f"(g1.z), "=f"(g1.w) : "r"(gaddr + 16));
    __half2* h = reinterpret_cast<__half2*>(&raw);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __half22float2(h[i]);
        h[i] = __floats2half2_rn(f.x * gg[2 * i], f.y * gg[2 * i + 1]);
    }
    return raw;
}

// OUT_H: the result is written as fp16 whatever T is (the expand conv feeding the HFMA2 depthwise kernel KD)
template <typename T, bool SWISH, int GATE, bool RESID, bool OUT_H = false>
__global__ void __launch_bounds__(128) pw_tc2_kernel(const T* __restrict__ A, const T* __restrict__ Wt,
                                                     const float* __restrict__ bias, const float* __restrict__ gate,
                                                     const T* __restrict__ resid, T* __restrict__ out,
                                                     int M, int K, int N, int hw,
                                                     int n_tile, int umma_n, int tmem_cols, int n_stages,
                                                     int tiles_per_crop,     // GATE == 2 only
                                                     uint32_t idesc, int* tflag) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t mbar[4];
    __shared__ uint32_t s_tmem_base;
    __shared__ int s_abort;

    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int w_stage_bytes = umma_n * BK * 2;
    const uint32_t stage_bytes = A_STAGE_BYTES + w_stage_bytes;
    const uint32_t sG = smem0 + n_stages * stage_bytes;          // gate rows: [<=4 crops][K] fp32 (GATE only)

    // ---- tile -> rows
    int m0, rows_valid, crop0;
    if (GATE == 2) {
        const int crop = blockIdx.y / tiles_per_crop, t = blockIdx.y - crop * tiles_per_crop;
        m0 = crop * hw + t * BM;
        rows_valid = min(BM, hw - t * BM);
        crop0 = crop;
    } else {
        m0 = blockIdx.y * BM;
        rows_valid = min(BM, M - m0);
        crop0 = GATE ? m0 / hw : 0;
    }
    const int n0 = blockIdx.x * n_tile;
    const int n_valid = min(n_tile, N - n0);
    const int nkb = (K + BK - 1) / BK;
    const int kchunks = K >> 3;

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&mbar[i], 1);
        s_abort = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"((uint32_t)tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // gate rows of the crops this tile touches -> smem (joins the first cp.async group)
    if (GATE) {
        const int ncrops = GATE == 2 ? 1 : ((m0 + rows_valid - 1) / hw - crop0 + 1);
        const int q = K >> 2;
        for (int idx = tid; idx < ncrops * q; idx += 128) {
            const int cr = idx / q, j = idx - cr * q;
            cp_async16_z(sG + (uint32_t)(cr * K + j * 4) * 4, gate + (long long)(crop0 + cr) * K + j * 4, true);
        }
    }
    auto fill = [&](int kb) {
        const int s = kb % n_stages;
        const uint32_t a_st = smem0 + s * stage_bytes, w_st = a_st + A_STAGE_BYTES;
        const int kc0 = kb * 8;
        const int cb = min(8, kchunks - kc0), cbp = (cb + 1) & ~1;
        if (cbp == 8) {
            // full block: item idx = tid + i*128 -> row tid/8 + 16 i, chunk tid%8 (no divisions in the hot loop)
            const int c = tid & 7, r0 = tid >> 3;
            const bool cvalid = c < cb;
            const T* asrc = A + (long long)(m0 + r0) * K + (kc0 + c) * 8;
            const uint32_t swz = (uint32_t)((r0 >> 3) * 1024 + (r0 & 7) * 128 + ((c ^ (r0 & 7)) << 4));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool valid = cvalid && (r0 + 16 * i) < rows_valid;
                cp_async16_z(a_st + swz + i * 2048, valid ? asrc + (long long)i * 16 * K : A, valid);
            }
            const T* wsrc = Wt + (long long)(n0 + r0) * K + (kc0 + c) * 8;
            for (int r = r0, i = 0; r < umma_n; r += 16, ++i) {
                const bool valid = cvalid && r < n_valid;
                cp_async16_z(w_st + swz + i * 2048, valid ? wsrc + (long long)i * 16 * K : Wt, valid);
            }
            return;
        }
        for (int idx = tid; idx < BM * cbp; idx += 128) {
            const int r = fdiv_small(idx, 1.0f / (float)cbp), c = idx - r * cbp;
            const bool valid = r < rows_valid && c < cb;
            cp_async16_z(a_st + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4),
                         valid ? A + (long long)(m0 + r) * K + (kc0 + c) * 8 : A, valid);
        }
        for (int idx = tid; idx < umma_n * cbp; idx += 128) {
            const int r = fdiv_small(idx, 1.0f / (float)cbp), c = idx - r * cbp;
            const bool valid = r < n_valid && c < cb;
            cp_async16_z(w_st + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4),
                         valid ? Wt + (long long)(n0 + r) * K + (kc0 + c) * 8 : Wt, valid);
        }
    };
    for (int j = 0; j < n_stages; ++j) {
        if (j < nkb) fill(j);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = s_tmem_base;

    // GATE == 1: byte offset of the gate row (crop) of each of the 8 rows this thread rescales, fixed for the whole K loop
    uint32_t g_row[8];
    if (GATE == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = (tid >> 3) + 16 * i;
            g_row[i] = r < rows_valid ? (uint32_t)(((m0 + r) / hw - crop0) * K) * 4u : 0u;
        }
    }
    for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % n_stages;
        const uint32_t a_st = smem0 + s * stage_bytes, w_st = a_st + A_STAGE_BYTES;
        // groups committed so far: n_stages + kb ; block kb sits in group kb (kb < n_stages) or kb+1 -> allow n_stages-2 pending
        if (n_stages >= 4) asm volatile("cp.async.wait_group 2;" ::: "memory");
        else if (n_stages == 3) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (GATE) {
            // the gate rows in sG were copied by ALL threads (group 0): one CTA barrier before their first use.
            // The operand chunks themselves need none: each thread rescales exactly the chunks it copied itself.
            if (kb == 0) __syncthreads();
            const int kc0 = kb * 8;
            const int cb = min(8, kchunks - kc0), cbp = (cb + 1) & ~1;
            if (GATE == 1) {
                if (cbp == 8) {
                    const int c = tid & 7, r0 = tid >> 3;
                    if (c < cb) {
                        const uint32_t a0 = a_st + (r0 >> 3) * 1024 + (r0 & 7) * 128 + ((c ^ (r0 & 7)) << 4);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (r0 + 16 * i < rows_valid)
                                sts128_(a0 + i * 2048, scale8s<T>(lds128(a0 + i * 2048), sG + g_row[i] + (uint32_t)((kc0 + c) * 8) * 4));
                        }
                    }
                } else {
                    for (int idx = tid; idx < BM * cbp; idx += 128) {
                        const int r = fdiv_small(idx, 1.0f / (float)cbp), c = idx - r * cbp;
                        if (r < rows_valid && c < cb) {
                            const uint32_t addr = a_st + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4);
                            const int cr = (m0 + r) / hw - crop0;
                            sts128_(addr, scale8s<T>(lds128(addr), sG + (uint32_t)(cr * K + (kc0 + c) * 8) * 4));
                        }
                    }
                }
            } else {
                if (cbp == 8) {
                    const int c = tid & 7, r0 = tid >> 3;
                    if (c < cb)
                        for (int r = r0, i = 0; r < n_valid; r += 16, ++i) {
                            const uint32_t addr = w_st + (r0 >> 3) * 1024 + (r0 & 7) * 128 + ((c ^ (r0 & 7)) << 4) + i * 2048;
                            sts128_(addr, scale8s<T>(lds128(addr), sG + (uint32_t)((kc0 + c) * 8) * 4));
                        }
                } else {
                    for (int idx = tid; idx < umma_n * cbp; idx += 128) {
                        const int r = fdiv_small(idx, 1.0f / (float)cbp), c = idx - r * cbp;
                        if (r < n_valid && c < cb) {
                            const uint32_t addr = w_st + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4);
                            sts128_(addr, scale8s<T>(lds128(addr), sG + (uint32_t)((kc0 + c) * 8) * 4));
                        }
                    }
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0 && !s_abort) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int krem = min(BK, K - kb * BK);
            const int ksteps = (krem + 15) >> 4;
            const uint64_t ad = make_desc(a_st), bd = make_desc(w_st);
            for (int k = 0; k < ksteps; ++k)
                umma_f16(tmem_d, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
            umma_commit(&mbar[s]);
        }
        // refill the stage block kb-1 used (its MMAs are done or about to be) with block kb-1+n_stages
        if (kb >= 1 && kb - 1 + n_stages < nkb) {
            const int pb = kb - 1;
            if (!mbar_wait(&mbar[pb % n_stages], (pb / n_stages) & 1, tflag)) s_abort = 1;
            fill(pb + n_stages);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    {
        const int last = nkb - 1;
        if (!mbar_wait(&mbar[last % n_stages], (last / n_stages) & 1, tflag)) s_abort = 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();

    // ---- epilogue: TMEM -> +shift, swish, +residual -> 16-bit -> stage -> coalesced stores
    const int nch = n_valid >> 3;
    const float inv_nch = 1.0f / (float)(nch > 0 ? nch : 1);
    const int pitch16 = nch | 1;
    uint4* stage = reinterpret_cast<uint4*>(smem_raw + (smem0 - smem_u32(smem_raw)));
    const bool row_ok = tid < rows_valid;
    const long long m = (long long)m0 + tid;
    if (!s_abort) {
        const uint32_t lane_base = tmem_d + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < n_valid; c0 += 16) {
            float v[16];
            tmem_ld16(lane_base + (uint32_t)c0, v);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = n0 + c0 + h * 8;
                if (c0 + h * 8 >= n_valid) break;
                float o[8];
                const float4 b0 = *reinterpret_cast<const float4*>(bias + n), b1 = *reinterpret_cast<const float4*>(bias + n + 4);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float x = v[h * 8 + j] + bb[j];
                    o[j] = SWISH ? swish_fast(x) : x;
                }
                if (RESID && row_ok) {
                    float r[8];
                    ld8<T>(resid + m * N + n, r);
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] += r[j];
                }
                if (OUT_H) st8<__half>(reinterpret_cast<__half*>(stage + tid * pitch16 + ((c0 >> 3) + h)), o);
                else st8<T>(reinterpret_cast<T*>(stage + tid * pitch16 + ((c0 >> 3) + h)), o);
            }
        }
    }
    __syncthreads();
    if (!s_abort) {
        for (int idx = tid; idx < rows_valid * nch; idx += 128) {
            const int r = fdiv_small(idx, inv_nch), j = idx - r * nch;
            *reinterpret_cast<uint4*>(out + ((long long)m0 + r) * N + n0 + j * 8) = stage[r * pitch16 + j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)tmem_cols) : "memory");
}

template <typename T>
int launch_pw_tc2(cudaStream_t stream, int* tflag, const T* A, const void* Wt16, const float* bias, const float* gate, const T* resid,
                  T* out, long long M, int K, int N, int hw, bool swish, int stage_cap = 0, int smem_budget_kb = 54, int min_ctas = 296,
                  bool out_half = false, int b_fmt = -1) {
    if (sizeof(T) != 2) return 1;
    if ((K & 7) || (N & 7) || M > 0x7fffffffLL) return 1;
    const bool per_crop = gate && hw >= 784;                 // gate on W, tiles stay inside a crop
    const int tpc = (hw + BM - 1) / BM;
    const long long m_tiles = per_crop ? (M / hw) * tpc : (M + BM - 1) / BM;
    int n_tile = N;
    if (N > 256) {
        int parts = (N + 255) / 256;
        while (true) {
            n_tile = ((N + parts - 1) / parts + 15) & ~15;
            if (n_tile <= 256) break;
            ++parts;
        }
    }
    while (n_tile > 48 && m_tiles * ((N + n_tile - 1) / n_tile) < min_ctas) {
        const int parts = (N + n_tile - 1) / n_tile + 1;
        const int nt = ((N + parts - 1) / parts + 15) & ~15;
        if (nt >= n_tile) break;
        n_tile = nt;
    }
    const int umma_n = (n_tile + 15) & ~15;
    int tmem_cols = 32;
    while (tmem_cols < umma_n) tmem_cols <<= 1;
    const uint32_t idesc = make_idesc(std::is_same<T, __nv_bfloat16>::value, umma_n, b_fmt);
    const int nkb = (K + BK - 1) / BK;
    const size_t stage_bytes = A_STAGE_BYTES + (size_t)umma_n * BK * 2;
    const int gate_crops = per_crop ? 1 : std::min(4, (BM - 1) / hw + 2);      // crops one 128-row tile can touch
    const size_t gate_bytes = gate ? (size_t)gate_crops * K * 4 : 0;
    const size_t out_bytes = (size_t)BM * ((size_t)(n_tile >> 3) | 1) * 16;
    int n_stages = nkb < 4 ? nkb : 4;
    if (stage_cap > 0 && n_stages > stage_cap) n_stages = stage_cap;
    // a grid that does not even fill the SMs once (single-crop latency path) gains nothing from co-residency: deepest ring
    if (m_tiles * ((N + n_tile - 1) / n_tile) < 148) smem_budget_kb = 180;
    // ring depth vs co-residency: a shallower ring lets more CTAs share the SM (budget = smem per CTA)
    while (n_stages > 2 && n_stages * stage_bytes + gate_bytes > (size_t)smem_budget_kb * 1024) --n_stages;
    size_t smem = n_stages * stage_bytes + gate_bytes;
    if (smem < out_bytes) smem = out_bytes;
    smem += 1024;
    if (smem > 225 * 1024) return 1;
    dim3 grid((unsigned)((N + n_tile - 1) / n_tile), (unsigned)m_tiles);
    const T* W = reinterpret_cast<const T*>(Wt16);
#define TC2(SW, GA, RE, OH)                                                                                          \
    do {                                                                                                             \
        auto kfn = pw_tc2_kernel<T, SW, GA, RE, OH>;                                                       \
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024) != cudaSuccess) return -1; \
        kfn<<<grid, 128, smem, stream>>>(A, W, bias, gate, resid, out, (int)M, K, N, hw, n_tile, umma_n, tmem_cols, n_stages, tpc, idesc, tflag); \
    } while (0)
    if (out_half && !(swish && !gate && !resid)) return 1;
    if (swish && !gate && !resid) { if (out_half) TC2(true, 0, false, true); else TC2(true, 0, false, false); }
    else if (!swish && !gate && !resid) TC2(false, 0, false, false);
    else if (!swish && !gate && resid) TC2(false, 0, true, false);           // project conv whose input K1 has already gated
    else if (!swish && gate && !resid) { if (per_crop) TC2(false, 2, false, false); else TC2(false, 1, false, false); }
    else if (!swish && gate && resid) { if (per_crop) TC2(false, 2, true, false); else TC2(false, 1, true, false); }
    else return 1;
#undef TC2
    return 0;
}


// ----------------------------------------------------------------------------- pw_tc3: gated projects of the large maps
// The gated project convs of blocks 1-5 are streams of tiny GEMMs (K = 32..144, N = 16..40, millions of rows): with one 128-row
// tile per CTA (pw_tc2) every tile pays TMEM allocation, the gate row, the W copy + its rescale and a cold load -> MMA ->
// epilogue chain - block 1 ran at 2.2 TB/s.  Here a CTA walks `tpc` consecutive tiles of ONE crop: gate row and W' = bf16(W * g)
// once per CTA (the same scale8s as pw_tc2's per-crop route, so the results are bit-identical), then a two-deep software
// pipeline over the tiles: cp.async of tile t+1 and the MMA of tile t (second TMEM accumulator) run under the epilogue of
// tile t-1.
template <typename T, bool RESID>
__global__ void __launch_bounds__(128) pw_tc3_kernel(const T* __restrict__ A, const T* __restrict__ Wt, const float* __restrict__ bias,
                                                     const float* __restrict__ gate, const T* __restrict__ resid, T* __restrict__ out,
                                                     int K, int N, int hw, int umma_n, int tmem_cols, int tiles_per_crop, int tpc, int groups,
                                                     uint32_t idesc, int* tflag) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ uint32_t s_tmem_base;
    __shared__ int s_abort;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int nkb = (K + BK - 1) / BK, kchunks = K >> 3;
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t w_bytes = (uint32_t)nkb * umma_n * 128, a_bytes = (uint32_t)nkb * A_STAGE_BYTES;
    const uint32_t sW = smem0, sA = sW + ((w_bytes + 1023u) & ~1023u), sG = sA + 2 * a_bytes;
    const int nch = N >> 3, pitch16 = nch | 1;
    uint4* stage = reinterpret_cast<uint4*>(smem_raw + (sG + (((uint32_t)K * 4 + 15u) & ~15u) - smem_u32(smem_raw)));
    const float inv_nch = 1.0f / (float)nch;

    const int crop = blockIdx.x / groups, g = blockIdx.x - crop * groups;
    const int t_begin = g * tpc, t_end = min(tiles_per_crop, t_begin + tpc);
    if (t_begin >= t_end) return;

    if (tid == 0) {
        mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1);
        s_abort = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"((uint32_t)tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // one K block of an operand: 16-byte chunk c of rows r0, r0+16, ... (the mapping pw_tc2 uses; ragged last block: zero fill)
    auto fill_rows = [&](uint32_t dst, const T* src0, int rows, int kb) {
        const int kc0 = kb * 8;
        const int cb = min(8, kchunks - kc0);
        const int c = tid & 7, r0 = tid >> 3;
        const uint32_t swz = (uint32_t)((r0 >> 3) * 1024 + (r0 & 7) * 128 + ((c ^ (r0 & 7)) << 4));
        const T* src = src0 + (long long)r0 * K + (kc0 + c) * 8;
        for (int r = r0, i = 0; r < rows; r += 16, ++i)
            cp_async16_z(dst + swz + i * 2048, c < cb ? src + (long long)i * 16 * K : src0, c < cb);
    };
    auto fill_a = [&](int t, int buf) {
        const int rows_valid = min(BM, hw - t * BM);
        const T* a0 = A + ((long long)crop * hw + (long long)t * BM) * K;
        for (int kb = 0; kb < nkb; ++kb) {
            const int kc0 = kb * 8;
            const int cb = min(8, kchunks - kc0);
            const int c = tid & 7, r0 = tid >> 3;
            const uint32_t dst = sA + buf * a_bytes + kb * A_STAGE_BYTES + (uint32_t)((r0 >> 3) * 1024 + (r0 & 7) * 128 + ((c ^ (r0 & 7)) << 4));
            const T* src = a0 + (long long)r0 * K + (kc0 + c) * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool valid = c < cb && (r0 + 16 * i) < rows_valid;
                cp_async16_z(dst + i * 2048, valid ? src + (long long)i * 16 * K : A, valid);
            }
        }
    };
    {   // gate row of this crop, the whole W, the first A tile
        const int q = K >> 2;
        for (int idx = tid; idx < q; idx += 128) cp_async16_z(sG + (uint32_t)idx * 16, gate + (long long)crop * K + idx * 4, true);
        for (int kb = 0; kb < nkb; ++kb) fill_rows(sW + (uint32_t)kb * umma_n * 128, Wt, umma_n <= N ? umma_n : N, kb);
        if (umma_n > N) {       // rows N..umma_n-1 of W (padding of the MMA's N): zeros
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            for (int kb = 0; kb < nkb; ++kb)
                for (int idx = tid; idx < (umma_n - N) * 8; idx += 128) {
                    const int r = N + (idx >> 3), c = idx & 7;
                    sts128_(sW + (uint32_t)kb * umma_n * 128 + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)), z);
                }
        }
        fill_a(t_begin, 0);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                         // gate row (copied by all threads) visible
        // W' = 16-bit(W * g): each thread rescales exactly the chunks it copied
        for (int kb = 0; kb < nkb; ++kb) {
            const int kc0 = kb * 8;
            const int cb = min(8, kchunks - kc0);
            const int c = tid & 7, r0 = tid >> 3;
            if (c < cb)
                for (int r = r0, i = 0; r < N; r += 16, ++i) {
                    const uint32_t addr = sW + (uint32_t)kb * umma_n * 128 + (uint32_t)((r0 >> 3) * 1024 + (r0 & 7) * 128 + ((c ^ (r0 & 7)) << 4)) + i * 2048;
                    sts128_(addr, scale8s<T>(lds128(addr), sG + (uint32_t)((kc0 + c) * 8) * 4));
                }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = s_tmem_base;

    auto epilogue = [&](int t, int buf, int it) {
        if (!mbar_wait(&mbar[buf], (uint32_t)(it >> 1) & 1u, tflag)) s_abort = 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int rows_valid = min(BM, hw - t * BM);
        const long long m0 = (long long)crop * hw + (long long)t * BM;
        const bool row_ok = tid < rows_valid;
        if (!s_abort) {
            const uint32_t lane_base = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * umma_n);
            for (int c0 = 0; c0 < N; c0 += 16) {
                float v[16];
                tmem_ld16(lane_base + (uint32_t)c0, v);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int n = c0 + h * 8;
                    if (n >= N) break;
                    float o[8];
                    const float4 b0 = *reinterpret_cast<const float4*>(bias + n), b1 = *reinterpret_cast<const float4*>(bias + n + 4);
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = v[h * 8 + j] + bb[j];
                    if (RESID && row_ok) {
                        float r[8];
                        ld8<T>(resid + (m0 + tid) * N + n, r);
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] += r[j];
                    }
                    st8<T>(reinterpret_cast<T*>(stage + tid * pitch16 + ((c0 >> 3) + h)), o);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (!s_abort)
            for (int idx = tid; idx < rows_valid * nch; idx += 128) {
                const int r = fdiv_small(idx, inv_nch), j = idx - r * nch;
                *reinterpret_cast<uint4*>(out + (m0 + r) * N + j * 8) = stage[r * pitch16 + j];
            }
    };

    const int ntiles = t_end - t_begin;
    for (int it = 0; it < ntiles; ++it) {
        const int t = t_begin + it, buf = it & 1;
        asm volatile("cp.async.wait_group 0;" ::: "memory");            // A(t) has landed (this thread's part)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                                // ... everyone's; the store loop of tile t-2 is done with `stage`
        if (tid == 0 && !s_abort) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int kb = 0; kb < nkb; ++kb) {
                const int krem = min(BK, K - kb * BK);
                const int ksteps = (krem + 15) >> 4;
                const uint64_t ad = make_desc(sA + buf * a_bytes + kb * A_STAGE_BYTES), bd = make_desc(sW + (uint32_t)kb * umma_n * 128);
                for (int k = 0; k < ksteps; ++k)
                    umma_f16(tmem_d + (uint32_t)(buf * umma_n), ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
            }
            umma_commit(&mbar[buf]);
        }
        if (it >= 1) {
            // tile t-1: its MMA has finished (mbar) -> its A buffer is free for tile t+1, its accumulator is ready
            if (!mbar_wait(&mbar[buf ^ 1], (uint32_t)((it - 1) >> 1) & 1u, tflag)) s_abort = 1;
        }
        if (it + 1 < ntiles) fill_a(t + 1, buf ^ 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (it >= 1) epilogue(t - 1, buf ^ 1, it - 1);
    }
    __syncthreads();            // the store loop of tile n-2 is done with `stage` (inside the loop the barrier at the top does this)
    epilogue(t_end - 1, (ntiles - 1) & 1, ntiles - 1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)tmem_cols) : "memory");
}

// Which gated projects pw_tc3 takes and how it walks them (host-side, tested without a GPU by tests/test_route_plans.py).
struct Pw3Plan { int tiles_per_crop, tpc, groups, umma_n, tmem_cols; size_t smem; };
inline bool plan_pw_tc3(long long M, int K, int N, int hw, bool has_gate, Pw3Plan* pl) {
    // K <= 64 (one K block): measured b01 (K = 32) 0.275 -> 0.170 ms, b02 (K = 96) unchanged, b03 / b04 (K = 144: three blocks,
    // 96 KB of A buffers, two CTAs per SM) 1.7x SLOWER than pw_tc2
    if (!has_gate || hw < 784 || (K & 7) || (N & 7) || K > 64 || N > 64 || M < 1 || M % hw) return false;
    const int crops = (int)(M / hw);
    pl->tiles_per_crop = (hw + BM - 1) / BM;
    int groups = (1536 + crops - 1) / crops;                      // enough CTAs for ~10 per SM
    if (groups > pl->tiles_per_crop) groups = pl->tiles_per_crop;
    if (groups < 1) groups = 1;
    pl->tpc = (pl->tiles_per_crop + groups - 1) / groups;
    pl->groups = (pl->tiles_per_crop + pl->tpc - 1) / pl->tpc;
    if (pl->tpc < 3) return false;                                // nothing to pipeline
    pl->umma_n = (N + 15) & ~15;
    pl->tmem_cols = 32;
    while (pl->tmem_cols < 2 * pl->umma_n) pl->tmem_cols <<= 1;
    const int nkb = (K + BK - 1) / BK;
    const size_t w_bytes = ((size_t)nkb * pl->umma_n * 128 + 1023) & ~(size_t)1023;
    pl->smem = w_bytes + 2 * (size_t)nkb * A_STAGE_BYTES + (((size_t)K * 4 + 15) & ~(size_t)15) + (size_t)BM * ((size_t)(N >> 3) | 1) * 16 + 1024;
    return pl->smem <= 200 * 1024;
}

// 0 = launched; 1 = shape not covered (caller falls back to pw_tc2)
template <typename T>
int launch_pw_tc3(cudaStream_t stream, int* tflag, const T* A, const void* Wt16, const float* bias, const float* gate, const T* resid, T* out,
                  long long M, int K, int N, int hw) {
    Pw3Plan pl{};
    if (sizeof(T) != 2 || !plan_pw_tc3(M, K, N, hw, gate != nullptr, &pl)) return 1;
    const int crops = (int)(M / hw);
    const int tiles_per_crop = pl.tiles_per_crop, tpc = pl.tpc, groups = pl.groups, umma_n = pl.umma_n, tmem_cols = pl.tmem_cols;
    const size_t smem = pl.smem;
    const uint32_t idesc = make_idesc(std::is_same<T, __nv_bfloat16>::value, umma_n);
    const T* W = reinterpret_cast<const T*>(Wt16);
    if (resid) {
        auto kfn = pw_tc3_kernel<T, true>;
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -1;
        kfn<<<crops * groups, 128, smem, stream>>>(A, W, bias, gate, resid, out, K, N, hw, umma_n, tmem_cols, tiles_per_crop, tpc, groups, idesc, tflag);
    } else {
        auto kfn = pw_tc3_kernel<T, false>;
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -1;
        kfn<<<crops * groups, 128, smem, stream>>>(A, W, bias, gate, resid, out, K, N, hw, umma_n, tmem_cols, tiles_per_crop, tpc, groups, idesc, tflag);
    }
    return 0;
}

}  // namespace tc
}  // namespace whenet
