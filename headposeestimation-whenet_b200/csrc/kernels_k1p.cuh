// kernels_k1p.cuh - K1P: persistent, warp-specialised variant of K1 for the blocks with several tiles per crop (2..6).
//
// K1 (kernels_fused.cuh) gives one CTA one tile and runs its phases behind CTA-wide barriers: operand load -> MMA ->
// TMEM-to-E epilogue (SFU bound: one tanh per expanded element) -> depthwise (FMA / ld.shared bound) -> ...  ncu shows
// the early blocks at ~50 % issue utilisation with a quarter of the instructions spent on per-CTA set-up.  K1P keeps
// ONE 512-thread CTA per SM for the whole launch and gives the phases to different warps, connected by mbarriers:
//
//   warp 0        loader     W / depthwise constants of ALL chunks once (they fit: <= 30 KB + 25 KB), then the A rows of
//                            item after item (cp.async) into a 2-deep ring
//   warp 1        MMA        tcgen05.mma of chunk after chunk into a 2-deep TMEM accumulator ring, tcgen05.commit -> mbarrier
//   warps 2-5     epilogue   TMEM -> swish -> 16-bit E tile in a 2-deep smem ring (one warp per TMEM lane quadrant)
//   warps 6-15    depthwise  k x k FMA out of E, + shift, swish, store, squeeze partial sums (named barrier inside the group)
//
// so the SFU-heavy epilogue of chunk i+1 overlaps the FMA-heavy depthwise of chunk i, the operand load of the next tile
// overlaps both, and TMEM allocation / barrier set-up / weight loads happen once per SM instead of once per tile.
// An item is one (crop, tile); items are dealt round-robin to the CTAs.  Arithmetic and summation order per tile are
// those of K1, so the two variants agree to rounding of the squeeze sums' association only.
#pragma once
#include "kernels_fused.cuh"

namespace whenet {
namespace fused {

constexpr int K1P_THREADS = 512;
constexpr int K1P_EPI0 = 64;          // first epilogue thread (warp 2); the epilogue group has 4 or 8 warps, the depthwise the rest

struct K1PParams {
    K1Params k;          // geometry, operand pointers and chunking of K1 (NB == 1)
    int items;           // crops * tiles
    int tiles;
    int epi_warps;       // 4 (one per TMEM lane quadrant) or 8 (two per quadrant)
    int ndw;             // depthwise threads = 512 - 64 - 32 * epi_warps
    int PYp;             // strip lanes of the depthwise group = ndw / (CC/4)
    int tbuf_cols;       // TMEM columns of one accumulator buffer = mtiles * CC
    int tmem_cols;       // power of two >= 2 * tbuf_cols
    int smem_A1;         // one A buffer
    int smem_Wall;       // W of every chunk (n_chunks * smem_W)
    int smem_Call;       // depthwise constants of every chunk (n_chunks * smem_C)
    int smem_E1;         // one E buffer
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
// bounded wait that degrades to a no-op once any thread of the CTA has timed out: the roles keep their control flow
// (named barriers stay matched), the kernel finishes quickly with garbage, and the host sees the timeout flag
__device__ __forceinline__ void k1p_wait(uint64_t* bar, uint32_t parity, volatile int* abort_flag, int* tflag) {
    if (*abort_flag) return;
    const uint32_t addr = tc::smem_u32(bar);
    for (uint32_t it = 0; it < (1u << 20); ++it) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
        if ((it & 1023u) == 1023u && *abort_flag) return;
    }
    *abort_flag = 1;
    *reinterpret_cast<volatile int*>(tflag) = 1;
}

template <typename T, int KS, int S, int R>
__global__ void __launch_bounds__(K1P_THREADS, 1) k1p_kernel(const K1PParams pp) {
    const K1Params& p = pp.k;
    const int CC = p.CC, pitchE = p.pitchE;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_a_full[2], bar_a_empty[2], bar_t_full[2], bar_t_empty[2], bar_e_full[2], bar_e_empty[2], bar_const;
    __shared__ uint32_t s_tmem_base;
    __shared__ int s_abort_mem;
    volatile int* s_abort = &s_abort_mem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = smem0;                                   // 2 x [nkb][rows_alloc][128 B]            swizzled
    const uint32_t sW = sA + 2 * pp.smem_A1;                     // [n_chunks][nkb][CC rows][128 B]         swizzled
    const uint32_t sC = sW + pp.smem_Wall;                       // [n_chunks]{ b_dw[CC], w_dw[KS*KS][CC] } fp32
    const uint32_t sE = sC + pp.smem_Call;                       // 2 x [e_rows][pitchE]
    const uint32_t sR = sE + 2 * pp.smem_E1;                     // 2 x [PYp][CC] fp32 squeeze partials

    const T* in = reinterpret_cast<const T*>(p.in);
    const T* wt = reinterpret_cast<const T*>(p.wt_aug);
    T* out = reinterpret_cast<T*>(p.out);
    const int npix = p.IH * p.IW;
    const int kchunks = p.Cin >> 3;
    const int Kaug = p.Cin + 8;
    const uint32_t a_kb_stride = (uint32_t)p.rows_alloc * 128u;
    const int n_chunks = p.n_chunks;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&bar_a_full[i], 32);
            tc::mbar_init(&bar_a_empty[i], 1);
            tc::mbar_init(&bar_t_full[i], 1);
            tc::mbar_init(&bar_t_empty[i], 32 * pp.epi_warps);
            tc::mbar_init(&bar_e_full[i], 32 * pp.epi_warps);
            tc::mbar_init(&bar_e_empty[i], pp.ndw);
        }
        tc::mbar_init(&bar_const, 32);
        s_abort_mem = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&s_tmem_base)), "r"((uint32_t)pp.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem_base;

    // geometry of one item (crop n, tile): output-tile origin, input-tile origin, the part of the halo inside the image
    struct Geo { int n, tile, ty0, tx0, iy0, ix0, ty_lo, tx_lo, IHin, IWin, npix_in, mtc; };
    const float inv_tiles = 1.0f / (float)pp.tiles, inv_tx = 1.0f / (float)p.tiles_x;
    auto geom = [&](int item) {
        Geo g;
        g.n = div_small(item, inv_tiles);
        g.tile = item - g.n * pp.tiles;
        const int tyi = div_small(g.tile, inv_tx);
        g.ty0 = tyi * p.TH;
        g.tx0 = (g.tile - tyi * p.tiles_x) * p.TW;
        g.iy0 = g.ty0 * S - p.pad;
        g.ix0 = g.tx0 * S - p.pad;
        g.ty_lo = max(0, -g.iy0);
        g.tx_lo = max(0, -g.ix0);
        g.IHin = min(p.IH, p.Hin - g.iy0) - g.ty_lo;
        g.IWin = min(p.IW, p.Hin - g.ix0) - g.tx_lo;
        g.npix_in = g.IHin * g.IWin;
        g.mtc = (g.npix_in + BM - 1) / BM;
        return g;
    };

    if (warp == 0) {
        // =========================================================================== loader
        {   // W and the depthwise constants of every chunk: resident for the life of the CTA
            const float inv_cpr = 1.0f / (float)p.cpr, inv_CC = 1.0f / (float)CC;
            for (int idx = lane; idx < p.Cexp * p.cpr; idx += 32) {
                const int r = div_small(idx, inv_cpr), c = idx - r * p.cpr;         // expanded channel, 16-byte chunk of its row
                const int ch = div_small(r, inv_CC), rr = r - ch * CC;
                const bool valid = c <= kchunks;
                cp_async16(sW + (uint32_t)ch * p.smem_W + (uint32_t)(c >> 3) * CC * 128 + sw128(rr, c & 7),
                           valid ? wt + (long long)r * Kaug + c * 8 : wt, valid);
            }
            const int q = CC >> 2, rows = KS * KS + 1;
            const float inv_q = 1.0f / (float)q, inv_rows = 1.0f / (float)rows;
            for (int idx = lane; idx < n_chunks * rows * q; idx += 32) {
                const int rw = div_small(idx, inv_q), j = idx - rw * q;              // (chunk, row), 16-byte piece
                const int ch = div_small(rw, inv_rows), row = rw - ch * rows;
                const float* src = row == 0 ? p.b_dw + ch * CC + j * 4 : p.w_dw + (long long)(row - 1) * p.Cexp + ch * CC + j * 4;
                cp_async16(sC + (uint32_t)ch * p.smem_C + (uint32_t)(row * CC + j * 4) * 4, src, true);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&bar_const);
        }
        const uint4 ones = make_uint4(ones2<T>(), 0u, 0u, 0u), zero = make_uint4(0u, 0u, 0u, 0u);
        int k = 0;
        for (int item = blockIdx.x; item < pp.items; item += gridDim.x, ++k) {
            const Geo g = geom(item);
            const int abuf = k & 1;
            const uint32_t a0 = sA + abuf * pp.smem_A1;
            k1p_wait(&bar_a_empty[abuf], ((k >> 1) & 1) ^ 1, s_abort, p.tflag);      // the MMAs of item k-2 have finished reading this buffer
            const float inv_IWin = 1.0f / (float)g.IWin;
            const T* in_t = in + ((long long)g.n * p.Hin * p.Hin + (long long)(g.iy0 + g.ty_lo) * p.Hin + (g.ix0 + g.tx_lo)) * p.Cin;
            // one GEMM row (= one inside pixel, Cin contiguous values in global memory) per lane and step
            for (int r = lane; r < g.npix_in; r += 32) {
                const int ty = div_small(r, inv_IWin), tx = r - ty * g.IWin;
                const T* src = in_t + ((long long)ty * p.Hin + tx) * p.Cin;
                const uint32_t row = a0 + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
                const int sw = r & 7;
                for (int c = 0; c < kchunks; ++c)
                    cp_async16(row + (uint32_t)(c >> 3) * a_kb_stride + (uint32_t)(((c & 7) ^ sw) << 4), src + c * 8, true);
                sts128(row + (uint32_t)(kchunks >> 3) * a_kb_stride + (uint32_t)(((kchunks & 7) ^ sw) << 4), ones);
                if (p.cpr > kchunks + 1)
                    sts128(row + (uint32_t)((kchunks + 1) >> 3) * a_kb_stride + (uint32_t)((((kchunks + 1) & 7) ^ sw) << 4), zero);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&bar_a_full[abuf]);
        }
    } else if (warp == 1) {
        // =========================================================================== MMA issue
        k1p_wait(&bar_const, 0, s_abort, p.tflag);
        const int ksteps_total = p.cpr >> 1;
        int k = 0, gi = 0;
        for (int item = blockIdx.x; item < pp.items; item += gridDim.x, ++k) {
            const Geo g = geom(item);
            const int abuf = k & 1;
            k1p_wait(&bar_a_full[abuf], (k >> 1) & 1, s_abort, p.tflag);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int ch = 0; ch < n_chunks; ++ch, ++gi) {
                const int tbuf = gi & 1;
                k1p_wait(&bar_t_empty[tbuf], ((gi >> 1) & 1) ^ 1, s_abort, p.tflag);   // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0 && !*s_abort) {
                    for (int mt = 0; mt < g.mtc; ++mt)
                        for (int ks = 0; ks < ksteps_total; ++ks) {
                            const int kb = ks >> 2, kk = ks & 3;
                            const uint64_t ad = tc::make_desc(sA + abuf * pp.smem_A1 + (uint32_t)kb * a_kb_stride + (uint32_t)mt * BM * 128);
                            const uint64_t bd = tc::make_desc(sW + (uint32_t)ch * p.smem_W + (uint32_t)kb * CC * 128);
                            tc::umma_f16(tmem_base + (uint32_t)(tbuf * pp.tbuf_cols + mt * CC), ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2),
                                         p.idesc, ks ? 1u : 0u);
                        }
                    tc::umma_commit(&bar_t_full[tbuf]);
                    if (ch == n_chunks - 1) tc::umma_commit(&bar_a_empty[abuf]);     // all MMAs of the item done -> A buffer free
                }
                __syncwarp();
            }
        }
    } else if (warp < 2 + pp.epi_warps) {
        // =========================================================================== epilogue 1: TMEM -> swish -> E
        const int q4 = warp & 3;                       // TMEM lane quadrant this warp may read
        const int grp = (warp - 2) >> 2;               // 0 or 1: the warps of one quadrant share its (M tile, unit) pairs
        const int NG = pp.epi_warps >> 2;
        const int et = tid - K1P_EPI0, n_et = 32 * pp.epi_warps;
        const int units = CC >> 4;
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        const float inv_IW = 1.0f / (float)p.IW;
        int gi = 0;
        for (int item = blockIdx.x; item < pp.items; item += gridDim.x) {
            const Geo g = geom(item);
            const float inv_IWin = 1.0f / (float)g.IWin;
            uint32_t e_off[3];
            bool e_valid[3];
#pragma unroll
            for (int mt = 0; mt < 3; ++mt) {
                const int r = mt * BM + q4 * 32 + lane;
                e_valid[mt] = r < g.npix_in;
                const int rc = e_valid[mt] ? r : 0;
                const int ty = div_small(rc, inv_IWin), tx = rc - ty * g.IWin;
                e_off[mt] = (uint32_t)((g.ty_lo + ty) * p.IW + g.tx_lo + tx) * pitchE;
            }
            // pairs f = mt * units + u of this warp: f = grp, grp + NG, ... over the M tiles that have rows in this quadrant
            int mt_count = 0;
#pragma unroll
            for (int mt = 0; mt < 3; ++mt) mt_count += (mt * BM + q4 * 32 < g.npix_in) ? 1 : 0;      // the valid tiles are a prefix
            const int n_pairs = mt_count * units;
            const bool border = g.npix_in != npix;
            for (int ch = 0; ch < n_chunks; ++ch, ++gi) {
                const int buf = gi & 1;
                k1p_wait(&bar_t_full[buf], (gi >> 1) & 1, s_abort, p.tflag);
                k1p_wait(&bar_e_empty[buf], ((gi >> 1) & 1) ^ 1, s_abort, p.tflag);     // the depthwise of chunk gi-2 has finished with this E
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t e0 = sE + buf * pp.smem_E1;
                if (border && ch < 2) {
                    // halo pixels outside the image are zero rows; both E buffers carry rows of other tiles' shapes
                    for (int r = et; r < npix; r += n_et) {
                        const int ty = div_small(r, inv_IW), tx = r - ty * p.IW;
                        if (ty < g.ty_lo || ty >= g.ty_lo + g.IHin || tx < g.tx_lo || tx >= g.tx_lo + g.IWin)
                            for (int b = 0; b < pitchE; b += 16) sts128(e0 + (uint32_t)r * pitchE + b, zero);
                    }
                }
                if (!*s_abort) {
                    const uint32_t t0 = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * pp.tbuf_cols);
                    auto col_of = [&](int f) { const int mt = f >= 2 * units ? 2 : (f >= units ? 1 : 0); return (uint32_t)(mt * CC + (f - mt * units) * 16); };
                    auto process = [&](uint32_t (&r)[16], int f) {
                        const int mt = f >= 2 * units ? 2 : (f >= units ? 1 : 0), u = f - mt * units;
                        const bool valid = mt == 0 ? e_valid[0] : (mt == 1 ? e_valid[1] : e_valid[2]);
                        if (valid) {
                            float v[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = swish_from_half(__uint_as_float(r[j]));
                            const uint4 lo = make_uint4(pack2<T>(v[0], v[1]), pack2<T>(v[2], v[3]), pack2<T>(v[4], v[5]), pack2<T>(v[6], v[7]));
                            const uint4 hi = make_uint4(pack2<T>(v[8], v[9]), pack2<T>(v[10], v[11]), pack2<T>(v[12], v[13]), pack2<T>(v[14], v[15]));
                            const uint32_t dst = e0 + (mt == 0 ? e_off[0] : (mt == 1 ? e_off[1] : e_off[2])) + u * 32;
                            sts128(dst, lo);
                            sts128(dst + 16, hi);
                        }
                    };
                    uint32_t ra[16], rb[16];
                    int f = grp;
                    if (f < n_pairs) tmem_ld16_issue(t0 + col_of(f), ra);
                    while (f < n_pairs) {
                        tmem_ld16_wait(ra);
                        const int f2 = f + NG;
                        if (f2 < n_pairs) tmem_ld16_issue(t0 + col_of(f2), rb);     // flies while ra is processed
                        process(ra, f);
                        if (f2 >= n_pairs) break;
                        tmem_ld16_wait(rb);
                        const int f3 = f2 + NG;
                        if (f3 < n_pairs) tmem_ld16_issue(t0 + col_of(f3), ra);
                        process(rb, f2);
                        f = f3;
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(&bar_t_empty[buf]);
                mbar_arrive(&bar_e_full[buf]);
            }
        }
    } else {
        // =========================================================================== depthwise on E
        const int dtid = tid - (K1P_EPI0 + 32 * pp.epi_warps);
        const int CVc = CC >> 2;
        const int py = div_small(dtid, 1.0f / (float)CVc), cv = dtid - py * CVc;
        const bool dw_active = py < pp.PYp;
        const int nstrips = p.TH << p.spr_log2;
        const uint32_t e_rowstride = (uint32_t)p.IW * pitchE;
        constexpr int NCOL = (R - 1) * S + KS;
        k1p_wait(&bar_const, 0, s_abort, p.tflag);
        int gi = 0;
        for (int item = blockIdx.x; item < pp.items; item += gridDim.x) {
            const Geo g = geom(item);
            T* const out_n = out + (long long)g.n * p.Ho * p.Ho * p.Cexp;
            for (int ch = 0; ch < n_chunks; ++ch, ++gi) {
                const int buf = gi & 1;
                const int cbase = ch * CC;
                k1p_wait(&bar_e_full[buf], (gi >> 1) & 1, s_abort, p.tflag);
                float sum[4] = {0.f, 0.f, 0.f, 0.f};
                if (dw_active) {
                    const int c0 = cbase + cv * 4;
                    const uint32_t cst = sC + (uint32_t)ch * p.smem_C + (uint32_t)cv * 16;
                    const float4 bq = lds_f4(cst);
                    const uint32_t e_cv = sE + buf * pp.smem_E1 + (uint32_t)cv * 8;
                    for (int sidx = py; sidx < nstrips; sidx += pp.PYp) {
                        const int oyl = sidx >> p.spr_log2, oxl0 = (sidx - (oyl << p.spr_log2)) * R;
                        float2 acc[R][2];      // (ch0,ch1), (ch2,ch3): one FFMA2 (fma.rn.f32x2) per pair - same IEEE FMAs, half the issue slots
#pragma unroll
                        for (int r = 0; r < R; ++r) { acc[r][0] = make_float2(bq.x, bq.y); acc[r][1] = make_float2(bq.z, bq.w); }
                        uint32_t erow = e_cv + (uint32_t)(oyl * S) * e_rowstride + (uint32_t)(oxl0 * S) * pitchE;
#pragma unroll
                        for (int ky = 0; ky < KS; ++ky) {
                            float2 wr[KS][2];
#pragma unroll
                            for (int kx = 0; kx < KS; ++kx) {
                                const float4 wq = lds_f4(cst + (uint32_t)((1 + ky * KS + kx) * CC) * 4);
                                wr[kx][0] = make_float2(wq.x, wq.y); wr[kx][1] = make_float2(wq.z, wq.w);
                            }
                            uint32_t ea = erow;
#pragma unroll
                            for (int col = 0; col < NCOL; ++col) {
                                uint32_t a, b;
                                lds64(ea, a, b);
                                ea += pitchE;
                                float2 x01, x23;
                                unpack2<T>(a, x01.x, x01.y);
                                unpack2<T>(b, x23.x, x23.y);
#pragma unroll
                                for (int r = 0; r < R; ++r) {
                                    const int kx = col - r * S;          // compile-time after unrolling
                                    if (kx >= 0 && kx < KS) {
                                        ffma2(acc[r][0], x01, wr[kx][0]);
                                        ffma2(acc[r][1], x23, wr[kx][1]);
                                    }
                                }
                            }
                            erow += e_rowstride;
                        }
                        const int oy = g.ty0 + oyl;
                        T* dst = out_n + ((long long)oy * p.Ho + g.tx0 + oxl0) * p.Cexp + c0;
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            if (oxl0 + r < p.TW && oy < p.Ho && g.tx0 + oxl0 + r < p.Ho) {
#pragma unroll
                                for (int i = 0; i < 2; ++i) {
                                    acc[r][i].x = swish_from_half(acc[r][i].x); acc[r][i].y = swish_from_half(acc[r][i].y);
                                    sum[2 * i] += acc[r][i].x; sum[2 * i + 1] += acc[r][i].y;
                                }
                                uint2 o;
                                o.x = pack2<T>(acc[r][0].x, acc[r][0].y);
                                o.y = pack2<T>(acc[r][1].x, acc[r][1].y);
                                *reinterpret_cast<uint2*>(dst + (long long)r * p.Cexp) = o;
                            }
                        }
                    }
                }
                mbar_arrive(&bar_e_empty[buf]);                   // every depthwise thread: its reads of this E are done
                const uint32_t r_buf = sR + (uint32_t)(buf * pp.PYp * CC) * 4;
                if (dw_active)
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(r_buf + (uint32_t)(py * CC + cv * 4) * 4),
                                 "f"(sum[0]), "f"(sum[1]), "f"(sum[2]), "f"(sum[3]) : "memory");
                asm volatile("bar.sync 1, %0;" ::"r"(pp.ndw) : "memory");
                if (dtid < CC) {
                    float s4[4] = {0.f, 0.f, 0.f, 0.f};
                    int y = 0;
                    for (; y + 3 < pp.PYp; y += 4) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float t;
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(r_buf + (uint32_t)((y + i) * CC + dtid) * 4));
                            s4[i] += t;
                        }
                    }
                    for (; y < pp.PYp; ++y) {
                        float t;
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(r_buf + (uint32_t)(y * CC + dtid) * 4));
                        s4[y & 3] += t;
                    }
                    p.partial[((long long)g.n * pp.tiles + g.tile) * p.Cexp + cbase + dtid] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
                }
                // the squeeze scratch of this parity is rewritten two chunks later, after the next named barrier
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)pp.tmem_cols) : "memory");
}

// Plan: K1's geometry for (TH, TW, R, CC) + the K1P shared-memory layout.  Only blocks with several tiles per crop.
inline bool plan_k1p(int Hin, int Ho, int Cin, int Cexp, int k, int s, int pad, bool is_bf16, int TH, int TW, int R, int CC, int epi_warps,
                     K1PParams* pp, size_t* smem_out) {
    if (epi_warps != 4 && epi_warps != 8) return false;
    pp->epi_warps = epi_warps;
    pp->ndw = K1P_THREADS - K1P_EPI0 - 32 * epi_warps;
    size_t dummy = 0;
    if (!plan_k1_candidate(Hin, Ho, Cin, Cexp, k, s, pad, is_bf16, TH, TW, R, CC, 512, 1, &pp->k, &dummy)) return false;
    K1Params& p = pp->k;
    pp->tiles = p.tiles_x * p.tiles_y;
    if (pp->tiles < 2) return false;
    pp->PYp = pp->ndw / (CC / 4);
    if (pp->PYp < 1 || CC > pp->ndw) return false;
    pp->tbuf_cols = p.mtiles * CC;
    int cols = 32;
    while (cols < 2 * pp->tbuf_cols) cols <<= 1;
    if (cols > 512) return false;
    pp->tmem_cols = cols;
    pp->smem_A1 = p.smem_A;
    pp->smem_Wall = p.n_chunks * p.smem_W;
    pp->smem_Call = p.n_chunks * p.smem_C;
    pp->smem_E1 = p.e_rows * p.pitchE;                  // multiple of 16
    const size_t total = (size_t)2 * pp->smem_A1 + pp->smem_Wall + pp->smem_Call + (size_t)2 * pp->smem_E1 + (size_t)2 * pp->PYp * CC * 4 + 1024;
    // the last M tile's UMMA may read past the staged rows of the second A buffer: that must stay inside the window
    if ((size_t)2 * pp->smem_A1 + (size_t)p.mtiles * BM * 128 + 1024 > total) return false;
    *smem_out = total;
    return total <= K1_MAX_SMEM;
}

template <typename T>
int launch_k1p(cudaStream_t stream, K1PParams pp, int k, int s, int R, size_t smem, int n_crops, int sm_count) {
    pp.k.N = n_crops;
    pp.items = n_crops * pp.tiles;
    const int ctas = pp.items < sm_count ? pp.items : sm_count;
    if (ctas < 1) return 0;
#define K1P(KS, S, RR)                                                                                                     \
    do {                                                                                                                   \
        auto kfn = k1p_kernel<T, KS, S, RR>;                                                                               \
        if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1_MAX_SMEM) != cudaSuccess) return -1;  \
        kfn<<<ctas, K1P_THREADS, smem, stream>>>(pp);                                                                      \
        return 0;                                                                                                          \
    } while (0)
    if (k == 3 && s == 2 && R == 4) K1P(3, 2, 4);
    if (k == 3 && s == 1 && R == 7) K1P(3, 1, 7);
    if (k == 5 && s == 1 && R == 7) K1P(5, 1, 7);
    if (k == 5 && s == 2 && R == 4) K1P(5, 2, 4);
    if (k == 3 && s == 2 && R == 7) K1P(3, 2, 7);
    if (k == 3 && s == 1 && R == 4) K1P(3, 1, 4);
    if (k == 5 && s == 1 && R == 4) K1P(5, 1, 4);
    if (k == 5 && s == 2 && R == 7) K1P(5, 2, 7);
#undef K1P
    return 1;
}

}  // namespace fused
}  // namespace whenet
