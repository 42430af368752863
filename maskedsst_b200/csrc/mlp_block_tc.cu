// Kernel (2) fused, bf16 mode: the whole FeedForward block of a pre-norm layer in ONE kernel -- the hidden activation never leaves
// the SM on its way from the first to the second linear:
//     u = h2 W1^T + b1            (saved, bf16: GELU' argument of the backward)
//     g = dropout(gelu_erf(u))    (saved, bf16: operand of the W2 weight gradient)          -> smem tile -> A operand of the 2nd GEMM
//     y = xmid + dropout(g W2^T + b2)                       (fp32 residual stream)
//     h1' = LayerNorm(y) of the NEXT layer's attention pre-norm (bf16) + its (mean, rstd)     (optional)
// Reference: FeedForward.forward, src/vit_spatial_spectral.py:35-44 (Linear -> GELU -> Dropout -> Linear -> Dropout), the residual
// add of Transformer.forward :103 and PreNorm :25-29 of the following layer.  Replaces gemm_tn_kernel<3> + gemm_tn_kernel<6>
// (two launches, h2 / g / xmid re-read between them) with identical arithmetic conventions: fp32 accumulate, u rounded to bf16
// for storage only, Abramowitz-Stegun erf, dropout quads indexed by (row * N + col) / 4, two-pass LayerNorm variance.
//
// One persistent CTA per SM, tiles of 128 token rows.  Warp 16 = TMA producer (W1 / W2 once; per tile the h2 rows as SWIZZLE_64B
// chunks and the xmid rows as [128][32] fp32 chunks, both double-buffered), warp 17 = MMA issuer (tcgen05: 128 x 64 x D, then
// 128 x D x 64 with the g tile the compute warps just wrote), warp 18 = TMA stores (u, g; then y and h1', which are staged IN PLACE in
// the tile's xmid / h2 buffers), warps 0-15 = 512 epilogue threads, thread = (row, column quarter).  HBM-bound: per token
// D*2 + D*4 bytes in, 2*M*2 + D*4 + D*2 out (1.4 KB at D = 96, M = 64).  The backward of the block (mlp_block_bwd_kernel) is below.
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"
#include <stdlib.h>
#include <stdio.h>

namespace msst {
using namespace ptx;
typedef __nv_bfloat16 bf16;
int make_tmap_nd(CUtensorMap* m, const void* base, int elem_bytes, int rank, const int64_t* dims, const int64_t* strides, const int* box, int swizzle_bytes);   // gemm_bf16.cu

namespace {

constexpr int MB_THREADS = 608;
constexpr uint32_t MB_T16 = 16384;
constexpr uint32_t COL_A2 = 0, COL_A3 = 64;

struct MlpParams {
    int64_t R, n_tiles;
    int D, nch;
    const float *b1, *b2, *ln_w, *ln_b;
    float* ln_stats;
    int has_ln;
    Drop drop_h, drop_o;
};
struct alignas(8) MlpBars {
    uint64_t w_full, in_h[2], in_x[2], buf_free[2], acc2_full, g_ready, ug_free, acc3_full, out_ready;
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t pack_bf(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}
constexpr uint64_t kHiK128 = ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (kLayoutSW128 << 61);
__device__ __forceinline__ uint64_t kdesc(uint32_t addr) { return kHiK128 | (1ull << 16) | (uint64_t)((addr >> 4) & 0x3FFF); }

template <int NCH>
__global__ void __launch_bounds__(MB_THREADS, 1)
mlp_block_fwd_kernel(const __grid_constant__ CUtensorMap tma_h2, const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_w1,
                     const __grid_constant__ CUtensorMap tma_w2, const __grid_constant__ CUtensorMap tma_u, const __grid_constant__ CUtensorMap tma_g,
                     const __grid_constant__ CUtensorMap tma_y, const __grid_constant__ CUtensorMap tma_h1, const MlpParams p) {
    constexpr int D = NCH * 32, M = 64;
    constexpr uint32_t HB = NCH * 8192, XB = NCH * MB_T16;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (smem_u32(smem) & 1023u) __trap();
    uint8_t* w1_s = smem;                        // NCH chunks [64 rows][32 cols] SWIZZLE_64B
    uint8_t* w2_s = w1_s + NCH * 4096;           // [D rows][64 cols] SWIZZLE_128B
    uint8_t* h_s = w2_s + D * 128;               // [2][NCH chunks [128][32] bf16 SWIZZLE_64B]: h2 in, h1' out (in place)
    uint8_t* x_s = h_s + 2 * HB;                 // [2][NCH chunks [128][32] fp32 SWIZZLE_128B]: xmid in, y out (in place)
    uint8_t* g_s = x_s + 2 * XB;                 // [128][64] bf16 SWIZZLE_128B
    uint8_t* u_s = g_s + MB_T16;
    float* xch = reinterpret_cast<float*>(u_s + MB_T16);   // [2][4][128]
    MlpBars* bars = reinterpret_cast<MlpBars*>(xch + 2 * 4 * 128);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 17 && elect_one()) {
        mbar_init(&bars->w_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->in_h[i], 1); mbar_init(&bars->in_x[i], 1); mbar_init(&bars->buf_free[i], 1); }
        mbar_init(&bars->acc2_full, 1); mbar_init(&bars->g_ready, 16); mbar_init(&bars->ug_free, 1); mbar_init(&bars->acc3_full, 1);
        mbar_init(&bars->out_ready, 16);
        fence_barrier_init();
    }
    if (warp == 16) {
        tmem_alloc(&bars->tmem_base, 256);
        if (elect_one()) {
            prefetch_tmap(&tma_h2); prefetch_tmap(&tma_x); prefetch_tmap(&tma_w1); prefetch_tmap(&tma_w2);
            prefetch_tmap(&tma_u); prefetch_tmap(&tma_g); prefetch_tmap(&tma_y); prefetch_tmap(&tma_h1);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    const int64_t my_tiles = blockIdx.x < p.n_tiles ? (p.n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    if (warp == 16) {
        // ===== TMA producer =====
        if (elect_one()) {
            mbar_arrive_expect_tx(&bars->w_full, (uint32_t)(NCH * 4096 + D * 128));
#pragma unroll
            for (int c = 0; c < NCH; ++c) tma_load_2d(w1_s + c * 4096, &tma_w1, &bars->w_full, c * 32, 0);
            tma_load_2d(w2_s, &tma_w2, &bars->w_full, 0, 0);
            for (int64_t t = 0; t < my_tiles; ++t) {
                const int b = (int)(t & 1);
                const int row0 = (int)((blockIdx.x + t * gridDim.x) * 128);
                mbar_wait(&bars->buf_free[b], ((uint32_t)(t >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&bars->in_h[b], HB);
#pragma unroll
                for (int c = 0; c < NCH; ++c) tma_load_2d(h_s + b * HB + c * 8192, &tma_h2, &bars->in_h[b], c * 32, row0);
                mbar_arrive_expect_tx(&bars->in_x[b], XB);
#pragma unroll
                for (int c = 0; c < NCH; ++c) tma_load_2d(x_s + b * XB + c * MB_T16, &tma_x, &bars->in_x[b], c * 32, row0);
            }
        }
    } else if (warp == 17) {
        // ===== MMA issuer =====
        if (elect_one() && my_tiles > 0) {
            const uint32_t idesc2 = make_idesc_bf16(128, M, 0, 0), idesc3 = make_idesc_bf16(128, D, 0, 0);
            const uint64_t w1d = make_smem_desc_sw64(smem_u32(w1_s)), w2d = kdesc(smem_u32(w2_s)), gd = kdesc(smem_u32(g_s));
            mbar_wait(&bars->w_full, 0);
            for (int64_t t = 0; t < my_tiles; ++t) {
                const int b = (int)(t & 1);
                mbar_wait(&bars->in_h[b], (uint32_t)(t >> 1) & 1);
                tc_fence_after();
                const uint64_t hd = make_smem_desc_sw64(smem_u32(h_s + b * HB));
#pragma unroll
                for (int ks = 0; ks < 2 * NCH; ++ks)                // u-accumulator = h2 W1^T  (the previous tile's was read before its g_ready)
                    umma_bf16(tmem + COL_A2, hd + (uint64_t)(((ks >> 1) * 8192 + (ks & 1) * 32) >> 4), w1d + (uint64_t)(((ks >> 1) * 4096 + (ks & 1) * 32) >> 4),
                              idesc2, ks != 0);
                umma_commit(&bars->acc2_full);
                mbar_wait(&bars->g_ready, (uint32_t)t & 1);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)                      // y-accumulator = g W2^T
                    umma_bf16(tmem + COL_A3, gd + (uint64_t)(ks * 2), w2d + (uint64_t)(ks * 2), idesc3, ks != 0);
                umma_commit(&bars->acc3_full);
            }
        }
    } else if (warp == 18) {
        // ===== TMA stores =====
        if (elect_one()) {
            for (int64_t t = 0; t < my_tiles; ++t) {
                const int b = (int)(t & 1);
                const int row0 = (int)((blockIdx.x + t * gridDim.x) * 128);
                mbar_wait(&bars->g_ready, (uint32_t)t & 1);
                tma_store_2d(&tma_u, u_s, 0, row0);
                tma_store_2d(&tma_g, g_s, 0, row0);
                tma_store_commit();
                tma_store_wait_read();
                mbar_arrive(&bars->ug_free);
                mbar_wait(&bars->out_ready, (uint32_t)t & 1);
#pragma unroll
                for (int c = 0; c < NCH; ++c) tma_store_2d(&tma_y, x_s + b * XB + c * MB_T16, c * 32, row0);
                if (p.has_ln) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) tma_store_2d(&tma_h1, h_s + b * HB + c * 8192, c * 32, row0);
                }
                tma_store_commit();
                tma_store_wait_read();
                mbar_arrive(&bars->buf_free[b]);
            }
        }
    } else {
        // ===== 512 epilogue threads: thread = (row L of the tile, column quarter cq) =====
        const int lq = warp & 3, cq = warp >> 2;
        const int L = lq * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;
        const uint32_t sw128 = (uint32_t)(L & 7), sw64 = (uint32_t)((L >> 1) & 3);
        float* xs0 = xch; float* xs1 = xch + 4 * 128;
        for (int64_t t = 0; t < my_tiles; ++t) {
            const int b = (int)(t & 1);
            const uint32_t ph = (uint32_t)t & 1;
            const int64_t row = (blockIdx.x + t * gridDim.x) * 128 + L;
            // ---- u = acc + b1 ; g = dropout(gelu(u)) : 16 columns per thread ----
            mbar_wait(&bars->acc2_full, ph);
            tc_fence_after();
            {
                uint32_t a[16];
                tmem_ld_32x16(tmem + lane_addr + COL_A2 + 16 * cq, a);
                tmem_ld_wait();
                uint32_t up[8], gp[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int col = 16 * cq + 4 * q;
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b1 + col));
                    float f[4] = {__uint_as_float(a[4 * q]) + bb.x, __uint_as_float(a[4 * q + 1]) + bb.y, __uint_as_float(a[4 * q + 2]) + bb.z,
                                  __uint_as_float(a[4 * q + 3]) + bb.w};
                    up[2 * q] = pack_bf(f[0], f[1]); up[2 * q + 1] = pack_bf(f[2], f[3]);
#pragma unroll
                    for (int e = 0; e < 4; ++e) f[e] = gelu_fast(f[e]);
                    if (p.drop_h.on()) {
                        float d4[4];
                        drop_factor4(p.drop_h, (uint64_t)(row * M + col) >> 2, d4);
#pragma unroll
                        for (int e = 0; e < 4; ++e) f[e] *= d4[e];
                    }
                    gp[2 * q] = pack_bf(f[0], f[1]); gp[2 * q + 1] = pack_bf(f[2], f[3]);
                }
                if (t > 0) mbar_wait(&bars->ug_free, (uint32_t)(t - 1) & 1);     // the previous tile's u / g stores have read the tiles
                uint8_t* ur = u_s + L * 128; uint8_t* gr = g_s + L * 128;
                const uint32_t c0 = (((uint32_t)(2 * cq)) ^ sw128) << 4, c1 = (((uint32_t)(2 * cq + 1)) ^ sw128) << 4;
                *reinterpret_cast<uint4*>(ur + c0) = make_uint4(up[0], up[1], up[2], up[3]);
                *reinterpret_cast<uint4*>(ur + c1) = make_uint4(up[4], up[5], up[6], up[7]);
                *reinterpret_cast<uint4*>(gr + c0) = make_uint4(gp[0], gp[1], gp[2], gp[3]);
                *reinterpret_cast<uint4*>(gr + c1) = make_uint4(gp[4], gp[5], gp[6], gp[7]);
            }
            tc_fence_before();
            fence_proxy_async();
            warp_arrive(&bars->g_ready, lane);
            // ---- y = xmid + dropout(acc + b2) ; LayerNorm of the next pre-norm : columns 32c + 8cq .. +7 of every chunk c ----
            mbar_wait(&bars->acc3_full, ph);
            mbar_wait(&bars->in_x[b], (uint32_t)(t >> 1) & 1);
            tc_fence_after();
            float y[NCH][8];
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                uint32_t a[8];
                tmem_ld_32x8(tmem + lane_addr + COL_A3 + 32 * c + 8 * cq, a);
                tmem_ld_wait();
                const int col = 32 * c + 8 * cq;
                uint8_t* xr = x_s + b * XB + c * MB_T16 + L * 128;
                const uint32_t c0 = (((uint32_t)(2 * cq)) ^ sw128) << 4, c1 = (((uint32_t)(2 * cq + 1)) ^ sw128) << 4;
                const float4 r0 = *reinterpret_cast<const float4*>(xr + c0), r1 = *reinterpret_cast<const float4*>(xr + c1);
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.b2 + col)), b1v = __ldg(reinterpret_cast<const float4*>(p.b2 + col + 4));
                float f[8] = {__uint_as_float(a[0]) + b0.x, __uint_as_float(a[1]) + b0.y, __uint_as_float(a[2]) + b0.z, __uint_as_float(a[3]) + b0.w,
                              __uint_as_float(a[4]) + b1v.x, __uint_as_float(a[5]) + b1v.y, __uint_as_float(a[6]) + b1v.z, __uint_as_float(a[7]) + b1v.w};
                if (p.drop_o.on()) {
                    float d4[4];
                    drop_factor4(p.drop_o, (uint64_t)(row * D + col) >> 2, d4);
                    f[0] *= d4[0]; f[1] *= d4[1]; f[2] *= d4[2]; f[3] *= d4[3];
                    drop_factor4(p.drop_o, (uint64_t)(row * D + col + 4) >> 2, d4);
                    f[4] *= d4[0]; f[5] *= d4[1]; f[6] *= d4[2]; f[7] *= d4[3];
                }
                y[c][0] = f[0] + r0.x; y[c][1] = f[1] + r0.y; y[c][2] = f[2] + r0.z; y[c][3] = f[3] + r0.w;
                y[c][4] = f[4] + r1.x; y[c][5] = f[5] + r1.y; y[c][6] = f[6] + r1.z; y[c][7] = f[7] + r1.w;
                *reinterpret_cast<float4*>(xr + c0) = make_float4(y[c][0], y[c][1], y[c][2], y[c][3]);      // staged in place for the TMA store
                *reinterpret_cast<float4*>(xr + c1) = make_float4(y[c][4], y[c][5], y[c][6], y[c][7]);
#pragma unroll
                for (int e = 0; e < 8; ++e) s += y[c][e];
            }
            tc_fence_before();
            if (p.has_ln) {
                xs0[cq * 128 + L] = s;
                asm volatile("bar.sync 1, 512;" ::: "memory");
                const float mean = ((xs0[L] + xs0[128 + L]) + (xs0[256 + L] + xs0[384 + L])) / D;
                float sq = 0.f;
#pragma unroll
                for (int c = 0; c < NCH; ++c)
#pragma unroll
                    for (int e = 0; e < 8; ++e) { y[c][e] -= mean; sq = fmaf(y[c][e], y[c][e], sq); }
                xs1[cq * 128 + L] = sq;
                asm volatile("bar.sync 1, 512;" ::: "memory");
                const float rstd = rsqrtf(((xs1[L] + xs1[128 + L]) + (xs1[256 + L] + xs1[384 + L])) / D + 1e-5f);
                if (cq == 0 && row < p.R) { p.ln_stats[2 * row] = mean; p.ln_stats[2 * row + 1] = rstd; }
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const int col = 32 * c + 8 * cq;
                    const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.ln_w + col)), w1v = __ldg(reinterpret_cast<const float4*>(p.ln_w + col + 4));
                    const float4 l0 = __ldg(reinterpret_cast<const float4*>(p.ln_b + col)), l1 = __ldg(reinterpret_cast<const float4*>(p.ln_b + col + 4));
                    // h1' chunk c: [128 rows][64 B] SWIZZLE_64B, this thread's 16 bytes at logical chunk cq
                    *reinterpret_cast<uint4*>(h_s + b * HB + c * 8192 + L * 64 + ((((uint32_t)cq) ^ sw64) << 4)) =
                        make_uint4(pack_bf(y[c][0] * rstd * w0.x + l0.x, y[c][1] * rstd * w0.y + l0.y), pack_bf(y[c][2] * rstd * w0.z + l0.z, y[c][3] * rstd * w0.w + l0.w),
                                   pack_bf(y[c][4] * rstd * w1v.x + l1.x, y[c][5] * rstd * w1v.y + l1.y), pack_bf(y[c][6] * rstd * w1v.z + l1.z, y[c][7] * rstd * w1v.w + l1.w));
                }
            }
            fence_proxy_async();
            warp_arrive(&bars->out_ready, lane);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}


// =========================================================================================================
// backward of the FeedForward block in ONE kernel (replaces gemm_wgrad(W2) + gemm_tn<4> + gemm_wgrad(W1) + gemm_tn<1>):
//     du  = (dyb W2) * gelu'(u) * hidden-dropout      (bf16; never leaves the SM)         db1 += colsum(du)
//     dW2 += dyb^T g        dW1 += du^T h2            (accumulated in TMEM over ALL tiles of the CTA, one atomic flush at the end)
//     dh  = du W1           (fp32 [R, D]: dy of the LayerNorm-2 backward)
// dyb = bf16(dropout_mlp_out(dy)) [R, D] is produced upstream (LayerNorm backward / cast) with db2.
// Per tile of 128 token rows, one shared-memory stage = dyb | g | h2 | u, every operand a [128 tokens][64] SWIZZLE_128B chunk:
// dyb is the K-major A operand of the du contraction AND the MN-major B operand of dW2 (reduction over the tokens); g / du / h2
// are MN-major operands of the weight gradients (M = 64 atoms, rows 16j + i -> TMEM lane 32j + i (+16 for dW1): both weight
// gradients share 96 TMEM columns); du overwrites u in place and is the K-major A operand of dh = du W1.  dh is staged in the
// (dead) dyb / g chunks of the stage and leaves by TMA store.  HBM per token: 2D + 2M + 2M + 2D in, 4D out (1 KB at D = 96).
// =========================================================================================================
constexpr uint32_t COL_DU = 0, COL_DHB = 64, COL_W = 192;
struct MlpBwdParams {
    int64_t R, n_tiles;
    float *dw1, *dw2, *db1;
    Drop drop_h;
    int dbg;                   // MSST_MLPB_DBG (profiling only): 1 = no GELU' math, 2 = no weight-gradient MMAs, 4 = no dh stores, 8 = timeline of CTA 0
    long long* tl;             // [tile][8] clock64 stamps
};
#define MB_TL(t, slot) do { if (p.tl && blockIdx.x == 0 && (t) < 32) p.tl[(t) * 8 + (slot)] = clock64(); } while (0)
struct alignas(8) MlpBwdBars {
    uint64_t w_full, in_full[2], stage_free[2], mma_done[2], acc1_full, du_ready, acc3_full, out_ready, all_done;
    uint32_t tmem_base;
};
// MN-major SWIZZLE_128B operand made of [tokens][64] chunks (16 KB apart): + 128 per K = 16 tokens
__device__ __forceinline__ uint64_t mn_desc(uint32_t addr) { return make_smem_desc(addr, MB_T16, 1024); }

template <int D>
__global__ void __launch_bounds__(MB_THREADS, 1)
mlp_block_bwd_kernel(const __grid_constant__ CUtensorMap tma_dy, const __grid_constant__ CUtensorMap tma_g, const __grid_constant__ CUtensorMap tma_h2,
                     const __grid_constant__ CUtensorMap tma_u, const __grid_constant__ CUtensorMap tma_w2t, const __grid_constant__ CUtensorMap tma_w1t,
                     const __grid_constant__ CUtensorMap tma_dh, const MlpBwdParams p) {
    constexpr int M = 64, NC = (D + 63) / 64, NR = D / 32;          // 64-column chunks of a [*, D] bf16 tile; 32-column rounds of the fp32 dh tile
    constexpr uint32_t O_G = NC * MB_T16, O_H2 = O_G + MB_T16, O_U = O_H2 + NC * MB_T16, STAGE = O_U + MB_T16;
    static_assert(NR * MB_T16 <= O_H2, "dh staging must fit the dead dyb | g chunks");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (smem_u32(smem) & 1023u) __trap();
    uint8_t* stage_s = smem;                             // [2][dyb NC | g | h2 NC | u/du]
    uint8_t* w2t_s = stage_s + 2 * STAGE;                // NC chunks [64 rows][64 cols]   (W2^T [M, D], K-major B of du)
    uint8_t* w1t_s = w2t_s + NC * 8192;                  // [D rows][64 cols]              (W1^T [D, M], K-major B of dh)
    float* red = reinterpret_cast<float*>(w1t_s + D * 128);   // [4 lane quarters][64] column sums
    MlpBwdBars* bars = reinterpret_cast<MlpBwdBars*>(red + 4 * 64);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 17 && elect_one()) {
        mbar_init(&bars->w_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->in_full[i], 1); mbar_init(&bars->stage_free[i], 1); mbar_init(&bars->mma_done[i], 1); }
        mbar_init(&bars->acc1_full, 1); mbar_init(&bars->du_ready, 16); mbar_init(&bars->acc3_full, 1); mbar_init(&bars->out_ready, 16);
        mbar_init(&bars->all_done, 1);
        fence_barrier_init();
    }
    if (warp == 16) {
        tmem_alloc(&bars->tmem_base, 512);
        if (elect_one()) {
            prefetch_tmap(&tma_dy); prefetch_tmap(&tma_g); prefetch_tmap(&tma_h2); prefetch_tmap(&tma_u);
            prefetch_tmap(&tma_w2t); prefetch_tmap(&tma_w1t); prefetch_tmap(&tma_dh);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    const int64_t my_tiles = blockIdx.x < p.n_tiles ? (p.n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    if (warp == 16) {
        // ===== TMA producer =====
        if (elect_one() && my_tiles > 0) {
            mbar_arrive_expect_tx(&bars->w_full, (uint32_t)(NC * 8192 + D * 128));
#pragma unroll
            for (int c = 0; c < NC; ++c) tma_load_2d(w2t_s + c * 8192, &tma_w2t, &bars->w_full, c * 64, 0);
            tma_load_2d(w1t_s, &tma_w1t, &bars->w_full, 0, 0);
            // two smem stages hold one tile in flight while the other is worked on; the tile after those is pulled into L2 already
            // (loaded HBM latency ~2.5 us >> the ~1.5 us a tile may take), so the stage's own load is an L2 hit
            auto prefetch_tile = [&](int row0) {
#pragma unroll
                for (int c = 0; c < NC; ++c) { tma_prefetch_l2_2d(&tma_dy, c * 64, row0); tma_prefetch_l2_2d(&tma_h2, c * 64, row0); }
                tma_prefetch_l2_2d(&tma_u, 0, row0); tma_prefetch_l2_2d(&tma_g, 0, row0);
            };
            for (int64_t t = 0; t < my_tiles; ++t) {
                const int s = (int)(t & 1);
                const uint32_t ph = ((uint32_t)(t >> 1) & 1) ^ 1;
                const int row0 = (int)((blockIdx.x + t * gridDim.x) * 128);
                uint8_t* st = stage_s + s * STAGE;
                mbar_wait(&bars->stage_free[s], ph);     // the dh stores of the stage's previous tile have read the staging chunks
                mbar_wait(&bars->mma_done[s], ph);       // ... and its contractions have read h2 / du
                MB_TL(t, 0);
                if (t + 2 < my_tiles && !(p.dbg & 16)) prefetch_tile((int)((blockIdx.x + (t + 2) * gridDim.x) * 128));
                mbar_arrive_expect_tx(&bars->in_full[s], STAGE);
#pragma unroll
                for (int c = 0; c < NC; ++c) tma_load_2d(st + c * MB_T16, &tma_dy, &bars->in_full[s], c * 64, row0);
                tma_load_2d(st + O_U, &tma_u, &bars->in_full[s], 0, row0);
                tma_load_2d(st + O_G, &tma_g, &bars->in_full[s], 0, row0);
#pragma unroll
                for (int c = 0; c < NC; ++c) tma_load_2d(st + O_H2 + c * MB_T16, &tma_h2, &bars->in_full[s], c * 64, row0);
            }
        }
    } else if (warp == 17) {
        // ===== MMA issuer =====
        if (elect_one() && my_tiles > 0) {
            const uint32_t idesc1 = make_idesc_bf16(128, M, 0, 0), idesc3 = make_idesc_bf16(128, D, 0, 0), idesc_w = make_idesc_bf16(64, D, 1, 1);
            const uint32_t w2a = smem_u32(w2t_s);
            const uint64_t w1d = kdesc(smem_u32(w1t_s));
            mbar_wait(&bars->w_full, 0);
            for (int64_t t = 0; t < my_tiles; ++t) {
                const int s = (int)(t & 1);
                const uint32_t st = smem_u32(stage_s + s * STAGE);
                mbar_wait(&bars->in_full[s], (uint32_t)(t >> 1) & 1);
                MB_TL(t, 1);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < D / 16; ++ks)                 // du accumulator = dyb W2   (K = D)
                    umma_bf16(tmem + COL_DU, kdesc(st + (ks >> 2) * MB_T16) + (uint64_t)((ks & 3) * 2), kdesc(w2a + (ks >> 2) * 8192) + (uint64_t)((ks & 3) * 2),
                              idesc1, ks != 0);
                umma_commit(&bars->acc1_full);
                {
                    const uint64_t ga = mn_desc(st + O_G), db = mn_desc(st);
#pragma unroll
                    for (int k = 0; k < 8; ++k)                     // dW2^T [64, D] += g^T dyb   (K = 128 tokens)
                        if (!(p.dbg & 2)) umma_bf16(tmem + COL_W, ga + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc_w, (t > 0) || k != 0);
                }
                mbar_wait(&bars->du_ready, (uint32_t)t & 1);
                tc_fence_after();
                {
                    const uint64_t dud = kdesc(st + O_U);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)                  // dh accumulator = du W1   (K = 64)
                        umma_bf16(tmem + COL_DHB, dud + (uint64_t)(ks * 2), w1d + (uint64_t)(ks * 2), idesc3, ks != 0);
                    umma_commit(&bars->acc3_full);
                    const uint64_t dua = mn_desc(st + O_U), hb = mn_desc(st + O_H2);
#pragma unroll
                    for (int k = 0; k < 8; ++k)                     // dW1 [64, D] += du^T h2
                        if (!(p.dbg & 2)) umma_bf16(tmem + COL_W + (16u << 16), dua + (uint64_t)(k * 128), hb + (uint64_t)(k * 128), idesc_w, (t > 0) || k != 0);
                }
                umma_commit(&bars->mma_done[s]);
                MB_TL(t, 7);
            }
            umma_commit(&bars->all_done);
        }
    } else if (warp == 18) {
        // ===== TMA stores of the staged dh rows =====
        if (elect_one()) {
            for (int64_t t = 0; t < my_tiles; ++t) {
                const int s = (int)(t & 1);
                const int row0 = (int)((blockIdx.x + t * gridDim.x) * 128);
                mbar_wait(&bars->out_ready, (uint32_t)t & 1);
#pragma unroll
                for (int c = 0; c < NR; ++c) if (!(p.dbg & 4)) tma_store_2d(&tma_dh, stage_s + s * STAGE + c * MB_T16, c * 32, row0);
                tma_store_commit();
                tma_store_wait_read();
                MB_TL(t, 6);
                mbar_arrive(&bars->stage_free[s]);
            }
        }
    } else {
        // ===== 512 epilogue threads: thread = (row L of the tile, column quarter cq) =====
        const int lq = warp & 3, cq = warp >> 2;
        const int L = lq * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;
        const uint32_t c0 = (((uint32_t)(2 * cq)) ^ (uint32_t)(L & 7)) << 4, c1 = (((uint32_t)(2 * cq + 1)) ^ (uint32_t)(L & 7)) << 4;
        float csum[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) csum[j] = 0.f;
        for (int64_t t = 0; t < my_tiles; ++t) {
            const int s = (int)(t & 1);
            const uint32_t ph = (uint32_t)t & 1;
            const int64_t row = (blockIdx.x + t * gridDim.x) * 128 + L;
            uint8_t* st = stage_s + s * STAGE;
            // ---- du = acc * gelu'(u) * dropout : 16 columns per thread, written over u ----
            mbar_wait(&bars->in_full[s], (uint32_t)(t >> 1) & 1);   // the u tile has landed (TMA writes -> this thread's reads)
            mbar_wait(&bars->acc1_full, ph);
            if (threadIdx.x == 0) MB_TL(t, 2);
            tc_fence_after();
            {
                uint32_t a[16];
                tmem_ld_32x16(tmem + lane_addr + COL_DU + 16 * cq, a);
                uint8_t* ur = st + O_U + L * 128;
                const uint4 u0 = *reinterpret_cast<const uint4*>(ur + c0), u1 = *reinterpret_cast<const uint4*>(ur + c1);
                const uint32_t uw[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
                tmem_ld_wait();
                uint32_t dp[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int col = 16 * cq + 4 * q;
                    const __nv_bfloat162 h0 = *reinterpret_cast<const __nv_bfloat162*>(&uw[2 * q]), h1 = *reinterpret_cast<const __nv_bfloat162*>(&uw[2 * q + 1]);
                    float f[4] = {__uint_as_float(a[4 * q]), __uint_as_float(a[4 * q + 1]), __uint_as_float(a[4 * q + 2]), __uint_as_float(a[4 * q + 3])};
                    if (!(p.dbg & 1)) {
                        f[0] *= gelu_fast_grad2(__low2float(h0)); f[1] *= gelu_fast_grad2(__high2float(h0));
                        f[2] *= gelu_fast_grad2(__low2float(h1)); f[3] *= gelu_fast_grad2(__high2float(h1));
                    }
                    if (p.drop_h.on()) {
                        float d4[4];
                        drop_factor4(p.drop_h, (uint64_t)(row * M + col) >> 2, d4);
#pragma unroll
                        for (int e = 0; e < 4; ++e) f[e] *= d4[e];
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) csum[4 * q + e] += f[e];
                    dp[2 * q] = pack_bf(f[0], f[1]); dp[2 * q + 1] = pack_bf(f[2], f[3]);
                }
                *reinterpret_cast<uint4*>(ur + c0) = make_uint4(dp[0], dp[1], dp[2], dp[3]);
                *reinterpret_cast<uint4*>(ur + c1) = make_uint4(dp[4], dp[5], dp[6], dp[7]);
            }
            tc_fence_before();
            fence_proxy_async();
            warp_arrive(&bars->du_ready, lane);
            if (threadIdx.x == 0) MB_TL(t, 3);
            // ---- dh rows: fp32 accumulator -> staging (the stage's dyb / g chunks: their readers completed before acc3_full) ----
            mbar_wait(&bars->acc3_full, ph);
            if (threadIdx.x == 0) MB_TL(t, 4);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < NR; ++c) {
                uint32_t v[8];
                tmem_ld_32x8(tmem + lane_addr + COL_DHB + 32 * c + 8 * cq, v);
                tmem_ld_wait();
                uint8_t* dr = st + c * MB_T16 + L * 128;
                *reinterpret_cast<uint4*>(dr + c0) = make_uint4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<uint4*>(dr + c1) = make_uint4(v[4], v[5], v[6], v[7]);
            }
            tc_fence_before();
            fence_proxy_async();
            warp_arrive(&bars->out_ready, lane);
            if (threadIdx.x == 0) MB_TL(t, 5);
        }
        if (my_tiles > 0) {
            // ---- db1: this thread's 16 column sums over its rows -> warp -> CTA -> one atomic per column ----
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float v = warp_sum(csum[j]);
                if (lane == 0) red[lq * 64 + 16 * cq + j] = v;
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (threadIdx.x < 64) atomicAdd(p.db1 + threadIdx.x, (red[threadIdx.x] + red[64 + threadIdx.x]) + (red[128 + threadIdx.x] + red[192 + threadIdx.x]));
            // ---- weight gradients: lane 32j + i holds row 16j + i of dW2^T (i < 16) or of dW1 (i >= 16); this thread's columns 8 (4c + cq) .. +7 ----
            mbar_wait(&bars->all_done, 0);
            tc_fence_after();
            const int m = 16 * lq + (lane & 15);
            const bool is_w1 = lane >= 16;
#pragma unroll
            for (int c = 0; c < NR; ++c) {
                uint32_t v[8];
                const int d0 = 8 * (4 * c + cq);
                tmem_ld_32x8(tmem + lane_addr + COL_W + d0, v);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    if (is_w1) atomicAdd(p.dw1 + m * D + d0 + e, __uint_as_float(v[e]));
                    else atomicAdd(p.dw2 + (d0 + e) * M + m, __uint_as_float(v[e]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

}  // namespace

bool mlp_block_supported(int D, int M) { return M == 64 && D % 32 == 0 && D >= 32 && D <= 96; }

// h2 [R,D] bf16, xmid [R,D] fp32, w1 [64,D] bf16, w2 [D,64] bf16 -> u, g [R,64] bf16, y [R,D] fp32, optionally h1 [R,D] bf16 + ln_stats [R,2]
int mlp_block_fwd(const bf16* h2, const float* xmid, const bf16* w1, const bf16* w2, const float* b1, const float* b2, bf16* u, bf16* g, float* y,
                  const float* ln_w, const float* ln_b, bf16* h1, float* ln_stats, int64_t R, int D, int M, Drop drop_h, Drop drop_o, cudaStream_t st) {
    MSST_REQUIRE(mlp_block_supported(D, M), "mlp_block_fwd: needs mlp_dim 64 and D in {32, 64, 96}");
    MSST_REQUIRE(R < (int64_t)2147483647 - 256, "mlp_block_fwd: too many rows for 32-bit TMA coordinates");
    if (R == 0) return MSST_OK;
    MlpParams p{};
    p.R = R; p.n_tiles = (R + 127) / 128; p.D = D; p.nch = D / 32;
    p.b1 = b1; p.b2 = b2; p.ln_w = ln_w; p.ln_b = ln_b; p.ln_stats = ln_stats; p.has_ln = h1 != nullptr;
    p.drop_h = drop_h; p.drop_o = drop_o;
    CUtensorMap t_h2, t_x, t_w1, t_w2, t_u, t_g, t_y, t_h1;
    auto map2 = [&](CUtensorMap* m, const void* base, int eb, int64_t rows, int64_t cols, int box_c, int box_r, int sw) {
        const int64_t dims[2] = {cols, rows}, strides[1] = {cols};
        const int box[2] = {box_c, box_r};
        return make_tmap_nd(m, base, eb, 2, dims, strides, box, sw);
    };
    if (int rc = map2(&t_h2, h2, 2, R, D, 32, 128, 64)) return rc;
    if (int rc = map2(&t_x, xmid, 4, R, D, 32, 128, 128)) return rc;
    if (int rc = map2(&t_w1, w1, 2, M, D, 32, M, 64)) return rc;
    if (int rc = map2(&t_w2, w2, 2, D, M, 64, D, 128)) return rc;
    if (int rc = map2(&t_u, u, 2, R, M, 64, 128, 128)) return rc;
    if (int rc = map2(&t_g, g, 2, R, M, 64, 128, 128)) return rc;
    if (int rc = map2(&t_y, y, 4, R, D, 32, 128, 128)) return rc;
    if (int rc = map2(&t_h1, h1 ? h1 : const_cast<bf16*>(h2), 2, R, D, 32, 128, 64)) return rc;
    const int nch = D / 32;
    const size_t smem = (size_t)nch * 4096 + (size_t)D * 128 + 2 * (size_t)nch * 8192 + 2 * (size_t)nch * MB_T16 + 2 * MB_T16 + 2 * 4 * 128 * sizeof(float) + sizeof(MlpBars);
    static PerDeviceOnce once;
    if (once.first()) {
        MSST_CUDA(cudaFuncSetAttribute(mlp_block_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        MSST_CUDA(cudaFuncSetAttribute(mlp_block_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        MSST_CUDA(cudaFuncSetAttribute(mlp_block_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    }
    const int grid = (int)(p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs);
    switch (nch) {
        case 1: mlp_block_fwd_kernel<1><<<grid, MB_THREADS, smem, st>>>(t_h2, t_x, t_w1, t_w2, t_u, t_g, t_y, t_h1, p); break;
        case 2: mlp_block_fwd_kernel<2><<<grid, MB_THREADS, smem, st>>>(t_h2, t_x, t_w1, t_w2, t_u, t_g, t_y, t_h1, p); break;
        default: mlp_block_fwd_kernel<3><<<grid, MB_THREADS, smem, st>>>(t_h2, t_x, t_w1, t_w2, t_u, t_g, t_y, t_h1, p); break;
    }
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}

// dyb [R,D] bf16 (= bf16(dropout_out(dy))), u / g [R,64] bf16 (saved by the forward), h2 [R,D] bf16, w2t = W2^T [64,D] bf16, w1t = W1^T [D,64] bf16
// -> dW1 [64,D] += , dW2 [D,64] += , db1 [64] += (fp32), dh [R,D] fp32
int mlp_block_bwd(const bf16* dyb, const bf16* u, const bf16* g, const bf16* h2, const bf16* w2t, const bf16* w1t, float* dw1, float* dw2, float* db1,
                  float* dh, int64_t R, int D, int M, Drop drop_h, cudaStream_t st) {
    MSST_REQUIRE(mlp_block_supported(D, M), "mlp_block_bwd: needs mlp_dim 64 and D in {32, 64, 96}");
    MSST_REQUIRE(R < (int64_t)2147483647 - 256, "mlp_block_bwd: too many rows for 32-bit TMA coordinates");
    if (R == 0) return MSST_OK;
    MlpBwdParams p{};
    p.R = R; p.n_tiles = (R + 127) / 128; p.dw1 = dw1; p.dw2 = dw2; p.db1 = db1; p.drop_h = drop_h;
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("MSST_MLPB_DBG"); dbg = e ? atoi(e) : 0; } p.dbg = dbg; }
    static long long* tl_dev = nullptr;
    if ((p.dbg & 8) && !tl_dev) cudaMalloc(&tl_dev, 32 * 8 * 8);
    if (p.dbg & 8) { p.tl = tl_dev; cudaMemsetAsync(tl_dev, 0, 32 * 8 * 8, st); }
    CUtensorMap t_dy, t_g, t_h2, t_u, t_w2t, t_w1t, t_dh;
    auto map2 = [&](CUtensorMap* m, const void* base, int eb, int64_t rows, int64_t cols, int box_c, int box_r, int sw) {
        const int64_t dims[2] = {cols, rows}, strides[1] = {cols};
        const int box[2] = {box_c, box_r};
        return make_tmap_nd(m, base, eb, 2, dims, strides, box, sw);
    };
    if (int rc = map2(&t_dy, dyb, 2, R, D, 64, 128, 128)) return rc;     // columns beyond D: zero-filled by TMA
    if (int rc = map2(&t_g, g, 2, R, M, 64, 128, 128)) return rc;
    if (int rc = map2(&t_h2, h2, 2, R, D, 64, 128, 128)) return rc;
    if (int rc = map2(&t_u, u, 2, R, M, 64, 128, 128)) return rc;
    if (int rc = map2(&t_w2t, w2t, 2, M, D, 64, M, 128)) return rc;
    if (int rc = map2(&t_w1t, w1t, 2, D, M, 64, D, 128)) return rc;
    if (int rc = map2(&t_dh, dh, 4, R, D, 32, 128, 128)) return rc;
    const int nc = (D + 63) / 64;
    const size_t smem = 2 * ((size_t)(2 * nc + 2) * MB_T16) + (size_t)nc * 8192 + (size_t)D * 128 + 4 * 64 * sizeof(float) + sizeof(MlpBwdBars);
    static PerDeviceOnce once;
    if (once.first()) {
        MSST_CUDA(cudaFuncSetAttribute(mlp_block_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        MSST_CUDA(cudaFuncSetAttribute(mlp_block_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        MSST_CUDA(cudaFuncSetAttribute(mlp_block_bwd_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    }
    const int grid = (int)(p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs);
    switch (D) {
        case 32: mlp_block_bwd_kernel<32><<<grid, MB_THREADS, smem, st>>>(t_dy, t_g, t_h2, t_u, t_w2t, t_w1t, t_dh, p); break;
        case 64: mlp_block_bwd_kernel<64><<<grid, MB_THREADS, smem, st>>>(t_dy, t_g, t_h2, t_u, t_w2t, t_w1t, t_dh, p); break;
        default: mlp_block_bwd_kernel<96><<<grid, MB_THREADS, smem, st>>>(t_dy, t_g, t_h2, t_u, t_w2t, t_w1t, t_dh, p); break;
    }
    if (p.dbg & 8) {
        static int n = 0;
        if (++n == 5) {   // one warm launch
            static long long h[32 * 8];
            cudaStreamSynchronize(st);
            cudaMemcpy(h, tl_dev, sizeof(h), cudaMemcpyDeviceToHost);
            printf("mlp_block_bwd timeline, CTA 0 (clks rel. to the first load): load issue | in_full | acc1_full | du_ready | acc3_full | out_ready | stage_free | MMAs issued\n");
            for (int i = 0; i < 20; ++i) {
                printf("%2d:", i);
                for (int k = 0; k < 8; ++k) printf(" %7lld", h[i * 8 + k] ? h[i * 8 + k] - h[0] : -1);
                printf("\n");
            }
            fflush(stdout);
        }
    }
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}

}  // namespace msst
