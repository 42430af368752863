// Kernel (3), bf16 mode, N <= 64: attention forward on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// One CTA works on 128 slots (two 64-slot groups of attn_geom.cuh) x one head at a time and loops persistently over
// (tile, head) items.  Per item:
//   producer       : TMA (cp.async.bulk.tensor, one 128x64 box per operand) when the tile's slots are consecutive rows
//                    (spatial stack), otherwise two warps gather the strided rows (spectral stack) with cp.async into the
//                    same SWIZZLE_128B K-major layout, completion signalled on an mbarrier (cp.async.mbarrier.arrive)
//   MMA warp       : S[128x128] = Q K^T  -> TMEM (tcgen05.mma, one elected thread);   later  O[128x64] = P V -> TMEM
//                    (P from smem K-major, V as MN-major B operand: no transpose anywhere)
//   softmax warps  : two ping-pong warpgroups (even / odd items), thread = row.  tcgen05.ld of the row's own 64-key block, mask (same sequence), exp2, row sum and Philox-free
//                    pair-hash dropout entirely thread-local (no shuffles), P (bf16) -> smem; epilogue: O row / l -> global, lse
// Only the two diagonal 64x64 blocks of S are meaningful (block-diagonal mask); the off-diagonal halves of the P tile are
// zeroed once and never written.  smem/TMEM stages are double-buffered so the tensor work of item i+1 overlaps the softmax of i.
// The kernel is HBM-bound: 3*128 B in + 128 B out per (slot, head).
#include "common.cuh"
#include "kernels.h"
#include "attn_geom.cuh"
#include "ptx.cuh"
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

namespace msst {
using namespace ptx;
typedef __nv_bfloat16 bf16;
int make_tmap_bf16(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);   // gemm_bf16.cu
int make_tmap_bf16_4d(CUtensorMap* m, const void* base, const int64_t dims[4], const int64_t strides[3], int box1, int box2);

constexpr int TC_ROWS = 128;
constexpr int TC_THREADS = 384;            // warps 0-3 softmax group A (even items), 4-5 producers, 6 MMA, 7 TMEM alloc, 8-11 softmax group B
constexpr int TC_PRODUCERS = 64;
constexpr uint32_t TC_TILE = TC_ROWS * 128;   // bytes of one [128 rows][64 bf16] tile (16 KB)

struct alignas(8) TcBars {
    uint64_t full[2], kv_empty[2], s_full[2], p_full[2], o_full[2], tmem_free[2];
    uint32_t tmem_base;
};

// smem: [2 stages][Q,K,V][16 KB] | [2 stages][P chunk0, chunk1][16 KB] | barriers
constexpr size_t kTcSmem = 2 * 3 * TC_TILE + 2 * 2 * TC_TILE + sizeof(TcBars) + 1024;

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// dropout pair indexing / hash identical to attention_bf16.cu (the backward kernel regenerates the same mask)
__device__ __forceinline__ uint64_t tile_pair_base_tc(const AttnGeom& g, int64_t group, int h) {
    return ((uint64_t)group * g.H + h) * (uint64_t)(TS * TS / 2);
}
__device__ __forceinline__ uint32_t tc_pair_hash(const Drop& d, uint64_t idx) {
    const uint64_t s = d.seed + (d.seed_dev ? __ldg(d.seed_dev) : 0ull);
    uint32_t x = ((uint32_t)idx + (uint32_t)(s >> 32)) * 0x9E3779B1u ^ ((uint32_t)(idx >> 32) * 0x85EBCA77u) ^ (uint32_t)s ^ (d.site * 0xC2B2AE3Du);
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
// global row (or -1) of slot r (0..127) of a tile
__device__ __forceinline__ int64_t tc_row(const AttnGeom& g, int64_t tile, int r, int& slot_lo) {
    const int64_t group = tile * 2 + (r >> 6);
    int64_t seq; int pos;
    const bool ok = group < g.groups && slot_to(g, group, 0, r & 63, seq, pos);
    slot_lo = ((r & 63) / g.N) * g.N;
    return ok ? row_of(g, seq, pos) : -1;
}

// use_tma != 0: the 128 slots of a tile are 128 consecutive rows (inner == 1, N divides 64) -> one TMA box per operand
__global__ void __launch_bounds__(TC_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tma_qkv, AttnGeom g, const bf16* __restrict__ qkv, bf16* __restrict__ out,
                   float* __restrict__ lse, Drop drop, int64_t n_tiles, int use_tma) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* qkv_s = smem;                                  // [2][3][16 KB]
    uint8_t* p_s = smem + 2 * 3 * TC_TILE;                  // [2][2][16 KB]
    TcBars* bars = reinterpret_cast<TcBars*>(p_s + 2 * 2 * TC_TILE);
    const int warp = threadIdx.x >> 5;
    const int I = g.H * 64;
    const int64_t ld = 3 * (int64_t)I;

    if (warp == 6 && elect_one()) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars->full[s], use_tma ? 1 : TC_PRODUCERS); mbar_init(&bars->kv_empty[s], 1); mbar_init(&bars->s_full[s], 1);
            mbar_init(&bars->p_full[s], 128); mbar_init(&bars->o_full[s], 1); mbar_init(&bars->tmem_free[s], 128);
        }
        fence_barrier_init();
    }
    if (warp == 4 && use_tma && elect_one()) prefetch_tmap(&tma_qkv);
    if (warp == 7) tmem_alloc(&bars->tmem_base, 512);
    // zero Q/K/V stages and the P tiles once: the off-diagonal halves of P are never written afterwards
    for (uint32_t i = threadIdx.x; i < (2 * 3 + 2 * 2) * TC_TILE / 16; i += TC_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0), idesc_o = make_idesc_bf16(128, 64, 0, 1);

    // items of this CTA: it = 0 .. n_items-1  <->  (tile = blockIdx.x + (it / H) * gridDim.x, head = it % H);
    // stage = it & 1, k-th use of a stage = it >> 1 (barrier phase parity (it >> 1) & 1)
    const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int64_t n_items = my_tiles * g.H;

    if (warp == 4 || warp == 5) {
        // ===== producers =====
        if (use_tma) {
            if (warp == 4 && elect_one()) {
                for (int64_t it = 0; it < n_items; ++it) {
                    const int st = (int)(it & 1); const uint32_t ph = (uint32_t)(it >> 1) & 1;
                    const int64_t tile = blockIdx.x + (it / g.H) * gridDim.x; const int h = (int)(it % g.H);
                    mbar_wait(&bars->kv_empty[st], ph ^ 1);
                    uint8_t* base = qkv_s + (size_t)st * 3 * TC_TILE;
                    mbar_arrive_expect_tx(&bars->full[st], 3 * TC_TILE);
                    tma_load_2d(base, &tma_qkv, &bars->full[st], h * 64, (int)(tile * TC_ROWS));
                    tma_load_2d(base + TC_TILE, &tma_qkv, &bars->full[st], I + h * 64, (int)(tile * TC_ROWS));
                    tma_load_2d(base + 2 * TC_TILE, &tma_qkv, &bars->full[st], 2 * I + h * 64, (int)(tile * TC_ROWS));
                }
            }
        } else {
            // gather: thread covers 16-byte piece pc of rows (pt >> 3) + 8k, k = 0..15
            const int pt = threadIdx.x - 128, pc = pt & 7;
            int64_t prow[16];
            int64_t cur_tile = -1;
            for (int64_t it = 0; it < n_items; ++it) {
                const int st = (int)(it & 1); const uint32_t ph = (uint32_t)(it >> 1) & 1;
                const int64_t tile = blockIdx.x + (it / g.H) * gridDim.x; const int h = (int)(it % g.H);
                if (tile != cur_tile) {
                    cur_tile = tile;
#pragma unroll
                    for (int k = 0; k < 16; ++k) { int lo; prow[k] = tc_row(g, tile, (pt >> 3) + 8 * k, lo); }
                }
                mbar_wait(&bars->kv_empty[st], ph ^ 1);
                const uint32_t base = smem_u32(qkv_s + (size_t)st * 3 * TC_TILE);
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int r = (pt >> 3) + 8 * k;
                    const bf16* src = qkv + (prow[k] < 0 ? 0 : prow[k]) * ld + h * 64 + pc * 8;
                    const uint32_t dst = base + r * 128 + ((pc ^ (r & 7)) << 4);
                    cp_async16_zfill(dst, src, prow[k] >= 0);
                    cp_async16_zfill(dst + TC_TILE, src + I, prow[k] >= 0);
                    cp_async16_zfill(dst + 2 * TC_TILE, src + 2 * I, prow[k] >= 0);
                }
                cp_async_mbar_arrive(&bars->full[st]);
            }
        }
    } else if (warp == 6) {
        // ===== MMA issuer: S(it+1) is issued before waiting for the softmax of item it =====
        if (elect_one()) {
            auto issue_s = [&](int64_t it) {
                const int st = (int)(it & 1); const uint32_t ph = (uint32_t)(it >> 1) & 1;
                const uint32_t qb = smem_u32(qkv_s + (size_t)st * 3 * TC_TILE);
                mbar_wait(&bars->tmem_free[st], ph ^ 1);       // S / O TMEM stage drained by the epilogue of item it - 2
                mbar_wait(&bars->full[st], ph);                // Q, K, V landed
                fence_proxy_async();
                tc_fence_after();
                const uint64_t dq = make_smem_desc(qb, 16, 1024), dk = make_smem_desc(qb + TC_TILE, 16, 1024);
                for (int k = 0; k < 4; ++k)                    // S = Q K^T, K = 64 = 4 x UMMA_K
                    umma_bf16(tmem_base + st * 128, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k != 0);
                umma_commit(&bars->s_full[st]);
            };
            if (n_items > 0) issue_s(0);
            for (int64_t it = 0; it < n_items; ++it) {
                const int st = (int)(it & 1); const uint32_t ph = (uint32_t)(it >> 1) & 1;
                if (it + 1 < n_items) issue_s(it + 1);
                const uint32_t qb = smem_u32(qkv_s + (size_t)st * 3 * TC_TILE);
                const uint32_t pb = smem_u32(p_s + (size_t)st * 2 * TC_TILE);
                mbar_wait(&bars->p_full[st], ph);              // P written by the softmax warps (generic proxy + their fence)
                fence_proxy_async();
                tc_fence_after();
                const uint64_t dv = make_smem_desc(qb + 2 * TC_TILE, TC_TILE, 1024);
                for (int k = 0; k < 8; ++k) {                  // O = P V, K = 128 keys = 8 x UMMA_K
                    const uint64_t dp = make_smem_desc(pb + (k >> 2) * TC_TILE, 16, 1024) + (uint64_t)((k & 3) * 2);
                    umma_bf16(tmem_base + 256 + st * 64, dp, dv + (uint64_t)(k * 128), idesc_o, k != 0);
                }
                umma_commit(&bars->o_full[st]);
                umma_commit(&bars->kv_empty[st]);              // Q, K, V (and P) of this stage are free again
            }
        }
    } else if (warp < 4 || warp >= 8) {
        // ===== two softmax + epilogue warpgroups (ping-pong): group A = even items / stage 0, group B = odd items / stage 1.
        // thread r = row r of the tile; everything row-wise is thread-local (no shuffles) =====
        const int grp = warp >= 8 ? 1 : 0;
        const int r = threadIdx.x & 127;
        const int blk = r >> 6;                                 // which 64-key block (== which slot group of the tile)
        const int st = grp;
        const float sl2 = g.scale * 1.4426950408889634f;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        int64_t cur_tile = -1, grow = -1; int klo = 0, khi = 0;
        for (int64_t it = grp; it < n_items; it += 2) {
            const uint32_t ph = (uint32_t)(it >> 1) & 1;
            const int64_t tile = blockIdx.x + (it / g.H) * gridDim.x; const int h = (int)(it % g.H);
            if (tile != cur_tile) {
                cur_tile = tile;
                grow = tc_row(g, tile, r, klo);
                khi = grow >= 0 ? klo + g.N : 0;
            }
            const int64_t group = tile * 2 + blk;
            mbar_wait(&bars->s_full[st], ph);
            tc_fence_after();
            float s[64];
            {
                uint32_t a[32], b[32];
                const uint32_t scol = tmem_base + lane_base + st * 128 + blk * 64;
                tmem_ld_32x32(scol, a);
                tmem_ld_32x32(scol + 32, b);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) { s[j] = __uint_as_float(a[j]); s[32 + j] = __uint_as_float(b[j]); }
            }
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                s[j] = (j >= klo && j < khi) ? s[j] * sl2 : -INFINITY;
                mx = fmaxf(mx, s[j]);
            }
            const float sub = mx == -INFINITY ? 0.f : mx;
            float l = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) { s[j] = exp2f(s[j] - sub); l += s[j]; }
            if (drop.on()) {
                const uint64_t base = tile_pair_base_tc(g, group, h) + (uint64_t)((r & 63) * 32);
                const uint32_t t16 = drop.thresh >> 16;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const uint32_t hsh = tc_pair_hash(drop, base + j);
                    s[2 * j] *= (hsh & 0xFFFFu) >= t16 ? drop.scale : 0.f;
                    s[2 * j + 1] *= (hsh >> 16) >= t16 ? drop.scale : 0.f;
                }
            }
            // P row (bf16) -> smem P tile of this stage: chunk = own key block, SWIZZLE_128B K-major
            {
                uint8_t* prow = p_s + (size_t)st * 2 * TC_TILE + (size_t)blk * TC_TILE + r * 128;
#pragma unroll
                for (int pc = 0; pc < 8; ++pc)
                    *reinterpret_cast<uint4*>(prow + ((pc ^ (r & 7)) << 4)) =
                        make_uint4(pack_bf(s[pc * 8], s[pc * 8 + 1]), pack_bf(s[pc * 8 + 2], s[pc * 8 + 3]),
                                   pack_bf(s[pc * 8 + 4], s[pc * 8 + 5]), pack_bf(s[pc * 8 + 6], s[pc * 8 + 7]));
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&bars->p_full[st]);
            // ---- epilogue of the same item (the other group works on the next item meanwhile) ----
            mbar_wait(&bars->o_full[st], ph);
            tc_fence_after();
            uint32_t v[32], w[32];
            tmem_ld_32x32(tmem_base + lane_base + 256 + st * 64, v);
            tmem_ld_32x32(tmem_base + lane_base + 256 + st * 64 + 32, w);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bars->tmem_free[st]);
            if (grow >= 0) {
                const float inv = l > 0.f ? 1.f / l : 0.f;
                bf16* dst = out + grow * I + h * 64;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<uint4*>(dst + c * 8) =
                        make_uint4(pack_bf(__uint_as_float(v[c * 8]) * inv, __uint_as_float(v[c * 8 + 1]) * inv),
                                   pack_bf(__uint_as_float(v[c * 8 + 2]) * inv, __uint_as_float(v[c * 8 + 3]) * inv),
                                   pack_bf(__uint_as_float(v[c * 8 + 4]) * inv, __uint_as_float(v[c * 8 + 5]) * inv),
                                   pack_bf(__uint_as_float(v[c * 8 + 6]) * inv, __uint_as_float(v[c * 8 + 7]) * inv));
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<uint4*>(dst + 32 + c * 8) =
                        make_uint4(pack_bf(__uint_as_float(w[c * 8]) * inv, __uint_as_float(w[c * 8 + 1]) * inv),
                                   pack_bf(__uint_as_float(w[c * 8 + 2]) * inv, __uint_as_float(w[c * 8 + 3]) * inv),
                                   pack_bf(__uint_as_float(w[c * 8 + 4]) * inv, __uint_as_float(w[c * 8 + 5]) * inv),
                                   pack_bf(__uint_as_float(w[c * 8 + 6]) * inv, __uint_as_float(w[c * 8 + 7]) * inv));
                lse[grow * g.H + h] = (sub + log2f(l)) * 0.6931471805599453f;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 7) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int attention_fwd_tc(const AttnGeom& g, const bf16* qkv, bf16* out, float* lse, Drop drop, cudaStream_t st) {
    const int64_t n_tiles = (g.groups + 1) / 2;
    static PerDeviceOnce once;
    if (once.first()) MSST_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
    // contiguous tiles (e.g. the spatial stack: inner == 1, full 64-slot groups): TMA; otherwise a cp.async row gather
    const int use_tma = (g.inner == 1 && 64 % g.N == 0 && n_tiles * TC_ROWS < (int64_t)2147483647) ? 1 : 0;
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    if (use_tma) {
        const int64_t R = g.n_seq * g.N;
        if (int rc = make_tmap_bf16(&tm, qkv, R, 3 * (int64_t)g.H * 64, 3 * (int64_t)g.H * 64, TC_ROWS)) return rc;
    }
    const int grid = (int)(n_tiles < kNumSMs ? n_tiles : kNumSMs);
    attn_fwd_tc_kernel<<<grid, TC_THREADS, kTcSmem, st>>>(tm, g, qkv, out, lse, drop, n_tiles, use_tma);
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}

// =========================================================================================================
// Backward on tcgen05 / TMEM for all packed short-sequence shapes (N <= 64): both the spatial and the spectral stack.
// A tile = two 64-slot groups; each group's rows arrive as ONE 4-D TMA box per operand (attn_geom.cuh: the group's slots are
// runs of consecutive rows in the (col, s, pos, blk) view of the activation matrix), so strided sequences cost nothing extra.
//
// Per (128-slot tile, head) item, all five contractions run on the tensor cores from one elected thread:
//   S  = Q K^T, dP = dO V^T            -> TMEM [128 x 128] fp32 each (K-major operands straight from the TMA tiles)
//   softmax warps (thread = query row): P = exp2(S*scale*log2e - lse), dropout mask regenerated from the pair hash,
//       D_i = sum_j P~_ij dP_ij in registers (no O re-read), P~ and dS = P (f dP - D) scale -> smem (bf16, K-major rows)
//   dV = P~^T dO, dK = dS^T Q          -> P~ / dS consumed as MN-major A operands, dO / Q as MN-major B (no transposes)
//   dQ = dS K                          -> dS K-major A, K as MN-major B
// dV | dK | dQ accumulate into the TMEM columns S / dP occupied (dead once the softmax warps hold them in registers), so an
// item needs 256 columns and two items are in flight: the S/dP MMAs of item i+1 overlap the softmax of item i (two
// ping-pong warpgroups).  Results leave through the dead Q/K/V tiles of the stage: bf16 rows -> smem -> TMA store.
// Only the diagonal 64x64 blocks of S / dP are read; the off-diagonal halves of the P~ / dS tiles stay zero.
// HBM-bound: (4 + 3) * 128 B per (slot, head).
// =========================================================================================================
struct alignas(8) TcbBars {
    uint64_t full[2], kv_empty[2], s_full[2], o_full[2], tmem_free[2];
    // events of the single P~/dS buffer and the single staging buffer, one barrier per item parity: with ONE barrier a waiter
    // that runs two items ahead would take the completion of item i - 2 for that of item i (same phase parity)
    uint64_t p_full[2], pds_free[2], stg_free[2], stg_full[2];
    uint32_t tmem_base;
};
constexpr int TCB_THREADS = 608;           // warps 0-7 softmax group A, 8-15 group B, 16 TMA producer (+ TMEM alloc), 17 MMA issuer, 18 TMA store
// P~ / dS tiles [128 q][128 k] are block diagonal.  Each is kept as two OVERLAPPING virtual [128 rows][64 keys] chunks
//   V0 = [C0: rows 0-63, keys 0-63][Z: 8 KB of zeros]      V1 = V0 + 8 KB = [Z][C1: rows 64-127, keys 64-127]
// (24 KB instead of 32 KB: both chunks share the zero block), so the chunk stride of the UMMA descriptors is 8 KB.
constexpr uint32_t TCB_CHUNK = 8192, TCB_PD = 3 * 8192;
constexpr int TCB_PREFETCH = 1;            // items ahead whose operand boxes are prefetched into L2 (MSST_ATTN_PF overrides; >= 3 thrashes)
// smem: [2 stages][Q, K, V, dO][16 KB] | P~ (24 KB) | dS (24 KB) | output staging dQ, dK, dV [16 KB] | partial row sums | barriers
// = 226.1 KB: no slack for re-aligning the base, the kernel checks that the dynamic window starts 1 KB aligned
constexpr size_t kTcbSmem = 2 * 4 * TC_TILE + 2 * TCB_PD + 3 * TC_TILE + 2 * 2 * 128 * sizeof(float) + sizeof(TcbBars);

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// One thread's share of an item: 32 of the 64 diagonal-block columns of row r.  p[] holds S, q[] holds dP on entry.
// Packs P~ (bf16 pairs) into pfp[], returns the partial row sum D; leaves p = P * scale and q = f * dP for the second pass.
template <bool FULL, bool DROP>
__device__ __forceinline__ float tcb_softmax_pass1(float (&p)[32], float (&q)[32], uint32_t (&pfp)[16], float sl2, float L, float scale,
                                                   uint32_t vmask, uint32_t hash_lo, uint32_t hash_hi, uint32_t t16, float dscale) {
    float D = 0.f;
#pragma unroll
    for (int pc = 0; pc < 4; ++pc) {
        float pf[8];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            const int j = pc * 8 + e;
            float f0 = 1.f, f1 = 1.f;
            if (DROP) {   // lowbias32 over (pair index, seed, site): identical to tc_pair_hash (the pair index never carries into the high word)
                uint32_t x = (hash_lo + (uint32_t)(j >> 1)) * 0x9E3779B1u ^ hash_hi;
                x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
                f0 = (x & 0xFFFFu) >= t16 ? dscale : 0.f;
                f1 = (x >> 16) >= t16 ? dscale : 0.f;
            }
            float p0 = ex2_approx(fmaf(p[j], sl2, -L)), p1 = ex2_approx(fmaf(p[j + 1], sl2, -L));
            if (!FULL) {   // vmask bit j: column j of this thread is a key of the row's own sequence
                p0 = (vmask >> j) & 1u ? p0 : 0.f;
                p1 = (vmask >> (j + 1)) & 1u ? p1 : 0.f;
            }
            pf[e] = DROP ? p0 * f0 : p0; pf[e + 1] = DROP ? p1 * f1 : p1;
            D = fmaf(pf[e], q[j], D); D = fmaf(pf[e + 1], q[j + 1], D);
            p[j] = p0 * scale; p[j + 1] = p1 * scale;
            if (DROP) { q[j] *= f0; q[j + 1] *= f1; }
        }
        pfp[pc * 4] = pack_bf(pf[0], pf[1]); pfp[pc * 4 + 1] = pack_bf(pf[2], pf[3]);
        pfp[pc * 4 + 2] = pack_bf(pf[4], pf[5]); pfp[pc * 4 + 3] = pack_bf(pf[6], pf[7]);
    }
    return D;
}

// TMA coordinates (c1, c2, c3) of a slot group's box (see attention_bwd_tc for the two tensor views); a group past the end maps
// to an out-of-range coordinate: zero-filled on load, clipped on store
__device__ __forceinline__ void tcb_group_coords(const AttnGeom& g, int64_t group, int& c1, int& c2, int& c3) {
    if (g.gpb == 0) { c1 = 0; c2 = (int)(group * g.G); c3 = 0; }
    else { c1 = (int)(group % g.gpb) * g.G; c2 = 0; c3 = (int)(group / g.gpb); }
}

__global__ void __launch_bounds__(TCB_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do,
                   const __grid_constant__ CUtensorMap tma_dqkv, AttnGeom g, const float* __restrict__ lse, Drop drop, int64_t n_tiles, int nbox, int pf_dist, long long* dbg) {
#define TCB_T(it, slot) do { if (dbg && blockIdx.x == 0 && (it) < 32) dbg[(it) * 16 + (slot)] = clock64(); } while (0)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (smem_u32(smem) & 1023u) __trap();                   // SWIZZLE_128B tiles need 1 KB alignment (see kTcbSmem)
    uint8_t* op_s = smem;                                   // [2][4][16 KB]
    uint8_t* p_s = smem + 2 * 4 * TC_TILE;                  // P~ (24 KB), dS (24 KB)
    uint8_t* stg_s = p_s + 2 * TCB_PD;                      // dQ, dK, dV staging for the TMA stores
    float* dpart = reinterpret_cast<float*>(stg_s + 3 * TC_TILE);   // [group][half][128] partial row sums
    TcbBars* bars = reinterpret_cast<TcbBars*>(dpart + 2 * 2 * 128);
    const int warp = threadIdx.x >> 5;
    const int I = g.H * 64;

    if (warp == 17 && elect_one()) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars->full[s], 1); mbar_init(&bars->kv_empty[s], 1); mbar_init(&bars->s_full[s], 1);
            mbar_init(&bars->o_full[s], 1); mbar_init(&bars->tmem_free[s], 256);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars->p_full[s], 256); mbar_init(&bars->pds_free[s], 1); mbar_init(&bars->stg_free[s], 1); mbar_init(&bars->stg_full[s], 256);
        }
        fence_barrier_init();
    }
    if (warp == 16) {
        tmem_alloc(&bars->tmem_base, 512);
        if (elect_one()) { prefetch_tmap(&tma_qkv); prefetch_tmap(&tma_do); prefetch_tmap(&tma_dqkv); }
    }
    // zero once: the operand rows a group's box does not cover (slots G*N .. 63) and the shared zero blocks of the P~ / dS tiles
    for (uint32_t i = threadIdx.x; i < (2 * 4 * TC_TILE + 2 * TCB_PD) / 16; i += TCB_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0), idesc_mn = make_idesc_bf16(128, 64, 1, 1), idesc_q = make_idesc_bf16(128, 64, 0, 1);

    const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int64_t n_items = my_tiles * g.H;
    // nbox == 2: one box per slot group; nbox == 1: the tile's 128 slots are 128 consecutive rows (inner == 1, N | 64) -> one box
    const uint32_t box_bytes = (uint32_t)(g.G * g.N) * 128u * (nbox == 1 ? 2u : 1u);

    if (warp == 16) {
        // ===== TMA producer =====
        if (elect_one()) {
            for (int64_t it = 0; it < n_items; ++it) {
                const int st = (int)(it & 1); const uint32_t ph = (uint32_t)(it >> 1) & 1;
                const int64_t tile = blockIdx.x + (it / g.H) * gridDim.x; const int h = (int)(it % g.H);
                mbar_wait(&bars->kv_empty[st], ph ^ 1);
                TCB_T(it, 0);
                uint8_t* base = op_s + (size_t)st * 4 * TC_TILE;
                mbar_arrive_expect_tx(&bars->full[st], 4 * nbox * box_bytes);
                if (nbox == 1) {                                 // 128 consecutive rows: plain 2-D boxes
                    const int row0 = (int)(tile * TC_ROWS);
                    tma_load_2d(base, &tma_qkv, &bars->full[st], h * 64, row0);
                    tma_load_2d(base + TC_TILE, &tma_qkv, &bars->full[st], I + h * 64, row0);
                    tma_load_2d(base + 2 * TC_TILE, &tma_qkv, &bars->full[st], 2 * I + h * 64, row0);
                    tma_load_2d(base + 3 * TC_TILE, &tma_do, &bars->full[st], h * 64, row0);
                } else
                for (int gi = 0; gi < nbox; ++gi) {              // the tile's two slot groups -> rows 0.. and 64.. of every operand tile
                    int c1, c2, c3;
                    tcb_group_coords(g, tile * 2 + gi, c1, c2, c3);
                    uint8_t* dst = base + gi * 64 * 128;
                    tma_load_4d(dst, &tma_qkv, &bars->full[st], h * 64, c1, c2, c3);
                    tma_load_4d(dst + TC_TILE, &tma_qkv, &bars->full[st], I + h * 64, c1, c2, c3);
                    tma_load_4d(dst + 2 * TC_TILE, &tma_qkv, &bars->full[st], 2 * I + h * 64, c1, c2, c3);
                    tma_load_4d(dst + 3 * TC_TILE, &tma_do, &bars->full[st], h * 64, c1, c2, c3);
                }
                if (pf_dist > 0 && it + pf_dist < n_items) {   // pull a later item's operand boxes into L2 now (its smem stage is still busy)
                    const int64_t it2 = it + pf_dist;
                    const int64_t tile2 = blockIdx.x + (it2 / g.H) * gridDim.x; const int h2 = (int)(it2 % g.H);
                    if (nbox == 1) {
                        const int row2 = (int)(tile2 * TC_ROWS);
                        tma_prefetch_l2_2d(&tma_qkv, h2 * 64, row2);
                        tma_prefetch_l2_2d(&tma_qkv, I + h2 * 64, row2);
                        tma_prefetch_l2_2d(&tma_qkv, 2 * I + h2 * 64, row2);
                        tma_prefetch_l2_2d(&tma_do, h2 * 64, row2);
                    } else
                    for (int gi = 0; gi < nbox; ++gi) {
                        int c1, c2, c3;
                        tcb_group_coords(g, tile2 * 2 + gi, c1, c2, c3);
                        tma_prefetch_l2_4d(&tma_qkv, h2 * 64, c1, c2, c3);
                        tma_prefetch_l2_4d(&tma_qkv, I + h2 * 64, c1, c2, c3);
                        tma_prefetch_l2_4d(&tma_qkv, 2 * I + h2 * 64, c1, c2, c3);
                        tma_prefetch_l2_4d(&tma_do, h2 * 64, c1, c2, c3);
                    }
                }
            }
        }
    } else if (warp == 17) {
        // ===== MMA issuer =====
        if (elect_one()) {
            // Two kinds of work, issued in whatever order their inputs become ready (never block on one while the other
            // could run): phase 1 of item next_s (S, dP; needs its operands + a drained TMEM stage) and phase 2 of item
            // next_p (dV, dK, dQ; needs P~ / dS from the softmax warps).  Phase 1 runs at most one item ahead.
            const uint32_t pb = smem_u32(p_s), dsb = pb + TCB_PD;
            int64_t next_s = 0, next_p = 0;
            while (next_p < n_items) {
                if (next_s < n_items && next_s <= next_p + 1) {
                    const int st = (int)(next_s & 1); const uint32_t ph = (uint32_t)(next_s >> 1) & 1;
                    if (mbar_try_wait(&bars->tmem_free[st], ph ^ 1) && mbar_try_wait(&bars->full[st], ph)) {
                        TCB_T(next_s, 1);
                        tc_fence_after();
                        const uint32_t qb = smem_u32(op_s + (size_t)st * 4 * TC_TILE);
                        const uint64_t dq = make_smem_desc(qb, 16, 1024), dk = make_smem_desc(qb + TC_TILE, 16, 1024);
                        const uint64_t dv = make_smem_desc(qb + 2 * TC_TILE, 16, 1024), dd = make_smem_desc(qb + 3 * TC_TILE, 16, 1024);
                        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + st * 256, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k != 0);          // S = Q K^T
                        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + st * 256 + 128, dd + (uint64_t)(k * 2), dv + (uint64_t)(k * 2), idesc_s, k != 0);    // dP = dO V^T
                        umma_commit(&bars->s_full[st]);
                        ++next_s;
                    }
                }
                if (next_p < next_s && mbar_try_wait(&bars->p_full[next_p & 1], (uint32_t)(next_p >> 1) & 1)) {   // P~ and dS written (generic proxy + the writers' fence)
                    const int st = (int)(next_p & 1);
                    TCB_T(next_p, 5);
                    const uint32_t qb = smem_u32(op_s + (size_t)st * 4 * TC_TILE);
                    fence_proxy_async();
                    tc_fence_after();
                    const uint64_t a_pf = make_smem_desc(pb, TCB_CHUNK, 1024), a_ds = make_smem_desc(dsb, TCB_CHUNK, 1024);
                    const uint64_t b_q = make_smem_desc(qb, TC_TILE, 1024), b_k = make_smem_desc(qb + TC_TILE, TC_TILE, 1024);
                    const uint64_t b_do = make_smem_desc(qb + 3 * TC_TILE, TC_TILE, 1024);
                    const uint32_t d0 = tmem_base + st * 256;
                    for (int k = 0; k < 8; ++k)   // dV[keys] = P~^T dO : reduction over the 128 queries, 16 rows (2 KB) per step
                        umma_bf16(d0, a_pf + (uint64_t)(k * 128), b_do + (uint64_t)(k * 128), idesc_mn, k != 0);
                    for (int k = 0; k < 8; ++k)   // dK[keys] = dS^T Q
                        umma_bf16(d0 + 64, a_ds + (uint64_t)(k * 128), b_q + (uint64_t)(k * 128), idesc_mn, k != 0);
                    for (int k = 0; k < 8; ++k) { // dQ[queries] = dS K : reduction over the 128 keys (2 chunks x 4 steps)
                        const uint64_t a = make_smem_desc(dsb + (k >> 2) * TCB_CHUNK, 16, 1024) + (uint64_t)((k & 3) * 2);
                        umma_bf16(d0 + 128, a, b_k + (uint64_t)(k * 128), idesc_q, k != 0);
                    }
                    umma_commit(&bars->o_full[st]);
                    umma_commit(&bars->pds_free[st]);
                    umma_commit(&bars->kv_empty[st]);          // Q, K, V, dO of this stage are dead: the next loads may start
                    ++next_p;
                }
            }
        }
    } else if (warp == 18) {
        // ===== TMA store of the staged dQ / dK / dV tiles (its own warp: nobody else ever waits for a store to drain) =====
        if (elect_one()) {
            for (int64_t it = 0; it < n_items; ++it) {
                const int64_t tile = blockIdx.x + (it / g.H) * gridDim.x; const int h = (int)(it % g.H);
                mbar_wait(&bars->stg_full[it & 1], (uint32_t)(it >> 1) & 1);
                TCB_T(it, 7);
                if (nbox == 1) {
                    const int row0 = (int)(tile * TC_ROWS);
                    tma_store_2d(&tma_dqkv, stg_s, h * 64, row0);
                    tma_store_2d(&tma_dqkv, stg_s + TC_TILE, I + h * 64, row0);
                    tma_store_2d(&tma_dqkv, stg_s + 2 * TC_TILE, 2 * I + h * 64, row0);
                } else
                for (int gi = 0; gi < nbox; ++gi) {
                    int c1, c2, c3;
                    tcb_group_coords(g, tile * 2 + gi, c1, c2, c3);
                    const uint8_t* src = stg_s + gi * 64 * 128;
                    tma_store_4d(&tma_dqkv, src, h * 64, c1, c2, c3);
                    tma_store_4d(&tma_dqkv, src + TC_TILE, I + h * 64, c1, c2, c3);
                    tma_store_4d(&tma_dqkv, src + 2 * TC_TILE, 2 * I + h * 64, c1, c2, c3);
                }
                tma_store_commit();
                tma_store_wait_read();
                TCB_T(it, 8);
                mbar_arrive(&bars->stg_free[it & 1]);
            }
        }
    } else if (warp < 16) {
        // ===== two softmax + epilogue groups of 8 warps (ping-pong over items).  Within a group, thread (r, half) owns
        // row r of the tile (query row in the softmax, key / query row in the epilogue) and 32 of its 64 columns. =====
        const int grp = warp >> 3, wl = warp & 7, half = wl >> 2;
        const int r = (wl & 3) * 32 + (threadIdx.x & 31);
        const int blk = r >> 6;
        const int st = grp;
        const int c0 = half * 32;
        const float sl2 = g.scale * 1.4426950408889634f;
        const uint32_t lane_base = (uint32_t)((wl & 3) * 32) << 16;
        const uint32_t swz = (uint32_t)(r & 7);
        const bool full_blocks = g.N == 64 && g.groups % 2 == 0;   // every row of every tile attends to its whole 64-key block (the spatial stack)
        const uint64_t seed = drop.seed + (drop.seed_dev ? __ldg(drop.seed_dev) : 0ull);
        const uint32_t t16 = drop.thresh >> 16;
        float* my_part = dpart + (grp * 2 + half) * 128 + r;
        const float* other_part = dpart + (grp * 2 + (half ^ 1)) * 128 + r;
        // which of this thread's 32 key columns belong to the sequence of slot (r & 63): sequence-major packing -> a run of N
        // slots, position-major -> every G-th slot.  Pure slot geometry: the same for every tile.
        uint32_t vmask_geom = 0;
        {
            const int slot = r & 63, used = g.G * g.N;
            for (int j = 0; j < 32; ++j) {
                const int c = c0 + j;
                const bool same = g.gpb == 0 ? (c / g.N == slot / g.N) : (c % g.G == slot % g.G);
                vmask_geom |= (uint32_t)(same && c < used && slot < used) << j;
            }
        }
        int64_t cur_tile = -1, grow = -1; uint32_t vmask = 0;
        for (int64_t it = grp; it < n_items; it += 2) {
            const uint32_t ph = (uint32_t)(it >> 1) & 1;
            const int64_t tile = blockIdx.x + (it / g.H) * gridDim.x; const int h = (int)(it % g.H);
            if (tile != cur_tile) {
                cur_tile = tile;
                int64_t seq; int pos;
                const int64_t group = tile * 2 + blk;
                const bool ok = group < g.groups && slot_to(g, group, 0, r & 63, seq, pos);
                grow = ok ? row_of(g, seq, pos) : -1;
                vmask = ok ? vmask_geom : 0u;
            }
            const float L = grow >= 0 ? __ldg(lse + grow * g.H + h) * 1.4426950408889634f : 0.f;
            // dropout pair index of (row, column pair jj): tile_pair_base_tc + (r & 63) * 32 + jj  -- split into the hash's 32-bit halves
            const uint64_t hidx = tile_pair_base_tc(g, tile * 2 + blk, h) + (uint64_t)((r & 63) * 32 + (c0 >> 1));
            const uint32_t hash_lo = (uint32_t)hidx + (uint32_t)(seed >> 32);
            const uint32_t hash_hi = ((uint32_t)(hidx >> 32) * 0x85EBCA77u) ^ (uint32_t)seed ^ (drop.site * 0xC2B2AE3Du);
            mbar_wait(&bars->s_full[st], ph);
            if (r == 0 && half == 0) TCB_T(it, 2);
            tc_fence_after();
            float p[32], q[32];
            {
                uint32_t a[32], b[32];
                const uint32_t scol = tmem_base + lane_base + st * 256 + blk * 64 + c0;
                tmem_ld_32x32(scol, a);
                tmem_ld_32x32(scol + 128, b);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) { p[j] = __uint_as_float(a[j]); q[j] = __uint_as_float(b[j]); }
            }
            if (r == 0 && half == 0) TCB_T(it, 3);
            uint32_t pfp[16], dsp[16];
            float D;
            if (drop.on()) {
                if (full_blocks) D = tcb_softmax_pass1<true, true>(p, q, pfp, sl2, L, g.scale, vmask, hash_lo, hash_hi, t16, drop.scale);
                else D = tcb_softmax_pass1<false, true>(p, q, pfp, sl2, L, g.scale, vmask, hash_lo, hash_hi, t16, drop.scale);
            } else {
                if (full_blocks) D = tcb_softmax_pass1<true, false>(p, q, pfp, sl2, L, g.scale, vmask, hash_lo, hash_hi, t16, drop.scale);
                else D = tcb_softmax_pass1<false, false>(p, q, pfp, sl2, L, g.scale, vmask, hash_lo, hash_hi, t16, drop.scale);
            }
            *my_part = D;
            asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory");
            D += *other_part;
#pragma unroll
            for (int j = 0; j < 16; ++j) dsp[j] = pack_bf(p[2 * j] * (q[2 * j] - D), p[2 * j + 1] * (q[2 * j + 1] - D));
            // everything above overlaps the previous item's dV / dK / dQ MMAs; the single P~ / dS buffer is free once they completed
            if (it > 0) mbar_wait(&bars->pds_free[(it - 1) & 1], (uint32_t)((it - 1) >> 1) & 1);
            uint8_t* prow = p_s + (size_t)blk * TCB_CHUNK + r * 128;
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) {
                const uint32_t off = ((uint32_t)(half * 4 + pc) ^ swz) << 4;
                *reinterpret_cast<uint4*>(prow + off) = make_uint4(pfp[pc * 4], pfp[pc * 4 + 1], pfp[pc * 4 + 2], pfp[pc * 4 + 3]);
                *reinterpret_cast<uint4*>(prow + TCB_PD + off) = make_uint4(dsp[pc * 4], dsp[pc * 4 + 1], dsp[pc * 4 + 2], dsp[pc * 4 + 3]);
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&bars->p_full[st]);
            if (r == 0 && half == 0) TCB_T(it, 4);
            // ---- epilogue: dQ / dK / dV rows (fp32, TMEM) -> bf16 -> staging tiles -> TMA store ----
            mbar_wait(&bars->o_full[st], ph);
            if (it > 0) mbar_wait(&bars->stg_free[(it - 1) & 1], (uint32_t)((it - 1) >> 1) & 1);   // the previous item's stores have read the staging tiles
            if (r == 0 && half == 0) TCB_T(it, 6);
            tc_fence_after();
            uint8_t* stage = stg_s;
#pragma unroll
            for (int m = 0; m < 3; ++m) {                       // TMEM columns: dV at +0, dK at +64, dQ at +128 -> tiles V, K, Q
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + lane_base + st * 256 + m * 64 + c0, v);
                tmem_ld_wait();
                uint8_t* trow = stage + (size_t)(2 - m) * TC_TILE + r * 128;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<uint4*>(trow + (((uint32_t)(half * 4 + c) ^ swz) << 4)) =
                        make_uint4(pack_bf(__uint_as_float(v[c * 8]), __uint_as_float(v[c * 8 + 1])), pack_bf(__uint_as_float(v[c * 8 + 2]), __uint_as_float(v[c * 8 + 3])),
                                   pack_bf(__uint_as_float(v[c * 8 + 4]), __uint_as_float(v[c * 8 + 5])), pack_bf(__uint_as_float(v[c * 8 + 6]), __uint_as_float(v[c * 8 + 7])));
            }
            tc_fence_before();
            mbar_arrive(&bars->tmem_free[st]);
            fence_proxy_async();
            mbar_arrive(&bars->stg_full[st]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
#undef TCB_T
}

bool attention_bwd_tc_supported(const AttnGeom& g) {
    const int64_t n_blk = g.inner > 0 ? g.n_seq / g.inner : 0;
    return g.tiles == 1 && g.groups > 0 && g.groups * g.G < (int64_t)2147483647 && n_blk < (int64_t)2147483647 &&
           ((g.groups + 1) / 2) * g.H < ((int64_t)1 << 40);
}

// 4-D view of an activation matrix [R, cols] whose box {64, box1, box2, 1} is one slot group (attn_geom.cuh):
//   inner == 1: (col, pos [N], seq [n_seq], 1)           box {64, N, G, 1}   -> slot = j * N + pos
//   inner  > 1: (col, s [inner], pos [N], blk [n_seq/inner])  box {64, G, N, 1}   -> slot = pos * G + j
static int tcb_tmap(CUtensorMap* m, const AttnGeom& g, const bf16* base, int64_t cols, int nbox) {
    if (nbox == 1) return make_tmap_bf16(m, base, g.n_seq * g.N, cols, cols, TC_ROWS);   // contiguous tiles: rank-2 map, 128-row box
    if (g.gpb == 0) {
        const int64_t dims[4] = {cols, g.N, g.n_seq, 1}, strides[3] = {cols, (int64_t)g.N * cols, g.n_seq * g.N * cols};
        return make_tmap_bf16_4d(m, base, dims, strides, g.N, g.G);
    }
    const int64_t dims[4] = {cols, g.inner, g.N, g.n_seq / g.inner};
    const int64_t strides[3] = {cols, (int64_t)g.inner * cols, (int64_t)g.N * g.inner * cols};
    return make_tmap_bf16_4d(m, base, dims, strides, g.G, g.N);
}

int attention_bwd_tc(const AttnGeom& g, const bf16* qkv, const float* lse, const bf16* d_out, bf16* d_qkv, Drop drop, cudaStream_t st) {
    MSST_REQUIRE(attention_bwd_tc_supported(g), "attention_bwd_tc: needs packed short sequences (N <= 64)");
    const int64_t n_tiles = (g.groups + 1) / 2;
    static PerDeviceOnce once;
    if (once.first()) MSST_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcbSmem));
    const int64_t I = (int64_t)g.H * 64;
    CUtensorMap t_qkv, t_do, t_dqkv;
    const int nbox = (g.gpb == 0 && g.G * g.N == 64) ? 1 : 2;
    if (int rc = tcb_tmap(&t_qkv, g, qkv, 3 * I, nbox)) return rc;
    if (int rc = tcb_tmap(&t_do, g, d_out, I, nbox)) return rc;
    if (int rc = tcb_tmap(&t_dqkv, g, d_qkv, 3 * I, nbox)) return rc;
    const int grid = (int)(n_tiles < kNumSMs ? n_tiles : kNumSMs);
    static int pf_dist = -1;
    if (pf_dist < 0) { const char* e = getenv("MSST_ATTN_PF"); pf_dist = e ? atoi(e) : TCB_PREFETCH; }
    static long long* dbg = nullptr;
    static int dbg_on = -1;
    if (dbg_on < 0) { const char* e = getenv("MSST_ATTN_DBG"); dbg_on = e ? atoi(e) : 0; if (dbg_on) { cudaMalloc(&dbg, 32 * 16 * 8); } }
    if (dbg_on) cudaMemsetAsync(dbg, 0, 32 * 16 * 8, st);
    attn_bwd_tc_kernel<<<grid, TCB_THREADS, kTcbSmem, st>>>(t_qkv, t_do, t_dqkv, g, lse, drop, n_tiles, nbox, pf_dist, dbg);
    MSST_LAUNCH_CHECK();
    if (dbg_on) {
        static int printed = 0;
        if (++printed == dbg_on) {
            long long h[32 * 16];
            cudaStreamSynchronize(st);
            cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
            long long t0 = h[0];
            printf("item: load_issue p1_issue sfull_seen ld_done pfull_arr p2_issue ofull_seen epi_done store_read (clks rel. to first load)\n");
            for (int i = 0; i < 24; ++i) { printf("%2d:", i); for (int k = 0; k < 9; ++k) printf(" %7lld", h[i * 16 + k] ? h[i * 16 + k] - t0 : -1); printf("\n"); }
            fflush(stdout);
        }
    }
    return MSST_OK;
}

}  // namespace msst
