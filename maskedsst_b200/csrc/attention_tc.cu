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

namespace msst {
using namespace ptx;
typedef __nv_bfloat16 bf16;
int make_tmap_bf16(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);   // gemm_bf16.cu

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

}  // namespace msst
