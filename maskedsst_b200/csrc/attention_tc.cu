// Kernel (3), bf16 mode, packed short sequences (N <= 64): attention forward AND backward on the 5th-generation tensor cores
// (tcgen05 + TMEM), operands by TMA, results by TMA store.
//
// One persistent CTA per SM works on 128 slots (two 64-slot groups of attn_geom.cuh) x one head at a time.  Warp roles:
//   warp 16  TMA producer (one elected thread; also allocates TMEM): one box per operand per slot group -- 2-D boxes when the
//            tile is 128 consecutive rows (spatial stack), 4-D boxes over the (col, s, pos, blk) view for strided sequences
//            (spectral stack) -- plus an L2 prefetch of the next item's boxes
//   warp 17  MMA issuer (one elected thread): a small state machine that issues whichever contraction has its inputs ready
//   warp 18  TMA store of the staged result tiles (nobody else ever waits for a store to drain)
//   warps 0-7 / 8-15  two softmax + epilogue groups that ping-pong over items; thread (r, half) owns row r and 32 of the 64
//            columns of the row's diagonal block (tcgen05.ld 32x32b: thread = TMEM lane), row statistics are exchanged between
//            the two halves through shared memory
// Only the two diagonal 64x64 blocks of S (and dP) are meaningful (block-diagonal mask): P~ / dS are kept as compact tiles
// whose off-diagonal parts share one block of zeros.  See the section comments for the per-direction data flow.
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
constexpr uint32_t TC_TILE = TC_ROWS * 128;   // bytes of one [128 rows][64 bf16] tile (16 KB)

__device__ __forceinline__ uint32_t pack_bf(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// dropout pair indexing / hash identical to attention_bf16.cu (every kernel regenerates the same mask)
__device__ __forceinline__ uint64_t tile_pair_base_tc(const AttnGeom& g, int64_t group, int h) {
    return ((uint64_t)group * g.H + h) * (uint64_t)(TS * TS / 2);
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
           g.n_seq * g.N < (int64_t)2147483647 - 256 &&   // TMA coordinates are 32-bit
          
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

// =========================================================================================================
// Forward on tcgen05 / TMEM, same tile / warp-role structure as the backward above.  Per (128-slot tile, head) item:
//   S = Q K^T -> TMEM [128 x 128];  softmax threads (row, half): row max and row sum exchanged between the two halves through
//   smem, P~ = dropout(exp2(S*scale*log2e - max)) (unnormalised, bf16) -> smem;  O = P~ V -> TMEM [128 x 64] (V as MN-major B);
//   epilogue: O / l -> bf16 -> staging tile -> TMA store, lse -> global.
// 3 operand stages (Q, K, V: 48 KB each), one P~ buffer and one S/O TMEM slot per softmax group, one staging tile.
// HBM-bound: 3 * 128 B in + 128 B out (+ 4 B lse) per (slot, head).
// =========================================================================================================
struct alignas(8) TcfBars {
    uint64_t full[3], kv_empty[3], s_full[2], p_full[2], o_full[2], tmem_free[2], stg_free[2], stg_full[2];
    uint32_t tmem_base;
};
constexpr int TCF_STAGES = 3;
// smem: [3 stages][Q, K, V][16 KB] | P~ per group (2 x 24 KB) | O staging (16 KB) | row max / row sum exchange | barriers
constexpr size_t kTcfSmem = TCF_STAGES * 3 * TC_TILE + 2 * TCB_PD + TC_TILE + 2 * 2 * 2 * 128 * sizeof(float) + sizeof(TcfBars);

__global__ void __launch_bounds__(TCB_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_out, AttnGeom g,
                   float* __restrict__ lse, Drop drop, int64_t n_tiles, int nbox, int pf_dist) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (smem_u32(smem) & 1023u) __trap();
    uint8_t* op_s = smem;                                   // [3][3][16 KB]
    uint8_t* p_s = smem + TCF_STAGES * 3 * TC_TILE;         // P~ of group 0, group 1 (24 KB each, see TCB_PD)
    uint8_t* stg_s = p_s + 2 * TCB_PD;                      // O staging for the TMA store
    float* xch = reinterpret_cast<float*>(stg_s + TC_TILE); // [group][max | sum][half][128]
    TcfBars* bars = reinterpret_cast<TcfBars*>(xch + 2 * 2 * 2 * 128);
    const int warp = threadIdx.x >> 5;
    const int I = g.H * 64;

    if (warp == 17 && elect_one()) {
        for (int s = 0; s < TCF_STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->kv_empty[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars->s_full[s], 1); mbar_init(&bars->p_full[s], 256); mbar_init(&bars->o_full[s], 1);
            mbar_init(&bars->tmem_free[s], 256); mbar_init(&bars->stg_free[s], 1); mbar_init(&bars->stg_full[s], 256);
        }
        fence_barrier_init();
    }
    if (warp == 16) {
        tmem_alloc(&bars->tmem_base, 512);
        if (elect_one()) { prefetch_tmap(&tma_qkv); prefetch_tmap(&tma_out); }
    }
    // zero once: operand rows a group's box does not cover and the shared zero blocks of the P~ tiles
    for (uint32_t i = threadIdx.x; i < (TCF_STAGES * 3 * TC_TILE + 2 * TCB_PD) / 16; i += TCB_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0), idesc_o = make_idesc_bf16(128, 64, 0, 1);

    const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int64_t n_items = my_tiles * g.H;
    const uint32_t box_bytes = (uint32_t)(g.G * g.N) * 128u * (nbox == 1 ? 2u : 1u);

    if (warp == 16) {
        // ===== TMA producer =====
        if (elect_one()) {
            int stg = 0; uint32_t ph = 0;
            for (int64_t it = 0; it < n_items; ++it) {
                const int64_t tile = blockIdx.x + (it / g.H) * gridDim.x; const int h = (int)(it % g.H);
                mbar_wait(&bars->kv_empty[stg], ph ^ 1);
                uint8_t* base = op_s + (size_t)stg * 3 * TC_TILE;
                mbar_arrive_expect_tx(&bars->full[stg], 3 * nbox * box_bytes);
                if (nbox == 1) {
                    const int row0 = (int)(tile * TC_ROWS);
                    tma_load_2d(base, &tma_qkv, &bars->full[stg], h * 64, row0);
                    tma_load_2d(base + TC_TILE, &tma_qkv, &bars->full[stg], I + h * 64, row0);
                    tma_load_2d(base + 2 * TC_TILE, &tma_qkv, &bars->full[stg], 2 * I + h * 64, row0);
                } else
                for (int gi = 0; gi < nbox; ++gi) {
                    int c1, c2, c3;
                    tcb_group_coords(g, tile * 2 + gi, c1, c2, c3);
                    uint8_t* dst = base + gi * 64 * 128;
                    tma_load_4d(dst, &tma_qkv, &bars->full[stg], h * 64, c1, c2, c3);
                    tma_load_4d(dst + TC_TILE, &tma_qkv, &bars->full[stg], I + h * 64, c1, c2, c3);
                    tma_load_4d(dst + 2 * TC_TILE, &tma_qkv, &bars->full[stg], 2 * I + h * 64, c1, c2, c3);
                }
                if (pf_dist > 0 && it + pf_dist < n_items) {
                    const int64_t it2 = it + pf_dist;
                    const int64_t tile2 = blockIdx.x + (it2 / g.H) * gridDim.x; const int h2 = (int)(it2 % g.H);
                    if (nbox == 1) {
                        const int row2 = (int)(tile2 * TC_ROWS);
                        tma_prefetch_l2_2d(&tma_qkv, h2 * 64, row2);
                        tma_prefetch_l2_2d(&tma_qkv, I + h2 * 64, row2);
                        tma_prefetch_l2_2d(&tma_qkv, 2 * I + h2 * 64, row2);
                    } else
                    for (int gi = 0; gi < nbox; ++gi) {
                        int c1, c2, c3;
                        tcb_group_coords(g, tile2 * 2 + gi, c1, c2, c3);
                        tma_prefetch_l2_4d(&tma_qkv, h2 * 64, c1, c2, c3);
                        tma_prefetch_l2_4d(&tma_qkv, I + h2 * 64, c1, c2, c3);
                        tma_prefetch_l2_4d(&tma_qkv, 2 * I + h2 * 64, c1, c2, c3);
                    }
                }
                if (++stg == TCF_STAGES) { stg = 0; ph ^= 1; }
            }
        }
    } else if (warp == 17) {
        // ===== MMA issuer: S of item next_s / O of item next_p, whichever has its inputs ready; S runs at most one item ahead =====
        if (elect_one()) {
            int64_t next_s = 0, next_p = 0;
            int stg_s1 = 0, stg_p = 0; uint32_t ph_s1 = 0;
            while (next_p < n_items) {
                if (next_s < n_items && next_s <= next_p + 1) {
                    const int slot = (int)(next_s & 1); const uint32_t phg = (uint32_t)(next_s >> 1) & 1;
                    if (mbar_try_wait(&bars->tmem_free[slot], phg ^ 1) && mbar_try_wait(&bars->full[stg_s1], ph_s1)) {
                        tc_fence_after();
                        const uint32_t qb = smem_u32(op_s + (size_t)stg_s1 * 3 * TC_TILE);
                        const uint64_t dq = make_smem_desc(qb, 16, 1024), dk = make_smem_desc(qb + TC_TILE, 16, 1024);
                        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + slot * 256, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k != 0);
                        umma_commit(&bars->s_full[slot]);
                        ++next_s;
                        if (++stg_s1 == TCF_STAGES) { stg_s1 = 0; ph_s1 ^= 1; }
                    }
                }
                if (next_p < next_s) {
                    const int slot = (int)(next_p & 1);
                    if (mbar_try_wait(&bars->p_full[slot], (uint32_t)(next_p >> 1) & 1)) {
                        fence_proxy_async();
                        tc_fence_after();
                        const uint32_t vb = smem_u32(op_s + (size_t)stg_p * 3 * TC_TILE + 2 * TC_TILE);
                        const uint32_t pb = smem_u32(p_s + (size_t)slot * TCB_PD);
                        const uint64_t b_v = make_smem_desc(vb, TC_TILE, 1024);
                        for (int k = 0; k < 8; ++k) {   // O = P~ V : reduction over the 128 keys (2 chunks x 4 steps), V as MN-major B
                            const uint64_t a = make_smem_desc(pb + (k >> 2) * TCB_CHUNK, 16, 1024) + (uint64_t)((k & 3) * 2);
                            umma_bf16(tmem_base + slot * 256 + 128, a, b_v + (uint64_t)(k * 128), idesc_o, k != 0);
                        }
                        umma_commit(&bars->o_full[slot]);
                        umma_commit(&bars->kv_empty[stg_p]);
                        ++next_p;
                        if (++stg_p == TCF_STAGES) stg_p = 0;
                    }
                }
            }
        }
    } else if (warp == 18) {
        // ===== TMA store of the staged O tile =====
        if (elect_one()) {
            for (int64_t it = 0; it < n_items; ++it) {
                const int64_t tile = blockIdx.x + (it / g.H) * gridDim.x; const int h = (int)(it % g.H);
                mbar_wait(&bars->stg_full[it & 1], (uint32_t)(it >> 1) & 1);
                if (nbox == 1) tma_store_2d(&tma_out, stg_s, h * 64, (int)(tile * TC_ROWS));
                else
                for (int gi = 0; gi < nbox; ++gi) {
                    int c1, c2, c3;
                    tcb_group_coords(g, tile * 2 + gi, c1, c2, c3);
                    tma_store_4d(&tma_out, stg_s + gi * 64 * 128, h * 64, c1, c2, c3);
                }
                tma_store_commit();
                tma_store_wait_read();
                mbar_arrive(&bars->stg_free[it & 1]);
            }
        }
    } else if (warp < 16) {
        // ===== two softmax + epilogue groups of 8 warps; thread (r, half) owns row r and 32 of its 64 columns =====
        const int grp = warp >> 3, wl = warp & 7, half = wl >> 2;
        const int r = (wl & 3) * 32 + (threadIdx.x & 31);
        const int blk = r >> 6;
        const int c0 = half * 32;
        const float sl2 = g.scale * 1.4426950408889634f;
        const uint32_t lane_base = (uint32_t)((wl & 3) * 32) << 16;
        const uint32_t swz = (uint32_t)(r & 7);
        const bool full_blocks = g.N == 64 && g.groups % 2 == 0;
        const uint64_t seed = drop.seed + (drop.seed_dev ? __ldg(drop.seed_dev) : 0ull);
        const uint32_t t16 = drop.thresh >> 16;
        float* my_max = xch + ((grp * 2 + 0) * 2 + half) * 128 + r;
        float* my_sum = xch + ((grp * 2 + 1) * 2 + half) * 128 + r;
        const float* other_max = xch + ((grp * 2 + 0) * 2 + (half ^ 1)) * 128 + r;
        const float* other_sum = xch + ((grp * 2 + 1) * 2 + (half ^ 1)) * 128 + r;
        uint8_t* prow = p_s + (size_t)grp * TCB_PD + (size_t)blk * TCB_CHUNK + r * 128;
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
                vmask = ok ? (full_blocks ? 0xFFFFFFFFu : vmask_geom) : 0u;
            }
            const uint64_t hidx = tile_pair_base_tc(g, tile * 2 + blk, h) + (uint64_t)((r & 63) * 32 + (c0 >> 1));
            const uint32_t hash_lo = (uint32_t)hidx + (uint32_t)(seed >> 32);
            const uint32_t hash_hi = ((uint32_t)(hidx >> 32) * 0x85EBCA77u) ^ (uint32_t)seed ^ (drop.site * 0xC2B2AE3Du);
            mbar_wait(&bars->s_full[grp], ph);
            tc_fence_after();
            float p[32];
            {
                uint32_t a[32];
                tmem_ld_32x32(tmem_base + lane_base + grp * 256 + blk * 64 + c0, a);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) p[j] = __uint_as_float(a[j]) * sl2;
            }
            float mx = -INFINITY;
            if (full_blocks) {
#pragma unroll
                for (int j = 0; j < 32; ++j) mx = fmaxf(mx, p[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) { p[j] = (vmask >> j) & 1u ? p[j] : -INFINITY; mx = fmaxf(mx, p[j]); }
            }
            *my_max = mx;
            asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory");
            mx = fmaxf(mx, *other_max);
            const float sub = mx == -INFINITY ? 0.f : mx;
            float l = 0.f;
            uint32_t pfp[16];
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                float p0 = ex2_approx(p[2 * jj] - sub), p1 = ex2_approx(p[2 * jj + 1] - sub);
                l += p0 + p1;
                if (drop.on()) {
                    uint32_t x = (hash_lo + (uint32_t)jj) * 0x9E3779B1u ^ hash_hi;
                    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
                    p0 *= (x & 0xFFFFu) >= t16 ? drop.scale : 0.f;
                    p1 *= (x >> 16) >= t16 ? drop.scale : 0.f;
                }
                pfp[jj] = pack_bf(p0, p1);
            }
            *my_sum = l;
            // this group's P~ buffer was last read by the O MMAs of item it - 2, whose completion this group awaited (o_full)
#pragma unroll
            for (int pc = 0; pc < 4; ++pc)
                *reinterpret_cast<uint4*>(prow + (((uint32_t)(half * 4 + pc) ^ swz) << 4)) = make_uint4(pfp[pc * 4], pfp[pc * 4 + 1], pfp[pc * 4 + 2], pfp[pc * 4 + 3]);
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&bars->p_full[grp]);
            asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory");
            l += *other_sum;
            const float inv = l > 0.f ? 1.f / l : 0.f;
            if (half == 0 && grow >= 0) lse[grow * g.H + h] = (sub + log2f(l)) * 0.6931471805599453f;
            // ---- epilogue: O row / l -> bf16 -> staging -> TMA store ----
            mbar_wait(&bars->o_full[grp], ph);
            if (it > 0) mbar_wait(&bars->stg_free[(it - 1) & 1], (uint32_t)((it - 1) >> 1) & 1);
            tc_fence_after();
            {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + lane_base + grp * 256 + 128 + c0, v);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&bars->tmem_free[grp]);
                uint8_t* trow = stg_s + r * 128;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<uint4*>(trow + (((uint32_t)(half * 4 + c) ^ swz) << 4)) =
                        make_uint4(pack_bf(__uint_as_float(v[c * 8]) * inv, __uint_as_float(v[c * 8 + 1]) * inv), pack_bf(__uint_as_float(v[c * 8 + 2]) * inv, __uint_as_float(v[c * 8 + 3]) * inv),
                                   pack_bf(__uint_as_float(v[c * 8 + 4]) * inv, __uint_as_float(v[c * 8 + 5]) * inv), pack_bf(__uint_as_float(v[c * 8 + 6]) * inv, __uint_as_float(v[c * 8 + 7]) * inv));
            }
            fence_proxy_async();
            mbar_arrive(&bars->stg_full[it & 1]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int attention_fwd_tc(const AttnGeom& g, const bf16* qkv, bf16* out, float* lse, Drop drop, cudaStream_t st) {
    MSST_REQUIRE(attention_bwd_tc_supported(g), "attention_fwd_tc: needs packed short sequences (N <= 64)");
    const int64_t n_tiles = (g.groups + 1) / 2;
    static PerDeviceOnce once;
    if (once.first()) MSST_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcfSmem));
    const int64_t I = (int64_t)g.H * 64;
    const int nbox = (g.gpb == 0 && g.G * g.N == 64) ? 1 : 2;
    CUtensorMap t_qkv, t_out;
    if (int rc = tcb_tmap(&t_qkv, g, qkv, 3 * I, nbox)) return rc;
    if (int rc = tcb_tmap(&t_out, g, out, I, nbox)) return rc;
    // no L2 prefetch in the forward: its three operand stages already keep enough loads in flight (measured: spectral 280 vs 294 us)
    static int pf_dist = -1;
    if (pf_dist < 0) { const char* e = getenv("MSST_ATTN_PF_FWD"); pf_dist = e ? atoi(e) : 0; }
    const int grid = (int)(n_tiles < kNumSMs ? n_tiles : kNumSMs);
    attn_fwd_tc_kernel<<<grid, TCB_THREADS, kTcfSmem, st>>>(t_qkv, t_out, g, lse, drop, n_tiles, nbox, pf_dist);
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}

// =========================================================================================================
// Long sequences (N > 64) forward on tcgen05 / TMEM: one work item = (sequence, head, 128-query tile), key tiles of 128
// streamed through a 3-slot TMA ring.  Two passes over the keys instead of an online-softmax rescale of the TMEM accumulator:
//   pass 1: S_j = Q K_j^T -> row max (all 16 softmax warps work on one S tile: thread = (row, 32 of its 128 columns))
//   pass 2: S_j again, P~_j = dropout(exp2(S_j - max)) -> smem (double-buffered), O += P~_j V_j accumulates in TMEM over ALL
//           key tiles (no correction step), row sums in registers; epilogue O / l -> staging -> TMA store.
// The second Q K^T costs 1/3 more tensor work, which is idle anyway: the kernel is bound by the exponentials (MUFU).
// Dropout pair indices are those of the mma.sync long kernels (64 x 64 tile coordinates), so the regular backward applies.
// =========================================================================================================
struct alignas(8) TclBars {
    uint64_t q_full[2], q_empty[2], kv_full[3], kv_empty[3], s_full[2], s_free[2], p_full[2], p_free[2], o_full[2], o_free[2],
             stg_full[2], stg_free[2];
    uint32_t tmem_base;
};
// smem: Q [2][16 KB] | ring [3][K, V][16 KB] | P~ [2][2 chunks][16 KB] | O staging [16 KB] | row statistics exchange [4][128] | barriers
constexpr size_t kTclSmem = 2 * TC_TILE + 3 * 2 * TC_TILE + 2 * 2 * TC_TILE + TC_TILE + 4 * 128 * sizeof(float) + sizeof(TclBars);

__device__ __forceinline__ void tcl_item(const AttnGeom& g, int64_t item, int qtiles, int64_t& seq, int& h, int& qt) {
    qt = (int)(item % qtiles); h = (int)((item / qtiles) % g.H); seq = item / ((int64_t)qtiles * g.H);
}
// TMA coordinates of rows [pos0, pos0 + 128) of sequence seq (see tcl_tmap)
__device__ __forceinline__ void tcl_coords(const AttnGeom& g, int64_t seq, int pos0, int& c1, int& c2, int& c3) {
    if (g.inner == 1) { c1 = pos0; c2 = (int)seq; c3 = 0; }
    else { c1 = (int)(seq % g.inner); c2 = pos0; c3 = (int)(seq / g.inner); }
}

__global__ void __launch_bounds__(TCB_THREADS, 1)
attn_fwd_tc_long_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_out, AttnGeom g,
                        float* __restrict__ lse, Drop drop, int64_t n_items, int qtiles, int nk) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (smem_u32(smem) & 1023u) __trap();
    uint8_t* q_s = smem;                                    // [2][16 KB]
    uint8_t* kv_s = q_s + 2 * TC_TILE;                      // [3][K | V][16 KB]
    uint8_t* p_s = kv_s + 3 * 2 * TC_TILE;                  // [2][keys 0-63 | keys 64-127][16 KB]
    uint8_t* stg_s = p_s + 2 * 2 * TC_TILE;
    float* xch = reinterpret_cast<float*>(stg_s + TC_TILE); // [4 column quarters][128 rows]
    TclBars* bars = reinterpret_cast<TclBars*>(xch + 4 * 128);
    const int warp = threadIdx.x >> 5;
    const int I = g.H * 64;

    if (warp == 17 && elect_one()) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars->q_full[s], 1); mbar_init(&bars->q_empty[s], 1); mbar_init(&bars->s_full[s], 1); mbar_init(&bars->s_free[s], 512);
            mbar_init(&bars->p_full[s], 512); mbar_init(&bars->p_free[s], 1); mbar_init(&bars->o_full[s], 1); mbar_init(&bars->o_free[s], 512);
            mbar_init(&bars->stg_full[s], 512); mbar_init(&bars->stg_free[s], 1);
        }
        for (int s = 0; s < 3; ++s) { mbar_init(&bars->kv_full[s], 1); mbar_init(&bars->kv_empty[s], 1); }
        fence_barrier_init();
    }
    if (warp == 16) {
        tmem_alloc(&bars->tmem_base, 512);
        if (elect_one()) { prefetch_tmap(&tma_qkv); prefetch_tmap(&tma_out); }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0), idesc_o = make_idesc_bf16(128, 64, 0, 1);
    const int64_t my_items = blockIdx.x < n_items ? (n_items - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    if (warp == 16) {
        // ===== TMA producer: Q once per item, then K_j (pass 1), then K_j + V_j (pass 2) through the ring =====
        if (elect_one()) {
            int ring = 0; uint32_t rph = 0;
            for (int64_t n = 0; n < my_items; ++n) {
                int64_t seq; int h, qt, c1, c2, c3;
                tcl_item(g, blockIdx.x + n * gridDim.x, qtiles, seq, h, qt);
                const int qs = (int)(n & 1);
                mbar_wait(&bars->q_empty[qs], (uint32_t)((n >> 1) & 1) ^ 1);
                tcl_coords(g, seq, qt * TC_ROWS, c1, c2, c3);
                mbar_arrive_expect_tx(&bars->q_full[qs], TC_TILE);
                tma_load_4d(q_s + (size_t)qs * TC_TILE, &tma_qkv, &bars->q_full[qs], h * 64, c1, c2, c3);
                for (int pass = 0; pass < 2; ++pass)
                    for (int j = 0; j < nk; ++j) {
                        mbar_wait(&bars->kv_empty[ring], rph ^ 1);
                        tcl_coords(g, seq, j * TC_ROWS, c1, c2, c3);
                        uint8_t* dst = kv_s + (size_t)ring * 2 * TC_TILE;
                        mbar_arrive_expect_tx(&bars->kv_full[ring], pass ? 2 * TC_TILE : TC_TILE);
                        tma_load_4d(dst, &tma_qkv, &bars->kv_full[ring], I + h * 64, c1, c2, c3);
                        if (pass) tma_load_4d(dst + TC_TILE, &tma_qkv, &bars->kv_full[ring], 2 * I + h * 64, c1, c2, c3);
                        if (++ring == 3) { ring = 0; rph ^= 1; }
                    }
            }
        }
    } else if (warp == 17) {
        // ===== MMA issuer.  Global step counter t over (item, pass, j); S of step t+1 is issued before the P~V of step t =====
        if (elect_one()) {
            const int64_t steps_per_item = 2 * (int64_t)nk, total = my_items * steps_per_item;
            int64_t ts = 0, tp = 0;          // next S step / next step whose pass-2 work (P~ V) is to be issued or skipped
            int ring_s = 0; uint32_t rph_s = 0; int ring_p = 0;
            int64_t npv = 0;                 // pass-2 steps issued so far (P~ buffer = npv & 1)
            // (item, step within the item) of ts / tp, advanced incrementally (a 64-bit division per poll of this loop is ~100 instructions
            // of the single issuing thread)
            const int spi = (int)steps_per_item;
            int64_t n_s = 0, n_p = 0; int rem_s = 0, rem_p = 0;
            while (tp < total) {
                if (ts < total && ts <= tp + 1) {        // two S slots (a third measured slower: the softmax warps are the bottleneck)
                    const int64_t n = n_s; const int rem = rem_s;
                    const int slot = (int)(ts & 1);
                    bool ok = mbar_try_wait(&bars->s_free[slot], (uint32_t)((ts >> 1) & 1) ^ 1) && mbar_try_wait(&bars->kv_full[ring_s], rph_s);
                    if (ok && rem == 0) ok = mbar_try_wait(&bars->q_full[n & 1], (uint32_t)(n >> 1) & 1);
                    if (ok) {
                        tc_fence_after();
                        const uint64_t dq = make_smem_desc(smem_u32(q_s + (size_t)(n & 1) * TC_TILE), 16, 1024);
                        const uint64_t dk = make_smem_desc(smem_u32(kv_s + (size_t)ring_s * 2 * TC_TILE), 16, 1024);
                        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + slot * 128, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k != 0);
                        umma_commit(&bars->s_full[slot]);
                        if (rem < nk) umma_commit(&bars->kv_empty[ring_s]);      // pass 1: K is not needed again
                        ++ts;
                        if (++rem_s == spi) { rem_s = 0; ++n_s; }
                        if (++ring_s == 3) { ring_s = 0; rph_s ^= 1; }
                    }
                }
                if (tp < ts) {
                    const int64_t n = n_p; const int rem = rem_p;
                    if (rem < nk) { ++tp; if (++rem_p == spi) { rem_p = 0; ++n_p; } if (++ring_p == 3) ring_p = 0; }       // pass-1 step: nothing to issue
                    else {
                        const int pb = (int)(npv & 1), os = (int)(n & 1), j = rem - nk;
                        bool ok = mbar_try_wait(&bars->p_full[pb], (uint32_t)(npv >> 1) & 1);
                        if (ok && j == 0) ok = mbar_try_wait(&bars->o_free[os], (uint32_t)((n >> 1) & 1) ^ 1);
                        if (ok) {
                            fence_proxy_async();
                            tc_fence_after();
                            const uint32_t pbase = smem_u32(p_s + (size_t)pb * 2 * TC_TILE);
                            const uint64_t b_v = make_smem_desc(smem_u32(kv_s + (size_t)ring_p * 2 * TC_TILE + TC_TILE), TC_TILE, 1024);
                            for (int k = 0; k < 8; ++k) {
                                const uint64_t a = make_smem_desc(pbase + (k >> 2) * TC_TILE, 16, 1024) + (uint64_t)((k & 3) * 2);
                                umma_bf16(tmem_base + 256 + os * 64, a, b_v + (uint64_t)(k * 128), idesc_o, (j | k) != 0);
                            }
                            umma_commit(&bars->kv_empty[ring_p]);
                            umma_commit(&bars->p_free[pb]);
                            if (j == nk - 1) { umma_commit(&bars->o_full[os]); umma_commit(&bars->q_empty[os]); }
                            ++npv; ++tp;
                            if (++rem_p == spi) { rem_p = 0; ++n_p; }
                            if (++ring_p == 3) ring_p = 0;
                        }
                    }
                }
            }
        }
    } else if (warp == 18) {
        // ===== TMA store of the staged O tile =====
        if (elect_one()) {
            for (int64_t n = 0; n < my_items; ++n) {
                int64_t seq; int h, qt, c1, c2, c3;
                tcl_item(g, blockIdx.x + n * gridDim.x, qtiles, seq, h, qt);
                tcl_coords(g, seq, qt * TC_ROWS, c1, c2, c3);
                mbar_wait(&bars->stg_full[n & 1], (uint32_t)(n >> 1) & 1);
                tma_store_4d(&tma_out, stg_s, h * 64, c1, c2, c3);
                tma_store_commit();
                tma_store_wait_read();
                mbar_arrive(&bars->stg_free[n & 1]);
            }
        }
    } else if (warp < 16) {
        // ===== 16 softmax warps on ONE S tile at a time: thread = (row r, column quarter cq) =====
        const int cq = warp >> 2;
        const int r = (warp & 3) * 32 + (threadIdx.x & 31);
        const float sl2 = g.scale * 1.4426950408889634f;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t swz = (uint32_t)(r & 7);
        const uint64_t seed = drop.seed + (drop.seed_dev ? __ldg(drop.seed_dev) : 0ull);
        const uint32_t t16 = drop.thresh >> 16;
        int64_t t = 0, npv = 0;
        for (int64_t n = 0; n < my_items; ++n) {
            int64_t seq; int h, qt;
            tcl_item(g, blockIdx.x + n * gridDim.x, qtiles, seq, h, qt);
            const int pos = qt * TC_ROWS + r;
            const bool row_ok = pos < g.N;
            float mx = -INFINITY, l = 0.f, sub = 0.f;
            for (int pass = 0; pass < 2; ++pass) {
                for (int j = 0; j < nk; ++j, ++t) {
                    const int slot = (int)(t & 1);
                    mbar_wait(&bars->s_full[slot], (uint32_t)(t >> 1) & 1);
                    tc_fence_after();
                    uint32_t a[32];
                    tmem_ld_32x32(tmem_base + lane_base + slot * 128 + cq * 32, a);
                    tmem_ld_wait();
                    tc_fence_before();
                    mbar_arrive(&bars->s_free[slot]);
                    const int nvalid = g.N - j * TC_ROWS - cq * 32;            // this thread's columns 0 .. nvalid-1 are real keys
                    if (pass == 0) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) if (c < nvalid) mx = fmaxf(mx, __uint_as_float(a[c]) * sl2);
                    } else {
                        uint32_t pk[16];
                        // 64 x 64 tile coordinates of the mma.sync kernels: qt64 = 2*qt + r/64, kt64 = 2*j + cq/2, pair = (cq & 1)*16 + jj
                        const uint64_t hidx = ((((uint64_t)seq * g.H + h) * g.tiles + (uint64_t)(2 * qt + (r >> 6))) * g.tiles + (uint64_t)(2 * j + (cq >> 1))) *
                                                  (uint64_t)(TS * TS / 2) + (uint64_t)((r & 63) * 32 + (cq & 1) * 16);
                        const uint32_t hash_lo = (uint32_t)hidx + (uint32_t)(seed >> 32);
                        const uint32_t hash_hi = ((uint32_t)(hidx >> 32) * 0x85EBCA77u) ^ (uint32_t)seed ^ (drop.site * 0xC2B2AE3Du);
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) {
                            float p0 = 2 * jj < nvalid ? ex2_approx(fmaf(__uint_as_float(a[2 * jj]), sl2, -sub)) : 0.f;
                            float p1 = 2 * jj + 1 < nvalid ? ex2_approx(fmaf(__uint_as_float(a[2 * jj + 1]), sl2, -sub)) : 0.f;
                            l += p0 + p1;
                            if (drop.on()) {
                                uint32_t x = (hash_lo + (uint32_t)jj) * 0x9E3779B1u ^ hash_hi;
                                x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
                                p0 *= (x & 0xFFFFu) >= t16 ? drop.scale : 0.f;
                                p1 *= (x >> 16) >= t16 ? drop.scale : 0.f;
                            }
                            pk[jj] = pack_bf(p0, p1);
                        }
                        const int pb = (int)(npv & 1);
                        mbar_wait(&bars->p_free[pb], (uint32_t)((npv >> 1) & 1) ^ 1);   // the P~ V of two steps ago has read this buffer
                        uint8_t* prow = p_s + (size_t)pb * 2 * TC_TILE + (size_t)(cq >> 1) * TC_TILE + r * 128;
#pragma unroll
                        for (int pc = 0; pc < 4; ++pc)
                            *reinterpret_cast<uint4*>(prow + (((uint32_t)((cq & 1) * 4 + pc) ^ swz) << 4)) = make_uint4(pk[pc * 4], pk[pc * 4 + 1], pk[pc * 4 + 2], pk[pc * 4 + 3]);
                        fence_proxy_async();
                        tc_fence_before();
                        mbar_arrive(&bars->p_full[pb]);
                        ++npv;
                    }
                }
                // combine the four column quarters of every row: max after pass 1, sum after pass 2
                xch[cq * 128 + r] = pass == 0 ? mx : l;
                asm volatile("bar.sync 1, 512;" ::: "memory");
                const float v0 = xch[r], v1 = xch[128 + r], v2 = xch[256 + r], v3 = xch[384 + r];
                asm volatile("bar.sync 1, 512;" ::: "memory");
                if (pass == 0) { mx = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3)); sub = mx == -INFINITY ? 0.f : mx; }
                else l = (v0 + v1) + (v2 + v3);
            }
            // ---- epilogue: thread (r, cq) owns O columns cq*16 .. +15 ----
            const int os = (int)(n & 1);
            const float inv = l > 0.f ? 1.f / l : 0.f;
            if (cq == 0 && row_ok) lse[row_of(g, seq, pos) * g.H + h] = (sub + log2f(l)) * 0.6931471805599453f;
            mbar_wait(&bars->o_full[os], (uint32_t)(n >> 1) & 1);
            if (n > 0) mbar_wait(&bars->stg_free[(n - 1) & 1], (uint32_t)((n - 1) >> 1) & 1);
            tc_fence_after();
            uint32_t v[16];
            tmem_ld_32x16(tmem_base + lane_base + 256 + os * 64 + cq * 16, v);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&bars->o_free[os]);
            uint8_t* trow = stg_s + r * 128;
#pragma unroll
            for (int c = 0; c < 2; ++c)
                *reinterpret_cast<uint4*>(trow + (((uint32_t)(cq * 2 + c) ^ swz) << 4)) =
                    make_uint4(pack_bf(__uint_as_float(v[c * 8]) * inv, __uint_as_float(v[c * 8 + 1]) * inv), pack_bf(__uint_as_float(v[c * 8 + 2]) * inv, __uint_as_float(v[c * 8 + 3]) * inv),
                               pack_bf(__uint_as_float(v[c * 8 + 4]) * inv, __uint_as_float(v[c * 8 + 5]) * inv), pack_bf(__uint_as_float(v[c * 8 + 6]) * inv, __uint_as_float(v[c * 8 + 7]) * inv));
            fence_proxy_async();
            mbar_arrive(&bars->stg_full[n & 1]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// 4-D view whose box is 128 consecutive positions of ONE sequence (rows beyond the sequence end: zero on load, clipped on store)
static int tcl_tmap(CUtensorMap* m, const AttnGeom& g, const bf16* base, int64_t cols) {
    if (g.inner == 1) {
        const int64_t dims[4] = {cols, g.N, g.n_seq, 1}, strides[3] = {cols, (int64_t)g.N * cols, g.n_seq * g.N * cols};
        return make_tmap_bf16_4d(m, base, dims, strides, TC_ROWS, 1);
    }
    const int64_t dims[4] = {cols, g.inner, g.N, g.n_seq / g.inner};
    const int64_t strides[3] = {cols, (int64_t)g.inner * cols, (int64_t)g.N * g.inner * cols};
    return make_tmap_bf16_4d(m, base, dims, strides, 1, TC_ROWS);
}

bool attention_fwd_tc_long_supported(const AttnGeom& g) {
    return g.tiles > 1 && g.n_seq > 0 && g.n_seq < (int64_t)2147483647 && (int64_t)g.N < (int64_t)2147483647 - 256;
}

int attention_fwd_tc_long(const AttnGeom& g, const bf16* qkv, bf16* out, float* lse, Drop drop, cudaStream_t st) {
    MSST_REQUIRE(attention_fwd_tc_long_supported(g), "attention_fwd_tc_long: needs N > 64");
    static PerDeviceOnce once;
    if (once.first()) MSST_CUDA(cudaFuncSetAttribute(attn_fwd_tc_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTclSmem));
    const int64_t I = (int64_t)g.H * 64;
    const int qtiles = (g.N + TC_ROWS - 1) / TC_ROWS, nk = qtiles;
    const int64_t n_items = g.n_seq * g.H * qtiles;
    CUtensorMap t_qkv, t_out;
    if (int rc = tcl_tmap(&t_qkv, g, qkv, 3 * I)) return rc;
    if (int rc = tcl_tmap(&t_out, g, out, I)) return rc;
    const int grid = (int)(n_items < kNumSMs ? n_items : kNumSMs);
    attn_fwd_tc_long_kernel<<<grid, TCB_THREADS, kTclSmem, st>>>(t_qkv, t_out, g, lse, drop, n_items, qtiles, nk);
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}

}  // namespace msst
