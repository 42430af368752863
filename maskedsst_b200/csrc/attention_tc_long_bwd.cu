// Long sequences (N > 64) backward on tcgen05 / TMEM, bf16 mode.  Reference: autograd of Attention.forward,
// src/vit_spatial_spectral.py:67-77 at the sequence lengths image_size = 64 allows (pretrain.py:99-107: S = 4096), where the
// reference materialises the N x N attention matrix (:71-76).
//
// Two passes, no atomics (as the mma.sync kernels they replace, attn_bwd_bf16_long_kernel<1|2>), both instances of ONE kernel:
//   PASS 0 (dK, dV): item = (sequence, head, 128-key tile j).  K_j, V_j stay in shared memory, the (Q_i, dO_i) tiles of all query
//                    tiles i stream through a TMA ring;  dV_j += P~^T dO_i and dK_j += dS^T Q_i accumulate in TMEM over i.
//   PASS 1 (dQ)    : item = (sequence, head, 128-query tile i).  Q_i, dO_i stay, the (K_j, V_j) tiles stream;  dQ_i += dS K_j
//                    accumulates in TMEM over j.
// A (query tile, key tile) pair is processed as two HALF steps of 64 keys: S_h = Q_i K_{j,h}^T and dP_h = dO_i V_{j,h}^T are
// 128 x 64 accumulators (128 TMEM columns per step, two steps in flight), so the S / dP contractions of step t + 1 run under the
// softmax arithmetic of step t, and the 512 softmax threads (thread = query row x 16 of the 64 key columns) keep 16 + 16 values
// in registers.  P~_h and dS_h (bf16, [128 queries][64 keys] SWIZZLE_128B tiles, double-buffered) are the MN-major A operands of the
// dV / dK contractions (two M = 64 atoms per key tile, interleaved in the same 64 TMEM columns: half h owns lanes 32q + 16h ..)
// and the K-major A operand of dQ.  The softmax backward needs D_i = sum_j P~_ij dP_ij = rowsum(dO_i * O_i): a small pre-pass.
// Dropout pair indices are those of the forward kernels (64 x 64 tile coordinates), so every kernel regenerates the same mask.
#include "common.cuh"
#include "kernels.h"
#include "attn_geom.cuh"
#include "ptx.cuh"
#include <stdlib.h>
#include <stdio.h>

namespace msst {
using namespace ptx;
typedef __nv_bfloat16 bf16;
int make_tmap_bf16_4d(CUtensorMap* m, const void* base, const int64_t dims[4], const int64_t strides[3], int box1, int box2);   // gemm_bf16.cu

namespace {

constexpr int LB_THREADS = 608;            // warps 0-15 softmax / epilogue, 16 TMA producer (+ TMEM alloc), 17 MMA issuer, 18 TMA store
constexpr int LB_ROWS = 128;
constexpr uint32_t LB_TILE = LB_ROWS * 128;   // one [128 rows][64 bf16] SWIZZLE_128B tile (16 KB)
constexpr int LB_RING = 2;

struct alignas(8) LbBars {
    uint64_t stat_full[2], stat_empty[2], ring_full[LB_RING], ring_empty[LB_RING], s_full[2], s_free[2], p_full[2], p_free[2],
             acc_full[2], acc_free[2], stg_full[2], stg_free[2];
    uint32_t tmem_base;
};
// smem: stationary [2][2 tiles] | ring [LB_RING][2 tiles] | P~ [2] | dS [2] | staging [2 tiles] | barriers
constexpr size_t kLbSmem = (size_t)(2 * 2 + LB_RING * 2 + 2 + 2 + 2) * LB_TILE + sizeof(LbBars);

__device__ __forceinline__ uint32_t pack_bf(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void lb_item(const AttnGeom& g, int64_t item, int nt, int64_t& seq, int& h, int& tile) {
    tile = (int)(item % nt); h = (int)((item / nt) % g.H); seq = item / ((int64_t)nt * g.H);
}
// TMA coordinates of rows [pos0, pos0 + 128) of sequence seq (see lb_tmap)
__device__ __forceinline__ void lb_coords(const AttnGeom& g, int64_t seq, int pos0, int& c1, int& c2, int& c3) {
    if (g.inner == 1) { c1 = pos0; c2 = (int)seq; c3 = 0; }
    else { c1 = (int)(seq % g.inner); c2 = pos0; c3 = (int)(seq / g.inner); }
}
// K-major / MN-major SWIZZLE_128B descriptors of a [rows][64 bf16] tile (8-row atoms of 1 KB)
__device__ __forceinline__ uint64_t kdesc(uint32_t addr) { return make_smem_desc(addr, 16, 1024); }          // + 2 per K = 16 columns
__device__ __forceinline__ uint64_t mndesc(uint32_t addr) { return make_smem_desc(addr, LB_TILE, 1024); }    // + 128 per K = 16 rows

// D[row, head] = sum_d dO[row, head*64 + d] * O[row, head*64 + d]   (one warp per row: lane = 16 columns, 4 lanes per head)
__global__ void __launch_bounds__(256) attn_rowdot_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, float* __restrict__ D, int64_t rows, int H) {
    const int lane = threadIdx.x & 31;
    const int I = H * 64;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
        for (int base = 0; base < I; base += 512) {             // warp-uniform trip count: the shuffles below need all 32 lanes
            const int c0 = base + lane * 16;
            float s = 0.f;
            if (c0 < I) {
                const uint4* po = reinterpret_cast<const uint4*>(o + r * I + c0);
                const uint4* pd = reinterpret_cast<const uint4*>(d_o + r * I + c0);
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    const uint4 a = __ldg(po + v), b = __ldg(pd + v);
                    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[e]));
                        const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bw[e]));
                        s = fmaf(fa.x, fb.x, s); s = fmaf(fa.y, fb.y, s);
                    }
                }
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (c0 < I && (lane & 3) == 0) D[r * H + (c0 >> 6)] = s;
        }
    }
}

template <int PASS, bool DROP>
__global__ void __launch_bounds__(LB_THREADS, 1)
attn_bwd_tc_long_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do, const __grid_constant__ CUtensorMap tma_dqkv,
                        AttnGeom g, const float* __restrict__ lse, const float* __restrict__ Dvec, Drop drop, int64_t n_items, int nt, long long* dbg) {
#define LB_T(t, slot) do { if (dbg && blockIdx.x == 0 && (t) < 64) dbg[(t) * 8 + (slot)] = clock64(); } while (0)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (smem_u32(smem) & 1023u) __trap();
    uint8_t* stat_s = smem;                                 // [2][X0 | X1]   PASS 0: K_j | V_j     PASS 1: Q_i | dO_i
    uint8_t* ring_s = stat_s + 2 * 2 * LB_TILE;             // [RING][Y0 | Y1] PASS 0: Q_i | dO_i    PASS 1: K_j | V_j
    uint8_t* p_s = ring_s + LB_RING * 2 * LB_TILE;          // P~ [2]
    uint8_t* ds_s = p_s + 2 * LB_TILE;                      // dS [2]
    uint8_t* stg_s = ds_s + 2 * LB_TILE;                    // staging: PASS 0 dK | dV, PASS 1 dQ
    LbBars* bars = reinterpret_cast<LbBars*>(stg_s + 2 * LB_TILE);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int I = g.H * 64;

    if (warp == 17 && elect_one()) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars->stat_full[s], 1); mbar_init(&bars->stat_empty[s], 1); mbar_init(&bars->s_full[s], 1); mbar_init(&bars->s_free[s], 8);
            mbar_init(&bars->p_full[s], 8); mbar_init(&bars->p_free[s], 1); mbar_init(&bars->acc_full[s], 1); mbar_init(&bars->acc_free[s], 16);
            mbar_init(&bars->stg_full[s], 16); mbar_init(&bars->stg_free[s], 1);
        }
        for (int s = 0; s < LB_RING; ++s) { mbar_init(&bars->ring_full[s], 1); mbar_init(&bars->ring_empty[s], 1); }
        fence_barrier_init();
    }
    if (warp == 16) {
        tmem_alloc(&bars->tmem_base, 512);
        if (elect_one()) { prefetch_tmap(&tma_qkv); prefetch_tmap(&tma_do); prefetch_tmap(&tma_dqkv); }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const int64_t my_items = blockIdx.x < n_items ? (n_items - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int64_t steps_per_item = 2 * (int64_t)nt, total = my_items * steps_per_item;
    // TMEM: step slots [0,128) / [128,256): S_h | dP_h;  accumulators [256,384) / [384,512): PASS 0 dV | dK (two M = 64 atoms), PASS 1 dQ
    constexpr uint32_t COL_ACC = 256;

    if (warp == 16) {
        // ===== TMA producer =====
        if (elect_one()) {
            int ring = 0; uint32_t rph = 0;
            for (int64_t n = 0; n < my_items; ++n) {
                int64_t seq; int h, tile, c1, c2, c3;
                lb_item(g, blockIdx.x + n * gridDim.x, nt, seq, h, tile);
                const int ss = (int)(n & 1);
                mbar_wait(&bars->stat_empty[ss], (uint32_t)((n >> 1) & 1) ^ 1);
                lb_coords(g, seq, tile * LB_ROWS, c1, c2, c3);
                uint8_t* sd = stat_s + (size_t)ss * 2 * LB_TILE;
                mbar_arrive_expect_tx(&bars->stat_full[ss], 2 * LB_TILE);
                if (PASS == 0) {
                    tma_load_4d(sd, &tma_qkv, &bars->stat_full[ss], I + h * 64, c1, c2, c3);
                    tma_load_4d(sd + LB_TILE, &tma_qkv, &bars->stat_full[ss], 2 * I + h * 64, c1, c2, c3);
                } else {
                    tma_load_4d(sd, &tma_qkv, &bars->stat_full[ss], h * 64, c1, c2, c3);
                    tma_load_4d(sd + LB_TILE, &tma_do, &bars->stat_full[ss], h * 64, c1, c2, c3);
                }
                for (int j = 0; j < nt; ++j) {
                    mbar_wait(&bars->ring_empty[ring], rph ^ 1);
                    lb_coords(g, seq, j * LB_ROWS, c1, c2, c3);
                    uint8_t* dst = ring_s + (size_t)ring * 2 * LB_TILE;
                    mbar_arrive_expect_tx(&bars->ring_full[ring], 2 * LB_TILE);
                    if (PASS == 0) {
                        tma_load_4d(dst, &tma_qkv, &bars->ring_full[ring], h * 64, c1, c2, c3);
                        tma_load_4d(dst + LB_TILE, &tma_do, &bars->ring_full[ring], h * 64, c1, c2, c3);
                    } else {
                        tma_load_4d(dst, &tma_qkv, &bars->ring_full[ring], I + h * 64, c1, c2, c3);
                        tma_load_4d(dst + LB_TILE, &tma_qkv, &bars->ring_full[ring], 2 * I + h * 64, c1, c2, c3);
                    }
                    if (++ring == LB_RING) { ring = 0; rph ^= 1; }
                }
            }
        }
    } else if (warp == 17) {
        // ===== MMA issuer: S_h | dP_h of step ts (at most one step ahead) and the accumulating contractions of step tp, whichever is ready =====
        if (elect_one()) {
            const uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0), idesc_mn = make_idesc_bf16(64, 64, 1, 1), idesc_q = make_idesc_bf16(128, 64, 0, 1);
            int64_t ts = 0, tp = 0;
            int ring_a = 0; uint32_t rph_a = 0; int ring_b = 0;
            // (item, step within the item) of ts and tp, advanced incrementally: a 64-bit division per poll of this loop made the issuing
            // thread the bottleneck of the kernel (~80 clks per MMA)
            const int spi = (int)steps_per_item;
            int64_t n_s = 0, n_p = 0; int rem_s = 0, rem_p = 0;
            while (tp < total) {
                if (ts < total && ts <= tp + 1) {
                    const int64_t n = n_s; const int rem = rem_s;
                    const int hh = rem & 1, slot = (int)(ts & 1);
                    bool ok = mbar_try_wait(&bars->s_free[slot], (uint32_t)((ts >> 1) & 1) ^ 1) && mbar_try_wait(&bars->ring_full[ring_a], rph_a);
                    if (ok && rem == 0) ok = mbar_try_wait(&bars->stat_full[n & 1], (uint32_t)(n >> 1) & 1);
                    if (ok) {
                        LB_T(ts, 0);
                        tc_fence_after();
                        const uint32_t xs = smem_u32(stat_s + (size_t)(n & 1) * 2 * LB_TILE), ys = smem_u32(ring_s + (size_t)ring_a * 2 * LB_TILE);
                        // queries are the M dimension (TMEM lane = query row), the half's 64 keys the N dimension
                        const uint32_t q_a = PASS == 0 ? ys : xs, do_a = q_a + LB_TILE;
                        const uint32_t k_b = (PASS == 0 ? xs : ys) + (uint32_t)hh * 8192u, v_b = k_b + LB_TILE;
                        const uint32_t d0 = tmem_base + (uint32_t)slot * 128u;
                        for (int k = 0; k < 4; ++k) umma_bf16(d0, kdesc(q_a) + (uint64_t)(k * 2), kdesc(k_b) + (uint64_t)(k * 2), idesc_s, k != 0);            // S_h = Q K_h^T
                        for (int k = 0; k < 4; ++k) umma_bf16(d0 + 64, kdesc(do_a) + (uint64_t)(k * 2), kdesc(v_b) + (uint64_t)(k * 2), idesc_s, k != 0);     // dP_h = dO V_h^T
                        umma_commit(&bars->s_full[slot]);
                        ++ts;
                        if (++rem_s == spi) { rem_s = 0; ++n_s; }
                        if (hh == 1 && ++ring_a == LB_RING) { ring_a = 0; rph_a ^= 1; }
                    }
                }
                if (tp < ts) {
                    const int64_t n = n_p; const int rem = rem_p;
                    const int hh = rem & 1, tile = rem >> 1, pb = (int)(tp & 1), os = (int)(n & 1);
                    bool ok = mbar_try_wait(&bars->p_full[pb], (uint32_t)(tp >> 1) & 1);
                    if (ok && rem == 0) ok = mbar_try_wait(&bars->acc_free[os], (uint32_t)((n >> 1) & 1) ^ 1);
                    if (ok) {
                        LB_T(tp, 1);
                        tc_fence_after();
                        const uint32_t xs = smem_u32(stat_s + (size_t)(n & 1) * 2 * LB_TILE), ys = smem_u32(ring_s + (size_t)ring_b * 2 * LB_TILE);
                        const uint32_t pa = smem_u32(p_s + (size_t)pb * LB_TILE), dsa = smem_u32(ds_s + (size_t)pb * LB_TILE);
                        const uint32_t acc = tmem_base + COL_ACC + (uint32_t)os * 128u;
                        if (PASS == 0) {
                            const uint32_t lo = (uint32_t)(16 * hh) << 16;      // half h of the key tile: lanes 32q + 16h ..
                            for (int k = 0; k < 8; ++k)   // dV_h [64 keys][64] += P~_h^T dO_i   (reduction over the 128 queries, 16 rows = 2 KB per step)
                                umma_bf16(acc + lo, mndesc(pa) + (uint64_t)(k * 128), mndesc(ys + LB_TILE) + (uint64_t)(k * 128), idesc_mn, (tile | k) != 0);
                            for (int k = 0; k < 8; ++k)   // dK_h += dS_h^T Q_i
                                umma_bf16(acc + 64 + lo, mndesc(dsa) + (uint64_t)(k * 128), mndesc(ys) + (uint64_t)(k * 128), idesc_mn, (tile | k) != 0);
                        } else {
                            const uint32_t k_b = ys + (uint32_t)hh * 8192u;
                            for (int k = 0; k < 4; ++k)   // dQ_i [128][64] += dS_h K_h   (reduction over the half's 64 keys)
                                umma_bf16(acc, kdesc(dsa) + (uint64_t)(k * 2), mndesc(k_b) + (uint64_t)(k * 128), idesc_q, (rem | k) != 0);
                        }
                        umma_commit(&bars->p_free[pb]);
                        if (hh == 1) { umma_commit(&bars->ring_empty[ring_b]); if (++ring_b == LB_RING) ring_b = 0; }
                        if (rem == spi - 1) { umma_commit(&bars->acc_full[os]); umma_commit(&bars->stat_empty[n & 1]); }
                        ++tp;
                        if (++rem_p == spi) { rem_p = 0; ++n_p; }
                    }
                }
            }
        }
    } else if (warp == 18) {
        // ===== TMA store of the staged result tiles =====
        if (elect_one()) {
            for (int64_t n = 0; n < my_items; ++n) {
                int64_t seq; int h, tile, c1, c2, c3;
                lb_item(g, blockIdx.x + n * gridDim.x, nt, seq, h, tile);
                lb_coords(g, seq, tile * LB_ROWS, c1, c2, c3);
                mbar_wait(&bars->stg_full[n & 1], (uint32_t)(n >> 1) & 1);
                if (PASS == 0) {
                    tma_store_4d(&tma_dqkv, stg_s, I + h * 64, c1, c2, c3);
                    tma_store_4d(&tma_dqkv, stg_s + LB_TILE, 2 * I + h * 64, c1, c2, c3);
                } else {
                    tma_store_4d(&tma_dqkv, stg_s, h * 64, c1, c2, c3);
                }
                tma_store_commit();
                tma_store_wait_read();
                mbar_arrive(&bars->stg_free[n & 1]);
            }
        }
    } else {
        // ===== two softmax groups of 8 warps ping-pong over the half steps (group = parity of the step = key half hh): thread = (query row r,
        // 32 of the half's 64 key columns, handled as two rounds of 16 to keep the register footprint of 16 + 16 values); while one group
        // waits for its S_h | dP_h or for its P~ / dS buffer, the other one issues.  The item epilogue uses all 16 warps (column quarter cq). =====
        const int lq = warp & 3, cq = warp >> 2, grp = warp >> 3, ch = (warp >> 2) & 1;
        const int r = lq * 32 + lane;
        const uint32_t lane_base = (uint32_t)(lq * 32) << 16;
        const uint32_t swz = (uint32_t)(r & 7);
        const uint32_t c0 = (((uint32_t)(2 * cq)) ^ swz) << 4, c1o = (((uint32_t)(2 * cq + 1)) ^ swz) << 4;
        const float sl2 = g.scale * 1.4426950408889634f;
        const uint64_t seed = drop.seed + (drop.seed_dev ? __ldg(drop.seed_dev) : 0ull);
        const uint32_t t16 = drop.thresh >> 16;
        for (int64_t n = 0; n < my_items; ++n) {
            int64_t seq; int h, stile;
            lb_item(g, blockIdx.x + n * gridDim.x, nt, seq, h, stile);
            float L = 0.f, Dv = 0.f; bool q_ok = false;
            for (int jt = 0; jt < nt; ++jt) {
                const int qt = PASS == 0 ? jt : stile, kt = PASS == 0 ? stile : jt;     // 128-row query / key tile of this pair
                if (PASS == 0 || jt == 0) {                     // row statistics of query row r of tile qt
                    const int qpos = qt * LB_ROWS + r;
                    q_ok = qpos < g.N;
                    if (q_ok) {
                        const int64_t row = row_of(g, seq, qpos);
                        L = __ldg(lse + row * g.H + h) * 1.4426950408889634f;
                        Dv = __ldg(Dvec + row * g.H + h);
                    }
                }
                const int hh = grp;
                const int64_t t = n * steps_per_item + 2 * jt + hh;
                const int slot = (int)(t & 1), pb = slot;
                const uint32_t ph = (uint32_t)(t >> 1) & 1;
                if (r == 0 && ch == 0) LB_T(t, 2);
                mbar_wait(&bars->s_full[slot], ph);
                if (r == 0 && ch == 0) LB_T(t, 3);
                tc_fence_after();
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    const int cs = 2 * ch + sub;                // 16-column group of the half's 64 keys
                    float sp[16], dp[16];
                    {
                        uint32_t a[16], b[16];
                        tmem_ld_32x16(tmem_base + lane_base + (uint32_t)slot * 128u + 16 * cs, a);
                        tmem_ld_32x16(tmem_base + lane_base + (uint32_t)slot * 128u + 64 + 16 * cs, b);
                        tmem_ld_wait();
                        if (sub == 1) { tc_fence_before(); warp_arrive(&bars->s_free[slot], lane); }
#pragma unroll
                        for (int j = 0; j < 16; ++j) { sp[j] = __uint_as_float(a[j]); dp[j] = __uint_as_float(b[j]); }
                    }
                    // 64 x 64 tile coordinates of the forward kernels: qt64 = 2 qt + r / 64, kt64 = 2 kt + hh, pair = (r & 63) * 32 + 8 cs + jj
                    uint32_t hash_lo = 0, hash_hi = 0;
                    if (DROP) {
                        const uint64_t hidx = ((((uint64_t)seq * g.H + h) * g.tiles + (uint64_t)(2 * qt + (r >> 6))) * g.tiles + (uint64_t)(2 * kt + hh)) *
                                                  (uint64_t)(TS * TS / 2) + (uint64_t)((r & 63) * 32 + 8 * cs);
                        hash_lo = (uint32_t)hidx + (uint32_t)(seed >> 32);
                        hash_hi = ((uint32_t)(hidx >> 32) * 0x85EBCA77u) ^ (uint32_t)seed ^ (drop.site * 0xC2B2AE3Du);
                    }
                    const int nvalid = q_ok ? g.N - kt * LB_ROWS - hh * 64 - 16 * cs : 0;     // this thread's columns 0 .. nvalid-1 are real keys
                    uint32_t pk[8], dk[8];
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const int j = 2 * jj;
                        float p0 = j < nvalid ? ex2_approx(fmaf(sp[j], sl2, -L)) : 0.f;
                        float p1 = j + 1 < nvalid ? ex2_approx(fmaf(sp[j + 1], sl2, -L)) : 0.f;
                        float d0 = dp[j], d1 = dp[j + 1];
                        if (DROP) {
                            uint32_t x = (hash_lo + (uint32_t)jj) * 0x9E3779B1u ^ hash_hi;
                            x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
                            const float f0 = (x & 0xFFFFu) >= t16 ? drop.scale : 0.f, f1 = (x >> 16) >= t16 ? drop.scale : 0.f;
                            if (PASS == 0) pk[jj] = pack_bf(p0 * f0, p1 * f1);
                            d0 *= f0; d1 *= f1;
                        } else if (PASS == 0) pk[jj] = pack_bf(p0, p1);
                        p0 *= g.scale; p1 *= g.scale;
                        dk[jj] = pack_bf(p0 * (d0 - Dv), p1 * (d1 - Dv));
                    }
                    if (sub == 0) { if (r == 0 && ch == 0) LB_T(t, 4); mbar_wait(&bars->p_free[pb], ph ^ 1); if (r == 0 && ch == 0) LB_T(t, 5); }   // the contractions of step t - 2 have read this P~ / dS buffer
                    const uint32_t o0 = (((uint32_t)(2 * cs)) ^ swz) << 4, o1 = (((uint32_t)(2 * cs + 1)) ^ swz) << 4;
                    if (PASS == 0) {
                        uint8_t* prow = p_s + (size_t)pb * LB_TILE + r * 128;
                        *reinterpret_cast<uint4*>(prow + o0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4*>(prow + o1) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
                    uint8_t* drow = ds_s + (size_t)pb * LB_TILE + r * 128;
                    *reinterpret_cast<uint4*>(drow + o0) = make_uint4(dk[0], dk[1], dk[2], dk[3]);
                    *reinterpret_cast<uint4*>(drow + o1) = make_uint4(dk[4], dk[5], dk[6], dk[7]);
                }
                fence_proxy_async();
                tc_fence_before();
                warp_arrive(&bars->p_full[pb], lane);
                if (r == 0 && ch == 0) LB_T(t, 6);
            }
            // ---- item epilogue: accumulators -> bf16 -> staging tiles -> TMA store ----
            const int os = (int)(n & 1);
            mbar_wait(&bars->acc_full[os], (uint32_t)(n >> 1) & 1);
            if (n > 0) mbar_wait(&bars->stg_free[(n - 1) & 1], (uint32_t)((n - 1) >> 1) & 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + lane_base + COL_ACC + (uint32_t)os * 128u;
            if (PASS == 0) {
                // M = 64 atoms: lane 32 q + 16 h + i holds key row 64 h + 16 q + i of the tile
                const int key = 64 * ((lane >> 4) & 1) + 16 * lq + (lane & 15);
                uint32_t a[16], b[16];
                tmem_ld_32x16(acc + 16 * cq, a);            // dV
                tmem_ld_32x16(acc + 64 + 16 * cq, b);       // dK
                tmem_ld_wait();
                tc_fence_before();
                warp_arrive(&bars->acc_free[os], lane);
                const uint32_t kz = (uint32_t)(key & 7);
                uint8_t* krow = stg_s + key * 128; uint8_t* vrow = krow + LB_TILE;
                const uint32_t k0 = (((uint32_t)(2 * cq)) ^ kz) << 4, k1 = (((uint32_t)(2 * cq + 1)) ^ kz) << 4;
                *reinterpret_cast<uint4*>(krow + k0) = make_uint4(pack_bf(__uint_as_float(b[0]), __uint_as_float(b[1])), pack_bf(__uint_as_float(b[2]), __uint_as_float(b[3])),
                                                                  pack_bf(__uint_as_float(b[4]), __uint_as_float(b[5])), pack_bf(__uint_as_float(b[6]), __uint_as_float(b[7])));
                *reinterpret_cast<uint4*>(krow + k1) = make_uint4(pack_bf(__uint_as_float(b[8]), __uint_as_float(b[9])), pack_bf(__uint_as_float(b[10]), __uint_as_float(b[11])),
                                                                  pack_bf(__uint_as_float(b[12]), __uint_as_float(b[13])), pack_bf(__uint_as_float(b[14]), __uint_as_float(b[15])));
                *reinterpret_cast<uint4*>(vrow + k0) = make_uint4(pack_bf(__uint_as_float(a[0]), __uint_as_float(a[1])), pack_bf(__uint_as_float(a[2]), __uint_as_float(a[3])),
                                                                  pack_bf(__uint_as_float(a[4]), __uint_as_float(a[5])), pack_bf(__uint_as_float(a[6]), __uint_as_float(a[7])));
                *reinterpret_cast<uint4*>(vrow + k1) = make_uint4(pack_bf(__uint_as_float(a[8]), __uint_as_float(a[9])), pack_bf(__uint_as_float(a[10]), __uint_as_float(a[11])),
                                                                  pack_bf(__uint_as_float(a[12]), __uint_as_float(a[13])), pack_bf(__uint_as_float(a[14]), __uint_as_float(a[15])));
            } else {
                uint32_t a[16];
                tmem_ld_32x16(acc + 16 * cq, a);            // dQ: lane = query row
                tmem_ld_wait();
                tc_fence_before();
                warp_arrive(&bars->acc_free[os], lane);
                uint8_t* qrow = stg_s + r * 128;
                *reinterpret_cast<uint4*>(qrow + c0) = make_uint4(pack_bf(__uint_as_float(a[0]), __uint_as_float(a[1])), pack_bf(__uint_as_float(a[2]), __uint_as_float(a[3])),
                                                                  pack_bf(__uint_as_float(a[4]), __uint_as_float(a[5])), pack_bf(__uint_as_float(a[6]), __uint_as_float(a[7])));
                *reinterpret_cast<uint4*>(qrow + c1o) = make_uint4(pack_bf(__uint_as_float(a[8]), __uint_as_float(a[9])), pack_bf(__uint_as_float(a[10]), __uint_as_float(a[11])),
                                                                   pack_bf(__uint_as_float(a[12]), __uint_as_float(a[13])), pack_bf(__uint_as_float(a[14]), __uint_as_float(a[15])));
            }
            fence_proxy_async();
            warp_arrive(&bars->stg_full[n & 1], lane);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
#undef LB_T
}

// 4-D view whose box is 128 consecutive positions of ONE sequence (rows beyond the sequence end: zero on load, clipped on store)
int lb_tmap(CUtensorMap* m, const AttnGeom& g, const bf16* base, int64_t cols) {
    if (g.inner == 1) {
        const int64_t dims[4] = {cols, g.N, g.n_seq, 1}, strides[3] = {cols, (int64_t)g.N * cols, g.n_seq * g.N * cols};
        return make_tmap_bf16_4d(m, base, dims, strides, LB_ROWS, 1);
    }
    const int64_t dims[4] = {cols, g.inner, g.N, g.n_seq / g.inner};
    const int64_t strides[3] = {cols, (int64_t)g.inner * cols, (int64_t)g.N * g.inner * cols};
    return make_tmap_bf16_4d(m, base, dims, strides, 1, LB_ROWS);
}

// per-device scratch for D [rows, H] fp32, grown on demand (cudaMalloc: call once outside any stream capture, e.g. a warm-up step)
float* rowdot_scratch(size_t floats) {
    static float* buf[64] = {};
    static size_t cap[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (cap[dev] < floats) {
        if (buf[dev]) cudaFree(buf[dev]);
        buf[dev] = nullptr; cap[dev] = 0;
        if (cudaMalloc(&buf[dev], floats * sizeof(float)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        cap[dev] = floats;
    }
    return buf[dev];
}

}  // namespace

bool attention_bwd_tc_long_supported(const AttnGeom& g) {
    return g.tiles > 1 && g.n_seq > 0 && g.n_seq < (int64_t)2147483647 && (int64_t)g.N < (int64_t)2147483647 - 256 &&
           g.n_seq * g.N * g.H < ((int64_t)1 << 40);
}

int attention_bwd_tc_long(const AttnGeom& g, const bf16* qkv, const bf16* out, const float* lse, const bf16* d_out, bf16* d_qkv, Drop drop, cudaStream_t st) {
    MSST_REQUIRE(attention_bwd_tc_long_supported(g), "attention_bwd_tc_long: needs N > 64");
    static PerDeviceOnce once;
    if (once.first()) {
        MSST_CUDA(cudaFuncSetAttribute(attn_bwd_tc_long_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLbSmem));
        MSST_CUDA(cudaFuncSetAttribute(attn_bwd_tc_long_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLbSmem));
        MSST_CUDA(cudaFuncSetAttribute(attn_bwd_tc_long_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLbSmem));
        MSST_CUDA(cudaFuncSetAttribute(attn_bwd_tc_long_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLbSmem));
    }
    const int64_t I = (int64_t)g.H * 64, rows = g.n_seq * g.N;
    float* D = rowdot_scratch((size_t)rows * g.H);
    MSST_REQUIRE(D != nullptr, "attention_bwd_tc_long: cannot allocate the %lld-float row-statistics scratch (run one step outside CUDA-graph capture first)",
                 (long long)(rows * g.H));
    int64_t rgrid = (rows + 7) / 8;
    if (rgrid > 16 * kNumSMs) rgrid = 16 * kNumSMs;
    attn_rowdot_kernel<<<(int)rgrid, 256, 0, st>>>(out, d_out, D, rows, g.H);
    MSST_LAUNCH_CHECK();
    const int nt = (g.N + LB_ROWS - 1) / LB_ROWS;
    const int64_t n_items = g.n_seq * g.H * nt;
    CUtensorMap t_qkv, t_do, t_dqkv;
    if (int rc = lb_tmap(&t_qkv, g, qkv, 3 * I)) return rc;
    if (int rc = lb_tmap(&t_do, g, d_out, I)) return rc;
    if (int rc = lb_tmap(&t_dqkv, g, d_qkv, 3 * I)) return rc;
    const int grid = (int)(n_items < kNumSMs ? n_items : kNumSMs);
    static int dbg_on = -1;
    static long long* dbg = nullptr;
    if (dbg_on < 0) { const char* e = getenv("MSST_LB_DBG"); dbg_on = e ? atoi(e) : 0; if (dbg_on) cudaMalloc(&dbg, 2 * 64 * 8 * 8); }
    if (dbg_on) cudaMemsetAsync(dbg, 0, 2 * 64 * 8 * 8, st);
    if (drop.on()) {
        attn_bwd_tc_long_kernel<0, true><<<grid, LB_THREADS, kLbSmem, st>>>(t_qkv, t_do, t_dqkv, g, lse, D, drop, n_items, nt, nullptr);
        MSST_LAUNCH_CHECK();
        attn_bwd_tc_long_kernel<1, true><<<grid, LB_THREADS, kLbSmem, st>>>(t_qkv, t_do, t_dqkv, g, lse, D, drop, n_items, nt, nullptr);
    } else {
        attn_bwd_tc_long_kernel<0, false><<<grid, LB_THREADS, kLbSmem, st>>>(t_qkv, t_do, t_dqkv, g, lse, D, drop, n_items, nt, dbg);
        MSST_LAUNCH_CHECK();
        attn_bwd_tc_long_kernel<1, false><<<grid, LB_THREADS, kLbSmem, st>>>(t_qkv, t_do, t_dqkv, g, lse, D, drop, n_items, nt, dbg ? dbg + 64 * 8 : nullptr);
    }
    MSST_LAUNCH_CHECK();
    if (dbg_on) {
        static int calls = 0;
        if (++calls == dbg_on) {
            static long long hbuf[2 * 64 * 8];
            cudaStreamSynchronize(st);
            cudaMemcpy(hbuf, dbg, sizeof(hbuf), cudaMemcpyDeviceToHost);
            for (int pass = 0; pass < 2; ++pass) {
                printf("attn_bwd_tc_long pass %d timeline, CTA 0, per half step: S issue | contractions issue | (group thread 0) wait S | S ready | computed | buffer free | written\n", pass);
                const long long* hb = hbuf + pass * 64 * 8; const long long t0 = hb[0];
                for (int i = 0; i < 40; ++i) { printf("%2d:", i); for (int k = 0; k < 7; ++k) printf(" %7lld", hb[i * 8 + k] ? hb[i * 8 + k] - t0 : -1); printf("\n"); }
            }
            fflush(stdout);
        }
    }
    return MSST_OK;
}

}  // namespace msst
