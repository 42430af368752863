// Kernels (2)+(3) fused, bf16 mode, packed short sequences (N <= 64): the per-head QKV projection runs INSIDE the attention
// kernels, so q / k / v never reach HBM in either direction.
//   forward :  h = LN1(x) [R, D] bf16  ->  per (128-slot tile, head):  [Q|K|V] = h W_h^T  ->  softmax(Q K^T dh^-0.5) V  ->  o [R, H*64] bf16
//   backward:  recomputes [Q|K|V] from the saved h (192 B / token instead of re-reading 3 KB / token of qkv), runs the five
//              attention-backward contractions, and folds the data gradient of the projection in:
//              dh += [dQ|dK|dV] W_h accumulated over the heads in TMEM.  Outputs: dqkv [R, 3*H*64] bf16 (for the weight-gradient
//              GEMM) and dh [R, D] fp32 (the LayerNorm backward's dy).
// Reference: Attention.forward, src/vit_spatial_spectral.py:67-77 (to_qkv :59,68; scores / softmax / dropout / PV :71-76) and
// its autograd.  Geometry, slot packing and the dropout pair hash are those of attention_tc.cu (attn_geom.cuh), so the fused
// kernels regenerate exactly the masks of the unfused ones.
//
// One persistent CTA per SM, 18 warps; an item = (128-slot tile, head), items of a tile run head after head:
//   warp 16  TMA producer: the tile's h rows (once per tile, double-buffered; K = D as 32-column SWIZZLE_64B chunks), the
//            head's weight sub-slices through a ring (W_q|k|v rows of the head, and in the backward the W^T column blocks for
//            the data gradient), dO of the item (backward)
//   warp 17  MMA issuer (one elected thread), everything on tcgen05:
//            prologue  [Q|K|V] (128 x 192, K = D)       -> TMEM            issued one item ahead (overlaps the softmax)
//            S_b = Q_b K_b^T, (dP_b = dO_b V_b^T)        two M = 64 atoms INTERLEAVED in the same 64 TMEM columns: block b of
//                                                        the tile (one 64-slot group) owns lanes 32j + 16b .. +15 -- no
//                                                        block-diagonal zero padding, no masked-away tensor work
//            O_b = P~_b V_b | dV_b, dK_b, dQ_b           same M = 64 form
//            dh += dQ W_q,h + dK W_k,h + dV W_v,h        A operand FROM TMEM (bf16 rows written back by the epilogue threads)
//   warps 0-15  512 threads work on ONE item at a time: thread = (TMEM lane, column quarter).  TMEM -> registers -> bf16 ->
//            swizzled smem tiles (Q, K, V, P~, dS) that the next contraction consumes; results leave with 32-byte row stores.
// TMEM columns: [0,192) prologue accumulators | forward: S [192,256), O [256,320) | backward: X = [192,384) holds S | dP, then
// dV | dK | dQ, then the bf16 A operand of the data gradient; dh accumulator [384, 384 + D).
#include "common.cuh"
#include "kernels.h"
#include "attn_geom.cuh"
#include "ptx.cuh"
#include <stdlib.h>

namespace msst {
using namespace ptx;
typedef __nv_bfloat16 bf16;
int make_tmap_bf16_nd(CUtensorMap* m, const void* base, int rank, const int64_t* dims, const int64_t* strides, const int* box, int swizzle_bytes);   // gemm_bf16.cu

namespace {

constexpr int AB_THREADS = 576;            // warps 0-15 compute, 16 TMA producer (+ TMEM alloc), 17 MMA issuer
constexpr int AB_MAXR = 6;                 // weight ring slots (upper bound)
constexpr uint32_t AB_T16 = 16384;         // one [128 rows][64 bf16] SWIZZLE_128B tile
constexpr uint32_t AB_BLK = 8192;          // one 64-row block of such a tile
constexpr uint32_t COL_P = 0, COL_S = 192, COL_O = 256, COL_X = 192, COL_DH = 384;

struct AbParams {
    AttnGeom g;
    int D, nch;                // model dim, number of 32-column chunks
    int nbox;                  // 1: a tile is 128 consecutive rows (2-D boxes); 2: one 4-D box per 64-slot group
    int NR;                    // weight ring slots in use
    int64_t n_tiles;
    uint32_t slot_bytes;       // D * 128: one weight sub-slice ([64 rows][D] as chunks, or [D rows][64])
    uint32_t hbuf_bytes;       // nch * 8192: one h tile
    const float* lse_in; float* lse_out;
    bf16* o; bf16* dqkv; float* dh;
    Drop drop;
};

__device__ __forceinline__ uint32_t pack_bf(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint64_t pair_base(const AttnGeom& g, int64_t group, int h) {   // == tile_pair_base_tc (attention_tc.cu)
    return ((uint64_t)group * g.H + h) * (uint64_t)(TS * TS / 2);
}
__device__ __forceinline__ void group_coords(const AttnGeom& g, int64_t group, int& c1, int& c2, int& c3) {
    if (g.gpb == 0) { c1 = 0; c2 = (int)(group * g.G); c3 = 0; }
    else { c1 = (int)(group % g.gpb) * g.G; c2 = 0; c3 = (int)(group / g.gpb); }
}
// 16 fp32 values (columns 16*cq .. +15 of row `row`) -> bf16 -> two 16-byte chunks of a SWIZZLE_128B K-major tile row
__device__ __forceinline__ void store_row16(uint8_t* tile, int row, int cq, const uint32_t (&v)[16]) {
    uint8_t* r = tile + row * 128;
    const uint32_t sw = (uint32_t)(row & 7);
    *reinterpret_cast<uint4*>(r + ((((uint32_t)(2 * cq)) ^ sw) << 4)) =
        make_uint4(pack_bf(__uint_as_float(v[0]), __uint_as_float(v[1])), pack_bf(__uint_as_float(v[2]), __uint_as_float(v[3])),
                   pack_bf(__uint_as_float(v[4]), __uint_as_float(v[5])), pack_bf(__uint_as_float(v[6]), __uint_as_float(v[7])));
    *reinterpret_cast<uint4*>(r + ((((uint32_t)(2 * cq + 1)) ^ sw) << 4)) =
        make_uint4(pack_bf(__uint_as_float(v[8]), __uint_as_float(v[9])), pack_bf(__uint_as_float(v[10]), __uint_as_float(v[11])),
                   pack_bf(__uint_as_float(v[12]), __uint_as_float(v[13])), pack_bf(__uint_as_float(v[14]), __uint_as_float(v[15])));
}
__device__ __forceinline__ void store_row16_packed(uint8_t* tile, int row, int cq, const uint32_t (&pk)[8]) {
    uint8_t* r = tile + row * 128;
    const uint32_t sw = (uint32_t)(row & 7);
    *reinterpret_cast<uint4*>(r + ((((uint32_t)(2 * cq)) ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(r + ((((uint32_t)(2 * cq + 1)) ^ sw) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
}
// which of this thread's 16 key columns (16*cq .. +15 of the row's 64-slot group) belong to the sequence of slot m
__device__ __forceinline__ uint32_t key_mask16(const AttnGeom& g, int m, int cq) {
    uint32_t vm = 0;
    const int used = g.G * g.N;
    for (int j = 0; j < 16; ++j) {
        const int c = 16 * cq + j;
        const bool same = g.gpb == 0 ? (c / g.N == m / g.N) : (c % g.G == m % g.G);
        vm |= (uint32_t)(same && c < used && m < used) << j;
    }
    return vm;
}

struct alignas(8) AbBars {
    uint64_t h_full[2], h_empty[2], w_full[AB_MAXR], w_empty[AB_MAXR], do_full[2], do_empty[2];
    uint64_t pro_full, conv_done, s_full, p_full, o_full, a_ready, dh_full, dh_free;
    uint32_t tmem_base;
};

// ---- pieces shared by the forward and the backward kernel ----
struct AbCommon {
    uint8_t *h_s, *w_s, *q_s, *k_s, *v_s;
    AbBars* bars;
    uint32_t tmem;
};

// producer: the h rows of tile `tile` -> buffer buf (nch chunks of [128 rows][32 cols], SWIZZLE_64B)
__device__ __forceinline__ void load_h_tile(const AbParams& p, const CUtensorMap* tma_h, uint8_t* dst, uint64_t* bar, int64_t tile) {
    const AttnGeom& g = p.g;
    const uint32_t rows = p.nbox == 1 ? 128u : (uint32_t)(g.G * g.N);
    mbar_arrive_expect_tx(bar, (uint32_t)p.nch * (uint32_t)p.nbox * rows * 64u);
    if (p.nbox == 1) {
        for (int c = 0; c < p.nch; ++c) tma_load_2d(dst + c * 8192, tma_h, bar, c * 32, (int)(tile * 128));
    } else {
        for (int gi = 0; gi < 2; ++gi) {
            int c1, c2, c3;
            group_coords(g, tile * 2 + gi, c1, c2, c3);
            for (int c = 0; c < p.nch; ++c) tma_load_4d(dst + c * 8192 + gi * 4096, tma_h, bar, c * 32, c1, c2, c3);
        }
    }
}
// producer: W rows [t*I + h*64, +64) x D columns as nch SWIZZLE_64B chunks -> ring slot
__device__ __forceinline__ void load_w_slice(const AbParams& p, const CUtensorMap* tma_w, AbBars* bars, uint8_t* w_s, int64_t& wc, int t, int h) {
    const int slot = (int)(wc % p.NR);
    mbar_wait(&bars->w_empty[slot], (uint32_t)((wc / p.NR) & 1) ^ 1);
    mbar_arrive_expect_tx(&bars->w_full[slot], (uint32_t)p.nch * 4096u);
    uint8_t* dst = w_s + (size_t)slot * p.slot_bytes;
    for (int c = 0; c < p.nch; ++c) tma_load_2d(dst + c * 4096, tma_w, &bars->w_full[slot], c * 32, t * p.g.H * 64 + h * 64);
    ++wc;
}
// MMA issuer: [Q|K|V] of item `it` = h tile x the head's three weight sub-slices -> TMEM columns [0, 192)
__device__ __forceinline__ void issue_prologue(const AbParams& p, const AbCommon& s, int64_t it, int64_t& wc) {
    const int H = p.g.H;
    const int64_t k = it / H; const int h = (int)(it % H); const int buf = (int)(k & 1);
    if (h == 0) mbar_wait(&s.bars->h_full[buf], (uint32_t)(k >> 1) & 1);
    const uint32_t hb = smem_u32(s.h_s + (size_t)buf * p.hbuf_bytes);
    const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    for (int t = 0; t < 3; ++t) {
        const int slot = (int)(wc % p.NR);
        mbar_wait(&s.bars->w_full[slot], (uint32_t)(wc / p.NR) & 1);
        tc_fence_after();
        const uint32_t wb = smem_u32(s.w_s + (size_t)slot * p.slot_bytes);
        for (int ks = 0; ks < 2 * p.nch; ++ks)
            umma_bf16(s.tmem + COL_P + 64 * t, make_smem_desc_sw64(hb + (ks >> 1) * 8192 + (ks & 1) * 32),
                      make_smem_desc_sw64(wb + (ks >> 1) * 4096 + (ks & 1) * 32), idesc, ks != 0);
        umma_commit(&s.bars->w_empty[slot]);
        ++wc;
    }
    if (h == H - 1) umma_commit(&s.bars->h_empty[buf]);
    umma_commit(&s.bars->pro_full);
}
// compute threads: Q, K, V accumulators (fp32, TMEM lane L = tile row L) -> bf16 smem tiles
__device__ __forceinline__ void convert_qkv(const AbCommon& s, uint32_t lane_addr, int L, int cq) {
    uint32_t a[16], b[16], c[16];
    tmem_ld_32x16(s.tmem + lane_addr + COL_P + 16 * cq, a);
    tmem_ld_32x16(s.tmem + lane_addr + COL_P + 64 + 16 * cq, b);
    tmem_ld_32x16(s.tmem + lane_addr + COL_P + 128 + 16 * cq, c);
    tmem_ld_wait();
    store_row16(s.q_s, L, cq, a);
    store_row16(s.k_s, L, cq, b);
    store_row16(s.v_s, L, cq, c);
}

__device__ __forceinline__ void ab_setup(AbBars* bars, int warp, uint8_t* zero_from, uint32_t zero_bytes, const CUtensorMap* m0, const CUtensorMap* m1,
                                         const CUtensorMap* m2, const CUtensorMap* m3) {
    if (warp == 17 && elect_one()) {
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->h_full[i], 1); mbar_init(&bars->h_empty[i], 1); mbar_init(&bars->do_full[i], 1); mbar_init(&bars->do_empty[i], 1); }
        for (int i = 0; i < AB_MAXR; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], 1); }
        mbar_init(&bars->pro_full, 1); mbar_init(&bars->conv_done, 512); mbar_init(&bars->s_full, 1); mbar_init(&bars->p_full, 512);
        mbar_init(&bars->o_full, 1); mbar_init(&bars->a_ready, 512); mbar_init(&bars->dh_full, 1); mbar_init(&bars->dh_free, 512);
        fence_barrier_init();
    }
    if (warp == 16) {
        tmem_alloc(&bars->tmem_base, 512);
        if (elect_one()) { prefetch_tmap(m0); prefetch_tmap(m1); if (m2) prefetch_tmap(m2); if (m3) prefetch_tmap(m3); }
    }
    // zero once: rows a slot group's box never writes (slots G*N .. 63) must stay finite (0 x NaN would poison the contractions)
    for (uint32_t i = threadIdx.x; i < zero_bytes / 16; i += AB_THREADS) reinterpret_cast<uint4*>(zero_from)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
}

// =========================================================================================================
// forward
// =========================================================================================================
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_block_fwd_kernel(const __grid_constant__ CUtensorMap tma_h, const __grid_constant__ CUtensorMap tma_w, const AbParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (smem_u32(smem) & 1023u) __trap();
    const AttnGeom& g = p.g;
    AbCommon s;
    s.h_s = smem; s.w_s = s.h_s + 2 * p.hbuf_bytes; s.q_s = s.w_s + (size_t)p.NR * p.slot_bytes; s.k_s = s.q_s + AB_T16; s.v_s = s.k_s + AB_T16;
    uint8_t* p_s = s.v_s + AB_T16;
    float* xch = reinterpret_cast<float*>(p_s + AB_T16);          // [max | sum][4 column quarters][128 lanes]
    s.bars = reinterpret_cast<AbBars*>(xch + 2 * 4 * 128);
    AbBars* bars = s.bars;
    const int warp = threadIdx.x >> 5;
    const int H = g.H, I = H * 64;
    ab_setup(bars, warp, s.h_s, 2 * p.hbuf_bytes, &tma_h, &tma_w, nullptr, nullptr);
    s.tmem = bars->tmem_base;

    const int64_t my_tiles = blockIdx.x < p.n_tiles ? (p.n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int64_t n_items = my_tiles * H;

    if (warp == 16) {
        // ===== TMA producer =====
        if (elect_one()) {
            int64_t wc = 0;
            for (int64_t it = 0; it < n_items; ++it) {
                const int64_t k = it / H; const int h = (int)(it % H);
                if (h == 0) {
                    const int buf = (int)(k & 1);
                    mbar_wait(&bars->h_empty[buf], ((uint32_t)(k >> 1) & 1) ^ 1);
                    load_h_tile(p, &tma_h, s.h_s + (size_t)buf * p.hbuf_bytes, &bars->h_full[buf], blockIdx.x + k * gridDim.x);
                }
                for (int t = 0; t < 3; ++t) load_w_slice(p, &tma_w, bars, s.w_s, wc, t, h);
            }
        }
    } else if (warp == 17) {
        // ===== MMA issuer =====
        if (elect_one() && n_items > 0) {
            int64_t wc = 0;
            const uint32_t idesc_s = make_idesc_bf16(64, 64, 0, 0), idesc_o = make_idesc_bf16(64, 64, 0, 1);
            const uint32_t qb = smem_u32(s.q_s), kb = smem_u32(s.k_s), vb = smem_u32(s.v_s), pb = smem_u32(p_s);
            issue_prologue(p, s, 0, wc);
            for (int64_t it = 0; it < n_items; ++it) {
                const uint32_t ph = (uint32_t)it & 1;
                mbar_wait(&bars->conv_done, ph);                  // Q, K, V tiles written (generic proxy + the writers' fence)
                fence_proxy_async();
                tc_fence_after();
                for (int b = 0; b < 2; ++b)
                    for (int ks = 0; ks < 4; ++ks)                // S_b = Q_b K_b^T
                        umma_bf16(s.tmem + COL_S + ((uint32_t)(16 * b) << 16), make_smem_desc(qb + b * AB_BLK, 16, 1024) + (uint64_t)(ks * 2),
                                  make_smem_desc(kb + b * AB_BLK, 16, 1024) + (uint64_t)(ks * 2), idesc_s, ks != 0);
                umma_commit(&bars->s_full);
                if (it + 1 < n_items) issue_prologue(p, s, it + 1, wc);   // the accumulators of item it were read before conv_done
                mbar_wait(&bars->p_full, ph);
                fence_proxy_async();
                tc_fence_after();
                for (int b = 0; b < 2; ++b)
                    for (int ks = 0; ks < 4; ++ks)                // O_b = P~_b V_b (V as MN-major B)
                        umma_bf16(s.tmem + COL_O + ((uint32_t)(16 * b) << 16), make_smem_desc(pb + b * AB_BLK, 16, 1024) + (uint64_t)(ks * 2),
                                  make_smem_desc(vb + b * AB_BLK, AB_BLK, 1024) + (uint64_t)(ks * 128), idesc_o, ks != 0);
                umma_commit(&bars->o_full);
            }
        }
    } else {
        // ===== 512 compute threads: thread = (TMEM lane L, column quarter cq) =====
        const int lq = warp & 3, cq = warp >> 2, lane = threadIdx.x & 31;
        const int L = lq * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;
        const int slot = 64 * ((lane >> 4) & 1) + 16 * lq + (lane & 15);      // tile row held by lane L of an interleaved M = 64 pair
        const int blk = slot >> 6, m = slot & 63;
        const float sl2 = g.scale * 1.4426950408889634f;
        const bool full_blocks = g.N == 64 && g.groups % 2 == 0;
        const uint64_t seed = p.drop.seed + (p.drop.seed_dev ? __ldg(p.drop.seed_dev) : 0ull);
        const uint32_t t16 = p.drop.thresh >> 16;
        const uint32_t vm_geom = full_blocks ? 0xFFFFu : key_mask16(g, m, cq);
        float* xmax = xch; float* xsum = xch + 4 * 128;
        int64_t cur_tile = -1, grow = -1; uint32_t vm = 0;
        for (int64_t it = 0; it < n_items; ++it) {
            const uint32_t ph = (uint32_t)it & 1;
            const int64_t tile = blockIdx.x + (it / H) * gridDim.x; const int h = (int)(it % H);
            if (tile != cur_tile) {
                cur_tile = tile;
                int64_t seq; int pos;
                const int64_t group = tile * 2 + blk;
                const bool ok = group < g.groups && slot_to(g, group, 0, m, seq, pos);
                grow = ok ? row_of(g, seq, pos) : -1;
                vm = ok ? vm_geom : 0u;
            }
            // ---- Q, K, V -> bf16 tiles ----
            mbar_wait(&bars->pro_full, ph);
            tc_fence_after();
            convert_qkv(s, lane_addr, L, cq);
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(&bars->conv_done);
            // ---- softmax of row `slot`, key columns 16*cq .. +15 of its block ----
            const uint64_t hidx = pair_base(g, tile * 2 + blk, h) + (uint64_t)(m * 32 + 8 * cq);
            const uint32_t hash_lo = (uint32_t)hidx + (uint32_t)(seed >> 32);
            const uint32_t hash_hi = ((uint32_t)(hidx >> 32) * 0x85EBCA77u) ^ (uint32_t)seed ^ (p.drop.site * 0xC2B2AE3Du);
            mbar_wait(&bars->s_full, ph);
            tc_fence_after();
            float sc[16];
            {
                uint32_t a[16];
                tmem_ld_32x16(s.tmem + lane_addr + COL_S + 16 * cq, a);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) sc[j] = __uint_as_float(a[j]) * sl2;
            }
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 16; ++j) { sc[j] = (vm >> j) & 1u ? sc[j] : -INFINITY; mx = fmaxf(mx, sc[j]); }
            xmax[cq * 128 + L] = mx;
            asm volatile("bar.sync 1, 512;" ::: "memory");
            mx = fmaxf(fmaxf(xmax[L], xmax[128 + L]), fmaxf(xmax[256 + L], xmax[384 + L]));
            const float sub = mx == -INFINITY ? 0.f : mx;
            float l = 0.f;
            uint32_t pk[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                float p0 = ex2_approx(sc[2 * jj] - sub), p1 = ex2_approx(sc[2 * jj + 1] - sub);
                l += p0 + p1;
                if (p.drop.on()) {
                    uint32_t x = (hash_lo + (uint32_t)jj) * 0x9E3779B1u ^ hash_hi;
                    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
                    p0 *= (x & 0xFFFFu) >= t16 ? p.drop.scale : 0.f;
                    p1 *= (x >> 16) >= t16 ? p.drop.scale : 0.f;
                }
                pk[jj] = pack_bf(p0, p1);
            }
            xsum[cq * 128 + L] = l;
            store_row16_packed(p_s + blk * AB_BLK, m, cq, pk);      // the previous item's O MMAs (its readers) completed before this item's S
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&bars->p_full);
            asm volatile("bar.sync 1, 512;" ::: "memory");
            l = (xsum[L] + xsum[128 + L]) + (xsum[256 + L] + xsum[384 + L]);
            const float inv = l > 0.f ? 1.f / l : 0.f;
            if (cq == 0 && grow >= 0) p.lse_out[grow * H + h] = (sub + log2f(l)) * 0.6931471805599453f;
            // ---- epilogue: O row / l -> bf16 -> global ----
            mbar_wait(&bars->o_full, ph);
            tc_fence_after();
            {
                uint32_t v[16];
                tmem_ld_32x16(s.tmem + lane_addr + COL_O + 16 * cq, v);
                tmem_ld_wait();
                if (grow >= 0) {
                    uint4* dst = reinterpret_cast<uint4*>(p.o + grow * I + h * 64 + 16 * cq);
                    dst[0] = make_uint4(pack_bf(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv), pack_bf(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv),
                                        pack_bf(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv), pack_bf(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv));
                    dst[1] = make_uint4(pack_bf(__uint_as_float(v[8]) * inv, __uint_as_float(v[9]) * inv), pack_bf(__uint_as_float(v[10]) * inv, __uint_as_float(v[11]) * inv),
                                        pack_bf(__uint_as_float(v[12]) * inv, __uint_as_float(v[13]) * inv), pack_bf(__uint_as_float(v[14]) * inv, __uint_as_float(v[15]) * inv));
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) { tc_fence_after(); tmem_dealloc(s.tmem, 512); }
}

// =========================================================================================================
// backward
// =========================================================================================================
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_block_bwd_kernel(const __grid_constant__ CUtensorMap tma_h, const __grid_constant__ CUtensorMap tma_w, const __grid_constant__ CUtensorMap tma_wt,
                      const __grid_constant__ CUtensorMap tma_do, const AbParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (smem_u32(smem) & 1023u) __trap();
    const AttnGeom& g = p.g;
    AbCommon s;
    s.h_s = smem; s.w_s = s.h_s + 2 * p.hbuf_bytes; s.q_s = s.w_s + (size_t)p.NR * p.slot_bytes; s.k_s = s.q_s + AB_T16; s.v_s = s.k_s + AB_T16;
    uint8_t* do_s = s.v_s + AB_T16;                               // [2][16 KB]
    uint8_t* p_s = do_s + 2 * AB_T16;                             // P~ : two [64 q][64 k] blocks
    uint8_t* ds_s = p_s + AB_T16;                                 // dS
    float* xch = reinterpret_cast<float*>(ds_s + AB_T16);         // [4 column quarters][128 lanes] partial row sums
    s.bars = reinterpret_cast<AbBars*>(xch + 4 * 128);
    AbBars* bars = s.bars;
    const int warp = threadIdx.x >> 5;
    const int H = g.H, I = H * 64, D = p.D;
    // zero: both h buffers ... and the dO stages (rows no box covers).  They are not adjacent: zero h here, dO below.
    ab_setup(bars, warp, s.h_s, 2 * p.hbuf_bytes, &tma_h, &tma_w, &tma_wt, &tma_do);
    for (uint32_t i = threadIdx.x; i < 2 * AB_T16 / 16; i += AB_THREADS) reinterpret_cast<uint4*>(do_s)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    __syncthreads();
    s.tmem = bars->tmem_base;

    const int64_t my_tiles = blockIdx.x < p.n_tiles ? (p.n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int64_t n_items = my_tiles * H;
    const uint32_t do_rows = p.nbox == 1 ? 128u : (uint32_t)(g.G * g.N);

    if (warp == 16) {
        // ===== TMA producer.  Ring order = consumption order: prologue(0), then per item: prologue(it + 1), data-gradient slices(it) =====
        if (elect_one()) {
            int64_t wc = 0;
            auto load_do = [&](int64_t it) {
                const int st = (int)(it & 1);
                const int64_t tile = blockIdx.x + (it / H) * gridDim.x; const int h = (int)(it % H);
                mbar_wait(&bars->do_empty[st], ((uint32_t)(it >> 1) & 1) ^ 1);
                uint8_t* dst = do_s + (size_t)st * AB_T16;
                mbar_arrive_expect_tx(&bars->do_full[st], (uint32_t)p.nbox * do_rows * 128u);
                if (p.nbox == 1) tma_load_2d(dst, &tma_do, &bars->do_full[st], h * 64, (int)(tile * 128));
                else
                    for (int gi = 0; gi < 2; ++gi) {
                        int c1, c2, c3;
                        group_coords(g, tile * 2 + gi, c1, c2, c3);
                        tma_load_4d(dst + gi * AB_BLK, &tma_do, &bars->do_full[st], h * 64, c1, c2, c3);
                    }
            };
            auto load_pro = [&](int64_t it) {
                const int64_t k = it / H; const int h = (int)(it % H);
                if (h == 0) {
                    const int buf = (int)(k & 1);
                    mbar_wait(&bars->h_empty[buf], ((uint32_t)(k >> 1) & 1) ^ 1);
                    load_h_tile(p, &tma_h, s.h_s + (size_t)buf * p.hbuf_bytes, &bars->h_full[buf], blockIdx.x + k * gridDim.x);
                }
                for (int t = 0; t < 3; ++t) load_w_slice(p, &tma_w, bars, s.w_s, wc, t, h);
            };
            if (n_items > 0) { load_pro(0); load_do(0); }
            for (int64_t it = 0; it < n_items; ++it) {
                if (it + 1 < n_items) { load_pro(it + 1); load_do(it + 1); }
                const int h = (int)(it % H);
                for (int t = 0; t < 3; ++t) {                    // W^T column block [D rows][64] of (t, h): the data gradient's B operand
                    const int slot = (int)(wc % p.NR);
                    mbar_wait(&bars->w_empty[slot], (uint32_t)((wc / p.NR) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bars->w_full[slot], (uint32_t)D * 128u);
                    tma_load_2d(s.w_s + (size_t)slot * p.slot_bytes, &tma_wt, &bars->w_full[slot], t * I + h * 64, 0);
                    ++wc;
                }
            }
        }
    } else if (warp == 17) {
        // ===== MMA issuer =====
        if (elect_one() && n_items > 0) {
            int64_t wc = 0;
            const uint32_t idesc_kk = make_idesc_bf16(64, 64, 0, 0), idesc_mn = make_idesc_bf16(64, 64, 1, 1), idesc_kmn = make_idesc_bf16(64, 64, 0, 1);
            const uint32_t idesc_dh = make_idesc_bf16(128, D, 0, 0);
            const uint32_t qb = smem_u32(s.q_s), kb = smem_u32(s.k_s), vb = smem_u32(s.v_s), pb = smem_u32(p_s), dsb = smem_u32(ds_s);
            issue_prologue(p, s, 0, wc);
            for (int64_t it = 0; it < n_items; ++it) {
                const uint32_t ph = (uint32_t)it & 1;
                const int st = (int)(it & 1);
                const int h = (int)(it % H); const int64_t k = it / H;
                const uint32_t dob = smem_u32(do_s + (size_t)st * AB_T16);
                mbar_wait(&bars->conv_done, ph);
                mbar_wait(&bars->do_full[st], (uint32_t)(it >> 1) & 1);
                fence_proxy_async();
                tc_fence_after();
                for (int b = 0; b < 2; ++b) {
                    const uint32_t lane_off = (uint32_t)(16 * b) << 16;
                    for (int ks = 0; ks < 4; ++ks)                // S_b = Q_b K_b^T
                        umma_bf16(s.tmem + COL_X + lane_off, make_smem_desc(qb + b * AB_BLK, 16, 1024) + (uint64_t)(ks * 2),
                                  make_smem_desc(kb + b * AB_BLK, 16, 1024) + (uint64_t)(ks * 2), idesc_kk, ks != 0);
                    for (int ks = 0; ks < 4; ++ks)                // dP_b = dO_b V_b^T
                        umma_bf16(s.tmem + COL_X + 64 + lane_off, make_smem_desc(dob + b * AB_BLK, 16, 1024) + (uint64_t)(ks * 2),
                                  make_smem_desc(vb + b * AB_BLK, 16, 1024) + (uint64_t)(ks * 2), idesc_kk, ks != 0);
                }
                umma_commit(&bars->s_full);
                if (it + 1 < n_items) issue_prologue(p, s, it + 1, wc);
                mbar_wait(&bars->p_full, ph);                     // P~ and dS written
                fence_proxy_async();
                tc_fence_after();
                for (int b = 0; b < 2; ++b) {
                    const uint32_t lane_off = (uint32_t)(16 * b) << 16;
                    for (int ks = 0; ks < 4; ++ks)                // dV_b = P~_b^T dO_b (reduction over the block's 64 queries, 16 rows = 2 KB per step)
                        umma_bf16(s.tmem + COL_X + lane_off, make_smem_desc(pb + b * AB_BLK, AB_BLK, 1024) + (uint64_t)(ks * 128),
                                  make_smem_desc(dob + b * AB_BLK, AB_BLK, 1024) + (uint64_t)(ks * 128), idesc_mn, ks != 0);
                    for (int ks = 0; ks < 4; ++ks)                // dK_b = dS_b^T Q_b
                        umma_bf16(s.tmem + COL_X + 64 + lane_off, make_smem_desc(dsb + b * AB_BLK, AB_BLK, 1024) + (uint64_t)(ks * 128),
                                  make_smem_desc(qb + b * AB_BLK, AB_BLK, 1024) + (uint64_t)(ks * 128), idesc_mn, ks != 0);
                    for (int ks = 0; ks < 4; ++ks)                // dQ_b = dS_b K_b
                        umma_bf16(s.tmem + COL_X + 128 + lane_off, make_smem_desc(dsb + b * AB_BLK, 16, 1024) + (uint64_t)(ks * 2),
                                  make_smem_desc(kb + b * AB_BLK, AB_BLK, 1024) + (uint64_t)(ks * 128), idesc_kmn, ks != 0);
                }
                umma_commit(&bars->o_full);
                umma_commit(&bars->do_empty[st]);
                // ---- data gradient of the projection: dh (+)= [dQ | dK | dV] (bf16 rows in TMEM) x W^T blocks ----
                mbar_wait(&bars->a_ready, ph);
                if (h == 0 && k > 0) mbar_wait(&bars->dh_free, (uint32_t)(k - 1) & 1);   // the previous tile's dh rows have been read
                tc_fence_after();
                for (int t = 0; t < 3; ++t) {
                    const int slot = (int)(wc % p.NR);
                    mbar_wait(&bars->w_full[slot], (uint32_t)(wc / p.NR) & 1);
                    tc_fence_after();
                    const uint64_t bw = make_smem_desc(smem_u32(s.w_s + (size_t)slot * p.slot_bytes), 16, 1024);
                    for (int ks = 0; ks < 4; ++ks)
                        umma_bf16_ts(s.tmem + COL_DH, s.tmem + COL_X + 32 * t + 8 * ks, bw + (uint64_t)(ks * 2), idesc_dh, (h | t | ks) != 0);
                    umma_commit(&bars->w_empty[slot]);
                    ++wc;
                }
                if (h == H - 1) umma_commit(&bars->dh_full);
            }
        }
    } else {
        // ===== 512 compute threads =====
        const int lq = warp & 3, cq = warp >> 2, lane = threadIdx.x & 31;
        const int L = lq * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;
        const int slot = 64 * ((lane >> 4) & 1) + 16 * lq + (lane & 15);
        const int blk = slot >> 6, m = slot & 63;
        const float sl2 = g.scale * 1.4426950408889634f;
        const bool full_blocks = g.N == 64 && g.groups % 2 == 0;
        const uint64_t seed = p.drop.seed + (p.drop.seed_dev ? __ldg(p.drop.seed_dev) : 0ull);
        const uint32_t t16 = p.drop.thresh >> 16;
        const uint32_t vm_geom = full_blocks ? 0xFFFFu : key_mask16(g, m, cq);
        const int dq4 = D / 4;                                    // dh columns per thread
        int64_t cur_tile = -1, grow = -1; uint32_t vm = 0;
        for (int64_t it = 0; it < n_items; ++it) {
            const uint32_t ph = (uint32_t)it & 1;
            const int64_t k = it / H;
            const int64_t tile = blockIdx.x + k * gridDim.x; const int h = (int)(it % H);
            if (tile != cur_tile) {
                cur_tile = tile;
                int64_t seq; int pos;
                const int64_t group = tile * 2 + blk;
                const bool ok = group < g.groups && slot_to(g, group, 0, m, seq, pos);
                grow = ok ? row_of(g, seq, pos) : -1;
                vm = ok ? vm_geom : 0u;
            }
            // ---- recomputed Q, K, V -> bf16 tiles ----
            mbar_wait(&bars->pro_full, ph);
            tc_fence_after();
            convert_qkv(s, lane_addr, L, cq);
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(&bars->conv_done);
            // ---- softmax backward for row `slot`, key columns 16*cq .. +15 ----
            const float Lse = grow >= 0 ? __ldg(p.lse_in + grow * H + h) * 1.4426950408889634f : 0.f;
            const uint64_t hidx = pair_base(g, tile * 2 + blk, h) + (uint64_t)(m * 32 + 8 * cq);
            const uint32_t hash_lo = (uint32_t)hidx + (uint32_t)(seed >> 32);
            const uint32_t hash_hi = ((uint32_t)(hidx >> 32) * 0x85EBCA77u) ^ (uint32_t)seed ^ (p.drop.site * 0xC2B2AE3Du);
            mbar_wait(&bars->s_full, ph);
            tc_fence_after();
            float sp[16], dp[16];
            {
                uint32_t a[16], b[16];
                tmem_ld_32x16(s.tmem + lane_addr + COL_X + 16 * cq, a);
                tmem_ld_32x16(s.tmem + lane_addr + COL_X + 64 + 16 * cq, b);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) { sp[j] = __uint_as_float(a[j]); dp[j] = __uint_as_float(b[j]); }
            }
            uint32_t pk[8], dk[8];
            float Dp = 0.f;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int j = 2 * jj;
                float f0 = 1.f, f1 = 1.f;
                if (p.drop.on()) {
                    uint32_t x = (hash_lo + (uint32_t)jj) * 0x9E3779B1u ^ hash_hi;
                    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
                    f0 = (x & 0xFFFFu) >= t16 ? p.drop.scale : 0.f;
                    f1 = (x >> 16) >= t16 ? p.drop.scale : 0.f;
                }
                float p0 = ex2_approx(fmaf(sp[j], sl2, -Lse)), p1 = ex2_approx(fmaf(sp[j + 1], sl2, -Lse));
                p0 = (vm >> j) & 1u ? p0 : 0.f;
                p1 = (vm >> (j + 1)) & 1u ? p1 : 0.f;
                const float pf0 = p0 * f0, pf1 = p1 * f1;
                Dp = fmaf(pf0, dp[j], Dp); Dp = fmaf(pf1, dp[j + 1], Dp);
                pk[jj] = pack_bf(pf0, pf1);
                sp[j] = p0 * g.scale; sp[j + 1] = p1 * g.scale;
                dp[j] *= f0; dp[j + 1] *= f1;
            }
            xch[cq * 128 + L] = Dp;
            asm volatile("bar.sync 1, 512;" ::: "memory");
            Dp = (xch[L] + xch[128 + L]) + (xch[256 + L] + xch[384 + L]);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) dk[jj] = pack_bf(sp[2 * jj] * (dp[2 * jj] - Dp), sp[2 * jj + 1] * (dp[2 * jj + 1] - Dp));
            store_row16_packed(p_s + blk * AB_BLK, m, cq, pk);
            store_row16_packed(ds_s + blk * AB_BLK, m, cq, dk);
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&bars->p_full);
            // ---- epilogue: dV, dK, dQ rows of `slot` -> bf16 -> dqkv (global) and the data gradient's A operand (TMEM) ----
            mbar_wait(&bars->o_full, ph);
            tc_fence_after();
            uint32_t qv[8], kv[8], vv[8];
            {
                uint32_t a[16], b[16], c[16];
                tmem_ld_32x16(s.tmem + lane_addr + COL_X + 16 * cq, a);          // dV
                tmem_ld_32x16(s.tmem + lane_addr + COL_X + 64 + 16 * cq, b);     // dK
                tmem_ld_32x16(s.tmem + lane_addr + COL_X + 128 + 16 * cq, c);    // dQ
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    vv[j] = pack_bf(__uint_as_float(a[2 * j]), __uint_as_float(a[2 * j + 1]));
                    kv[j] = pack_bf(__uint_as_float(b[2 * j]), __uint_as_float(b[2 * j + 1]));
                    qv[j] = pack_bf(__uint_as_float(c[2 * j]), __uint_as_float(c[2 * j + 1]));
                }
            }
            if (grow >= 0) {
                bf16* base = p.dqkv + grow * (int64_t)(3 * I) + h * 64 + 16 * cq;
                uint4* dq_ = reinterpret_cast<uint4*>(base);
                uint4* dk_ = reinterpret_cast<uint4*>(base + I);
                uint4* dv_ = reinterpret_cast<uint4*>(base + 2 * I);
                dq_[0] = make_uint4(qv[0], qv[1], qv[2], qv[3]); dq_[1] = make_uint4(qv[4], qv[5], qv[6], qv[7]);
                dk_[0] = make_uint4(kv[0], kv[1], kv[2], kv[3]); dk_[1] = make_uint4(kv[4], kv[5], kv[6], kv[7]);
                dv_[0] = make_uint4(vv[0], vv[1], vv[2], vv[3]); dv_[1] = make_uint4(vv[4], vv[5], vv[6], vv[7]);
            }
            // every thread has read its accumulator columns: the bf16 rows may overwrite X (lane L keeps holding row `slot`)
            tc_fence_before();
            asm volatile("bar.sync 2, 512;" ::: "memory");
            tc_fence_after();
            tmem_st_32x8(s.tmem + lane_addr + COL_X + 8 * cq, qv);
            tmem_st_32x8(s.tmem + lane_addr + COL_X + 32 + 8 * cq, kv);
            tmem_st_32x8(s.tmem + lane_addr + COL_X + 64 + 8 * cq, vv);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars->a_ready);
            if (h == H - 1) {
                // ---- the tile's dh rows: accumulated over the heads -> fp32 -> global ----
                mbar_wait(&bars->dh_full, (uint32_t)k & 1);
                tc_fence_after();
                for (int c = 0; c < p.nch; ++c) {
                    uint32_t v[8];
                    tmem_ld_32x8(s.tmem + lane_addr + COL_DH + cq * dq4 + 8 * c, v);
                    tmem_ld_wait();
                    if (grow >= 0) {
                        float4* dst = reinterpret_cast<float4*>(p.dh + grow * D + cq * dq4 + 8 * c);
                        dst[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
                        dst[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
                    }
                }
                tc_fence_before();
                mbar_arrive(&bars->dh_free);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) { tc_fence_after(); tmem_dealloc(s.tmem, 512); }
}

// ---- host side ----
// view of an activation matrix [R, cols] whose box is one slot group (nbox == 2) or 128 consecutive rows (nbox == 1)
int act_tmap(CUtensorMap* m, const AttnGeom& g, const bf16* base, int64_t cols, int nbox, int box_cols, int swizzle) {
    if (nbox == 1) {
        const int64_t dims[2] = {cols, g.n_seq * g.N}, strides[1] = {cols};
        const int box[2] = {box_cols, 128};
        return make_tmap_bf16_nd(m, base, 2, dims, strides, box, swizzle);
    }
    if (g.gpb == 0) {
        const int64_t dims[4] = {cols, g.N, g.n_seq, 1}, strides[3] = {cols, (int64_t)g.N * cols, g.n_seq * g.N * cols};
        const int box[4] = {box_cols, g.N, g.G, 1};
        return make_tmap_bf16_nd(m, base, 4, dims, strides, box, swizzle);
    }
    const int64_t dims[4] = {cols, g.inner, g.N, g.n_seq / g.inner};
    const int64_t strides[3] = {cols, (int64_t)g.inner * cols, (int64_t)g.N * g.inner * cols};
    const int box[4] = {box_cols, g.G, g.N, 1};
    return make_tmap_bf16_nd(m, base, 4, dims, strides, box, swizzle);
}

int fill_params(AbParams& p, const AttnGeom& g, int D, bool bwd, size_t& smem_bytes) {
    p.g = g; p.D = D; p.nch = D / 32;
    p.nbox = (g.gpb == 0 && g.G * g.N == 64) ? 1 : 2;
    p.n_tiles = (g.groups + 1) / 2;
    p.slot_bytes = (uint32_t)D * 128u;
    p.hbuf_bytes = (uint32_t)p.nch * 8192u;
    const size_t fixed = 2 * (size_t)p.hbuf_bytes + 3 * AB_T16 + (bwd ? 4 * AB_T16 + 4 * 128 * sizeof(float) : AB_T16 + 2 * 4 * 128 * sizeof(float)) + sizeof(AbBars);
    const size_t cap = 232448;   // 227 KB of dynamic shared memory per CTA
    int nr = AB_MAXR;
    while (nr > 3 && fixed + (size_t)nr * p.slot_bytes > cap) --nr;
    MSST_REQUIRE(fixed + (size_t)nr * p.slot_bytes <= cap, "attn_block: D = %d does not fit shared memory", D);
    p.NR = nr;
    smem_bytes = fixed + (size_t)nr * p.slot_bytes;
    return MSST_OK;
}

}  // namespace

bool attn_block_supported(const AttnGeom& g, int D) {
    return attention_bwd_tc_supported(g) && g.dh == 64 && D % 32 == 0 && D >= 32 && D <= 128 && g.H * 64 * 3 < 65536;
}

int attn_block_fwd(const AttnGeom& g, int D, const bf16* h, const bf16* w_qkv, bf16* out, float* lse, Drop drop, cudaStream_t st) {
    MSST_REQUIRE(attn_block_supported(g, D), "attn_block_fwd: needs packed short sequences (N <= 64), dim_head 64, D in {32, 64, 96, 128}");
    AbParams p{};
    size_t smem = 0;
    if (int rc = fill_params(p, g, D, false, smem)) return rc;
    p.lse_out = lse; p.o = out; p.drop = drop;
    static PerDeviceOnce once;
    if (once.first()) MSST_CUDA(cudaFuncSetAttribute(attn_block_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    CUtensorMap t_h, t_w;
    if (int rc = act_tmap(&t_h, g, h, D, p.nbox, 32, 64)) return rc;
    const int64_t I3 = (int64_t)g.H * 64 * 3;
    { const int64_t dims[2] = {D, I3}, strides[1] = {D}; const int box[2] = {32, 64};
      if (int rc = make_tmap_bf16_nd(&t_w, w_qkv, 2, dims, strides, box, 64)) return rc; }
    const int grid = (int)(p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs);
    attn_block_fwd_kernel<<<grid, AB_THREADS, smem, st>>>(t_h, t_w, p);
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}

int attn_block_bwd(const AttnGeom& g, int D, const bf16* h, const bf16* w_qkv, const bf16* w_qkv_t, const bf16* d_out, const float* lse,
                   bf16* d_qkv, float* d_h, Drop drop, cudaStream_t st) {
    MSST_REQUIRE(attn_block_supported(g, D), "attn_block_bwd: needs packed short sequences (N <= 64), dim_head 64, D in {32, 64, 96, 128}");
    AbParams p{};
    size_t smem = 0;
    if (int rc = fill_params(p, g, D, true, smem)) return rc;
    p.lse_in = lse; p.dqkv = d_qkv; p.dh = d_h; p.drop = drop;
    static PerDeviceOnce once;
    if (once.first()) MSST_CUDA(cudaFuncSetAttribute(attn_block_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    CUtensorMap t_h, t_w, t_wt, t_do;
    const int64_t I = (int64_t)g.H * 64, I3 = 3 * I;
    if (int rc = act_tmap(&t_h, g, h, D, p.nbox, 32, 64)) return rc;
    if (int rc = act_tmap(&t_do, g, d_out, I, p.nbox, 64, 128)) return rc;
    { const int64_t dims[2] = {D, I3}, strides[1] = {D}; const int box[2] = {32, 64};
      if (int rc = make_tmap_bf16_nd(&t_w, w_qkv, 2, dims, strides, box, 64)) return rc; }
    { const int64_t dims[2] = {I3, D}, strides[1] = {I3}; const int box[2] = {64, D};
      if (int rc = make_tmap_bf16_nd(&t_wt, w_qkv_t, 2, dims, strides, box, 128)) return rc; }
    const int grid = (int)(p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs);
    attn_block_bwd_kernel<<<grid, AB_THREADS, smem, st>>>(t_h, t_w, t_wt, t_do, p);
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}

}  // namespace msst
