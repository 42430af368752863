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
#include <stdio.h>
#include <type_traits>

namespace msst {
using namespace ptx;
typedef __nv_bfloat16 bf16;
int make_tmap_bf16_nd(CUtensorMap* m, const void* base, int rank, const int64_t* dims, const int64_t* strides, const int* box, int swizzle_bytes);   // gemm_bf16.cu
int make_tmap_nd(CUtensorMap* m, const void* base, int elem_bytes, int rank, const int64_t* dims, const int64_t* strides, const int* box, int swizzle_bytes);

namespace {

constexpr int AB_THREADS = 608;            // warps 0-15 compute, 16 TMA producer (+ TMEM alloc), 17 MMA issuer, 18 TMA store
constexpr int AB_MAXR = 8;                 // weight ring slots (upper bound)
constexpr uint32_t AB_T16 = 16384;         // one [128 rows][64 bf16] SWIZZLE_128B tile
constexpr uint32_t AB_BLK = 8192;          // one 64-row block of such a tile
constexpr uint32_t COL_P = 0, COL_S = 192, COL_O = 256, COL_X = 192, COL_DH = 384;
constexpr uint32_t COL_Y = 320, Y_COLS = 96;   // forward with the fused out-projection: two Y accumulators (tile parity), D <= 96 columns each

struct AbParams {
    AttnGeom g;
    int D, nch;                // model dim, number of 32-column chunks
    int nbox;                  // 1: a tile is 128 consecutive rows (2-D boxes); 2: one 4-D box per 64-slot group
    int NR;                    // weight ring slots in use
    int hbufs;                 // h tile buffers (2 forward, 1 backward)
    int64_t n_tiles;
    uint32_t slot_bytes;       // D * 128: one weight sub-slice ([64 rows][D] as chunks, or [D rows][64])
    uint32_t hbuf_bytes;       // nch * 8192: one h tile
    const float* lse_in; float* lse_out;
    bf16* o; bf16* dqkv; float* dh;
    Drop drop;
    // forward, fused out-projection (AttnBlockOut): xmid = x + drop(o Wo^T + b_out), h2 = LN2(xmid), ln_stats = (mean, rstd)
    const float *x_res, *b_out, *ln_w, *ln_b;
    float* ln_stats;
    Drop drop_out;
    int tail_dbg;              // MSST_AB_TAILDBG (profiling only): 1 = no out-projection MMAs, 2 = no phase arithmetic
    long long* dbg;            // MSST_AB_DBG: clock64 timeline of CTA 0 ([item][16] slots)
};
#define AB_T(it, slot) do { if (p.dbg && blockIdx.x == 0 && (it) < 48) p.dbg[(it) * 16 + (slot)] = clock64(); } while (0)

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
    uint64_t pro_full, conv_done, s_full, sdp_read, p_full, o_full, a_ready, dh_full, dh_free, stg_full, stg_free, dhs_full, dhs_free;
    uint64_t y_full[2], out_full, out_free, xin_full, t_kfree[2], t_vfree[2], kv_free[2];   // forward with the fused out-projection (pairs: one per tile / item parity)
    uint32_t tmem_base;
};

// ---- pieces shared by the forward and the backward kernel ----
struct AbCommon {
    uint8_t *h_s, *w_s;
    AbBars* bars;
    uint32_t tmem, h_addr, w_addr;
};
// all lanes of the warp have finished (and fenced) their part: one elected arrival for the warp
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}
struct Ring {   // weight ring position of one agent (producer or MMA issuer): slot + phase parity, no div / mod on the hot path
    int slot; uint32_t ph; int NR;
    __device__ __forceinline__ void next() { if (++slot == NR) { slot = 0; ph ^= 1u; } }
};
constexpr uint64_t kDescHiK128 = ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (kLayoutSW128 << 61);
// K-major SWIZZLE_128B tile (8-row atoms of 1 KB); + 2 per K = 16 step
__device__ __forceinline__ uint64_t kdesc(uint32_t addr) { return kDescHiK128 | (1ull << 16) | (uint64_t)((addr >> 4) & 0x3FFF); }
// MN-major SWIZZLE_128B tile, 64 MN elements wide (one atom): + 128 per K = 16 step (16 rows = 2 KB)
__device__ __forceinline__ uint64_t mndesc(uint32_t addr) { return kDescHiK128 | ((uint64_t)(AB_BLK >> 4) << 16) | (uint64_t)((addr >> 4) & 0x3FFF); }

// producer: the h rows of tile `tile` -> buffer dst (NCH chunks of [128 rows][32 cols], SWIZZLE_64B)
template <int NCH>
__device__ __forceinline__ void load_h_tile(const AbParams& p, const CUtensorMap* tma_h, uint8_t* dst, uint64_t* bar, int64_t tile) {
    const AttnGeom& g = p.g;
    const uint32_t rows = p.nbox == 1 ? 128u : (uint32_t)(g.G * g.N);
    mbar_arrive_expect_tx(bar, (uint32_t)NCH * (uint32_t)p.nbox * rows * 64u);
    if (p.nbox == 1) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) tma_load_2d(dst + c * 8192, tma_h, bar, c * 32, (int)(tile * 128));
    } else {
        for (int gi = 0; gi < 2; ++gi) {
            int c1, c2, c3;
            group_coords(g, tile * 2 + gi, c1, c2, c3);
#pragma unroll
            for (int c = 0; c < NCH; ++c) tma_load_4d(dst + c * 8192 + gi * 4096, tma_h, bar, c * 32, c1, c2, c3);
        }
    }
}
template <int NCH>
__device__ __forceinline__ void prefetch_h_tile(const AbParams& p, const CUtensorMap* tma_h, int64_t tile) {
    if (p.nbox == 1) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) tma_prefetch_l2_2d(tma_h, c * 32, (int)(tile * 128));
    } else {
        for (int gi = 0; gi < 2; ++gi) {
            int c1, c2, c3;
            group_coords(p.g, tile * 2 + gi, c1, c2, c3);
#pragma unroll
            for (int c = 0; c < NCH; ++c) tma_prefetch_l2_4d(tma_h, c * 32, c1, c2, c3);
        }
    }
}
// producer: W rows [t*I + h*64, +64) x D columns as NCH SWIZZLE_64B chunks -> ring slot
template <int NCH>
__device__ __forceinline__ void load_w_slice(const AbParams& p, const CUtensorMap* tma_w, AbBars* bars, uint8_t* w_s, Ring& r, int t, int h) {
    mbar_wait(&bars->w_empty[r.slot], r.ph ^ 1u);
    mbar_arrive_expect_tx(&bars->w_full[r.slot], (uint32_t)NCH * 4096u);
    uint8_t* dst = w_s + (size_t)r.slot * p.slot_bytes;
#pragma unroll
    for (int c = 0; c < NCH; ++c) tma_load_2d(dst + c * 4096, tma_w, &bars->w_full[r.slot], c * 32, t * p.g.H * 64 + h * 64);
    r.next();
}
// MMA issuer: [Q|K|V] of the item (tile counter k, head h) = h tile x the head's three weight sub-slices -> TMEM columns [0, 192)
template <int NCH>
__device__ __forceinline__ void issue_prologue(const AbParams& p, const AbCommon& s, int64_t k, int h, Ring& r, bool lead) {
    const int buf = (int)(k & 1);
    if (h == 0) mbar_wait(&s.bars->h_full[buf], (uint32_t)(k >> 1) & 1);
    const uint64_t ha = make_smem_desc_sw64(s.h_addr + (uint32_t)buf * p.hbuf_bytes);
    const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        mbar_wait(&s.bars->w_full[r.slot], r.ph);
        tc_fence_after();
        const uint64_t wa = make_smem_desc_sw64(s.w_addr + (uint32_t)r.slot * p.slot_bytes);
#pragma unroll
        for (int ks = 0; ks < 2 * NCH; ++ks)
            if (lead) umma_bf16(s.tmem + COL_P + 64 * t, ha + (uint64_t)(((ks >> 1) * 8192 + (ks & 1) * 32) >> 4), wa + (uint64_t)(((ks >> 1) * 4096 + (ks & 1) * 32) >> 4),
                      idesc, ks != 0);
        if (lead) umma_commit(&s.bars->w_empty[r.slot]);
        r.next();
    }
    if (lead && h == p.g.H - 1) umma_commit(&s.bars->h_empty[buf]);
    if (lead) umma_commit(&s.bars->pro_full);
}
// backward variant: ONE weight layout serves both uses.  The ring holds W^T column blocks [D rows][64 features] (SWIZZLE_128B) of
// (t, h): here they are the MN-major B operand of the recomputation (N = 64 features contiguous, K = D rows, 16 rows = 2 KB per
// step); the data gradient reads the same blocks as a K-major B operand.  The blocks are fetched TWICE (L2 hits) rather than held
// from the recomputation of an item to its data gradient two phases later: holding them pins 6 of the 7 ring slots and the next
// item's blocks then arrive ~2 k clks late (measured); fetched per use, every load has a whole item of latency budget.
// Single h buffer: tile k + 1 is loaded while the last items of tile k (whose recomputation was issued earlier) still run.
template <int NCH>
__device__ __forceinline__ void issue_prologue_bwd(const AbParams& p, const AbCommon& s, int64_t k, int h, Ring& r, bool lead) {
    if (h == 0) mbar_wait(&s.bars->h_full[0], (uint32_t)k & 1);
    const uint64_t ha = make_smem_desc_sw64(s.h_addr);
    const uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        mbar_wait(&s.bars->w_full[r.slot], r.ph);
        tc_fence_after();
        const uint64_t wa = mndesc(s.w_addr + (uint32_t)r.slot * p.slot_bytes);
#pragma unroll
        for (int ks = 0; ks < 2 * NCH; ++ks)
            if (lead) umma_bf16(s.tmem + COL_P + 64 * t, ha + (uint64_t)(((ks >> 1) * 8192 + (ks & 1) * 32) >> 4), wa + (uint64_t)(ks * 128), idesc, ks != 0);
        if (lead) umma_commit(&s.bars->w_empty[r.slot]);
        r.next();
    }
    if (lead && h == p.g.H - 1) umma_commit(&s.bars->h_empty[0]);
    if (lead) umma_commit(&s.bars->pro_full);
}
// compute threads: Q, K, V accumulators (fp32, TMEM lane L = tile row L) -> bf16 smem tiles at qkv (Q | K | V, 16 KB each)
__device__ __forceinline__ void convert_qkv(uint32_t tmem, uint8_t* qkv, uint32_t lane_addr, int L, int cq) {
    uint32_t a[16], b[16], c[16];
    tmem_ld_32x16(tmem + lane_addr + COL_P + 16 * cq, a);
    tmem_ld_32x16(tmem + lane_addr + COL_P + 64 + 16 * cq, b);
    tmem_ld_32x16(tmem + lane_addr + COL_P + 128 + 16 * cq, c);
    tmem_ld_wait();
    store_row16(qkv, L, cq, a);
    store_row16(qkv + AB_T16, L, cq, b);
    store_row16(qkv + 2 * AB_T16, L, cq, c);
}

__device__ __forceinline__ void ab_setup(AbBars* bars, int warp, uint8_t* z0, uint32_t z0_bytes, uint8_t* z1, uint32_t z1_bytes, const CUtensorMap* m0,
                                         const CUtensorMap* m1, const CUtensorMap* m2, const CUtensorMap* m3, const CUtensorMap* m4,
                                         uint32_t stg_free_count = 1) {
    if (warp == 17 && elect_one()) {
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->h_full[i], 1); mbar_init(&bars->h_empty[i], 1); mbar_init(&bars->do_full[i], 1); mbar_init(&bars->do_empty[i], 1); }
        for (int i = 0; i < AB_MAXR; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], 1); }
        // barriers the 16 compute warps arrive on: ONE arrival per warp (elected lane after __syncwarp) -- 512 arrivals on one
        // shared-memory word serialise (~2 clks each) and that cost sat on the critical path four times per item
        mbar_init(&bars->pro_full, 1); mbar_init(&bars->conv_done, 16); mbar_init(&bars->s_full, 1); mbar_init(&bars->p_full, 16);
        mbar_init(&bars->o_full, 1); mbar_init(&bars->a_ready, 16); mbar_init(&bars->dh_full, 1); mbar_init(&bars->dh_free, 16);
        mbar_init(&bars->stg_full, 16); mbar_init(&bars->stg_free, stg_free_count); mbar_init(&bars->sdp_read, 16);
        mbar_init(&bars->dhs_full, 16); mbar_init(&bars->dhs_free, 1);
        mbar_init(&bars->y_full[0], 1); mbar_init(&bars->y_full[1], 1);
        mbar_init(&bars->out_full, 4); mbar_init(&bars->out_free, 1); mbar_init(&bars->xin_full, 1);
        mbar_init(&bars->t_kfree[0], 1); mbar_init(&bars->t_kfree[1], 1); mbar_init(&bars->t_vfree[0], 1); mbar_init(&bars->t_vfree[1], 1); mbar_init(&bars->kv_free[0], 1); mbar_init(&bars->kv_free[1], 1);
        fence_barrier_init();
    }
    if (warp == 16) {
        tmem_alloc(&bars->tmem_base, 512);
        if (elect_one()) { prefetch_tmap(m0); prefetch_tmap(m1); prefetch_tmap(m2); if (m3) prefetch_tmap(m3); if (m4) prefetch_tmap(m4); }
    }
    // zero once: rows a slot group's box never writes (slots G*N .. 63) must stay finite (0 x NaN would poison the contractions)
    for (uint32_t i = threadIdx.x; i < z0_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(z0)[i] = make_uint4(0, 0, 0, 0);
    for (uint32_t i = threadIdx.x; i < z1_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(z1)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
}
// store warp: one staged tile of 128 rows (row pitch `pitch` bytes) -> columns col0.. of the tile's rows
__device__ __forceinline__ void store_tile_rows(const AbParams& p, const CUtensorMap* tm, const uint8_t* src, int pitch, int col0, int64_t tile) {
    if (p.nbox == 1) tma_store_2d(tm, src, col0, (int)(tile * 128));
    else
        for (int gi = 0; gi < 2; ++gi) {
            int c1, c2, c3;
            group_coords(p.g, tile * 2 + gi, c1, c2, c3);
            tma_store_4d(tm, src + gi * 64 * pitch, col0, c1, c2, c3);
        }
}
__device__ __forceinline__ void store_tile(const AbParams& p, const CUtensorMap* tm, const uint8_t* src, int col0, int64_t tile) {
    store_tile_rows(p, tm, src, 128, col0, tile);
}

// =========================================================================================================
// forward.  Software pipeline of the 512 compute threads over items:  softmax(i) -> Q,K,V conversion of item i+1 (double-buffered
// tiles) -> epilogue(i), so the S / O contractions run under the conversion / epilogue instead of being waited for.
// smem: h [2] | weight ring | Q,K,V [2][3][16 KB] | P~ [16 KB] | row statistics | barriers.  The O tile is staged in the item's own
// (dead) Q tile and leaves by TMA store (warp 18).
//
// OUT = true additionally folds the attention block's tail in (reference Attention.to_out, src/vit_spatial_spectral.py:62-65,77, the
// residual add of Transformer.forward :102 and the PreNorm LayerNorm of the FeedForward branch :25-29):
//     xmid = x + dropout(concat_h(O_h) Wo^T + b_out),   h2 = LayerNorm(xmid),   stats2 = (mean, rstd)
// The staged bf16 O tile of every item is ALSO the K-major A operand of  Y (+)= O_h Wo[:, 64h .. 64h+63]^T  (M = 128, N = D), which
// accumulates over the heads of the tile in TMEM (two accumulators of D columns, alternating with the tile); the head's Wo column
// block travels through the weight ring as a fourth slot per item.  The tile epilogue (bias, dropout, residual, two-pass LayerNorm:
// the arithmetic of gemm_tn_kernel<6> / mlp_block_fwd_kernel) belongs to FOUR EXTRA WARPS (thread = tile row = TMEM lane of
// the M = 128 accumulator; warps 19-22, warp % 4 = lane quarter), not to the 512 softmax threads: folded into their loop (three variants were measured) it cost every item
// ~0.7 k clks of worse code on top of its own work.  It runs in NCH + 1 phases under the first items of the NEXT tile and moves its
// data through the K (| V) tile of the hosting item, which is dead from the item's S (O) contraction to the conversion two items
// later -- so no shared memory is added and nothing is accessed row-per-thread in global memory (one LSU wavefront per row):
//   phase c < NCH (item c):  S done (the MMA issuer commits t_kfree) -> TMA load of the [128 rows][32] fp32 chunk c of the residual x
//                            into the K tile -> xmid = x + drop(Y + b) formed IN PLACE -> warp 18 TMA-stores the chunk; the values
//                            are parked in the Y accumulator's own TMEM columns (tcgen05.st), the row sum rides in a register
//   phase NCH (item NCH):    O done (t_vfree) -> mean / rstd over the parked row (two-pass, thread-local) -> h2 into the K | V tiles
//                            as SWIZZLE_64B chunks -> TMA store
// The softmax threads only wait (kv_free) before the conversion that overwrites a buffer a phase used.  The kernel runs 23 warps
// (736 threads, 80 registers each: warp slots are granted four at a time, so 23 warps cost what 24 do); a setmaxnreg split (softmax
// warps 96, the others 48) was measured and gave the same time -- ptxas kept the softmax loop at ~84 registers either way.
// o [R, I] is still written (the Wo weight gradient of the backward reads it) but never re-read in the forward, and the separate
// out-projection GEMM launch (gemm_tn_kernel<6>) is gone.
// =========================================================================================================
constexpr int AB_THREADS_OUT = 736;            // 16 softmax warps + producer, MMA issuer, store + 4 tile-epilogue warps (19-22: warp % 4 = TMEM lane quarter)
// one thread: chunk c of the residual rows of `tile` -> dst ([128 rows][32 fp32] SWIZZLE_128B)
__device__ __forceinline__ void tail_load_x(const AbParams& p, AbBars* bars, const CUtensorMap* tma_x, uint8_t* dst, int c, int64_t tile) {
    const AttnGeom& g = p.g;
    const uint32_t rows = p.nbox == 1 ? 128u : (uint32_t)(g.G * g.N);
    mbar_arrive_expect_tx(&bars->xin_full, (uint32_t)p.nbox * rows * 128u);
    if (p.nbox == 1) tma_load_2d(dst, tma_x, &bars->xin_full, c * 32, (int)(tile * 128));
    else
        for (int gi = 0; gi < 2; ++gi) {
            int c1, c2, c3;
            group_coords(g, tile * 2 + gi, c1, c2, c3);
            tma_load_4d(dst + gi * AB_BLK, tma_x, &bars->xin_full, c * 32, c1, c2, c3);
        }
}
__device__ __forceinline__ void tail_prefetch_x(const AbParams& p, const CUtensorMap* tma_x, int c, int64_t tile) {
    if (tile >= p.n_tiles) return;
    if (p.nbox == 1) tma_prefetch_l2_2d(tma_x, c * 32, (int)(tile * 128));
    else
        for (int gi = 0; gi < 2; ++gi) {
            int c1, c2, c3;
            group_coords(p.g, tile * 2 + gi, c1, c2, c3);
            tma_prefetch_l2_4d(tma_x, c * 32, c1, c2, c3);
        }
}
// phases of the previous tile's epilogue that run at head h of the current tile: j in [lo, hi]; phase j runs at item min(j, H - 1)
__device__ __forceinline__ void tail_schedule(int h, int H, int nch, int& lo, int& hi) {
    lo = h; hi = (h == H - 1) ? nch : h;
    if (hi > nch) hi = nch;                                       // (lo > hi: nothing at this item)
}

template <int NCH, bool OUT>
__global__ void __launch_bounds__(OUT ? AB_THREADS_OUT : AB_THREADS, 1)
attn_block_fwd_kernel(const __grid_constant__ CUtensorMap tma_h, const __grid_constant__ CUtensorMap tma_w, const __grid_constant__ CUtensorMap tma_o,
                      const __grid_constant__ CUtensorMap tma_wo, const __grid_constant__ CUtensorMap tma_xmid, const __grid_constant__ CUtensorMap tma_h2,
                      const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ AbParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (smem_u32(smem) & 1023u) __trap();
    const AttnGeom& g = p.g;
    AbCommon s;
    s.h_s = smem; s.w_s = s.h_s + 2 * p.hbuf_bytes;
    uint8_t* qkv_s = s.w_s + (size_t)p.NR * p.slot_bytes;        // [2][Q | K | V]
    uint8_t* p_s = qkv_s + 6 * AB_T16;
    float* xch = reinterpret_cast<float*>(p_s + AB_T16);          // [max | sum][4 column quarters][128 lanes]
    s.bars = reinterpret_cast<AbBars*>(xch + 2 * 4 * 128);
    AbBars* bars = s.bars;
    const int warp = threadIdx.x >> 5;
    const int H = g.H;
    constexpr int D = NCH * 32;
    ab_setup(bars, warp, s.h_s, 2 * p.hbuf_bytes, nullptr, 0, &tma_h, &tma_w, &tma_o, OUT ? &tma_wo : nullptr, OUT ? &tma_xmid : nullptr, OUT ? 2u : 1u);
    s.tmem = bars->tmem_base; s.h_addr = smem_u32(s.h_s); s.w_addr = smem_u32(s.w_s);

    const int64_t my_tiles = blockIdx.x < p.n_tiles ? (p.n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int64_t n_items = my_tiles * H;
    if (warp < 16) {
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
        int64_t k = 0; int h = 0;
        bool hosted = false; uint32_t n_kvf0 = 0, n_kvf1 = 0;     // OUT: the previous item hosted epilogue phases / kv_free completions consumed
        if (n_items > 0) {                                        // pipeline prologue: Q, K, V of item 0
            mbar_wait(&bars->pro_full, 0);
            tc_fence_after();
            convert_qkv(s.tmem, qkv_s, lane_addr, L, cq);
            tc_fence_before();
            fence_proxy_async();
            warp_arrive(&bars->conv_done, lane);
        }
        for (int64_t it = 0; it < n_items; ++it) {
            const uint32_t ph = (uint32_t)it & 1;
            const int64_t tile = blockIdx.x + k * gridDim.x;
            if (tile != cur_tile) {
                cur_tile = tile;
                int64_t seq; int pos;
                const int64_t group = tile * 2 + blk;
                const bool ok = group < g.groups && slot_to(g, group, 0, m, seq, pos);
                grow = ok ? row_of(g, seq, pos) : -1;
                vm = ok ? vm_geom : 0u;
            }
            // ---- softmax of row `slot`, key columns 16*cq .. +15 of its block ----
            const uint64_t hidx = pair_base(g, tile * 2 + blk, h) + (uint64_t)(m * 32 + 8 * cq);
            const uint32_t hash_lo = (uint32_t)hidx + (uint32_t)(seed >> 32);
            const uint32_t hash_hi = ((uint32_t)(hidx >> 32) * 0x85EBCA77u) ^ (uint32_t)seed ^ (p.drop.site * 0xC2B2AE3Du);
            if (threadIdx.x == 0) AB_T(it, 0);
            mbar_wait(&bars->s_full, ph);
            if (threadIdx.x == 0) AB_T(it, 1);
            tc_fence_after();
            float sc[16];                                        // raw scores; the softmax scale is folded into the exponent (scale > 0: the max commutes)
            {
                uint32_t a[16];
                tmem_ld_32x16(s.tmem + lane_addr + COL_S + 16 * cq, a);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) sc[j] = __uint_as_float(a[j]);
            }
            float mx = -INFINITY;
            if (vm == 0xFFFFu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) mx = fmaxf(mx, sc[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) { sc[j] = (vm >> j) & 1u ? sc[j] : -INFINITY; mx = fmaxf(mx, sc[j]); }
            }
            xmax[cq * 128 + L] = mx;
            asm volatile("bar.sync 1, 512;" ::: "memory");
            mx = fmaxf(fmaxf(xmax[L], xmax[128 + L]), fmaxf(xmax[256 + L], xmax[384 + L]));
            const float sub = mx == -INFINITY ? 0.f : mx * sl2;
            float l = 0.f;
            uint32_t pk[8];
            // (the dropout test is hoisted by hand: in the larger OUT kernel the compiler no longer unswitches the unrolled loop on it)
            auto exp_row = [&](auto drop_tag) {
                constexpr bool DROP = decltype(drop_tag)::value;
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    float p0 = ex2_approx(fmaf(sc[2 * jj], sl2, -sub)), p1 = ex2_approx(fmaf(sc[2 * jj + 1], sl2, -sub));
                    l += p0 + p1;
                    if (DROP) {
                        uint32_t x = (hash_lo + (uint32_t)jj) * 0x9E3779B1u ^ hash_hi;
                        x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
                        p0 *= (x & 0xFFFFu) >= t16 ? p.drop.scale : 0.f;
                        p1 *= (x >> 16) >= t16 ? p.drop.scale : 0.f;
                    }
                    pk[jj] = pack_bf(p0, p1);
                }
            };
            if (p.drop.on()) exp_row(std::true_type{}); else exp_row(std::false_type{});
            xsum[cq * 128 + L] = l;
            store_row16_packed(p_s + blk * AB_BLK, m, cq, pk);      // its last readers (the O MMAs of item it - 1) completed before S(it) did
            fence_proxy_async();
            tc_fence_before();
            warp_arrive(&bars->p_full, lane);
            if (threadIdx.x == 0) AB_T(it, 2);
            asm volatile("bar.sync 1, 512;" ::: "memory");
            l = (xsum[L] + xsum[128 + L]) + (xsum[256 + L] + xsum[384 + L]);
            const float inv = l > 0.f ? 1.f / l : 0.f;
            if (cq == 0 && grow >= 0) p.lse_out[grow * H + h] = (sub + log2f(l)) * 0.6931471805599453f;
            // ---- Q, K, V of item it + 1 -> the other tile buffer (runs under the O MMAs of item it) ----
            // the other buffer's Q tile staged O(it - 1): its store (and, OUT, the out-projection MMA) has read it.  Waited for at EVERY item, the
            // last one included: it keeps stg_full in lockstep with the store warp -- a waiter two phases behind deadlocks on the parity test
            if (it >= 1) mbar_wait(&bars->stg_free, (uint32_t)(it - 1) & 1);
            if (it + 1 < n_items) {
                if (OUT && hosted) {                              // ... and the epilogue phases its K | V tiles staged at item it - 1 have left
                    if (it & 1) { mbar_wait(&bars->kv_free[0], n_kvf0 & 1u); ++n_kvf0; }   // (one barrier per item parity: a completion can never run two ahead of this wait)
                    else { mbar_wait(&bars->kv_free[1], n_kvf1 & 1u); ++n_kvf1; }
                }
                mbar_wait(&bars->pro_full, ph ^ 1u);
                if (threadIdx.x == 0) AB_T(it, 3);
                tc_fence_after();
                convert_qkv(s.tmem, qkv_s + (size_t)((it + 1) & 1) * 3 * AB_T16, lane_addr, L, cq);
                tc_fence_before();
                fence_proxy_async();
                warp_arrive(&bars->conv_done, lane);
                if (threadIdx.x == 0) AB_T(it, 4);
            }
            // ---- epilogue: O row / l -> bf16 -> staging (the item's own Q tile, dead since S) -> TMA store ----
            mbar_wait(&bars->o_full, ph);
            if (threadIdx.x == 0) AB_T(it, 5);
            tc_fence_after();
            {
                uint32_t v[16], o8[8];
                tmem_ld_32x16(s.tmem + lane_addr + COL_O + 16 * cq, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) o8[j] = pack_bf(__uint_as_float(v[2 * j]) * inv, __uint_as_float(v[2 * j + 1]) * inv);
                store_row16_packed(qkv_s + (size_t)(it & 1) * 3 * AB_T16, slot, cq, o8);
            }
            tc_fence_before();
            fence_proxy_async();
            warp_arrive(&bars->stg_full, lane);
            if (threadIdx.x == 0) AB_T(it, 6);
            if (OUT) { int lo, hi; tail_schedule(h, H, NCH, lo, hi); hosted = k > 0 && lo <= hi; }   // did THIS item host phases? (tested at the next one)
            if (++h == H) { h = 0; ++k; }
        }
    } else
    if (warp >= 16 && warp < 19) {
    if (warp == 16) {
        // ===== TMA producer.  Ring order = the MMA issuer's consumption order: P(0), P(1), then per item: [Wo(it - 1)], P(it + 2); [Wo(last)] =====
        if (elect_one()) {
            Ring r{0, 0u, p.NR};
            int64_t pk = 0; int phd = 0;                          // (tile counter, head) of the next prologue
            auto load_pro = [&]() {
                if (phd == 0) {
                    const int buf = (int)(pk & 1);
                    mbar_wait(&bars->h_empty[buf], ((uint32_t)(pk >> 1) & 1) ^ 1);
                    load_h_tile<NCH>(p, &tma_h, s.h_s + (size_t)buf * p.hbuf_bytes, &bars->h_full[buf], blockIdx.x + pk * gridDim.x);
                }
                for (int t = 0; t < 3; ++t) load_w_slice<NCH>(p, &tma_w, bars, s.w_s, r, t, phd);
                if (++phd == H) { phd = 0; ++pk; }
            };
            int yh = 0;
            auto load_wo = [&]() {                                // Wo[:, 64 yh .. +63]: [D rows][64 cols] SWIZZLE_128B, K-major B of the out-projection
                mbar_wait(&bars->w_empty[r.slot], r.ph ^ 1u);
                mbar_arrive_expect_tx(&bars->w_full[r.slot], (uint32_t)D * 128u);
                tma_load_2d(s.w_s + (size_t)r.slot * p.slot_bytes, &tma_wo, &bars->w_full[r.slot], yh * 64, 0);
                r.next();
                if (++yh == H) yh = 0;
            };
            if (n_items > 0) load_pro();
            if (n_items > 1) load_pro();
            for (int64_t it = 0; it < n_items; ++it) {
                if (OUT && it >= 1) load_wo();
                if (it + 2 < n_items) load_pro();
            }
            if (OUT && n_items > 0) load_wo();
        }
    } else if (warp == 17) {
        // ===== MMA issuer: prologue(0), S(0), prologue(1); then per item: [Y(i-1)], O(i), S(i+1), prologue(i+2); [Y(last)] =====
        // (the whole warp runs the control flow, one elected lane issues: see the backward kernel)
        if (n_items > 0) {
            const bool lead = elect_one();
            Ring r{0, 0u, p.NR};
            const uint32_t idesc_s = make_idesc_bf16(64, 64, 0, 0), idesc_o = make_idesc_bf16(64, 64, 0, 1), idesc_y = make_idesc_bf16(128, D, 0, 0);
            const uint32_t qkv_addr = smem_u32(qkv_s);
            const uint64_t pd0 = kdesc(smem_u32(p_s)), pd1 = kdesc(smem_u32(p_s) + AB_BLK);
            int64_t pk = 0; int phd = 0;                          // (tile counter, head) of the next prologue
            auto next_pro = [&]() { issue_prologue<NCH>(p, s, pk, phd, r, lead); if (++phd == H) { phd = 0; ++pk; } };
            int64_t sk = 0; int sh = 0;                           // (tile counter, head) of the next S item
            auto issue_s = [&](int64_t it) {
                const uint32_t qa = qkv_addr + (uint32_t)(it & 1) * 3 * AB_T16;
                mbar_wait(&bars->conv_done, (uint32_t)it & 1);    // Q, K, V tiles written (generic proxy + the writers' fence)
                if (lead) AB_T(it, 8);
                tc_fence_after();
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const uint64_t dq = kdesc(qa + b * AB_BLK), dk = kdesc(qa + AB_T16 + b * AB_BLK);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)                // S_b = Q_b K_b^T
                        if (lead) umma_bf16(s.tmem + COL_S + ((uint32_t)(16 * b) << 16), dq + (uint64_t)(ks * 2), dk + (uint64_t)(ks * 2), idesc_s, ks != 0);
                }
                if (lead) umma_commit(&bars->s_full);
                if (OUT) {                                        // the item hosts phases of the previous tile's epilogue: its K tile is theirs once S is done
                    int lo, hi;
                    tail_schedule(sh, H, NCH, lo, hi);
                    if (lead && sk > 0 && lo <= hi) umma_commit(&bars->t_kfree[it & 1]);
                    if (++sh == H) { sh = 0; ++sk; }
                }
            };
            int yh = 0; uint32_t ycol = COL_Y;                    // head / accumulator (tile parity) of the next out-projection item
            auto issue_y = [&](int64_t it) {                      // Y (+)= O(it) Wo_h^T: the staged O tile [128 rows][64] is the K-major A operand
                mbar_wait(&bars->stg_full, (uint32_t)it & 1);
                mbar_wait(&bars->w_full[r.slot], r.ph);
                if (lead) AB_T(it, 11);
                tc_fence_after();
                const uint64_t ad = kdesc(qkv_addr + (uint32_t)(it & 1) * 3 * AB_T16), bd = kdesc(s.w_addr + (uint32_t)r.slot * p.slot_bytes);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    if (lead && !(p.tail_dbg & 1)) umma_bf16(s.tmem + ycol, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), idesc_y, (yh | ks) != 0);
                if (lead) {
                    umma_commit(&bars->w_empty[r.slot]);
                    umma_commit(&bars->stg_free);                 // second arrival (the first is the O store's): the Q tile may be overwritten
                    if (yh == H - 1) umma_commit(&bars->y_full[ycol != COL_Y]);   // (one barrier per tile parity, like the accumulator)
                }
                r.next();
                if (++yh == H) { yh = 0; ycol ^= (COL_Y ^ (COL_Y + Y_COLS)); }
            };
            next_pro();
            issue_s(0);
            if (n_items > 1) next_pro();
            int64_t k = 0; int h = 0;
            for (int64_t it = 0; it < n_items; ++it) {
                if (OUT && it >= 1) issue_y(it - 1);
                const uint32_t va = qkv_addr + (uint32_t)(it & 1) * 3 * AB_T16 + 2 * AB_T16;
                mbar_wait(&bars->p_full, (uint32_t)it & 1);
                if (lead) AB_T(it, 10);
                tc_fence_after();
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const uint64_t dv = mndesc(va + b * AB_BLK), dp = b ? pd1 : pd0;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)                // O_b = P~_b V_b (V as MN-major B)
                        if (lead) umma_bf16(s.tmem + COL_O + ((uint32_t)(16 * b) << 16), dp + (uint64_t)(ks * 2), dv + (uint64_t)(ks * 128), idesc_o, ks != 0);
                }
                if (lead) umma_commit(&bars->o_full);
                if (OUT) {                                        // the item hosts the LayerNorm phase: its V tile is free as well once O is done
                    int lo, hi;
                    tail_schedule(h, H, NCH, lo, hi);
                    if (lead && k > 0 && lo <= hi && hi == NCH) umma_commit(&bars->t_vfree[it & 1]);
                    if (++h == H) { h = 0; ++k; }
                }
                if (it + 1 < n_items) {
                    issue_s(it + 1);
                    if (it + 2 < n_items) next_pro();             // the accumulators of item it + 1 were read before its conv_done
                }
                if (lead) AB_T(it, 9);
            }
            if (OUT) {
                issue_y(n_items - 1);
                if (lead) { umma_commit(&bars->t_kfree[n_items & 1]); umma_commit(&bars->t_vfree[n_items & 1]); }   // the CTA's last tile: its phases run after the last item
            }
        }
    } else if (warp == 18) {
        // ===== TMA store of the staged O tiles (and, OUT, of the staged xmid / h2 chunks of the previous tile's epilogue phases) =====
        if (elect_one()) {
            uint32_t n_out = 0;                                   // staging rounds stored so far
            auto out_round = [&](int j, int64_t tile, const uint8_t* kv) {   // phase j of the epilogue of `tile`, staged at kv
                mbar_wait(&bars->out_full, n_out & 1u);
                if (j < NCH) store_tile(p, &tma_xmid, kv, 32 * j, tile);
                else {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) store_tile_rows(p, &tma_h2, kv + c * 8192, 64, 32 * c, tile);
                }
                tma_store_commit();
                tma_store_wait_read();
                mbar_arrive(&bars->out_free);
                ++n_out;
            };
            int64_t k = 0; int h = 0;
            for (int64_t it = 0; it < n_items; ++it) {
                const int64_t tile = blockIdx.x + k * gridDim.x;
                const uint8_t* buf = qkv_s + (size_t)(it & 1) * 3 * AB_T16;
                if (OUT && k > 0) {
                    int lo, hi;
                    tail_schedule(h, H, NCH, lo, hi);
                    for (int j = lo; j <= hi; ++j) out_round(j, tile - gridDim.x, buf + AB_T16);
                    if (lo <= hi) mbar_arrive(&bars->kv_free[it & 1]);   // the softmax threads may convert into this buffer again
                }
                mbar_wait(&bars->stg_full, (uint32_t)it & 1);
                if (p.o) {
                    store_tile(p, &tma_o, buf, h * 64, tile);
                    tma_store_commit();
                    tma_store_wait_read();
                }
                mbar_arrive(&bars->stg_free);
                if (++h == H) { h = 0; ++k; }
            }
            if (OUT && n_items > 0)
                for (int j = 0; j <= NCH; ++j) out_round(j, blockIdx.x + (my_tiles - 1) * gridDim.x, qkv_s + (size_t)(n_items & 1) * 3 * AB_T16 + AB_T16);
        }
    }
    } else if (OUT && warp >= 19) {
        // ===== tile-epilogue warps: thread = tile row T = TMEM lane T of the out-projection accumulator =====
        const int lane = threadIdx.x & 31, tq = warp & 3;
        const int T = tq * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(tq * 32) << 16;
        const uint32_t sw128 = (uint32_t)(T & 7), sw64 = (uint32_t)((T >> 1) & 3);
        uint32_t n_in = 0, n_out = 0, n_kf0 = 0, n_kf1 = 0, n_vf0 = 0, n_vf1 = 0;   // completions consumed so far per barrier
        // b_out | ln_w | ln_b in shared memory: every thread needs all D of each, and a global load per use misses the (tiny) L1 -- ~600 clks apiece,
        // which made a chunk phase take 6 k clks
        float* cst = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + ((sizeof(AbBars) + 15) & ~size_t(15)));
        if (T < D) { cst[T] = p.b_out[T]; cst[D + T] = p.ln_w[T]; cst[2 * D + T] = p.ln_b[T]; }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (T == 0 && my_tiles > 0)                               // the first tile's residual rows -> L2 while its items run
            for (int c = 0; c < NCH; ++c) tail_prefetch_x(p, &tma_x, c, blockIdx.x);
        Drop dro = p.drop_out;                                    // (device-resident seed offset folded in once)
        dro.seed += dro.seed_dev ? __ldg(dro.seed_dev) : 0ull;
        dro.seed_dev = nullptr;
        for (int64_t kk = 0; kk < my_tiles; ++kk) {
            const bool last = kk + 1 == my_tiles;
            const int64_t ptile = blockIdx.x + kk * gridDim.x;
            int64_t row;                                          // global row of tile row T, or -1
            {
                int64_t seq; int pos;
                const int64_t gl = ptile * 2 + (T >> 6);
                row = (gl < g.groups && slot_to(g, gl, 0, T & 63, seq, pos)) ? row_of(g, seq, pos) : -1;
            }
            const uint32_t ty = s.tmem + lane_addr + COL_Y + (uint32_t)(kk & 1) * Y_COLS;
            float sum = 0.f;
            int64_t host_prev = -1;
            for (int j = 0; j <= NCH; ++j) {
                const int64_t host = last ? n_items : (kk + 1) * H + (j < H - 1 ? j : H - 1);   // the item whose K | V tiles stage this phase
                uint8_t* kv = qkv_s + (size_t)(host & 1) * 3 * AB_T16 + AB_T16;
                const bool same_host = host == host_prev;
                if (same_host && n_out > 0) mbar_wait(&bars->out_free, (n_out - 1u) & 1u);   // the K tile still stages the previous round
                if (!same_host) {                                 // S of the hosting item is done: its K tile is dead
                    if (host & 1) { mbar_wait(&bars->t_kfree[1], n_kf1 & 1u); ++n_kf1; }
                    else { mbar_wait(&bars->t_kfree[0], n_kf0 & 1u); ++n_kf0; }
                    host_prev = host;
                }
                if (T == 0 && !last) AB_T(host, 13);
                // (a different hosting item = the other buffer: the chunk load below overlaps the previous round's TMA store)
                if (j < NCH && warp == 19 && lane == 0) {
                    tail_load_x(p, bars, &tma_x, kv, j, ptile);
                    // pull what the next phase / the next tile's first phase will load into L2 meanwhile
                    if (j + 1 < NCH) tail_prefetch_x(p, &tma_x, j + 1, ptile);
                    else if (!last) tail_prefetch_x(p, &tma_x, 0, ptile + gridDim.x);
                }
                if (!same_host && n_out > 0) mbar_wait(&bars->out_free, (n_out - 1u) & 1u);   // (completions are observed in order)
                if (j < NCH) {
                    const int c = j;
                    if (j == 0) mbar_wait(&bars->y_full[kk & 1], (uint32_t)(kk >> 1) & 1u);
                    mbar_wait(&bars->xin_full, n_in & 1u);
                    ++n_in;
                    if (T == 0 && !last) AB_T(host, 14);
                    tc_fence_after();
                    uint8_t* xr = kv + T * 128;
#pragma unroll 1
                    for (int i = 0; i < 2; ++i) {                 // 16 columns at a time
                        uint32_t a[16];
                        tmem_ld_32x16(ty + 32 * c + 16 * i, a);
                        tmem_ld_wait();
                        const int col = 32 * c + 16 * i;
#pragma unroll
                        for (int hf = 0; hf < 4; ++hf) {
                            float4* xp = reinterpret_cast<float4*>(xr + ((((uint32_t)(4 * i + hf)) ^ sw128) << 4));
                            const float4 r = *xp;
                            const float4 bb = *reinterpret_cast<const float4*>(cst + col + 4 * hf);
                            float f0 = __uint_as_float(a[4 * hf]) + bb.x, f1 = __uint_as_float(a[4 * hf + 1]) + bb.y, f2 = __uint_as_float(a[4 * hf + 2]) + bb.z,
                                  f3 = __uint_as_float(a[4 * hf + 3]) + bb.w;
                            if (dro.on()) {
                                float d4[4];
                                drop_factor4(dro, (uint64_t)((row >= 0 ? row : 0) * D + col + 4 * hf) >> 2, d4);
                                f0 *= d4[0]; f1 *= d4[1]; f2 *= d4[2]; f3 *= d4[3];
                            }
                            f0 += r.x; f1 += r.y; f2 += r.z; f3 += r.w;
                            *xp = make_float4(f0, f1, f2, f3);    // xmid, staged in place for the TMA store
                            sum += (f0 + f1) + (f2 + f3);
                            a[4 * hf] = __float_as_uint(f0); a[4 * hf + 1] = __float_as_uint(f1); a[4 * hf + 2] = __float_as_uint(f2); a[4 * hf + 3] = __float_as_uint(f3);
                        }
                        tmem_st_32x16(ty + 32 * c + 16 * i, a);   // parked for the LayerNorm phase
                    }
                    tmem_st_wait();
                } else {
                    // O of the hosting item is done: the V tile is dead as well.  (Barrier pairs by tile / item parity throughout: with few heads
                    //  the next completion of a single barrier can overtake this wait by two phases, and the parity test then never returns)
                    if (host & 1) { mbar_wait(&bars->t_vfree[1], n_vf1 & 1u); ++n_vf1; }
                    else { mbar_wait(&bars->t_vfree[0], n_vf0 & 1u); ++n_vf0; }
                    tc_fence_after();
                    const float mean = sum / D;
                    float sq = 0.f;
#pragma unroll 1
                    for (int i = 0; i < D / 16; ++i) {
                        uint32_t a[16];
                        tmem_ld_32x16(ty + 16 * i, a);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) { const float d = __uint_as_float(a[e]) - mean; sq = fmaf(d, d, sq); }
                    }
                    const float rstd = rsqrtf(sq / D + 1e-5f);
                    if (row >= 0) *reinterpret_cast<float2*>(p.ln_stats + 2 * row) = make_float2(mean, rstd);
#pragma unroll 1
                    for (int i = 0; i < D / 16; ++i) {            // h2 chunk i / 2 ([128 rows][64 B] SWIZZLE_64B at kv + 8 KB * chunk), 16-byte units 2 (i % 2), + 1
                        uint32_t a[16];
                        tmem_ld_32x16(ty + 16 * i, a);
                        tmem_ld_wait();
                        const int col = 16 * i;
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            uint32_t o[4];
#pragma unroll
                            for (int hf = 0; hf < 2; ++hf) {
                                const int e0 = 8 * u + 4 * hf;
                                const float4 w = *reinterpret_cast<const float4*>(cst + D + col + e0), lb = *reinterpret_cast<const float4*>(cst + 2 * D + col + e0);
                                o[2 * hf] = pack_bf((__uint_as_float(a[e0]) - mean) * rstd * w.x + lb.x, (__uint_as_float(a[e0 + 1]) - mean) * rstd * w.y + lb.y);
                                o[2 * hf + 1] = pack_bf((__uint_as_float(a[e0 + 2]) - mean) * rstd * w.z + lb.z, (__uint_as_float(a[e0 + 3]) - mean) * rstd * w.w + lb.w);
                            }
                            *reinterpret_cast<uint4*>(kv + (i >> 1) * 8192 + T * 64 + ((((uint32_t)(2 * (i & 1) + u)) ^ sw64) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                        }
                    }
                }
                tc_fence_before();
                fence_proxy_async();
                warp_arrive(&bars->out_full, lane);
                ++n_out;
                if (T == 0 && !last) AB_T(host, 15);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) { tc_fence_after(); tmem_dealloc(s.tmem, 512); }
}

// =========================================================================================================
// backward (one item at a time; the prologue of item i + 1 runs under the softmax of item i, the data-gradient MMAs of item i under
// the conversion of item i + 1).  smem: h [2] | weight ring | Q,K,V [48 KB] | dO [2][16 KB] | P~ | dS | partial row sums | barriers.
// dQ / dK / dV leave through the dead P~ / dS / dO tiles by TMA store (warp 18).
// =========================================================================================================
template <int NCH>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_block_bwd_kernel(const __grid_constant__ CUtensorMap tma_h, const __grid_constant__ CUtensorMap tma_wt, const __grid_constant__ CUtensorMap tma_do,
                      const __grid_constant__ CUtensorMap tma_dqkv, const __grid_constant__ CUtensorMap tma_dh, const AbParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (smem_u32(smem) & 1023u) __trap();
    const AttnGeom& g = p.g;
    AbCommon s;
    s.h_s = smem; s.w_s = s.h_s + p.hbuf_bytes;
    uint8_t* qkv_s = s.w_s + (size_t)p.NR * p.slot_bytes;        // Q | K | V
    uint8_t* do_s = qkv_s + 3 * AB_T16;                           // [2][16 KB]
    uint8_t* p_s = do_s + 2 * AB_T16;                             // P~ : two [64 q][64 k] blocks
    uint8_t* ds_s = p_s + AB_T16;                                 // dS
    uint8_t* dhs_s = ds_s + AB_T16;                               // dh staging: [128 rows][32 fp32], SWIZZLE_128B (one 32-column round at a time)
    float* xch = reinterpret_cast<float*>(dhs_s + AB_T16);        // [4 column quarters][128 lanes] partial row sums
    s.bars = reinterpret_cast<AbBars*>(xch + 4 * 128);
    AbBars* bars = s.bars;
    const int warp = threadIdx.x >> 5;
    const int H = g.H, I = H * 64;
    constexpr int D = NCH * 32;
    ab_setup(bars, warp, s.h_s, p.hbuf_bytes, do_s, 2 * AB_T16, &tma_h, &tma_wt, &tma_do, &tma_dqkv, &tma_dh);
    s.tmem = bars->tmem_base; s.h_addr = smem_u32(s.h_s); s.w_addr = smem_u32(s.w_s);

    const int64_t my_tiles = blockIdx.x < p.n_tiles ? (p.n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int64_t n_items = my_tiles * H;
    const uint32_t do_rows = p.nbox == 1 ? 128u : (uint32_t)(g.G * g.N);

    if (warp == 16) {
        // ===== TMA producer: per item the three W^T column blocks of its head (ring), h once per tile, dO per item =====
        if (elect_one()) {
            Ring r{0, 0u, p.NR};
            auto load_do = [&](int64_t it, int64_t tile, int h) {
                const int st = (int)(it & 1);
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
            auto load_blocks = [&](int h) {                       // the three W^T column blocks [D rows][64] of head h
                for (int t = 0; t < 3; ++t) {
                    mbar_wait(&bars->w_empty[r.slot], r.ph ^ 1u);
                    mbar_arrive_expect_tx(&bars->w_full[r.slot], (uint32_t)D * 128u);
                    tma_load_2d(s.w_s + (size_t)r.slot * p.slot_bytes, &tma_wt, &bars->w_full[r.slot], t * I + h * 64, 0);
                    r.next();
                }
            };
            int64_t pk = 0; int phd = 0;                          // item of the next recomputation ("P") group
            auto load_pro = [&]() {
                if (phd == 0) {
                    mbar_wait(&bars->h_empty[0], ((uint32_t)pk & 1) ^ 1);
                    load_h_tile<NCH>(p, &tma_h, s.h_s, &bars->h_full[0], blockIdx.x + pk * gridDim.x);
                    if (pk + 1 < my_tiles) prefetch_h_tile<NCH>(p, &tma_h, blockIdx.x + (pk + 1) * gridDim.x);   // single smem buffer: cut the next load to an L2 hit
                }
                load_blocks(phd);
                if (++phd == H) { phd = 0; ++pk; }
            };
            // ring order = the MMA issuer's consumption order: P(0), P(1), then per item: G(it), P(it + 2)
            if (n_items > 0) { load_pro(); load_do(0, blockIdx.x, 0); }
            if (n_items > 1) load_pro();
            int64_t k = 0; int h = 0;
            for (int64_t it = 0; it < n_items; ++it) {
                int64_t k1 = k; int h1 = h + 1;
                if (h1 == H) { h1 = 0; ++k1; }
                if (it + 1 < n_items) load_do(it + 1, blockIdx.x + k1 * gridDim.x, h1);
                load_blocks(h);                                   // G(it)
                if (it + 2 < n_items) load_pro();                 // P(it + 2)
                k = k1; h = h1;
            }
        }
    } else if (warp == 17) {
        // ===== MMA issuer.  Order: prologue(0), S|dP(0), phase2(0), prologue(1); then per item i: S|dP(i+1), dgrad(i), phase2(i+1), prologue(i+2) =====
        // The whole warp runs the control flow (waits, ring / descriptor arithmetic: warp-uniform, so it lives in uniform registers) and one
        // elected lane issues the tcgen05 instructions: with everything under elect_one() every descriptor crossed from vector to uniform
        // registers before each UMMA group (~13 instructions per UMMA on a single dependent-issue warp that sits on the item's critical path)
        if (n_items > 0) {
            const bool lead = elect_one();
            Ring r{0, 0u, p.NR};
            const uint32_t idesc_kk = make_idesc_bf16(64, 64, 0, 0), idesc_mn = make_idesc_bf16(64, 64, 1, 1), idesc_kmn = make_idesc_bf16(64, 64, 0, 1);
            const uint32_t idesc_dh = make_idesc_bf16(128, D, 0, 0);
            const uint32_t qa = smem_u32(qkv_s), ka = qa + AB_T16, va = ka + AB_T16, pa = smem_u32(p_s), dsa = smem_u32(ds_s);
            int64_t pk = 0; int phd = 0;                          // (tile counter, head) of the next prologue
            auto next_pro = [&]() { issue_prologue_bwd<NCH>(p, s, pk, phd, r, lead); if (++phd == H) { phd = 0; ++pk; } };
            auto issue_sdp = [&](int64_t it) {                    // S_b = Q_b K_b^T | dP_b = dO_b V_b^T  -> the (dead) prologue columns [0, 128)
                const int st = (int)(it & 1);
                const uint32_t doa = smem_u32(do_s) + (uint32_t)st * AB_T16;
                mbar_wait(&bars->conv_done, (uint32_t)it & 1);
                mbar_wait(&bars->do_full[st], (uint32_t)(it >> 1) & 1);
                if (lead) AB_T(it, 8);
                tc_fence_after();
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const uint32_t lane_off = (uint32_t)(16 * b) << 16;
                    const uint64_t dq = kdesc(qa + b * AB_BLK), dk = kdesc(ka + b * AB_BLK), dd = kdesc(doa + b * AB_BLK), dv = kdesc(va + b * AB_BLK);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        if (lead) umma_bf16(s.tmem + COL_P + lane_off, dq + (uint64_t)(ks * 2), dk + (uint64_t)(ks * 2), idesc_kk, ks != 0);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        if (lead) umma_bf16(s.tmem + COL_P + 64 + lane_off, dd + (uint64_t)(ks * 2), dv + (uint64_t)(ks * 2), idesc_kk, ks != 0);
                }
                if (lead) umma_commit(&bars->s_full);
            };
            auto issue_phase2 = [&](int64_t it) {                 // dV | dK | dQ of item it -> X
                const int st = (int)(it & 1);
                const uint32_t doa = smem_u32(do_s) + (uint32_t)st * AB_T16;
                mbar_wait(&bars->p_full, (uint32_t)it & 1);       // P~ and dS written
                if (lead) AB_T(it, 10);
                tc_fence_after();
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const uint32_t lane_off = (uint32_t)(16 * b) << 16;
                    const uint64_t mp = mndesc(pa + b * AB_BLK), mdo = mndesc(doa + b * AB_BLK), mds = mndesc(dsa + b * AB_BLK), mq = mndesc(qa + b * AB_BLK);
                    const uint64_t kds = kdesc(dsa + b * AB_BLK), mk = mndesc(ka + b * AB_BLK);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)                // dV_b = P~_b^T dO_b (reduction over the block's 64 queries, 16 rows = 2 KB per step)
                        if (lead) umma_bf16(s.tmem + COL_X + lane_off, mp + (uint64_t)(ks * 128), mdo + (uint64_t)(ks * 128), idesc_mn, ks != 0);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)                // dK_b = dS_b^T Q_b
                        if (lead) umma_bf16(s.tmem + COL_X + 64 + lane_off, mds + (uint64_t)(ks * 128), mq + (uint64_t)(ks * 128), idesc_mn, ks != 0);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)                // dQ_b = dS_b K_b
                        if (lead) umma_bf16(s.tmem + COL_X + 128 + lane_off, kds + (uint64_t)(ks * 2), mk + (uint64_t)(ks * 128), idesc_kmn, ks != 0);
                }
                if (lead) umma_commit(&bars->o_full);
            };
            // the recomputation of item i + 2 is issued as soon as the softmax threads hold S | dP of item i + 1 in registers (sdp_read), i.e.
            // BEFORE the dV | dK | dQ contractions of item i + 1: behind them in the in-order tensor pipe it finished ~1 k clks after o_full
            auto pro_after_read = [&](int64_t it) {
                mbar_wait(&bars->sdp_read, (uint32_t)it & 1);
                tc_fence_after();
                next_pro();
            };
            next_pro();
            issue_sdp(0);
            if (n_items > 1) pro_after_read(0);
            issue_phase2(0);
            int64_t k = 0; int h = 0;
            for (int64_t it = 0; it < n_items; ++it) {
                if (it + 1 < n_items) issue_sdp(it + 1);
                // ---- data gradient of the projection: dh (+)= [dQ | dK | dV] (bf16 rows in TMEM) x W^T blocks ----
                mbar_wait(&bars->a_ready, (uint32_t)it & 1);
                if (lead) AB_T(it, 11);
                if (h == 0 && k > 0) mbar_wait(&bars->dh_free, (uint32_t)(k - 1) & 1);   // the previous tile's dh rows have been read
                tc_fence_after();
#pragma unroll
                for (int t = 0; t < 3; ++t) {                    // the head's weight blocks again, now as K-major B
                    mbar_wait(&bars->w_full[r.slot], r.ph);
                    tc_fence_after();
                    const uint64_t bw = kdesc(s.w_addr + (uint32_t)r.slot * p.slot_bytes);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        if (lead) umma_bf16_ts(s.tmem + COL_DH, s.tmem + COL_X + 32 * t + 8 * ks, bw + (uint64_t)(ks * 2), idesc_dh, (h | t | ks) != 0);
                    if (lead) umma_commit(&bars->w_empty[r.slot]);
                    r.next();
                }
                if (h == H - 1) if (lead) umma_commit(&bars->dh_full);
                if (lead) AB_T(it, 12);
                if (it + 1 < n_items) {
                    if (it + 2 < n_items) pro_after_read(it + 1);
                    issue_phase2(it + 1);
                    if (lead) AB_T(it, 9);
                }
                if (++h == H) { h = 0; ++k; }
            }
        }
    } else if (warp == 18) {
        // ===== TMA store of the staged dQ (P~ tile) / dK (dS tile) / dV (dO stage) =====
        if (elect_one()) {
            for (int64_t it = 0; it < n_items; ++it) {
                const int64_t tile = blockIdx.x + (it / H) * gridDim.x; const int h = (int)(it % H);
                const int st = (int)(it & 1);
                mbar_wait(&bars->stg_full, (uint32_t)it & 1);
                store_tile(p, &tma_dqkv, p_s, h * 64, tile);
                store_tile(p, &tma_dqkv, ds_s, I + h * 64, tile);
                store_tile(p, &tma_dqkv, do_s + (size_t)st * AB_T16, 2 * I + h * 64, tile);
                tma_store_commit();
                tma_store_wait_read();
                mbar_arrive(&bars->stg_free);
                mbar_arrive(&bars->do_empty[st]);                 // the dO stage is free for the load of item it + 2
                if (h == H - 1) {                                 // the tile's dh rows, NCH rounds of 32 columns (one [128 rows][128 B] fp32 tile each)
                    const int64_t rd0 = (it / H) * NCH;
                    for (int c = 0; c < NCH; ++c) {
                        mbar_wait(&bars->dhs_full, (uint32_t)(rd0 + c) & 1);
                        store_tile(p, &tma_dh, dhs_s, 32 * c, tile);
                        tma_store_commit();
                        tma_store_wait_read();
                        mbar_arrive(&bars->dhs_free);
                    }
                }
            }
        }
    } else {
        // ===== 512 compute threads.  Software pipeline: after the five contractions of item i complete, first convert Q, K, V of item
        // i + 1 (its S | dP contractions then run under the epilogue of item i), then the epilogue of item i, then the softmax of i + 1 =====
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
        struct Item { int64_t k, tile, grow; int h; uint32_t vm; };
        auto locate = [&](Item& x) {                              // row / validity of this thread's slot in tile x.tile
            int64_t seq; int pos;
            const int64_t group = x.tile * 2 + blk;
            const bool ok = group < g.groups && slot_to(g, group, 0, m, seq, pos);
            x.grow = ok ? row_of(g, seq, pos) : -1;
            x.vm = ok ? vm_geom : 0u;
        };
        auto advance = [&](Item& x) {
            if (++x.h == H) { x.h = 0; ++x.k; x.tile = blockIdx.x + x.k * gridDim.x; locate(x); }
        };
        auto convert = [&](int64_t it) {
            mbar_wait(&bars->pro_full, (uint32_t)it & 1);
            if (threadIdx.x == 0) AB_T(it, 1);
            tc_fence_after();
            convert_qkv(s.tmem, qkv_s, lane_addr, L, cq);
            tc_fence_before();
            fence_proxy_async();
            warp_arrive(&bars->conv_done, lane);
            if (threadIdx.x == 0) AB_T(it, 2);
        };
        auto softmax = [&](int64_t it, const Item& x) {           // softmax backward for row `slot`, key columns 16*cq .. +15
            const float Lse = x.grow >= 0 ? __ldg(p.lse_in + x.grow * H + x.h) * 1.4426950408889634f : 0.f;
            const uint64_t hidx = pair_base(g, x.tile * 2 + blk, x.h) + (uint64_t)(m * 32 + 8 * cq);
            const uint32_t hash_lo = (uint32_t)hidx + (uint32_t)(seed >> 32);
            const uint32_t hash_hi = ((uint32_t)(hidx >> 32) * 0x85EBCA77u) ^ (uint32_t)seed ^ (p.drop.site * 0xC2B2AE3Du);
            const uint32_t vm = x.vm;
            mbar_wait(&bars->s_full, (uint32_t)it & 1);
            if (threadIdx.x == 0) AB_T(it, 3);
            tc_fence_after();
            float sp[16], dp[16];
            {
                uint32_t a[16], b[16];
                tmem_ld_32x16(s.tmem + lane_addr + COL_P + 16 * cq, a);
                tmem_ld_32x16(s.tmem + lane_addr + COL_P + 64 + 16 * cq, b);
                tmem_ld_wait();
                tc_fence_before();
                warp_arrive(&bars->sdp_read, lane);               // the prologue columns may take the next recomputation
#pragma unroll
                for (int j = 0; j < 16; ++j) { sp[j] = __uint_as_float(a[j]); dp[j] = __uint_as_float(b[j]); }
            }
            uint32_t pk[8], dk[8];
            float D0 = 0.f, D1 = 0.f;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int j = 2 * jj;
                float f0 = 1.f, f1 = 1.f;
                if (p.drop.on()) {
                    uint32_t y = (hash_lo + (uint32_t)jj) * 0x9E3779B1u ^ hash_hi;
                    y ^= y >> 16; y *= 0x7feb352du; y ^= y >> 15; y *= 0x846ca68bu; y ^= y >> 16;
                    f0 = (y & 0xFFFFu) >= t16 ? p.drop.scale : 0.f;
                    f1 = (y >> 16) >= t16 ? p.drop.scale : 0.f;
                }
                float p0 = ex2_approx(fmaf(sp[j], sl2, -Lse)), p1 = ex2_approx(fmaf(sp[j + 1], sl2, -Lse));
                if (vm != 0xFFFFu) {
                    p0 = (vm >> j) & 1u ? p0 : 0.f;
                    p1 = (vm >> (j + 1)) & 1u ? p1 : 0.f;
                }
                const float pf0 = p0 * f0, pf1 = p1 * f1;
                D0 = fmaf(pf0, dp[j], D0); D1 = fmaf(pf1, dp[j + 1], D1);
                pk[jj] = pack_bf(pf0, pf1);
                sp[j] = p0 * g.scale; sp[j + 1] = p1 * g.scale;
                dp[j] *= f0; dp[j + 1] *= f1;
            }
            xch[cq * 128 + L] = D0 + D1;
            asm volatile("bar.sync 1, 512;" ::: "memory");
            const float Dp = (xch[L] + xch[128 + L]) + (xch[256 + L] + xch[384 + L]);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) dk[jj] = pack_bf(sp[2 * jj] * (dp[2 * jj] - Dp), sp[2 * jj + 1] * (dp[2 * jj + 1] - Dp));
            if (it >= 1) mbar_wait(&bars->stg_free, (uint32_t)(it - 1) & 1);   // the P~ / dS tiles staged dQ / dK of item it - 1: their stores have read them
            store_row16_packed(p_s + blk * AB_BLK, m, cq, pk);
            store_row16_packed(ds_s + blk * AB_BLK, m, cq, dk);
            fence_proxy_async();
            tc_fence_before();
            warp_arrive(&bars->p_full, lane);
            if (threadIdx.x == 0) AB_T(it, 4);
        };
        Item cur{0, (int64_t)blockIdx.x, -1, 0, 0u};
        if (n_items > 0) { locate(cur); convert(0); softmax(0, cur); }
        Item nxt = cur;
        for (int64_t it = 0; it < n_items; ++it) {
            const uint32_t ph = (uint32_t)it & 1;
            const int st = (int)(it & 1);
            if (threadIdx.x == 0) AB_T(it, 0);
            mbar_wait(&bars->o_full, ph);                         // the item's five contractions are complete: Q,K,V, P~, dS and its dO stage are dead
            if (threadIdx.x == 0) AB_T(it, 5);
            tc_fence_after();
            if (it + 1 < n_items) { advance(nxt); convert(it + 1); }
            // ---- epilogue of item it: dV, dK, dQ rows of `slot` -> bf16 -> staging tiles (TMA store) and the data gradient's A operand (TMEM) ----
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
            store_row16_packed(p_s, slot, cq, qv);
            store_row16_packed(ds_s, slot, cq, kv);
            store_row16_packed(do_s + (size_t)st * AB_T16, slot, cq, vv);
            fence_proxy_async();
            warp_arrive(&bars->stg_full, lane);
            // every thread has read its accumulator columns: the bf16 rows may overwrite X (lane L keeps holding row `slot`)
            tc_fence_before();
            asm volatile("bar.sync 2, 512;" ::: "memory");
            tc_fence_after();
            tmem_st_32x8(s.tmem + lane_addr + COL_X + 8 * cq, qv);
            tmem_st_32x8(s.tmem + lane_addr + COL_X + 32 + 8 * cq, kv);
            tmem_st_32x8(s.tmem + lane_addr + COL_X + 64 + 8 * cq, vv);
            tmem_st_wait();
            tc_fence_before();
            warp_arrive(&bars->a_ready, lane);
            if (threadIdx.x == 0) AB_T(it, 6);
            if (it + 1 < n_items) softmax(it + 1, nxt);           // (the data-gradient MMAs of item it run under it)
            if (cur.h == H - 1) {
                // ---- the tile's dh rows: accumulated over the heads -> fp32 -> global ----
                // (thread-per-row global stores cost one LSU wavefront per row and instruction -- ~5 k clks per tile and they starved the
                //  MMA issuer's MIO slots; staged 8 columns at a time and written by TMA instead)
                mbar_wait(&bars->dh_full, (uint32_t)cur.k & 1);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) {
                    const int64_t rd = cur.k * NCH + c;
                    uint32_t v[8];
                    tmem_ld_32x8(s.tmem + lane_addr + COL_DH + 32 * c + 8 * cq, v);   // round c: columns 32c .. 32c + 31 of the row, this thread's 8
                    tmem_ld_wait();
                    if (c == NCH - 1) { tc_fence_before(); warp_arrive(&bars->dh_free, lane); }   // the accumulator may take the next tile
                    if (rd > 0) mbar_wait(&bars->dhs_free, (uint32_t)(rd - 1) & 1);
                    store_row16_packed(dhs_s, slot, cq, v);       // 32 bytes of a SWIZZLE_128B row, exactly like a bf16 tile row
                    fence_proxy_async();
                    warp_arrive(&bars->dhs_full, lane);
                }
            }
            cur = nxt;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) { tc_fence_after(); tmem_dealloc(s.tmem, 512); }
}

// ---- host side ----
// view of an activation matrix [R, cols] whose box is one slot group (nbox == 2) or 128 consecutive rows (nbox == 1)
int act_tmap(CUtensorMap* m, const AttnGeom& g, const void* base, int64_t cols, int nbox, int box_cols, int swizzle, int elem_bytes = 2) {
    if (nbox == 1) {
        const int64_t dims[2] = {cols, g.n_seq * g.N}, strides[1] = {cols};
        const int box[2] = {box_cols, 128};
        return make_tmap_nd(m, base, elem_bytes, 2, dims, strides, box, swizzle);
    }
    if (g.gpb == 0) {
        const int64_t dims[4] = {cols, g.N, g.n_seq, 1}, strides[3] = {cols, (int64_t)g.N * cols, g.n_seq * g.N * cols};
        const int box[4] = {box_cols, g.N, g.G, 1};
        return make_tmap_nd(m, base, elem_bytes, 4, dims, strides, box, swizzle);
    }
    const int64_t dims[4] = {cols, g.inner, g.N, g.n_seq / g.inner};
    const int64_t strides[3] = {cols, (int64_t)g.inner * cols, (int64_t)g.N * g.inner * cols};
    const int box[4] = {box_cols, g.G, g.N, 1};
    return make_tmap_nd(m, base, elem_bytes, 4, dims, strides, box, swizzle);
}

long long* g_dbg = nullptr;
int dbg_level() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MSST_AB_DBG"); v = e ? atoi(e) : 0; if (v) cudaMalloc(&g_dbg, 48 * 16 * 8); }
    return v;
}
void dbg_dump(const char* what, cudaStream_t st, int& count) {
    if (++count != dbg_level()) return;
    static long long h[48 * 16];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, g_dbg, sizeof(h), cudaMemcpyDeviceToHost);
    const long long t0 = h[0];
    printf("%s timeline, CTA 0 (clks rel. to item 0), compute slots 0-6 | MMA-issuer slots 8-12 (see the AB_T calls)\n", what);
    for (int i = 0; i < 40; ++i) {
        printf("%2d:", i);
        for (int k = 0; k < 16; ++k) { if (k == 7) { printf("  |"); continue; } printf(" %7lld", h[i * 16 + k] ? h[i * 16 + k] - t0 : -1); }
        printf("\n");
    }
    fflush(stdout);
}

int fill_params(AbParams& p, const AttnGeom& g, int D, bool bwd, size_t& smem_bytes) {
    p.g = g; p.D = D; p.nch = D / 32;
    p.dbg = dbg_level() ? g_dbg : nullptr;
    p.nbox = (g.gpb == 0 && g.G * g.N == 64) ? 1 : 2;
    p.n_tiles = (g.groups + 1) / 2;
    p.slot_bytes = (uint32_t)D * 128u;
    p.hbuf_bytes = (uint32_t)p.nch * 8192u;
    // forward: h [2], Q,K,V double-buffered (6 tiles) + P~ + 2 x [4][128] statistics; backward: h [1], Q,K,V + dO [2] + P~ + dS + dh staging + [4][128] partial sums
    p.hbufs = bwd ? 1 : 2;
    const size_t fixed = (size_t)p.hbufs * p.hbuf_bytes + (bwd ? 8 * AB_T16 + 4 * 128 * sizeof(float) : 7 * AB_T16 + 2 * 4 * 128 * sizeof(float)) + sizeof(AbBars) +
                         (bwd ? 0 : 16 + 3 * 96 * sizeof(float));   // forward: + b_out | ln_w | ln_b of the fused tail
    const size_t cap = 232448;   // 227 KB of dynamic shared memory per CTA
    int nr = AB_MAXR;
    const int nr_min = bwd ? 6 : 3;   // backward: the data-gradient blocks of item i and the recomputation blocks of item i + 2 are consumed back to back
    while (nr > nr_min && fixed + (size_t)nr * p.slot_bytes > cap) --nr;
    MSST_REQUIRE(fixed + (size_t)nr * p.slot_bytes <= cap, "attn_block: D = %d does not fit shared memory", D);
    p.NR = nr;
    smem_bytes = fixed + (size_t)nr * p.slot_bytes;
    return MSST_OK;
}

}  // namespace

bool attn_block_supported(const AttnGeom& g, int D) {
    return attention_bwd_tc_supported(g) && g.dh == 64 && D % 32 == 0 && D >= 32 && D <= 96 && g.H * 64 * 3 < 65536;   // D = 128: the backward's ring does not fit
}

template <int NCH, bool OUT>
static int launch_fwd(const CUtensorMap& t_h, const CUtensorMap& t_w, const CUtensorMap& t_o, const CUtensorMap& t_wo, const CUtensorMap& t_xmid,
                      const CUtensorMap& t_h2, const CUtensorMap& t_x, const AbParams& p, int grid, size_t smem, cudaStream_t st) {
    static PerDeviceOnce once;
    if (once.first()) MSST_CUDA(cudaFuncSetAttribute(attn_block_fwd_kernel<NCH, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attn_block_fwd_kernel<NCH, OUT><<<grid, OUT ? AB_THREADS_OUT : AB_THREADS, smem, st>>>(t_h, t_w, t_o, t_wo, t_xmid, t_h2, t_x, p);
    return MSST_OK;
}

int attn_block_fwd(const AttnGeom& g, int D, const bf16* h, const bf16* w_qkv, bf16* out, float* lse, Drop drop, cudaStream_t st, const AttnBlockOut* tail) {
    MSST_REQUIRE(attn_block_supported(g, D), "attn_block_fwd: needs packed short sequences (N <= 64), dim_head 64, D in {32, 64, 96}");
    AbParams p{};
    size_t smem = 0;
    if (int rc = fill_params(p, g, D, false, smem)) return rc;
    p.lse_out = lse; p.o = out; p.drop = drop;
    CUtensorMap t_h, t_w, t_o, t_wo, t_xmid, t_h2, t_x;
    const int64_t I = (int64_t)g.H * 64;
    if (int rc = act_tmap(&t_h, g, h, D, p.nbox, 32, 64)) return rc;
    if (out) { if (int rc = act_tmap(&t_o, g, out, I, p.nbox, 64, 128)) return rc; }
    else t_o = t_h;
    const int64_t I3 = I * 3;
    { const int64_t dims[2] = {D, I3}, strides[1] = {D}; const int box[2] = {32, 64};
      if (int rc = make_tmap_bf16_nd(&t_w, w_qkv, 2, dims, strides, box, 64)) return rc; }
    if (tail) {
        MSST_REQUIRE(g.n_seq * g.N < (int64_t)2147483647, "attn_block_fwd: the fused out-projection indexes rows with 32 bits");
        MSST_REQUIRE(tail->w_out && tail->b_out && tail->x && tail->xmid && tail->ln_w && tail->ln_b && tail->h2 && tail->ln_stats,
                     "attn_block_fwd: the fused out-projection needs w_out, b_out, x, xmid, ln_w, ln_b, h2 and ln_stats");
        p.x_res = tail->x; p.b_out = tail->b_out; p.ln_w = tail->ln_w; p.ln_b = tail->ln_b; p.ln_stats = tail->ln_stats; p.drop_out = tail->drop;
        { static int v = -1; if (v < 0) { const char* e = getenv("MSST_AB_TAILDBG"); v = e ? atoi(e) : 0; } p.tail_dbg = v; }
        { const int64_t dims[2] = {I, D}, strides[1] = {I}; const int box[2] = {64, D};
          if (int rc = make_tmap_bf16_nd(&t_wo, tail->w_out, 2, dims, strides, box, 128)) return rc; }
        if (int rc = act_tmap(&t_xmid, g, tail->xmid, D, p.nbox, 32, 128, 4)) return rc;
        if (int rc = act_tmap(&t_x, g, tail->x, D, p.nbox, 32, 128, 4)) return rc;
        if (int rc = act_tmap(&t_h2, g, tail->h2, D, p.nbox, 32, 64)) return rc;
    } else {
        MSST_REQUIRE(out, "attn_block_fwd: null output");
        t_wo = t_h; t_xmid = t_h; t_h2 = t_h; t_x = t_h;
    }
    const int grid = (int)(p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs);
    if (p.dbg) cudaMemsetAsync(p.dbg, 0, 48 * 16 * 8, st);
    int rc = MSST_OK;
    switch (p.nch * 2 + (tail ? 1 : 0)) {
        case 2: rc = launch_fwd<1, false>(t_h, t_w, t_o, t_wo, t_xmid, t_h2, t_x, p, grid, smem, st); break;
        case 3: rc = launch_fwd<1, true>(t_h, t_w, t_o, t_wo, t_xmid, t_h2, t_x, p, grid, smem, st); break;
        case 4: rc = launch_fwd<2, false>(t_h, t_w, t_o, t_wo, t_xmid, t_h2, t_x, p, grid, smem, st); break;
        case 5: rc = launch_fwd<2, true>(t_h, t_w, t_o, t_wo, t_xmid, t_h2, t_x, p, grid, smem, st); break;
        case 6: rc = launch_fwd<3, false>(t_h, t_w, t_o, t_wo, t_xmid, t_h2, t_x, p, grid, smem, st); break;
        default: rc = launch_fwd<3, true>(t_h, t_w, t_o, t_wo, t_xmid, t_h2, t_x, p, grid, smem, st); break;
    }
    if (rc) return rc;
    MSST_LAUNCH_CHECK();
    if (p.dbg) { static int n = 0; dbg_dump("attn_block_fwd", st, n); }
    return MSST_OK;
}

int attn_block_bwd(const AttnGeom& g, int D, const bf16* h, const bf16* w_qkv, const bf16* w_qkv_t, const bf16* d_out, const float* lse,
                   bf16* d_qkv, float* d_h, Drop drop, cudaStream_t st) {
    MSST_REQUIRE(attn_block_supported(g, D), "attn_block_bwd: needs packed short sequences (N <= 64), dim_head 64, D in {32, 64, 96}");
    AbParams p{};
    size_t smem = 0;
    if (int rc = fill_params(p, g, D, true, smem)) return rc;
    p.lse_in = lse; p.dqkv = d_qkv; p.dh = d_h; p.drop = drop;
    static PerDeviceOnce once;
    if (once.first()) {
        MSST_CUDA(cudaFuncSetAttribute(attn_block_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        MSST_CUDA(cudaFuncSetAttribute(attn_block_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        MSST_CUDA(cudaFuncSetAttribute(attn_block_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    }
    CUtensorMap t_h, t_wt, t_do, t_dqkv, t_dh;
    (void)w_qkv;   // q / k / v are recomputed from the transposed copy (one weight tile serves the recomputation and the data gradient)
    const int64_t I = (int64_t)g.H * 64, I3 = 3 * I;
    if (int rc = act_tmap(&t_h, g, h, D, p.nbox, 32, 64)) return rc;
    if (int rc = act_tmap(&t_do, g, d_out, I, p.nbox, 64, 128)) return rc;
    if (int rc = act_tmap(&t_dqkv, g, d_qkv, I3, p.nbox, 64, 128)) return rc;
    if (int rc = act_tmap(&t_dh, g, d_h, D, p.nbox, 32, 128, 4)) return rc;
    { const int64_t dims[2] = {I3, D}, strides[1] = {I3}; const int box[2] = {64, D};
      if (int rc = make_tmap_bf16_nd(&t_wt, w_qkv_t, 2, dims, strides, box, 128)) return rc; }
    const int grid = (int)(p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs);
    if (p.dbg) cudaMemsetAsync(p.dbg, 0, 48 * 16 * 8, st);
    switch (p.nch) {
        case 1: attn_block_bwd_kernel<1><<<grid, AB_THREADS, smem, st>>>(t_h, t_wt, t_do, t_dqkv, t_dh, p); break;
        case 2: attn_block_bwd_kernel<2><<<grid, AB_THREADS, smem, st>>>(t_h, t_wt, t_do, t_dqkv, t_dh, p); break;
        default: attn_block_bwd_kernel<3><<<grid, AB_THREADS, smem, st>>>(t_h, t_wt, t_do, t_dqkv, t_dh, p); break;
    }
    MSST_LAUNCH_CHECK();
    if (p.dbg) { static int n = 0; dbg_dump("attn_block_bwd", st, n); }
    return MSST_OK;
}

}  // namespace msst
