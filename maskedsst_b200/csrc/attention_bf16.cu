// Kernel (3), bf16 mode: fused flash-style attention forward / backward on tensor cores.
// softmax(q k^T * dh^-0.5) v per (sequence, head); dh = 64; q/k/v/o bf16, statistics + accumulation fp32.
// Reference: Attention.forward, src/vit_spatial_spectral.py:67-78 (and its autograd).
//
// Same tiling / packing geometry as attention_f32.cu (64 query slots x 64 key slots, short sequences packed with a
// block-diagonal mask, strided row addressing for the spectral stack).  Each warp owns 16 query rows (FlashAttention-2
// register layout): S and P never leave registers, O / dQ accumulate in registers; the model's tiles are 64x64x64 -- too
// small for a tcgen05 pipeline to pay off (one UMMA M=64 tile per CTA), so the contractions use
// mma.sync.m16n8k16 with ldmatrix-fed fragments.  Scores never reach HBM.  Three kernel families:
//   N <= 64 : *_heads_kernel  -- one CTA per slot group loops over all heads, cp.async double buffer across heads
//   N  > 64 : *_long_kernel   -- forward: 128 query rows per CTA, K/V tiles streamed through a cp.async double buffer;
//                                backward: a dQ pass over key tiles and a dK/dV pass over query tiles (no atomics)
// Algorithmic HBM bytes per (slot, head): fwd 3*128 B in + 128 B out + 4 B lse; bwd 5*128 B in + 3*128 B out.
#include "common.cuh"
#include "kernels.h"
#include "attn_geom.cuh"
#include <stdlib.h>

namespace msst {

constexpr int BT = 128;         // threads per CTA (4 warps)
// smem tiles are [64 rows][64 bf16] with the 16-byte chunks of a row XOR-swizzled by (row & 7): conflict-free ldmatrix /
// cp.async / fragment stores without padding (8 KB per tile -> the backward kernel fits three CTAs per SM)
// Two layouts: padded rows (cheapest addressing; forward kernels) and XOR-swizzled (no padding: 8 KB tiles; backward).
struct Pad { static constexpr int TILE = TS * 72; static __device__ __forceinline__ int at(int r, int c) { return r * 72 + c; } };
struct Swz { static constexpr int TILE = TS * 64;
             static __device__ __forceinline__ int at(int r, int c) { return r * 64 + ((((c >> 3) ^ r) & 7) << 3) + (c & 7); } };
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const bf16* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const bf16* p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// C[16 x 64] (8 n-tiles) += A[16 rows of As starting at row0][64] . B^T, B stored [n][k] in Bs (rows = n, 64 k each)
template <class L>
__device__ __forceinline__ void gemm_a_bnk(float (&c)[8][4], const bf16* As, int row0, const bf16* Bs, int lane) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[4];
        ldsm_x4(a, As + L::at(row0 + (lane & 7) + 8 * ((lane >> 3) & 1), ks * 16 + 8 * (lane >> 4)));
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b[4];
            ldsm_x4(b, Bs + L::at(np * 16 + (lane & 7) + 8 * (lane >> 4), ks * 16 + 8 * ((lane >> 3) & 1)));
            mma16816(c[2 * np], a, b[0], b[1]);
            mma16816(c[2 * np + 1], a, b[2], b[3]);
        }
    }
}
// C[16 x 64] += P[16 x 64 (regs, C-fragment layout)] . B, B stored [k][n] in Bs (rows = k, 64 n each)
template <class L>
__device__ __forceinline__ void gemm_p_bkn(float (&c)[8][4], const float (&p)[8][4], const bf16* Bs, int lane) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t a[4] = {pack2(p[2 * j][0], p[2 * j][1]), pack2(p[2 * j][2], p[2 * j][3]),
                         pack2(p[2 * j + 1][0], p[2 * j + 1][1]), pack2(p[2 * j + 1][2], p[2 * j + 1][3])};
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b[4];
            ldsm_x4_t(b, Bs + L::at(j * 16 + (lane & 7) + 8 * ((lane >> 3) & 1), np * 16 + 8 * (lane >> 4)));
            mma16816(c[2 * np], a, b[0], b[1]);
            mma16816(c[2 * np + 1], a, b[2], b[3]);
        }
    }
}
// C[16 x 64] += A^T . B with A stored [k][m] in As (this warp's m = col0..col0+15), B stored [k][n] in Bs; k = 0..63
template <class L>
__device__ __forceinline__ void gemm_at_bkn(float (&c)[8][4], const bf16* As, int col0, const bf16* Bs, int lane) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[4];
        ldsm_x4_t(a, As + L::at(ks * 16 + (lane & 7) + 8 * (lane >> 4), col0 + 8 * ((lane >> 3) & 1)));
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b[4];
            ldsm_x4_t(b, Bs + L::at(ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1), np * 16 + 8 * (lane >> 4)));
            mma16816(c[2 * np], a, b[0], b[1]);
            mma16816(c[2 * np + 1], a, b[2], b[3]);
        }
    }
}

// attention-probability dropout: one 32-bit hash decides two neighbouring key slots (16 bits each)
__device__ __forceinline__ uint32_t pair_hash(const Drop& d, uint64_t idx) {
    const uint64_t s = d.seed + (d.seed_dev ? __ldg(d.seed_dev) : 0ull);
    // one round of the lowbias32 integer hash over (index, seed, site): full avalanche, ~12 instructions per key pair
    uint32_t x = ((uint32_t)idx + (uint32_t)(s >> 32)) * 0x9E3779B1u ^ ((uint32_t)(idx >> 32) * 0x85EBCA77u) ^ (uint32_t)s ^ (d.site * 0xC2B2AE3Du);
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ void pair_factors(const Drop& d, uint64_t idx, float& f0, float& f1) {
    const uint32_t h = pair_hash(d, idx), t16 = d.thresh >> 16;
    f0 = (h & 0xFFFFu) >= t16 ? d.scale : 0.f;
    f1 = (h >> 16) >= t16 ? d.scale : 0.f;
}
__device__ __forceinline__ uint64_t tile_pair_base(const AttnGeom& g, int64_t group, int h, int qt, int kt) {
    return ((((uint64_t)group * g.H + h) * g.tiles + qt) * g.tiles + kt) * (uint64_t)(TS * TS / 2);
}


// =========================================================================================================
// Single-tile (N <= 64) fast path: one CTA per slot group loops over ALL heads; the Q/K/V(/dO) tiles of head h+1 are
// prefetched with cp.async into a second buffer while head h is computed, and everything that only depends on the
// slot geometry (row addresses, block-diagonal mask bits) is computed once per CTA instead of once per (CTA, head).
// =========================================================================================================
__device__ __forceinline__ void cp_async16(bf16* dst, const bf16* src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int sz = valid ? 16 : 0;   // src-size 0 -> the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct HeadsShared {
    int64_t row[TS];     // global row of each slot (or -1)
    int seq[TS];         // local sequence id of each slot (or -1)
    int pos[TS];
};

__device__ __forceinline__ void heads_setup(const AttnGeom& g, int64_t group, HeadsShared& hs) {
    if (threadIdx.x < TS) {
        int64_t seq; int pos;
        const bool ok = slot_to(g, group, 0, threadIdx.x, seq, pos);
        hs.seq[threadIdx.x] = ok ? (int)(seq - group_seq0(g, group)) : -1;
        hs.pos[threadIdx.x] = pos;
        hs.row[threadIdx.x] = ok ? row_of(g, seq, pos) : -1;
    }
}
// this thread's 4 chunks of a [64 x 64] tile: rows (tid>>3) + 16k, column chunk (tid&7)*8
template <class L>
__device__ __forceinline__ void prefetch_tile(bf16* dst, const bf16* __restrict__ base, int64_t ld, int col0, const HeadsShared& hs) {
    const int c = (threadIdx.x & 7) * 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = (threadIdx.x >> 3) + 16 * k;
        const int64_t row = hs.row[r];
        cp_async16(dst + L::at(r, c), base + (row < 0 ? 0 : row) * ld + col0 + c, row >= 0);
    }
}
template <class L>
__device__ __forceinline__ void store_rows16_hs(const bf16* src, int row0, bf16* __restrict__ base, int64_t ld, int col0,
                                                const HeadsShared& hs, int lane) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = lane + 32 * k;
        const int r = row0 + (i >> 3), c = (i & 7) * 8;
        const int64_t row = hs.row[r];
        if (row >= 0) *reinterpret_cast<uint4*>(base + row * ld + col0 + c) = *reinterpret_cast<const uint4*>(src + L::at(r, c));
    }
}
// bit (4*nt + e) set when element e of n-tile nt of this thread's C fragment pairs a query and a key of the same sequence
__device__ __forceinline__ uint32_t fragment_mask(const HeadsShared& hs, int r0, int r1, int tq) {
    const int qs0 = hs.seq[r0], qs1 = hs.seq[r1];
    uint32_t m = 0;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int c = nt * 8 + 2 * tq;
        const int ks0 = hs.seq[c], ks1 = hs.seq[c + 1];
        m |= (uint32_t)(qs0 >= 0 && qs0 == ks0) << (4 * nt);
        m |= (uint32_t)(qs0 >= 0 && qs0 == ks1) << (4 * nt + 1);
        m |= (uint32_t)(qs1 >= 0 && qs1 == ks0) << (4 * nt + 2);
        m |= (uint32_t)(qs1 >= 0 && qs1 == ks1) << (4 * nt + 3);
    }
    return m;
}

constexpr size_t kFwdHeadsSmem = sizeof(bf16) * 2 * 3 * Pad::TILE + sizeof(HeadsShared);
constexpr size_t kBwdHeadsSmem = sizeof(bf16) * (2 * 4 + 1) * Swz::TILE + sizeof(HeadsShared);

__global__ void __launch_bounds__(BT, 4) attn_fwd_bf16_heads_kernel(AttnGeom g, const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                                    float* __restrict__ lse, Drop drop) {
    using L = Pad;
    extern __shared__ __align__(16) uint8_t smem_h[];
    bf16* buf = reinterpret_cast<bf16*>(smem_h);                                   // [2][3][TILE]
    HeadsShared& hs = *reinterpret_cast<HeadsShared*>(smem_h + sizeof(bf16) * 6 * L::TILE);
    const int I = g.H * 64;
    const int64_t ld = 3 * (int64_t)I, group = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
    const int r0 = warp * 16 + gq, r1 = r0 + 8;
    const float sl2 = g.scale * 1.4426950408889634f;
    heads_setup(g, group, hs);
    __syncthreads();
    const uint32_t mask = fragment_mask(hs, r0, r1, tq);
    const bool full_tile = __syncthreads_and(mask == 0xFFFFFFFFu);   // e.g. the spatial stack (one 64-token sequence per tile)
    const int64_t grow0 = hs.row[r0], grow1 = hs.row[r1];
    auto prefetch = [&](int h, int b) {
        bf16* t = buf + (size_t)b * 3 * L::TILE;
        prefetch_tile<L>(t, qkv, ld, h * 64, hs);
        prefetch_tile<L>(t + L::TILE, qkv, ld, I + h * 64, hs);
        prefetch_tile<L>(t + 2 * L::TILE, qkv, ld, 2 * I + h * 64, hs);
        cp_async_commit();
    };
    prefetch(0, 0);
    for (int h = 0; h < g.H; ++h) {
        const int b = h & 1;
        cp_async_wait_all();
        __syncthreads();
        if (h + 1 < g.H) prefetch(h + 1, b ^ 1);
        bf16* Qs = buf + (size_t)b * 3 * L::TILE; bf16* Ks = Qs + L::TILE; bf16* Vs = Ks + L::TILE;
        float s[8][4] = {};
        gemm_a_bnk<L>(s, Qs, warp * 16, Ks, lane);
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) s[nt][e] = (full_tile || ((mask >> (4 * nt + e)) & 1u)) ? s[nt][e] * sl2 : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1])); mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float sub0 = mx0 == -INFINITY ? 0.f : mx0, sub1 = mx1 == -INFINITY ? 0.f : mx1;
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = exp2f(s[nt][0] - sub0); s[nt][1] = exp2f(s[nt][1] - sub0);
            s[nt][2] = exp2f(s[nt][2] - sub1); s[nt][3] = exp2f(s[nt][3] - sub1);
            l0 += s[nt][0] + s[nt][1]; l1 += s[nt][2] + s[nt][3];
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
        if (drop.on()) {
            const uint64_t base = tile_pair_base(g, group, h, 0, 0);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                float f0, f1;
                pair_factors(drop, base + (uint64_t)(r0 * 32 + nt * 4 + tq), f0, f1); s[nt][0] *= f0 * i0; s[nt][1] *= f1 * i0;
                pair_factors(drop, base + (uint64_t)(r1 * 32 + nt * 4 + tq), f0, f1); s[nt][2] *= f0 * i1; s[nt][3] *= f1 * i1;
            }
        } else {
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) { s[nt][0] *= i0; s[nt][1] *= i0; s[nt][2] *= i1; s[nt][3] *= i1; }
        }
        float o[8][4] = {};
        gemm_p_bkn<L>(o, s, Vs, lane);
        // stage this warp's 16 output rows in its own Q rows (only this warp reads them), then coalesced stores
        __syncwarp();
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            *reinterpret_cast<uint32_t*>(Qs + L::at(r0, nt * 8 + 2 * tq)) = pack2(o[nt][0], o[nt][1]);
            *reinterpret_cast<uint32_t*>(Qs + L::at(r1, nt * 8 + 2 * tq)) = pack2(o[nt][2], o[nt][3]);
        }
        __syncwarp();
        store_rows16_hs<L>(Qs, warp * 16, out, I, h * 64, hs, lane);
        if (tq == 0) {
            if (grow0 >= 0) lse[grow0 * g.H + h] = (mx0 + log2f(l0)) * 0.6931471805599453f;
            if (grow1 >= 0) lse[grow1 * g.H + h] = (mx1 + log2f(l1)) * 0.6931471805599453f;
        }
    }
}

__global__ void __launch_bounds__(BT, 3) attn_bwd_bf16_heads_kernel(AttnGeom g, const bf16* __restrict__ qkv, const float* __restrict__ lse,
                                                                    const bf16* __restrict__ d_out, bf16* __restrict__ d_qkv, Drop drop) {
    using L = Swz;
    extern __shared__ __align__(16) uint8_t smem_h[];
    bf16* buf = reinterpret_cast<bf16*>(smem_h);                                   // [2][4][TILE]: Q K V dO
    bf16* dSs = buf + (size_t)8 * L::TILE;
    HeadsShared& hs = *reinterpret_cast<HeadsShared*>(smem_h + sizeof(bf16) * 9 * L::TILE);
    const int I = g.H * 64;
    const int64_t ld = 3 * (int64_t)I, group = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
    const int r0 = warp * 16 + gq, r1 = r0 + 8;
    const float sl2 = g.scale * 1.4426950408889634f;
    heads_setup(g, group, hs);
    __syncthreads();
    const uint32_t mask = fragment_mask(hs, r0, r1, tq);
    const bool full_tile = __syncthreads_and(mask == 0xFFFFFFFFu);
    const int64_t grow0 = hs.row[r0], grow1 = hs.row[r1];
    auto prefetch = [&](int h, int b) {
        bf16* t = buf + (size_t)b * 4 * L::TILE;
        prefetch_tile<L>(t, qkv, ld, h * 64, hs);
        prefetch_tile<L>(t + L::TILE, qkv, ld, I + h * 64, hs);
        prefetch_tile<L>(t + 2 * L::TILE, qkv, ld, 2 * I + h * 64, hs);
        prefetch_tile<L>(t + 3 * L::TILE, d_out, I, h * 64, hs);
        cp_async_commit();
    };
    prefetch(0, 0);
    for (int h = 0; h < g.H; ++h) {
        const int b = h & 1;
        cp_async_wait_all();
        __syncthreads();
        if (h + 1 < g.H) prefetch(h + 1, b ^ 1);
        bf16* Qs = buf + (size_t)b * 4 * L::TILE; bf16* Ks = Qs + L::TILE; bf16* Vs = Ks + L::TILE; bf16* dOs = Vs + L::TILE;
        const float L0 = grow0 >= 0 ? lse[grow0 * g.H + h] * 1.4426950408889634f : 0.f;
        const float L1 = grow1 >= 0 ? lse[grow1 * g.H + h] * 1.4426950408889634f : 0.f;
        float s[8][4] = {}, dp[8][4] = {};
        gemm_a_bnk<L>(s, Qs, warp * 16, Ks, lane);
        gemm_a_bnk<L>(dp, dOs, warp * 16, Vs, lane);
        // P, Pf = P*f (kept in dp[]), D_i = sum_j Pf_ij dP_ij  (== dO_i . O_i), dS (kept in s[])
        float D0 = 0.f, D1 = 0.f;
        uint32_t pf_pack[8][2];
        const uint64_t base = tile_pair_base(g, group, h, 0, 0);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float f[4] = {1.f, 1.f, 1.f, 1.f};
            if (drop.on()) {
                pair_factors(drop, base + (uint64_t)(r0 * 32 + nt * 4 + tq), f[0], f[1]);
                pair_factors(drop, base + (uint64_t)(r1 * 32 + nt * 4 + tq), f[2], f[3]);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float p = (full_tile || ((mask >> (4 * nt + e)) & 1u)) ? exp2f(s[nt][e] * sl2 - (e < 2 ? L0 : L1)) : 0.f;
                const float pf = p * f[e];
                const float t = pf * dp[nt][e];
                if (e < 2) D0 += t; else D1 += t;
                s[nt][e] = p;          // P
                dp[nt][e] = f[e] * dp[nt][e];   // f * dP
                f[e] = pf;
            }
            // f[] now holds Pf for this n-tile; keep it as bf16 pairs for the smem write after barrier (A)
            pf_pack[nt][0] = pack2(f[0], f[1]);
            pf_pack[nt][1] = pack2(f[2], f[3]);
        }
        D0 += __shfl_xor_sync(0xffffffffu, D0, 1); D0 += __shfl_xor_sync(0xffffffffu, D0, 2);
        D1 += __shfl_xor_sync(0xffffffffu, D1, 1); D1 += __shfl_xor_sync(0xffffffffu, D1, 2);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = s[nt][0] * (dp[nt][0] - D0) * g.scale; s[nt][1] = s[nt][1] * (dp[nt][1] - D0) * g.scale;
            s[nt][2] = s[nt][2] * (dp[nt][2] - D1) * g.scale; s[nt][3] = s[nt][3] * (dp[nt][3] - D1) * g.scale;
        }
        __syncthreads();   // (A) every warp has finished reading V: its region now receives Pf
        bf16* Ps = Vs;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int c = nt * 8 + 2 * tq;
            *reinterpret_cast<uint32_t*>(Ps + L::at(r0, c)) = pf_pack[nt][0];
            *reinterpret_cast<uint32_t*>(Ps + L::at(r1, c)) = pf_pack[nt][1];
            *reinterpret_cast<uint32_t*>(dSs + L::at(r0, c)) = pack2(s[nt][0], s[nt][1]);
            *reinterpret_cast<uint32_t*>(dSs + L::at(r1, c)) = pack2(s[nt][2], s[nt][3]);
        }
        float dq[8][4] = {};
        gemm_p_bkn<L>(dq, s, Ks, lane);                       // dQ[16 rows] = dS . K
        __syncthreads();   // (B) Pf / dS tiles complete
        float dv[8][4] = {}, dk[8][4] = {};
        gemm_at_bkn<L>(dv, Ps, warp * 16, dOs, lane);         // dV[16 keys] = (P f)^T . dO
        gemm_at_bkn<L>(dk, dSs, warp * 16, Qs, lane);         // dK[16 keys] = dS^T . Q
        __syncthreads();   // (C) Q / K / Pf tiles are dead: reuse them as staging for coalesced stores
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int c = nt * 8 + 2 * tq;
            *reinterpret_cast<uint32_t*>(Qs + L::at(r0, c)) = pack2(dq[nt][0], dq[nt][1]);
            *reinterpret_cast<uint32_t*>(Qs + L::at(r1, c)) = pack2(dq[nt][2], dq[nt][3]);
            *reinterpret_cast<uint32_t*>(Ks + L::at(r0, c)) = pack2(dk[nt][0], dk[nt][1]);
            *reinterpret_cast<uint32_t*>(Ks + L::at(r1, c)) = pack2(dk[nt][2], dk[nt][3]);
            *reinterpret_cast<uint32_t*>(Vs + L::at(r0, c)) = pack2(dv[nt][0], dv[nt][1]);
            *reinterpret_cast<uint32_t*>(Vs + L::at(r1, c)) = pack2(dv[nt][2], dv[nt][3]);
        }
        __syncwarp();
        store_rows16_hs<L>(Qs, warp * 16, d_qkv, ld, h * 64, hs, lane);
        store_rows16_hs<L>(Ks, warp * 16, d_qkv, ld, I + h * 64, hs, lane);
        store_rows16_hs<L>(Vs, warp * 16, d_qkv, ld, 2 * I + h * 64, hs, lane);
    }
}


// =========================================================================================================
// Long sequences (N > 64) forward: one CTA = 128 query rows (8 warps) of one (sequence, head); the 64-key K/V tiles are
// streamed through a cp.async double buffer (tile kt+1 lands while tile kt is multiplied), online softmax in registers.
// =========================================================================================================
constexpr int LQ = 128, LT = 256;
constexpr size_t kFwdLongSmem = sizeof(bf16) * (LQ * 72 + 4 * Pad::TILE);

__global__ void __launch_bounds__(LT, 2) attn_fwd_bf16_long_kernel(AttnGeom g, const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                                   float* __restrict__ lse, Drop drop) {
    using L = Pad;
    extern __shared__ __align__(16) uint8_t smem_l[];
    bf16* Qs = reinterpret_cast<bf16*>(smem_l);            // [128][72]
    bf16* KV = Qs + LQ * 72;                                // [2 buffers][K, V][64][72]
    const int I = g.H * 64, h = blockIdx.y;
    const int64_t ld = 3 * (int64_t)I;
    const int qtiles = (g.N + LQ - 1) / LQ;
    const int64_t seq = blockIdx.x / qtiles;
    const int qt = blockIdx.x % qtiles;
    const int64_t row_base = (seq / g.inner) * g.N * g.inner + (seq % g.inner);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
    const int r0 = warp * 16 + gq, r1 = r0 + 8;
    const float sl2 = g.scale * 1.4426950408889634f;
    const int c8 = (threadIdx.x & 7) * 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) {                           // Q tile: 128 rows x 8 chunks
        const int r = (threadIdx.x >> 3) + 32 * k, pos = qt * LQ + r;
        const bool ok = pos < g.N;
        cp_async16(Qs + L::at(r, c8), qkv + (row_base + (int64_t)(ok ? pos : 0) * g.inner) * ld + h * 64 + c8, ok);
    }
    auto prefetch_kv = [&](int kt, int b) {
        bf16* Kb = KV + (size_t)b * 2 * L::TILE; bf16* Vb = Kb + L::TILE;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int r = (threadIdx.x >> 3) + 32 * k, pos = kt * TS + r;
            const bool ok = pos < g.N;
            const bf16* src = qkv + (row_base + (int64_t)(ok ? pos : 0) * g.inner) * ld + h * 64 + c8;
            cp_async16(Kb + L::at(r, c8), src + I, ok);
            cp_async16(Vb + L::at(r, c8), src + 2 * I, ok);
        }
        cp_async_commit();
    };
    prefetch_kv(0, 0);
    float o[8][4] = {};
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    for (int kt = 0; kt < g.tiles; ++kt) {
        const int b = kt & 1;
        cp_async_wait_all();
        __syncthreads();
        if (kt + 1 < g.tiles) prefetch_kv(kt + 1, b ^ 1);
        const bf16* Ks = KV + (size_t)b * 2 * L::TILE; const bf16* Vs = Ks + L::TILE;
        float s[8][4] = {};
        gemm_a_bnk<L>(s, Qs, warp * 16, Ks, lane);
        const int kvalid = g.N - kt * TS;                  // keys of this tile with column index < kvalid exist
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int c = nt * 8 + 2 * tq;
            s[nt][0] = c < kvalid ? s[nt][0] * sl2 : -INFINITY; s[nt][1] = c + 1 < kvalid ? s[nt][1] * sl2 : -INFINITY;
            s[nt][2] = c < kvalid ? s[nt][2] * sl2 : -INFINITY; s[nt][3] = c + 1 < kvalid ? s[nt][3] * sl2 : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1])); mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);     // every tile has >= 1 valid key: mn is finite
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = exp2f(s[nt][0] - mn0); s[nt][1] = exp2f(s[nt][1] - mn0);
            s[nt][2] = exp2f(s[nt][2] - mn1); s[nt][3] = exp2f(s[nt][3] - mn1);
            rs0 += s[nt][0] + s[nt][1]; rs1 += s[nt][2] + s[nt][3];
        }
        rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
        rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
        l0 = l0 * c0 + rs0; l1 = l1 * c1 + rs1; m0 = mn0; m1 = mn1;
        if (drop.on()) {   // same (64-row query tile, key tile, row, pair) indexing as the backward kernels
            const uint64_t base = tile_pair_base(g, seq, h, 2 * qt + (warp >> 2), kt);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                float f0, f1;
                pair_factors(drop, base + (uint64_t)((r0 & 63) * 32 + nt * 4 + tq), f0, f1); s[nt][0] *= f0; s[nt][1] *= f1;
                pair_factors(drop, base + (uint64_t)((r1 & 63) * 32 + nt * 4 + tq), f0, f1); s[nt][2] *= f0; s[nt][3] *= f1;
            }
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { o[nt][0] *= c0; o[nt][1] *= c0; o[nt][2] *= c1; o[nt][3] *= c1; }
        gemm_p_bkn<L>(o, s, Vs, lane);
    }
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {                        // stage through this warp's own Q rows
        *reinterpret_cast<uint32_t*>(Qs + L::at(r0, nt * 8 + 2 * tq)) = pack2(o[nt][0] * i0, o[nt][1] * i0);
        *reinterpret_cast<uint32_t*>(Qs + L::at(r1, nt * 8 + 2 * tq)) = pack2(o[nt][2] * i1, o[nt][3] * i1);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = lane + 32 * k;
        const int r = warp * 16 + (i >> 3), c = (i & 7) * 8, pos = qt * LQ + r;
        if (pos < g.N)
            *reinterpret_cast<uint4*>(out + (row_base + (int64_t)pos * g.inner) * I + h * 64 + c) = *reinterpret_cast<const uint4*>(Qs + L::at(r, c));
    }
    if (tq == 0) {
        const int p0 = qt * LQ + r0, p1 = qt * LQ + r1;
        if (p0 < g.N) lse[(row_base + (int64_t)p0 * g.inner) * g.H + h] = (m0 + log2f(l0)) * 0.6931471805599453f;
        if (p1 < g.N) lse[(row_base + (int64_t)p1 * g.inner) * g.H + h] = (m1 + log2f(l1)) * 0.6931471805599453f;
    }
}


// =========================================================================================================
// Long sequences backward, two passes without atomics:
//   MODE 1: CTA = one 64-row query tile, loops the key tiles (cp.async double buffer)  -> dQ
//   MODE 2: CTA = one 64-key tile, loops the query tiles (cp.async double buffer)      -> dK, dV
// D_i = dO_i . O_i is computed in-kernel per query tile (dO tile in smem, O rows from global / L2).
// =========================================================================================================
constexpr size_t kBwdLongSmem = sizeof(bf16) * 8 * Pad::TILE + sizeof(float) * 2 * TS;

template <int MODE>
__global__ void __launch_bounds__(BT, 2) attn_bwd_bf16_long_kernel(AttnGeom g, const bf16* __restrict__ qkv, const bf16* __restrict__ out,
                                                                   const float* __restrict__ lse, const bf16* __restrict__ d_out,
                                                                   bf16* __restrict__ d_qkv, Drop drop) {
    using L = Pad;
    extern __shared__ __align__(16) uint8_t smem_l[];
    bf16* F0 = reinterpret_cast<bf16*>(smem_l);     // fixed tiles: MODE 1 -> Q, dO ; MODE 2 -> K, V
    bf16* F1 = F0 + L::TILE;
    bf16* LB = F1 + L::TILE;                         // looped tiles [2 buffers][2]: MODE 1 -> K, V ; MODE 2 -> Q, dO
    bf16* Ps = LB + 4 * L::TILE;
    bf16* dSs = Ps + L::TILE;
    float* Drow = reinterpret_cast<float*>(dSs + L::TILE);
    float* lse_s = Drow + TS;
    const int I = g.H * 64, h = blockIdx.y;
    const int64_t ld = 3 * (int64_t)I;
    const int64_t seq = blockIdx.x / g.tiles;
    const int own = blockIdx.x % g.tiles;
    const int64_t row_base = (seq / g.inner) * g.N * g.inner + (seq % g.inner);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
    const int r0 = warp * 16 + gq, r1 = r0 + 8;
    const float sl2 = g.scale * 1.4426950408889634f;
    const int c8 = (threadIdx.x & 7) * 8;
    // tile loaders: which = 0 (q), 1 (k), 2 (v) columns of qkv, 3 = d_out
    auto load_tile = [&](bf16* dst, int which, int tile) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = (threadIdx.x >> 3) + 16 * k, pos = tile * TS + r;
            const bool ok = pos < g.N;
            const int64_t row = row_base + (int64_t)(ok ? pos : 0) * g.inner;
            const bf16* src = which == 3 ? d_out + row * I + h * 64 + c8 : qkv + row * ld + which * I + h * 64 + c8;
            cp_async16(dst + L::at(r, c8), src, ok);
        }
    };
    auto compute_D = [&](const bf16* dOs, int qt) {   // Drow / lse_s of query tile qt (dO tile already in smem)
        for (int k = 0; k < 8; ++k) {
            const int r = warp * 16 + k * 2 + (lane >> 4), c = (lane & 15) * 4, pos = qt * TS + r;
            float dsum = 0.f, l = 0.f;
            if (pos < g.N) {
                const int64_t row = row_base + (int64_t)pos * g.inner;
                const uint2 ov = *reinterpret_cast<const uint2*>(out + row * I + h * 64 + c);
                const __nv_bfloat162 o01 = *reinterpret_cast<const __nv_bfloat162*>(&ov.x), o23 = *reinterpret_cast<const __nv_bfloat162*>(&ov.y);
                const bf16* dp = dOs + L::at(r, c);
                dsum = __bfloat162float(dp[0]) * __low2float(o01) + __bfloat162float(dp[1]) * __high2float(o01) +
                       __bfloat162float(dp[2]) * __low2float(o23) + __bfloat162float(dp[3]) * __high2float(o23);
                l = lse[row * g.H + h];
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
            if ((lane & 15) == 0) { Drow[r] = dsum; lse_s[r] = l * 1.4426950408889634f; }
        }
    };
    if (MODE == 1) { load_tile(F0, 0, own); load_tile(F1, 3, own); load_tile(LB, 1, 0); load_tile(LB + L::TILE, 2, 0); }
    else           { load_tile(F0, 1, own); load_tile(F1, 2, own); load_tile(LB, 0, 0); load_tile(LB + L::TILE, 3, 0); }
    cp_async_commit();
    float acc0[8][4] = {}, acc1[8][4] = {};          // MODE 1: acc0 = dQ ; MODE 2: acc0 = dK, acc1 = dV
    for (int it = 0; it < g.tiles; ++it) {
        const int b = it & 1;
        cp_async_wait_all();
        __syncthreads();
        if (it + 1 < g.tiles) {
            bf16* nb = LB + (size_t)(b ^ 1) * 2 * L::TILE;
            if (MODE == 1) { load_tile(nb, 1, it + 1); load_tile(nb + L::TILE, 2, it + 1); }
            else           { load_tile(nb, 0, it + 1); load_tile(nb + L::TILE, 3, it + 1); }
            cp_async_commit();
        }
        const bf16* L0 = LB + (size_t)b * 2 * L::TILE; const bf16* L1 = L0 + L::TILE;
        const bf16* Qs = MODE == 1 ? F0 : L0; const bf16* dOs = MODE == 1 ? F1 : L1;
        const bf16* Ks = MODE == 1 ? L0 : F0; const bf16* Vs = MODE == 1 ? L1 : F1;
        const int qt = MODE == 1 ? own : it, kt = MODE == 1 ? it : own;
        if (MODE == 2 || it == 0) { compute_D(dOs, qt); __syncthreads(); }
        float s[8][4] = {}, dp[8][4] = {};
        gemm_a_bnk<L>(s, Qs, warp * 16, Ks, lane);
        gemm_a_bnk<L>(dp, dOs, warp * 16, Vs, lane);
        const float Lr0 = lse_s[r0], Lr1 = lse_s[r1], D0 = Drow[r0], D1 = Drow[r1];
        const bool q0 = qt * TS + r0 < g.N, q1 = qt * TS + r1 < g.N;
        const int kvalid = g.N - kt * TS;
        const uint64_t base = tile_pair_base(g, seq, h, qt, kt);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int c = nt * 8 + 2 * tq;
            const bool k0 = c < kvalid, k1 = c + 1 < kvalid;
            const float p0 = (q0 && k0) ? exp2f(s[nt][0] * sl2 - Lr0) : 0.f, p1 = (q0 && k1) ? exp2f(s[nt][1] * sl2 - Lr0) : 0.f;
            const float p2 = (q1 && k0) ? exp2f(s[nt][2] * sl2 - Lr1) : 0.f, p3 = (q1 && k1) ? exp2f(s[nt][3] * sl2 - Lr1) : 0.f;
            float f0 = 1.f, f1 = 1.f, f2 = 1.f, f3 = 1.f;
            if (drop.on()) {
                pair_factors(drop, base + (uint64_t)(r0 * 32 + nt * 4 + tq), f0, f1);
                pair_factors(drop, base + (uint64_t)(r1 * 32 + nt * 4 + tq), f2, f3);
            }
            s[nt][0] = p0 * (f0 * dp[nt][0] - D0) * g.scale; s[nt][1] = p1 * (f1 * dp[nt][1] - D0) * g.scale;
            s[nt][2] = p2 * (f2 * dp[nt][2] - D1) * g.scale; s[nt][3] = p3 * (f3 * dp[nt][3] - D1) * g.scale;
            if (MODE == 2) {
                *reinterpret_cast<uint32_t*>(Ps + L::at(r0, c)) = pack2(p0 * f0, p1 * f1);
                *reinterpret_cast<uint32_t*>(Ps + L::at(r1, c)) = pack2(p2 * f2, p3 * f3);
                *reinterpret_cast<uint32_t*>(dSs + L::at(r0, c)) = pack2(s[nt][0], s[nt][1]);
                *reinterpret_cast<uint32_t*>(dSs + L::at(r1, c)) = pack2(s[nt][2], s[nt][3]);
            }
        }
        if (MODE == 1) {
            gemm_p_bkn<L>(acc0, s, Ks, lane);                    // dQ[16 rows] += dS . K
        } else {
            __syncthreads();
            gemm_at_bkn<L>(acc1, Ps, warp * 16, dOs, lane);      // dV[16 keys] += (P f)^T . dO
            gemm_at_bkn<L>(acc0, dSs, warp * 16, Qs, lane);      // dK[16 keys] += dS^T . Q
        }
    }
    __syncthreads();
    // stage through the (dead) P / dS tiles and store coalesced: this warp's 16 rows of the CTA's own tile
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int c = nt * 8 + 2 * tq;
        *reinterpret_cast<uint32_t*>(Ps + L::at(r0, c)) = pack2(acc0[nt][0], acc0[nt][1]);
        *reinterpret_cast<uint32_t*>(Ps + L::at(r1, c)) = pack2(acc0[nt][2], acc0[nt][3]);
        if (MODE == 2) {
            *reinterpret_cast<uint32_t*>(dSs + L::at(r0, c)) = pack2(acc1[nt][0], acc1[nt][1]);
            *reinterpret_cast<uint32_t*>(dSs + L::at(r1, c)) = pack2(acc1[nt][2], acc1[nt][3]);
        }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = lane + 32 * k;
        const int r = warp * 16 + (i >> 3), c = (i & 7) * 8, pos = own * TS + r;
        if (pos >= g.N) continue;
        bf16* dst = d_qkv + (row_base + (int64_t)pos * g.inner) * ld + h * 64 + c;
        if (MODE == 1) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(Ps + L::at(r, c));
        else {
            *reinterpret_cast<uint4*>(dst + I) = *reinterpret_cast<const uint4*>(Ps + L::at(r, c));
            *reinterpret_cast<uint4*>(dst + 2 * I) = *reinterpret_cast<const uint4*>(dSs + L::at(r, c));
        }
    }
}

int attention_fwd_bf16(const msst_attn_dims* d, const bf16* qkv, bf16* out, float* lse, cudaStream_t st) {
    AttnGeom g;
    if (int rc = make_attn_geom(d, g, false)) return rc;
    if (g.n_seq == 0) return MSST_OK;
    const Drop drop = make_drop(d->drop_p, d->seed, d->site, d->seed_dev);
    if (g.tiles == 1) {
        static int use_tc = -1;
        if (use_tc < 0) { const char* e = getenv("MSST_ATTN_TC"); use_tc = e ? atoi(e) : 1; }   // default: tcgen05 kernel; 0 = mma.sync
        if (use_tc && attention_bwd_tc_supported(g)) return attention_fwd_tc(g, qkv, out, lse, drop, st);   // tcgen05 / TMEM forward (attention_tc.cu)
    }
    if (g.tiles == 1) {   // N <= 64: head-looping, cp.async double-buffered kernel
        static PerDeviceOnce attr_set;
        if (attr_set.first()) {
            MSST_CUDA(cudaFuncSetAttribute(attn_fwd_bf16_heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdHeadsSmem));
        }
        attn_fwd_bf16_heads_kernel<<<(unsigned)g.groups, BT, kFwdHeadsSmem, st>>>(g, qkv, out, lse, drop);
        MSST_LAUNCH_CHECK();
        return MSST_OK;
    }
    {   // N > 64: tcgen05 kernel (attention_tc.cu) from N = 512 on, where it beats the mma.sync kernel (318 vs 267 TF/s at N = 1024,
        // 337 vs 278 at 4096; on par at 256).  MSST_ATTN_TC_LONG=1 forces it for every N > 64, =0 disables it.
        static int use_tcl = -2;
        if (use_tcl == -2) { const char* e = getenv("MSST_ATTN_TC_LONG"); use_tcl = e ? atoi(e) : -1; }
        if ((use_tcl == 1 || (use_tcl == -1 && g.N >= 512)) && attention_fwd_tc_long_supported(g))
            return attention_fwd_tc_long(g, qkv, out, lse, drop, st);
    }
    {   // N > 64: 128-row query tiles, K/V streamed through a cp.async double buffer
        static PerDeviceOnce lattr;
        if (lattr.first()) {
            MSST_CUDA(cudaFuncSetAttribute(attn_fwd_bf16_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdLongSmem));
        }
        const int64_t qtiles = (g.N + LQ - 1) / LQ;
        MSST_REQUIRE(g.n_seq * qtiles < (int64_t)2147483647, "attention: grid too large");
        attn_fwd_bf16_long_kernel<<<dim3((unsigned)(g.n_seq * qtiles), g.H), LT, kFwdLongSmem, st>>>(g, qkv, out, lse, drop);
        MSST_LAUNCH_CHECK();
    }
    return MSST_OK;
}

int attention_bwd_bf16(const msst_attn_dims* d, const bf16* qkv, const bf16* out, const float* lse, const bf16* d_out, bf16* d_qkv,
                       cudaStream_t st) {
    AttnGeom g;
    if (int rc = make_attn_geom(d, g, false)) return rc;
    if (g.n_seq == 0) return MSST_OK;
    const Drop drop = make_drop(d->drop_p, d->seed, d->site, d->seed_dev);
    const dim3 grid((unsigned)(g.groups * g.tiles), g.H);
    if (attention_bwd_tc_supported(g)) {   // default: tcgen05 / TMEM backward (attention_tc.cu, 1.5x the mma.sync kernel); MSST_ATTN_BWD_TC=0 selects mma.sync
        static int use_tc = -1;
        if (use_tc < 0) { const char* e = getenv("MSST_ATTN_BWD_TC"); use_tc = e ? atoi(e) : 1; }
        if (use_tc) return attention_bwd_tc(g, qkv, lse, d_out, d_qkv, drop, st);
    }
    if (g.tiles == 1) {
        static PerDeviceOnce hattr;
        if (hattr.first()) {
            MSST_CUDA(cudaFuncSetAttribute(attn_bwd_bf16_heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdHeadsSmem));
        }
        attn_bwd_bf16_heads_kernel<<<(unsigned)g.groups, BT, kBwdHeadsSmem, st>>>(g, qkv, lse, d_out, d_qkv, drop);
        MSST_LAUNCH_CHECK();
    } else {   // long sequences: dQ pass over key tiles, dK/dV pass over query tiles (no atomics)
        // tcgen05 / TMEM kernels (attention_tc_long_bwd.cu) from N = MSST_ATTN_BWD_TC_LONG (default 256; 0 = never); else mma.sync, cp.async double-buffered
        static int tc_from = -1;
        if (tc_from < 0) { const char* e = getenv("MSST_ATTN_BWD_TC_LONG"); tc_from = e ? atoi(e) : 256; }
        if (tc_from > 0 && g.N >= tc_from && attention_bwd_tc_long_supported(g)) return attention_bwd_tc_long(g, qkv, out, lse, d_out, d_qkv, drop, st);
        static PerDeviceOnce lattr;
        if (lattr.first()) {
            MSST_CUDA(cudaFuncSetAttribute(attn_bwd_bf16_long_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdLongSmem));
            MSST_CUDA(cudaFuncSetAttribute(attn_bwd_bf16_long_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdLongSmem));
        }
        attn_bwd_bf16_long_kernel<1><<<grid, BT, kBwdLongSmem, st>>>(g, qkv, out, lse, d_out, d_qkv, drop);
        MSST_LAUNCH_CHECK();
        attn_bwd_bf16_long_kernel<2><<<grid, BT, kBwdLongSmem, st>>>(g, qkv, out, lse, d_out, d_qkv, drop);
        MSST_LAUNCH_CHECK();
    }
    return MSST_OK;
}

}  // namespace msst
