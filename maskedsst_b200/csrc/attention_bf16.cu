// Kernel (3), bf16 mode: fused flash-style attention forward / backward on tensor cores.
// softmax(q k^T * dh^-0.5) v per (sequence, head); dh = 64; q/k/v/o bf16, statistics + accumulation fp32.
// Reference: Attention.forward, src/vit_spatial_spectral.py:67-78 (and its autograd).
//
// Same tiling / packing geometry as attention_f32.cu (64 query slots x 64 key slots, short sequences packed with a
// block-diagonal mask, strided row addressing for the spectral stack).  One CTA = 4 warps, each warp owns 16 query
// rows (FlashAttention-2 register layout): S and P never leave registers, O / dQ accumulate in registers; the tiles
// here are 64x64x64 -- too small for a tcgen05 pipeline to pay off (one UMMA M=64 tile per CTA), so the
// contractions use mma.sync.m16n8k16 with ldmatrix-fed fragments.  Scores never reach HBM.
// Algorithmic HBM bytes per (slot, head): fwd 3*128 B in + 128 B out + 4 B lse; bwd 5*128 B in + 3*128 B out.
#include "common.cuh"
#include "kernels.h"
#include "attn_geom.cuh"

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

// [64 slots x 64] bf16 tile of one of q/k/v/o (column offset col0) -> smem, zero rows for invalid slots
template <class L>
__device__ __forceinline__ void load_tile_bf16(const AttnGeom& g, const bf16* __restrict__ base, int64_t ld, int col0, int64_t group,
                                               int tile, bf16* dst) {
    for (int i = threadIdx.x; i < TS * 8; i += BT) {
        const int r = i >> 3, c = (i & 7) * 8;
        int64_t seq; int pos;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (slot_to(g, group, tile, r, seq, pos)) v = *reinterpret_cast<const uint4*>(base + row_of(g, seq, pos) * ld + col0 + c);
        *reinterpret_cast<uint4*>(dst + L::at(r, c)) = v;
    }
}
// smem tile rows [row0, row0+16) (bf16) -> global rows (one warp, coalesced 16-byte stores)
template <class L>
__device__ __forceinline__ void store_rows16(const AttnGeom& g, const bf16* src, int row0, bf16* __restrict__ base, int64_t ld, int col0,
                                             const int* seq_s, const int* pos_s, int64_t group, int lane) {
    for (int i = lane; i < 16 * 8; i += 32) {
        const int r = row0 + (i >> 3), c = (i & 7) * 8;
        if (seq_s[r] < 0) continue;
        const int64_t row = row_of(g, group * g.G + seq_s[r], pos_s[r]);
        *reinterpret_cast<uint4*>(base + row * ld + col0 + c) = *reinterpret_cast<const uint4*>(src + L::at(r, c));
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

struct AttnSmemIdx { int qseq[TS], qpos[TS], kseq[TS], kpos[TS]; };

__device__ __forceinline__ void fill_idx(const AttnGeom& g, int64_t group, int tile, int* seq_s, int* pos_s, int invalid) {
    if (threadIdx.x < TS) {
        int64_t seq; int pos;
        const bool ok = slot_to(g, group, tile, threadIdx.x, seq, pos);
        seq_s[threadIdx.x] = ok ? (int)(seq - group * g.G) : invalid;
        pos_s[threadIdx.x] = pos;
    }
}

__global__ void __launch_bounds__(BT) attn_fwd_bf16_kernel(AttnGeom g, const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                           float* __restrict__ lse, Drop drop) {
    using L = Pad;
    __shared__ __align__(16) bf16 Qs[L::TILE];
    __shared__ __align__(16) bf16 Ks[L::TILE];
    __shared__ __align__(16) bf16 Vs[L::TILE];
    __shared__ AttnSmemIdx ix;
    const int I = g.H * 64, h = blockIdx.y;
    const int64_t ld = 3 * (int64_t)I;
    const int64_t group = blockIdx.x / g.tiles;
    const int qt = blockIdx.x % g.tiles;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
    const int r0 = warp * 16 + gq, r1 = r0 + 8;
    const float sl2 = g.scale * 1.4426950408889634f;   // scores are kept in log2 units: exp2f(s*scale*log2e - m)

    load_tile_bf16<L>(g, qkv, ld, h * 64, group, qt, Qs);
    fill_idx(g, group, qt, ix.qseq, ix.qpos, -1);
    float o[8][4] = {};
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    for (int kt = 0; kt < g.tiles; ++kt) {
        __syncthreads();
        load_tile_bf16<L>(g, qkv, ld, I + h * 64, group, kt, Ks);
        load_tile_bf16<L>(g, qkv, ld, 2 * I + h * 64, group, kt, Vs);
        fill_idx(g, group, kt, ix.kseq, ix.kpos, -2);
        __syncthreads();
        float s[8][4] = {};
        gemm_a_bnk<L>(s, Qs, warp * 16, Ks, lane);
        const int qs0 = ix.qseq[r0], qs1 = ix.qseq[r1];
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int c = nt * 8 + 2 * tq;
            const int ks0 = ix.kseq[c], ks1 = ix.kseq[c + 1];
            s[nt][0] = qs0 == ks0 ? s[nt][0] * sl2 : -INFINITY; s[nt][1] = qs0 == ks1 ? s[nt][1] * sl2 : -INFINITY;
            s[nt][2] = qs1 == ks0 ? s[nt][2] * sl2 : -INFINITY; s[nt][3] = qs1 == ks1 ? s[nt][3] * sl2 : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1])); mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float c0 = mn0 == -INFINITY ? 1.f : exp2f(m0 - mn0), c1 = mn1 == -INFINITY ? 1.f : exp2f(m1 - mn1);
        const float sub0 = mn0 == -INFINITY ? 0.f : mn0, sub1 = mn1 == -INFINITY ? 0.f : mn1;
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = exp2f(s[nt][0] - sub0); s[nt][1] = exp2f(s[nt][1] - sub0);
            s[nt][2] = exp2f(s[nt][2] - sub1); s[nt][3] = exp2f(s[nt][3] - sub1);
            rs0 += s[nt][0] + s[nt][1]; rs1 += s[nt][2] + s[nt][3];
        }
        rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
        rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
        l0 = l0 * c0 + rs0; l1 = l1 * c1 + rs1; m0 = mn0; m1 = mn1;
        if (drop.on()) {
            const uint64_t base = tile_pair_base(g, group, h, qt, kt);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                float f0, f1;
                pair_factors(drop, base + (uint64_t)(r0 * 32 + nt * 4 + tq), f0, f1); s[nt][0] *= f0; s[nt][1] *= f1;
                pair_factors(drop, base + (uint64_t)(r1 * 32 + nt * 4 + tq), f0, f1); s[nt][2] *= f0; s[nt][3] *= f1;
            }
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { o[nt][0] *= c0; o[nt][1] *= c0; o[nt][2] *= c1; o[nt][3] *= c1; }
        gemm_p_bkn<L>(o, s, Vs, lane);
    }
    // normalise, stage this warp's 16 rows through its own Q rows (no other warp reads them), coalesced store
    const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        *reinterpret_cast<uint32_t*>(Qs + L::at(r0, nt * 8 + 2 * tq)) = pack2(o[nt][0] * i0, o[nt][1] * i0);
        *reinterpret_cast<uint32_t*>(Qs + L::at(r1, nt * 8 + 2 * tq)) = pack2(o[nt][2] * i1, o[nt][3] * i1);
    }
    __syncwarp();
    store_rows16<L>(g, Qs, warp * 16, out, I, h * 64, ix.qseq, ix.qpos, group, lane);
    if (tq == 0) {
        if (ix.qseq[r0] >= 0) lse[row_of(g, group * g.G + ix.qseq[r0], ix.qpos[r0]) * g.H + h] = (m0 + log2f(l0)) * 0.6931471805599453f;
        if (ix.qseq[r1] >= 0) lse[row_of(g, group * g.G + ix.qseq[r1], ix.qpos[r1]) * g.H + h] = (m1 + log2f(l1)) * 0.6931471805599453f;
    }
}

// MODE 0: one tile (N <= 64): dQ, dK, dV in one CTA.  MODE 1: dQ of q tile blockIdx (loops key tiles).
// MODE 2: dK, dV of key tile blockIdx (loops query tiles).
template <int MODE>
__global__ void __launch_bounds__(BT) attn_bwd_bf16_kernel(AttnGeom g, const bf16* __restrict__ qkv, const bf16* __restrict__ out,
                                                           const float* __restrict__ lse, const bf16* __restrict__ d_out,
                                                           bf16* __restrict__ d_qkv, Drop drop) {
    using L = Pad;
    extern __shared__ __align__(16) uint8_t smem_bwd[];
    bf16* Qs = reinterpret_cast<bf16*>(smem_bwd);
    bf16* Ks = Qs + L::TILE;
    bf16* Vs = Ks + L::TILE;
    bf16* dOs = Vs + L::TILE;
    bf16* Ps = dOs + L::TILE;     // P * dropout factor   [q][key]
    bf16* dSs = Ps + L::TILE;     // dS                   [q][key]
    float* Drow = reinterpret_cast<float*>(dSs + L::TILE);
    float* lse_s = Drow + TS;
    AttnSmemIdx& ix = *reinterpret_cast<AttnSmemIdx*>(lse_s + TS);

    const int I = g.H * 64, h = blockIdx.y;
    const int64_t ld = 3 * (int64_t)I;
    const int64_t group = blockIdx.x / g.tiles;
    const int own = blockIdx.x % g.tiles;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
    const int r0 = warp * 16 + gq, r1 = r0 + 8;
    const float sl2 = g.scale * 1.4426950408889634f;
    const int n_inner = MODE == 0 ? 1 : g.tiles;

    float dq[8][4] = {}, dk[8][4] = {}, dv[8][4] = {};
    for (int it = 0; it < n_inner; ++it) {
        const int qt = MODE == 2 ? it : own, kt = MODE == 1 ? it : own;
        __syncthreads();
        if (MODE != 2 || it == 0) {
            load_tile_bf16<L>(g, qkv, ld, I + h * 64, group, kt, Ks);
            load_tile_bf16<L>(g, qkv, ld, 2 * I + h * 64, group, kt, Vs);
            fill_idx(g, group, kt, ix.kseq, ix.kpos, -2);
        }
        if (MODE != 1 || it == 0) {
            load_tile_bf16<L>(g, qkv, ld, h * 64, group, qt, Qs);
            load_tile_bf16<L>(g, d_out, I, h * 64, group, qt, dOs);
            fill_idx(g, group, qt, ix.qseq, ix.qpos, -1);
        }
        __syncthreads();
        if (MODE != 1 || it == 0) {
            // D_i = dO_i . O_i, lse_i  (warp: 16 rows, two rows per pass, 16 lanes x 4 elements per row)
            for (int k = 0; k < 8; ++k) {
                const int r = warp * 16 + k * 2 + (lane >> 4), c = (lane & 15) * 4;
                float dsum = 0.f, l = 0.f;
                if (ix.qseq[r] >= 0) {
                    const int64_t row = row_of(g, group * g.G + ix.qseq[r], ix.qpos[r]);
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
            __syncthreads();
        }
        float s[8][4] = {}, dp[8][4] = {};
        gemm_a_bnk<L>(s, Qs, warp * 16, Ks, lane);
        gemm_a_bnk<L>(dp, dOs, warp * 16, Vs, lane);
        const int qs0 = ix.qseq[r0], qs1 = ix.qseq[r1];
        const float L0 = lse_s[r0], L1 = lse_s[r1], D0 = Drow[r0], D1 = Drow[r1];
        const uint64_t base = tile_pair_base(g, group, h, qt, kt);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int c = nt * 8 + 2 * tq;
            const int ks0 = ix.kseq[c], ks1 = ix.kseq[c + 1];
            float p0 = qs0 == ks0 ? exp2f(s[nt][0] * sl2 - L0) : 0.f, p1 = qs0 == ks1 ? exp2f(s[nt][1] * sl2 - L0) : 0.f;
            float p2 = qs1 == ks0 ? exp2f(s[nt][2] * sl2 - L1) : 0.f, p3 = qs1 == ks1 ? exp2f(s[nt][3] * sl2 - L1) : 0.f;
            float f0 = 1.f, f1 = 1.f, f2 = 1.f, f3 = 1.f;
            if (drop.on()) {
                pair_factors(drop, base + (uint64_t)(r0 * 32 + nt * 4 + tq), f0, f1);
                pair_factors(drop, base + (uint64_t)(r1 * 32 + nt * 4 + tq), f2, f3);
            }
            // dS = P * (f * dP - D) * scale ; Pf = P * f
            s[nt][0] = p0 * (f0 * dp[nt][0] - D0) * g.scale; s[nt][1] = p1 * (f1 * dp[nt][1] - D0) * g.scale;
            s[nt][2] = p2 * (f2 * dp[nt][2] - D1) * g.scale; s[nt][3] = p3 * (f3 * dp[nt][3] - D1) * g.scale;
            if (MODE != 1) {
                *reinterpret_cast<uint32_t*>(Ps + L::at(r0, c)) = pack2(p0 * f0, p1 * f1);
                *reinterpret_cast<uint32_t*>(Ps + L::at(r1, c)) = pack2(p2 * f2, p3 * f3);
                *reinterpret_cast<uint32_t*>(dSs + L::at(r0, c)) = pack2(s[nt][0], s[nt][1]);
                *reinterpret_cast<uint32_t*>(dSs + L::at(r1, c)) = pack2(s[nt][2], s[nt][3]);
            }
        }
        if (MODE != 2) gemm_p_bkn<L>(dq, s, Ks, lane);          // dQ[16 rows] += dS . K
        if (MODE != 1) {
            __syncthreads();
            gemm_at_bkn<L>(dv, Ps, warp * 16, dOs, lane);       // dV[16 keys] += (P f)^T . dO
            gemm_at_bkn<L>(dk, dSs, warp * 16, Qs, lane);       // dK[16 keys] += dS^T . Q
        }
    }
    __syncthreads();
    // stage results through smem (Q/K/V tiles are dead now) and store coalesced
    if (MODE != 2) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            *reinterpret_cast<uint32_t*>(Qs + L::at(r0, nt * 8 + 2 * tq)) = pack2(dq[nt][0], dq[nt][1]);
            *reinterpret_cast<uint32_t*>(Qs + L::at(r1, nt * 8 + 2 * tq)) = pack2(dq[nt][2], dq[nt][3]);
        }
    }
    if (MODE != 1) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            *reinterpret_cast<uint32_t*>(Ks + L::at(r0, nt * 8 + 2 * tq)) = pack2(dk[nt][0], dk[nt][1]);
            *reinterpret_cast<uint32_t*>(Ks + L::at(r1, nt * 8 + 2 * tq)) = pack2(dk[nt][2], dk[nt][3]);
            *reinterpret_cast<uint32_t*>(Vs + L::at(r0, nt * 8 + 2 * tq)) = pack2(dv[nt][0], dv[nt][1]);
            *reinterpret_cast<uint32_t*>(Vs + L::at(r1, nt * 8 + 2 * tq)) = pack2(dv[nt][2], dv[nt][3]);
        }
    }
    __syncwarp();
    if (MODE != 2) store_rows16<L>(g, Qs, warp * 16, d_qkv, ld, h * 64, ix.qseq, ix.qpos, group, lane);
    if (MODE != 1) {
        store_rows16<L>(g, Ks, warp * 16, d_qkv, ld, I + h * 64, ix.kseq, ix.kpos, group, lane);
        store_rows16<L>(g, Vs, warp * 16, d_qkv, ld, 2 * I + h * 64, ix.kseq, ix.kpos, group, lane);
    }
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
        hs.seq[threadIdx.x] = ok ? (int)(seq - group * g.G) : -1;
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

constexpr size_t kBwdSmem = sizeof(bf16) * 6 * Pad::TILE + sizeof(float) * 2 * TS + sizeof(AttnSmemIdx);

int attention_fwd_bf16(const msst_attn_dims* d, const bf16* qkv, bf16* out, float* lse, cudaStream_t st) {
    AttnGeom g;
    if (int rc = make_attn_geom(d, g, false)) return rc;
    if (g.n_seq == 0) return MSST_OK;
    const Drop drop = make_drop(d->drop_p, d->seed, d->site, d->seed_dev);
    if (g.tiles == 1) {   // N <= 64: head-looping, cp.async double-buffered kernel
        static bool attr_set = false;
        if (!attr_set) {
            MSST_CUDA(cudaFuncSetAttribute(attn_fwd_bf16_heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdHeadsSmem));
            attr_set = true;
        }
        attn_fwd_bf16_heads_kernel<<<(unsigned)g.groups, BT, kFwdHeadsSmem, st>>>(g, qkv, out, lse, drop);
        MSST_LAUNCH_CHECK();
        return MSST_OK;
    }
    attn_fwd_bf16_kernel<<<dim3((unsigned)(g.groups * g.tiles), g.H), BT, 0, st>>>(g, qkv, out, lse, drop);
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}

int attention_bwd_bf16(const msst_attn_dims* d, const bf16* qkv, const bf16* out, const float* lse, const bf16* d_out, bf16* d_qkv,
                       cudaStream_t st) {
    AttnGeom g;
    if (int rc = make_attn_geom(d, g, false)) return rc;
    if (g.n_seq == 0) return MSST_OK;
    const Drop drop = make_drop(d->drop_p, d->seed, d->site, d->seed_dev);
    const dim3 grid((unsigned)(g.groups * g.tiles), g.H);
    static bool attr_set = false;
    if (!attr_set) {
        MSST_CUDA(cudaFuncSetAttribute(attn_bwd_bf16_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));
        MSST_CUDA(cudaFuncSetAttribute(attn_bwd_bf16_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));
        MSST_CUDA(cudaFuncSetAttribute(attn_bwd_bf16_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));
        attr_set = true;
    }
    if (g.tiles == 1) {
        static bool hattr = false;
        if (!hattr) {
            MSST_CUDA(cudaFuncSetAttribute(attn_bwd_bf16_heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdHeadsSmem));
            hattr = true;
        }
        attn_bwd_bf16_heads_kernel<<<(unsigned)g.groups, BT, kBwdHeadsSmem, st>>>(g, qkv, lse, d_out, d_qkv, drop);
        MSST_LAUNCH_CHECK();
    } else {   // long sequences: dQ pass over key tiles, dK/dV pass over query tiles (no atomics)
        attn_bwd_bf16_kernel<1><<<grid, BT, kBwdSmem, st>>>(g, qkv, out, lse, d_out, d_qkv, drop);
        MSST_LAUNCH_CHECK();
        attn_bwd_bf16_kernel<2><<<grid, BT, kBwdSmem, st>>>(g, qkv, out, lse, d_out, d_qkv, drop);
        MSST_LAUNCH_CHECK();
    }
    return MSST_OK;
}

}  // namespace msst
