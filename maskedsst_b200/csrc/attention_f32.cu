// fp32 flash-style attention forward/backward for the MSST_PREC_FP32 parity mode.
// softmax(q k^T * dh^-0.5) v per (sequence, head); scores live only in shared memory.
// Reference: Attention.forward, src/vit_spatial_spectral.py:67-78 (and its autograd).
//
// Tiling: 64 query slots x 64 key slots per CTA step, one CTA per (slot group, head).
//   N <= 64 : G = 64/N whole sequences are packed into one 64-slot tile with a block-diagonal mask
//             (spatial N = 64 -> G = 1; spectral N = 20 -> G = 3; N = 5 -> G = 12)
//   N  > 64 : one sequence per CTA column, online softmax over ceil(N/64) key tiles.
// Row addressing row(seq,pos) = (seq/inner)*N*inner + seq%inner + pos*inner folds the reference's
// '(b c)(h w) d -> (b h w) c d' transpose copies into index arithmetic.
#include "common.cuh"
#include "kernels.h"
#include "attn_geom.cuh"

namespace msst {

// loads a [64 x DH] tile (rows = slots) of one of q/k/v (column offset col0) into smem, zero for invalid slots
template <int DH>
__device__ __forceinline__ void load_rows(const AttnGeom& g, const float* __restrict__ base, int64_t ld, int col0, int64_t group,
                                          int tile, float (*dst)[DH + 1]) {
    constexpr int VPR = DH / 4;   // float4 per row
    for (int i = threadIdx.x; i < TS * VPR; i += AT) {
        const int r = i / VPR, c4 = (i % VPR) * 4;
        int64_t seq; int pos;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (slot_to(g, group, tile, r, seq, pos)) v = *reinterpret_cast<const float4*>(base + row_of(g, seq, pos) * ld + col0 + c4);
        dst[r][c4] = v.x; dst[r][c4 + 1] = v.y; dst[r][c4 + 2] = v.z; dst[r][c4 + 3] = v.w;
    }
}

template <int DH>
__global__ void __launch_bounds__(AT) attn_fwd_kernel(AttnGeom g, const float* __restrict__ qkv, float* __restrict__ out,
                                                      float* __restrict__ lse, Drop drop) {
    constexpr int CW = DH / 16;
    extern __shared__ float smem[];
    float (*Qs)[DH + 1] = reinterpret_cast<float (*)[DH + 1]>(smem);
    float (*Ks)[DH + 1] = Qs + TS;
    float (*Vs)[DH + 1] = Ks + TS;
    float (*Ss)[TS + 1] = reinterpret_cast<float (*)[TS + 1]>(&Vs[TS][0]);
    float* m_run = &Ss[TS][0];
    float* l_run = m_run + TS;
    float* corr = l_run + TS;
    int* qseq = reinterpret_cast<int*>(corr + TS);   // per q slot: local sequence id (or -1) and position
    int* qpos = qseq + TS;
    int* kseq = qpos + TS;
    int* kpos = kseq + TS;

    const int I = g.H * g.dh, h = blockIdx.y;
    const int64_t ld = 3 * (int64_t)I;
    const int64_t group = blockIdx.x / g.tiles;
    const int qtile = blockIdx.x % g.tiles;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    load_rows<DH>(g, qkv, ld, h * DH, group, qtile, Qs);
    if (threadIdx.x < TS) {
        int64_t seq; int pos;
        const bool ok = slot_to(g, group, qtile, threadIdx.x, seq, pos);
        qseq[threadIdx.x] = ok ? (int)(seq - group_seq0(g, group)) : -1;
        qpos[threadIdx.x] = pos;
        m_run[threadIdx.x] = -INFINITY; l_run[threadIdx.x] = 0.f;
    }
    float acc[4][CW] = {};
    for (int kt = 0; kt < g.tiles; ++kt) {
        __syncthreads();
        load_rows<DH>(g, qkv, ld, I + h * DH, group, kt, Ks);
        load_rows<DH>(g, qkv, ld, 2 * I + h * DH, group, kt, Vs);
        if (threadIdx.x < TS) {
            int64_t seq; int pos;
            const bool ok = slot_to(g, group, kt, threadIdx.x, seq, pos);
            kseq[threadIdx.x] = ok ? (int)(seq - group_seq0(g, group)) : -2;
            kpos[threadIdx.x] = pos;
        }
        __syncthreads();
        float s[4][4] = {};
#pragma unroll 8
        for (int d = 0; d < DH; ++d) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = Qs[ty * 4 + i][d]; b[i] = Ks[tx * 4 + i][d]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) s[i][j] = fmaf(a[i], b[j], s[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool ok = qseq[ty * 4 + i] == kseq[tx * 4 + j];
                Ss[ty * 4 + i][tx * 4 + j] = ok ? s[i][j] * g.scale : -INFINITY;
            }
        __syncthreads();
        // online softmax: warp handles 8 rows, lane handles columns lane, lane+32
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int r = warp * 8 + k;
            const float s0 = Ss[r][lane], s1 = Ss[r][lane + 32];
            const float mx = warp_max(fmaxf(s0, s1));
            const float m_old = m_run[r];
            const float m_new = fmaxf(m_old, mx);
            float p0 = 0.f, p1 = 0.f, c = 1.f;
            if (m_new != -INFINITY) {
                p0 = expf(s0 - m_new); p1 = expf(s1 - m_new);
                c = expf(m_old - m_new);
            }
            const float rs = warp_sum(p0 + p1);
            if (drop.on() && qseq[r] >= 0) {
                const int64_t seq = group_seq0(g, group) + qseq[r];
                const uint64_t e0 = (((uint64_t)seq * g.H + h) * g.N + qpos[r]) * g.N;
                if (kseq[lane] == qseq[r]) p0 *= drop_factor(drop, e0 + kpos[lane]);
                if (kseq[lane + 32] == qseq[r]) p1 *= drop_factor(drop, e0 + kpos[lane + 32]);
            }
            Ss[r][lane] = p0; Ss[r][lane + 32] = p1;
            if (lane == 0) { m_run[r] = m_new; l_run[r] = l_run[r] * c + rs; corr[r] = c; }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float c = corr[ty * 4 + i];
#pragma unroll
            for (int cc = 0; cc < CW; ++cc) acc[i][cc] *= c;
        }
#pragma unroll 8
        for (int j = 0; j < TS; ++j) {
            float p[4], v[CW];
#pragma unroll
            for (int i = 0; i < 4; ++i) p[i] = Ss[ty * 4 + i][j];
#pragma unroll
            for (int cc = 0; cc < CW; ++cc) v[cc] = Vs[j][tx * CW + cc];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int cc = 0; cc < CW; ++cc) acc[i][cc] = fmaf(p[i], v[cc], acc[i][cc]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = ty * 4 + i;
        if (qseq[r] < 0) continue;
        const int64_t row = row_of(g, group_seq0(g, group) + qseq[r], qpos[r]);
        const float inv = 1.f / l_run[r];
#pragma unroll
        for (int cc = 0; cc < CW; ++cc) out[row * I + h * DH + tx * CW + cc] = acc[i][cc] * inv;
        if (tx == 0) lse[row * g.H + h] = m_run[r] + logf(l_run[r]);
    }
}

template <int DH>
__global__ void __launch_bounds__(AT) attn_bwd_kernel(AttnGeom g, const float* __restrict__ qkv, const float* __restrict__ out,
                                                      const float* __restrict__ lse, const float* __restrict__ d_out,
                                                      float* __restrict__ d_qkv, Drop drop) {
    constexpr int CW = DH / 16;
    extern __shared__ float smem[];
    float (*Qs)[DH + 1] = reinterpret_cast<float (*)[DH + 1]>(smem);
    float (*Ks)[DH + 1] = Qs + TS;
    float (*Vs)[DH + 1] = Ks + TS;
    float (*dOs)[DH + 1] = Vs + TS;
    float (*Ss)[TS + 1] = reinterpret_cast<float (*)[TS + 1]>(&dOs[TS][0]);
    float* Drow = &Ss[TS][0];
    float* lse_s = Drow + TS;
    int* qseq = reinterpret_cast<int*>(lse_s + TS);
    int* qpos = qseq + TS;
    int* kseq = qpos + TS;
    int* kpos = kseq + TS;

    const int I = g.H * g.dh, h = blockIdx.y;
    const int64_t ld = 3 * (int64_t)I;
    const int64_t group = blockIdx.x / g.tiles;
    const int ktile = blockIdx.x % g.tiles;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    load_rows<DH>(g, qkv, ld, I + h * DH, group, ktile, Ks);
    load_rows<DH>(g, qkv, ld, 2 * I + h * DH, group, ktile, Vs);
    if (threadIdx.x < TS) {
        int64_t seq; int pos;
        const bool ok = slot_to(g, group, ktile, threadIdx.x, seq, pos);
        kseq[threadIdx.x] = ok ? (int)(seq - group_seq0(g, group)) : -2;
        kpos[threadIdx.x] = pos;
    }
    float dK[4][CW] = {}, dV[4][CW] = {};
    for (int qt = 0; qt < g.tiles; ++qt) {
        __syncthreads();
        load_rows<DH>(g, qkv, ld, h * DH, group, qt, Qs);
        load_rows<DH>(g, d_out, I, h * DH, group, qt, dOs);
        if (threadIdx.x < TS) {
            int64_t seq; int pos;
            const bool ok = slot_to(g, group, qt, threadIdx.x, seq, pos);
            qseq[threadIdx.x] = ok ? (int)(seq - group_seq0(g, group)) : -1;
            qpos[threadIdx.x] = pos;
        }
        __syncthreads();
        // D_i = dO_i . O_i and lse_i  (warp: 8 rows, lanes over DH)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int r = warp * 8 + k;
            float dsum = 0.f, l = 0.f;
            if (qseq[r] >= 0) {
                const int64_t row = row_of(g, group_seq0(g, group) + qseq[r], qpos[r]);
                for (int d = lane; d < DH; d += 32) dsum += dOs[r][d] * out[row * I + h * DH + d];
                l = lse[row * g.H + h];
            }
            dsum = warp_sum(dsum);
            if (lane == 0) { Drow[r] = dsum; lse_s[r] = l; }
        }
        __syncthreads();
        float s[4][4] = {}, dp[4][4] = {};
#pragma unroll 4
        for (int d = 0; d < DH; ++d) {
            float a[4], b[4], e[4], v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = Qs[ty * 4 + i][d]; b[i] = Ks[tx * 4 + i][d]; e[i] = dOs[ty * 4 + i][d]; v[i] = Vs[tx * 4 + i][d]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) { s[i][j] = fmaf(a[i], b[j], s[i][j]); dp[i][j] = fmaf(e[i], v[j], dp[i][j]); }
        }
        // P (softmax probs), Pf = P * dropout factor -> smem for dV; dS kept in registers
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = ty * 4 + i;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int cidx = tx * 4 + j;
                const bool ok = qseq[r] == kseq[cidx];
                float p = ok ? expf(s[i][j] * g.scale - lse_s[r]) : 0.f;
                float f = 1.f;
                if (drop.on() && ok) {
                    const int64_t seq = group_seq0(g, group) + qseq[r];
                    f = drop_factor(drop, (((uint64_t)seq * g.H + h) * g.N + qpos[r]) * g.N + kpos[cidx]);
                }
                Ss[r][cidx] = p * f;
                s[i][j] = p * (f * dp[i][j] - Drow[r]) * g.scale;   // dS
            }
        }
        __syncthreads();
        // dV[j][c] += sum_i Pf[i][j] dO[i][c]     (thread owns key rows ty*4.., cols tx*CW..)
#pragma unroll 8
        for (int i = 0; i < TS; ++i) {
            float p[4], e[CW];
#pragma unroll
            for (int j = 0; j < 4; ++j) p[j] = Ss[i][ty * 4 + j];
#pragma unroll
            for (int cc = 0; cc < CW; ++cc) e[cc] = dOs[i][tx * CW + cc];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int cc = 0; cc < CW; ++cc) dV[j][cc] = fmaf(p[j], e[cc], dV[j][cc]);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) Ss[ty * 4 + i][tx * 4 + j] = s[i][j];
        __syncthreads();
        // dK[j][c] += sum_i dS[i][j] Q[i][c];   dQ[i][c] = sum_j dS[i][j] K[j][c]
        float dQ[4][CW] = {};
#pragma unroll 8
        for (int i = 0; i < TS; ++i) {
            float p[4], qv[CW];
#pragma unroll
            for (int j = 0; j < 4; ++j) p[j] = Ss[i][ty * 4 + j];
#pragma unroll
            for (int cc = 0; cc < CW; ++cc) qv[cc] = Qs[i][tx * CW + cc];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int cc = 0; cc < CW; ++cc) dK[j][cc] = fmaf(p[j], qv[cc], dK[j][cc]);
        }
#pragma unroll 8
        for (int j = 0; j < TS; ++j) {
            float p[4], kv[CW];
#pragma unroll
            for (int i = 0; i < 4; ++i) p[i] = Ss[ty * 4 + i][j];
#pragma unroll
            for (int cc = 0; cc < CW; ++cc) kv[cc] = Ks[j][tx * CW + cc];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int cc = 0; cc < CW; ++cc) dQ[i][cc] = fmaf(p[i], kv[cc], dQ[i][cc]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = ty * 4 + i;
            if (qseq[r] < 0) continue;
            const int64_t row = row_of(g, group_seq0(g, group) + qseq[r], qpos[r]);
#pragma unroll
            for (int cc = 0; cc < CW; ++cc) {
                float* dst = d_qkv + row * ld + h * DH + tx * CW + cc;
                if (g.tiles == 1) *dst = dQ[i][cc];
                else atomicAdd(dst, dQ[i][cc]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = ty * 4 + j;
        if (kseq[r] < 0) continue;
        const int64_t row = row_of(g, group_seq0(g, group) + kseq[r], kpos[r]);
#pragma unroll
        for (int cc = 0; cc < CW; ++cc) {
            d_qkv[row * ld + I + h * DH + tx * CW + cc] = dK[j][cc];
            d_qkv[row * ld + 2 * I + h * DH + tx * CW + cc] = dV[j][cc];
        }
    }
}

template <int DH>
static int launch_fwd(const AttnGeom& g, const float* qkv, float* out, float* lse, Drop drop, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)3 * TS * (DH + 1) + (size_t)TS * (TS + 1) + 3 * TS) + sizeof(int) * 4 * TS;
    MSST_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_fwd_kernel<DH><<<dim3((unsigned)(g.groups * g.tiles), g.H), AT, smem, st>>>(g, qkv, out, lse, drop);
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}
template <int DH>
static int launch_bwd(const AttnGeom& g, const float* qkv, const float* out, const float* lse, const float* d_out, float* d_qkv,
                      Drop drop, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)4 * TS * (DH + 1) + (size_t)TS * (TS + 1) + 2 * TS) + sizeof(int) * 4 * TS;
    MSST_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_kernel<DH><<<dim3((unsigned)(g.groups * g.tiles), g.H), AT, smem, st>>>(g, qkv, out, lse, d_out, d_qkv, drop);
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}

int attention_fwd_f32(const msst_attn_dims* d, const float* qkv, float* out, float* lse, cudaStream_t st) {
    AttnGeom g;
    if (int rc = make_attn_geom(d, g, true)) return rc;
    if (g.n_seq == 0) return MSST_OK;
    const Drop drop = make_drop(d->drop_p, d->seed, d->site, d->seed_dev);
    switch (g.dh) {
        case 32: return launch_fwd<32>(g, qkv, out, lse, drop, st);
        case 64: return launch_fwd<64>(g, qkv, out, lse, drop, st);
        default: return launch_fwd<128>(g, qkv, out, lse, drop, st);
    }
}

int attention_bwd_f32(const msst_attn_dims* d, const float* qkv, const float* out, const float* lse, const float* d_out,
                      float* d_qkv, cudaStream_t st) {
    AttnGeom g;
    if (int rc = make_attn_geom(d, g, true)) return rc;
    if (g.n_seq == 0) return MSST_OK;
    const Drop drop = make_drop(d->drop_p, d->seed, d->site, d->seed_dev);
    if (g.tiles > 1)   // dQ is accumulated across key tiles with atomics
        MSST_CUDA(cudaMemsetAsync(d_qkv, 0, sizeof(float) * (size_t)g.n_seq * g.N * 3 * g.H * g.dh, st));
    switch (g.dh) {
        case 32: return launch_bwd<32>(g, qkv, out, lse, d_out, d_qkv, drop, st);
        case 64: return launch_bwd<64>(g, qkv, out, lse, d_out, d_qkv, drop, st);
        default: return launch_bwd<128>(g, qkv, out, lse, d_out, d_qkv, drop, st);
    }
}

}  // namespace msst
