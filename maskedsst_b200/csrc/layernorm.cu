// Kernel (4a): warp-shuffle LayerNorm forward/backward (nn.LayerNorm semantics: biased variance, eps in sqrt).
// Reference: PreNorm, src/vit_spatial_spectral.py:22-29 (16 calls per forward).  HBM-bound:
// fwd reads rows*D*4 B and writes rows*D*(4|2) B; one warp per row, D/32 values per lane in registers.
#include "common.cuh"

namespace msst {

constexpr int kLnThreads = 256;

template <int NJ, bool BF16>
__global__ void __launch_bounds__(kLnThreads)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, void* __restrict__ y,
              float* __restrict__ stats, int64_t rows, int D, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (kLnThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kLnThreads / 32);
    float wj[NJ], bj[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) { const int f = lane + 32 * j; wj[j] = f < D ? w[f] : 0.f; bj[j] = f < D ? b[f] : 0.f; }
    for (int64_t r = warp0; r < rows; r += nwarps) {
        float v[NJ], s = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) { const int f = lane + 32 * j; v[j] = f < D ? x[r * D + f] : 0.f; s += v[j]; }
        const float mean = warp_sum(s) / D;
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) { const int f = lane + 32 * j; const float dl = f < D ? v[j] - mean : 0.f; sq += dl * dl; }
        const float rstd = rsqrtf(warp_sum(sq) / D + eps);
        if (stats && lane == 0) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int f = lane + 32 * j;
            if (f < D) {
                const float o = (v[j] - mean) * rstd * wj[j] + bj[j];
                if (BF16) reinterpret_cast<__nv_bfloat16*>(y)[r * D + f] = __float2bfloat16(o);
                else reinterpret_cast<float*>(y)[r * D + f] = o;
            }
        }
    }
}

// dx = dx_add + rstd * (dy*w - mean(dy*w) - xhat * mean(dy*w*xhat)); dw += sum dy*xhat; db += sum dy
template <int NJ>
__global__ void __launch_bounds__(kLnThreads)
ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ stats,
              const float* __restrict__ dy, const float* __restrict__ dx_add, float* __restrict__ dx,
              float* __restrict__ dw, float* __restrict__ db, int64_t rows, int D) {
    __shared__ float red[2][kLnThreads / 32][NJ * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warp0 = (int64_t)blockIdx.x * (kLnThreads / 32) + warp;
    const int64_t nwarps = (int64_t)gridDim.x * (kLnThreads / 32);
    float wj[NJ], aw[NJ], ab[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) { const int f = lane + 32 * j; wj[j] = f < D ? w[f] : 0.f; aw[j] = ab[j] = 0.f; }
    for (int64_t r = warp0; r < rows; r += nwarps) {
        const float mean = stats[2 * r], rstd = stats[2 * r + 1];
        float xh[NJ], g[NJ], c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int f = lane + 32 * j;
            const float d = f < D ? dy[r * D + f] : 0.f;
            xh[j] = f < D ? (x[r * D + f] - mean) * rstd : 0.f;
            aw[j] += d * xh[j]; ab[j] += d;
            g[j] = d * wj[j];
            c1 += g[j]; c2 += g[j] * xh[j];
        }
        c1 = warp_sum(c1) / D; c2 = warp_sum(c2) / D;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int f = lane + 32 * j;
            if (f < D) {
                float o = rstd * (g[j] - c1 - xh[j] * c2);
                if (dx_add) o += dx_add[r * D + f];
                dx[r * D + f] = o;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) { red[0][warp][lane + 32 * j] = aw[j]; red[1][warp][lane + 32 * j] = ab[j]; }
    __syncthreads();
    for (int f = threadIdx.x; f < D; f += kLnThreads) {
        float sw = 0.f, sb = 0.f;
        for (int k = 0; k < kLnThreads / 32; ++k) { sw += red[0][k][f]; sb += red[1][k][f]; }
        atomicAdd(dw + f, sw);
        atomicAdd(db + f, sb);
    }
}

// ---- generic row width (D > 256, e.g. the spectral-MLP head's LayerNorm(D*C)): lanes stride over the row, re-reading it
// from L1/L2 instead of caching it in registers ----
template <bool BF16>
__global__ void __launch_bounds__(kLnThreads)
ln_fwd_wide_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, void* __restrict__ y,
                   float* __restrict__ stats, int64_t rows, int D, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (kLnThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kLnThreads / 32);
    for (int64_t r = warp0; r < rows; r += nwarps) {
        const float* xr = x + r * D;
        float s = 0.f;
        for (int f = lane; f < D; f += 32) s += xr[f];
        const float mean = warp_sum(s) / D;
        float sq = 0.f;
        for (int f = lane; f < D; f += 32) { const float dl = xr[f] - mean; sq += dl * dl; }
        const float rstd = rsqrtf(warp_sum(sq) / D + eps);
        if (stats && lane == 0) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
        for (int f = lane; f < D; f += 32) {
            const float o = (xr[f] - mean) * rstd * w[f] + b[f];
            if (BF16) reinterpret_cast<__nv_bfloat16*>(y)[r * D + f] = __float2bfloat16(o);
            else reinterpret_cast<float*>(y)[r * D + f] = o;
        }
    }
}

__global__ void __launch_bounds__(kLnThreads)
ln_bwd_wide_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ stats,
                   const float* __restrict__ dy, const float* __restrict__ dx_add, float* __restrict__ dx,
                   float* __restrict__ dw, float* __restrict__ db, int64_t rows, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (kLnThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kLnThreads / 32);
    for (int64_t r = warp0; r < rows; r += nwarps) {
        const float mean = stats[2 * r], rstd = stats[2 * r + 1];
        const float* xr = x + r * D; const float* dr = dy + r * D;
        float c1 = 0.f, c2 = 0.f;
        for (int f = lane; f < D; f += 32) {
            const float xh = (xr[f] - mean) * rstd, g = dr[f] * w[f];
            c1 += g; c2 += g * xh;
            atomicAdd(dw + f, dr[f] * xh);     // wide rows are a rarely used head variant: plain atomics
            atomicAdd(db + f, dr[f]);
        }
        c1 = warp_sum(c1) / D; c2 = warp_sum(c2) / D;
        for (int f = lane; f < D; f += 32) {
            const float xh = (xr[f] - mean) * rstd;
            float o = rstd * (dr[f] * w[f] - c1 - xh * c2);
            if (dx_add) o += dx_add[r * D + f];
            dx[r * D + f] = o;
        }
    }
}

static inline int ln_grid(int64_t rows) {
    int64_t blocks = ceil_div(rows, kLnThreads / 32);
    const int64_t cap = (int64_t)kNumSMs * 8;   // 8 resident CTAs of 256 threads per SM
    return (int)(blocks < cap ? blocks : cap);
}

int layernorm_fwd(const float* x, const float* w, const float* b, void* y, int y_bf16, float* stats, int64_t rows, int D,
                  float eps, cudaStream_t st) {
    MSST_REQUIRE(D >= 1, "layernorm: D=%d out of range", D);
    if (rows == 0) return MSST_OK;
    const int nj = (D + 31) / 32, grid = ln_grid(rows);
    if (D > 256) {
        if (y_bf16) ln_fwd_wide_kernel<true><<<grid, kLnThreads, 0, st>>>(x, w, b, y, stats, rows, D, eps);
        else ln_fwd_wide_kernel<false><<<grid, kLnThreads, 0, st>>>(x, w, b, y, stats, rows, D, eps);
        MSST_LAUNCH_CHECK();
        return MSST_OK;
    }
#define MSST_LN(NJ)                                                                                            \
    if (nj <= NJ) {                                                                                            \
        if (y_bf16) ln_fwd_kernel<NJ, true><<<grid, kLnThreads, 0, st>>>(x, w, b, y, stats, rows, D, eps);      \
        else ln_fwd_kernel<NJ, false><<<grid, kLnThreads, 0, st>>>(x, w, b, y, stats, rows, D, eps);            \
        MSST_LAUNCH_CHECK();                                                                                   \
        return MSST_OK;                                                                                        \
    }
    MSST_LN(1) MSST_LN(2) MSST_LN(3) MSST_LN(4) MSST_LN(8)
#undef MSST_LN
    return MSST_ERR_ARG;
}

int layernorm_bwd(const float* x, const float* w, const float* stats, const float* dy, const float* dx_add, float* dx,
                  float* dw, float* db, int64_t rows, int D, cudaStream_t st) {
    MSST_REQUIRE(D >= 1, "layernorm: D=%d out of range", D);
    if (rows == 0) return MSST_OK;
    const int nj = (D + 31) / 32;
    if (D > 256) {
        ln_bwd_wide_kernel<<<ln_grid(rows), kLnThreads, 0, st>>>(x, w, stats, dy, dx_add, dx, dw, db, rows, D);
        MSST_LAUNCH_CHECK();
        return MSST_OK;
    }
    const int grid = ln_grid(rows);   // up to 8 CTAs/SM: the row loop is a dependent chain, memory-level parallelism comes from warps
#define MSST_LN(NJ)                                                                                  \
    if (nj <= NJ) {                                                                                  \
        ln_bwd_kernel<NJ><<<grid, kLnThreads, 0, st>>>(x, w, stats, dy, dx_add, dx, dw, db, rows, D); \
        MSST_LAUNCH_CHECK();                                                                         \
        return MSST_OK;                                                                              \
    }
    MSST_LN(1) MSST_LN(2) MSST_LN(3) MSST_LN(4) MSST_LN(8)
#undef MSST_LN
    return MSST_ERR_ARG;
}

}  // namespace msst

extern "C" int msst_layernorm_fwd(const float* x, const float* w, const float* b, void* y, int y_bf16, float* stats,
                                  int64_t rows, int D, float eps, msst_stream_t stream) {
    return msst::layernorm_fwd(x, w, b, y, y_bf16, stats, rows, D, eps, (cudaStream_t)stream);
}
extern "C" int msst_layernorm_bwd(const float* x, const float* w, const float* stats, const float* dy, const float* dx_add,
                                  float* dx, float* dw, float* db, int64_t rows, int D, msst_stream_t stream) {
    return msst::layernorm_bwd(x, w, stats, dy, dx_add, dx, dw, db, rows, D, (cudaStream_t)stream);
}
