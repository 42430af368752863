// Kernel (4c): SimMIM decoder + masked-pixel L1 loss, forward and backward.
// Reference: src/vit_simmim_original.py:314-338 (gather of masked tokens / patches, to_pixels, F.l1_loss / num_masked)
// and BlockwiseToPixels :9-40 (one Linear(D->P) per spectral block, chosen by idx // S).
// The gathers are fused: masked rows are read straight from the encoder output and target pixels straight
// from the input cube.  HBM-bound: per masked token D + P floats in, nothing out (the loss is a scalar).
#include "common.cuh"
#include "pixel_source.cuh"

namespace msst {

constexpr int DT = 256;
struct DecGeom { int B, C, G, p0, p1, D, nm, n_wb, P, S, T, HW, Wimg; PixelSource src; };

__device__ __forceinline__ float target_pixel(const DecGeom& g, int b, int t, int p) {
    const int c = t / g.S, s = t % g.S;
    if (g.p1 == 1) return g.src.load(b, c * g.p0 + p, s / g.Wimg, s % g.Wimg);
    const int pp = g.p1 * g.p1, p0i = p / pp, r = p % pp, p1i = r / g.p1, p2i = r % g.p1, h = s / g.G, w = s % g.G;
    return g.src.load(b, c * g.p0 + p0i, h * g.p1 + p1i, w * g.p1 + p2i);
}

// lane p < P fetches target element p and bias p of the item up front (one load instruction each instead of P dependent
// scalar loads inside the reduction loop); the P dot products are unrolled so their shuffles interleave.
template <int NJ, int PMAX>
__global__ void __launch_bounds__(DT) decode_fwd_kernel(DecGeom g, const float* __restrict__ enc, const int64_t* __restrict__ idx,
                                                        const float* __restrict__ img, const float* __restrict__ tgt_tok,
                                                        const float* __restrict__ W, const float* __restrict__ bias,
                                                        float* __restrict__ pred, float* __restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const int64_t total = (int64_t)g.B * g.nm;
    // the gather chain idx -> encoder row / target pixels is pure latency: the NEXT item's loads are issued before this item's math
    const int64_t stride = (int64_t)gridDim.x * (DT / 32);
    int64_t it = (int64_t)blockIdx.x * (DT / 32) + (threadIdx.x >> 5);
    float e_n[NJ], tg_n = 0.f;
    int t_n = 0;
    auto fetch = [&](int64_t i) {
        const int b = (int)(i / g.nm);
        t_n = (int)idx[i];
#pragma unroll
        for (int j = 0; j < NJ; ++j) e_n[j] = enc[((int64_t)b * g.T + t_n) * g.D + lane + 32 * j];
        tg_n = 0.f;
        if (lane < g.P) tg_n = tgt_tok ? tgt_tok[((int64_t)b * g.T + t_n) * g.P + lane] : target_pixel(g, b, t_n, lane);
    };
    if (it < total) fetch(it);
    for (; it < total; it += stride) {
        const int t = t_n;
        const int blk = g.n_wb == 1 ? 0 : t / g.S;
        float e[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) e[j] = e_n[j];
        float tgb = 0.f;   // target_p - bias_p held by lane p
        if (lane < g.P) tgb = tg_n - bias[blk * g.P + lane];
        if (it + stride < total) fetch(it + stride);
        float mine = 0.f;  // lane p keeps dot product p
#pragma unroll
        for (int p = 0; p < PMAX; ++p) {
            if (p < g.P) {
                const float* w = W + ((int64_t)blk * g.P + p) * g.D;
                float a = 0.f;
#pragma unroll
                for (int j = 0; j < NJ; ++j) a = fmaf(e[j], __ldg(w + lane + 32 * j), a);
                a = warp_sum(a);
                if (lane == p) mine = a;
            }
        }
        if (pred && lane < g.P) pred[it * g.P + lane] = mine + bias[blk * g.P + lane];
        const float l1 = warp_sum(lane < g.P ? fabsf(mine - tgb) : 0.f);
        if (lane == 0) partial[it] = l1;
    }
}

// deterministic single-CTA reduction of the per-token partial sums
__global__ void __launch_bounds__(1024) loss_reduce_kernel(const float* __restrict__ partial, int64_t n, float scale, float* __restrict__ loss) {
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) s += (double)partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = red[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) *loss = (float)(s * (double)scale);
    }
}

// Backward.  Each warp walks a CONTIGUOUS run of (sample, masked-token) items; gradients of the decoder weights are
// accumulated in registers for the current spectral block and flushed (shared-memory atomics when the [n_wb][P][D] table
// fits, else global atomics) only when the block changes -- with the reference's mask generators the indices of a row are
// ascending, so a run of ~nm/C items shares one block.  d_enc is a scatter-add (duplicate indices are legal).
template <int NJ, int PMAX>
__global__ void __launch_bounds__(DT) decode_bwd_kernel(DecGeom g, const float* __restrict__ enc, const int64_t* __restrict__ idx,
                                                        const float* __restrict__ img, const float* __restrict__ tgt_tok,
                                                        const float* __restrict__ W, const float* __restrict__ bias,
                                                        const float* __restrict__ d_loss, float coef, float* __restrict__ d_enc,
                                                        float* __restrict__ d_W, float* __restrict__ d_bias,
                                                        float* __restrict__ d_tgt, int use_smem, int items_per_warp) {
    extern __shared__ float acc[];     // [n_wb][P][D] + [n_wb][P] when use_smem
    const int nacc = g.n_wb * g.P * g.D, nb = g.n_wb * g.P;
    if (use_smem) {
        for (int i = threadIdx.x; i < nacc + nb; i += DT) acc[i] = 0.f;
        __syncthreads();
    }
    float* accW = use_smem ? acc : d_W;
    float* accB = use_smem ? acc + nacc : d_bias;
    const int lane = threadIdx.x & 31;
    const float gscale = coef * d_loss[0];
    const int64_t total = (int64_t)g.B * g.nm;
    const int64_t warp_id = (int64_t)blockIdx.x * (DT / 32) + (threadIdx.x >> 5);
    const int64_t it0 = warp_id * items_per_warp;
    const int64_t it1 = it0 + items_per_warp < total ? it0 + items_per_warp : total;
    float aW[PMAX][NJ], aB[PMAX];
#pragma unroll
    for (int p = 0; p < PMAX; ++p) {
        aB[p] = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) aW[p][j] = 0.f;
    }
    int cur_blk = -1;
    auto flush = [&]() {
        if (cur_blk < 0) return;
#pragma unroll
        for (int p = 0; p < PMAX; ++p) {
            if (p < g.P) {
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    if (aW[p][j] != 0.f) atomicAdd(accW + ((int64_t)cur_blk * g.P + p) * g.D + lane + 32 * j, aW[p][j]);
                    aW[p][j] = 0.f;
                }
                if (lane == 0 && aB[p] != 0.f) atomicAdd(accB + cur_blk * g.P + p, aB[p]);
                aB[p] = 0.f;
            }
        }
    };
    float e_n[NJ], tg_n = 0.f;
    int t_n = 0;
    auto fetch = [&](int64_t i) {     // next item's gather chain is issued before this item's math (see decode_fwd_kernel)
        const int b = (int)(i / g.nm);
        t_n = (int)idx[i];
#pragma unroll
        for (int j = 0; j < NJ; ++j) e_n[j] = enc[((int64_t)b * g.T + t_n) * g.D + lane + 32 * j];
        tg_n = 0.f;
        if (lane < g.P) tg_n = tgt_tok ? tgt_tok[((int64_t)b * g.T + t_n) * g.P + lane] : target_pixel(g, b, t_n, lane);
    };
    if (it0 < it1) fetch(it0);
    for (int64_t it = it0; it < it1; ++it) {
        const int b = (int)(it / g.nm);
        const int t = t_n;
        const int blk = g.n_wb == 1 ? 0 : t / g.S;
        if (blk != cur_blk) { flush(); cur_blk = blk; }
        float e[NJ], de[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) { e[j] = e_n[j]; de[j] = 0.f; }
        float tgb = 0.f;   // lane p: target_p - bias_p (fetched before the reduction loop)
        if (lane < g.P) tgb = tg_n - bias[blk * g.P + lane];
        if (it + 1 < it1) fetch(it + 1);
#pragma unroll
        for (int p = 0; p < PMAX; ++p) {
            if (p < g.P) {
                const float* w = W + ((int64_t)blk * g.P + p) * g.D;
                float wv[NJ], a = 0.f;
#pragma unroll
                for (int j = 0; j < NJ; ++j) { wv[j] = __ldg(w + lane + 32 * j); a = fmaf(e[j], wv[j], a); }
                const float diff = warp_sum(a) - __shfl_sync(0xffffffffu, tgb, p);
                const float gp = diff > 0.f ? gscale : (diff < 0.f ? -gscale : 0.f);
#pragma unroll
                for (int j = 0; j < NJ; ++j) { aW[p][j] = fmaf(gp, e[j], aW[p][j]); de[j] = fmaf(gp, wv[j], de[j]); }
                aB[p] += gp;
                if (d_tgt && lane == 0 && gp != 0.f) atomicAdd(d_tgt + ((int64_t)b * g.T + t) * g.P + p, -gp);
            }
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) atomicAdd(d_enc + ((int64_t)b * g.T + t) * g.D + lane + 32 * j, de[j]);
    }
    flush();
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < nacc; i += DT) if (acc[i] != 0.f) atomicAdd(d_W + i, acc[i]);
        for (int i = threadIdx.x; i < nb; i += DT) if (acc[nacc + i] != 0.f) atomicAdd(d_bias + i, acc[nacc + i]);
    }
}

static int make_geom(const msst_decode_dims* d, const float* img, const float* target_tokens, DecGeom& g) {
    MSST_REQUIRE(d && d->B > 0 && d->C > 0 && d->G > 0 && d->p0 > 0 && d->p1 > 0 && d->nm > 0, "simmim_decode: bad dims");
    MSST_REQUIRE(d->D % 32 == 0 && d->D >= 32 && d->D <= 256, "simmim_decode: D=%d must be a multiple of 32 in [32,256]", d->D);
    MSST_REQUIRE(d->n_weight_blocks == 1 || d->n_weight_blocks == d->C, "simmim_decode: n_weight_blocks must be 1 or C");
    g.B = d->B; g.C = d->C; g.G = d->G; g.p0 = d->p0; g.p1 = d->p1; g.D = d->D; g.nm = d->nm; g.n_wb = d->n_weight_blocks;
    g.P = d->p0 * d->p1 * d->p1; g.S = d->G * d->G; g.T = g.C * g.S; g.Wimg = d->G * d->p1; g.HW = g.Wimg * g.Wimg;
    g.src = PixelSource{};
    if (!target_tokens) {   // targets are gathered from the cube (fp32 or raw tiles)
        if (const char* e = make_pixel_source(g.src, img, d->raw, g.C * g.p0, g.Wimg, g.Wimg)) { set_error("simmim_decode: %s", e); return MSST_ERR_ARG; }
    }
    return MSST_OK;
}

}  // namespace msst
using namespace msst;

extern "C" int msst_simmim_decode_l1_fwd(const msst_decode_dims* d, const float* enc, const int64_t* idx, const float* img,
                                         const float* target_tokens, const float* W, const float* bias, float* pred,
                                         float* partial, float* loss, msst_stream_t stream) {
    DecGeom g;
    if (int rc = make_geom(d, img, target_tokens, g)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = (int64_t)g.B * g.nm;
    int64_t grid = ceil_div(total, DT / 32);
    if (grid > 8 * kNumSMs) grid = 8 * kNumSMs;
    MSST_REQUIRE(g.P <= 16, "simmim_decode: pixels per patch %d > 16 unsupported", g.P);
#define MSST_D(NJ) case NJ: decode_fwd_kernel<NJ, 16><<<(int)grid, DT, 0, st>>>(g, enc, idx, img, target_tokens, W, bias, pred, partial); break;
    switch (g.D / 32) { MSST_D(1) MSST_D(2) MSST_D(3) MSST_D(4) default: set_error("simmim_decode: D unsupported"); return MSST_ERR_ARG; }
#undef MSST_D
    MSST_LAUNCH_CHECK();
    // loss = mean over B*nm*P elements, then / num_masked (reference double normalisation, :338)
    const float scale = (float)(1.0 / ((double)total * g.P) / (double)g.nm);
    loss_reduce_kernel<<<1, 1024, 0, st>>>(partial, total, scale, loss);
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}

extern "C" int msst_simmim_decode_l1_bwd(const msst_decode_dims* d, const float* enc, const int64_t* idx, const float* img,
                                         const float* target_tokens, const float* W, const float* bias, const float* d_loss,
                                         float* d_enc, float* d_W, float* d_bias, float* d_target_tokens, msst_stream_t stream) {
    DecGeom g;
    if (int rc = make_geom(d, img, target_tokens, g)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = (int64_t)g.B * g.nm;
    MSST_REQUIRE(g.P <= 16, "simmim_decode_bwd: pixels per patch %d > 16 unsupported", g.P);
    const size_t smem = sizeof(float) * ((size_t)g.n_wb * g.P * g.D + (size_t)g.n_wb * g.P);
    const int use_smem = smem <= 100 * 1024;     // two CTAs per SM
    // contiguous runs of items per warp: long enough to amortise the flushes, short enough to fill 2 CTAs x 148 SMs
    int64_t ipw = ceil_div(total, (int64_t)2 * kNumSMs * (DT / 32));
    if (ipw < 8) ipw = 8;
    const int64_t grid = ceil_div(ceil_div(total, ipw), DT / 32);
    const float coef = (float)(1.0 / ((double)total * g.P) / (double)g.nm);
#define MSST_D(NJ)                                                                                                              \
    case NJ:                                                                                                                    \
        if (use_smem) MSST_CUDA(cudaFuncSetAttribute(decode_bwd_kernel<NJ, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        decode_bwd_kernel<NJ, 16><<<(int)grid, DT, use_smem ? smem : 0, st>>>(g, enc, idx, img, target_tokens, W, bias, d_loss, coef, \
                                                                             d_enc, d_W, d_bias, d_target_tokens, use_smem, (int)ipw); \
        break;
    switch (g.D / 32) { MSST_D(1) MSST_D(2) MSST_D(3) MSST_D(4) default: set_error("simmim_decode: D unsupported"); return MSST_ERR_ARG; }
#undef MSST_D
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}
