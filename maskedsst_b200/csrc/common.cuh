// Shared device/host helpers for the MaskedSST B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/msst.h"

namespace msst {

// ---- error plumbing: no exception crosses the C boundary (SURVEY.md §8(b) "Errors") ----
void set_error(const char* fmt, ...);
inline int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
    return MSST_ERR_CUDA;
}
#define MSST_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) return ::msst::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)
void count_launch();
#define MSST_LAUNCH_CHECK() do { ::msst::count_launch(); MSST_CUDA(cudaPeekAtLastError()); } while (0)
#define MSST_REQUIRE(cond, ...)                                                \
    do {                                                                       \
        if (!(cond)) { ::msst::set_error(__VA_ARGS__); return MSST_ERR_ARG; }  \
    } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// cudaFuncSetAttribute is per device: one-time setup guards are kept per device ordinal (one process may drive several GPUs)
struct PerDeviceOnce {
    unsigned long long done = 0ull;
    bool first() {
        int dev = 0;
        cudaGetDevice(&dev);
        const unsigned long long bit = 1ull << (dev & 63);
        if (done & bit) return false;
        done |= bit;
        return true;
    }
};

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- warp helpers ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// exact (erf) GELU and its derivative: nn.GELU() default, reference src/vit_spatial_spectral.py:37
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// bf16-mode variants: erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7, far below bf16 resolution), ~3x fewer instructions
__device__ __forceinline__ float erf_as(float x) {
    const float ax = fabsf(x);
    const float t = __frcp_rn(fmaf(0.3275911f, ax, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f); p = fmaf(p, t, -0.284496736f); p = fmaf(p, t, 0.254829592f);
    const float r = 1.0f - p * t * __expf(-ax * ax);
    return copysignf(r, x);
}
__device__ __forceinline__ float gelu_fast(float x) { return 0.5f * x * (1.0f + erf_as(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_fast_grad(float x) {
    const float cdf = 0.5f * (1.0f + erf_as(x * 0.70710678118654752440f));
    return fmaf(x * 0.39894228040143267794f, __expf(-0.5f * x * x), cdf);
}

// ---- counter-based dropout RNG (Philox4x32-7: the 7-round variant of Salmon et al., passes BigCrush) ----
// Masks are never stored: forward and backward regenerate them from (seed, site, element index).
// One call yields four 32-bit words for elements 4q..4q+3 of site `site`.
struct Philox {
    static constexpr uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
    __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#ifdef __CUDA_ARCH__
        const uint32_t hi0 = __umulhi(kM0, c[0]), hi1 = __umulhi(kM1, c[2]);
#else
        const uint32_t hi0 = (uint32_t)(((uint64_t)kM0 * c[0]) >> 32), hi1 = (uint32_t)(((uint64_t)kM1 * c[2]) >> 32);
#endif
        const uint32_t lo0 = kM0 * c[0], lo1 = kM1 * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    __host__ __device__ static inline void gen(uint64_t seed, uint32_t site, uint64_t quad, uint32_t (&out)[4]) {
        uint32_t c[4] = {(uint32_t)quad, (uint32_t)(quad >> 32), site, 0x5e5eed5u};
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
        for (int r = 0; r < 7; ++r) { round(c, k0, k1); k0 += kW0; k1 += kW1; }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
};

// Inverted-dropout descriptor. keep element iff rand32 >= thresh, thresh = p * 2^32.
struct Drop {
    uint64_t seed;
    const uint64_t* seed_dev;   // optional device-resident seed offset (lets a captured CUDA graph draw fresh masks per replay)
    uint32_t site;
    uint32_t thresh;   // 0 => dropout disabled
    float scale;       // 1/(1-p)
    __host__ __device__ bool on() const { return thresh != 0u; }
};
inline Drop make_drop(float p, uint64_t seed, uint32_t site, const uint64_t* seed_dev = nullptr) {
    Drop d; d.seed = seed; d.site = site; d.seed_dev = seed_dev;
    if (p <= 0.f) { d.thresh = 0u; d.scale = 1.f; }
    else {
        double t = (double)p * 4294967296.0;
        d.thresh = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
        if (d.thresh == 0u) d.thresh = 1u;
        d.scale = 1.f / (1.f - p);
    }
    return d;
}
// factor (0 or scale) for a single element index e of the site
__device__ __forceinline__ float drop_factor(const Drop& d, uint64_t e) {
    uint32_t r[4];
    Philox::gen(d.seed + (d.seed_dev ? __ldg(d.seed_dev) : 0ull), d.site, e >> 2, r);
    return r[e & 3] >= d.thresh ? d.scale : 0.f;
}
// factors for the aligned quad 4q..4q+3
__device__ __forceinline__ void drop_factor4(const Drop& d, uint64_t quad, float (&f)[4]) {
    uint32_t r[4];
    Philox::gen(d.seed + (d.seed_dev ? __ldg(d.seed_dev) : 0ull), d.site, quad, r);
#pragma unroll
    for (int i = 0; i < 4; ++i) f[i] = r[i] >= d.thresh ? d.scale : 0.f;
}

// dropout site ids (distinct Philox streams); per transformer layer: site = base + layer*8 + k
enum DropSite : uint32_t { kSiteEmb = 1, kSiteLayerBase = 16, kSiteAttnProb = 0, kSiteAttnOut = 1, kSiteMlpHidden = 2, kSiteMlpOut = 3 };

}  // namespace msst
