// Small HBM-bound helpers of the bf16 path:
//   cast_rows   : fp32|bf16 [M,N] -> (optional counter-hash dropout) -> bf16 copy (optional) + column sums (optional, bias grads)
//   weights_bf16: fp32 weight matrices -> bf16 copy and bf16 transposed copy (operands of the forward / data-gradient GEMMs)
#include "common.cuh"
#include "kernels.h"

namespace msst {

// One thread owns a fixed set of 4 consecutive columns: block = (N/4) x rows_per_block threads, grid-stride over rows,
// so one RNG call serves one aligned quad and the column sums stay in registers until one flush per block.
template <typename TIn>
__global__ void cast_rows_kernel(const TIn* __restrict__ x, __nv_bfloat16* __restrict__ y, float* __restrict__ colsum, int64_t M, int N,
                                 int rows_per_block, Drop drop) {
    extern __shared__ float red[];   // [rows_per_block][N]
    const int tpr = N / 4;
    const int cq = threadIdx.x % tpr, rb = threadIdx.x / tpr;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t r = (int64_t)blockIdx.x * rows_per_block + rb; r < M; r += (int64_t)gridDim.x * rows_per_block) {
        const int64_t off = r * N + cq * 4;
        float v[4];
        if (sizeof(TIn) == 4) {
            const float4 t = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + off);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
            const uint2 t = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(x) + off);
            const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x), b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
            v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
        }
        if (drop.on()) {
            float f[4];
            drop_factor4(drop, (uint64_t)off >> 2, f);
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] *= f[i];
        }
        if (y) {
            __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
            uint2 t;
            t.x = *reinterpret_cast<uint32_t*>(&a); t.y = *reinterpret_cast<uint32_t*>(&b);
            *reinterpret_cast<uint2*>(y + off) = t;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] += v[i];
    }
    if (colsum) {
#pragma unroll
        for (int i = 0; i < 4; ++i) red[rb * N + cq * 4 + i] = acc[i];
        __syncthreads();
        for (int c = threadIdx.x; c < N; c += blockDim.x) {
            float s = 0.f;
            for (int k = 0; k < rows_per_block; ++k) s += red[k * N + c];
            atomicAdd(colsum + c, s);
        }
    }
}

template <typename TIn>
static int cast_rows_launch(const TIn* x, __nv_bfloat16* y, float* colsum, int64_t M, int N, Drop drop, cudaStream_t st) {
    MSST_REQUIRE(N % 4 == 0 && N <= 4096, "cast_rows: N=%d must be a multiple of 4 and <= 4096", N);
    if (M == 0) return MSST_OK;
    const int tpr = N / 4;
    int rpb = 256 / tpr; if (rpb < 1) rpb = 1;
    const int threads = tpr * rpb;
    int64_t grid = ceil_div(M, rpb);
    const int64_t cap = colsum ? 4 * kNumSMs : 16 * kNumSMs;
    if (grid > cap) grid = cap;
    const size_t smem = colsum ? sizeof(float) * (size_t)rpb * N : 0;
    cast_rows_kernel<TIn><<<(int)grid, threads, smem, st>>>(x, y, colsum, M, N, rpb, drop);
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}
int cast_rows_f32(const float* x, __nv_bfloat16* y, float* colsum, int64_t M, int N, Drop drop, cudaStream_t st) {
    return cast_rows_launch<float>(x, y, colsum, M, N, drop, st);
}
int cast_rows_bf16(const __nv_bfloat16* x, __nv_bfloat16* y, float* colsum, int64_t M, int N, Drop drop, cudaStream_t st) {
    return cast_rows_launch<__nv_bfloat16>(x, y, colsum, M, N, drop, st);
}

__global__ void weights_bf16_kernel(WeightCastTable t) {
    // blockIdx.y = matrix, 32x32 tiles through smem so both the straight and the transposed copy store coalesced
    __shared__ float tile[32][33];
    const WeightCastEntry e = t.e[blockIdx.y];
    const int tiles_c = (e.cols + 31) / 32, tiles_r = (e.rows + 31) / 32;
    for (int tl = blockIdx.x; tl < tiles_c * tiles_r; tl += gridDim.x) {
        const int r0 = (tl / tiles_c) * 32, c0 = (tl % tiles_c) * 32;
        __syncthreads();
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            const int r = r0 + i, c = c0 + threadIdx.x;
            float v = 0.f;
            if (r < e.rows && c < e.cols) { v = e.src[(int64_t)r * e.cols + c]; e.dst[(int64_t)r * e.cols + c] = __float2bfloat16(v); }
            tile[i][threadIdx.x] = v;
        }
        __syncthreads();
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            const int c = c0 + i, r = r0 + threadIdx.x;
            if (r < e.rows && c < e.cols) e.dst_t[(int64_t)c * e.rows + r] = __float2bfloat16(tile[threadIdx.x][i]);
        }
    }
}

int weights_to_bf16(const WeightCastTable& t, cudaStream_t st) {
    if (t.n == 0) return MSST_OK;
    weights_bf16_kernel<<<dim3(48, t.n), dim3(32, 8), 0, st>>>(t);
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}

}  // namespace msst
