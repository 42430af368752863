// fp32 (FFMA) GEMM family for the MSST_PREC_FP32 parity mode: forward linear, data-gradient, weight-gradient
// (split-K), with fused bias / GELU(erf) / dropout / residual / GELU-backward epilogues.
// Reference call sites: nn.Linear in Attention.to_qkv/.to_out and FeedForward.net (src/vit_spatial_spectral.py:35-41,59-65)
// and their autograd (mm backward).  This is the correctness mode (rel-err <= 1e-5); the throughput path is the
// tcgen05 bf16 GEMM in gemm_bf16_tcgen05.cu.
#include "common.cuh"
#include "kernels.h"

namespace msst {

constexpr int BM = 64, BN = 64, BK = 16, GT = 256;

struct GemmArgs {
    const float* A; const float* B; float* C;
    int64_t M; int N; int64_t K;          // C[M,N] = sum_k A(m,k) B(k,n)
    int64_t lda, ldb, ldc;
    int64_t k_per_split;                  // multiple of BK
    int atomic;                           // split-K: atomicAdd into C (epilogue ops other than bias-free sum are off)
    const float* bias; const float* residual; float* pre_act; const float* aux;
    int act;                              // 0 none, 1 gelu, 2 multiply by gelu'(aux)
    Drop drop;
};

template <bool KC, bool VEC>
__device__ __forceinline__ void load_tile(const float* __restrict__ P, int64_t ld, int64_t r0, int64_t rmax, int64_t k0,
                                          int64_t kmax, float (&reg)[4]) {
    // fetches 4 elements of a [64 rows x 16 k] operand tile into registers.
    // KC  (k contiguous):  thread -> row = t/4,  k = (t%4)*4 .. +3   element (row,k) = P[row*ld + k]
    // !KC (row contiguous): thread -> k = t/16, row = (t%16)*4 .. +3  element (row,k) = P[k*ld + row]
    const int t = threadIdx.x;
    if (KC) {
        const int64_t row = r0 + t / 4, k = k0 + (t % 4) * 4;
        if (row < rmax) {
            const float* p = P + row * ld + k;
            if (VEC && k + 3 < kmax) { const float4 v = *reinterpret_cast<const float4*>(p); reg[0] = v.x; reg[1] = v.y; reg[2] = v.z; reg[3] = v.w; }
            else {
#pragma unroll
                for (int i = 0; i < 4; ++i) reg[i] = (k + i < kmax) ? p[i] : 0.f;
            }
        } else { reg[0] = reg[1] = reg[2] = reg[3] = 0.f; }
    } else {
        const int64_t k = k0 + t / 16, row = r0 + (t % 16) * 4;
        if (k < kmax) {
            const float* p = P + k * ld + row;
            if (VEC && row + 3 < rmax) { const float4 v = *reinterpret_cast<const float4*>(p); reg[0] = v.x; reg[1] = v.y; reg[2] = v.z; reg[3] = v.w; }
            else {
#pragma unroll
                for (int i = 0; i < 4; ++i) reg[i] = (row + i < rmax) ? p[i] : 0.f;
            }
        } else { reg[0] = reg[1] = reg[2] = reg[3] = 0.f; }
    }
}

template <bool KC>
__device__ __forceinline__ void store_tile(float (*S)[BM + 4], const float (&reg)[4]) {
    const int t = threadIdx.x;
    if (KC) {
        const int row = t / 4, k = (t % 4) * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) S[k + i][row] = reg[i];
    } else {
        const int k = t / 16, row = (t % 16) * 4;
        *reinterpret_cast<float4*>(&S[k][row]) = make_float4(reg[0], reg[1], reg[2], reg[3]);
    }
}

template <bool A_KC, bool B_KC, bool VEC>
__global__ void __launch_bounds__(GT) sgemm_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int64_t kb = (int64_t)blockIdx.z * g.k_per_split;
    const int64_t ke = kb + g.k_per_split < g.K ? kb + g.k_per_split : g.K;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    float acc[4][4] = {};
    float ra[4], rb[4];
    if (kb < ke) {
        load_tile<A_KC, VEC>(g.A, g.lda, m0, g.M, kb, ke, ra);
        load_tile<B_KC, VEC>(g.B, g.ldb, n0, g.N, kb, ke, rb);
    }
    for (int64_t k0 = kb; k0 < ke; k0 += BK) {
        __syncthreads();
        store_tile<A_KC>(As, ra);
        store_tile<B_KC>(Bs, rb);
        __syncthreads();
        if (k0 + BK < ke) {
            load_tile<A_KC, VEC>(g.A, g.lda, m0, g.M, k0 + BK, ke, ra);
            load_tile<B_KC, VEC>(g.B, g.ldb, n0, g.N, k0 + BK, ke, rb);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
    // ---- epilogue ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
        const int n = n0 + tx * 4;
        if (n >= g.N) continue;
        float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
        const int64_t off = m * g.ldc + n;
        if (g.atomic) {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (n + j < g.N) atomicAdd(g.C + off + j, v[j]);
            continue;
        }
        const bool full = VEC && (n + 3 < g.N);
        if (g.bias) {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (n + j < g.N) v[j] += g.bias[n + j];
        }
        if (g.pre_act) {
            if (full) *reinterpret_cast<float4*>(g.pre_act + off) = make_float4(v[0], v[1], v[2], v[3]);
            else for (int j = 0; j < 4; ++j) if (n + j < g.N) g.pre_act[off + j] = v[j];
        }
        if (g.act == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = gelu_erf(v[j]);
        } else if (g.act == 2) {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (n + j < g.N) v[j] *= gelu_erf_grad(g.aux[off + j]);
        }
        if (g.drop.on()) {
            // element index = m*N + n (ldc == N for every dropout site); quads are aligned because N % 4 == 0
            if ((g.N & 3) == 0) {
                float f[4];
                drop_factor4(g.drop, (uint64_t)(m * g.N + n) >> 2, f);
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] *= f[j];
            } else {
                for (int j = 0; j < 4; ++j) if (n + j < g.N) v[j] *= drop_factor(g.drop, (uint64_t)(m * g.N + n + j));
            }
        }
        if (g.residual) {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (n + j < g.N) v[j] += g.residual[off + j];
        }
        if (full) *reinterpret_cast<float4*>(g.C + off) = make_float4(v[0], v[1], v[2], v[3]);
        else for (int j = 0; j < 4; ++j) if (n + j < g.N) g.C[off + j] = v[j];
    }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <bool A_KC, bool B_KC>
static int launch(const GemmArgs& g, int splits, cudaStream_t st) {
    const bool vec = (g.lda % 4 == 0) && (g.ldb % 4 == 0) && (g.ldc % 4 == 0) && aligned16(g.A) && aligned16(g.B) &&
                     aligned16(g.C) && (!g.residual || aligned16(g.residual)) && (!g.pre_act || aligned16(g.pre_act));
    dim3 grid((unsigned)ceil_div(g.M, BM), (unsigned)ceil_div(g.N, BN), (unsigned)splits);
    MSST_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "gemm: grid too large");
    if (vec) sgemm_kernel<A_KC, B_KC, true><<<grid, GT, 0, st>>>(g);
    else sgemm_kernel<A_KC, B_KC, false><<<grid, GT, 0, st>>>(g);
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}

// y[M,N] = epi(x[M,K] @ W[N,K]^T)
int linear_fwd_f32(const float* x, const float* W, const float* bias, const float* residual, float* y, float* pre_act,
                   int64_t M, int N, int K, int act, Drop drop, cudaStream_t st) {
    if (M == 0) return MSST_OK;
    GemmArgs g{};
    g.A = x; g.B = W; g.C = y; g.M = M; g.N = N; g.K = K; g.lda = K; g.ldb = K; g.ldc = N;
    g.k_per_split = ceil_div(K, BK) * BK; g.atomic = 0;
    g.bias = bias; g.residual = residual; g.pre_act = pre_act; g.aux = nullptr; g.act = act; g.drop = drop;
    return launch<true, true>(g, 1, st);
}

// dx[M,K] = epi(dy[M,N] @ W[N,K]); epi: * gelu'(pre_act) (act=2), * dropout factor (hidden site), + dx_add
int linear_bwd_data_f32(const float* dy, const float* W, const float* pre_act, const float* dx_add, float* dx, int64_t M,
                        int N, int K, Drop drop, cudaStream_t st) {
    if (M == 0) return MSST_OK;
    GemmArgs g{};
    g.A = dy; g.B = W; g.C = dx; g.M = M; g.N = K; g.K = N; g.lda = N; g.ldb = K; g.ldc = K;
    g.k_per_split = ceil_div(N, BK) * BK; g.atomic = 0;
    g.bias = nullptr; g.residual = dx_add; g.pre_act = nullptr; g.aux = pre_act; g.act = pre_act ? 2 : 0; g.drop = drop;
    return launch<true, false>(g, 1, st);
}

__global__ void colsum_kernel(const float* __restrict__ dy, float* __restrict__ db, int64_t M, int N, int64_t rows_per_block) {
    // grid.x tiles columns in groups of 32, grid.y splits rows; block (32, 8)
    __shared__ float red[8][33];
    const int n = blockIdx.x * 32 + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
    float s = 0.f;
    if (n < N) for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) s += dy[r * N + n];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
        atomicAdd(db + n, t);
    }
}

int colsum_f32(const float* dy, float* db, int64_t M, int N, cudaStream_t st) {
    if (M == 0) return MSST_OK;
    const int gx = (int)ceil_div(N, 32);
    int gy = (int)ceil_div(4 * kNumSMs, gx);
    const int64_t max_gy = ceil_div(M, 64);
    if (gy > max_gy) gy = (int)max_gy;
    if (gy < 1) gy = 1;
    const int64_t rpb = ceil_div(M, gy);
    colsum_kernel<<<dim3(gx, gy), dim3(32, 8), 0, st>>>(dy, db, M, N, rpb);
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}

// dW[N,K] += dy[M,N]^T @ x[M,K]  (reduction over M split across CTAs, fp32 atomics), db[N] += colsum(dy)
int linear_bwd_weight_f32(const float* dy, const float* x, float* dW, float* db, int64_t M, int N, int K, cudaStream_t st) {
    if (M == 0) return MSST_OK;
    GemmArgs g{};
    g.A = dy; g.B = x; g.C = dW; g.M = N; g.N = K; g.K = M; g.lda = N; g.ldb = K; g.ldc = K;
    const int64_t tiles = ceil_div(N, BM) * ceil_div(K, BN);
    int64_t splits = ceil_div(4 * kNumSMs, tiles);
    const int64_t max_splits = ceil_div(M, 4 * BK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    g.k_per_split = ceil_div(ceil_div(M, splits), BK) * BK;
    splits = ceil_div(M, g.k_per_split);
    g.atomic = 1;
    g.bias = nullptr; g.residual = nullptr; g.pre_act = nullptr; g.aux = nullptr; g.act = 0; g.drop = make_drop(0.f, 0, 0);
    if (int rc = launch<false, false>(g, (int)splits, st)) return rc;
    if (db) return colsum_f32(dy, db, M, N, st);
    return MSST_OK;
}

__global__ void dropout_apply_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n4, Drop drop) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (int64_t)gridDim.x * blockDim.x) {
        float f[4];
        drop_factor4(drop, (uint64_t)q, f);
        const float4 v = reinterpret_cast<const float4*>(x)[q];
        reinterpret_cast<float4*>(y)[q] = make_float4(v.x * f[0], v.y * f[1], v.z * f[2], v.w * f[3]);
    }
}

// y = x * dropout_factor(site, element index); n must be a multiple of 4 (every site has N % 4 == 0)
int dropout_apply_f32(const float* x, float* y, int64_t n, Drop drop, cudaStream_t st) {
    MSST_REQUIRE(n % 4 == 0, "dropout_apply: n must be a multiple of 4");
    if (n == 0) return MSST_OK;
    int64_t blocks = ceil_div(n / 4, 256);
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    dropout_apply_kernel<<<(int)blocks, 256, 0, st>>>(x, y, n / 4, drop);
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}

}  // namespace msst
