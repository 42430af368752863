// Kernel (2): tcgen05 / TMEM tensor-core GEMMs fed by TMA (sm_100a), bf16 operands, fp32 accumulation.
//
//   gemm_tn_kernel   C[M,N] = epilogue(A[M,K] . B[N,K]^T)        both operands K-major (row-major, K contiguous)
//                    -> forward linears (A = activations, B = W) and data gradients (A = dY, B = W^T copy)
//                    epilogue: +bias, GELU(erf) / GELU' (x aux), Philox dropout, +fp32 residual, bf16 or fp32 store
//   gemm_wgrad_kernel  dW[N,K] += dY[M,N]^T . X[M,K]              both operands MN-major (the reduction dim M is the
//                    row dim in memory), split over M across CTAs, fp32 red.global epilogue
//
// Reference call sites: nn.Linear to_qkv / to_out / net.0 / net.3 (src/vit_spatial_spectral.py:35-41,59-65) and their
// autograd.  Structure: persistent warp-specialised CTA (1 per SM): warp 0 = TMA producer, warp 1 = MMA issuer
// (one elected thread), warp 2 = TMEM allocator, warps 4-7 = epilogue (TMEM -> registers -> global).
// smem ring of 4 stages (A 128x64 + B BNx64 bf16, SWIZZLE_128B), two TMEM accumulator stages so the epilogue of
// tile i overlaps the loads + MMAs of tile i+1.  K is tiny in this model (64..512, 1536 for one dgrad), the kernels
// are HBM-bound: algorithmic bytes per launch = 2(MK + NK) + out_bytes*MN (+ 4MN residual).
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace msst {
using namespace ptx;

constexpr int GB_M = 128;          // UMMA M (rows of A per tile) -- accumulator row i lives in TMEM lane i
constexpr int GB_K = 64;           // bf16 elements per k-block = 128 B = one SWIZZLE_128B row
constexpr int GB_STAGES = 4;
constexpr int GB_THREADS = 256;
constexpr uint32_t GB_A_BYTES = GB_M * GB_K * 2;   // 16 KB

struct GemmTnParams {
    int64_t M; int N, K;
    int block_n, num_kb, tiles_n; int64_t tiles_m;
    uint32_t idesc, tmem_cols;
    const float* bias; const float* residual; void* out; int out_fp32;
    __nv_bfloat16* pre_act; const __nv_bfloat16* aux; int act; Drop drop;
};

struct alignas(8) GemmBars {
    uint64_t full[GB_STAGES], empty[GB_STAGES], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(GB_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmTnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t b_bytes = (uint32_t)p.block_n * GB_K * 2;
    const uint32_t stage_bytes = GB_A_BYTES + b_bytes;
    GemmBars* bars = reinterpret_cast<GemmBars*>(smem + (size_t)GB_STAGES * stage_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t num_tiles = p.tiles_m * p.tiles_n;

    if (warp == 0 && elect_one()) { prefetch_tmap(&tma_a); prefetch_tmap(&tma_b); }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < GB_STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&bars->tmem_full[s], 1); mbar_init(&bars->tmem_empty[s], 4); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&bars->tmem_base, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int n_blk = (int)(tile % p.tiles_n);
                const int64_t m_blk = tile / p.tiles_n;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    mbar_arrive_expect_tx(&bars->full[stage], stage_bytes);
                    tma_load_2d(sa, &tma_a, &bars->full[stage], kb * GB_K, (int)(m_blk * GB_M));
                    tma_load_2d(sa + GB_A_BYTES, &tma_b, &bars->full[stage], kb * GB_K, n_blk * p.block_n);
                    if (++stage == GB_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (single elected thread) =====
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0; int it = 0;
            for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1; const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.block_n);
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&bars->full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t da = make_smem_desc(sa, 16, 1024), db = make_smem_desc(sa + GB_A_BYTES, 16, 1024);
                    const int rem = p.K - kb * GB_K;
                    const int ksteps = rem >= GB_K ? GB_K / 16 : (rem + 15) / 16;
                    for (int k = 0; k < ksteps; ++k)   // +32 B per UMMA_K step inside the 128 B swizzle row
                        umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), p.idesc, (kb | k) != 0);
                    umma_commit(&bars->empty[stage]);
                    if (++stage == GB_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&bars->tmem_full[acc]);
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> global =====
        const int q = warp - 4;   // TMEM lane quarter == warp_idx % 4
        int it = 0;
        const bool vec_ok = (p.N % 8 == 0);
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1; const uint32_t acc_phase = (it >> 1) & 1;
            const int n_blk = (int)(tile % p.tiles_n);
            const int64_t m_blk = tile / p.tiles_n;
            mbar_wait(&bars->tmem_full[acc], acc_phase);
            tc_fence_after();
            const int64_t row = m_blk * GB_M + q * 32 + lane;
            for (int c = 0; c < p.block_n / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.block_n + c * 32), v);
                tmem_ld_wait();
                const int col0 = n_blk * p.block_n + c * 32;
                if (row >= p.M || col0 >= p.N) continue;
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                const int64_t off = row * p.N + col0;
                const bool full = vec_ok && (col0 + 32 <= p.N);
                if (p.bias) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) if (full || col0 + j < p.N) f[j] += __ldg(p.bias + col0 + j);
                }
                if (p.pre_act) {
                    if (full) {
#pragma unroll
                        for (int g = 0; g < 4; ++g)
                            *reinterpret_cast<uint4*>(p.pre_act + off + g * 8) =
                                make_uint4(pack_bf16(f[g * 8], f[g * 8 + 1]), pack_bf16(f[g * 8 + 2], f[g * 8 + 3]),
                                           pack_bf16(f[g * 8 + 4], f[g * 8 + 5]), pack_bf16(f[g * 8 + 6], f[g * 8 + 7]));
                    } else {
                        for (int j = 0; j < 32; ++j) if (col0 + j < p.N) p.pre_act[off + j] = __float2bfloat16(f[j]);
                    }
                }
                if (p.act == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
                } else if (p.act == 2) {
                    if (full) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const uint4 u = *reinterpret_cast<const uint4*>(p.aux + off + g * 8);
                            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[t]);
                                f[g * 8 + 2 * t] *= gelu_erf_grad(__low2float(h));
                                f[g * 8 + 2 * t + 1] *= gelu_erf_grad(__high2float(h));
                            }
                        }
                    } else {
                        for (int j = 0; j < 32; ++j) if (col0 + j < p.N) f[j] *= gelu_erf_grad(__bfloat162float(p.aux[off + j]));
                    }
                }
                if (p.drop.on()) {
                    if ((p.N & 3) == 0) {
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            float d4[4];
                            drop_factor4(p.drop, (uint64_t)(off + g * 4) >> 2, d4);
#pragma unroll
                            for (int t = 0; t < 4; ++t) f[g * 4 + t] *= d4[t];
                        }
                    } else {
                        for (int j = 0; j < 32; ++j) if (col0 + j < p.N) f[j] *= drop_factor(p.drop, (uint64_t)(off + j));
                    }
                }
                if (p.residual) {
                    if (full) {
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            const float4 r = *reinterpret_cast<const float4*>(p.residual + off + g * 4);
                            f[g * 4] += r.x; f[g * 4 + 1] += r.y; f[g * 4 + 2] += r.z; f[g * 4 + 3] += r.w;
                        }
                    } else {
                        for (int j = 0; j < 32; ++j) if (col0 + j < p.N) f[j] += p.residual[off + j];
                    }
                }
                if (p.out_fp32) {
                    float* o = reinterpret_cast<float*>(p.out) + off;
                    if (full) {
#pragma unroll
                        for (int g = 0; g < 8; ++g) *reinterpret_cast<float4*>(o + g * 4) = make_float4(f[g * 4], f[g * 4 + 1], f[g * 4 + 2], f[g * 4 + 3]);
                    } else {
                        for (int j = 0; j < 32; ++j) if (col0 + j < p.N) o[j] = f[j];
                    }
                } else {
                    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
                    if (full) {
#pragma unroll
                        for (int g = 0; g < 4; ++g)
                            *reinterpret_cast<uint4*>(o + g * 8) =
                                make_uint4(pack_bf16(f[g * 8], f[g * 8 + 1]), pack_bf16(f[g * 8 + 2], f[g * 8 + 3]),
                                           pack_bf16(f[g * 8 + 4], f[g * 8 + 5]), pack_bf16(f[g * 8 + 6], f[g * 8 + 7]));
                    } else {
                        for (int j = 0; j < 32; ++j) if (col0 + j < p.N) o[j] = __float2bfloat16(f[j]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, p.tmem_cols); }
}

// ---------------------------------------------------------------------------------------------------------
// weight gradient: dW[N,K] += dY[M,N]^T X[M,K].  UMMA view: D[128 rows of N][BN cols of K] += A^T B with the
// reduction dim = tokens; both operand tiles are [64 tokens][64-element chunks] in smem -> MN-major descriptors
// (LBO = 8 KB between 64-element chunks, SBO = 1 KB between 8-token groups).  grid = (output tiles, splits).
// ---------------------------------------------------------------------------------------------------------
constexpr int WG_TOK = 64;                         // tokens per stage
constexpr uint32_t WG_CHUNK = WG_TOK * 128;        // bytes of one [64 tokens][64 elements] chunk = 8 KB

struct WgradParams {
    int64_t M; int N, K;
    int block_n, n_chunks_b, tiles_n_out, tiles_k_out;     // block_n = columns of K per tile (<= 256)
    int64_t tok_blocks_per_split;
    uint32_t idesc, tmem_cols;
    float* dW;
};

__global__ void __launch_bounds__(GB_THREADS, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tma_dy, const __grid_constant__ CUtensorMap tma_x, const WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = 2 * WG_CHUNK, b_bytes = (uint32_t)p.n_chunks_b * WG_CHUNK;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    GemmBars* bars = reinterpret_cast<GemmBars*>(smem + (size_t)GB_STAGES * stage_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_n = blockIdx.x / p.tiles_k_out, tile_k = blockIdx.x % p.tiles_k_out;
    const int64_t total_tb = (p.M + WG_TOK - 1) / WG_TOK;
    const int64_t tb0 = (int64_t)blockIdx.y * p.tok_blocks_per_split;
    const int64_t tb1 = tb0 + p.tok_blocks_per_split < total_tb ? tb0 + p.tok_blocks_per_split : total_tb;

    if (warp == 0 && elect_one()) { prefetch_tmap(&tma_dy); prefetch_tmap(&tma_x); }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < GB_STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        mbar_init(&bars->tmem_full[0], 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(&bars->tmem_base, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    if (tb0 < tb1) {
        if (warp == 0) {
            if (elect_one()) {
                int stage = 0; uint32_t phase = 0;
                for (int64_t tb = tb0; tb < tb1; ++tb) {
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    uint8_t* s = smem + (size_t)stage * stage_bytes;
                    mbar_arrive_expect_tx(&bars->full[stage], stage_bytes);
                    const int tok = (int)(tb * WG_TOK);
                    tma_load_2d(s, &tma_dy, &bars->full[stage], tile_n * GB_M, tok);
                    tma_load_2d(s + WG_CHUNK, &tma_dy, &bars->full[stage], tile_n * GB_M + 64, tok);
                    for (int c = 0; c < p.n_chunks_b; ++c)
                        tma_load_2d(s + a_bytes + c * WG_CHUNK, &tma_x, &bars->full[stage], tile_k * p.block_n + c * 64, tok);
                    if (++stage == GB_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp == 1) {
            if (elect_one()) {
                int stage = 0; uint32_t phase = 0;
                for (int64_t tb = tb0; tb < tb1; ++tb) {
                    mbar_wait(&bars->full[stage], phase);
                    tc_fence_after();
                    const uint32_t s = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t da = make_smem_desc(s, WG_CHUNK, 1024), db = make_smem_desc(s + a_bytes, WG_CHUNK, 1024);
                    for (int k = 0; k < WG_TOK / 16; ++k)   // 16 tokens = 2 KB per UMMA_K step
                        umma_bf16(tmem_base, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), p.idesc, (tb > tb0) || k != 0);
                    umma_commit(&bars->empty[stage]);
                    if (++stage == GB_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&bars->tmem_full[0]);
            }
        } else if (warp >= 4) {
            const int q = warp - 4;
            mbar_wait(&bars->tmem_full[0], 0);
            tc_fence_after();
            const int n = tile_n * GB_M + q * 32 + lane;     // row of dW
            for (int c = 0; c < p.block_n / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
                tmem_ld_wait();
                const int k0 = tile_k * p.block_n + c * 32;
                if (n < p.N) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (k0 + j < p.K) atomicAdd(p.dW + (int64_t)n * p.K + k0 + j, __uint_as_float(v[j]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, p.tmem_cols); }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// 2-D bf16 tensor [rows, cols] (row pitch ld elements), box [box_rows, 64 cols], SWIZZLE_128B, OOB -> zeros
static int make_tmap(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    MSST_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    MSST_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 2) % 16 == 0, "TMA operand must be 16-byte aligned with a 16-byte multiple pitch");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSST_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d (rows %lld cols %lld ld %lld box %d)", (int)r,
                 (long long)rows, (long long)cols, (long long)ld, box_rows);
    return MSST_OK;
}

static uint32_t pow2_cols(int c) { uint32_t v = 32; while ((int)v < c) v <<= 1; return v; }

int gemm_tn_bf16(const GemmBf16Args& a, cudaStream_t st) {
    if (a.M == 0) return MSST_OK;
    MSST_REQUIRE(a.K % 8 == 0 && a.K >= 8, "bf16 GEMM: K=%d must be a multiple of 8", a.K);
    MSST_REQUIRE(a.N >= 1, "bf16 GEMM: bad N");
    GemmTnParams p{};
    p.M = a.M; p.N = a.N; p.K = a.K;
    const int n32 = (a.N + 31) / 32 * 32;
    p.block_n = n32 < 256 ? n32 : 256;
    // balance tiles along N (e.g. N = 1536 -> 6 x 256)
    p.tiles_n = (a.N + p.block_n - 1) / p.block_n;
    p.tiles_m = (a.M + GB_M - 1) / GB_M;
    p.num_kb = (a.K + GB_K - 1) / GB_K;
    p.idesc = make_idesc_bf16(GB_M, p.block_n, 0, 0);
    p.tmem_cols = pow2_cols(2 * p.block_n);
    p.bias = a.bias; p.residual = a.residual; p.out = a.out; p.out_fp32 = a.out_fp32; p.pre_act = a.pre_act; p.aux = a.aux;
    p.act = a.act; p.drop = a.drop;
    CUtensorMap ta, tb;
    if (int rc = make_tmap(&ta, a.A, a.M, a.K, a.K, GB_M)) return rc;
    if (int rc = make_tmap(&tb, a.B, a.N, a.K, a.K, p.block_n)) return rc;
    const size_t smem = (size_t)GB_STAGES * (GB_A_BYTES + (size_t)p.block_n * GB_K * 2) + sizeof(GemmBars) + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        MSST_CUDA(cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const int64_t tiles = p.tiles_m * p.tiles_n;
    const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
    gemm_tn_kernel<<<grid, GB_THREADS, smem, st>>>(ta, tb, p);
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}

int gemm_wgrad_bf16(const __nv_bfloat16* dy, const __nv_bfloat16* x, float* dW, int64_t M, int N, int K, cudaStream_t st) {
    if (M == 0) return MSST_OK;
    MSST_REQUIRE(N % 8 == 0 && K % 8 == 0, "bf16 wgrad: N=%d and K=%d must be multiples of 8", N, K);
    WgradParams p{};
    p.M = M; p.N = N; p.K = K; p.dW = dW;
    const int k32 = (K + 31) / 32 * 32;
    p.block_n = k32 < 256 ? k32 : 256;
    p.n_chunks_b = (p.block_n + 63) / 64;
    p.tiles_n_out = (N + GB_M - 1) / GB_M;
    p.tiles_k_out = (K + p.block_n - 1) / p.block_n;
    p.idesc = make_idesc_bf16(GB_M, p.block_n, 1, 1);
    p.tmem_cols = pow2_cols(p.block_n);
    const int out_tiles = p.tiles_n_out * p.tiles_k_out;
    const int64_t total_tb = (M + WG_TOK - 1) / WG_TOK;
    int64_t splits = kNumSMs / out_tiles;
    if (splits < 1) splits = 1;
    if (splits > total_tb) splits = total_tb;
    p.tok_blocks_per_split = (total_tb + splits - 1) / splits;
    splits = (total_tb + p.tok_blocks_per_split - 1) / p.tok_blocks_per_split;
    CUtensorMap ta, tb;
    if (int rc = make_tmap(&ta, dy, M, N, N, WG_TOK)) return rc;
    if (int rc = make_tmap(&tb, x, M, K, K, WG_TOK)) return rc;
    const size_t smem = (size_t)GB_STAGES * (2 * WG_CHUNK + (size_t)p.n_chunks_b * WG_CHUNK) + sizeof(GemmBars) + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        MSST_CUDA(cudaFuncSetAttribute(gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    gemm_wgrad_kernel<<<dim3(out_tiles, (unsigned)splits), GB_THREADS, smem, st>>>(ta, tb, p);
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}

}  // namespace msst
