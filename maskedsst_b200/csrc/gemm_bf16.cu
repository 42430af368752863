// Kernel (2): tcgen05 / TMEM tensor-core GEMMs fed by TMA (sm_100a), bf16 operands, fp32 accumulation.
//
//   gemm_tn_kernel   C[M,N] = epilogue(A[M,K] . B[N,K]^T)        both operands K-major (row-major, K contiguous)
//                    -> forward linears (A = activations, B = W) and data gradients (A = dY, B = W^T copy)
//                    epilogue: +bias, GELU(erf) / GELU' (x aux), counter-hash dropout, +fp32 residual, bf16 or fp32 store
//   gemm_wgrad_kernel  dW[N,K] += dY[M,N]^T . X[M,K]              both operands MN-major (the reduction dim M is the
//                    row dim in memory), split over M across CTAs, fp32 red.global epilogue
//
// Reference call sites: nn.Linear to_qkv / to_out / net.0 / net.3 (src/vit_spatial_spectral.py:35-41,59-65) and their
// autograd.  Structure: persistent warp-specialised CTA (1 per SM): warp 0 = TMA producer, warp 1 = MMA issuer
// (one elected thread), warp 2 = TMEM allocator, warps 4-19 = epilogue (TMEM -> registers -> smem transpose -> global).
// smem ring of 4 stages (A 128x64 + B BNx64 bf16, SWIZZLE_128B), two TMEM accumulator stages so the epilogue of
// tile i overlaps the loads + MMAs of tile i+1.  K is tiny in this model (64..512, 1536 for one dgrad), the kernels
// are HBM-bound: algorithmic bytes per launch = 2(MK + NK) + out_bytes*MN (+ 4MN residual).
#include "common.cuh"
#include "kernels.h"
#include "ptx.cuh"
#include <stdlib.h>

namespace msst {
using namespace ptx;

constexpr int GB_M = 128;          // UMMA M (rows of A per tile) -- accumulator row i lives in TMEM lane i
constexpr int GB_K = 64;           // bf16 elements per k-block = 128 B = one SWIZZLE_128B row
constexpr int GB_STAGES = 6;            // upper bound of the operand ring; the depth actually used (p.stages) is what fits next to the epilogue staging
constexpr int GB_EPI_WARPS = 16;      // 4 per TMEM lane quarter: the epilogue math (GELU, dropout hash) is the throughput limiter
constexpr int GB_THREADS = 128 + 32 * GB_EPI_WARPS;   // TMA, MMA, TMEM-alloc, spare + epilogue warps
// accumulator columns per epilogue step: 16 for the math-heavy epilogues (GELU / dropout / residual: more warps busy on the
// N = 64 / 96 tiles), 32 for the store-only ones (whole 64/128-byte row segments per store instruction)
__host__ __device__ constexpr int gb_chunk(int mode) { return (mode == 2 || mode == 3 || mode == 4 || mode == 6) ? 16 : 32; }
__host__ __device__ constexpr int gb_epi_smem(int mode) { return (mode == 7 || mode == 8) ? 0 : GB_EPI_WARPS * 32 * (gb_chunk(mode) + 4) * 4; }
constexpr uint32_t GB_A_BYTES = GB_M * GB_K * 2;   // 16 KB

struct GemmTnParams {
    int64_t M; int N, K;
    int block_n, num_kb, tiles_n; int64_t tiles_m;
    uint32_t idesc, tmem_cols;
    const float* bias; const float* residual; void* out; int out_fp32;
    __nv_bfloat16* pre_act; const __nv_bfloat16* aux; int act; Drop drop;
    const float* ln_w; const float* ln_b; __nv_bfloat16* ln_out; float* ln_stats;   // MODE 6: fused LayerNorm of the output rows
    const float* lnb_x; const float* lnb_stats; float* lnb_dw; float* lnb_db; __nv_bfloat16* lnb_cast; float* lnb_colsum;   // MODE 8: fused LayerNorm backward
    int debug;   // MSST_GEMM_DEBUG bits (profiling experiments only): 1 skip global stores, 2 skip epilogue body, 4 skip MMA issue
    float* colsum;   // MODE 4, block_n <= 64: column sums of the (fp32, post-dropout) output rows += here (bias gradient of the hidden layer)
    int stages;  // operand ring depth actually used (<= GB_STAGES; MODE 7 trades one stage for the output staging tiles)
};

struct alignas(8) GemmBars {
    uint64_t full[GB_STAGES], empty[GB_STAGES], tmem_full[2], tmem_empty[2];
    uint64_t stg_full[2], stg_free[2];     // MODE 7: staging tiles written / read by the TMA store
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}


// Epilogue body for one staged chunk (32 rows x 16 fp32 in smem): lane -> (row sub_r + 8i, columns c4..c4+3).
// MODE selects a compile-time specialisation of the common operator shapes (lean instruction stream -- the epilogue is
// the throughput limiter of these small-K GEMMs); MODE 5 is the fully general path.
//   0: bf16 out                         (QKV projection, dO data-gradient)
//   1: fp32 out                         (data-gradients that feed LayerNorm backward)
//   2: fp32 out = acc + bias [dropout] + residual      (out-projection, MLP second linear)
//   3: bf16 out = [dropout] gelu(acc + bias), pre_act = acc + bias   (MLP first linear)
//   4: bf16 out = acc * gelu'(aux) [dropout]            (data-gradient through the MLP hidden layer)
//   7: bf16 out through swizzled staging tiles + TMA stores (wide store-only outputs)
//   8: fp32 out = LayerNorm-backward(acc) + residual, LN weight/bias gradients, optional bf16 dropout cast (see GemmBf16Args)
//   6: MODE 2 + LayerNorm of the finished rows (bf16) + its statistics: the pre-norm of the NEXT sub-block is produced by
//      the GEMM that finishes the residual stream row (needs the whole row in one tile: N == block_n <= 128)
// residual (MODE 2) / GELU' argument (MODE 4) of this lane's 8 row segments, fetched BEFORE the accumulator is waited for:
// the loads must not sit between the stores of the main loop (possible aliasing would serialise them on DRAM latency).
template <int MODE>
__device__ __forceinline__ void epilogue_prefetch(const GemmTnParams& p, int64_t row_base, int col0, int sub_r, int c4,
                                                  float4 (&pre)[gb_chunk(MODE) / 4]) {
    if (MODE != 2 && MODE != 4 && MODE != 6) return;
    constexpr int ITER = gb_chunk(MODE) / 4, RPI = 128 / gb_chunk(MODE);   // row iterations, rows per iteration
    const int col = col0 + c4;
    const int64_t rows_left = p.M - row_base;
#pragma unroll
    for (int i = 0; i < ITER; ++i) {
        const int rr = i * RPI + sub_r;
        pre[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rr < rows_left && col < p.N) {
            const int64_t off = (row_base + rr) * p.N + col;
            if (MODE == 2 || MODE == 6) pre[i] = *reinterpret_cast<const float4*>(p.residual + off);
            else { const uint2 u = *reinterpret_cast<const uint2*>(p.aux + off); pre[i].x = __uint_as_float(u.x); pre[i].y = __uint_as_float(u.y); }
        }
    }
}

template <int MODE>
__device__ __forceinline__ void epilogue_rows(const GemmTnParams& p, const float* stg, int64_t row_base, int col0, int sub_r, int c4,
                                              const float4 (&pre)[gb_chunk(MODE) / 4], float* rowbuf, int rowbuf_pitch, float (&csum)[4]) {
    constexpr int ITER = gb_chunk(MODE) / 4, RPI = 128 / gb_chunk(MODE), PITCH = gb_chunk(MODE) + 4;
    if (MODE == 0 && (p.N & 7) == 0) {
        // store-only bf16 epilogue: lane -> 8 consecutive columns (one 16-byte store), 4 lanes per row, 8 rows per instruction
        const int lane = threadIdx.x & 31, r8 = lane >> 2, c8 = (lane & 3) * 8;
        const int col8 = col0 + c8;
        if (col8 >= p.N) return;
        const int64_t rows_left = p.M - row_base;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + r8;
            if (rr >= rows_left) continue;
            const float4 a = *reinterpret_cast<const float4*>(stg + rr * PITCH + c8);
            const float4 b = *reinterpret_cast<const float4*>(stg + rr * PITCH + c8 + 4);
            if (p.debug & 1) continue;
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (row_base + rr) * p.N + col8) =
                make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
        }
        return;
    }
    const int col = col0 + c4;
    if (MODE != 5) {
        if (col >= p.N) return;            // N % 4 == 0 in the specialised modes: a float4 is all-in or all-out
        float b4[4] = {0.f, 0.f, 0.f, 0.f};
        if (MODE == 2 || MODE == 3 || MODE == 6) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + col));
            b4[0] = t.x; b4[1] = t.y; b4[2] = t.z; b4[3] = t.w;
        }
        const bool drop_on = (MODE >= 2) && p.drop.on();
        const int64_t rows_left = p.M - row_base;
#pragma unroll
        for (int i = 0; i < ITER; ++i) {
            const int rr = i * RPI + sub_r;
            if (rr >= rows_left) continue;
            const float4 a4 = *reinterpret_cast<const float4*>(stg + rr * PITCH + c4);
            float f[4] = {a4.x + b4[0], a4.y + b4[1], a4.z + b4[2], a4.w + b4[3]};
            const int64_t off = (row_base + rr) * p.N + col;
            if (MODE == 3) {
                *reinterpret_cast<uint2*>(p.pre_act + off) = make_uint2(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]));
#pragma unroll
                for (int t = 0; t < 4; ++t) f[t] = gelu_fast(f[t]);
            }
            if (MODE == 4) {
                const uint32_t ux = __float_as_uint(pre[i].x), uy = __float_as_uint(pre[i].y);
                const __nv_bfloat162 h0 = *reinterpret_cast<const __nv_bfloat162*>(&ux), h1 = *reinterpret_cast<const __nv_bfloat162*>(&uy);
                f[0] *= gelu_fast_grad(__low2float(h0)); f[1] *= gelu_fast_grad(__high2float(h0));
                f[2] *= gelu_fast_grad(__low2float(h1)); f[3] *= gelu_fast_grad(__high2float(h1));
            }
            if (drop_on) {
                float d4[4];
                drop_factor4(p.drop, (uint64_t)off >> 2, d4);
#pragma unroll
                for (int t = 0; t < 4; ++t) f[t] *= d4[t];
            }
            if (MODE == 4) { csum[0] += f[0]; csum[1] += f[1]; csum[2] += f[2]; csum[3] += f[3]; }
            if (MODE == 2 || MODE == 6) { f[0] += pre[i].x; f[1] += pre[i].y; f[2] += pre[i].z; f[3] += pre[i].w; }
            if (MODE == 6) *reinterpret_cast<float4*>(rowbuf + rr * rowbuf_pitch + col) = make_float4(f[0], f[1], f[2], f[3]);
            if (p.debug & 1) continue;
            if (MODE == 1 || MODE == 2 || MODE == 6) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off) = make_float4(f[0], f[1], f[2], f[3]);
            else *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off) = make_uint2(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]));
        }
        return;
    }
    // ---- general path ----
    const bool n4 = (p.N & 3) == 0;
    const bool cfull = n4 && (col + 4 <= p.N);
    float bias4[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.bias) {
#pragma unroll
        for (int t = 0; t < 4; ++t) if (col + t < p.N) bias4[t] = __ldg(p.bias + col + t);
    }
#pragma unroll 2
    for (int i = 0; i < ITER; ++i) {
        const int rr = i * RPI + sub_r;
        const int64_t row = row_base + rr;
        if (row >= p.M || col >= p.N) continue;
        const float4 a4 = *reinterpret_cast<const float4*>(stg + rr * PITCH + c4);
        float f[4] = {a4.x + bias4[0], a4.y + bias4[1], a4.z + bias4[2], a4.w + bias4[3]};
        const int64_t off = row * p.N + col;
        if (p.pre_act) {
            if (cfull) *reinterpret_cast<uint2*>(p.pre_act + off) = make_uint2(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]));
            else for (int t = 0; t < 4; ++t) if (col + t < p.N) p.pre_act[off + t] = __float2bfloat16(f[t]);
        }
        if (p.act == 1) {
#pragma unroll
            for (int t = 0; t < 4; ++t) f[t] = gelu_erf(f[t]);
        } else if (p.act == 2) {
            for (int t = 0; t < 4; ++t) if (col + t < p.N) f[t] *= gelu_erf_grad(__bfloat162float(p.aux[off + t]));
        }
        if (p.drop.on()) {
            if (n4) {
                float d4[4];
                drop_factor4(p.drop, (uint64_t)off >> 2, d4);
#pragma unroll
                for (int t = 0; t < 4; ++t) f[t] *= d4[t];
            } else {
                for (int t = 0; t < 4; ++t) if (col + t < p.N) f[t] *= drop_factor(p.drop, (uint64_t)(off + t));
            }
        }
        if (p.residual) for (int t = 0; t < 4; ++t) if (col + t < p.N) f[t] += p.residual[off + t];
        if (p.debug & 1) continue;
        if (p.out_fp32) {
            float* o = reinterpret_cast<float*>(p.out) + off;
            if (cfull) *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
            else for (int t = 0; t < 4; ++t) if (col + t < p.N) o[t] = f[t];
        } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
            if (cfull) *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]));
            else for (int t = 0; t < 4; ++t) if (col + t < p.N) o[t] = __float2bfloat16(f[t]);
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(GB_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const __grid_constant__ CUtensorMap tma_c,
               const GemmTnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t b_bytes = (uint32_t)p.block_n * GB_K * 2;
    const uint32_t stage_bytes = GB_A_BYTES + b_bytes;
    // MODE 7 (bf16 store-only output through TMA): [stages][A|B] | 2 x [block_n/64][128 rows][128 B] staging | barriers
    const uint32_t stg_bytes = (uint32_t)(p.block_n / 64) * GB_A_BYTES;
    uint8_t* stg_base = smem + (size_t)p.stages * stage_bytes;
    GemmBars* bars = reinterpret_cast<GemmBars*>(smem + (size_t)p.stages * stage_bytes + (MODE == 7 ? 2 * (size_t)stg_bytes : 0));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t num_tiles = p.tiles_m * p.tiles_n;

    if (warp == 0 && elect_one()) { prefetch_tmap(&tma_a); prefetch_tmap(&tma_b); if (MODE == 7) prefetch_tmap(&tma_c); }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < GB_STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars->tmem_full[s], 1); mbar_init(&bars->tmem_empty[s], GB_EPI_WARPS);
            mbar_init(&bars->stg_full[s], GB_EPI_WARPS); mbar_init(&bars->stg_free[s], 1);
        }
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
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
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
                    if (!(p.debug & 4))
                    for (int k = 0; k < ksteps; ++k)   // +32 B per UMMA_K step inside the 128 B swizzle row
                        umma_bf16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), p.idesc, (kb | k) != 0);
                    umma_commit(&bars->empty[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&bars->tmem_full[acc]);
            }
        }
    } else if (MODE == 7 && warp == 3) {
        // ===== MODE 7: TMA store of the staged bf16 output tiles (one 128 x 64 box per 64 columns) =====
        if (elect_one()) {
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                const int n_blk = (int)(tile % p.tiles_n);
                const int64_t m_blk = tile / p.tiles_n;
                mbar_wait(&bars->stg_full[buf], (uint32_t)(it >> 1) & 1);
                const uint8_t* src = stg_base + (size_t)buf * stg_bytes;
                for (int t = 0; t < p.block_n / 64; ++t) {
                    const int col = n_blk * p.block_n + t * 64;
                    if (col < p.N) tma_store_2d(&tma_c, src + (size_t)t * GB_A_BYTES, col, (int)(m_blk * GB_M));
                }
                tma_store_commit();
                if (it > 0) {   // the PREVIOUS tile's stores have read their staging buffer: hand it back (this tile's may still be in flight)
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    mbar_arrive(&bars->stg_free[buf ^ 1]);
                }
            }
            tma_store_wait_read();
        }
    } else if (MODE == 7 && warp >= 4) {
        // ===== MODE 7 epilogue: TMEM -> registers -> bf16 -> swizzled staging tile (thread = output row) =====
        const int q = (warp - 4) & 3, sub = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const uint32_t swz = (uint32_t)(row & 7);
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1; const uint32_t acc_phase = (it >> 1) & 1;
            uint8_t* dst = stg_base + (size_t)acc * stg_bytes + (size_t)row * 128;
            mbar_wait(&bars->tmem_full[acc], acc_phase);
            if (it >= 2) mbar_wait(&bars->stg_free[acc], (uint32_t)((it - 2) >> 1) & 1);   // stores of tile it - 2 have read this buffer
            tc_fence_after();
            for (int c = sub; c < p.block_n / 32; c += GB_EPI_WARPS / 4) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.block_n + c * 32), v);
                tmem_ld_wait();
                uint8_t* trow = dst + (size_t)(c >> 1) * GB_A_BYTES;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<uint4*>(trow + ((((uint32_t)((c & 1) * 4 + i)) ^ swz) << 4)) =
                        make_uint4(pack_bf16(__uint_as_float(v[i * 8]), __uint_as_float(v[i * 8 + 1])), pack_bf16(__uint_as_float(v[i * 8 + 2]), __uint_as_float(v[i * 8 + 3])),
                                   pack_bf16(__uint_as_float(v[i * 8 + 4]), __uint_as_float(v[i * 8 + 5])), pack_bf16(__uint_as_float(v[i * 8 + 6]), __uint_as_float(v[i * 8 + 7])));
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&bars->tmem_empty[acc]); mbar_arrive(&bars->stg_full[acc]); }
        }
    } else if (MODE == 8 && warp >= 4) {
        // ===== MODE 8 epilogue: the accumulator rows are dy of a LayerNorm -> LN backward in place of a dy round trip =====
        // phase A: thread = row copies its TMEM row into rowbuf; phase B: warp `sub` of quarter q owns rows 8*sub..8*sub+7,
        // lane -> 4 columns (same arithmetic as ln_bwd4_kernel), x / skip-gradient rows prefetched before the accumulator wait
        const int q = (warp - 4) & 3, sub = (warp - 4) >> 2;
        const int rb_pitch = p.block_n + 4;
        float* rowbuf = reinterpret_cast<float*>(smem + (size_t)p.stages * stage_bytes + 256) + (size_t)q * 32 * rb_pitch;
        const int c = lane * 4;
        const bool act = c < p.N;
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 w4 = act ? *reinterpret_cast<const float4*>(p.ln_w + c) : z4;
        const float ws[4] = {w4.x, w4.y, w4.z, w4.w};
        const float inv_n = 1.f / (float)p.N;
        float aw[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f}, ac[4] = {0.f, 0.f, 0.f, 0.f};
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1; const uint32_t acc_phase = (it >> 1) & 1;
            const int64_t row_base = tile * GB_M + q * 32;      // tiles_n == 1
            const int64_t r0 = row_base + sub * 8;
            float4 xv[8], av[8];
            float mean[8], rstd[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const bool ok = r0 + k < p.M;
                mean[k] = ok ? p.lnb_stats[2 * (r0 + k)] : 0.f; rstd[k] = ok ? p.lnb_stats[2 * (r0 + k) + 1] : 0.f;
                xv[k] = (ok && act) ? *reinterpret_cast<const float4*>(p.lnb_x + (r0 + k) * p.N + c) : z4;
            }
            mbar_wait(&bars->tmem_full[acc], acc_phase);
            tc_fence_after();
            for (int ch = sub; ch < p.block_n / 32; ch += GB_EPI_WARPS / 4) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.block_n + ch * 32), v);
                tmem_ld_wait();
                float* dst = rowbuf + lane * rb_pitch + ch * 32;
#pragma unroll
                for (int g8 = 0; g8 < 8; ++g8)
                    *reinterpret_cast<float4*>(dst + g8 * 4) = make_float4(__uint_as_float(v[g8 * 4]), __uint_as_float(v[g8 * 4 + 1]),
                                                                            __uint_as_float(v[g8 * 4 + 2]), __uint_as_float(v[g8 * 4 + 3]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
#pragma unroll
            for (int k = 0; k < 8; ++k)   // skip-connection gradient rows (issued here: the TMEM registers are dead, x rows already landed)
                av[k] = (r0 + k < p.M && act && p.residual) ? *reinterpret_cast<const float4*>(p.residual + (r0 + k) * p.N + c) : z4;
            asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int64_t r = r0 + k;
                if (r >= p.M) break;
                const float4 d4 = act ? *reinterpret_cast<const float4*>(rowbuf + (sub * 8 + k) * rb_pitch + c) : z4;
                const float xs[4] = {xv[k].x, xv[k].y, xv[k].z, xv[k].w}, ds[4] = {d4.x, d4.y, d4.z, d4.w};
                const float as[4] = {av[k].x, av[k].y, av[k].z, av[k].w};
                float xh[4], g[4], c1 = 0.f, c2 = 0.f;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    xh[t] = act ? (xs[t] - mean[k]) * rstd[k] : 0.f;
                    aw[t] += ds[t] * xh[t]; ab[t] += ds[t];
                    g[t] = ds[t] * ws[t];
                    c1 += g[t]; c2 += g[t] * xh[t];
                }
                c1 = warp_sum(c1) * inv_n; c2 = warp_sum(c2) * inv_n;
                if (act) {
                    float o[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) o[t] = rstd[k] * (g[t] - c1 - xh[t] * c2) + as[t];
                    const int64_t off = r * p.N + c;
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off) = make_float4(o[0], o[1], o[2], o[3]);
                    if (p.lnb_cast) {
                        if (p.drop.on()) {
                            float f[4];
                            drop_factor4(p.drop, (uint64_t)off >> 2, f);
#pragma unroll
                            for (int t = 0; t < 4; ++t) o[t] *= f[t];
                        }
                        *reinterpret_cast<uint2*>(p.lnb_cast + off) = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
#pragma unroll
                        for (int t = 0; t < 4; ++t) ac[t] += o[t];
                    }
                }
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");   // rowbuf is free for the next tile
        }
        // column sums of this CTA: 16 warps -> smem -> one atomic per column
        asm volatile("bar.sync 5, 512;" ::: "memory");
        float* red = reinterpret_cast<float*>(smem + (size_t)p.stages * stage_bytes + 256);   // [3][16][128], aliases rowbuf
        const int ew = warp - 4;
        if (act) {
#pragma unroll
            for (int t = 0; t < 4; ++t) { red[(0 * 16 + ew) * 128 + c + t] = aw[t]; red[(1 * 16 + ew) * 128 + c + t] = ab[t]; red[(2 * 16 + ew) * 128 + c + t] = ac[t]; }
        }
        asm volatile("bar.sync 5, 512;" ::: "memory");
        for (int f = threadIdx.x - 128; f < p.N; f += GB_EPI_WARPS * 32) {
            float sw = 0.f, sb = 0.f, sc = 0.f;
            for (int k = 0; k < 16; ++k) { sw += red[(0 * 16 + k) * 128 + f]; sb += red[(1 * 16 + k) * 128 + f]; sc += red[(2 * 16 + k) * 128 + f]; }
            atomicAdd(p.lnb_dw + f, sw);
            atomicAdd(p.lnb_db + f, sb);
            if (p.lnb_colsum) atomicAdd(p.lnb_colsum + f, sc);
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> smem transpose -> coalesced global =====
        // 16 warps: warp (4 + q + 4*sub) owns TMEM lane quarter q and the 16-column chunks c == sub (mod 4).
        // A chunk (32 rows x 16 fp32) is written row-per-thread into a padded smem tile and read back with 4 lanes per
        // row (float4 each), so every global access of the epilogue (residual / aux loads, stores) covers whole 64-byte
        // (fp32) or 32-byte (bf16) row segments instead of 32 scattered 16-byte pieces.
        constexpr int CH = gb_chunk(MODE), PITCH = CH + 4, LPR = CH / 4;   // lanes per row in the read-back phase
        const int q = (warp - 4) & 3, sub = (warp - 4) >> 2;
        float* stg = reinterpret_cast<float*>(smem + (size_t)p.stages * stage_bytes + 256) + (size_t)(warp - 4) * (32 * PITCH);
        const int sub_r = lane / LPR, c4 = (lane % LPR) * 4;
        // MODE 6: finished fp32 rows of this TMEM lane quarter, [32 rows][block_n + 4]
        const int rb_pitch = p.block_n + 4;
        float* rowbuf = reinterpret_cast<float*>(smem + (size_t)p.stages * stage_bytes + 256 + gb_epi_smem(MODE)) + (size_t)q * 32 * rb_pitch;
        float4 lnw4 = make_float4(0.f, 0.f, 0.f, 0.f), lnb4 = lnw4;
        if (MODE == 6 && lane * 4 < p.N) { lnw4 = *reinterpret_cast<const float4*>(p.ln_w + lane * 4); lnb4 = *reinterpret_cast<const float4*>(p.ln_b + lane * 4); }
        float csum[4] = {0.f, 0.f, 0.f, 0.f};   // MODE 4: this lane's 4 output columns (fixed when block_n <= 64: one chunk per warp)
        int it = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1; const uint32_t acc_phase = (it >> 1) & 1;
            const int n_blk = (int)(tile % p.tiles_n);
            const int64_t m_blk = tile / p.tiles_n;
            bool first_chunk = true;     // the accumulator is waited for after the first chunk's operand prefetch is in flight
            const int64_t row_base = m_blk * GB_M + q * 32;
            for (int c = sub; c < ((p.debug & 2) ? 0 : p.block_n / CH); c += GB_EPI_WARPS / 4) {
                const int col0 = n_blk * p.block_n + c * CH;
                float4 pre[CH / 4];
                epilogue_prefetch<MODE>(p, row_base, col0, sub_r, c4, pre);
                if (first_chunk) { mbar_wait(&bars->tmem_full[acc], acc_phase); tc_fence_after(); first_chunk = false; }
                uint32_t v[CH];
                if (CH == 16) tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.block_n + c * CH), reinterpret_cast<uint32_t(&)[16]>(v));
                else tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.block_n + c * CH), reinterpret_cast<uint32_t(&)[32]>(v));
                tmem_ld_wait();
                if (row_base >= p.M || col0 >= p.N) continue;      // warp-uniform
                __syncwarp();
#pragma unroll
                for (int g8 = 0; g8 < CH / 4; ++g8)
                    *reinterpret_cast<float4*>(stg + lane * PITCH + g8 * 4) =
                        make_float4(__uint_as_float(v[g8 * 4]), __uint_as_float(v[g8 * 4 + 1]), __uint_as_float(v[g8 * 4 + 2]), __uint_as_float(v[g8 * 4 + 3]));
                __syncwarp();
                epilogue_rows<MODE>(p, stg, row_base, col0, sub_r, c4, pre, rowbuf, rb_pitch, csum);
            }
            if (first_chunk) { mbar_wait(&bars->tmem_full[acc], acc_phase); tc_fence_after(); }   // warp had no chunk in this tile
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
            if (MODE == 6) {
                // the 4 warps of this quarter have written its 32 finished rows: warp `sub` normalises rows 8*sub .. 8*sub+7
                // (lane -> 4 columns, two-pass variance like nn.LayerNorm), bf16 output + (mean, rstd) for the backward
                asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
                const int c = lane * 4;
                const bool act = c < p.N;
#pragma unroll 4
                for (int k = 0; k < 8; ++k) {
                    const int rr = sub * 8 + k;
                    const int64_t row = row_base + rr;
                    if (row >= p.M) break;
                    const float4 v = act ? *reinterpret_cast<const float4*>(rowbuf + rr * rb_pitch + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float mean = warp_sum(v.x + v.y + v.z + v.w) / p.N;
                    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
                    const float rstd = rsqrtf(warp_sum(act ? d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3 : 0.f) / p.N + 1e-5f);
                    if (lane == 0) { p.ln_stats[2 * row] = mean; p.ln_stats[2 * row + 1] = rstd; }
                    if (act)
                        *reinterpret_cast<uint2*>(p.ln_out + row * p.N + c) =
                            make_uint2(pack_bf16(d0 * rstd * lnw4.x + lnb4.x, d1 * rstd * lnw4.y + lnb4.y),
                                       pack_bf16(d2 * rstd * lnw4.z + lnb4.z, d3 * rstd * lnw4.w + lnb4.w));
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");   // rowbuf is free for the next tile
            }
        }
        if (MODE == 4 && p.colsum) {
            // lanes that share c4 (they differ in the row they served) -> one value per column group, then one atomic per column
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float v = csum[t];
                for (int o = LPR; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                const int col = sub * CH + c4 + t;
                if (lane < LPR && col < p.N) atomicAdd(p.colsum + col, v);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, p.tmem_cols); }
}

// ---------------------------------------------------------------------------------------------------------
// weight gradient: dW[N,K] += dY[M,N]^T X[M,K].  UMMA view: D[128 rows of N][BN cols of K] += A^T B with the
// reduction dim = tokens; both operand tiles are [64 tokens][64-element chunks] in smem -> MN-major descriptors
// (LBO = 16 KB between 64-element chunks, SBO = 1 KB between 8-token groups).  grid = (output tiles, splits).
// ---------------------------------------------------------------------------------------------------------
constexpr int WG_TOK = 128;                        // tokens per stage
constexpr int WG_STAGES = 3;
constexpr uint32_t WG_CHUNK = WG_TOK * 128;        // bytes of one [128 tokens][64 elements] chunk = 16 KB

struct WgradParams {
    int64_t M; int N, K;
    int block_n, n_chunks_b, tiles_n_out, tiles_k_out;     // block_n = columns of K per tile (<= 256)
    int64_t tok_blocks_per_split;
    uint32_t idesc, tmem_cols;
    float* dW;
};

constexpr int WG_THREADS = 256;   // TMA, MMA, TMEM-alloc, spare + 4 epilogue warps
__global__ void __launch_bounds__(WG_THREADS, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tma_dy, const __grid_constant__ CUtensorMap tma_x, const WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = 2 * WG_CHUNK, b_bytes = (uint32_t)p.n_chunks_b * WG_CHUNK;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    GemmBars* bars = reinterpret_cast<GemmBars*>(smem + (size_t)WG_STAGES * stage_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_n = blockIdx.x / p.tiles_k_out, tile_k = blockIdx.x % p.tiles_k_out;
    const int64_t total_tb = (p.M + WG_TOK - 1) / WG_TOK;
    const int64_t tb0 = (int64_t)blockIdx.y * p.tok_blocks_per_split;
    const int64_t tb1 = tb0 + p.tok_blocks_per_split < total_tb ? tb0 + p.tok_blocks_per_split : total_tb;

    if (warp == 0 && elect_one()) { prefetch_tmap(&tma_dy); prefetch_tmap(&tma_x); }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
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
                    if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
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
                    if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
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

int make_tmap_bf16(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    return make_tmap(m, base, rows, cols, ld, box_rows);
}

// 4-D bf16 tensor view (dims[0] = contiguous columns; strides in ELEMENTS for dims 1..3), box {64, box1, box2, 1}, SWIZZLE_128B,
// OOB -> zeros on load / clipped on store.  Used by the attention kernels to fetch the strided rows of a slot group as one box.
int make_tmap_bf16_4d(CUtensorMap* m, const void* base, const int64_t dims[4], const int64_t strides[3], int box1, int box2) {
    EncodeTiledFn enc = get_encode();
    MSST_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    MSST_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA operand must be 16-byte aligned");
    cuuint64_t d[4], sb[3];
    for (int i = 0; i < 4; ++i) d[i] = (cuuint64_t)dims[i];
    for (int i = 0; i < 3; ++i) { sb[i] = (cuuint64_t)strides[i] * 2; MSST_REQUIRE(sb[i] % 16 == 0, "TMA strides must be multiples of 16 bytes"); }
    cuuint32_t box[4] = {64u, (cuuint32_t)box1, (cuuint32_t)box2, 1u};
    cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), d, sb, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSST_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (4-D) failed with code %d (dims %lld %lld %lld %lld, box 64 %d %d 1)", (int)r,
                 (long long)dims[0], (long long)dims[1], (long long)dims[2], (long long)dims[3], box1, box2);
    return MSST_OK;
}

// generic bf16 view of rank 2..4 (dims[0] = contiguous columns; strides in ELEMENTS for dims 1..rank-1), arbitrary box,
// swizzle_bytes 64 or 128 (the box's inner extent must equal the swizzle span), OOB -> zeros on load
int make_tmap_nd(CUtensorMap* m, const void* base, int elem_bytes, int rank, const int64_t* dims, const int64_t* strides, const int* box, int swizzle_bytes) {
    EncodeTiledFn enc = get_encode();
    MSST_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
    MSST_REQUIRE(rank >= 2 && rank <= 4 && (elem_bytes == 2 || elem_bytes == 4) && (swizzle_bytes == 0 || swizzle_bytes == 64 || swizzle_bytes == 128) &&
                 (swizzle_bytes == 0 ? (box[0] * elem_bytes) % 16 == 0 : box[0] * elem_bytes == swizzle_bytes), "make_tmap_nd: bad rank / element / swizzle / box");
    MSST_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA operand must be 16-byte aligned");
    cuuint64_t d[4], sb[3];
    cuuint32_t bx[4], estr[4] = {1u, 1u, 1u, 1u};
    for (int i = 0; i < rank; ++i) { d[i] = (cuuint64_t)dims[i]; bx[i] = (cuuint32_t)box[i]; }
    for (int i = 0; i + 1 < rank; ++i) { sb[i] = (cuuint64_t)strides[i] * elem_bytes; MSST_REQUIRE(sb[i] % 16 == 0, "TMA strides must be multiples of 16 bytes"); }
    CUresult r = enc(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), d, sb, bx, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MSST_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (rank %d, swizzle %d) failed with code %d", rank, swizzle_bytes, (int)r);
    return MSST_OK;
}
int make_tmap_bf16_nd(CUtensorMap* m, const void* base, int rank, const int64_t* dims, const int64_t* strides, const int* box, int swizzle_bytes) {
    return make_tmap_nd(m, base, 2, rank, dims, strides, box, swizzle_bytes);
}

// MODE 6 / 8: finished fp32 rows of the four TMEM lane quarters; MODE 8 re-uses the area for its final [3][16][128] column reduction
static size_t gb_rowbuf_smem(int mode, int block_n) {
    if (mode != 6 && mode != 8) return 0;
    const size_t rb = (size_t)4 * 32 * (block_n + 4) * 4;
    return (mode == 8 && rb < (size_t)3 * 16 * 128 * 4) ? (size_t)3 * 16 * 128 * 4 : rb;
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
    p.colsum = nullptr;
    p.ln_w = a.ln_w; p.ln_b = a.ln_b; p.ln_out = a.ln_out; p.ln_stats = a.ln_stats;
    p.lnb_x = a.lnb_x; p.lnb_stats = a.lnb_stats; p.lnb_dw = a.lnb_dw; p.lnb_db = a.lnb_db; p.lnb_cast = a.lnb_cast; p.lnb_colsum = a.lnb_colsum;
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("MSST_GEMM_DEBUG"); dbg = e ? atoi(e) : 0; } p.debug = dbg; }
    CUtensorMap ta, tb;
    if (int rc = make_tmap(&ta, a.A, a.M, a.K, a.K, GB_M)) return rc;
    if (int rc = make_tmap(&tb, a.B, a.N, a.K, a.K, p.block_n)) return rc;
    
    int mode = 5;
    const bool vec = (a.N % 4 == 0);
    if (vec && !a.pre_act && !a.bias && !a.residual && a.act == 0 && !a.drop.on()) mode = a.out_fp32 ? 1 : 0;
    else if (vec && a.out_fp32 && a.bias && a.residual && a.act == 0 && !a.pre_act) mode = 2;
    if (a.lnb_x) {
        MSST_REQUIRE(a.out_fp32 && !a.bias && !a.pre_act && a.act == 0 && !a.ln_out && p.tiles_n == 1 && a.N <= 128 && a.N % 4 == 0 && a.ln_w &&
                     a.lnb_stats && a.lnb_dw && a.lnb_db, "bf16 GEMM: the fused LayerNorm-backward epilogue needs a plain fp32-out GEMM with N <= 128 (N=%d)", a.N);
        mode = 8;
    }
    if (a.ln_out) {
        MSST_REQUIRE(mode == 2 && p.tiles_n == 1 && a.N <= 128 && a.ln_w && a.ln_b && a.ln_stats,
                     "bf16 GEMM: the fused LayerNorm epilogue needs fp32 out + bias + residual and N <= 128 (N=%d)", a.N);
        mode = 6;
    }
    else if (vec && !a.out_fp32 && a.bias && a.act == 1 && a.pre_act && !a.residual) mode = 3;
    else if (vec && !a.out_fp32 && !a.bias && a.act == 2 && a.aux && !a.residual && !a.pre_act) mode = 4;
    if (a.colsum) {
        MSST_REQUIRE(mode == 4 && p.tiles_n == 1 && p.block_n <= 64, "bf16 GEMM: the fused column sum needs the GELU' data-gradient epilogue with N <= 64 (N=%d)", a.N);
        p.colsum = a.colsum;
    }
    // store-only bf16 output (QKV projection, dO data gradient): the epilogue packs rows into swizzled staging tiles and a
    // dedicated warp streams them out with TMA stores (MSST_GEMM_TMA_STORE=0 selects the st.global epilogue, MODE 0)
    CUtensorMap tc = ta;
    {   // as many operand stages as fit (long-K data gradients keep more loads in flight), at least 3 worth of the old layout
        const size_t stage_b = GB_A_BYTES + (size_t)p.block_n * GB_K * 2;
        const size_t fixed = 256 + gb_epi_smem(mode) + 1024 + gb_rowbuf_smem(mode, p.block_n);
        int st_fit = (int)((227 * 1024 - fixed) / stage_b);
        if (st_fit > GB_STAGES) st_fit = GB_STAGES;
        if (st_fit > p.num_kb + 1) st_fit = p.num_kb + 1 > 2 ? p.num_kb + 1 : 2;
        MSST_REQUIRE(st_fit >= 2, "bf16 GEMM: shared memory budget");
        p.stages = st_fit;
    }
    if (mode == 0 && a.N % 64 == 0 && p.block_n == 256 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0) {   // wide outputs only (measured: no gain at N = 64)
        static int use_tma_store = -1;
        if (use_tma_store < 0) { const char* e = getenv("MSST_GEMM_TMA_STORE"); use_tma_store = e ? atoi(e) : 1; }
        if (use_tma_store) {
            if (int rc = make_tmap(&tc, a.out, a.M, a.N, a.N, GB_M)) return rc;
            mode = 7;
            p.stages = 2;
        }
    }
    const size_t smem = mode == 7 ? (size_t)p.stages * (GB_A_BYTES + (size_t)p.block_n * GB_K * 2) + 2 * (size_t)(p.block_n / 64) * GB_A_BYTES + 256 + 1024
                                  : (size_t)p.stages * (GB_A_BYTES + (size_t)p.block_n * GB_K * 2) + 256 + gb_epi_smem(mode) + 1024 +
                                        gb_rowbuf_smem(mode, p.block_n);
    const int64_t tiles = p.tiles_m * p.tiles_n;
    const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        MSST_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MSST_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MSST_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MSST_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MSST_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MSST_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MSST_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MSST_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MSST_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    switch (mode) {
        case 0: gemm_tn_kernel<0><<<grid, GB_THREADS, smem, st>>>(ta, tb, tc, p); break;
        case 1: gemm_tn_kernel<1><<<grid, GB_THREADS, smem, st>>>(ta, tb, tc, p); break;
        case 2: gemm_tn_kernel<2><<<grid, GB_THREADS, smem, st>>>(ta, tb, tc, p); break;
        case 3: gemm_tn_kernel<3><<<grid, GB_THREADS, smem, st>>>(ta, tb, tc, p); break;
        case 4: gemm_tn_kernel<4><<<grid, GB_THREADS, smem, st>>>(ta, tb, tc, p); break;
        case 6: gemm_tn_kernel<6><<<grid, GB_THREADS, smem, st>>>(ta, tb, tc, p); break;
        case 7: gemm_tn_kernel<7><<<grid, GB_THREADS, smem, st>>>(ta, tb, tc, p); break;
        case 8: gemm_tn_kernel<8><<<grid, GB_THREADS, smem, st>>>(ta, tb, tc, p); break;
        default: gemm_tn_kernel<5><<<grid, GB_THREADS, smem, st>>>(ta, tb, tc, p); break;
    }
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}

int gemm_wgrad_bf16(const __nv_bfloat16* dy, const __nv_bfloat16* x, float* dW, int64_t M, int N, int K, cudaStream_t st) {
    if (M == 0) return MSST_OK;
    MSST_REQUIRE(N % 8 == 0 && K % 8 == 0, "bf16 wgrad: N=%d and K=%d must be multiples of 8", N, K);
    WgradParams p{};
    p.M = M; p.N = N; p.K = K; p.dW = dW;
    const int k32 = (K + 31) / 32 * 32;
    p.block_n = k32 < 128 ? k32 : 128;     // <= 2 chunks of B per stage: 3 stages x 64 KB of smem
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
    const size_t smem = (size_t)WG_STAGES * (2 * WG_CHUNK + (size_t)p.n_chunks_b * WG_CHUNK) + sizeof(GemmBars) + 1024;
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        MSST_CUDA(cudaFuncSetAttribute(gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    gemm_wgrad_kernel<<<dim3(out_tiles, (unsigned)splits), WG_THREADS, smem, st>>>(ta, tb, p);
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}

}  // namespace msst
