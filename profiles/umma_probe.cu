// Stand-alone probe of the tcgen05 features the fused attention-block kernels (csrc/attn_block_tc.cu) rely on, each checked against
// a CPU result.  Build + run (B200):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_probe profiles/umma_probe.cu -lcuda && /tmp/umma_probe
//   T1  K = 96 operands as three 32-column chunks, SWIZZLE_64B, written by TMA, read by UMMA (M = 128, N = 64)
//   T2  two M = 64 MMAs interleaved in the same TMEM columns (second atom at lane offset 16): lane <-> row map
//   T3  M = 64 with MN-major (transposed) A and MN-major B  (dV = P^T dO per 64-slot block)
//   T4  A operand from TMEM (bf16 pairs written with tcgen05.st), M = 128
//   T5  M = 64, K-major A, MN-major B (O = P V per block) -- the forward's second contraction
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include "../maskedsst_b200/csrc/ptx.cuh"

using namespace msst::ptx;
typedef __nv_bfloat16 bf16;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr uint64_t kSW64 = 4;
__device__ __host__ inline uint64_t desc_sw(uint32_t addr, uint32_t lbo, uint32_t sbo, uint64_t layout) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// manual SWIZZLE_128B tile [rows][64 bf16]: 16-byte chunk c of row r lives at chunk (c ^ (r & 7))
__device__ void fill_sw128(uint8_t* tile, const bf16* src, int rows, int ld) {
    for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(src + (size_t)r * ld + c * 8);
    }
}

struct Bars { uint64_t full, done; uint32_t tmem; };

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const bf16* Q, const bf16* K, const bf16* P,
             const bf16* dO, float* out1, float* out2, float* out3, float* out4, float* out5) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_s = smem;                 // T1: 3 chunks [128][32] SW64 = 3 x 8 KB
    uint8_t* b_s = smem + 24576;         // T1: 3 chunks [64][32] SW64 = 3 x 4 KB
    uint8_t* q_s = smem + 40960;         // [128][64] SW128 16 KB
    uint8_t* k_s = q_s + 16384;
    uint8_t* p_s = k_s + 16384;          // two blocks [64 q][64 k]
    uint8_t* do_s = p_s + 16384;
    Bars* bars = reinterpret_cast<Bars*>(do_s + 16384);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bars->full, 1); mbar_init(&bars->done, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&bars->tmem, 512);
    fill_sw128(q_s, Q, 128, 64); fill_sw128(k_s, K, 128, 64); fill_sw128(p_s, P, 128, 64); fill_sw128(do_s, dO, 128, 64);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = bars->tmem;
    const uint32_t L = threadIdx.x;                       // TMEM lane of this thread
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    uint32_t ph = 0;

    // ---- T1: SW64 chunks by TMA ----
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bars->full, 24576 + 12288);
        for (int c = 0; c < 3; ++c) {
            tma_load_2d(a_s + c * 8192, &tmA, &bars->full, c * 32, 0);
            tma_load_2d(b_s + c * 4096, &tmB, &bars->full, c * 32, 0);
        }
        mbar_wait(&bars->full, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
        for (int k = 0; k < 6; ++k) {
            const uint64_t da = desc_sw(smem_u32(a_s + (k >> 1) * 8192) + (k & 1) * 32, 16, 512, kSW64);
            const uint64_t db = desc_sw(smem_u32(b_s + (k >> 1) * 4096) + (k & 1) * 32, 16, 512, kSW64);
            umma_bf16(tm, da, db, idesc, k != 0);
        }
        umma_commit(&bars->done);
    }
    mbar_wait(&bars->done, ph); ph ^= 1;
    tc_fence_after();
    {
        uint32_t v[32];
        for (int c = 0; c < 2; ++c) {
            tmem_ld_32x32(tm + lane_addr + c * 32, v); tmem_ld_wait();
            for (int j = 0; j < 32; ++j) out1[L * 64 + c * 32 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();

    // ---- T2: S_b = Q_b K_b^T, two M = 64 atoms interleaved in columns [64, 128) ----
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(64, 64, 0, 0);
        for (int b = 0; b < 2; ++b)
            for (int k = 0; k < 4; ++k) {
                const uint64_t da = make_smem_desc(smem_u32(q_s + b * 8192), 16, 1024) + (uint64_t)(k * 2);
                const uint64_t db = make_smem_desc(smem_u32(k_s + b * 8192), 16, 1024) + (uint64_t)(k * 2);
                umma_bf16(tm + 64 + ((uint32_t)(16 * b) << 16), da, db, idesc, k != 0);
            }
        umma_commit(&bars->done);
    }
    mbar_wait(&bars->done, ph); ph ^= 1;
    tc_fence_after();
    {
        uint32_t v[32];
        for (int c = 0; c < 2; ++c) {
            tmem_ld_32x32(tm + lane_addr + 64 + c * 32, v); tmem_ld_wait();
            for (int j = 0; j < 32; ++j) out2[L * 64 + c * 32 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();

    // ---- T3: dV_b = P_b^T dO_b (A MN-major, B MN-major, M = 64) into columns [128, 192) ----
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(64, 64, 1, 1);
        for (int b = 0; b < 2; ++b)
            for (int k = 0; k < 4; ++k) {   // reduction over the 64 queries of the block, 16 rows (2 KB) per step
                const uint64_t da = make_smem_desc(smem_u32(p_s + b * 8192), 8192, 1024) + (uint64_t)(k * 128);
                const uint64_t db = make_smem_desc(smem_u32(do_s + b * 8192), 8192, 1024) + (uint64_t)(k * 128);
                umma_bf16(tm + 128 + ((uint32_t)(16 * b) << 16), da, db, idesc, k != 0);
            }
        umma_commit(&bars->done);
    }
    mbar_wait(&bars->done, ph); ph ^= 1;
    tc_fence_after();
    {
        uint32_t v[32];
        for (int c = 0; c < 2; ++c) {
            tmem_ld_32x32(tm + lane_addr + 128 + c * 32, v); tmem_ld_wait();
            for (int j = 0; j < 32; ++j) out3[L * 64 + c * 32 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();

    // ---- T4: A from TMEM: thread = row L writes Q[L][0..63] as 32 packed columns at [192, 224); D = Q K^T(all 128 keys) -> [256, 384) ----
    {
        for (int c = 0; c < 4; ++c) {
            uint32_t v[8];
            for (int j = 0; j < 8; ++j) v[j] = *reinterpret_cast<const uint32_t*>(Q + (size_t)L * 64 + c * 16 + j * 2);
            tmem_st_32x8(tm + lane_addr + 192 + c * 8, v);
        }
        tmem_st_wait();
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
        for (int k = 0; k < 4; ++k) {
            const uint64_t db = make_smem_desc(smem_u32(k_s), 16, 1024) + (uint64_t)(k * 2);
            umma_bf16_ts(tm + 256, tm + 192 + k * 8, db, idesc, k != 0);
        }
        umma_commit(&bars->done);
    }
    mbar_wait(&bars->done, ph); ph ^= 1;
    tc_fence_after();
    {
        uint32_t v[32];
        for (int c = 0; c < 4; ++c) {
            tmem_ld_32x32(tm + lane_addr + 256 + c * 32, v); tmem_ld_wait();
            for (int j = 0; j < 32; ++j) out4[L * 128 + c * 32 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();

    // ---- T5: O_b = P_b dO_b (A K-major, B MN-major, M = 64) into columns [384, 448) ----
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(64, 64, 0, 1);
        for (int b = 0; b < 2; ++b)
            for (int k = 0; k < 4; ++k) {
                const uint64_t da = make_smem_desc(smem_u32(p_s + b * 8192), 16, 1024) + (uint64_t)(k * 2);
                const uint64_t db = make_smem_desc(smem_u32(do_s + b * 8192), 8192, 1024) + (uint64_t)(k * 128);
                umma_bf16(tm + 384 + ((uint32_t)(16 * b) << 16), da, db, idesc, k != 0);
            }
        umma_commit(&bars->done);
    }
    mbar_wait(&bars->done, ph); ph ^= 1;
    tc_fence_after();
    {
        uint32_t v[32];
        for (int c = 0; c < 2; ++c) {
            tmem_ld_32x32(tm + lane_addr + 384 + c * 32, v); tmem_ld_wait();
            for (int j = 0; j < 32; ++j) out5[L * 64 + c * 32 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float bfr(float x) { return __bfloat162float(__float2bfloat16(x)); }
static int slot_of_lane(int L) { return 64 * ((L >> 4) & 1) + 16 * (L >> 5) + (L & 15); }

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    EncodeFn enc = (EncodeFn)fn;
    srand(5);
    auto rnd = [](int n) { std::vector<float> v(n); for (auto& x : v) x = bfr((float)rand() / RAND_MAX - 0.5f); return v; };
    std::vector<float> A = rnd(128 * 96), B = rnd(64 * 96), Q = rnd(128 * 64), K = rnd(128 * 64), P = rnd(128 * 64), dO = rnd(128 * 64);
    auto up = [](const std::vector<float>& v) { std::vector<bf16> h(v.size()); for (size_t i = 0; i < v.size(); ++i) h[i] = __float2bfloat16(v[i]); bf16* d; CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice)); return d; };
    bf16 *dA = up(A), *dB = up(B), *dQ = up(Q), *dK = up(K), *dP = up(P), *ddO = up(dO);
    CUtensorMap tmA, tmB;
    auto mk = [&](CUtensorMap* m, void* base, int rows, int box_rows) {
        cuuint64_t d[2] = {96, (cuuint64_t)rows}, sb[1] = {96 * 2};
        cuuint32_t box[2] = {32, (cuuint32_t)box_rows}, es[2] = {1, 1};
        CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, d, sb, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    };
    mk(&tmA, dA, 128, 128); mk(&tmB, dB, 64, 64);
    float *o1, *o2, *o3, *o4, *o5;
    CK(cudaMalloc(&o1, 128 * 64 * 4)); CK(cudaMalloc(&o2, 128 * 64 * 4)); CK(cudaMalloc(&o3, 128 * 64 * 4)); CK(cudaMalloc(&o4, 128 * 128 * 4)); CK(cudaMalloc(&o5, 128 * 64 * 4));
    const size_t smem = 40960 + 4 * 16384 + 64;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<<<1, 128, smem>>>(tmA, tmB, dQ, dK, dP, ddO, o1, o2, o3, o4, o5);
    CK(cudaDeviceSynchronize());
    std::vector<float> h1(128 * 64), h2(128 * 64), h3(128 * 64), h4(128 * 128), h5(128 * 64);
    CK(cudaMemcpy(h1.data(), o1, h1.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h2.data(), o2, h2.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h3.data(), o3, h3.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h4.data(), o4, h4.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h5.data(), o5, h5.size() * 4, cudaMemcpyDeviceToHost));
    double e1 = 0, e2 = 0, e3 = 0, e4 = 0, e5 = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) { double s = 0; for (int k = 0; k < 96; ++k) s += (double)A[m * 96 + k] * B[n * 96 + k]; e1 = fmax(e1, fabs(s - h1[m * 64 + n])); }
    for (int L = 0; L < 128; ++L) {
        const int slot = slot_of_lane(L), b = slot >> 6;
        for (int n = 0; n < 64; ++n) {
            double s = 0; for (int k = 0; k < 64; ++k) s += (double)Q[slot * 64 + k] * K[(64 * b + n) * 64 + k];
            e2 = fmax(e2, fabs(s - h2[L * 64 + n]));
            double t = 0; for (int q = 0; q < 64; ++q) t += (double)P[(64 * b + q) * 64 + (slot & 63)] * dO[(64 * b + q) * 64 + n];   // dV[key = slot][n]
            e3 = fmax(e3, fabs(t - h3[L * 64 + n]));
            double u = 0; for (int k = 0; k < 64; ++k) u += (double)P[slot * 64 + k] * dO[(64 * b + k) * 64 + n];                      // O[q = slot][n]
            e5 = fmax(e5, fabs(u - h5[L * 64 + n]));
        }
    }
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) { double s = 0; for (int k = 0; k < 64; ++k) s += (double)Q[m * 64 + k] * K[n * 64 + k]; e4 = fmax(e4, fabs(s - h4[m * 128 + n])); }
    printf("T1 SW64 TMA chunks K=96 M=128 N=64      max abs err %.3e  %s\n", e1, e1 < 1e-3 ? "OK" : "FAIL");
    printf("T2 M=64 interleaved atoms (K-major)     max abs err %.3e  %s\n", e2, e2 < 1e-3 ? "OK" : "FAIL");
    printf("T3 M=64 MN-major A, MN-major B          max abs err %.3e  %s\n", e3, e3 < 1e-3 ? "OK" : "FAIL");
    printf("T4 A from TMEM (bf16 pairs) M=128 N=128 max abs err %.3e  %s\n", e4, e4 < 1e-3 ? "OK" : "FAIL");
    printf("T5 M=64 K-major A, MN-major B           max abs err %.3e  %s\n", e5, e5 < 1e-3 ? "OK" : "FAIL");
    if (e2 >= 1e-3) {   // diagnose the lane map: which slot does each lane hold (match on column 0..3)?
        for (int L = 0; L < 128; L += 1) {
            int best = -1; double be = 1e9;
            for (int s = 0; s < 128; ++s) { double err = 0; for (int n = 0; n < 8; ++n) { double v = 0; for (int k = 0; k < 64; ++k) v += (double)Q[s * 64 + k] * K[(64 * (s >> 6) + n) * 64 + k]; err += fabs(v - h2[L * 64 + n]); } if (err < be) { be = err; best = s; } }
            printf("lane %3d -> slot %3d (err %.2e)%s", L, best, be, (L % 4 == 3) ? "\n" : "   ");
        }
    }
    return 0;
}
