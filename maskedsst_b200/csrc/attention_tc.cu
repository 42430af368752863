// Kernel (3), bf16 mode, N <= 64: attention forward on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// One CTA works on 128 slots (two 64-slot groups of attn_geom.cuh) x one head at a time and loops persistently over
// (tile, head) items.  Per item:
//   producer warps : gather the Q, K, V rows (strided row addressing of the spectral stack included) with cp.async into a
//                    SWIZZLE_128B K-major tile each, completion signalled on an mbarrier (cp.async.mbarrier.arrive)
//   MMA warp       : S[128x128] = Q K^T  -> TMEM (tcgen05.mma, one elected thread);   later  O[128x64] = P V -> TMEM
//                    (P from smem K-major, V as MN-major B operand: no transpose anywhere)
//   softmax warps  : thread = row.  tcgen05.ld of the row's own 64-key block, mask (same sequence), exp2, row sum and Philox-free
//                    pair-hash dropout entirely thread-local (no shuffles), P (bf16) -> smem; epilogue: O row / l -> global, lse
// Only the two diagonal 64x64 blocks of S are meaningful (block-diagonal mask); the off-diagonal halves of the P tile are
// zeroed once and never written.  smem/TMEM stages are double-buffered so the tensor work of item i+1 overlaps the softmax of i.
// The kernel is HBM-bound: 3*128 B in + 128 B out per (slot, head).
#include "common.cuh"
#include "kernels.h"
#include "attn_geom.cuh"
#include "ptx.cuh"

namespace msst {
using namespace ptx;
typedef __nv_bfloat16 bf16;

constexpr int TC_ROWS = 128;
constexpr int TC_THREADS = 256;            // warps 0-3 softmax/epilogue, 4-5 producers, 6 MMA, 7 TMEM alloc
constexpr int TC_PRODUCERS = 64;
constexpr uint32_t TC_TILE = TC_ROWS * 128;   // bytes of one [128 rows][64 bf16] tile (16 KB)

struct TcRowMeta {
    int64_t row[TC_ROWS];      // global row of each slot or -1
    int16_t klo[TC_ROWS];      // valid key range [klo, khi) inside the slot's own 64-key block
    int16_t khi[TC_ROWS];
};
struct alignas(8) TcBars {
    uint64_t full[2], kv_empty[2], s_full[2], p_full[2], o_full[2], tmem_free[2];
    uint32_t tmem_base;
};

// smem: [2 stages][Q,K,V][16 KB] | [2 stages][P chunk0, chunk1][16 KB] | meta | barriers
constexpr size_t kTcSmem = 2 * 3 * TC_TILE + 2 * 2 * TC_TILE + sizeof(TcRowMeta) + sizeof(TcBars) + 1024;

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// dropout pair indexing / hash identical to attention_bf16.cu (the backward kernel regenerates the same mask)
__device__ __forceinline__ uint64_t tile_pair_base_tc(const AttnGeom& g, int64_t group, int h) {
    return ((uint64_t)group * g.H + h) * (uint64_t)(TS * TS / 2);
}
__device__ __forceinline__ uint32_t tc_pair_hash(const Drop& d, uint64_t idx) {
    const uint64_t s = d.seed + (d.seed_dev ? __ldg(d.seed_dev) : 0ull);
    uint32_t x = ((uint32_t)idx + (uint32_t)(s >> 32)) * 0x9E3779B1u ^ ((uint32_t)(idx >> 32) * 0x85EBCA77u) ^ (uint32_t)s ^ (d.site * 0xC2B2AE3Du);
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
attn_fwd_tc_kernel(AttnGeom g, const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse, Drop drop, int64_t n_tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* qkv_s = smem;                                  // [2][3][16 KB]
    uint8_t* p_s = smem + 2 * 3 * TC_TILE;                  // [2][2][16 KB]
    TcRowMeta* meta = reinterpret_cast<TcRowMeta*>(p_s + 2 * 2 * TC_TILE);
    TcBars* bars = reinterpret_cast<TcBars*>(meta + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int I = g.H * 64;
    const int64_t ld = 3 * (int64_t)I;

    if (warp == 6 && elect_one()) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars->full[s], TC_PRODUCERS); mbar_init(&bars->kv_empty[s], 1); mbar_init(&bars->s_full[s], 1);
            mbar_init(&bars->p_full[s], 128); mbar_init(&bars->o_full[s], 1); mbar_init(&bars->tmem_free[s], 128);
        }
        fence_barrier_init();
    }
    if (warp == 7) tmem_alloc(&bars->tmem_base, 512);
    // zero the P tiles once: the off-diagonal halves are never written afterwards
    for (uint32_t i = threadIdx.x; i < 2 * 2 * TC_TILE / 16; i += TC_THREADS) reinterpret_cast<uint4*>(p_s)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0), idesc_o = make_idesc_bf16(128, 64, 0, 1);

    // every role walks the same item sequence: tiles blockIdx.x, +gridDim.x, ...; heads 0..H-1 inside a tile
    int tile_iter = 0;                  // items of this CTA are numbered tile_iter * H + h (stage = item & 1, phase = (item >> 1) & 1)
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_iter) {
        const int item0 = tile_iter * g.H;
        // ---- per-tile row metadata (all threads), visible to everybody after the barrier ----
        __syncthreads();                // previous tile completely finished (its epilogue read meta)
        if (threadIdx.x < TC_ROWS) {
            const int r = threadIdx.x;
            const int64_t group = tile * 2 + (r >> 6);
            int64_t seq; int pos;
            bool ok = group < g.groups && slot_to(g, group, 0, r & 63, seq, pos);
            meta->row[r] = ok ? row_of(g, seq, pos) : -1;
            const int lo = ok ? ((r & 63) / g.N) * g.N : 0;     // slots of the same sequence are contiguous inside the group
            meta->klo[r] = (int16_t)lo;
            meta->khi[r] = (int16_t)(ok ? lo + g.N : 0);
        }
        __syncthreads();

        if (warp >= 4 && warp <= 5) {
            // ===== producers: gather Q, K, V rows of head h into stage (item & 1) =====
            const int pt = threadIdx.x - 128;                      // 0..63
            for (int h = 0; h < g.H; ++h) {
                const int item = item0 + h;
                const int st = item & 1; const uint32_t ph = (item >> 1) & 1;
                mbar_wait(&bars->kv_empty[st], ph ^ 1);
                const uint32_t base = smem_u32(qkv_s + (size_t)st * 3 * TC_TILE);
#pragma unroll 4
                for (int i = pt; i < TC_ROWS * 8; i += TC_PRODUCERS) {
                    const int r = i >> 3, pc = i & 7;
                    const int64_t row = meta->row[r];
                    const bf16* src = qkv + (row < 0 ? 0 : row) * ld + h * 64 + pc * 8;
                    const uint32_t dst = base + r * 128 + ((pc ^ (r & 7)) << 4);
                    cp_async16_zfill(dst, src, row >= 0);
                    cp_async16_zfill(dst + TC_TILE, src + I, row >= 0);
                    cp_async16_zfill(dst + 2 * TC_TILE, src + 2 * I, row >= 0);
                }
                cp_async_mbar_arrive(&bars->full[st]);
            }
        } else if (warp == 6) {
            // ===== MMA issuer =====
            if (elect_one()) {
                for (int h = 0; h < g.H; ++h) {
                    const int item = item0 + h;
                    const int st = item & 1; const uint32_t ph = (item >> 1) & 1;
                    const uint32_t qb = smem_u32(qkv_s + (size_t)st * 3 * TC_TILE);
                    const uint32_t pb = smem_u32(p_s + (size_t)st * 2 * TC_TILE);
                    mbar_wait(&bars->tmem_free[st], ph ^ 1);       // S / O TMEM stage drained by the epilogue of item - 2
                    mbar_wait(&bars->full[st], ph);                // Q, K, V landed (generic-proxy writes by cp.async)
                    fence_proxy_async();
                    tc_fence_after();
                    const uint64_t dq = make_smem_desc(qb, 16, 1024), dk = make_smem_desc(qb + TC_TILE, 16, 1024);
                    for (int k = 0; k < 4; ++k)                    // S = Q K^T, K = 64 = 4 x UMMA_K
                        umma_bf16(tmem_base + st * 128, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k != 0);
                    umma_commit(&bars->s_full[st]);
                    mbar_wait(&bars->p_full[st], ph);              // P written by the softmax warps (generic proxy + their fence)
                    fence_proxy_async();
                    tc_fence_after();
                    const uint64_t dv = make_smem_desc(qb + 2 * TC_TILE, TC_TILE, 1024);
                    for (int k = 0; k < 8; ++k) {                  // O = P V, K = 128 keys = 8 x UMMA_K
                        const uint64_t dp = make_smem_desc(pb + (k >> 2) * TC_TILE, 16, 1024) + (uint64_t)((k & 3) * 2);
                        umma_bf16(tmem_base + 256 + st * 64, dp, dv + (uint64_t)(k * 128), idesc_o, k != 0);
                    }
                    umma_commit(&bars->o_full[st]);
                    umma_commit(&bars->kv_empty[st]);              // Q, K, V (and P) of this stage are free again
                }
            }
        } else if (warp < 4) {
            // ===== softmax + epilogue warps: thread r = row r of the tile =====
            const int r = threadIdx.x;
            const int blk = r >> 6;                                 // which 64-key block (== which slot group of the tile)
            const int64_t grow = meta->row[r];
            const int klo = meta->klo[r], khi = meta->khi[r];
            const int64_t group = tile * 2 + blk;
            const float sl2 = g.scale * 1.4426950408889634f;
            const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
            float l_prev = 0.f, m_prev = 0.f; int h_prev = -1, st_prev = 0; uint32_t ph_prev = 0;
            auto epilogue = [&](int st, uint32_t ph, int h, float l, float m) {
                mbar_wait(&bars->o_full[st], ph);
                tc_fence_after();
                uint32_t v[32], w[32];
                tmem_ld_32x32(tmem_base + lane_base + 256 + st * 64, v);
                tmem_ld_32x32(tmem_base + lane_base + 256 + st * 64 + 32, w);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&bars->tmem_free[st]);
                if (grow >= 0) {
                    const float inv = l > 0.f ? 1.f / l : 0.f;
                    bf16* dst = out + grow * I + h * 64;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *reinterpret_cast<uint4*>(dst + c * 8) =
                            make_uint4(pack_bf(__uint_as_float(v[c * 8]) * inv, __uint_as_float(v[c * 8 + 1]) * inv),
                                       pack_bf(__uint_as_float(v[c * 8 + 2]) * inv, __uint_as_float(v[c * 8 + 3]) * inv),
                                       pack_bf(__uint_as_float(v[c * 8 + 4]) * inv, __uint_as_float(v[c * 8 + 5]) * inv),
                                       pack_bf(__uint_as_float(v[c * 8 + 6]) * inv, __uint_as_float(v[c * 8 + 7]) * inv));
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *reinterpret_cast<uint4*>(dst + 32 + c * 8) =
                            make_uint4(pack_bf(__uint_as_float(w[c * 8]) * inv, __uint_as_float(w[c * 8 + 1]) * inv),
                                       pack_bf(__uint_as_float(w[c * 8 + 2]) * inv, __uint_as_float(w[c * 8 + 3]) * inv),
                                       pack_bf(__uint_as_float(w[c * 8 + 4]) * inv, __uint_as_float(w[c * 8 + 5]) * inv),
                                       pack_bf(__uint_as_float(w[c * 8 + 6]) * inv, __uint_as_float(w[c * 8 + 7]) * inv));
                    lse[grow * g.H + h] = (m + log2f(l)) * 0.6931471805599453f;
                }
            };
            for (int h = 0; h < g.H; ++h) {
                const int item = item0 + h;
                const int st = item & 1; const uint32_t ph = (item >> 1) & 1;
                mbar_wait(&bars->s_full[st], ph);
                tc_fence_after();
                uint32_t a[32], b[32];
                const uint32_t scol = tmem_base + lane_base + st * 128 + blk * 64;
                tmem_ld_32x32(scol, a);
                tmem_ld_32x32(scol + 32, b);
                tmem_ld_wait();
                float s[64];
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                    const float v = __uint_as_float(j < 32 ? a[j] : b[j - 32]);
                    s[j] = (j >= klo && j < khi) ? v * sl2 : -INFINITY;
                    mx = fmaxf(mx, s[j]);
                }
                const float sub = mx == -INFINITY ? 0.f : mx;
                float l = 0.f;
#pragma unroll
                for (int j = 0; j < 64; ++j) { s[j] = exp2f(s[j] - sub); l += s[j]; }
                if (drop.on()) {
                    const uint64_t base = tile_pair_base_tc(g, group, h) + (uint64_t)((r & 63) * 32);
                    const uint32_t t16 = drop.thresh >> 16;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const uint32_t hsh = tc_pair_hash(drop, base + j);
                        s[2 * j] *= (hsh & 0xFFFFu) >= t16 ? drop.scale : 0.f;
                        s[2 * j + 1] *= (hsh >> 16) >= t16 ? drop.scale : 0.f;
                    }
                }
                // P row (bf16) -> smem P tile of this stage: chunk = own key block, SWIZZLE_128B K-major
                {
                    uint8_t* prow = p_s + (size_t)st * 2 * TC_TILE + (size_t)blk * TC_TILE + r * 128;
#pragma unroll
                    for (int pc = 0; pc < 8; ++pc)
                        *reinterpret_cast<uint4*>(prow + ((pc ^ (r & 7)) << 4)) =
                            make_uint4(pack_bf(s[pc * 8], s[pc * 8 + 1]), pack_bf(s[pc * 8 + 2], s[pc * 8 + 3]),
                                       pack_bf(s[pc * 8 + 4], s[pc * 8 + 5]), pack_bf(s[pc * 8 + 6], s[pc * 8 + 7]));
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(&bars->p_full[st]);
                // epilogue of the previous item overlaps this item's O = P V
                if (h_prev >= 0) epilogue(st_prev, ph_prev, h_prev, l_prev, m_prev);
                l_prev = l; m_prev = sub; h_prev = h; st_prev = st; ph_prev = ph;
            }
            epilogue(st_prev, ph_prev, h_prev, l_prev, m_prev);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 7) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int attention_fwd_tc(const AttnGeom& g, const bf16* qkv, bf16* out, float* lse, Drop drop, cudaStream_t st) {
    const int64_t n_tiles = (g.groups + 1) / 2;
    static PerDeviceOnce once;
    if (once.first()) MSST_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
    const int grid = (int)(n_tiles < kNumSMs ? n_tiles : kNumSMs);
    attn_fwd_tc_kernel<<<grid, TC_THREADS, kTcSmem, st>>>(g, qkv, out, lse, drop, n_tiles);
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}

}  // namespace msst
