// SimMIM block / tube mask generation ON THE DEVICE, one launch per step.
// Replaces MaskGenerator.__call__ / get_batch / get_batch_tube_masked / bool_mask_to_indices (reference
// src/vit_simmim_original.py:343-416), whose python loops (B*C numpy permutations) and nonzero() sync dominate a B200 step.
//   draw  : a uniformly random subset of `mask_count` of the rand_size^2 cells (the reference: permutation(...)[:mask_count]) --
//           here the mask_count smallest of i.i.d. counter-hash keys (seed, draw, cell), ties broken by the cell index
//   mask  : cells upsampled by `scale` to the side x side token grid; one draw per sample shared by all C spectral blocks (tube) or one
//           per (sample, block)
//   idx   : the reference's slicing (quirk C3): the ascending set positions of ALL samples concatenated, cut into consecutive runs of
//           num_masked -- flat position j = sample * per_row + rank lands in idx[j / nm][j % nm] (per_row = set positions per sample)
// The host generator (numpy global RNG, bit-compatible with the reference) remains the default for parity runs.
#include "common.cuh"

namespace msst {

constexpr int MG_THREADS = 256;

__global__ void __launch_bounds__(MG_THREADS)
mask_draw_kernel(msst_maskgen_dims d, uint8_t* __restrict__ mask, int64_t* __restrict__ idx) {
    extern __shared__ uint32_t sm[];
    const int cells = d.rand_size * d.rand_size, side = d.rand_size * d.scale, S = side * side, T = d.C * S;
    const int n_draws = d.tube ? 1 : d.C;
    uint32_t* keys = sm;                                   // [n_draws][cells]
    uint8_t* chosen = reinterpret_cast<uint8_t*>(sm + n_draws * cells);   // [n_draws][cells]
    __shared__ int warp_tot[MG_THREADS / 32];
    __shared__ int running;
    const int b = blockIdx.x;
    const uint64_t seed = d.seed + (d.seed_dev ? __ldg(d.seed_dev) : 0ull);
    for (int i = threadIdx.x; i < n_draws * cells; i += MG_THREADS) {
        const uint64_t draw = (uint64_t)b * n_draws + i / cells;
        uint32_t a, c;
        drop_bits64(seed, 0x4D41534Bu /* "MASK" */, draw * (uint64_t)cells + (uint64_t)(i % cells), a, c);
        keys[i] = a ^ (c >> 7);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_draws * cells; i += MG_THREADS) {
        const int dr = i / cells, me = i % cells;
        const uint32_t k = keys[i];
        int rank = 0;
        for (int j = 0; j < cells; ++j) { const uint32_t o = keys[dr * cells + j]; rank += (o < k) || (o == k && j < me); }
        chosen[i] = rank < d.mask_count;
    }
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    const int64_t per_row = (int64_t)d.mask_count * d.scale * d.scale * d.C;
    const int64_t total = (int64_t)d.B * d.nm;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int p0 = 0; p0 < T; p0 += MG_THREADS) {
        const int p = p0 + threadIdx.x;
        int set = 0;
        if (p < T) {
            const int c = p / S, s = p % S, y = s / side, x = s % side;
            set = chosen[(d.tube ? 0 : c) * cells + (y / d.scale) * d.rand_size + x / d.scale];
            mask[(int64_t)b * T + p] = (uint8_t)set;
        }
        // block-wide exclusive scan of the set flags (ascending position order)
        const unsigned bal = __ballot_sync(0xffffffffu, set);
        const int in_warp = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int before = running;
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
        if (set) {
            const int64_t j = (int64_t)b * per_row + before + in_warp;
            if (j < total) idx[j] = p;
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < MG_THREADS / 32; ++w) t += warp_tot[w]; running += t; }
        __syncthreads();
    }
}

}  // namespace msst
using namespace msst;

extern "C" int msst_draw_masks(const msst_maskgen_dims* d, uint8_t* mask, int64_t* idx, msst_stream_t stream) {
    MSST_REQUIRE(d && d->B > 0 && d->C > 0 && d->rand_size > 0 && d->scale > 0 && d->nm > 0, "draw_masks: bad dims");
    const int cells = d->rand_size * d->rand_size;
    MSST_REQUIRE(d->mask_count > 0 && d->mask_count <= cells, "draw_masks: mask_count %d outside (0, %d]", d->mask_count, cells);
    const int64_t per_row = (int64_t)d->mask_count * d->scale * d->scale * d->C;
    MSST_REQUIRE((int64_t)d->B * per_row >= (int64_t)d->B * d->nm, "draw_masks: mask has fewer set positions than batch * num_masked");
    const size_t smem = (size_t)(d->tube ? 1 : d->C) * cells * 5;
    MSST_REQUIRE(smem <= 48 * 1024, "draw_masks: %d cells x %d draws exceed shared memory", cells, d->tube ? 1 : d->C);
    mask_draw_kernel<<<d->B, MG_THREADS, (smem + 15) & ~(size_t)15, (cudaStream_t)stream>>>(*d, mask, idx);
    MSST_LAUNCH_CHECK();
    return MSST_OK;
}
