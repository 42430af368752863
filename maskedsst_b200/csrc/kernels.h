// Internal (C++) launch functions shared between translation units; the public surface is include/msst.h.
#pragma once
#include "common.cuh"

namespace msst {

// layernorm.cu
int layernorm_fwd(const float* x, const float* w, const float* b, void* y, int y_bf16, float* stats, int64_t rows, int D,
                  float eps, cudaStream_t st);
// cast_out / cast_drop / cast_colsum (optional, D % 4 == 0 && D <= 128): also emit bf16(dropout(dx)) and its column sums
int layernorm_bwd(const float* x, const float* w, const float* stats, const float* dy, const float* dx_add, float* dx,
                  float* dw, float* db, int64_t rows, int D, cudaStream_t st, __nv_bfloat16* cast_out = nullptr,
                  Drop cast_drop = Drop{0, nullptr, 0, 0, 1.f}, float* cast_colsum = nullptr);

// gemm_f32.cu
int linear_fwd_f32(const float* x, const float* W, const float* bias, const float* residual, float* y, float* pre_act,
                   int64_t M, int N, int K, int act, Drop drop, cudaStream_t st);
int linear_bwd_data_f32(const float* dy, const float* W, const float* pre_act, const float* dx_add, float* dx, int64_t M,
                        int N, int K, Drop drop, cudaStream_t st);
int linear_bwd_weight_f32(const float* dy, const float* x, float* dW, float* db, int64_t M, int N, int K, cudaStream_t st);
int colsum_f32(const float* dy, float* db, int64_t M, int N, cudaStream_t st);
int dropout_apply_f32(const float* x, float* y, int64_t n, Drop drop, cudaStream_t st);

// gemm_bf16.cu (tcgen05 / TMA)
struct GemmBf16Args {
    const __nv_bfloat16* A; const __nv_bfloat16* B;   // A [M,K], B [N,K], both K contiguous
    int64_t M; int N, K;
    const float* bias; const float* residual;          // fp32 [N], fp32 [M,N]
    void* out; int out_fp32;                           // [M,N] fp32 or bf16
    __nv_bfloat16* pre_act; const __nv_bfloat16* aux;  // optional bf16 [M,N] out (pre-activation) / in (GELU' argument)
    int act;                                           // 0 none, 1 GELU(erf), 2 multiply by GELU'(aux)
    Drop drop;
    // optional fused LayerNorm of the finished output rows (fp32 out + bias + residual GEMMs with N <= 128):
    const float* ln_w; const float* ln_b; __nv_bfloat16* ln_out; float* ln_stats;
    // optional fused LayerNorm BACKWARD of the output rows (plain fp32-out data-gradient GEMMs with N <= 128): the accumulator row
    // is dy of a LayerNorm whose input rows were lnb_x (statistics lnb_stats, weight ln_w).  The GEMM then writes
    //   out[r,:] = LN'(dy)[r,:] + residual[r,:]   (fp32; `residual` = gradient arriving over the skip connection)
    // accumulates lnb_dw += sum_r dy*xhat, lnb_db += sum_r dy, and optionally emits lnb_cast = bf16(dropout(out)) (Drop = `drop`)
    // with its column sums added to lnb_colsum -- i.e. layernorm_bwd() including its fused cast, without the dy round trip.
    const float* lnb_x; const float* lnb_stats; float* lnb_dw; float* lnb_db; __nv_bfloat16* lnb_cast; float* lnb_colsum;
    // optional (act == 2 data gradient with N <= 64): column sums of the output rows (before the bf16 rounding) += colsum
    float* colsum;
};
int gemm_tn_bf16(const GemmBf16Args& a, cudaStream_t st);
int gemm_wgrad_bf16(const __nv_bfloat16* dy, const __nv_bfloat16* x, float* dW, int64_t M, int N, int K, cudaStream_t st);

// bf16_util.cu
int cast_rows_f32(const float* x, __nv_bfloat16* y, float* colsum, int64_t M, int N, Drop drop, cudaStream_t st);
int cast_rows_bf16(const __nv_bfloat16* x, __nv_bfloat16* y, float* colsum, int64_t M, int N, Drop drop, cudaStream_t st);
struct WeightCastEntry { const float* src; __nv_bfloat16* dst; __nv_bfloat16* dst_t; int rows, cols; };
struct WeightCastTable { WeightCastEntry e[32]; int n; };
int weights_to_bf16(const WeightCastTable& t, cudaStream_t st);

// attention_bf16.cu (mma.sync, dh = 64)
int attention_fwd_bf16(const msst_attn_dims* d, const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse, cudaStream_t st);
int attention_bwd_bf16(const msst_attn_dims* d, const __nv_bfloat16* qkv, const __nv_bfloat16* out, const float* lse,
                       const __nv_bfloat16* d_out, __nv_bfloat16* d_qkv, cudaStream_t st);

// attention_tc.cu (tcgen05 / TMEM forward and backward, N <= 64)
struct AttnGeom;
int attention_fwd_tc(const AttnGeom& g, const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse, Drop drop, cudaStream_t st);
// tcgen05 backward for packed short sequences (N <= 64), contiguous or strided rows (both transformer stacks)
bool attention_bwd_tc_supported(const AttnGeom& g);
// tcgen05 forward for long sequences (N > 64): two passes over the key tiles, O accumulated in TMEM without rescaling
bool attention_fwd_tc_long_supported(const AttnGeom& g);
int attention_fwd_tc_long(const AttnGeom& g, const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse, Drop drop, cudaStream_t st);
int attention_bwd_tc(const AttnGeom& g, const __nv_bfloat16* qkv, const float* lse, const __nv_bfloat16* d_out, __nv_bfloat16* d_qkv,
                     Drop drop, cudaStream_t st);

// attention_tc_long_bwd.cu (tcgen05 / TMEM backward for long sequences, N > 64): dK/dV pass + dQ pass, D = rowsum(dO * O) pre-pass
bool attention_bwd_tc_long_supported(const AttnGeom& g);
int attention_bwd_tc_long(const AttnGeom& g, const __nv_bfloat16* qkv, const __nv_bfloat16* out, const float* lse, const __nv_bfloat16* d_out,
                          __nv_bfloat16* d_qkv, Drop drop, cudaStream_t st);

// attn_block_tc.cu (tcgen05): per-head QKV projection fused into the attention kernels, N <= 64 -- q/k/v never reach HBM
bool attn_block_supported(const AttnGeom& g, int D);
// optional tail of the forward kernel (the attention block's out-projection, residual add and the FeedForward pre-norm in the same launch):
//   xmid [R, D] fp32 = x + dropout(out . w_out^T + b_out)   (w_out [D, H*64] bf16, Drop = `drop`)
//   h2 [R, D] bf16 = LayerNorm(xmid; ln_w, ln_b), ln_stats [R, 2] = (mean, rstd);   `out` may then be null (inference: o is not needed)
struct AttnBlockOut {
    const __nv_bfloat16* w_out; const float* b_out; const float* x; float* xmid;
    const float* ln_w; const float* ln_b; __nv_bfloat16* h2; float* ln_stats; Drop drop;
};
int attn_block_fwd(const AttnGeom& g, int D, const __nv_bfloat16* h, const __nv_bfloat16* w_qkv, __nv_bfloat16* out, float* lse, Drop drop,
                   cudaStream_t st, const AttnBlockOut* tail = nullptr);
// recomputes q/k/v from h; d_qkv [R, 3*H*64] bf16 (for the weight gradient) and d_h [R, D] fp32 = d_qkv . W_qkv (data gradient, accumulated over heads in TMEM)
int attn_block_bwd(const AttnGeom& g, int D, const __nv_bfloat16* h, const __nv_bfloat16* w_qkv, const __nv_bfloat16* w_qkv_t,
                   const __nv_bfloat16* d_out, const float* lse, __nv_bfloat16* d_qkv, float* d_h, Drop drop, cudaStream_t st);

// mlp_block_tc.cu (tcgen05): Linear -> GELU -> dropout -> Linear -> dropout -> residual (-> next pre-norm LayerNorm) in one kernel
bool mlp_block_supported(int D, int M);
int mlp_block_fwd(const __nv_bfloat16* h2, const float* xmid, const __nv_bfloat16* w1, const __nv_bfloat16* w2, const float* b1, const float* b2,
                  __nv_bfloat16* u, __nv_bfloat16* g, float* y, const float* ln_w, const float* ln_b, __nv_bfloat16* h1, float* ln_stats,
                  int64_t R, int D, int M, Drop drop_h, Drop drop_o, cudaStream_t st);
// backward of the block: du = (dyb W2) gelu'(u) drop stays on the SM; dW1 / dW2 / db1 accumulate (+=), dh [R,D] fp32 = du W1
int mlp_block_bwd(const __nv_bfloat16* dyb, const __nv_bfloat16* u, const __nv_bfloat16* g, const __nv_bfloat16* h2, const __nv_bfloat16* w2t,
                  const __nv_bfloat16* w1t, float* dw1, float* dw2, float* db1, float* dh, int64_t R, int D, int M, Drop drop_h, cudaStream_t st);

// attention_f32.cu
int attention_fwd_f32(const msst_attn_dims* d, const float* qkv, float* out, float* lse, cudaStream_t st);
int attention_bwd_f32(const msst_attn_dims* d, const float* qkv, const float* out, const float* lse, const float* d_out,
                      float* d_qkv, cudaStream_t st);

}  // namespace msst
