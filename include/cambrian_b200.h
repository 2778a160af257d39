/* cambrian_b200 — C ABI of the B200-native Cambrian-1 hot-path kernels (libcambrian_b200.so).
 *
 * The reference (cambrian-mllm/cambrian) has NO FFI / plugin layer: its hot path is Python
 * nn.Modules that dispatch torch ops.  This header therefore declares the
 * boundary a maintainer binds from Python with ctypes: one entry point per
 * operator the reference modules execute, each citing the reference call site it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch allocations); the library never
 *     allocates, frees or retains memory (SURVEY.md §8b "Memory ownership");
 *   - `stream` is the caller's cudaStream_t (torch.cuda.current_stream().cuda_stream) passed as void*;
 *   - all entry points are re-entrant per stream, keep no global mutable state besides one-time
 *     function-attribute caches, and return 0 on success or a CB_ERR_* code; cb_last_error()
 *     returns a thread-local message.  The Python wrappers re-raise ValueError / RuntimeError as
 *     the reference modules do (vision_sampler.py:202-206, builder.py:147);
 *   - activations and parameters are bf16 (uint16 storage) unless a parameter says otherwise;
 *     statistics, LSE and losses are fp32; token ids / labels / positions are int64.
 */
#ifndef CAMBRIAN_B200_H
#define CAMBRIAN_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_OK 0
#define CB_ERR_INVALID 1
#define CB_ERR_CUDA 2
#define CB_ERR_UNSUPPORTED 3

/* activation codes for cb_gemm_bf16 / cb_act_* */
#define CB_ACT_NONE 0
#define CB_ACT_GELU_ERF 1   /* nn.GELU()            vision_sampler.py:241, cambrian_arch.py:49,56 */
#define CB_ACT_GELU_TANH 2  /* gelu(approximate='tanh') (SigLIP HF variant)                       */
#define CB_ACT_QUICK_GELU 3 /* CLIP quick_gelu      clip_encoder.py:47 -> HF CLIPMLP              */
#define CB_ACT_SILU 4       /* LLaMA SwiGLU gate    cambrian_llama.py:142-164 -> HF LlamaMLP      */

int cb_version(void);
const char* cb_last_error(void);
int cb_sm_count(void);

/* ---- dense contraction: tcgen05 + TMA persistent GEMM --------------------------------------
 * C[b] (M x N row-major, ldc)  (+)=  epi( alpha * opA(A[b]) (M x K) * opB(B[b]) (K x N) )
 *   a_mn = 0: A is [M, K] with K contiguous (lda);  a_mn = 1: A is [K, M] with M contiguous
 *   b_mn = 0: B is [N, K] with K contiguous (ldb) — the nn.Linear weight layout;  b_mn = 1: [K, N]
 *   epi(v) = act(v + bias[n]) * colscale[n] + residual[m, n]; each of bias/colscale/residual may be NULL
 *   out_fp32: C is fp32 instead of bf16; accumulate: C += epi(..)
 *   force_bn: 0 = heuristic tile width, else 64 / 128 / 256
 * Replaces every nn.Linear / matmul of the hot path: HF CLIP/DINOv2/LLaMA linears, timm SigLIP /
 * ConvNeXt linears, vision_sampler.py:170-175,254-257 (q/k/v/o/proj_* Linear), cambrian_arch.py:49,56
 * (mm_projector, mm_projector_aux), cambrian_llama.py:402-409 (lm_head), and their autograd backward.
 * Requires lda/ldb and the batch strides to be multiples of 8 elements and 16-byte aligned bases (TMA). */
int cb_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int batch,
                 int64_t lda, int64_t ldb, int64_t ldc, int64_t bsa, int64_t bsb, int64_t bsc,
                 int a_mn, int b_mn, const void* bias, const void* colscale, const void* residual,
                 int64_t ldr, int64_t bsr, float alpha, int act, int out_fp32, int accumulate,
                 int force_bn, void* stream);

/* ---- SVA window attention -------------------------------------------------
 * Fused replacement of rearrange_vision_tower_features_train (cambrian_arch.py:271-287) + the SDPA
 * inside MultiKVCrossAttention.forward (vision_sampler.py:191-230).
 *   q, out : [batch*q_side*q_side, hidden] bf16 (hidden = 1024 = 16 heads x 64)
 *   k[t], v[t] : [batch, (r[t]*q_side)^2, hidden] bf16 projected key/value grids in natural layout
 *   mask[t] : [batch*q_side*q_side, r[t]*r[t]] bool (1 byte) or NULL (= all true); `mask` itself may be NULL
 *   lse : [batch*q_side*q_side, 16] fp32 log2-domain log-sum-exp, needed by the backward (may be NULL) */
int cb_sva_window_attn_fwd(const void* q, void* out, float* lse, int num_towers, const void* const* k,
                           const void* const* v, const void* const* mask, const int* r, int batch,
                           int q_side, int hidden, void* stream);
int cb_sva_window_attn_bwd(const void* q, const void* out, const void* dout, const float* lse, void* dq,
                           int num_towers, const void* const* k, const void* const* v,
                           const void* const* mask, void* const* dk, void* const* dv, const int* r,
                           int batch, int q_side, int hidden, void* stream);

/* ---- normalisation ----------------------------------------------------------------------------
 * LayerNorm over the last dim C (nn.LayerNorm, eps in fp32 statistics).  `pos` (may be NULL) is the
 * SVA pos_embed [r*r, C] added to row x BEFORE normalising, indexed by the cell's position in its
 * r x r window of a side x side grid (vision_sampler.py:304-309 fused with :170-175).
 * mean/rstd [rows] fp32 are written when non-NULL (needed by the backward). */
int cb_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd,
                     int64_t rows, int C, float eps, const void* pos, int side, int r, void* stream);
int cb_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
                     void* dx, void* dgamma, void* dbeta, float* workspace, int64_t workspace_floats,
                     int64_t rows, int C, const void* pos, int side, int r, void* stream);
/* LLaMA RMSNorm.  hf_cast = 0: (w * x_hat_fp32).to(bf16) — the variant the reference trains with
 * (train_fsdp.py:1429-1435);  hf_cast = 1: w * x_hat.to(bf16) — stock HF LlamaRMSNorm (inference). */
int cb_rmsnorm_fwd(const void* x, const void* gamma, void* y, float* rstd, int64_t rows, int C, float eps,
                   int hf_cast, void* stream);
int cb_rmsnorm_bwd(const void* dy, const void* x, const void* gamma, const float* rstd, void* dx,
                   void* dgamma, float* workspace, int64_t workspace_floats, int64_t rows, int C,
                   void* stream);
int64_t cb_norm_bwd_workspace_floats(int64_t rows, int C);

#ifdef __cplusplus
}
#endif
#endif /* CAMBRIAN_B200_H */
