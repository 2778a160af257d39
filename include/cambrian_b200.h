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
/* number of kernels this library has launched in this process (bench.py reports the per-step delta as gpu_launches) */
int64_t cb_launch_count(void);

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
 *   lse : [batch*q_side*q_side, 16] fp32 log2-domain log-sum-exp, needed by the backward (may be NULL)
 *   windowed = 1: k/v are the reference's window-rearranged [N, r*r, hidden] tensors (vision_sampler.py API);
 *   windowed = 0: natural grid layout (the fast path used by cambrian_arch: no permute/contiguous copies) */
int cb_sva_window_attn_fwd(const void* q, void* out, float* lse, int num_towers, const void* const* k,
                           const void* const* v, const void* const* mask, const int* r, int batch,
                           int q_side, int hidden, int windowed, void* stream);
int cb_sva_window_attn_bwd(const void* q, const void* out, const void* dout, const float* lse, void* dq,
                           int num_towers, const void* const* k, const void* const* v,
                           const void* const* mask, void* const* dk, void* const* dv, const int* r,
                           int batch, int q_side, int hidden, int windowed, void* stream);

/* ---- normalisation ----------------------------------------------------------------------------
 * LayerNorm over the last dim C (nn.LayerNorm, eps in fp32 statistics).  `pos` (may be NULL) is the
 * SVA pos_embed [r*r, C] added to row x BEFORE normalising, indexed by the cell's position in its
 * r x r window of a side x side grid (vision_sampler.py:304-309 fused with :170-175).
 * mean/rstd [rows] fp32 are written when non-NULL (needed by the backward). */
int cb_layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd,
                     int64_t rows, int C, float eps, const void* pos, int side, int r, void* stream);
/* dres (may be NULL): gradient arriving on the residual branch, fused as dx = dres + d(norm input).
 * side = 0 selects the window-rearranged layout [N, r*r, C] for `pos` (row % (r*r)); side > 0 the natural grid. */
int cb_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
                     void* dx, const void* dres, void* dgamma, void* dbeta, float* workspace,
                     int64_t workspace_floats, int64_t rows, int C, const void* pos, int side, int r, void* stream);
/* LLaMA RMSNorm.  hf_cast = 0: (w * x_hat_fp32).to(bf16) — the variant the reference trains with
 * (train_fsdp.py:1429-1435);  hf_cast = 1: w * x_hat.to(bf16) — stock HF LlamaRMSNorm (inference). */
int cb_rmsnorm_fwd(const void* x, const void* gamma, void* y, float* rstd, int64_t rows, int C, float eps,
                   int hf_cast, void* stream);
int cb_rmsnorm_bwd(const void* dy, const void* x, const void* gamma, const float* rstd, void* dx,
                   const void* dres, void* dgamma, float* workspace, int64_t workspace_floats, int64_t rows, int C,
                   void* stream);
int64_t cb_norm_bwd_workspace_floats(int64_t rows, int C);

/* ---- softmax attention: tcgen05 flash attention ------------------------------------------------
 * q/k/v/o element (b, s, head, d) lives at base + b*bs + s*ss + head*hd + d (element strides), so the packed
 * QKV GEMM output is consumed in place.  nh % nkv == 0 (GQA).  kmask [B, Skv] bool (1 byte, 1 = attend) or NULL.
 * causal: key k visible to query i iff k <= i + (Skv - Sq).  lse [B, nh, Sq] fp32 (log2 domain) may be NULL.
 * Replaces torch SDPA inside HF CLIPAttention / Dinov2SelfAttention / timm Attention (clip_encoder.py:104,
 * dino_encoder.py:159, siglip_encoder.py:97) and HF LlamaSdpaAttention with the 4-D causal+padding mask of
 * cambrian_llama.py:123-128.  head_dim: any multiple of 8 up to 128 (fwd); 64 or 128 (bwd). */
int cb_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, const void* kmask, int B, int nh,
                int nkv, int Sq, int Skv, int hd, int64_t q_bs, int64_t q_ss, int64_t k_bs, int64_t k_ss,
                int64_t v_bs, int64_t v_ss, int64_t o_bs, int64_t o_ss, float scale, int causal, void* stream);
/* delta [B, nh, Sq] fp32 scratch; dq_acc [B, Sq, nh, hd] fp32 ZERO-INITIALISED by the caller, receives the
 * unscaled sum dS K (multiply by `scale` when converting, cb_f32_to_bf16); dk/dv written (bf16, strided). */
int cb_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                float* delta, float* dq_acc, void* dk, void* dv, const void* kmask, int B, int nh, int nkv, int Sq,
                int Skv, int hd, int64_t q_bs, int64_t q_ss, int64_t k_bs, int64_t k_ss, int64_t v_bs, int64_t v_ss,
                int64_t o_bs, int64_t o_ss, int64_t do_bs, int64_t do_ss, int64_t dk_bs, int64_t dk_ss,
                int64_t dv_bs, int64_t dv_ss, float scale, int causal, void* stream);

/* ---- elementwise / gather / reduction kernels (HBM-bound) -------------------------------------- */
/* y = act(x), dx = dy * act'(x); n elements (n % 8 == 0).  nn.GELU of vision_sampler.py:241, cambrian_arch.py:49,56 */
int cb_act_fwd(const void* x, void* y, int64_t n, int act, void* stream);
int cb_act_bwd(const void* dy, const void* x, void* dx, int64_t n, int act, void* stream);
/* out = silu(gate) * up (HF LlamaMLP); gate/up rows may live in one [rows, 2I] buffer (ld_in) */
int cb_swiglu_fwd(const void* gate, const void* up, void* out, int64_t rows, int I, int64_t ld_in, int64_t ld_out,
                  void* stream);
int cb_swiglu_bwd(const void* dout, const void* gate, const void* up, void* dgate, void* dup, int64_t rows, int I,
                  int64_t ld_in, int64_t ld_dout, int64_t ld_dgu, void* stream);
/* in-place rotary embedding of the first n_heads heads of each row of a packed [rows, ld] buffer;
 * pos int64 [rows]; cos/sin fp32 [max_pos, hd/2]; inverse = 1 applies the transpose (backward).
 * HF apply_rotary_pos_emb as used by LlamaAttention (cambrian_llama.py:142-164). */
int cb_rope(void* buf, const int64_t* pos, const float* cos_t, const float* sin_t, int64_t rows, int n_heads, int hd,
            int64_t ld, int max_pos, int inverse, void* stream);
/* inputs_embeds [B,S,H]: embed_tokens gather + image-span replace + image_newline column
 * (cambrian_arch.py:413-420, :457-490).  img [B, q*q, H] or NULL (text only); img_start int32 [B] (<0: no image) */
int cb_embed_splice(const int64_t* ids, const int32_t* img_start, const void* embed, const void* img,
                    const void* newline, void* out, int B, int S, int H, int q_side, int64_t vocab, void* stream);
int cb_embed_splice_bwd(const void* dout, const int64_t* ids, const int32_t* img_start, void* d_embed, void* d_img,
                        void* d_newline_rows, int B, int S, int H, int q_side, int64_t vocab, void* stream);
/* ViT token assembly: out[b] = [cls + pos[0]] ++ (patch[b] + pos[1:])  (cls may be NULL: SigLIP) */
int cb_add_pos_tokens(const void* patch, const void* cls, const void* pos, void* out, int B, int N, int C,
                      void* stream);
/* fp32 bilinear token-grid resize, align_corners=False (clip_encoder.py:83-88 and siblings, cambrian_arch.py:397-400)
 * in: [B, h, w, C] rows at in_bs batch stride; out row (b, oy, ox) at out + b*out_bs + (oy*tw+ox)*out_ld + out_col0 */
int cb_bilinear(const void* in, void* out, int B, int h, int w, int th, int tw, int C, int64_t in_bs, int64_t out_bs,
                int out_ld, int out_col0, void* stream);
/* im2col for strided patch convolutions feeding cb_gemm_bf16 */
int cb_patchify_nchw(const void* img, void* out, int B, int Cin, int R, int p, int Kpad, void* stream);
int cb_patchify_nhwc(const void* in, void* out, int B, int H, int W, int C, int p, void* stream);
/* depthwise 7x7 conv, NHWC, weights [7,7,C] (timm ConvNeXtBlock.conv_dw via clip_convnext_encoder.py:121-144) */
int cb_dwconv7(const void* in, const void* w, const void* bias, void* out, int B, int H, int W, int C, void* stream);
int cb_add_inplace(void* dst, const void* src, int64_t n, void* stream);
/* out[g, c] = scale * sum_r x[g*rows_per_group + r, c]; either output may be NULL (cambrian_arch.py:377 mean; bias grads) */
int cb_group_colsum(const void* x, void* out_bf16, float* out_f32, int groups, int64_t rows_per_group, int C,
                    float scale, int accumulate, void* stream);
int cb_group_broadcast(const void* dmean, void* dx, int groups, int64_t rows_per_group, int C, float scale,
                       int accumulate, void* stream);
/* d pos_embed [r*r, C] from the gradient of the (x + pos) rows on the natural grid layout */
int cb_pos_grad(const void* dx, void* dpos, int B, int side, int r, int C, int accumulate, void* stream);
/* fp32 [rows, cols] contiguous -> bf16 rows at stride out_ld, times scale (dQ of cb_attn_bwd into a packed dQKV buffer) */
int cb_f32_to_bf16(const float* in, void* out, int64_t rows, int cols, int64_t out_ld, float scale, void* stream);
/* per-row cross entropy on bf16 logits [rows, V] (fp32 math, cambrian_llama.py:408-422); loss_rows [rows];
 * loss_acc (may be NULL) += {sum of losses, number of non-ignored rows}; write_grad overwrites the logits in place
 * with (softmax - onehot) * grad_scale. */
int cb_cross_entropy(void* logits, const int64_t* labels, float* loss_rows, float* loss_acc, int64_t rows, int64_t V,
                     int64_t ld, float grad_scale, int write_grad, int64_t ignore_index, void* stream);
/* same, with an additional gradient scale read from DEVICE memory (grad_scale_dev[0], may be NULL): the mean over
 * non-ignored labels (cambrian_llama.py:411-422) without counting them on the host */
int cb_cross_entropy_ex(void* logits, const int64_t* labels, float* loss_rows, float* loss_acc, int64_t rows, int64_t V,
                        int64_t ld, float grad_scale, const float* grad_scale_dev, int write_grad, int64_t ignore_index,
                        void* stream);
/* in-LLM SVA site (cambrian_llama.py:168-207): gather the q*q latent rows of the image span [start, start+q*(q+1))
 * of hidden [B,S,H] into lat [B*q*q, H] / scatter updated rows back in place (newline rows untouched) */
int cb_span_gather(const void* hidden, void* lat, int B, int S, int H, int start, int q_side, void* stream);
int cb_span_scatter(void* hidden, const void* lat, int B, int S, int H, int start, int q_side, void* stream);
/* Dynamic-shape (per-sample, non-square) variant of the two above: the span holds q_h rows of (q_w queries + 1 newline)
 * (cambrian_llama.py:208-253, final_vision_feature_size[b] = (q_h, q_w)). */
int cb_span_gather_hw(const void* hidden, void* lat, int B, int S, int H, int start, int q_h, int q_w, void* stream);
int cb_span_scatter_hw(void* hidden, const void* lat, int B, int S, int H, int start, int q_h, int q_w, void* stream);
/* Window rearrangement with the `unpad_image` crop of the q x q window grid
 * (rearrange_vision_tower_features_inference, cambrian_arch.py:289-330; full range = _train, :271-287):
 * feat [B, q*r, q*r, C] -> out [B*(y1-y0)*(x1-x0), r*r, C] for query rows [y0,y1) x columns [x0,x1). */
int cb_window_gather(const void* feat, void* out, int B, int q_side, int r, int C, int y0, int y1, int x0, int x1,
                     void* stream);
/* Ragged embed + splice of the dynamic branch (cambrian_arch.py:493-609): out[row] = embed[src[row]] if src >= 0,
 * zeros if src == -1 (padding), newline if src == INT32_MIN, img[-2 - src] otherwise; rows = B * max_len. */
int cb_embed_splice_ragged(void* out, const void* embed, const void* img, const void* newline, const int32_t* src,
                           int64_t rows, int H, void* stream);
/* LLaMA MLP first half in one launch (HF LlamaMLP, reached from cambrian_llama.py:142-166): W = [gate_proj; up_proj]
 * [2F, K]; gu_out [M, 2F] = A W^T (bf16 pre-activations, saved for backward), act_out [M, F] = silu(gate) * up.
 * F % 128 == 0.  CTA-pair tcgen05 kernel whose tile pairs 128 gate columns with the matching 128 up columns. */
int cb_gemm_swiglu_bf16(const void* A, const void* W, void* gu_out, void* act_out, int M, int F, int K, int64_t lda,
                        int64_t ldw, int64_t ld_gu, int64_t ld_act, void* stream);
/* Image preprocessing on the GPU — replaces, per tower, the host chain of `process_images`
 * (mm_utils.py:186-201): expand2square(img, int(mean*255)) -> PIL Image.resize((R,R)) [bicubic, antialiased, uint8,
 * bit-exact with Pillow's Resample.c] -> x/255 -> (x-mean)/std.  img: device uint8 [H,W,3] RGB; out: bf16 [3,R,R];
 * out_u8 (optional, may be NULL): the resized uint8 image [R,R,3]; pad_rgb / mean / std: HOST arrays of 3.
 * cb_resample_ksize / cb_resample_coeffs are the host-side coefficient generator (bounds [out,2], kk [out,ksize],
 * 22-bit fixed point) exposed for tests. */
int cb_resample_ksize(int in_size, int out_size);
int cb_resample_coeffs(int in_size, int out_size, int32_t* bounds, int32_t* kk);
int64_t cb_preprocess_workspace_bytes(int H, int W, int R);
int cb_preprocess_image(const uint8_t* img, int H, int W, int R, const int32_t* pad_rgb, const float* mean,
                        const float* std, void* out, uint8_t* out_u8, void* workspace, int64_t workspace_bytes,
                        void* stream);
/* AdamW on fp32 master weights / moments with bf16 gradients, writing the bf16 compute copy */
int cb_adamw(float* p, float* m, float* v, const void* g, void* p16, int64_t n, float lr, float beta1, float beta2,
             float eps, float weight_decay, int step, float grad_scale, void* stream);
/* cb_adamw with (a) the gradient scale read from DEVICE memory (clip_coef[0], written by cb_clip_coef; null = use
 * grad_scale) so a clipped step needs no host sync, and (b) a background launch shape (one small block per SM) that
 * co-resides with a persistent GEMM CTA on every SM.  Replaces HF Trainer's clip_grad_norm_ + AdamW.step
 * (cambrian_trainer.py:242-381; transformers Trainer max_grad_norm default 1.0). */
int cb_adamw_ex(float* p, float* m, float* v, const void* g, void* p16, int64_t n, float lr, float beta1, float beta2,
                float eps, float weight_decay, int step, float grad_scale, const float* clip_coef, int background,
                void* stream);
/* Deterministic gradient of the embedding rows (backward of embed_tokens inside cb_embed_splice): rows of dout [n, H] with
 * equal keys[t] (token id; >= vocab = no gradient, e.g. the image span) are summed in position order by one block and
 * written once to d_embed [vocab, H] (pre-zeroed).  `order` = positions stably sorted by key.  Replaces the racing
 * bf16x2 atomics of cb_embed_splice_bwd (pass d_embed = NULL there). */
int cb_embed_grad_sorted(const void* dout, const int64_t* keys, const int32_t* order, void* d_embed, int64_t n, int H,
                         int64_t vocab, void* stream);
/* Decode-shaped projection, M <= 8 rows: y[M,N] = x[M,K] W[N,K]^T (+ bias[N]) (+ residual[M,N]); W in nn.Linear layout.
 * Weight-streaming CUDA-core kernel (each weight byte read once, fp32 accumulation) used by the KV-cache decode step of
 * generate() (cambrian_llama.py:437-483) where a 128-row tensor-core tile would idle; y bf16 or fp32 (lm_head logits). */
int cb_gemv_bf16(const void* x, const void* w, void* y, int M, int N, int K, int64_t ldx, int64_t ldw, int64_t ldy,
                 const void* bias, const void* residual, int64_t ldr, int out_fp32, void* stream);
/* In-place sum all-reduce of the bf16 range [offset_bytes, +nbytes) of a SYMMETRIC buffer (same offset on every rank) —
 * the bucketed gradient reduction of the data-parallel step (reference: inside torch_xla FSDP; explicit helper
 * cambrian_trainer.py:181-190).  multicast_base != 0: NVLS path (multimem.ld_reduce / multimem.st, the NVSwitch adds and
 * fans out); 0: peer loads / stores over NVLink.  buffer_ptrs / signal_pad_ptrs: HOST arrays of `world` peer-mapped device
 * addresses (this rank's own included).  epoch: strictly increasing across launches on all ranks alike, advance by 2 per
 * call.  nbytes % (16 * world) == 0.  Every rank must issue the same calls in the same order. */
int cb_allreduce_symm_bf16(uint64_t multicast_base, const uint64_t* buffer_ptrs, const uint64_t* signal_pad_ptrs,
                           int64_t offset_bytes, int64_t nbytes, int rank, int world, uint32_t epoch, int ctas, void* stream);
/* GEMM tile scheduling: 1 (default; env CB_GEMM_CLC=0 to start with 0) = one CTA / CTA pair per tile in the grid, tiles
 * handed out by Cluster Launch Control so the GEMM tolerates SMs held by collectives / the background optimizer;
 * 0 = static persistent walk.  Returns the previous setting.  Results are bit-identical either way. */
int cb_gemm_set_dynamic_scheduling(int on);
/* acc[0] += sum of squares of n bf16 values (deterministic two-stage reduction; workspace >= grid floats, 4096 suffices) */
int cb_sumsq_bf16(const void* g, int64_t n, float* acc, float* workspace, int64_t workspace_floats, int background,
                  void* stream);
/* coef[0] = inv_world * min(1, max_norm / (sqrt(sumsq[0]) * inv_world + 1e-6)); coef[1] = that norm; sumsq[0] = 0
 * (torch.nn.utils.clip_grad_norm_ on the rank-averaged gradient, all on the device) */
int cb_clip_coef(float* sumsq, float max_norm, float inv_world, float* coef, void* stream);
/* `sep` aggregator layer (VisionAggregationLayer.forward, vision_sampler.py:368-398): per-query softmax over the towers
 * of the weight_mlp logits and the weighted sum of the per-tower aggregates added to the query stream,
 *   out[n,:] = q_in[n,:] + sum_t softmax(logits[n,:T])[t] * aggs[t][n,:]        (replaces .softmax(-1), torch.stack,
 * (agg * weight).sum(2) and the residual add).  logits [N, ld_logits >= T] bf16 (columns >= T are padding), aggs = HOST
 * array of T device pointers to [N, C] bf16.  Backward: daggs[t] = w_t * dout, dlogits = softmax adjoint of
 * g_t = <dout, agg_t> (padding columns zeroed); d q_in = dout. */
int cb_tower_combine_fwd(const void* logits, int ld_logits, const void* const* aggs, const void* q_in, void* out, int64_t N,
                         int C, int num_towers, void* stream);
int cb_tower_combine_bwd(const void* logits, int ld_logits, const void* const* aggs, const void* dout, void* const* daggs,
                         void* dlogits, int64_t N, int C, int num_towers, void* stream);
/* adjoint of cb_bilinear on contiguous grids: dout [B, th, tw, C] -> din [B, h, w, C] (backward of the query-grid resize
 * cambrian_arch.py:394-401 and of the towers' token interpolation, clip_encoder.py:83-88 and siblings); deterministic */
int cb_bilinear_bwd(const void* dout, void* din, int B, int h, int w, int th, int tw, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CAMBRIAN_B200_H */
