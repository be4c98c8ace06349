/* lecb.h — C ABI of liblecb.so: the B200 (sm_100a) kernels behind the dual-prompt CLIP scoring path.
 *
 * The reference (JarvisUSTC/Language-Enhanced-CLIP-For-Multi-label-Image-Recognition) is pure
 * Python/PyTorch and has no FFI of its own: its "operator interface" for this path is
 * the ATen call issued at each line cited below.  Every entry point replaces the cited call site(s).
 * Paths are relative to project/my_code/:  T = trainers/Caption_distill_double.py,
 * M = clip/model.py, U = trainers/utils.py.
 *
 * Conventions (all entry points):
 *   - plain C: raw DEVICE pointers + explicit sizes; no torch / C++ types in any signature;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised;
 *   - never allocates, never frees, never owns memory; scratch comes in through `ws` arguments;
 *   - returns 0 on success, a negative lecb_status otherwise; `lecb_last_error()` gives the
 *     thread-local message of the last failure on this host thread;
 *   - activations are row-major bf16 "pixel-major" matrices: NHWC images == [B*H*W, C] rows;
 *   - there is NO CPU fallback: without a CUDA device every compute call returns LECB_ERR_CUDA;
 *   - launches use programmatic dependent launch (cudaLaunchKernelEx + the stream-serialization attribute) where it was
 *     measured to pay — the tensor-core kernels always, row kernels with small grids — and every kernel executes
 *     griddepcontrol.wait before its first global access, so ordinary stream order holds for the caller: work enqueued on
 *     `stream` before a call is complete and visible to it, and its results to whatever follows.  Works under CUDA-graph
 *     capture.  Environment LECB_NO_PDL=1 launches everything the ordinary way (LECB_PDL_MODE=0..4: experiment knob).
 */
#ifndef LECB_H_
#define LECB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LECB_ABI_VERSION 1

enum lecb_status {
  LECB_OK = 0,
  LECB_ERR_ARG = -1,     /* bad shape / alignment / null pointer */
  LECB_ERR_CUDA = -2,    /* a CUDA runtime or driver call failed */
  LECB_ERR_UNSUPPORTED = -3
};

/* epilogue flags for lecb_gemm_bf16 / lecb_conv3x3_bf16 */
#define LECB_EPI_RELU 1u        /* y = max(y, 0) after bias (+ residual)                     */
#define LECB_EPI_QUICKGELU 2u   /* y = y * sigmoid(1.702 y)   (M:202-204)                    */
#define LECB_EPI_OUT_F32 4u     /* store fp32 instead of bf16                                 */
#define LECB_EPI_RES_F32 8u     /* `residual` is fp32 [M,N] instead of bf16                   */
#define LECB_GEMM_F16_OPERANDS 16u /* A and W hold IEEE fp16 instead of bf16 (retrieval, T:445)  */
#define LECB_EPI_AVGPOOL2 32u   /* lecb_conv3x3_bf16 only: 2x2 average pool after the activation (M:147,177 stem avgpool; */
                                /* M:27,46 Bottleneck avgpool) fused into the epilogue; out is [B,H/2,W/2,Cout]            */
#define LECB_EPI_MUL_QGELU_GRAD 64u /* lecb_gemm_bf16 only, bf16 out: `residual` holds the MLP pre-activation v (bf16 [M,N]) and  */
                                /* y = (A W^T + bias) * QuickGELU'(v): the backward of M:202-204 fused into the data-gradient */
                                /* GEMM of c_proj (autograd of M:226-227); excludes RELU / QUICKGELU / OUT_F32 / RES_F32      */

int lecb_abi_version(void);
const char* lecb_last_error(void);
/* Number of kernels this library has launched from this process (bench.py's gpu_launches). */
unsigned long long lecb_launch_count(void);

/* ---- tensor-core GEMM (tcgen05 + TMA + TMEM) --------------------------------------------------
 * out[M,N] = epi( A[M,K] · W[N,K]^T + bias[N] (+ residual[M,N]) ), bf16 operands, fp32 accumulate.
 * Replaces every F.linear / 1x1 nn.Conv2d (+ folded eval BatchNorm + ReLU + residual add) on the path:
 * T:409-410 (v_proj, c_proj per patch), M:43,46,49-52 (Bottleneck conv1/conv3/downsample),
 * M:225-227 (transformer in_proj/out_proj/c_fc/c_proj), T:95,100 (text_projection).
 * Requirements: K % 32 == 0, N % 8 == 0, A/W 16-byte aligned, lda == K, ldw == K.
 * `bias` may be NULL; `residual` (bf16 [M,N]) may be NULL; `row_sumsq` (fp32 [M], pre-zeroed) if
 * non-NULL receives sum_n out[m,n]^2 of the stored (rounded) values — feeds the L2 normalisation. */
int lecb_gemm_bf16(const void* A, const void* W, const float* bias, const void* residual, void* out,
                   float* row_sumsq, int64_t M, int N, int K, unsigned flags, void* stream);

/* ---- the same GEMM over TWO A operands (K-concatenation without the concatenated tensor) ----------
 * out[M,N] = epi( A1[M,K1] · W[:, :K1]^T + A2[M,K2] · W[:, K1:]^T + bias[N] ), W bf16 [N, K1+K2], bf16 output.
 * Replaces the tail of a Bottleneck with a projection shortcut, M:46-52:
 *   out = bn3(conv3(out)); identity = downsample(x); out += identity; out = relu(out)
 * as ONE GEMM of [y | x_pooled] against [W3 | Wd] (eval BatchNorm folded, bias = b3 + bd): the [M,N] shortcut tensor is
 * neither written nor read back as a residual.  k blocks [0, K1/64) are TMA-loaded from A1, the rest from A2.
 * Requirements: K1 % 64 == 0, K2 % 64 == 0, N % 8 == 0, 16-byte aligned operands; flags: 0 or LECB_EPI_RELU. */
int lecb_gemm_bf16_dual(const void* A1, int K1, const void* A2, int K2, const void* W, const float* bias, void* out,
                        int64_t M, int N, unsigned flags, void* stream);

/* ---- implicit-GEMM 3x3 convolution, stride 1, pad 1 (TMA im2col + tcgen05) ---------------------
 * x: NHWC bf16 [B,H,W,Cin]; w: bf16 [Cout][3][3][Cin] with eval-BatchNorm folded in; bias fp32 [Cout];
 * out: NHWC bf16 [B,H,W,Cout].  Replaces M:44 (Bottleneck conv2+bn2+relu) and M:174-175 (stem
 * conv2/conv3).  Requirements: Cin % 32 == 0, Cout % 8 == 0. */
int lecb_conv3x3_bf16(const void* x, const void* w, const float* bias, void* out, int B, int H, int Wd,
                      int Cin, int Cout, unsigned flags, void* stream);
/* Wide layers (N >= 256, K >= 256 and a multiple of 64, bf16 output, at least one 256 x 256 tile per SM) run on the CTA-pair
 * variant of the kernel: two CTAs of a cluster own a 256 x 256 tile, each stages its 128 rows of A and half of the W tile,
 * one tcgen05.mma.cta_group::2 (M = 256) reads both halves (gemm_pair.cu).  lecb_set_pair_gemm(0 / 1) switches that path off /
 * on at run time (default on; environment LECB_NO_PAIR=1 starts with it off) and returns the previous setting. */
int lecb_set_pair_gemm(int enable);
/* 1 if lecb_conv3x3_bf16 accepts LECB_EPI_AVGPOOL2 for this problem (halo-tile mode: Cin 32 / 64, Cout <= 128,
 * even H and W, at least two 128-pixel patches per SM), else 0 — the caller then runs lecb_avgpool2x2 itself. */
int lecb_conv3x3_pool_fusable(int B, int H, int Wd, int Cin, int Cout);

/* ---- stem conv1: 3x3, stride 2, pad 1, 3 -> Cout channels, folded BN + ReLU (M:144-145,174-175) ----
 * x: NCHW fp32 [B,3,H,W] (the reference's input layout, T:401); w: fp32 [27][Cout] with row index
 * ci*9+ky*3+kx; out: NHWC bf16 [B,H/2,W/2,Cout].  Cout in {8, 32, 48}. */
int lecb_stem_conv1(const float* x, const float* w, const float* bias, void* out, int B, int H, int W, int Cout,
                    void* stream);
/* The same convolution reading RAW uint8 pixels, NHWC [B,H,W,3] (the output of lecb_crop_resize_u8, or a host batch that
 * skips the float conversion: 4x fewer bytes over PCIe and HBM).  ToTensor + Normalize (transforms.py:396-402: v / 255 then
 * (x - mean) / std, fp32) is applied on the fly from a 3 x 256 table built with those very operations, so the bf16 operand
 * equals the one lecb_stem_conv1 forms from the reference's float tensor.  mean / std: HOST float[3].  Cout = 32. */
int lecb_stem_conv1_u8(const uint8_t* x, const float* w, const float* bias, const float* mean, const float* stdv, void* out,
                       int B, int H, int W, int Cout, void* stream);

/* ---- 2x2 average pool, NHWC bf16 (M:150 stem avgpool, M:23,35 anti-aliased stride) ---- */
int lecb_avgpool2x2(const void* x, void* out, int B, int H, int W, int C, void* stream);

/* ---- mean over the P tokens of each image: x bf16 [B,P,C] -> bf16 and/or fp32 [B,C] (M:92) ---- */
int lecb_token_mean(const void* x, void* out_bf16, float* out_f32, int B, int P, int C, void* stream);

/* ---- row L2 normalisation y = x / ||x||, no epsilon (T:441-442, T:431-436, T:485-488) ---- */
int lecb_l2norm_rows(const void* x, void* y, int64_t rows, int D, int in_is_bf16, int out_is_bf16, void* stream);

/* ---- LayerNorm forward, fp32 statistics (M:193-199); x fp32 [rows,D] -> bf16 and/or fp32 y;
 * optional per-row mean / rstd outputs for the backward ---- */
int lecb_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32,
                       float* mean, float* rstd, int64_t rows, int D, float eps, void* stream);

/* ---- AttentionPool2d, if_pos=False, query = mean token only (M:89-127 via T:413) ----
 * q fp32 [B,C] = q_proj(mean token) (unscaled); kmat, vmat bf16 [B*P,C] = k_proj / v_proj of the patch
 * tokens; out bf16 [B,C] = attention output of token 0 before c_proj.  Head dim 64, heads % 8 == 0. */
int lecb_attnpool_query0(const float* q, const void* kmat, const void* vmat, void* out, int B, int P, int C,
                         int heads, void* stream);

/* ---- causal multi-head self-attention forward (M:221-223 with the mask of M:364-370) ----
 * qkv bf16 [N*L, 3W] (q | k | v), out bf16 [N*L, W]; head dim 64; L <= 128. */
int lecb_causal_attn_fwd(const void* qkv, void* out, int N, int L, int W, int heads, void* stream);

/* ---- multi-head self-attention forward on tcgen05 tensor cores, head dim 64 ----
 * The ViT visual encoder's attention (nn.MultiheadAttention inside ResidualAttentionBlock, M:207-228, called
 * from VisionTransformer.forward M:259-276) and, with causal != 0, the text transformer's masked attention.
 * qkv bf16 [B*T, 3W] (q | k | v per token, head h = columns h*64..h*64+63 of each third); out bf16 [B*T, W].
 * Only the first `q_rows` query rows of every sequence are computed and stored (q_rows == T: all of them;
 * q_rows == 1: the class token only, used by the dense last block); other rows of `out` are left untouched. */
int lecb_attn_fwd(const void* qkv, void* out, int B, int T, int W, int heads, int q_rows, int causal, void* stream);
/* Share of the softmax's exponentials evaluated on the FMA pipe instead of the MUFU: every n-th score (n in {2, 3, 4}; 0 = none;
 * sequences shorter than 256 tokens always use 0) becomes 2^round(x) * cubic(x - round(x)) (relative error 7.7e-5, 4 % of half a
 * bf16 ulp of the probability).  Default 0: it paid +5-7 % while the kernel still exponentiated the dead rows / columns of
 * ragged tiles, and costs 3-4 % since it does not (DESIGN.md section 3 xviii).  lecb_set_attn_poly returns the previous setting. */
int lecb_set_attn_poly(int n);

/* ---- ViT visual tower row kernels (VisionTransformer.forward, M:259-276) ----
 * lecb_patchify: x NCHW fp32 [B,3,H,W] -> bf16 [B*(H/p)*(W/p), Kpad], column k = c*p*p + py*p + px (the flattening
 *   of conv1.weight [width,3,p,p]), zero beyond 3*p*p: the A operand of the patch-embedding GEMM (M:261).
 * lecb_vit_embed_ln: out fp32 [B*T, D] = ln_pre( [class_embedding ; emb rows of image b] + positional_embedding )
 *   (M:262-266); emb bf16 [B*(T-1), D].
 * lecb_copy_cols: dst[r, 0:cols] = src[r, col0:col0+cols] (bf16): takes the value third of a packed qkv matrix. */
int lecb_patchify(const float* x, void* out, int B, int H, int W, int patch, int Kpad, void* stream);
int lecb_vit_embed_ln(const void* emb, const float* cls, const float* pos, const float* gamma, const float* beta,
                      float* out, int B, int T, int D, float eps, void* stream);
int lecb_copy_cols(const void* src, int64_t ld_src, int col0, void* dst, int64_t ld_dst, int64_t rows, int cols,
                   void* stream);

/* ---- dual-prompt head aggregation (T:456-470 test, T:496-514 train) ----
 * dots fp32 [B*P, ldn]: raw dot products of the UN-normalised local features with the unit prompt
 * features, columns [0,K) = positive, [K,2K) = negative, [2K,3K) = evidence (n_txt == 3);
 * row_sumsq fp32 [B*P] (or NULL if rows are already unit) supplies 1/||feature||;
 * row_mask u8 [B*P] (or NULL): 1 = padded caption token (T:491), skipped.
 * K <= 128 classes.  Map rows of masked tokens are left untouched (the host wrapper zero-fills the maps when a mask is given).
 * -> logits_local fp32 [B,K]; optional neg_map / pos_map fp32 [P,B,K] (3rd/4th return values, T:472;
 * neg_map is post winner-take-all when evidence is used, as in the reference). */
int lecb_head_aggregate(const float* dots, int ldn, const float* row_sumsq, const uint8_t* row_mask,
                        float* logits_local, float* neg_map, float* pos_map, int B, int P, int K, int n_txt,
                        float logit_scale, float spatial_scale, void* stream);

/* ---- global logits (T:448,453-455): out[b,k] = scale * <x_b, tpos_k>, x = g_unit or (g_unit+g_add)/2 ---- */
int lecb_global_logits(const float* g_unit, const float* g_add, const float* tpos, float* out, int B, int D, int K,
                       float scale, void* stream);

/* ---- asymmetric loss forward + backward (U:126-173): loss (scalar, device) and dloss/dlogits ----
 * partial != 0: -sum/B (dualcoop_loss, U:175-181); else -mean (ASL_loss, U:184-190). grad may be NULL.
 * exp / log / reciprocal are the hardware approximations (ex2 / lg2 / rcp.approx.ftz: <= 2 ulp); against the float64 formula the
 * loss agrees to 2e-5 relative and every gradient element to 2e-5 of the largest one (tests/test_train_gpu.py).  gamma_pos = 1,
 * gamma_neg = 2 with thresh_neg <= thresh_pos (the two shipped call sites) run a branch-free variant with one logarithm per
 * element; environment LECB_ASL_GENERAL=1 keeps the general variant for them as well. */
int lecb_asl_fwd_bwd(const float* logits, const float* targets, float* grad, float* loss, int64_t B, int K,
                     float gamma_neg, float gamma_pos, float clip, float eps, float thresh_pos, float thresh_neg,
                     int partial, void* stream);

/* ---- pairwise ranking hinge forward + backward (U:85-93), logits are NOT modified ---- */
int lecb_ranking_fwd_bwd(const float* logits, const float* targets, float* grad, float* loss, int B, int K,
                         float scale, float margin, void* stream);

/* ---- co-occurrence weighted ranking hinge (U:95-110, called at T:842-850): pair (i, j) weighted by pair_weights[i*K + j]
 * (fp32 [K,K]: the normalised log inverse co-occurrence the caller derives from freq_stats.pkl, U:99-102) ---- */
int lecb_ranking_cooc_fwd_bwd(const float* logits, const float* targets, const float* pair_weights, float* grad,
                              float* loss, int B, int K, float scale, float margin, void* stream);

/* ---- EMA consistency term (T:809-813): weight * KLDivLoss(batchmean)(log_softmax(logits), softmax(logits_target)) and
 * its gradient w.r.t. logits, weight * (softmax(logits) - softmax(logits_target)) / B; the target is a constant ---- */
int lecb_kl_softmax_fwd_bwd(const float* logits, const float* logits_target, float* grad, float* loss, int64_t B, int K,
                            float weight, void* stream);

/* ---- distribution-balanced loss: `ResampleLoss` (trainers/dbl.py:263-445; LOSSFUNC 'dbl', T:818-841) with use_sigmoid=True,
 * partial=False: re-balanced weights sigmoid(beta (freq_inv_k / sum_k y_k freq_inv_k - gamma)) + alpha (dbl.py:411-416;
 * freq_inv NULL = no re-weighting), logit regularisation (init_bias fp32 [K] or NULL, neg_scale or 0; dbl.py:401-409),
 * weighted BCE-with-logits averaged over B*K, and the optional focal factor balance_param (1 - e^-L0)^gamma on the two
 * scalar means (dbl.py:373-383).  Writes loss (device scalar) and dloss/dlogits; scratch2 = two floats (focal only). ---- */
int lecb_resample_bce_fwd_bwd(const float* logits, const float* labels, const float* freq_inv, const float* init_bias,
                              float* grad, float* loss, float* scratch2, int64_t B, int K, float map_alpha, float map_beta,
                              float map_gamma, float neg_scale, int focal, float focal_gamma, float balance_param,
                              float loss_weight, void* stream);

/* ---- multi-tensor updates over the prompt-learner parameter list: HOST arrays of `count` (<= 16) device pointers and
 * element counts, one launch each ----
 * lecb_ema_update: twin_i <- twin_i * momentum + live_i * one_minus_momentum (_momentum_update, T:554-559; the caller
 *   passes `1. - momentum` evaluated in double and rounded once, as Python / ATen do: the update is bit-identical)
 * lecb_pack_f32: flat <- concat_i src_i (a NULL source contributes zeros: a parameter without gradient, what DDP's
 *   find_unused_parameters=True covers at T:787); lecb_unpack_scale_f32: dst_i <- scale * its slice of flat — the flat
 *   gradient bucket that is all-reduced once per step (T:786-787)
 * lecb_sgd_step: p_i <- p_i - lr * (buf_i <- momentum * buf_i + grad_scale * flat_i + weight_decay * p_i), torch.optim.SGD
 *   with zero-initialised momentum buffers (the optimiser built at T:773) */
int lecb_ema_update(const float* const* live, float* const* twin, const long long* n, int count, float momentum,
                    float one_minus_momentum, void* stream);
int lecb_pack_f32(const float* const* src, const long long* n, int count, float* flat, void* stream);
int lecb_unpack_scale_f32(const float* flat, float* const* dst, const long long* n, int count, float scale, void* stream);
int lecb_sgd_step(const float* flat_grad, float* const* params, float* const* momentum_buf, const long long* n, int count,
                  float grad_scale, float lr, float momentum, float weight_decay, void* stream);

/* ---- caption retrieval (T:444-448) ----
 * Fused path: lecb_split_f16_hilo writes the exact fp16 pair q = hi + lo of the fp32 queries as one [B, 2D] operand;
 * lecb_gemm_topk10 multiplies both halves against the fp16 bank [N, D] on tcgen05 (each bank tile is staged once per
 * query block and used for both halves) and keeps, per query row and per CTA, the ten largest similarities in registers:
 * part_val / part_idx [B][slots][10] (slots >= *slots_used = the grid size; reset inside the call) — the [B, N] fp32
 * similarity matrix of T:445 (225 MB at 256 x 220 000) is never written; lecb_topk10_merge reduces the partial lists to the
 * row's top-10 (descending; of equal values the lower index first); lecb_gather_mean10 averages the selected rows (T:447).
 * N is arbitrary (>= 10).  The unfused helpers (lecb_split_f16 + two lecb_gemm_bf16 with LECB_GEMM_F16_OPERANDS into an
 * explicit similarity matrix + lecb_topk10) remain as the cross-check. */
int lecb_split_f16_hilo(const float* x, void* out, int64_t rows, int D, void* stream);
int lecb_gemm_topk10(const void* A_hilo, const void* bank, int64_t M, int N, int K, float* part_val, int* part_idx, int slots,
                     int* slots_used, void* stream);
int lecb_topk10_merge(const float* part_val, const int* part_idx, int slots, int B, float* out_val, int* out_idx,
                      void* stream);
int lecb_split_f16(const float* x, void* hi, void* lo, int64_t n, void* stream);   /* x = hi + lo, fp16 each */
int lecb_topk10(const float* sim, int64_t ld, int B, int N, float* out_val, int* out_idx, void* stream);
int lecb_gather_mean10(const void* bank, int bank_is_f16, const int* idx, float* out, int B, int D, void* stream);

/* ---- prompt-tuning backward (T:473-545 + loss.backward(); all CLIP weights frozen, T:763-765) ----
 * data-gradient kernels only: the dgrad GEMMs are lecb_gemm_bf16 against pre-transposed weights. */
int lecb_quick_gelu_fwd(const void* v, void* u, int64_t n, void* stream);                 /* bf16, M:202-204 */
int lecb_quick_gelu_bwd(const void* du, const void* v, void* dv, int64_t n, void* stream);
/* dx = dx_in + dLN/dx(dy); dy, x fp32 [rows,D]; outputs fp32 and/or bf16 (M:193-199, weights frozen) */
int lecb_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                       const float* dx_in, float* dx_f32, void* dx_bf16, int64_t rows, int D, void* stream);
/* causal attention backward: qkv, dqkv bf16 [N*L,3W]; dout bf16 [N*L,W]; L <= 96 (M:221-223) */
int lecb_causal_attn_bwd(const void* qkv, const void* dout, void* dqkv, int N, int L, int W, int heads, void* stream);
/* the same on tcgen05 tensor cores (five M=128 MMAs per (sequence, head), transposed operands read in place through
 * MN-major descriptors); L <= 128.  Probabilities and dS are rounded to bf16 for the MMAs, like lecb_attn_fwd. */
int lecb_attn_causal_bwd(const void* qkv, const void* dout, void* dqkv, int N, int L, int W, int heads, void* stream);
/* the frozen residual ReLU adapter of the adapter trainer (`x + Adapter(x)`, trainers/Caption_distill_double_adapter.py:304-317,
 * :109; its two bias-free linears are lecb_gemm_bf16): out = x + max(z, 0) on fp32; dz = dy * 1[z > 0] rounded to bf16 */
int lecb_residual_relu_fwd(const float* x, const float* z, float* out, int64_t n, void* stream);
int lecb_relu_bwd(const float* dy, const void* z, int z_is_bf16, void* dz_bf16, int64_t n, void* stream);
/* backward of y = x/||x|| on fp32 rows (T:487-488,503) */
int lecb_l2norm_bwd(const float* x, const float* dy, float* dx, int64_t rows, int D, void* stream);
/* gradient of logits_local w.r.t. the raw dot products (same operands as lecb_head_aggregate; T:496-514) */
int lecb_head_aggregate_bwd(const float* dots, int ldn, const float* row_sumsq, const uint8_t* row_mask,
                            const float* grad_local, float* d_dots, int B, int P, int K, int n_txt, float logit_scale,
                            float spatial_scale, void* stream);
/* out[J,D] (+)= alpha * sum_r a[r,J] * b[r,D]: prompt-feature gradients (a fp32 [R,lda], b bf16|fp32 [R,D]) */
int lecb_tn_gemm_small(const float* a, int lda, const void* b, int b_is_bf16, float* out, int R, int J, int D,
                       float alpha, int accumulate, void* stream);

/* ---- test-time score fusion around the scoring path ----
 * lecb_block_fuse: data fp32 [B,NB,K] per-window scores -> out[b,k] = weight * s_ag + base[b,k] (base may be NULL),
 *   s_ag = max_n d if max_n d > threshold else min_n d.  mode 0: d = data (Caption_distill_double.py:655-662);
 *   mode 1 / 2: windows re-weighted by 1 + mean(sims[b,n,:]) and 1 + unbiased var_k as in gen_final_ans.py `fuse`
 *   (18-36) / `fuse6` (38-71); sims fp32 [B,NB,sims_ld].
 * lecb_cooc_adjust: out = pred + weight * pred @ P, P fp32 [K,K] row-normalised co-occurrence (T:611-618). */
int lecb_block_fuse(const float* data, const float* sims, int sims_ld, const float* base, float* out, int B, int NB,
                    int K, int mode, float threshold, float weight, void* stream);
int lecb_cooc_adjust(const float* pred, const float* P, float* out, int B, int K, float weight, void* stream);

/* ---- test-time window pipeline, host side: Pillow-compatible resampling taps ----
 * The reference resizes the image and every sliding window with PIL (transforms.py:384-394 via data_manager.py:348-492).
 * lecb_resize_plan fills, for resizing one axis of in_size pixels to out_size, bounds int32 [out_size][2] = (first source
 * index, tap count) and coeffs int32 [out_size][ksize] = the taps in 22-bit fixed point (zero padded), exactly as Pillow's
 * 8-bit resampler computes them (Resample.c precompute_coeffs / normalize_coeffs_8bpc): an output byte is then
 * clamp((2^21 + sum_t coeffs[t] * src[first + t]) >> 22, 0, 255), horizontal pass first.  HOST pointers, no CUDA call.
 * lecb_resize_ksize returns the taps per output pixel a plan needs (> 0), or a negative status. */
#define LECB_RESIZE_BILINEAR 0
#define LECB_RESIZE_BICUBIC 1
int lecb_resize_ksize(int in_size, int out_size, int filter);
int lecb_resize_plan(int in_size, int out_size, int filter, int* bounds, int* coeffs, int ksize);

/* ---- test-time window pipeline, device side: crop + Pillow-compatible resize + ToTensor + Normalize of `n` windows of
 * ONE decoded image (data_manager.py:348-492 `_transform_image` + transforms.py:379-411), bit-exact against Pillow ----
 * wins: HOST int32 [n][6] = (top, left, height, width, pad_top, pad_bottom) per window (lecb200.windows.Window: group-1
 *   rows counted through the reflect padding / crop of the image, F.pad quirk included).
 * lecb_window_plan_size -> ints of the plan blob and bytes of the [height, S, 3] intermediates;
 * lecb_window_plan fills the HOST blob (window records + both axes' lecb_resize_plan arrays); the caller uploads it.
 * lecb_crop_resize_u8: img uint8 [H,W,3], plan and tmp in device memory -> out_u8 [n,S,S,3] (NHWC, the operand of
 *   lecb_stem_conv1_u8) and / or out_f32 [n,3,S,S] = ((v / 255) - mean) / std (the reference's tensor); mean / std: HOST
 *   float[3].  Two launches (horizontal pass rounded to uint8, then vertical, like Resample.c). */
int lecb_window_plan_size(const int* wins, int n, int H, int W, int S, int filter, long long* plan_ints,
                          long long* tmp_bytes);
int lecb_window_plan(const int* wins, int n, int H, int W, int S, int filter, int* plan, long long plan_ints);
int lecb_crop_resize_u8(const uint8_t* img, int H, int W, const int* plan, int n, int S, uint8_t* tmp, uint8_t* out_u8,
                        float* out_f32, const float* mean, const float* stdv, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LECB_H_ */
