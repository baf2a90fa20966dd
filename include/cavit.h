/* cavit.h — C ABI of libcavit_sm100a.so, the B200 (sm_100a) cross-attention ViT hot path.
 *
 * The reference (vsahni3/cross-attention-ViT) has no FFI: its boundary is the Python module API
 * (`ModelCross.forward`, /root/reference/model_cross.py:186-212; `ModelVIT.forward`,
 * /root/reference/modelv3.py:123-147), every op of which is a torch/ATen call. Each entry point
 * below replaces the ATen call sites named in its comment; `INTEGRATION.md` shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; all pointers are DEVICE pointers owned by the caller unless marked host;
 *   - enqueue-only on `stream` (a cudaStream_t passed as void*), no host synchronisation, safe
 *     under CUDA-graph capture; the library keeps no pointers past a call except a cache of TMA
 *     descriptors keyed by value;
 *   - return 0 on success, <0 on error (CAVIT_E_*); `cavit_last_error()` gives a thread-local
 *     message. Unsupported shapes / devices are hard errors: there is NO fallback path;
 *   - bf16 = 2-byte bfloat16, row-major, "ld" = leading dimension in ELEMENTS, "gs" = stride
 *     between groups (token streams / fusions) in ELEMENTS.
 */
#ifndef CAVIT_H_
#define CAVIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAVIT_ABI_VERSION 3

enum {
  CAVIT_OK = 0,
  CAVIT_E_BADARG = -1,
  CAVIT_E_UNSUPPORTED_SHAPE = -2,
  CAVIT_E_ARCH = -3,
  CAVIT_E_LAUNCH = -4,
  CAVIT_E_DEVICE = -5
};

/* GEMM epilogues (fused into the tcgen05 kernel's TMEM read-out). */
enum {
  CAVIT_EPI_NONE = 0,          /* out = acc                                             */
  CAVIT_EPI_BIAS = 1,          /* out = acc + bias[n]                                   */
  CAVIT_EPI_BIAS_GELU = 2,     /* aux = u = acc + bias (bf16, pre-activation); out = GELU_erf(u) */
  CAVIT_EPI_BIAS_RESID = 3,    /* out = acc + bias[n] + resid[m][n]   (fp32 residual stream) */
  CAVIT_EPI_GELU_BWD = 4,      /* out = acc * GELU'(aux[m][n])        (fc2 dgrad)        */
  CAVIT_EPI_EMBED = 5,         /* out[row_map(m)][n] = acc + bias[n] + pos[1 + m % Np][n] (patch embedding) */
  CAVIT_EPI_BIAS_RELU = 6,     /* out = max(acc + bias[n], 0)   (nn.TransformerEncoderLayer linear1, modelv2.py:72-78) */
  CAVIT_EPI_RELU_BWD = 7       /* out = acc * [aux[m][n] > 0]   (aux = the stored relu output; linear2 dgrad) */
};

int cavit_abi_version(void);
const char* cavit_last_error(void);
/* 1 if device `dev` is compute capability 10.x (B200), else 0. */
int cavit_device_ok(int dev);
/* Synchronises the device and returns the kernel-side status word (0 = OK; 1xx = a bounded
 * mbarrier wait timed out inside a tcgen05 kernel). `reset` != 0 clears it. */
int cavit_device_status(int reset);
/* Number of kernel launches issued through this library since load (bench.py's gpu_launches). */
long long cavit_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * cavit_gemm — D[g][m][n] = sum_k A_g(m,k) * B_g(n,k), bf16 operands, fp32 accumulation in TMEM.
 * Replaces: every nn.Linear on the path — to_qkv / to_out / FeedForward / wk,wv / patch_to_embedding
 *   forward `addmm`s (/root/reference/model_cross.py:23-26,43-47,81-85,168) and their autograd
 *   dgrad / wgrad `mm`s.
 * Operand storage:  *_mn = 0  stored [rows = M or N][K], K contiguous, ld = row stride;
 *                   *_mn = 1  stored [K][M or N],  M/N contiguous, ld = stride between k-rows.
 *   forward  Y = X W^T      : A = X (mn 0), B = W[N][K]        (mn 0)
 *   dgrad    dX = dY W      : A = dY (mn 0), B = W[Nout][Kin]   (mn 1, reduction over Nout)
 *   wgrad    dW = dY^T X    : A = dY[T][Nout] (mn 1), B = X[T][Kin] (mn 1), reduction over T
 * Requirements: K % 8 == 0, ld % 8 == 0, 16-byte aligned bases. `groups` independent problems with
 * identical shapes are batched in one launch via the *_gs strides.
 * ------------------------------------------------------------------------------------------- */
typedef struct cavit_gemm_args {
  int32_t M, N, K, groups;
  int32_t a_mn, b_mn;
  int32_t epi;        /* CAVIT_EPI_*                      */
  int32_t out_fp32;   /* 0: out is bf16, 1: out is fp32   */
  const void* A; int64_t lda, a_gs;
  const void* B; int64_t ldb, b_gs;
  void* out; int64_t ldo, out_gs;
  const float* bias; int64_t bias_gs;               /* [groups][N] fp32 (NULL allowed for EPI_NONE) */
  const float* resid; int64_t ldr, resid_gs;        /* EPI_BIAS_RESID: fp32 [M][N]; EPI_EMBED: pos [1+Np][N] */
  void* aux; int64_t ldaux, aux_gs;                 /* EPI_BIAS_GELU: bf16 out u; EPI_GELU_BWD: bf16 in u; EPI_RELU_BWD: bf16 in h */
  int32_t accumulate;  /* 1: out (fp32 only) += result */
  int32_t split_k;     /* >1: the K range is split over that many CTAs whose partial tiles are combined with
                          fp32 atomics into `out` (fp32, EPI_NONE only; zeroed by the call unless accumulate).
                          0/1: no split. Used by wgrad, whose reduction axis (tokens) is the long one. */
  int32_t embed_np;    /* EPI_EMBED: patches per sample (row m -> out row (m / Np) * (Np + 1) + 1 + m % Np) */
  /* fp32-tolerance mode (ABI 2): when non-NULL (both or neither) the operands are bf16 hi + lo pairs, A ~ A + A_lo and
   * B ~ B + B_lo (same layout / strides as the hi planes), and D = A B^T + A_lo B^T + A B_lo^T is accumulated in ONE pass
   * over TMEM (three MMAs per product, ~2^-16 relative operand error instead of 2^-9). The reference computes every
   * nn.Linear in fp32 (L.Trainer without precision=, /root/reference/main_mist.py:211-218). */
  const void* A_lo;
  const void* B_lo;
} cavit_gemm_args;
int cavit_gemm(const cavit_gemm_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LayerNorm over the last axis (biased variance, affine), fp32 statistics.
 * Replaces: nn.LayerNorm inside PreNorm and the per-stream final norm
 *   (/root/reference/model_cross.py:14-17,174,203; modelv3.py:21,115) and
 *   `native_layer_norm_backward`.
 * x is fp32 with arbitrary row stride; rows are split into `groups` of `rows_per_group`, group g
 * uses gamma/beta[g][C]. y is bf16 [groups*rows_per_group][C] dense.
 * ------------------------------------------------------------------------------------------- */
int cavit_ln_fwd(const float* x, int64_t x_row_stride, int64_t x_gs, int32_t rows_per_group, int32_t groups,
                 int32_t C, const float* gamma, const float* beta, float eps, void* y_bf16, float* mean,
                 float* rstd, void* stream);
/* dx[row] = (dresid ? dresid[row] : 0) + LN'(dy)[row]; optionally also written as bf16 (dx_bf16).
 * dgamma/dbeta [groups][C] are fully reduced inside the launch (per-block partial rows in `partials`, the last
 * block of a group sums them in a fixed order: deterministic). `partials` is a fp32 workspace of
 * cavit_ln_bwd_workspace_floats(groups, C) elements that must be ZERO-INITIALISED once (ticket counters live at
 * its end and are left at zero by every launch). dcol (nullable) [groups][C] receives the column sums of dx,
 * i.e. the bias gradient of the Linear layer whose output this LayerNorm normalised (out-proj / fc2 bias of
 * model_cross.py:45,26). dx may alias dresid.                                                                  */
size_t cavit_ln_bwd_workspace_floats(int32_t groups, int32_t C);
int cavit_ln_bwd(const void* dy_bf16, const float* x, int64_t x_row_stride, int64_t x_gs, const float* mean,
                 const float* rstd, const float* gamma, int32_t rows_per_group, int32_t groups, int32_t C,
                 const float* dresid, float* dx, int64_t dx_row_stride, int64_t dx_gs, void* dx_bf16,
                 float* dgamma, float* dbeta, float* dcol, float* partials, void* stream);

/* Same as cavit_ln_fwd with an additional (or only) fp32 copy of the normalised rows: in the post-norm
 * nn.TransformerEncoderLayer of `ViT3D` (/root/reference/modelv2.py:72-78) the LayerNorm output IS the next
 * residual stream, which stays fp32. Either of y_bf16 / y_f32 may be NULL, not both. Also the eps = 1e-6 norms of
 * /root/reference/model.py:179-180,202 whose CLS rows feed the fp32 BCE tail. */
int cavit_ln_fwd_dual(const float* x, int64_t x_row_stride, int64_t x_gs, int32_t rows_per_group, int32_t groups,
                      int32_t C, const float* gamma, const float* beta, float eps, void* y_bf16, float* y_f32,
                      float* mean, float* rstd, void* stream);
/* cavit_ln_bwd whose incoming gradient is fp32 (the gradient of the post-norm residual stream itself). */
int cavit_ln_bwd_f32(const float* dy_f32, const float* x, int64_t x_row_stride, int64_t x_gs, const float* mean,
                     const float* rstd, const float* gamma, int32_t rows_per_group, int32_t groups, int32_t C,
                     const float* dresid, float* dx, int64_t dx_row_stride, int64_t dx_gs, void* dx_bf16,
                     float* dgamma, float* dbeta, float* dcol, float* partials, void* stream);

/* Fused gather + LayerNorm for the cross-modal fusion input `cat(cls_i, patches_j)`
 * (/root/reference/model_cross.py:140, 109): row 0 of every sample is read from x_cls[k][b]
 * (the saved CLS row of stream cls_src[k]), rows 1.. from stream tok_src[k]; fusion k uses
 * gamma/beta[k]. streams: fp32 [M][B*N][C]; x_cls: fp32 [K][B][C]; y: bf16 [K][B*N][C].
 * cls_src/tok_src are HOST int arrays of length K (K <= 16). */
int cavit_ln_fusion_fwd(const float* streams, int64_t stream_gs, const float* x_cls, int32_t B, int32_t N,
                        int32_t C, int32_t K, const int32_t* cls_src, const int32_t* tok_src, const float* gamma,
                        const float* beta, float eps, void* y_bf16, float* mean, float* rstd, void* stream);
/* Backward of the above: adds LN'(dy) into the fp32 stream gradients with atomics: row 0 into
 * stream cls_src[k], rows 1.. into stream tok_src[k] (a stream can donate patches to several
 * fusions). dy_cls (optional, fp32 [K][B][C]) is added to dy on the CLS rows (the query path). */
int cavit_ln_fusion_bwd(const void* dy_bf16, const float* dy_cls, const float* streams, int64_t stream_gs,
                        const float* x_cls, const float* mean, const float* rstd, const float* gamma, int32_t B,
                        int32_t N, int32_t C, int32_t K, const int32_t* cls_src, const int32_t* tok_src,
                        float* dstreams, float* dgamma, float* dbeta, float* partials, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Self-attention core softmax(Q K^T * scale) V, head_dim 64, fused online softmax on tcgen05;
 * the N x N score matrix never leaves the SM.
 * Replaces: Attention.forward's permute / bmm / softmax / bmm / permute chain
 *   (/root/reference/model_cross.py:53-60, modelv3.py:59-66) and its autograd.
 * qkv: bf16 [G][B*N][3*C] packed exactly as to_qkv writes it (q | k | v thirds, each (h d));
 * out: bf16 [G][B*N][C] merged heads ('b h n d -> b n (h d)'); lse: fp32 [G][B][H][N].
 * ------------------------------------------------------------------------------------------- */
int cavit_attn_fwd(const void* qkv, void* out, float* lse, int32_t G, int32_t B, int32_t N, int32_t H,
                   float scale, void* stream);
/* dqkv: bf16 [G][B*N][3*C]. `delta` fp32 [G][B][H][N] workspace (rowsum(dO*O)).
 * `dq_acc` fp32 [G][B*N][C] workspace (zero-initialised by the call). */
int cavit_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                   float* delta, float* dq_acc, int32_t G, int32_t B, int32_t N, int32_t H, float scale,
                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * Single-query cross attention (L_q = 1): one CLS query per (sample, head) against N keys/values.
 * Replaces: CrossAttention.forward's matmul / softmax / matmul (/root/reference/model_cross.py:95-99).
 * q: fp32 [K][B][C]; kv: bf16 [K][B*N][2*C] (k | v halves); out: fp32 [K][B][C];
 * probs: fp32 [K][B][H][N] saved for backward.
 * -------------------------------------------------------------------------------------------
 * p_drop > 0 applies the attention-probability dropout of CrossAttention (attn_drop, model_cross.py:84,97)
 * with the counter-based mask of cavit_dropout (site / seed as there); probs are saved pre-dropout.
 */
int cavit_xattn_fwd(const float* q, const void* kv, float* out, float* probs, int32_t K, int32_t B, int32_t N,
                    int32_t H, float scale, float p_drop, const uint64_t* seed_dev, uint32_t site, void* stream);
/* dq: fp32 [K][B][C]; dkv: bf16 [K][B*N][2*C]. */
int cavit_xattn_bwd(const float* q, const void* kv, const float* probs, const float* dout, float* dq,
                    void* dkv, int32_t K, int32_t B, int32_t N, int32_t H, float scale, float p_drop,
                    const uint64_t* seed_dev, uint32_t site, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Folded single-query cross attention (csrc/xfold.cu). Replaces, for the engine's fusion path,
 * LayerNorm(cat(cls_i, patches_j)) + wk / wv projections of all N tokens + matmul / softmax / matmul of
 * CrossAttention.forward (/root/reference/model_cross.py:88-99, 111-112) and their autograd, by one pass over
 * the fp32 token streams:  q'_h = Wk_h^T q_h is projected on the host side by a GEMM against the head-expanded
 * weight (cavit_expand_heads), z_h = gamma o (sum_n p_n xhat_n) + beta comes back and o = Wv z + bv is again a GEMM.
 *   x: fp32 [M][B*N][C] token streams; cls: fp32 [K][B][C] (row 0 of each fused sequence);
 *   qp / gz / zhat / dqp: fp32 [K][B][H][C]; z: bf16 [K][B][H][C]; probs: fp32 [K][B][H][N];
 *   mean / rstd: fp32 [K][B][N]; scratch: cavit_xfold_scratch_floats(K, B, N, H) floats;
 *   cls_src / tok_src: HOST arrays of K stream indices.
 * Backward accumulates atomically into dx (rows n >= 1 of stream tok_src[k]; row 0 of stream cls_src[k]),
 * dgamma and dbeta ([K][C]) and stores dqp. p_drop must be 0: with attention dropout the probabilities no longer sum
 * to one and the fold is not exact (CAVIT_E_UNSUPPORTED_SHAPE; use cavit_xattn_* then).
 * ------------------------------------------------------------------------------------------- */
int64_t cavit_xfold_scratch_floats(int32_t K, int32_t B, int32_t N, int32_t H);
/* Variant of the folded kernels in bf16 mode (forward: z_lo == NULL; backward: exact_fp32 == 0): 1 (default, or environment
 * CAVIT_XFOLD_TC=0 to start with 0) = the contractions run on tcgen05 from one bf16 xhat tile in shared memory when N <= 256,
 * C % 128 == 0, C <= 512 (backward: 384) and the tile fits (xhat in bf16 is what the reference-equivalent unfolded route feeds
 * its K / V GEMMs; fp32 accumulation); 0 = the all-fp32 CUDA-core kernels (always used in the fp32 mode and for other shapes).
 * on < 0 only queries. Returns the previous setting; process-wide. */
int cavit_xfold_tensor_cores(int on);
int cavit_xfold_fwd(const float* x, const float* cls, const float* qp, const float* gamma, const float* beta,
                    float* zhat, void* z, void* z_lo /* nullable: second bf16 plane of z, fp32-tolerance mode (ABI 2) */,
                    float* probs, float* mean, float* rstd, float* scratch, int32_t K,
                    int32_t B, int32_t N, int32_t C, int32_t H, const int32_t* cls_src, const int32_t* tok_src,
                    float scale, float eps, float p_drop, const uint64_t* seed_dev, uint32_t site, void* stream);
int cavit_xfold_bwd(const float* x, const float* cls, const float* qp, const float* gamma, const float* zhat,
                    const float* probs, const float* mean, const float* rstd, const float* gz, float* scratch,
                    float* dx, float* dqp, float* dgamma, float* dbeta, int32_t K, int32_t B,
                    int32_t N, int32_t C, int32_t H, const int32_t* cls_src, const int32_t* tok_src, float scale,
                    float p_drop, const uint64_t* seed_dev, uint32_t site, int32_t exact_fp32, void* stream);
/* E[g][c_out][h*C + c_in] = W[g][c_out][c_in] if c_out is a row of head h (c_out / 64 == h) else 0   (bf16)
 * dW[g][c_out][c_in] = dE[g][c_out][(c_out / 64)*C + c_in]                                          (fp32) */
int cavit_expand_heads(const void* W, void* E, int32_t groups, int32_t C, int32_t H, void* stream);
int cavit_fold_heads(const float* dE, float* dW, int32_t groups, int32_t C, int32_t H, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Patch extraction: 'b c (d p1)(h p2)(w p3) -> b (h w d)(p1 p2 p3 c)' gather of one [B, M, 1, D, H, W]
 * fp32 batch into bf16 patch rows [M][B*Np][P] (bit-exact index map).
 * Replaces: einops.rearrange at /root/reference/model_cross.py:193, modelv3.py:129.
 * ------------------------------------------------------------------------------------------- */
/* sample_major = 0: rows ordered [m][b][t] (ModelCross, one token stream per modality);
 * sample_major = 1: rows ordered [b][m][t] (ModelVIT, streams concatenated on the token axis). */
int cavit_patchify(const float* img, void* patches_bf16, int32_t B, int32_t M, int32_t D, int32_t H, int32_t W,
                   int32_t dp, int32_t hp, int32_t wp, int32_t sample_major, void* stream);
/* tokens[m][b*N + 0][:] = cls + pos[0]  for all streams/samples (/root/reference/model_cross.py:195-197). */
int cavit_cls_rows(const float* cls, const float* pos, float* tokens, int32_t M, int32_t B, int32_t N,
                   int32_t C, void* stream);
/* d(pos)[n][c] = sum_{m,b} dtokens[m][b*N+n][c];  d(cls)[c] = sum_{m,b} dtokens[m][b*N][c]. */
int cavit_embed_param_grads(const float* dtokens, float* dpos, float* dcls, int32_t M, int32_t B, int32_t N,
                            int32_t C, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Token assembly and tails of the CNN-stem encoders (csrc/enc.cu).
 * ------------------------------------------------------------------------------------------- */
/* tokens[b][off + s][c] = feat[b][c][s] + pos[off + s][c], off = has_cls; row 0 = cls + pos[0] when has_cls.
 * Replaces flatten / cat / transpose / cls cat / `x + self.pos_embed` of ViT3D.forward
 * (/root/reference/modelv2.py:203-224). feat: fp32 [B][C][S] (the per-modality stem outputs concatenated on the
 * token axis); pos: fp32 [off + S][C]; tokens: fp32 [B][off + S][C]. */
int cavit_tokens_from_channels(const float* feat, const float* cls, const float* pos, float* tokens, int32_t B,
                               int32_t C, int32_t S, int32_t has_cls, void* stream);
/* Adjoint: dfeat[b][c][s] = dtokens[b][off + s][c] (d(pos), d(cls) come from cavit_embed_param_grads). */
int cavit_tokens_to_channels(const float* dtokens, float* dfeat, int32_t B, int32_t C, int32_t S, int32_t has_cls,
                             void* stream);
/* Conv3d(kernel = stride = grid) patch embedding as a permutation into GEMM rows
 * (/root/reference/model.py:84,95-100: `patch_embed`, `flatten(-3)`, `transpose(-2,-1)`, modality concat of :258).
 * feat: fp32 [M*B][Cin][A][Bd][Cd], sample index m*B + b; rows: bf16 [(b*M + m)*Np + t][P] with
 * t = (a'*Bn + b')*Cn + c' (conv output order) and P index ((cin*g0 + i0)*g1 + i1)*g2 + i2 (= Conv3d weight
 * flattened over its last four axes). Voxels beyond the last full patch are ignored (zero gradient). */
int cavit_conv_patch_rows(const float* feat, void* rows_bf16, int32_t M, int32_t B, int32_t Cin, int32_t A, int32_t Bd,
                          int32_t Cd, int32_t g0, int32_t g1, int32_t g2, void* stream);
int cavit_conv_patch_rows_bwd(const void* drows_bf16, float* dfeat, int32_t M, int32_t B, int32_t Cin, int32_t A,
                              int32_t Bd, int32_t Cd, int32_t g0, int32_t g1, int32_t g2, void* stream);
/* Mean over the token axis and its adjoint: the head input of `ViT3D(add_cls_token=False)`
 * (/root/reference/modelv2.py:233-235, `x.mean(dim=1)`). x / dx: fp32 [B][N][C]; out / dmean: fp32 [B][C]. */
int cavit_token_mean_fwd(const float* x, float* out, int32_t B, int32_t N, int32_t C, void* stream);
int cavit_token_mean_bwd(const float* dmean, float* dx, int32_t B, int32_t N, int32_t C, void* stream);
/* Single-logit tail of `ViT` (/root/reference/model.py:224,279-286): logits[b] = x[b].w + b0,
 * loss = BCEWithLogitsLoss(mean)(logits, targets). x: fp32 [B][C]; targets fp32 [B] (NULL: logits only).
 * Backward: dx fp32 [B][C], dw [C], db [1]; upstream gradient as in cavit_head_loss_bwd. */
int cavit_bce_head_fwd(const float* x, const float* w, const float* b0, const float* targets, float* logits,
                       float* loss, int32_t B, int32_t C, void* stream);
int cavit_bce_head_bwd(const float* x, const float* w, const float* targets, const float* logits, float loss_scale,
                       const float* loss_scale_dev, float* dx, float* dw, float* db, int32_t B, int32_t C,
                       void* stream);

/* ---------------------------------------------------------------------------------------------
 * Small utilities.
 * ------------------------------------------------------------------------------------------- */
/* dst_bf16[i] = (bf16) src[i] */
int cavit_cast_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);
/* out[g][c] (+)= sum_r x[g][r][c];  x bf16 [groups][rows][C] (bias gradients). */
int cavit_colsum_bf16(const void* x, int64_t ldx, int64_t x_gs, int32_t rows, int32_t C, int32_t groups,
                      float* out, int64_t out_gs, void* stream);
/* Strided row copy / gather: dst[g][r][:] (+)= src[g][r][:C] (fp32), independent row / group strides.
 * zero_src != 0 additionally clears the source rows ("move"). */
int cavit_gather_rows_f32(float* src, int64_t src_row_stride, int64_t src_gs, float* dst,
                          int64_t dst_row_stride, int64_t dst_gs, int32_t rows, int32_t C, int32_t groups,
                          int32_t accumulate, int32_t zero_src, void* stream);
/* Same with HOST index arrays (length `groups` <= 16, NULL = identity): group g reads source group src_group[g] and
 * writes destination group dst_group[g] — the CLS rows of the K fusions live in K different token streams
 * (/root/reference/model_cross.py:136-142), one launch moves all of them. Destination groups must be distinct. */
int cavit_gather_rows_f32_indexed(float* src, int64_t src_row_stride, int64_t src_gs, const int32_t* src_group,
                                  float* dst, int64_t dst_row_stride, int64_t dst_gs, const int32_t* dst_group,
                                  int32_t rows, int32_t C, int32_t groups, int32_t accumulate, int32_t zero_src,
                                  void* stream);
/* out[i] = a[i] + (float) b_bf16[i]   (residual add for the heads == 1 case where the reference's
 * attention has no output projection, /root/reference/model_cross.py:37,44-48). out may alias a. */
int cavit_add_bf16_f32(const float* a, const void* b_bf16, float* out, int64_t n, void* stream);
/* du[i] = dh[i] * GELU'(u[i])  (bf16 in/out; the classification head's GELU backward). */
int cavit_gelu_bwd_bf16(const void* dh, const void* u, void* du, int64_t n, void* stream);
/* Compact the patch rows of a token tensor (drop the CLS row of every sample):
 * out[s*Np + t][:] = in[s*(Np+1) + 1 + t][:], bf16, s in [0, S). (embedding wgrad operand) */
int cavit_compact_patch_rows_bf16(const void* in, void* out, int32_t S, int32_t Np, int32_t C, void* stream);

/* Classification tail: logits = mean_m(h_m W2_m^T + b2_m); loss = CE(logits, labels, smoothing).
 * Replaces: mlp_head[*][3], torch.mean, F.cross_entropy (/root/reference/model_cross.py:181,205-211).
 * h: bf16 [M][B][F]; W2: fp32 [M][classes][F]; b2: fp32 [M][classes]; labels: int64 [B].
 * logits fp32 [B][classes]; loss fp32 [1]. Backward: dh bf16 [M][B][F], dW2, db2 (overwritten);
 * the upstream gradient of the loss is loss_scale * (loss_scale_dev ? *loss_scale_dev : 1), the
 * device scalar lets autograd's d(loss) be consumed without a host synchronisation. */
int cavit_head_loss_fwd(const void* h, const float* W2, const float* b2, const int64_t* labels, float* logits,
                        float* loss, int32_t M, int32_t B, int32_t F, int32_t classes, float smoothing,
                        float p_drop, const uint64_t* seed_dev, uint32_t site, void* stream);
int cavit_head_loss_bwd(const void* h, const float* W2, const int64_t* labels, const float* logits,
                        float loss_scale, const float* loss_scale_dev, void* dh, float* dW2, float* db2, int32_t M,
                        int32_t B, int32_t F, int32_t classes, float smoothing, float p_drop,
                        const uint64_t* seed_dev, uint32_t site, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused Adam step over the flat parameter / gradient slabs (the same slabs the data-parallel all-reduce uses).
 * Replaces: torch.optim.Adam(self.parameters(), lr, weight_decay).step() of configure_optimizers
 * (/root/reference/model_cross.py:276-279, modelv3.py:211-214, modelv2.py:280-281, model.py:322-323):
 *   g = grad_scale * grad + weight_decay * p;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
 *   p -= lr / (1 - b1^step) * m / (sqrt(v) / sqrt(1 - b2^step) + eps)
 * params_bf16 (nullable) receives the refreshed bf16 operand copy in the same pass. n % 4 == 0, step >= 1. */
int cavit_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, void* params_bf16, int64_t n,
                    float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                    void* stream);

/* ---------------------------------------------------------------------------------------------
 * Dropout (nn.Dropout on the path: model_cross.py:25,27,47,84,86,170,180,182). Masks are counter-based:
 * keep(i) = hash(*seed_dev, site, i) >= p * 2^32, a pure function of the per-step device seed, the
 * dropout module ("site") and the element index, so backward regenerates the forward mask instead
 * of storing it. m = keep / (1 - p).
 *   mode 0: out_f32  = m * a_f32              mode 1: out_bf16 = m * a_bf16
 *   mode 2: out_f32  = a_f32 + m * b_bf16     (residual add of a dropped branch)
 *   mode 3: out_bf16 = bf16(m * a_f32)        (masked gradient operand)
 *   mode 4: out_u8   = keep                   (mask export for tests)
 * out may alias a. The reference's ATen Philox stream cannot be reproduced bit-for-bit;
 * parity with p > 0 is checked by replaying these masks in the CPU oracle.
 * ------------------------------------------------------------------------------------------- */
int cavit_dropout(int32_t mode, const void* a, const void* b, void* out, int64_t n, float p,
                  const uint64_t* seed_dev, uint32_t site, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Input staging for stored volumes.
 * Replaces, per batch, the host side of the deterministic transform chain of BrainDataset
 * (/root/reference/dataset_ucsf.py:81-89 train without augmentation, :121-134 test, :157-158 .to(torch.float)):
 * nibabel's read scaling (stored * scl_slope + scl_inter in float64), MONAI's ResizeWithPadOrCropd(img_size,
 * constant_values = pad_value) and the conversion to a C-contiguous fp32 tensor.
 *   raw   device copy of the stored voxel bytes of all volumes (native byte order, file order: axis 0 fastest);
 *   desc  device array of `volumes` descriptors; byte_offset must be a multiple of the voxel size;
 *   out   fp32 [volumes][D][H][W] (the [B][M][1][D][H][W] model input with volumes = B * M, sample-major).
 * Axis a of the stored volume maps to axis a of (D, H, W): centre crop where dims[a] > target (start dims[a]/2 - target/2),
 * symmetric pad where smaller (before = (target - dims[a]) / 2). slope == 1 and inter == 0 means "take stored values"
 * (the caller maps nibabel's invalid-slope rule, slope 0 / NaN / inf, to that pair). */
enum { CAVIT_VOX_U8 = 0, CAVIT_VOX_I16 = 1, CAVIT_VOX_I32 = 2, CAVIT_VOX_F32 = 3, CAVIT_VOX_F64 = 4, CAVIT_VOX_I8 = 5,
       CAVIT_VOX_U16 = 6, CAVIT_VOX_U32 = 7 };
typedef struct cavit_volume_desc {
  int64_t byte_offset; /* of the volume's first voxel inside `raw` */
  int32_t dims[3];     /* stored extents, axis 0 fastest */
  int32_t dtype;       /* CAVIT_VOX_* */
  float slope, inter;  /* scl_slope, scl_inter after the validity rule */
} cavit_volume_desc;   /* 32 bytes */
int cavit_stage_volumes(const void* raw, const cavit_volume_desc* desc, float* out, int32_t volumes, int32_t D, int32_t H,
                        int32_t W, float pad_value, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Per-step classification metrics, accumulated on the device.
 * Replaces log_stats (/root/reference/model_cross.py:243-255: argmax, compute_metrics of /root/reference/utils.py:18-62,
 * softmax + torchmetrics.functional.auroc) and the train_loss / val_loss logging of :262-273, which cost one host
 * synchronisation per metric per step; Lightning's on_epoch mean weights each batch by its size.
 *   accum[0..7] += B * (accuracy, precision, recall, specificity, F1, NPV, AUROC, *loss);  accum[8] += B;  accum[9] += 1
 * accum holds 16 doubles, zero-initialised by the caller: [10], [11] are scratch words the launch leaves at zero
 * (cross-block pair count and ticket), [12..15] are reserved.
 * logits fp32 [B][2], labels int64 [B] (non-zero = positive), loss nullable device scalar; 0 / 0 = 0; AUROC = 0 when the batch
 * holds a single class (torchmetrics' convention). 1 <= B <= 8192. */
int cavit_batch_metrics(const float* logits, const int64_t* labels, const float* loss, double* accum, int32_t B,
                        int32_t classes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K-EMBED (csrc/embed.cu): patch extraction as a TMA-staged unfold FUSED with the embedding GEMM and the positional
 * add — the unfolded [B, Np, P] patch tensor never exists. Replaces einops.rearrange + patch_to_embedding + cat(cls) +
 * `x += pos_embedding` of /root/reference/model_cross.py:189-198 and modelv3.py:125-140 (forward) and the
 * patch_to_embedding weight gradient of their autograd (the input volumes need no gradient).
 *   img: fp32 [B][M][1][D][H][W];  W: fp32 [C][P] (the master weights), P = dp*hp*wp, feature order (a, b, c);
 *   bias fp32 [C];  pos: fp32 [Ntok][C];  tokens: fp32, sample_major = 0: [M][B*(Np+1)][C] (ModelCross),
 *   1: [B*(M*Np+1)][C] (ModelVIT); token t = (hi*Wn + wi)*Dn + di of a volume goes to row 1 + t (+ m*Np) of its sequence —
 *   bit-exact index map.
 * TMA fetches fp32 bricks {c: wp, h: 32/wp rows, wi: W/wp, z: 1, v: volumes} of the volume per 32-feature k-block straight
 * into the 128-byte-swizzled operand stage; tcgen05.mma.kind::tf32 reads the fp32 words as TF32 (10 mantissa bits), fp32
 * accumulation — no conversion pass, no bf16 copy of the weights. The CLS rows are cavit_cls_rows'.
 * *_supported() != 0 iff the geometry is inside the kernels' reach (wp in {8,16,32} with hp % (32/wp) == 0, or wp % 32 == 0;
 * W % 4 == 0, W/wp <= 128, C % 32 == 0; wgrad: wp in {8,16,32,64}, M * W/wp <= 64 and 128-feature tiles made of whole
 * (a, b) rows); otherwise the calls fail with CAVIT_E_UNSUPPORTED_SHAPE and the caller uses cavit_patchify + cavit_gemm.
 * wgrad: dtokens bf16 in the tokens layout, volume bricks staged in fp32 and rounded to bf16 by converter warps (tcgen05
 * reads MN-major TF32 only from 128-byte rows, which 8- / 16-wide patch rows cannot form); dW fp32 [C][P] is overwritten
 * (zeroed, then fp32 red.add of the partial tiles).
 * cavit_embed_bias_grad: db[c] = sum_{n >= 1} dpos[n][c] (dpos from cavit_embed_param_grads).
 * ------------------------------------------------------------------------------------------- */
int cavit_embed_fused_supported(int32_t B, int32_t M, int32_t D, int32_t H, int32_t W, int32_t dp, int32_t hp, int32_t wp,
                                int32_t C);
int cavit_embed_fused_wgrad_supported(int32_t B, int32_t M, int32_t D, int32_t H, int32_t W, int32_t dp, int32_t hp,
                                      int32_t wp, int32_t C);
int cavit_embed_fused_fwd(const float* img, const float* W_f32, const float* bias, const float* pos, float* tokens, int32_t B,
                          int32_t M, int32_t D, int32_t H, int32_t W, int32_t dp, int32_t hp, int32_t wp, int32_t C,
                          int32_t sample_major, void* stream);
int cavit_embed_fused_wgrad(const float* img, const void* dtokens_bf16, float* dW, int32_t B, int32_t M, int32_t D, int32_t H,
                            int32_t W, int32_t dp, int32_t hp, int32_t wp, int32_t C, int32_t sample_major, void* stream);
int cavit_embed_bias_grad(const float* dpos, float* db, int32_t N, int32_t C, void* stream);

/* =============================================================================================
 * fp32-tolerance mode (north_star: ~1e-3 on logits and attention outputs "in fp32"; the reference runs fp32 throughout,
 * /root/reference/main_mist.py:211-218). GEMM operands travel as TWO bf16 planes, x ~ hi + lo with hi = bf16(x),
 * lo = bf16(x - hi); cavit_gemm (A_lo / B_lo) multiplies them with three tensor-core MMAs per product; everything
 * between the GEMMs (attention, GELU, LayerNorm, head, loss) is fp32. The entry points below produce / consume the
 * planes. `n` counts elements.
 * ============================================================================================= */
/* hi / lo planes of an fp32 array (weights once per step, activations that no fused producer splits). */
int cavit_cast_split(const float* src, void* hi, void* lo, int64_t n, void* stream);
/* Exact-erf GELU (nn.GELU(), /root/reference/model_cross.py:24,179) of fp32 pre-activations u: h = gelu(u) as hi / lo
 * planes (NULL pair allowed) and / or as fp32 (NULL allowed); n % 4 == 0. */
int cavit_gelu_split(const float* u, void* h_hi, void* h_lo, float* h_f32, int64_t n, void* stream);
/* du = dh * gelu'(u), fp32 in, hi / lo planes out (autograd of the above). */
int cavit_gelu_bwd_split(const float* dh, const float* u, void* du_hi, void* du_lo, int64_t n, void* stream);
/* cavit_ln_fwd with the normalised rows as hi / lo planes. */
int cavit_ln_fwd_split(const float* x, int64_t x_row_stride, int64_t x_gs, int32_t rows_per_group, int32_t groups,
                       int32_t C, const float* gamma, const float* beta, float eps, void* y_hi, void* y_lo, float* mean,
                       float* rstd, void* stream);
/* cavit_ln_bwd_f32 (fp32 incoming gradient) whose bf16 copy of dx is a hi / lo pair (NULL pair allowed). */
int cavit_ln_bwd_split(const float* dy_f32, const float* x, int64_t x_row_stride, int64_t x_gs, const float* mean,
                       const float* rstd, const float* gamma, int32_t rows_per_group, int32_t groups, int32_t C,
                       const float* dresid, float* dx, int64_t dx_row_stride, int64_t dx_gs, void* dx_hi, void* dx_lo,
                       float* dgamma, float* dbeta, float* dcol, float* partials, void* stream);
/* cavit_colsum_bf16 over hi + lo rows (bias gradients). */
int cavit_colsum_split(const void* x_hi, const void* x_lo, int64_t ldx, int64_t x_gs, int32_t rows, int32_t C,
                       int32_t groups, float* out, int64_t out_gs, void* stream);
/* cavit_patchify with the patch rows as hi / lo planes (raw MRI intensities reach 1.7e4: 8 mantissa bits are not enough). */
int cavit_patchify_split(const float* img, void* patches_hi, void* patches_lo, int32_t B, int32_t M, int32_t D, int32_t H,
                         int32_t W, int32_t dp, int32_t hp, int32_t wp, int32_t sample_major, void* stream);
/* Self-attention softmax(q k^T * scale) v in fp32 on the CUDA cores (head_dim 64), online softmax, no N x N matrix in
 * memory. Replaces Attention.forward's matmul / softmax / matmul and their autograd
 * (/root/reference/model_cross.py:50-61) in the fp32 mode. qkv: fp32 [G][B*N][3C] (q | k | v thirds, each (h d));
 * out / dout: fp32 [G][B*N][C]; lse, delta: fp32 [G][B][H][N] (delta is scratch); dqkv like qkv. */
int cavit_attn_fwd_f32(const float* qkv, float* out, float* lse, int32_t G, int32_t B, int32_t N, int32_t H, float scale,
                       void* stream);
int cavit_attn_bwd_f32(const float* qkv, const float* out, const float* dout, const float* lse, float* dqkv, float* delta,
                       int32_t G, int32_t B, int32_t N, int32_t H, float scale, void* stream);
/* cavit_head_loss_fwd / _bwd with fp32 hidden activations h [M][B][F] (and fp32 dh); no dropout in this mode. */
int cavit_head_loss_fwd_f32(const float* h, const float* W2, const float* b2, const int64_t* labels, float* logits,
                            float* loss, int32_t M, int32_t B, int32_t F, int32_t classes, float smoothing, void* stream);
int cavit_head_loss_bwd_f32(const float* h, const float* W2, const int64_t* labels, const float* logits, float loss_scale,
                            const float* loss_scale_dev, float* dh, float* dW2, float* db2, int32_t M, int32_t B, int32_t F,
                            int32_t classes, float smoothing, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CAVIT_H_ */
