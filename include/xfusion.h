/* xfusion.h — C ABI of libxfusion_sm100a.so: hand-written sm_100a (B200) kernels for the
 * TransFusion cross_fusion hot path (reference: modeling/cross_fusion/ego_fusion/
 * cross_f_box_wrapper.py:165-230, cross_f_box_layers.py:69-108, torch18_adapters.py:108-113,
 * 544-608, 789-798, modeling/cross_fusion/utils.py:35-46,114-119,209-218).
 *
 * The reference has no FFI (100 % Python over ATen); each entry point below replaces the
 * ATen library calls the cited reference lines dispatch.  Conventions:
 *   - the caller owns every buffer (device pointers + explicit shapes/strides); nothing is
 *     allocated, nothing is synchronised: all work is enqueued on the passed stream;
 *   - return 0 = ok, < 0 = invalid argument / unsupported shape (see xf_last_error()),
 *     > 0 = cudaError_t;
 *   - bf16 tensors are row-major with an explicit leading dimension in ELEMENTS;
 *   - no torch types cross this boundary.
 */
#ifndef XFUSION_H_
#define XFUSION_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* xf_stream_t; /* cudaStream_t */

#define XF_ABI_VERSION 2

int xf_version(void);
const char* xf_last_error(void);
/* number of kernel launches issued through this library since load (bench.py: gpu_launches) */
int64_t xf_launch_count(void);
/* TMA descriptor cache (descriptors are keyed by address / extents / pitches / box / swizzle under a mutex): which = 0 hits,
 * 1 misses (descriptors actually encoded) since load */
int64_t xf_tmap_cache_stats(int which);

/* ------------------------------------------------------------------------------------------
 * GEMM  D[M,N] (+)= A(M x K) * B(N x K)^T  on tcgen05 tensor cores (bf16 in, fp32 accumulate
 * in TMEM, TMA-fed, 128 x tile_n x 64 tiles, persistent CTAs), with a fused epilogue.
 * Replaces: nn.Conv2d patch-embed as im2col GEMM (cross_f_box_wrapper.py:268-274,183),
 * F.linear in-proj (torch18_adapters.py:685), out_proj (:608), linear1/linear2 (:111),
 * RegroupPatchesLayerBox.linear (utils.py:116) and their autograd backward (dgrad, wgrad).
 *
 * Operand storage (bf16, row-major):
 *   a_mn_major = 0: A stored [M, K] (K contiguous)      = 1: A stored [K, M] (M contiguous)
 *   b_mn_major = 0: B stored [N, K] (K contiguous)      = 1: B stored [K, N] (N contiguous)
 * so  forward  y = x W^T      : a = x [M,K],  b = W [N,K]           (0,0)
 *     wgrad    dW = dy^T x    : a = dy [tokens, N_out] (1), b = x [tokens, K_in] (1)
 *     dgrad    dx = dy W      : a = dy [M,N_out] (0), b = W [N_out, K_in] stored [K,N] (1)
 *
 * Epilogue order per element (m, n):  v = acc (+ bias[n]) (+ pos_table[m % rows_in, n])
 *   -> [drop_first: dropout] -> act / dact -> [!drop_first: dropout] -> (+ residual[m_out, n])
 *   -> store / atomic-add to out[m_out, n],  m_out = (m / rows_in) * rows_out + m % rows_in + row_off
 *   (rows_in = 0: m_out = m).
 * The common combinations (bf16 out with [bias] [dropout] [residual]; bias + saved pre-activation + GELU
 * [+ dropout]; GELU' [+ dropout mask]; fp32 accumulate for wgrad) run specialised kernels that need
 * N % 8 == 0 and 16-byte aligned operands; everything else takes a generic element-wise epilogue.
 * Dropout: keep(m_out, n) = rowhash(seed, stream, m_out) * colhash(n) >= p * 2^32 (csrc/ptx.cuh) -- the same
 * (seed, stream) reproduces the mask in any kernel of the library (GEMM, LayerNorm, row gather).
 * ------------------------------------------------------------------------------------------ */
typedef struct XfGemm {
  const void* a; int64_t a_ld;
  const void* b; int64_t b_ld;
  int32_t a_mn_major, b_mn_major;
  int64_t M, N, K;
  int32_t tile_n;      /* 0 = auto; else multiple of 32 in [32, 256] */
  int32_t split_k;     /* <= 1: none; > 1 requires out_dtype = 1 and accumulate = 1 */
  const float* bias;        /* [N] fp32 or NULL */
  const float* pos_table;   /* [>= rows_in, N] fp32, row = m % rows_in, or NULL (utils.py:209-214) */
  int64_t rows_in, rows_out, row_off;
  int32_t act;              /* 0 none, 1 GELU erf (F.gelu), 2 ReLU (TwoMLPHead fc6 / fc7, torchvision faster_rcnn.py) */
  void* preact_out;         /* bf16, indexed like out: value before act (saved for backward), or NULL */
  const void* dact_in;      /* bf16, indexed like out, or NULL.  act 0: v *= gelu'(dact_in[m,n]) (GELU backward, dact_in = saved
                               pre-activation); act 2: v = dact_in[m,n] > 0 ? v : 0 (ReLU backward, dact_in = forward output) */
  const void* residual;     /* bf16 [m_out, n], leading dim ldr, or NULL */
  int64_t ldr;
  void* out; int64_t ldc;
  int32_t out_dtype;        /* 0 bf16, 1 fp32 */
  int32_t accumulate;       /* 0 store, 1 fp32 atomic add into out */
  float drop_p;             /* 0 = no dropout */
  uint32_t drop_seed, drop_stream;
  int32_t drop_first;
  int32_t max_ctas;         /* 0 = one per SM */
  int32_t cta_group;        /* 0 = auto (CTA pairs, 256 x tile_n tiles via cta_group::2), 1 = single-CTA 128 x tile_n tiles, 2 = force pairs */
  /* Batched GEMM (batch1 > 0): batch1 x max(batch2, 1) independent problems of the same M, N, K; entry (i, j) reads A at
   * a + i * a_bs1 + j * a_bs2 (elements), B and out likewise (residual / preact_out / dact_in use the out strides).
   * The two-level index lets a per-(sample, head) problem address token-major [B*S, H*dp] tensors: bs1 = S * ld,
   * bs2 = dp.  Rows / columns outside M, N, K are zero-filled per entry (TMA bounds), so neighbouring entries may
   * overlap in memory.  Specialised epilogues only (split_k: the fp32 reduction epilogue, as without batching). */
  int32_t batch1, batch2;
  int64_t a_bs1, a_bs2, b_bs1, b_bs2, out_bs1, out_bs2;
} XfGemm;

int xf_gemm(const XfGemm* g, xf_stream_t stream);
/* Per-host-thread cap on the persistent grid of xf_gemm calls that leave max_ctas = 0 (also the GEMMs xf_attn_bwd issues);
 * 0 removes the cap.  Returns the previous cap.  For callers that run independent problems on several streams and give
 * each a fixed share of the SMs. */
int xf_set_gemm_cta_cap(int ctas);


/* ------------------------------------------------------------------------------------------
 * Layout passes (HBM-bound, shared-memory staged so both sides are coalesced).
 * Token matrix T[(b,i,j), (c,u,v)] <-> feature map F[b,c,i*p+u,j*p+v].
 * xf_patchify replaces the im2col of nn.Conv2d(k=stride=p) (cross_f_box_wrapper.py:268-274) and
 * patchify_image(.,1,1) (utils.py:35-39); xf_fold replaces regroup_patches / F.fold
 * (utils.py:42-46).  Each is the other's backward.  feat_dtype: 0 bf16, 1 fp32; + 4 = the map is stored channels_last
 * (NHWC memory: F[b, h, w, c], torch.channels_last), SURVEY 8f N3 -- a channels_last backbone feeds the patch-embed and receives
 * the fused map without any layout conversion pass.
 * ------------------------------------------------------------------------------------------ */
int xf_patchify(const void* feat, int feat_dtype, void* tok_bf16, int64_t tok_ld, int B, int C, int H, int W, int p,
                xf_stream_t stream);
int xf_fold(const void* tok_bf16, int64_t tok_ld, void* feat, int feat_dtype, int accumulate, int B, int C, int H, int W,
            int p, xf_stream_t stream);

/* z[b, n+j, :] = bf16(lang[b,j,:] + kind[:])   (cross_f_box_layers.py:76,86) and its backward
 * (dlang += dz rows, may be NULL; dkind += column sums). */
int xf_lang_rows_fwd(const float* lang, const float* kind, void* z_bf16, int B, int L, int D, int n, int S, xf_stream_t stream);
int xf_lang_rows_bwd(const void* dz_bf16, float* dlang, float* dkind, int B, int L, int D, int n, int S, xf_stream_t stream);

/* LayerNorm(D, eps, affine) — norm1 / norm2 (torch18_adapters.py:110,113) and final_norm_layer
 * (cross_f_box_layers.py:104-107).  Logical row r maps to x row / y row through the same block
 * remap as the GEMM epilogue ((r / rows_in) * rows_out + r % rows_in + row_off; rows_in = 0: r),
 * so the final LN reads only the visual rows of the [B,S,D] sequence.  Optional dropout on y
 * (RegroupPatchesLayerBox.back_dropout, utils.py:115). */
typedef struct XfLayerNorm {
  const void* x; int64_t ldx;
  void* y; int64_t ldy;
  const float* gamma; const float* beta;
  float* mean; float* rstd;          /* [rows] saved for backward, or NULL */
  int32_t rows, D;
  int32_t in_rows_in, in_rows_out, in_row_off;
  int32_t out_rows_in, out_rows_out, out_row_off;
  float eps;
  float drop_p; uint32_t drop_seed, drop_stream;
} XfLayerNorm;
int xf_layernorm_fwd(const XfLayerNorm* a, xf_stream_t stream);

/* LayerNorm backward: dx, dgamma += , dbeta += , optional dbias += colsum(dx2 or dx) and an optional
 * dropout-masked copy dx2 of dx (gradient entering the linear whose output was dropped out). */
typedef struct XfLayerNormBwd {
  const void* dy; int64_t lddy;
  const void* x; int64_t ldx;
  const float* gamma; const float* mean; const float* rstd;
  void* dx; int64_t lddx;
  void* dx2;
  float* dgamma; float* dbeta; float* dbias;
  int32_t rows, D;
  int32_t in_rows_in, in_rows_out, in_row_off;
  int32_t out_rows_in, out_rows_out, out_row_off;
  float dy_drop_p; uint32_t dy_drop_seed, dy_drop_stream;
  float dx2_drop_p; uint32_t dx2_drop_seed, dx2_drop_stream;
} XfLayerNormBwd;
int xf_layernorm_bwd(const XfLayerNormBwd* a, xf_stream_t stream);

/* out[n] += sum_m x[m,n]  (bias gradients of in_proj / linear1) */
int xf_colsum(const void* x_bf16, int64_t ld, int rows, int cols, float* out, xf_stream_t stream);

/* fp32 -> bf16 weight cast with optional block padding of rows (rin -> rout) / columns (cin -> cout):
 * head_dim 178 -> 192 for the Ego4Dv1 shape; xf_unpad_add is the inverse for weight gradients. */
int xf_cast_pad(const float* src, int64_t lds, void* dst_bf16, int64_t ldd, int rows, int cols, int rin, int rout, int cin,
                int cout, xf_stream_t stream);
int xf_unpad_add(const float* src_padded, int64_t lds, float* dst, int64_t ldd, int rows, int cols, int rin, int rout, int cin,
                 int cout, xf_stream_t stream);

/* dst[i] = scale * float(src_bf16[i]), n % 8 == 0: unpacks a bf16-compressed gradient arena after its all-reduce (and
 * averages it) -- the optional compression of the data-parallel exchange (run_experiment.py:452 runs DDP in fp32). */
int xf_bf16_to_f32(const void* src_bf16, float* dst, int64_t n, float scale, xf_stream_t stream);

/* The same cast for up to XF_CAST_MAX_JOBS tensors in ONE launch (all bf16 weight copies of an FPN level:
 * 18 small tensors whose separate launches were latency-bound).  Replaces the implicit per-module
 * fp32 -> autocast-bf16 weight conversions of ego_fusion/cross_f_box_layers.py:75-103. */
#define XF_CAST_MAX_JOBS 32
typedef struct XfCastJob {
  const float* src; int64_t lds;
  void* dst_bf16;   int64_t ldd;
  int32_t rows, cols, rin, rout, cin, cout;
} XfCastJob;
int xf_cast_pad_multi(const XfCastJob* jobs, int n_jobs, xf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * LM head: PoolPredictor, modeling/cross_fusion/ego_fusion/lm_layers.py:30-81, in fp32.
 *   pool   : pooled[b,d] = mean_l | max_l (tok[b,l,d] * mask[b,l])   (:60-66; mean divides by the padded L;
 *            type 0 = mean, 1 = max; mask uint8 [B,L] 1 = valid, may be NULL; argmax [B,D] only for max)
 *   rowln  : nn.LayerNorm over the rows of a small fp32 [rows, D] matrix (:68-69)
 *   linear : y[r,c] = sum_d act(x[r,d]) W[c,d] + bias[c], act 0 = identity, 1 = GELU(erf) applied to the INPUT
 *            (nn.Sequential(GELU, Linear) :43-45, mlp_noun / mlp_verb :47-55).  The backward ACCUMULATES into
 *            dW / dbias (either may be NULL together with dx to skip that part).
 * ------------------------------------------------------------------------------------------ */
int xf_lm_pool_fwd(const float* tok, const uint8_t* mask, int B, int L, int D, int type, float* pooled, int32_t* argmax,
                   xf_stream_t stream);
int xf_lm_pool_bwd(const float* dpooled, const uint8_t* mask, const int32_t* argmax, int B, int L, int D, int type, float* dtok,
                   xf_stream_t stream);
int xf_rowln_fwd(const float* x, const float* gamma, const float* beta, int rows, int D, float eps, float* y, float* mean,
                 float* rstd, xf_stream_t stream);
int xf_rowln_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd, int rows, int D,
                 float* dx, float* dgamma, float* dbeta, xf_stream_t stream);
int xf_small_linear_fwd(const float* x, const float* W, const float* bias, int rows, int C, int D, int act, float* y,
                        xf_stream_t stream);
int xf_small_linear_bwd(const float* dy, const float* x, const float* W, int rows, int C, int D, int act, float* dx, float* dW,
                        float* dbias, xf_stream_t stream);

/* delta[b, h, s] = sum_e O[b*S+s, h*dp+e] * dO[b*S+s, h*dp+e]  (attention backward pre-pass);
 * delta is [B, H, stat_stride] like the LSE. */
int xf_attn_delta(const void* o_bf16, const void* do_bf16, int64_t ld, int B, int S, int heads, int dp, int stat_stride,
                  float* delta, xf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused multi-head attention (flash-style: no S x S tensor in HBM), tcgen05 + TMA + TMEM.
 * Replaces torch18_adapters.py:544-555 (head split), :578-597 (key_padding_mask -> -inf),
 * :789-798 (_scaled_dot_product_attention) and :607 (head merge); general Sq != Sk, so the
 * QKVEncoder cross-attention (cross_qkv_layers.py:70-77) is the same call.
 * q/k/v/out: bf16, token-major [B*S, ld], head hd in columns [hd*dp, hd*dp+dp) (dp = head dim
 * padded to a multiple of 32, <= 256; pad columns must be zero).  For a fused in-proj output
 * [B*S, 3*H*dp] pass k = q + H*dp, v = q + 2*H*dp with ldq = ldk = ldv = 3*H*dp.
 * ------------------------------------------------------------------------------------------ */
typedef struct XfAttnFwd {
  const void* q; int64_t ldq;
  const void* k; int64_t ldk;
  const void* v; int64_t ldv;
  void* out; int64_t ldo;
  float* lse;                        /* [B,H,lse_stride] fp32, log2-domain logsumexp of the scaled scores (for backward), or NULL */
  int32_t lse_stride;                /* 0 = Sq; the backward wants a multiple of 32 >= Sq */
  const uint8_t* key_padding_mask;   /* [B,Sk] nonzero = ignore (True of src_key_padding_mask), or NULL */
  int32_t kpm_start;                 /* keys < kpm_start are never masked (visual tokens): lets tiles skip the mask */
  int32_t B, H, Sq, Sk, dp;
  float scale;                       /* 1/sqrt(head_dim) */
  float drop_p; uint32_t drop_seed, drop_stream;   /* dropout on the attention probabilities */
  void* debug_timeline;              /* dev aid: NULL, or int64[2][64][8] device buffer of clock64 stamps */
} XfAttnFwd;
int xf_attn_fwd(const XfAttnFwd* a, xf_stream_t stream);

/* Attention backward (recomputes the probabilities from q, k and the saved LSE): dq, dk, dv in the
 * same token-major layout as q, k, v (pad columns come out zero).  Three tcgen05 passes with one TMEM
 * accumulator each: a query-stationary dQ pass and key-stationary dK and dV passes.  dp must be a multiple
 * of 32, <= 224; stat_stride a multiple of 64 >= Sq rounded up to 64. */
typedef struct XfAttnBwd {
  const void* q; int64_t ldq;
  const void* k; int64_t ldk;
  const void* v; int64_t ldv;
  const void* d_out; int64_t lddo;
  const float* lse; const float* delta; int32_t stat_stride;   /* [B,H,stat_stride] each */
  void* dq; int64_t lddq;
  void* dk; int64_t lddk;
  void* dv; int64_t lddv;
  const uint8_t* key_padding_mask;
  int32_t kpm_start;                               /* keys < kpm_start are never masked */
  int32_t B, H, Sq, Sk, dp;
  float scale;
  float drop_p; uint32_t drop_seed, drop_stream;   /* must equal the forward's */
  void* debug_timeline;                            /* dev aid: NULL, or int64[2][2][64][8] device buffer of clock64 stamps */
  /* Optional scratch of >= xf_attn_bwd_workspace_bytes(B, H, Sq, Sk) bytes (16-byte aligned, contents undefined on
   * entry and exit).  With it the backward computes the score tiles ONCE: a key-stationary tcgen05 pass produces dV and
   * streams E = scale * dS^T as bf16 [B, H, Sk, Sq] tiles to the scratch, and dQ = E^T K, dK = E Q run as two batched
   * GEMMs (5 GEMM units in total, the algorithmic minimum).  Without it (NULL): three tcgen05
   * passes that recompute the scores (8 units) and never leave the chip: no S x S bytes in HBM. */
  void* workspace; int64_t workspace_bytes;
} XfAttnBwd;
int xf_attn_bwd(const XfAttnBwd* a, xf_stream_t stream);
int64_t xf_attn_bwd_workspace_bytes(int B, int H, int Sq, int Sk);

/* out[r,:] = in[(r / rin) * rout + r % rin + roff, :] with the dropout mask of the forward write at
 * the source position, colsum += column sums of out (may be NULL).  Patch-embed backward: gathers the
 * visual rows of d(sequence) (cross_f_box_layers.py:72-74 backward; colsum = image_kind_embedding grad). */
int xf_rows_gather(const void* in_bf16, int64_t ldi, void* out_bf16, int64_t ldo, int rows, int D, int rin, int rout, int roff,
                   float* colsum, float drop_p, uint32_t drop_seed, uint32_t drop_stream, xf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * fp32-tolerance mode helpers (csrc/fp32_mode.cu; the reference's Ego4Dv2 config runs precision 32,
 * ego_nao_res50_ego4dv2.yml:124).  xf_split3 writes the 3-term bf16 split of an fp32 [rows, cols] matrix concatenated
 * along K as [rows, 6*cols]: pattern 0 = (a0 a0 a0 a1 a1 a2) for the A side, 1 = (b0 b1 b2 b0 b1 b0) for the B side, so
 * xf_gemm on the two (K' = 6K, fp32 output) returns the fp32-accurate product on the bf16 tensor cores.  act = 1 applies the
 * exact GELU(erf) to the source first; bias_cols = 8 appends eight columns that carry a bias through the GEMM (A side:
 * 1 1 1 0.., B side: the 3-term split of bias[row]), so the split-K fp32-reduction epilogue suffices: split-K matters
 * because the tensor core truncates when it adds into its fp32 accumulator -- a 6K-long chain loses ~1e-5 relative, twelve
 * short chains summed by round-to-nearest reductions ~1e-6.
 * xf_softmax_rows_f32: in-place softmax_k(scale * s[b,h,q,k]) over k < Sk of a padded [B,H,Sq,Sp] fp32 score tensor with
 * key padding kpm[B,Sk] (nonzero = masked); pad columns come out 0 (torch18_adapters.py:578-597,789-798).
 * ------------------------------------------------------------------------------------------ */
int xf_split3(const float* src, int64_t lds, int rows, int cols, void* dst_bf16, int pattern, int act, int bias_cols,
              const float* bias, xf_stream_t stream);
int xf_softmax_rows_f32(float* s, int B, int H, int Sq, int Sk, int Sp, const uint8_t* kpm, float scale, xf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused optimizer step over the path's parameters: global-norm clipping (Lightning gradient_clip_val,
 * runner/run_experiment.py:445-446 -> torch clip_grad_norm_) + RAdam (runner/metrics_losses/radam_optim.py:30-104) +
 * the bf16 weight copy of the next forward, one HBM pass.  Up to XF_OPT_MAX_JOBS tensors per call; all fp32 pointers
 * 16-byte aligned.  xf_grad_sqnorm ADDS sum(grad^2) over the jobs to *out (zero it first; add the squared norm of any
 * gradient outside this path before the step so the clip is global).  xf_radam_step: the moments are always updated;
 * the parameter moves per radam_optim.py:89-102 (N_sma >= 5: adaptive; else SGD-like iff degenerated_to_sgd; else not
 * at all) with gradients pre-scaled by min(1, max_grad_norm / (sqrt(*grad_sqnorm) + 1e-6)) when max_grad_norm > 0.
 * Gradients themselves are NOT rewritten (clip_grad_norm_ scales them in place; nothing on this path reads them again).
 * ------------------------------------------------------------------------------------------ */
#define XF_OPT_MAX_JOBS 32
typedef struct XfRAdamJob {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
  void* param_bf16;      /* optional: bf16 copy of the updated parameter (same flat indexing), or NULL */
  int64_t n;
} XfRAdamJob;
typedef struct XfRAdam {
  /* doubles: the reference evaluates its scalar coefficients (1 - beta, weight_decay * lr, step_size * lr, N_sma) with
   * Python floats before they meet an fp32 tensor; doing that arithmetic in fp32 is 5e-5 off (1.f - 0.999f) */
  double lr_d, beta1_d, beta2_d, weight_decay_d;
  float eps;
  int32_t degenerated_to_sgd;
  int64_t step;                 /* 1-based step count of these tensors (state["step"] after the increment, :65) */
  float max_grad_norm;          /* <= 0: no clipping */
  const float* grad_sqnorm;     /* device scalar (see xf_grad_sqnorm), or NULL */
} XfRAdam;
int xf_grad_sqnorm(const XfRAdamJob* jobs, int n_jobs, float* out, xf_stream_t stream);
int xf_radam_step(const XfRAdamJob* jobs, int n_jobs, const XfRAdam* a, xf_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Test aids (not on the hot path): materialise the counter-based dropout keep masks that the kernels above
 * recompute on the fly, so a CPU checker can replay a train-mode step with the SAME masks (the reference draws
 * its masks from torch's RNG inside F.dropout: cross_f_box_layers.py:74, torch18_adapters.py:109-112,796-797,
 * utils.py:115).  out is uint8, 1 = keep.
 *   xf_debug_dropout_mask     : GEMM-epilogue / LayerNorm / row-gather sites; out[r, c] for rows row0 .. row0+rows-1
 *                               (row = the OUTPUT row index m_out of the site), columns 0 .. cols-1.
 *   xf_debug_attn_dropout_mask: attention probabilities; out[bh, q, k], bh = b * H + head.
 * ------------------------------------------------------------------------------------------ */
int xf_debug_dropout_mask(float p, uint32_t seed, uint32_t stream_id, int64_t row0, int rows, int cols, uint8_t* out,
                          xf_stream_t stream);
int xf_debug_attn_dropout_mask(float p, uint32_t seed, uint32_t stream_id, int BH, int Sq, int Sk, uint8_t* out,
                               xf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* XFUSION_H_ */
