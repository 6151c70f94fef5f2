/*
 * xqb200.h -- C ABI of libxqb200.so: the B200 (sm_100a) quantizer hot path of the XQ-GAN /
 * ImageFolder image tokenizer.
 *
 * The reference (lxa9867/ImageFolder) is pure Python; its "operator boundary" for this path is
 * the nn.Module surface of its quantizers.  Each entry point below replaces
 * the arithmetic of one reference method and is what a binding for that method calls
 * (INTEGRATION.md shows the ctypes stubs; imagefolder_b200/_capi.py is the in-repo binding).
 *
 * Conventions
 *   - plain pointers and sizes only; every tensor pointer is a DEVICE pointer owned by the caller
 *     (PyTorch), contiguous, fp32 unless stated; indices are int64 like torch.long.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises.
 *   - no allocation, no global state: scratch memory is a caller-provided workspace whose size the
 *     matching *_workspace_bytes() call returns; calls are re-entrant across streams when the
 *     workspaces differ.
 *   - return value: 0 = ok, negative = error (xq_strerror()); the Python side maps it to
 *     RuntimeError / ValueError, mirroring the reference's exceptions/asserts.
 *   - tensors named `*_nchw` are [B, C, H*W] exactly as the reference passes them (B,C,H,W
 *     contiguous); "rows" n = b*HW + p follow the reference's 'b c h w -> b h w c' flattening
 *     (xqgan_model.py:750-751).
 */
#ifndef XQB200_H_
#define XQB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XQ_OK 0
#define XQ_ERR_ARG (-1)         /* bad shape / null pointer / unsupported size */
#define XQ_ERR_WORKSPACE (-2)   /* workspace too small */
#define XQ_ERR_CUDA (-3)        /* a CUDA runtime call or launch failed */
#define XQ_ERR_UNSUPPORTED (-4) /* valid in the reference, not built here */

#define XQ_MAX_SCALES 32

const char *xq_strerror(int code);
int xq_abi_version(void);
/* last CUDA error string recorded on this thread by a failed call (for diagnostics) */
const char *xq_last_cuda_error(void);

/* ------------------------------------------------------------------------------------------
 * Single-scale VectorQuantizer   (VQ-4096 / VQ-8192 / VP2-* / RobustTok)
 *   replaces VectorQuantizer.forward           tokenizer/tokenizer_image/xqgan_model.py:745-801
 *            VectorQuantizer.f_to_idxBl_or_fhat                      xqgan_model.py:803-833
 * ------------------------------------------------------------------------------------------ */
size_t xq_vq_workspace_bytes(int B, int C, int HW, int V);

/*
 * Fused normalise -> distance -> argmin -> gather -> normalise -> STE value -> MSE partials
 * -> usage histogram.  The N x V distance matrix is never written to memory.
 *   z_nchw        [B,C,HW]   encoder latent (input of the reference forward)
 *   E             [V,C]      embedding.weight (raw)
 *   codebook_norm 1: rows and codes are L2-normalised first (xqgan_model.py:753-756)
 *   ste_value     1: out = zn + (q - zn)  (forward, :796)   0: out = q  (f_to_idxBl_or_fhat :826-831)
 *   idx           [B*HW]     argmin index per row (first index on ties)
 *   out_nchw      [B,C,HW]
 *   loss          [2]        {vq_loss, commit_loss} = {mse, beta*mse} (:792-793); may be NULL
 *   hist          [V]        += bincount(idx) as float (:774); may be NULL
 */
int xq_vq_forward(const float *z_nchw, const float *E, int B, int C, int HW, int V, int codebook_norm,
                  int ste_value, float beta, int64_t *idx, float *out_nchw, float *loss, float *hist,
                  void *workspace, size_t workspace_bytes, void *stream);

/*
 * Backward of (out, vq_loss, commit_loss) wrt z and E.
 *   g_out_nchw [B,C,HW] or NULL; g_vq / g_commit: device scalars or NULL (treated as 0)
 *   gz_nchw    [B,C,HW]  written;   gE [V,C] overwritten (zeroed, then scatter-added)
 */
int xq_vq_backward(const float *z_nchw, const float *E, const int64_t *idx, const float *g_out_nchw,
                   const float *g_vq, const float *g_commit, int B, int C, int HW, int V, int codebook_norm,
                   float beta, float *gz_nchw, float *gE, void *stream);

/* ------------------------------------------------------------------------------------------
 * Latent perturbation   (RobustTok)
 *   replaces add_perturbation      tokenizer/tokenizer_image/latent_perturbation.py:4-35
 * The two random tensors the reference draws (torch.rand(N) :21, torch.randint(0,delta,(N,)) :22)
 * are INPUTS, so the host keeps the reference's RNG stream.
 * ------------------------------------------------------------------------------------------ */
size_t xq_perturb_workspace_bytes(int B, int C, int HW, int V);
/*   n_perturb = int(B * beta) evaluated by the host (Python double arithmetic, :32)
 *   out_nchw [B,C,HW] = where(b < n_perturb, zn + (normalize(E[sel]) - zn), zq)
 *   sel      [n_perturb*HW] chosen code per perturbed row (may be NULL) */
int xq_perturb_forward(const float *z_nchw, const float *zq_nchw, const float *E, const float *rand_u,
                       const int64_t *rand_j, int B, int C, int HW, int V, int codebook_norm, float alpha,
                       int n_perturb, int delta, float *out_nchw, int64_t *sel, void *workspace,
                       size_t workspace_bytes, void *stream);
/*   g [B,C,HW] -> gz (through the normalisation Jacobian, perturbed samples only), gzq (the rest) */
int xq_perturb_backward(const float *z_nchw, const float *g_nchw, int B, int C, int HW, int codebook_norm,
                        int n_perturb, float *gz_nchw, float *gzq_nchw, void *stream);

/* ------------------------------------------------------------------------------------------
 * Multi-scale residual quantizers (MSVR*, MSBR*)
 *   replaces VectorQuantizer2.forward / f_to_idxBl_or_fhat   tokenizer/tokenizer_image/quant.py:64-223
 *            LFQ.forward / f_to_idxBl_or_fhat    tokenizer/tokenizer_image/lookup_free_quantize.py:149-380
 *            Phi.forward                                                   quant.py:261-268
 * One CTA owns one image: residual, accumulated f_hat and the upsampled code map stay in shared
 * memory across all scales.
 * ------------------------------------------------------------------------------------------ */
#define XQ_MS_VQ_ZNORM 0 /* VectorQuantizer2, using_znorm=True  (argmax cosine)        */
#define XQ_MS_VQ_L2 1    /* VectorQuantizer2, using_znorm=False (argmin L2)            */
#define XQ_MS_BSQ 2      /* LFQ: sign bits, code = +-scaler[si]                        */

typedef struct {
    int B, C, H, W;       /* f is [B,C,H,W]                                             */
    int V;                /* codebook size (BSQ: 2^C)                                   */
    int K;                /* number of Phi modules (0 = identity)                       */
    int SN;               /* number of scales                                           */
    int mode;             /* XQ_MS_*                                                    */
    int patch_nums[XQ_MAX_SCALES];
    int phi_map[XQ_MAX_SCALES]; /* scale -> Phi index (PhiPartiallyShared, quant.py:279-288) */
    float scaler[XQ_MAX_SCALES]; /* BSQ code magnitude per scale (lookup_free_quantize.py:124-128) */
    float resi_ratio;     /* Phi blend r (quant.py:265)                                 */
    float beta;           /* commit weight                                              */
    int loss_div_sn_all;  /* 0: only vq is divided by SN (quant.py:134)  1: all (LFQ :238-240) */
    int channel_norm;     /* 1: f is L2-normalised over C first (LFQ using_znorm, :153) */
    float entropy_weight, w_sample, w_batch; /* LFQ entropy term                        */
} xq_ms_desc;

size_t xq_ms_workspace_bytes(const xq_ms_desc *d);
size_t xq_ms_saved_bytes(const xq_ms_desc *d); /* bytes of `saved` (forward -> backward) */
int64_t xq_ms_total_tokens(const xq_ms_desc *d); /* sum_si B*pn^2 */

/*
 *   f            [B,C,H,W]
 *   E            [V,C] raw codebook (NULL for BSQ)
 *   phi_w/phi_b  [K,C,C,3,3] / [K,C]
 *   n_quantizers [B] float, scale si contributes to sample b iff si < n_quantizers[b]
 *                (quant.py:79-86,115); NULL = no quantizer dropout
 *   with_losses  0: inference (f_to_idxBl_or_fhat): no masks, no losses
 *   out          [B,C,H,W]  forward: (f_hat - f) + f (quant.py:135); inference: f_hat
 *   idx_all      [sum_si B*pn^2] int64, scale-major, then (b, y, x)
 *   fhat_scales  [SN,B,C,H,W] cumulative f_hat after each scale, or NULL (to_fhat=True lists)
 *   loss         [3] {vq, commit, entropy}
 *   hist         [SN,V] += bincount per scale, or NULL
 *   saved        xq_ms_saved_bytes(): state the backward needs (final masked f_hat, ...)
 */
int xq_ms_forward(const xq_ms_desc *d, const float *f, const float *E, const float *phi_w, const float *phi_b,
                  const float *n_quantizers, int with_losses, float *out, int64_t *idx_all, float *fhat_scales,
                  float *loss, float *hist, void *saved, void *workspace, size_t workspace_bytes, void *stream);

/*
 * Backward wrt f, E, phi_w, phi_b.
 *   g_out [B,C,H,W] or NULL; g_vq/g_commit/g_entropy device scalars or NULL
 *   gf [B,C,H,W]; gE [V,C] (NULL for BSQ); gphi_w [K,C,C,3,3]; gphi_b [K,C]  -- all overwritten
 */
int xq_ms_backward(const xq_ms_desc *d, const float *f, const float *E, const float *phi_w, const float *phi_b,
                   const float *n_quantizers, const int64_t *idx_all, const void *saved, const float *g_out,
                   const float *g_vq, const float *g_commit, const float *g_entropy, float *gf, float *gE,
                   float *gphi_w, float *gphi_b, void *workspace, size_t workspace_bytes, void *stream);

/*
 * VAR-side helpers built from the same primitives (quant.py:148-180, 226-258):
 * given token indices per scale, rebuild f_hat (all scales) and the next-scale inputs.
 *   var_input [B, sum_{si>=1} pn_si^2, C] (idxBl_to_var_input) or NULL
 *   fhat_scales [SN,B,C,H,W] or NULL ; out [B,C,H,W] final f_hat or NULL
 */
int xq_ms_decode(const xq_ms_desc *d, const int64_t *idx_all, const float *E, const float *phi_w,
                 const float *phi_b, float *out, float *fhat_scales, float *var_input, void *stream);

/*
 * The same step in FEATURE-MAP form (the VAR generator's per-step loop and the VAE's embed_to_fhat):
 *   VectorQuantizer2.embed_to_fhat(all_to_max_scale=True)  quant.py:148-166   -> si0 = 0, si1 = SN, fhat_scales / out
 *   VectorQuantizer2.get_next_autoregressive_input         quant.py:247-258   -> si1 = si0 + 1, fhat_in = out (in place), next
 *   (LFQ: lookup_free_quantize.py:311-343, 404-415)
 * for si in [si0, si1):  f_hat += Phi_si(bicubic_up(h_si))   (no interpolation at the last scale)
 *   h_all       scales si0..si1-1 packed back to back, each [B,C,pn_si,pn_si] fp32
 *   fhat_in     [B,C,H,W] running f_hat or NULL (= zeros); may alias out
 *   out         [B,C,H,W] f_hat after scale si1-1, or NULL
 *   fhat_scales [si1-si0,B,C,H,W] cumulative f_hat after every scale, or NULL
 *   next        [B,C,pn_si1,pn_si1] = area-pool of the final f_hat (ignored when si1 == SN), or NULL
 */
int xq_ms_embed(const xq_ms_desc *d, int si0, int si1, const float *h_all, const float *phi_w, const float *phi_b,
                const float *fhat_in, float *out, float *fhat_scales, float *next, void *stream);

/* ------------------------------------------------------------------------------------------
 * Codebook-usage EMA (xqgan_model.py:777-788, quant.py:121-127,137-141)
 *   ema[rows,V], hit[rows,V]: row i <- copy | 0.9/0.1 | 0.99/0.01 blend of hit[i], chosen by
 *   (record_hit + i) == 0 | < 100 | otherwise  (the reference bumps record_hit once per scale).
 *   usage_out[rows] (device, may be NULL) = 100 * mean(ema[i] >= margin)
 * ------------------------------------------------------------------------------------------ */
int xq_usage_ema(float *ema, const float *hit, int rows, int V, int record_hit, float margin,
                 float *usage_out, void *stream);
/* same, with the step counter on the device: record_hit_dev[0] = `record_hit` (read, then advanced by `rows` by the kernel),
 * record_hit_dev[1] = scratch (zero-initialised once).  The host neither reads nor writes the counter, which keeps the call
 * CUDA-graph capturable and lets torch.compile trace the module without specialising on a Python int. */
int xq_usage_ema_dev(float *ema, const float *hit, int rows, int V, int64_t *record_hit_dev, float margin,
                     float *usage_out, void *stream);

/* ------------------------------------------------------------------------------------------
 * ViT block glue (DINOv2Encoder / DINOv2Decoder blocks)
 *   replaces the non-GEMM ops of Block.forward   tokenizer/tokenizer_image/dino_enc/vision_transformer.py:336-339
 *   (LayerNorm :301,316 ; LayerScale :280-292 ; DropPath ; residual add) and Mlp's GELU.
 * Residual stream fp32 [M,D], GEMM operands bf16 (what bf16 autocast gives the reference).
 * D in {384, 768, 1024}.  `branch`, `y`, `g_y`, `g_branch` are bf16 [M,D].
 * ------------------------------------------------------------------------------------------ */
/*   x_out = x + rowscale[row / rows_per_sample] * ls_gamma[d] * (branch + branch_bias[d])
 *           (branch may be NULL: x_out = x; branch_bias = bias of the GEMM that produced `branch`, folded here
 *            so that its gradient is a free column sum of the backward kernel; may be NULL)
 *   y     = LayerNorm(x_out; eps) * ln_w + ln_b   (bf16; may be NULL)   mean / rstd [M] saved for backward
 *   rowscale [B] = DropPath keep mask / keep_prob (NULL = 1);  x_out may alias x or be NULL */
int xq_vit_residual_ln_fwd(const float *x, const void *branch, const float *branch_bias, const float *ls_gamma,
                           const float *rowscale, int rows_per_sample, const float *ln_w, const float *ln_b, float eps,
                           int M, int D, float *x_out, void *y, float *mean, float *rstd, void *stream);
size_t xq_vit_ln_bwd_workspace_bytes(int D);
/*   G = g_xout + LayerNorm^T(g_y)  -> g_x [M,D] fp32 ; g_branch = G * rowscale * ls_gamma (bf16)
 *   g_ln_w, g_ln_b, g_ls_gamma, g_branch_bias [D] overwritten (any may be NULL) */
int xq_vit_residual_ln_bwd(const float *g_xout, const void *g_y, const float *x_out, const float *mean,
                           const float *rstd, const float *ln_w, const void *branch, const float *branch_bias,
                           const float *ls_gamma, const float *rowscale, int rows_per_sample, int M, int D, float *g_x,
                           void *g_branch, float *g_ln_w, float *g_ln_b, float *g_ls_gamma, float *g_branch_bias,
                           void *workspace, size_t workspace_bytes, void *stream);
/*   gradient re-packing of the fused qkv projection (vision_transformer.py:175-176): dq, dk, dv [M,C] bf16
 *   dense -> dqkv [M,3C]; replaces autograd's stack + permute + contiguous copies.  g_bias [3C] fp32 (may be NULL)
 *   receives the column sums of dqkv = the gradient of the qkv bias (nn.Linear's backward `sum(0)` pass, fused). */
size_t xq_vit_pack_workspace_bytes(void);
int xq_vit_pack_qkv(const void *dq, const void *dk, const void *dv, void *dqkv, float *g_bias, size_t M, int C,
                    void *workspace, size_t workspace_bytes, void *stream);
/*   im2col of the patch embedding (timm PatchEmbed = Conv2d(kernel = stride = p), vision_transformer.py PatchEmbed.forward):
 *   x fp32 [B,Cin,H,W] -> patches bf16 [B*(H/p)*(W/p), Cin*p*p] (K index = (c*p + ky)*p + kx = the flattened conv
 *   weight), so that tokens = patches @ weight.view(D,-1)^T + bias is a plain GEMM.  p % 4 == 0, H % p == W % p == 0. */
int xq_vit_patchify(const float *x, void *patches, int B, int Cin, int H, int W, int p, void *stream);
/*   token assembly of the ViT encoder / decoder input (dino_enc/dinov2.py:151-170, 318-336):
 *     out[b,t,:] = table[t,:] + (t0 <= t < t0+Ls ? src[b,t-t0,:] : 0)   out fp32 [B,T,D], table fp32 [T,D] (the batch-
 *   independent part: cls / mask / latent tokens + positional + level embeddings), src [B,Ls,D] fp32 or bf16.
 *   backward: d_src = g[:, t0:t0+Ls] in the source dtype (may be NULL), d_table = sum_b g (may be NULL); one read of g. */
int xq_vit_assemble_fwd(const void *src, int src_is_bf16, const float *table, int B, int Ls, int T, int D, int t0, float *out,
                        void *stream);
int xq_vit_assemble_bwd(const float *g, int B, int Ls, int T, int D, int t0, void *d_src, int src_is_bf16, float *d_table,
                        void *stream);
/*   y = GELU(x + bias) exact-erf form (timm Mlp act_layer=nn.GELU), x / y bf16 [M,C], bias fp32 [C] or NULL,
 *   C % 8 == 0.  Backward also returns g_bias [C] = column sums of gx (may be NULL). */
int xq_vit_gelu_fwd(const void *x, const float *bias, void *y, int M, int C, void *stream);
int xq_vit_gelu_bwd(const void *x, const float *bias, const void *gy, void *gx, float *g_bias, int M, int C,
                    void *stream);

/*   Flash attention of the ViT blocks, head_dim 64, no mask, no dropout (Attention.forward,
 *   tokenizer/tokenizer_image/dino_enc/vision_transformer.py:173-197: F.scaled_dot_product_attention on
 *   qkv.reshape(B,N,3,H,hd).permute(2,0,3,1,4), then x.transpose(1,2).reshape(B,N,C)).  tcgen05 / TMEM / TMA kernel.
 *     qkv   bf16 [B,N,3,H,64]  the packed projection, read in place (no q/k/v copies)
 *     out   bf16 [B,N,H*64]    head-merged attention output (what `proj` consumes)
 *     lse2  fp32 [B,H,N]       base-2 log-sum-exp of the scaled scores (scale*log2(e)*q.k), saved for backward
 *   scale = head_dim^-0.5 (Attention.scale).  Any N >= 1; qkv / out 16-byte aligned. */
int xq_vit_attn_fwd(const void *qkv, void *out, float *lse2, int B, int N, int H, int head_dim, float scale, void *stream);

/*   Backward of xq_vit_attn_fwd: d_out bf16 [B,N,H*64] -> dqkv bf16 [B,N,3,H,64] (the gradient of the packed projection,
 *   written in place of autograd's three permuted tensors + stack).  `out` and `lse2` are the forward's results.
 *   workspace (256-byte aligned, xq_vit_attn_bwd_workspace_bytes): fp32 dQ accumulator [B*H,N,64] (TMA reduce-add across
 *   the key blocks) + the padded statistics.  g_bias fp32 [3*H*64] (may be NULL) receives the column sums of dqkv = the
 *   gradient of the qkv bias (nn.Linear's backward `sum(0)` pass, fused into the epilogues).  3 launches + memsets. */
size_t xq_vit_attn_bwd_workspace_bytes(int B, int N, int H);
int xq_vit_attn_bwd(const void *qkv, const void *out, const void *d_out, const float *lse2, void *dqkv, float *g_bias, int B, int N,
                    int H, int head_dim, float scale, void *workspace, size_t workspace_bytes, void *stream);

/*
 * ---- loss stack -------------------------------------------------------------------
 * LPIPS stage distance (tokenizer/tokenizer_image/lpips.py:79-90): for one VGG stage with feature maps f0, f1
 * [B,C,H*W] (fp32, or bf16 when is_bf16) and the stage's `lin` weights lin_w [C]:
 *   out[b] (+)= mean_p sum_c lin_w[c] * ( f0/(|f0|_c + eps) - f1/(|f1|_c + eps) )^2      (accumulate != 0: add to out)
 * backward returns the gradient w.r.t. f1 (swap the maps for f0); g_out [B] is d loss / d out.
 * The fp64 per-CTA partials live in the caller's workspace (xq_lpips_workspace_bytes).
 */
size_t xq_lpips_workspace_bytes(int B, int HW);
int xq_lpips_layer_forward(const void *f0, const void *f1, int is_bf16, const float *lin_w, int B, int C, int HW, float eps,
                           int accumulate, float *out, void *workspace, size_t workspace_bytes, void *stream);
int xq_lpips_layer_backward(const void *f0, const void *f1, int is_bf16, const float *lin_w, int B, int C, int HW, float eps,
                            const float *g_out, void *g_f1, void *stream);
/*
 * DiffAug.aug without the warm-up blur (tokenizer/tokenizer_image/diffaug.py:60-118): translation (zero fill), colour
 * (brightness, saturation about the per-pixel channel mean, contrast about the per-sample mean) and cutout.
 *   x, y, g, gx  [B,C,H,W] fp32, C <= 8 ;  rand01 [7,B] = the reference's torch.rand(7,B,1,1) (:64) ;
 *   flags: bit0 translation, bit1 colour, bit2 cutout (the reference's three `torch.rand(3) <= prob` draws, :61) ;
 *   cut_h, cut_w = round(H*cutout), round(W*cutout) ; sums [B] scratch.
 * backward is the exact transpose of the (per-sample affine) forward map.
 */
int xq_diffaug_forward(const float *x, const float *rand01, int B, int C, int H, int W, int flags, int cut_h, int cut_w,
                       float *y, float *sums, void *stream);
int xq_diffaug_backward(const float *g, const float *rand01, int B, int C, int H, int W, int flags, int cut_h, int cut_w,
                        float *gx, float *sums, void *stream);

/* ---- ViT MLP with the element-wise work fused into a hand-written tcgen05 GEMM (csrc/gemm_kernel.cu) --------------------
 * Replaces, inside timm's Mlp as called by Block.forward (tokenizer/tokenizer_image/dino_enc/vision_transformer.py:336-339):
 *   forward   F.linear(y, W1) [cuBLAS] + GELU(. + b1) [xq_vit_gelu_fwd]            -> xq_vit_fc1_gelu_fwd
 *   backward  d_act = d_out W2 [cuBLAS] + d_act * GELU'(pre + b1), d_b1 [xq_vit_gelu_bwd] -> xq_vit_fc2_dgelu_bwd
 * All matrices row-major bf16; bias / d_bias fp32.  N % 256 == 0, K % 64 == 0 (else XQ_ERR_UNSUPPORTED: the caller keeps the
 * library GEMM + stand-alone kernel), any M.  Results are bit-identical to that two-call sequence.
 *   x [M,K], w [N,K] (fc1.weight as bf16)  ->  pre [M,N] = x w^T ,  act [M,N] = GELU(pre + bias)                               */
int xq_vit_fc1_gelu_fwd(const void *x, const void *w, const float *bias, void *pre, void *act, int M, int N, int K, void *stream);
/*  d_out [M,K] (gradient of the fc2 output), w2t [N,K] (fc2.weight TRANSPOSED, bf16), pre [M,N] (saved by the forward)
 *   ->  d_pre [M,N] = (d_out w2t^T) * GELU'(pre + bias) ,  d_bias [N] = column sums of the rounded d_pre                      */
int xq_vit_fc2_dgelu_bwd(const void *d_out, const void *w2t, const void *pre, const float *bias, void *d_pre, float *d_bias,
                         int M, int N, int K, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* XQB200_H_ */
