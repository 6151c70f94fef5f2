/*
 * xq_oracle.c -- CPU ORACLE for the XQ-GAN quantizer hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under imagefolder_b200/ may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it (as the checker / the timed CPU
 * baseline, never as the product).
 *
 * It restates, in plain C, the arithmetic of the reference's quantizers:
 *   VectorQuantizer.forward           tokenizer/tokenizer_image/xqgan_model.py:745-801
 *   VectorQuantizer.f_to_idxBl_or_fhat                       xqgan_model.py:803-833
 *   VectorQuantizer2.forward          tokenizer/tokenizer_image/quant.py:64-144
 *   VectorQuantizer2.f_to_idxBl_or_fhat                      quant.py:182-223
 *   LFQ.forward                       tokenizer/tokenizer_image/lookup_free_quantize.py:149-250
 *   add_perturbation                  tokenizer/tokenizer_image/latent_perturbation.py:4-35
 *   Phi.forward                       quant.py:261-268
 * and of the ATen ops those call (F.normalize, F.interpolate area / bicubic).
 *
 * Parity pinning: the reference ships no golden vectors, so this
 * oracle is pinned against outputs of the reference's own Python modules run in the
 * build container (tests/golden/make_golden.py -> tests/golden/*.npz,
 * tests/test_oracle_golden.py).
 *
 * CANONICAL ARITHMETIC (DESIGN.md "Canonical arithmetic"): every value that feeds an index
 * decision is computed in IEEE fp32, round-to-nearest, in a fixed operation order,
 * with fused multiply-add only where written as fmaf().  The CUDA kernels follow the
 * same order, which is what makes token indices bit-exact between GPU and oracle.
 * Compile with -ffp-contract=off (see oracle/build.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define XQ_EPS 1e-12f

/* ---- canonical primitives ------------------------------------------------------- */

/* dot(a,b) = fma chain in ascending k, starting from +0. */
static inline float xq_dot(const float *a, const float *b, int C) {
    float acc = 0.0f;
    for (int k = 0; k < C; ++k) acc = fmaf(a[k], b[k], acc);
    return acc;
}

/* F.normalize(x, p=2, dim=-1, eps=1e-12): y = x / max(sqrt(sum x^2), eps)
 * (xqgan_model.py:753-756, quant.py:93).  Returns the clamped denominator. */
static inline float xq_l2norm(const float *x, int C, float *y) {
    float den = fmaxf(sqrtf(xq_dot(x, x, C)), XQ_EPS);
    for (int k = 0; k < C; ++k) y[k] = x[k] / den;
    return den;
}

/* ---- exported helpers --------------------------------------------------------------- */

/* rows: [n, C] row-major; out y [n, C], den [n] (may be NULL). */
void xqo_l2norm_rows(const float *x, int64_t n, int C, float *y, float *den) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float d = xq_l2norm(x + i * C, C, y + i * C);
        if (den) den[i] = d;
    }
}

/* NCHW [B,C,HW] -> rows [B*HW, C] (the reference's 'b c h w -> b h w c' permute). */
void xqo_nchw_to_rows(const float *x, int B, int C, int HW, float *rows) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)B * HW; ++n) {
        int b = (int)(n / HW), p = (int)(n % HW);
        for (int c = 0; c < C; ++c) rows[n * C + c] = x[((int64_t)b * C + c) * HW + p];
    }
}

void xqo_rows_to_nchw(const float *rows, int B, int C, int HW, float *x) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)B * HW; ++n) {
        int b = (int)(n / HW), p = (int)(n % HW);
        for (int c = 0; c < C; ++c) x[((int64_t)b * C + c) * HW + p] = rows[n * C + c];
    }
}

/*
 * Codebook search over rows.
 *   metric 0 (L2):  d = (zz + ee[v]) - 2*dot   -> argmin, first index on ties
 *                   (xqgan_model.py:761-766, latent_perturbation.py:16-18, quant.py:98-101)
 *   metric 1 (DOT): s = dot                    -> argmax, first index on ties (quant.py:94)
 * rows [n,C], codes [V,C] are used as given (normalise beforehand when the caller does).
 * Outputs: idx [n]; best/second [n] (may be NULL) = best and runner-up score, so that
 * callers can classify reference mismatches as near-ties.
 */
void xqo_search(const float *rows, int64_t n, const float *codes, int V, int C, int metric,
                int64_t *idx, float *best, float *second) {
    float *ee = NULL;
    if (metric == 0) {
        ee = (float *)malloc(sizeof(float) * (size_t)V);
        for (int v = 0; v < V; ++v) ee[v] = xq_dot(codes + (size_t)v * C, codes + (size_t)v * C, C);
    }
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; ++i) {
        const float *z = rows + i * C;
        float zz = (metric == 0) ? xq_dot(z, z, C) : 0.0f;
        float b1 = INFINITY, b2 = INFINITY;
        int64_t bi = 0;
        for (int v = 0; v < V; ++v) {
            float dot = xq_dot(z, codes + (size_t)v * C, C);
            /* key: smaller is better for both metrics */
            float key = (metric == 0) ? fmaf(-2.0f, dot, zz + ee[v]) : -dot;
            if (key < b1) { b2 = b1; b1 = key; bi = v; }
            else if (key < b2) { b2 = key; }
        }
        idx[i] = bi;
        if (best) best[i] = (metric == 0) ? b1 : -b1;
        if (second) second[i] = (metric == 0) ? b2 : -b2;
    }
    free(ee);
}

typedef struct { float d; int v; } xq_pair;
static int xq_pair_cmp(const void *a, const void *b) {
    const xq_pair *x = (const xq_pair *)a, *y = (const xq_pair *)b;
    if (x->d < y->d) return -1;
    if (x->d > y->d) return 1;
    return (x->v > y->v) - (x->v < y->v);
}

/*
 * Rank-select used by add_perturbation (latent_perturbation.py:20-24):
 * topk(d, delta, largest=False) is sorted ascending, the reference then takes column
 * rank[i].  Canonical order = (d, index) lexicographic.  out_idx[i] = code at that rank.
 * If topk_out != NULL it receives the first `delta` codes per row ([n, delta]).
 */
void xqo_rank_select(const float *rows, int64_t n, const float *codes, int V, int C,
                     const int64_t *rank, int delta, int64_t *out_idx, int64_t *topk_out) {
    float *ee = (float *)malloc(sizeof(float) * (size_t)V);
    for (int v = 0; v < V; ++v) ee[v] = xq_dot(codes + (size_t)v * C, codes + (size_t)v * C, C);
#pragma omp parallel
    {
        xq_pair *p = (xq_pair *)malloc(sizeof(xq_pair) * (size_t)V);
#pragma omp for schedule(dynamic, 4)
        for (int64_t i = 0; i < n; ++i) {
            const float *z = rows + i * C;
            float zz = xq_dot(z, z, C);
            for (int v = 0; v < V; ++v) {
                float dot = xq_dot(z, codes + (size_t)v * C, C);
                p[v].d = fmaf(-2.0f, dot, zz + ee[v]);
                p[v].v = v;
            }
            qsort(p, (size_t)V, sizeof(xq_pair), xq_pair_cmp);
            out_idx[i] = p[rank[i]].v;
            if (topk_out) for (int j = 0; j < delta; ++j) topk_out[i * delta + j] = p[j].v;
        }
        free(p);
    }
    free(ee);
}

/* ---- resampling (ATen semantics) ----------------------------- */

/* F.interpolate(mode='area') H->P on one [H,W] plane == adaptive_avg_pool2d:
 * window [floor(i*H/P), ceil((i+1)*H/P)); plain sequential sum (row-major), / count. */
static inline float xq_area_px(const float *plane, int H, int W, int P, int oy, int ox) {
    int y0 = (oy * H) / P, y1 = ((oy + 1) * H + P - 1) / P;
    int x0 = (ox * W) / P, x1 = ((ox + 1) * W + P - 1) / P;
    float acc = 0.0f;
    for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) acc = acc + plane[y * W + x];
    return acc / (float)((y1 - y0) * (x1 - x0));
}

/* f [B,C,H,W] -> rows [B*P*P, C] (NHWC rows, as quant.py:91 permute(0,2,3,1).reshape(-1,C)). */
void xqo_area_pool_rows(const float *f, int B, int C, int H, int W, int P, float *rows) {
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b)
        for (int oy = 0; oy < P; ++oy)
            for (int ox = 0; ox < P; ++ox)
                for (int c = 0; c < C; ++c) {
                    const float *plane = f + ((size_t)b * C + c) * H * W;
                    float v = (P == H && P == W) ? plane[oy * W + ox] : xq_area_px(plane, H, W, P, oy, ox);
                    rows[(((size_t)b * P + oy) * P + ox) * C + c] = v;
                }
}

/* bicubic taps, A=-0.75, align_corners=False (ATen UpSample.h: cubic_convolution1/2,
 * get_cubic_upsample_coefficients, area_pixel_compute_source_index(cubic=true)). */
static inline void xq_cubic_taps(int dst, int in_size, int out_size, int idx[4], float w[4]) {
    const float A = -0.75f;
    float scale = (float)in_size / (float)out_size;
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    float fl = floorf(src);
    float t = src - fl;
    int i0 = (int)fl;
    float x;
    x = t + 1.0f; w[0] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
    x = t;        w[1] = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
    x = 1.0f - t; w[2] = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
    x = (1.0f - t) + 1.0f; w[3] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
    for (int k = 0; k < 4; ++k) {
        int j = i0 - 1 + k;
        idx[k] = j < 0 ? 0 : (j > in_size - 1 ? in_size - 1 : j);
    }
}

void xqo_cubic_table(int in_size, int out_size, int32_t *idx /*[out,4]*/, float *w /*[out,4]*/) {
    for (int d = 0; d < out_size; ++d) {
        int ii[4];
        xq_cubic_taps(d, in_size, out_size, ii, w + d * 4);
        for (int k = 0; k < 4; ++k) idx[d * 4 + k] = ii[k];
    }
}

/* src rows [B,P,P,C] (gathered codes, NHWC) -> u [B,C,H,W] bicubic (quant.py:107).
 * out = fma-chain_i wy[i] * (fma-chain_j wx[j]*src[yi][xj]).  P==H -> plain transpose. */
void xqo_bicubic_up(const float *src, int B, int C, int P, int H, int W, float *u) {
    int32_t *iy = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)(H + W));
    float *wy = (float *)malloc(sizeof(float) * 4 * (size_t)(H + W));
    int32_t *ix = iy + 4 * H;
    float *wx = wy + 4 * H;
    xqo_cubic_table(P, H, iy, wy);
    xqo_cubic_table(P, W, ix, wx);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    float out;
                    if (P == H && P == W) {
                        out = src[(((size_t)b * P + y) * P + x) * C + c];
                    } else {
                        out = 0.0f;
                        for (int i = 0; i < 4; ++i) {
                            float inner = 0.0f;
                            for (int j = 0; j < 4; ++j)
                                inner = fmaf(wx[x * 4 + j],
                                             src[(((size_t)b * P + iy[y * 4 + i]) * P + ix[x * 4 + j]) * C + c], inner);
                            out = fmaf(wy[y * 4 + i], inner, out);
                        }
                    }
                    u[(((size_t)b * C + c) * H + y) * W + x] = out;
                }
    free(iy); free(wy);
}

/* Phi (quant.py:261-268): h = u*(1-r) + (conv3x3(u; w) + b)*r, zero padding 1.
 * conv accumulates from the bias, ci-major then ky, kx, as an fma chain. */
void xqo_phi(const float *u, int B, int C, int H, int W, const float *w /*[C,C,3,3]*/, const float *bias /*[C]*/,
             float r, float *h) {
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b)
        for (int co = 0; co < C; ++co)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    float acc = bias[co];
                    for (int ci = 0; ci < C; ++ci) {
                        const float *plane = u + ((size_t)b * C + ci) * H * W;
                        const float *wk = w + ((size_t)co * C + ci) * 9;
                        for (int ky = 0; ky < 3; ++ky) {
                            int yy = y + ky - 1;
                            if (yy < 0 || yy >= H) continue;
                            for (int kx = 0; kx < 3; ++kx) {
                                int xx = x + kx - 1;
                                if (xx < 0 || xx >= W) continue;
                                acc = fmaf(wk[ky * 3 + kx], plane[yy * W + xx], acc);
                            }
                        }
                    }
                    float uv = u[(((size_t)b * C + co) * H + y) * W + x];
                    h[(((size_t)b * C + co) * H + y) * W + x] = uv * (1.0f - r) + acc * r;
                }
}

/*
 * Multi-scale residual loop shared by VectorQuantizer2 (quant.py:88-118, :196-221) and
 * LFQ (lookup_free_quantize.py:176-201, :362-378).
 *   mode 0: VQ2 using_znorm=True   (argmax of normalised-row . normalised-code; gathers RAW code rows)
 *   mode 1: VQ2 using_znorm=False  (argmin (zz+ee) - 2 dot on raw rows/codes)
 *   mode 2: BSQ / LFQ              (bit c = pooled_c > 0; idx = sum bit_c 2^c; code = +-scaler[si])
 * f       [B,C,H,W]   input (for LFQ: already channel-normalised by the caller)
 * E       [V,C]       raw codebook (modes 0/1)
 * phi_w   [K,C,C,3,3], phi_b [K,C], phi_map [SN] -> which Phi each scale uses; r = resi ratio
 *                     (phi_map[si] < 0 means identity, quant_resi ~ 0)
 * outputs idx_all  [sum_si B*pn^2]   int64, scale-major then (b, y, x)
 *         u_all    [SN,B,C,H,W]      upsampled gathered codes (Phi input)
 *         h_all    [SN,B,C,H,W]      Phi output h_si  (unmasked)
 *         f_rest   [B,C,H,W]         final residual
 *         margin   [sum_si B*pn^2]   best - runner-up score gap (modes 0/1; may be NULL)
 */
void xqo_multiscale(const float *f, int B, int C, int H, int W, const float *E, int V, int mode,
                    const int32_t *patch_nums, int SN, const float *scaler,
                    const float *phi_w, const float *phi_b, const int32_t *phi_map, float r,
                    int64_t *idx_all, float *u_all, float *h_all, float *f_rest, float *margin) {
    size_t plane = (size_t)B * C * H * W;
    memcpy(f_rest, f, sizeof(float) * plane);
    float *En = NULL;
    if (mode == 0) {
        En = (float *)malloc(sizeof(float) * (size_t)V * C);
        xqo_l2norm_rows(E, V, C, En, NULL);
    }
    size_t off = 0;
    for (int si = 0; si < SN; ++si) {
        int P = patch_nums[si];
        int64_t n = (int64_t)B * P * P;
        float *rows = (float *)malloc(sizeof(float) * (size_t)n * C);
        float *gath = (float *)malloc(sizeof(float) * (size_t)n * C);
        xqo_area_pool_rows(f_rest, B, C, H, W, P, rows);
        int64_t *idx = idx_all + off;
        if (mode == 2) {
            for (int64_t i = 0; i < n; ++i) {
                int64_t code = 0;
                for (int c = 0; c < C; ++c) {
                    int bit = rows[i * C + c] > 0.0f;
                    code |= ((int64_t)bit) << c;
                    gath[i * C + c] = bit ? scaler[si] : -scaler[si];
                }
                idx[i] = code;
                if (margin) margin[off + i] = 0.0f;
            }
        } else {
            float *best = (float *)malloc(sizeof(float) * (size_t)n);
            float *second = (float *)malloc(sizeof(float) * (size_t)n);
            if (mode == 0) {
                float *rn = (float *)malloc(sizeof(float) * (size_t)n * C);
                xqo_l2norm_rows(rows, n, C, rn, NULL);
                xqo_search(rn, n, En, V, C, 1, idx, best, second);
                free(rn);
            } else {
                xqo_search(rows, n, E, V, C, 0, idx, best, second);
            }
            for (int64_t i = 0; i < n; ++i) {
                memcpy(gath + i * C, E + (size_t)idx[i] * C, sizeof(float) * C);
                if (margin) margin[off + i] = fabsf(best[i] - second[i]);
            }
            free(best); free(second);
        }
        float *u = u_all + (size_t)si * plane;
        float *h = h_all + (size_t)si * plane;
        xqo_bicubic_up(gath, B, C, P, H, W, u);
        if (phi_map[si] >= 0) {
            int k = phi_map[si];
            xqo_phi(u, B, C, H, W, phi_w + (size_t)k * C * C * 9, phi_b + (size_t)k * C, r, h);
        } else {
            memcpy(h, u, sizeof(float) * plane);
        }
        for (size_t i = 0; i < plane; ++i) f_rest[i] = f_rest[i] - h[i];
        free(rows); free(gath);
        off += (size_t)n;
    }
    free(En);
}

int xqo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void xqo_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
