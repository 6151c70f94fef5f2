// vit_kernels.cu -- HBM-bound glue kernels of the ViT encoder/decoder blocks (sm_100a).
//
// The reference block (dino_enc/vision_transformer.py:336-339) is
//     x = x + drop_path(ls1(attn(norm1(x))));   x = x + drop_path(ls2(mlp(norm2(x))))
// with the residual stream in fp32 and GEMM operands in bf16 under autocast.  Eager PyTorch spends
// one kernel per arrow (LayerNorm, cast, LayerScale mul, DropPath mul, add, GELU ...), each a full
// HBM round trip.  Here the whole non-GEMM glue between two GEMMs is ONE pass:
//
//   residual_ln_fwd : x_new = x + rowscale[b] * gamma_ls[d] * branch[m,d]      (fp32 stream, bf16 branch)
//                     y     = LayerNorm(x_new) * w + b  -> bf16 (next GEMM operand), mean / rstd saved
//   residual_ln_bwd : G = g_xnew + LN^T(g_y);  g_branch = G * rowscale * gamma_ls -> bf16;
//                     d ln_w, d ln_b, d gamma_ls column sums (per-CTA partials, reduced deterministically)
//   gelu_fwd / gelu_bwd : exact (erf) GELU on bf16, 16-byte vectors
//
// Algorithmic bytes per element (row x channel): fwd 4 (x) + 2 (branch) + 4 (x_new) + 2 (y) = 12 B;
// bwd 4 (g_xnew) + 2 (g_y) + 4 (x_new) + 2 (branch) + 4 (G) + 2 (g_branch) = 18 B.
// These TUs do not carry index decisions, so they are built with the default -fmad=true.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/xqb200.h"

namespace xqv {

constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct bf16x4 { __nv_bfloat162 a, b; };

__device__ __forceinline__ float4 load_bf16x4(const __nv_bfloat16 *p) {
    bf16x4 v = *reinterpret_cast<const bf16x4 *>(p);
    float2 lo = __bfloat1622float2(v.a), hi = __bfloat1622float2(v.b);
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ void store_bf16x4(__nv_bfloat16 *p, float4 f) {
    bf16x4 v;
    v.a = __floats2bfloat162_rn(f.x, f.y);
    v.b = __floats2bfloat162_rn(f.z, f.w);
    *reinterpret_cast<bf16x4 *>(p) = v;
}

// One warp handles FR = 2 rows (loads of both issued before any arithmetic); NV = D / 128 float4 chunks per lane.
//   x_new = x + rowscale * gamma_ls * (branch + branch_bias)
constexpr int FR = 2;
template <int NV>
__global__ void __launch_bounds__(THREADS)
residual_ln_fwd_kernel(const float *__restrict__ x, const __nv_bfloat16 *__restrict__ branch,
                       const float *__restrict__ branch_bias, const float *__restrict__ ls_gamma,
                       const float *__restrict__ rowscale, int rows_per_sample, const float *__restrict__ ln_w,
                       const float *__restrict__ ln_b, float eps, int M, float *__restrict__ x_out,
                       __nv_bfloat16 *__restrict__ y, float *__restrict__ mean_out, float *__restrict__ rstd_out) {
    constexpr int D = NV * 128;
    const int lane = threadIdx.x & 31;
    const int rbase = (blockIdx.x * WARPS + (threadIdx.x >> 5)) * FR;
    if (rbase >= M) return;
    float4 v[FR][NV], bv[FR][NV];
    bool ok[FR];
#pragma unroll
    for (int u = 0; u < FR; ++u) {
        ok[u] = rbase + u < M;
        const size_t base = (size_t)(ok[u] ? rbase + u : rbase) * D;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int col = (i * 32 + lane) * 4;
            v[u][i] = *reinterpret_cast<const float4 *>(x + base + col);
            bv[u][i] = branch ? load_bf16x4(branch + base + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
#pragma unroll
    for (int u = 0; u < FR; ++u) {
        if (!ok[u]) continue;
        const int row = rbase + u;
        const size_t base = (size_t)row * D;
        if (branch) {
            const float s = rowscale ? rowscale[row / rows_per_sample] : 1.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int col = (i * 32 + lane) * 4;
                float4 b = bv[u][i];
                if (branch_bias) {
                    float4 bb = *reinterpret_cast<const float4 *>(branch_bias + col);
                    b.x += bb.x; b.y += bb.y; b.z += bb.z; b.w += bb.w;
                }
                float4 g = ls_gamma ? *reinterpret_cast<const float4 *>(ls_gamma + col) : make_float4(1.f, 1.f, 1.f, 1.f);
                v[u][i].x += s * g.x * b.x; v[u][i].y += s * g.y * b.y; v[u][i].z += s * g.z * b.z; v[u][i].w += s * g.w * b.w;
            }
        }
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) sum += v[u][i].x + v[u][i].y + v[u][i].z + v[u][i].w;
        const float mean = warp_sum(sum) * (1.f / D);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float a = v[u][i].x - mean, b = v[u][i].y - mean, c = v[u][i].z - mean, d = v[u][i].w - mean;
            sq += a * a + b * b + c * c + d * d;
        }
        const float rstd = rsqrtf(warp_sum(sq) * (1.f / D) + eps);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int col = (i * 32 + lane) * 4;
            if (x_out) *reinterpret_cast<float4 *>(x_out + base + col) = v[u][i];
            if (y) {
                float4 w = *reinterpret_cast<const float4 *>(ln_w + col);
                float4 b = *reinterpret_cast<const float4 *>(ln_b + col);
                float4 o;
                o.x = (v[u][i].x - mean) * rstd * w.x + b.x; o.y = (v[u][i].y - mean) * rstd * w.y + b.y;
                o.z = (v[u][i].z - mean) * rstd * w.z + b.z; o.w = (v[u][i].w - mean) * rstd * w.w + b.w;
                store_bf16x4(y + base + col, o);
            }
        }
        if (lane == 0 && mean_out) { mean_out[row] = mean; rstd_out[row] = rstd; }
    }
}

// Backward.  Persistent grid; each warp walks rows with a grid stride, RPW = 2 rows per iteration so that twice as
// many loads are in flight (the kernel is bound by global-load latency: long-scoreboard stalls dominate at 16 warps/SM).
// The four column partial sums (d ln_w, d ln_b, sum G*s*branch, sum G*s) live in a per-warp SHARED-MEMORY accumulator
// (each lane owns its columns -> plain load-add-store, no atomics).  CTA partials -> part[blockIdx][4][D].
//   g_xout may be null (no later residual gradient), g_y may be null (LN output unused).
constexpr int NACC = 4;
constexpr int RPW = 2;
template <int NV>
__global__ void __launch_bounds__(THREADS, 2)
residual_ln_bwd_kernel(const float *__restrict__ g_xout, const __nv_bfloat16 *__restrict__ g_y,
                       const float *__restrict__ x_out, const float *__restrict__ mean_in,
                       const float *__restrict__ rstd_in, const float *__restrict__ ln_w,
                       const __nv_bfloat16 *__restrict__ branch, const float *__restrict__ branch_bias,
                       const float *__restrict__ ls_gamma, const float *__restrict__ rowscale, int rows_per_sample,
                       int M, float *__restrict__ g_x, __nv_bfloat16 *__restrict__ g_branch,
                       float *__restrict__ part) {
    constexpr int D = NV * 128;
    extern __shared__ __align__(16) float acc_s[];  // [WARPS][NACC][D]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *acc = acc_s + (size_t)warp * NACC * D;
    for (int i = lane * 4; i < NACC * D; i += 128) *reinterpret_cast<float4 *>(acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    const int stride = gridDim.x * WARPS;
    for (int row0 = blockIdx.x * WARPS + warp; row0 < M; row0 += stride * RPW) {
        int rows[RPW];
        bool ok[RPW];
        float mean[RPW], rstd[RPW], sc[RPW], c1[RPW], c2[RPW];
        float4 xh[RPW][NV], gy[RPW][NV];
#pragma unroll
        for (int u = 0; u < RPW; ++u) {
            rows[u] = row0 + u * stride;
            ok[u] = rows[u] < M;
            const int r = ok[u] ? rows[u] : row0;
            mean[u] = mean_in[r];
            rstd[u] = rstd_in[r];
            sc[u] = (branch && rowscale) ? rowscale[r / rows_per_sample] : 1.f;
            c1[u] = c2[u] = 0.f;
        }
        // phase 1: issue the loads of both rows back to back, then the per-row statistics
#pragma unroll
        for (int u = 0; u < RPW; ++u) {
            const size_t base = (size_t)(ok[u] ? rows[u] : row0) * D;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int col = (i * 32 + lane) * 4;
                xh[u][i] = *reinterpret_cast<const float4 *>(x_out + base + col);
                gy[u][i] = g_y ? load_bf16x4(g_y + base + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < RPW; ++u) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int col = (i * 32 + lane) * 4;
                float4 xv = xh[u][i];
                xh[u][i] = make_float4((xv.x - mean[u]) * rstd[u], (xv.y - mean[u]) * rstd[u], (xv.z - mean[u]) * rstd[u],
                                       (xv.w - mean[u]) * rstd[u]);
                if (g_y && ok[u]) {
                    float4 g = gy[u][i];
                    float4 w = *reinterpret_cast<const float4 *>(ln_w + col);
                    float4 a0 = *reinterpret_cast<float4 *>(acc + 0 * D + col);
                    float4 a1 = *reinterpret_cast<float4 *>(acc + 1 * D + col);
                    a0.x += g.x * xh[u][i].x; a0.y += g.y * xh[u][i].y; a0.z += g.z * xh[u][i].z; a0.w += g.w * xh[u][i].w;
                    a1.x += g.x; a1.y += g.y; a1.z += g.z; a1.w += g.w;
                    *reinterpret_cast<float4 *>(acc + 0 * D + col) = a0;
                    *reinterpret_cast<float4 *>(acc + 1 * D + col) = a1;
                    gy[u][i] = make_float4(g.x * w.x, g.y * w.y, g.z * w.z, g.w * w.w);
                    c1[u] += gy[u][i].x + gy[u][i].y + gy[u][i].z + gy[u][i].w;
                    c2[u] += gy[u][i].x * xh[u][i].x + gy[u][i].y * xh[u][i].y + gy[u][i].z * xh[u][i].z + gy[u][i].w * xh[u][i].w;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < RPW; ++u) {
            c1[u] = warp_sum(c1[u]) * (1.f / D);
            c2[u] = warp_sum(c2[u]) * (1.f / D);
        }
        // phase 2
#pragma unroll
        for (int u = 0; u < RPW; ++u) {
            if (!ok[u]) continue;
            const size_t base = (size_t)rows[u] * D;
            float4 rr[NV], bv[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int col = (i * 32 + lane) * 4;
                rr[i] = g_xout ? *reinterpret_cast<const float4 *>(g_xout + base + col) : make_float4(0.f, 0.f, 0.f, 0.f);
                bv[i] = branch ? load_bf16x4(branch + base + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int col = (i * 32 + lane) * 4;
                float4 G;
                G.x = rstd[u] * (gy[u][i].x - c1[u] - xh[u][i].x * c2[u]) + rr[i].x;
                G.y = rstd[u] * (gy[u][i].y - c1[u] - xh[u][i].y * c2[u]) + rr[i].y;
                G.z = rstd[u] * (gy[u][i].z - c1[u] - xh[u][i].z * c2[u]) + rr[i].z;
                G.w = rstd[u] * (gy[u][i].w - c1[u] - xh[u][i].w * c2[u]) + rr[i].w;
                if (g_x) *reinterpret_cast<float4 *>(g_x + base + col) = G;
                if (branch) {
                    float4 gm = ls_gamma ? *reinterpret_cast<const float4 *>(ls_gamma + col) : make_float4(1.f, 1.f, 1.f, 1.f);
                    float4 a2 = *reinterpret_cast<float4 *>(acc + 2 * D + col);
                    float4 a3 = *reinterpret_cast<float4 *>(acc + 3 * D + col);
                    float4 Gs = make_float4(G.x * sc[u], G.y * sc[u], G.z * sc[u], G.w * sc[u]);
                    a2.x += Gs.x * bv[i].x; a2.y += Gs.y * bv[i].y; a2.z += Gs.z * bv[i].z; a2.w += Gs.w * bv[i].w;
                    a3.x += Gs.x; a3.y += Gs.y; a3.z += Gs.z; a3.w += Gs.w;
                    *reinterpret_cast<float4 *>(acc + 2 * D + col) = a2;
                    *reinterpret_cast<float4 *>(acc + 3 * D + col) = a3;
                    if (g_branch) store_bf16x4(g_branch + base + col, make_float4(Gs.x * gm.x, Gs.y * gm.y, Gs.z * gm.z, Gs.w * gm.w));
                }
            }
        }
    }
    __syncthreads();
    float *outp = part + (size_t)blockIdx.x * NACC * D;
    for (int e = threadIdx.x; e < NACC * D; e += THREADS) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) a += acc_s[(size_t)w * NACC * D + e];
        outp[e] = a;
    }
}

// sums the CTA partials and finishes the four vectors:
//   d ln_w = P0 ; d ln_b = P1 ; d gamma_ls = P2 + bias * P3 ; d branch_bias = gamma_ls * P3
// one thread per column, 8 independent loads in flight per accumulator (the serial version was latency-bound)
__global__ void reduce_parts_kernel(const float *__restrict__ part, int nblocks, int D, const float *__restrict__ ls_gamma,
                                    const float *__restrict__ branch_bias, float *__restrict__ g_ln_w,
                                    float *__restrict__ g_ln_b, float *__restrict__ g_ls_gamma,
                                    float *__restrict__ g_branch_bias) {
    __shared__ float sh[NACC][8][32];
    const int lane = threadIdx.x & 31, sub = threadIdx.x >> 5;       // 8 sub-ranges of the block list per column
    const int d = blockIdx.x * 32 + lane;
    float p[NACC] = {0.f, 0.f, 0.f, 0.f};
    if (d < D) {
        for (int b = sub; b < nblocks; b += 8) {
#pragma unroll
            for (int q = 0; q < NACC; ++q) p[q] += part[((size_t)b * NACC + q) * D + d];
        }
    }
#pragma unroll
    for (int q = 0; q < NACC; ++q) sh[q][sub][lane] = p[q];
    __syncthreads();
    if (sub == 0 && d < D) {
#pragma unroll
        for (int q = 0; q < NACC; ++q) {
            float a = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) a += sh[q][w][lane];
            p[q] = a;
        }
        if (g_ln_w) g_ln_w[d] = p[0];
        if (g_ln_b) g_ln_b[d] = p[1];
        if (g_ls_gamma) g_ls_gamma[d] = p[2] + (branch_bias ? branch_bias[d] * p[3] : 0.f);
        if (g_branch_bias) g_branch_bias[d] = (ls_gamma ? ls_gamma[d] : 1.f) * p[3];
    }
}

// erf via Abramowitz-Stegun 7.1.28:  erf(a) = 1 - (1 + a1 a + ... + a6 a^6)^-16,  |abs err| <= 3e-7 (far below the bf16
// output rounding).  ONE special-function op (the reciprocal) per element -- the 7.1.26 form used before needed exp and
// reciprocal, and the kernel was bound by the 16-lane/clk MUFU pipe, not by HBM.
__device__ __forceinline__ float rcp_fast(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float erf_as(float z) {
    const float a = fabsf(z);
    float p = fmaf(a, 0.0000430638f, 0.0002765672f);
    p = fmaf(p, a, 0.0001520143f);
    p = fmaf(p, a, 0.0092705272f);
    p = fmaf(p, a, 0.0422820123f);
    p = fmaf(p, a, 0.0705230784f);
    p = fmaf(p, a, 1.0f);
    p = p * p; p = p * p; p = p * p; p = p * p;          // ^16 (overflows to +inf for |z| > ~9 -> erf = 1, as it should)
    return copysignf(1.0f - rcp_fast(p), z);
}
__device__ __forceinline__ float gelu_f(float x) {
    const float h = 0.5f * x;
    return fmaf(h, erf_as(x * 0.70710678118654752f), h);
}
__device__ __forceinline__ float dgelu_f(float x) {
    const float e = ex2_fast(x * x * -0.72134752044448170f);      // exp(-x^2 / 2)
    return fmaf(0.5f, erf_as(x * 0.70710678118654752f), 0.5f) + x * 0.3989422804014327f * e;
}

// y = gelu(x + bias).  A thread owns column chunk c (8 bf16 = 16 B) and walks rows with a grid stride, RU rows
// per iteration so that RU independent 16-byte loads are in flight (one load per iteration leaves HBM idle).
constexpr int GELU_RU = 4;
__device__ __forceinline__ void load_bias8(const float *bias, int c, float (&bb)[8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) bb[k] = 0.f;
    if (bias) {
        float4 b0 = *reinterpret_cast<const float4 *>(bias + c * 8), b1 = *reinterpret_cast<const float4 *>(bias + c * 8 + 4);
        bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
    }
}
__global__ void gelu_fwd_kernel(const uint4 *__restrict__ x, const float *__restrict__ bias, uint4 *__restrict__ y,
                                int M, int C8) {
    // NON-persistent: one CTA per GELU_RU rows.  tools/mb/stream_mb.cu on B200: fresh small CTAs stream at 6.1 TB/s where
    // the persistent grid-stride form of the same loop reaches 5.4 (lock-step load/store phases + SM imbalance).
    const int row0 = blockIdx.x * GELU_RU;
    for (int c = threadIdx.x; c < C8; c += blockDim.x) {
        float bb[8];
        load_bias8(bias, c, bb);
        uint4 v[GELU_RU];
#pragma unroll
        for (int u = 0; u < GELU_RU; ++u)
            if (row0 + u < M) v[u] = x[(size_t)(row0 + u) * C8 + c];
#pragma unroll
        for (int u = 0; u < GELU_RU; ++u) {
            if (row0 + u >= M) continue;
            __nv_bfloat162 *p = reinterpret_cast<__nv_bfloat162 *>(&v[u]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float2 f = __bfloat1622float2(p[k]);
                p[k] = __floats2bfloat162_rn(gelu_f(f.x + bb[2 * k]), gelu_f(f.y + bb[2 * k + 1]));
            }
            y[(size_t)(row0 + u) * C8 + c] = v[u];
        }
    }
}

// gx = gy * gelu'(x + bias); column sums of gx (= d bias) accumulate per thread, one atomicAdd per column
// per CTA at the end (g_bias must be zeroed by the caller).
constexpr int GELU_BWD_RU = 2;
__global__ void gelu_bwd_kernel(const uint4 *__restrict__ x, const float *__restrict__ bias, const uint4 *__restrict__ gy,
                                uint4 *__restrict__ gx, float *__restrict__ g_bias, int M, int C8) {
    // persistent (the column sums stay in registers, one atomicAdd per column per CTA) and SOFTWARE-PIPELINED: the loads
    // of iteration i+1 are issued before the math / stores of iteration i, so a warp always has loads in flight.
    constexpr int RU = GELU_BWD_RU;
    const int step = gridDim.x * RU;
    for (int c = threadIdx.x; c < C8; c += blockDim.x) {
        float bb[8], acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        load_bias8(bias, c, bb);
        uint4 v[RU], g[RU], vn[RU], gn[RU];
        int row0 = blockIdx.x * RU;
#pragma unroll
        for (int u = 0; u < RU; ++u)
            if (row0 + u < M) { v[u] = x[(size_t)(row0 + u) * C8 + c]; g[u] = gy[(size_t)(row0 + u) * C8 + c]; }
        for (; row0 < M; row0 += step) {
            const int nxt = row0 + step;
#pragma unroll
            for (int u = 0; u < RU; ++u)
                if (nxt + u < M) { vn[u] = x[(size_t)(nxt + u) * C8 + c]; gn[u] = gy[(size_t)(nxt + u) * C8 + c]; }
#pragma unroll
            for (int u = 0; u < RU; ++u) {
                if (row0 + u >= M) continue;
                __nv_bfloat162 *p = reinterpret_cast<__nv_bfloat162 *>(&v[u]);
                __nv_bfloat162 *q = reinterpret_cast<__nv_bfloat162 *>(&g[u]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float2 f = __bfloat1622float2(p[k]), h = __bfloat1622float2(q[k]);
                    float r0 = h.x * dgelu_f(f.x + bb[2 * k]), r1 = h.y * dgelu_f(f.y + bb[2 * k + 1]);
                    acc[2 * k] += r0; acc[2 * k + 1] += r1;
                    p[k] = __floats2bfloat162_rn(r0, r1);
                }
                gx[(size_t)(row0 + u) * C8 + c] = v[u];
            }
#pragma unroll
            for (int u = 0; u < RU; ++u) { v[u] = vn[u]; g[u] = gn[u]; }
        }
        if (g_bias) {
#pragma unroll
            for (int k = 0; k < 8; ++k) atomicAdd(g_bias + c * 8 + k, acc[k]);
        }
    }
}

// dq, dk, dv: each [M, C] bf16 dense (what the SDPA backward returns for q/k/v views of a packed
// [M, 3C] projection) -> dqkv [M, 3C], and (optionally) g_bias [3C] = column sums of dqkv = the gradient of the
// qkv projection bias, which autograd would otherwise compute with one more full pass over dqkv.
// A thread owns one 16-byte column chunk of the packed row and walks rows with a grid stride, PACK_RU rows in flight.
constexpr int PACK_RU = 4;
__global__ void pack_qkv_kernel(const uint4 *__restrict__ dq, const uint4 *__restrict__ dk, const uint4 *__restrict__ dv,
                                uint4 *__restrict__ out, float *__restrict__ g_bias, int M, int C8) {
    const int per_row = 3 * C8, step = gridDim.x * PACK_RU;
    for (int c = threadIdx.x; c < per_row; c += blockDim.x) {
        const int which = c / C8, cc = c - which * C8;
        const uint4 *src = which == 0 ? dq : (which == 1 ? dk : dv);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        uint4 v[PACK_RU], vn[PACK_RU];
        int row0 = blockIdx.x * PACK_RU;
#pragma unroll
        for (int u = 0; u < PACK_RU; ++u)
            if (row0 + u < M) v[u] = src[(size_t)(row0 + u) * C8 + cc];
        for (; row0 < M; row0 += step) {            // software-pipelined like gelu_bwd_kernel
            const int nxt = row0 + step;
#pragma unroll
            for (int u = 0; u < PACK_RU; ++u)
                if (nxt + u < M) vn[u] = src[(size_t)(nxt + u) * C8 + cc];
#pragma unroll
            for (int u = 0; u < PACK_RU; ++u) {
                if (row0 + u >= M) continue;
                out[(size_t)(row0 + u) * per_row + c] = v[u];
                if (g_bias) {
                    const __nv_bfloat162 *p = reinterpret_cast<const __nv_bfloat162 *>(&v[u]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float2 f = __bfloat1622float2(p[k]);
                        acc[2 * k] += f.x; acc[2 * k + 1] += f.y;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < PACK_RU; ++u) v[u] = vn[u];
        }
        if (g_bias) {
#pragma unroll
            for (int k = 0; k < 8; ++k) atomicAdd(g_bias + c * 8 + k, acc[k]);
        }
    }
}

// Patch embedding as a GEMM (timm PatchEmbed: Conv2d(kernel = stride = p) -> flatten -> NLC, vision_transformer.py:
// patch_embed): non-overlapping patches make im2col a pure permutation, so the conv is
//   tokens[B*gh*gw, D] = patches[B*gh*gw, Cin*p*p] @ W[D, Cin*p*p]^T + b .
// This kernel writes `patches` in bf16 (GEMM operand) from the fp32 NCHW image: 4 pixels (16 B in, 8 B out) per thread,
// output-major indexing -> fully coalesced stores, 64-byte-segment loads.  (cuDNN's implicit-GEMM for Cin = 3 pads the
// channel dimension to 8 and adds two layout conversions: 4.3 ms per step at B = 256 vs 0.1 ms here + a 0.06 ms GEMM.)
__global__ void patchify_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ out, int Cin, int H, int W, int p,
                                size_t total4) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int p4 = p >> 2, gw = W / p, gh = H / p;
    size_t t = i;
    const int kx4 = (int)(t % p4); t /= p4;
    const int ky = (int)(t % p); t /= p;
    const int c = (int)(t % Cin); t /= Cin;
    const int px = (int)(t % gw); t /= gw;
    const int py = (int)(t % gh); t /= gh;
    const size_t b = t;
    const float4 v = *reinterpret_cast<const float4 *>(x + ((b * Cin + c) * H + (size_t)py * p + ky) * W + (size_t)px * p + kx4 * 4);
    store_bf16x4(out + i * 4, v);
}

// persistent grids = (SM count) x (CTAs of this kernel that are actually co-resident on one SM)
template <typename K>
static int persistent_grid(K kernel, int threads, size_t smem = 0) {
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    return sms * per_sm;
}

static int bwd_grid() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms * 2;  // 2 CTAs / SM (shared-memory accumulators: 96 KB per CTA at D=768)
}

}  // namespace xqv

using namespace xqv;

#define XQV_DISPATCH(D, ...)                  \
    switch (D) {                               \
        case 384: { constexpr int NV = 3; __VA_ARGS__; break; }   \
        case 768: { constexpr int NV = 6; __VA_ARGS__; break; }   \
        case 1024: { constexpr int NV = 8; __VA_ARGS__; break; }  \
        default: return XQ_ERR_UNSUPPORTED;    \
    }

extern "C" {

size_t xq_vit_ln_bwd_workspace_bytes(int D) { return sizeof(float) * (size_t)bwd_grid() * NACC * D; }

int xq_vit_residual_ln_fwd(const float *x, const void *branch, const float *branch_bias, const float *ls_gamma,
                           const float *rowscale, int rows_per_sample, const float *ln_w, const float *ln_b, float eps,
                           int M, int D, float *x_out, void *y, float *mean, float *rstd, void *stream) {
    if (!x || M <= 0 || (y && (!ln_w || !ln_b)) || (!x_out && !y)) return XQ_ERR_ARG;
    if (branch && rowscale && rows_per_sample <= 0) return XQ_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int grid = (M + WARPS * FR - 1) / (WARPS * FR);
    XQV_DISPATCH(D, (residual_ln_fwd_kernel<NV><<<grid, THREADS, 0, st>>>(
                        x, (const __nv_bfloat16 *)branch, branch_bias, ls_gamma, rowscale, rows_per_sample, ln_w, ln_b,
                        eps, M, x_out, (__nv_bfloat16 *)y, mean, rstd)));
    return cudaGetLastError() == cudaSuccess ? XQ_OK : XQ_ERR_CUDA;
}

int xq_vit_residual_ln_bwd(const float *g_xout, const void *g_y, const float *x_out, const float *mean,
                           const float *rstd, const float *ln_w, const void *branch, const float *branch_bias,
                           const float *ls_gamma, const float *rowscale, int rows_per_sample, int M, int D, float *g_x,
                           void *g_branch, float *g_ln_w, float *g_ln_b, float *g_ls_gamma, float *g_branch_bias,
                           void *workspace, size_t workspace_bytes, void *stream) {
    if (!x_out || !mean || !rstd || M <= 0 || !workspace) return XQ_ERR_ARG;
    if (g_y && !ln_w) return XQ_ERR_ARG;
    const int grid = bwd_grid();
    if (workspace_bytes < sizeof(float) * (size_t)grid * NACC * D) return XQ_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float *part = (float *)workspace;
    const size_t smem = sizeof(float) * (size_t)WARPS * NACC * D;
    XQV_DISPATCH(D, {
        if (cudaFuncSetAttribute(residual_ln_bwd_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return XQ_ERR_CUDA;
        residual_ln_bwd_kernel<NV><<<grid, THREADS, smem, st>>>(
            g_xout, (const __nv_bfloat16 *)g_y, x_out, mean, rstd, ln_w, (const __nv_bfloat16 *)branch, branch_bias,
            ls_gamma, rowscale, rows_per_sample, M, g_x, (__nv_bfloat16 *)g_branch, part);
    });
    if (cudaGetLastError() != cudaSuccess) return XQ_ERR_CUDA;
    reduce_parts_kernel<<<(D + 31) / 32, 256, 0, st>>>(part, grid, D, ls_gamma, branch_bias, g_ln_w, g_ln_b,
                                                        branch ? g_ls_gamma : nullptr, branch ? g_branch_bias : nullptr);
    return cudaGetLastError() == cudaSuccess ? XQ_OK : XQ_ERR_CUDA;
}

int xq_vit_pack_qkv(const void *dq, const void *dk, const void *dv, void *dqkv, float *g_bias, size_t M, int C,
                    void *stream) {
    if (!dq || !dk || !dv || !dqkv || M == 0 || M > 0x7fffffffu || C <= 0 || (C & 7)) return XQ_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int C8 = C / 8, per_row = 3 * C8;
    const int threads = per_row >= 288 ? 288 : (per_row >= 192 ? 192 : 96);
    const int rows4 = (int)((M + PACK_RU - 1) / PACK_RU);
    int grid = persistent_grid(pack_qkv_kernel, threads);
    if (rows4 < grid) grid = rows4;
    if (g_bias && cudaMemsetAsync(g_bias, 0, sizeof(float) * 3 * (size_t)C, st) != cudaSuccess) return XQ_ERR_CUDA;
    pack_qkv_kernel<<<grid, threads, 0, st>>>((const uint4 *)dq, (const uint4 *)dk, (const uint4 *)dv, (uint4 *)dqkv,
                                              g_bias, (int)M, C8);
    return cudaGetLastError() == cudaSuccess ? XQ_OK : XQ_ERR_CUDA;
}

int xq_vit_patchify(const float *x, void *patches, int B, int Cin, int H, int W, int p, void *stream) {
    if (!x || !patches || B <= 0 || Cin <= 0 || H <= 0 || W <= 0 || p <= 0) return XQ_ERR_ARG;
    if ((p & 3) || H % p || W % p) return XQ_ERR_UNSUPPORTED;
    const size_t total4 = (size_t)B * Cin * H * W / 4;
    patchify_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16 *)patches, Cin, H, W,
                                                                                       p, total4);
    return cudaGetLastError() == cudaSuccess ? XQ_OK : XQ_ERR_CUDA;
}

int xq_vit_gelu_fwd(const void *x, const float *bias, void *y, int M, int C, void *stream) {
    if (!x || !y || M <= 0 || C <= 0 || (C & 7)) return XQ_ERR_ARG;
    int C8 = C / 8;
    int threads = C8 >= 384 ? 384 : (C8 >= 192 ? 192 : 128);
    int grid = (M + GELU_RU - 1) / GELU_RU;
    gelu_fwd_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>((const uint4 *)x, bias, (uint4 *)y, M, C8);
    return cudaGetLastError() == cudaSuccess ? XQ_OK : XQ_ERR_CUDA;
}

int xq_vit_gelu_bwd(const void *x, const float *bias, const void *gy, void *gx, float *g_bias, int M, int C, void *stream) {
    if (!x || !gy || !gx || M <= 0 || C <= 0 || (C & 7)) return XQ_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int C8 = C / 8;
    int threads = C8 >= 384 ? 384 : (C8 >= 192 ? 192 : 128);
    int grid = persistent_grid(gelu_bwd_kernel, threads);
    if ((M + GELU_BWD_RU - 1) / GELU_BWD_RU < grid) grid = (M + GELU_BWD_RU - 1) / GELU_BWD_RU;
    if (g_bias && cudaMemsetAsync(g_bias, 0, sizeof(float) * (size_t)C, st) != cudaSuccess) return XQ_ERR_CUDA;
    gelu_bwd_kernel<<<grid, threads, 0, st>>>((const uint4 *)x, bias, (const uint4 *)gy, (uint4 *)gx, g_bias, M, C8);
    return cudaGetLastError() == cudaSuccess ? XQ_OK : XQ_ERR_CUDA;
}

}  // extern "C"
