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

// One warp per row; NV = D / 128 float4 chunks per lane.
//   x_new = x + rowscale * gamma_ls * (branch + branch_bias)
template <int NV>
__global__ void __launch_bounds__(THREADS)
residual_ln_fwd_kernel(const float *__restrict__ x, const __nv_bfloat16 *__restrict__ branch,
                       const float *__restrict__ branch_bias, const float *__restrict__ ls_gamma,
                       const float *__restrict__ rowscale, int rows_per_sample, const float *__restrict__ ln_w,
                       const float *__restrict__ ln_b, float eps, int M, float *__restrict__ x_out,
                       __nv_bfloat16 *__restrict__ y, float *__restrict__ mean_out, float *__restrict__ rstd_out) {
    constexpr int D = NV * 128;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (row >= M) return;
    const size_t base = (size_t)row * D;
    float4 v[NV];
    const float s = (branch && rowscale) ? rowscale[row / rows_per_sample] : 1.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int col = (i * 32 + lane) * 4;
        v[i] = *reinterpret_cast<const float4 *>(x + base + col);
        if (branch) {
            float4 b = load_bf16x4(branch + base + col);
            if (branch_bias) {
                float4 bb = *reinterpret_cast<const float4 *>(branch_bias + col);
                b.x += bb.x; b.y += bb.y; b.z += bb.z; b.w += bb.w;
            }
            float4 g = ls_gamma ? *reinterpret_cast<const float4 *>(ls_gamma + col) : make_float4(1.f, 1.f, 1.f, 1.f);
            v[i].x += s * g.x * b.x; v[i].y += s * g.y * b.y; v[i].z += s * g.z * b.z; v[i].w += s * g.w * b.w;
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) sum += v[i].x + v[i].y + v[i].z + v[i].w;
    const float mean = warp_sum(sum) * (1.f / D);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        sq += a * a + b * b + c * c + d * d;
    }
    const float rstd = rsqrtf(warp_sum(sq) * (1.f / D) + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int col = (i * 32 + lane) * 4;
        if (x_out) *reinterpret_cast<float4 *>(x_out + base + col) = v[i];
        if (y) {
            float4 w = *reinterpret_cast<const float4 *>(ln_w + col);
            float4 b = *reinterpret_cast<const float4 *>(ln_b + col);
            float4 o;
            o.x = (v[i].x - mean) * rstd * w.x + b.x; o.y = (v[i].y - mean) * rstd * w.y + b.y;
            o.z = (v[i].z - mean) * rstd * w.z + b.z; o.w = (v[i].w - mean) * rstd * w.w + b.w;
            store_bf16x4(y + base + col, o);
        }
    }
    if (lane == 0 && mean_out) { mean_out[row] = mean; rstd_out[row] = rstd; }
}

// Backward.  Persistent grid; each warp walks rows with a grid stride.  The four column partial sums
// (d ln_w, d ln_b, sum G*s*branch, sum G*s) live in a per-warp SHARED-MEMORY accumulator (each lane owns
// its columns, so plain load-add-store, no atomics): registers stay low enough for 3 CTAs / SM, which is
// what keeps enough loads in flight to run at HBM speed.  CTA partials -> part[blockIdx][4][D].
//   g_xout may be null (no later residual gradient), g_y may be null (LN output unused).
constexpr int NACC = 4;
constexpr int NACC_S = 2;   // accumulators kept in shared memory (the other two stay in registers)
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
    extern __shared__ __align__(16) float acc_s[];  // [WARPS][NACC_S][D]  (sum G*s*branch, sum G*s)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *acc = acc_s + (size_t)warp * NACC_S * D;
    for (int i = lane * 4; i < NACC_S * D; i += 128) *reinterpret_cast<float4 *>(acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 aw[NV], ab[NV];   // d ln_w, d ln_b partials (registers)
#pragma unroll
    for (int i = 0; i < NV; ++i) { aw[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f); }
    __syncwarp();
    for (int row = blockIdx.x * WARPS + warp; row < M; row += gridDim.x * WARPS) {
        const size_t base = (size_t)row * D;
        const float mean = mean_in[row], rstd = rstd_in[row];
        const float s = (branch && rowscale) ? rowscale[row / rows_per_sample] : 1.f;
        float4 xh[NV], gy[NV];
        float c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int col = (i * 32 + lane) * 4;
            float4 xv = *reinterpret_cast<const float4 *>(x_out + base + col);
            xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
            if (g_y) {
                float4 g = load_bf16x4(g_y + base + col);
                float4 w = *reinterpret_cast<const float4 *>(ln_w + col);
                aw[i].x += g.x * xh[i].x; aw[i].y += g.y * xh[i].y; aw[i].z += g.z * xh[i].z; aw[i].w += g.w * xh[i].w;
                ab[i].x += g.x; ab[i].y += g.y; ab[i].z += g.z; ab[i].w += g.w;
                gy[i] = make_float4(g.x * w.x, g.y * w.y, g.z * w.z, g.w * w.w);
                c1 += gy[i].x + gy[i].y + gy[i].z + gy[i].w;
                c2 += gy[i].x * xh[i].x + gy[i].y * xh[i].y + gy[i].z * xh[i].z + gy[i].w * xh[i].w;
            } else {
                gy[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        c1 = warp_sum(c1) * (1.f / D);
        c2 = warp_sum(c2) * (1.f / D);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int col = (i * 32 + lane) * 4;
            float4 G;
            G.x = rstd * (gy[i].x - c1 - xh[i].x * c2); G.y = rstd * (gy[i].y - c1 - xh[i].y * c2);
            G.z = rstd * (gy[i].z - c1 - xh[i].z * c2); G.w = rstd * (gy[i].w - c1 - xh[i].w * c2);
            if (g_xout) {
                float4 r = *reinterpret_cast<const float4 *>(g_xout + base + col);
                G.x += r.x; G.y += r.y; G.z += r.z; G.w += r.w;
            }
            if (g_x) *reinterpret_cast<float4 *>(g_x + base + col) = G;
            if (branch) {
                float4 b = load_bf16x4(branch + base + col);
                float4 gm = ls_gamma ? *reinterpret_cast<const float4 *>(ls_gamma + col) : make_float4(1.f, 1.f, 1.f, 1.f);
                float4 a2 = *reinterpret_cast<float4 *>(acc + 0 * D + col);
                float4 a3 = *reinterpret_cast<float4 *>(acc + 1 * D + col);
                float4 Gs = make_float4(G.x * s, G.y * s, G.z * s, G.w * s);
                a2.x += Gs.x * b.x; a2.y += Gs.y * b.y; a2.z += Gs.z * b.z; a2.w += Gs.w * b.w;
                a3.x += Gs.x; a3.y += Gs.y; a3.z += Gs.z; a3.w += Gs.w;
                *reinterpret_cast<float4 *>(acc + 0 * D + col) = a2;
                *reinterpret_cast<float4 *>(acc + 1 * D + col) = a3;
                if (g_branch) store_bf16x4(g_branch + base + col, make_float4(Gs.x * gm.x, Gs.y * gm.y, Gs.z * gm.z, Gs.w * gm.w));
            }
        }
    }
    __syncthreads();
    float *outp = part + (size_t)blockIdx.x * NACC * D;
    for (int e = threadIdx.x; e < NACC_S * D; e += THREADS) {     // P2, P3 from the shared accumulators
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) a += acc_s[(size_t)w * NACC_S * D + e];
        outp[2 * D + e] = a;
    }
    __syncthreads();
    // P0, P1 from the register accumulators, staged through the (now free) shared buffer warp by warp
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int col = (i * 32 + lane) * 4;
        *reinterpret_cast<float4 *>(acc + 0 * D + col) = aw[i];
        *reinterpret_cast<float4 *>(acc + 1 * D + col) = ab[i];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 2 * D; e += THREADS) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) a += acc_s[(size_t)w * NACC_S * D + e];
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

// erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below the bf16 output rounding);
// e = exp(-z^2) is shared with the derivative (z = x / sqrt(2) -> e = exp(-x^2 / 2)).
__device__ __forceinline__ float erf_as(float z, float e) {
    float az = fabsf(z);
    float t = __fdividef(1.f, fmaf(0.3275911f, az, 1.f));
    float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
    float r = fmaf(-poly, e, 1.f);
    return copysignf(r, z);
}
__device__ __forceinline__ float gelu_f(float x) {
    float e = __expf(-0.5f * x * x);
    return 0.5f * x * (1.f + erf_as(x * 0.70710678118654752f, e));
}
__device__ __forceinline__ float dgelu_f(float x) {
    float e = __expf(-0.5f * x * x);
    return 0.5f * (1.f + erf_as(x * 0.70710678118654752f, e)) + x * 0.3989422804014327f * e;
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
    for (int c = threadIdx.x; c < C8; c += blockDim.x) {
        float bb[8];
        load_bias8(bias, c, bb);
        for (int row0 = blockIdx.x * GELU_RU; row0 < M; row0 += gridDim.x * GELU_RU) {
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
}

// gx = gy * gelu'(x + bias); column sums of gx (= d bias) accumulate per thread, one atomicAdd per column
// per CTA at the end (g_bias must be zeroed by the caller).
__global__ void gelu_bwd_kernel(const uint4 *__restrict__ x, const float *__restrict__ bias, const uint4 *__restrict__ gy,
                                uint4 *__restrict__ gx, float *__restrict__ g_bias, int M, int C8) {
    for (int c = threadIdx.x; c < C8; c += blockDim.x) {
        float bb[8], acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        load_bias8(bias, c, bb);
        for (int row0 = blockIdx.x * GELU_RU; row0 < M; row0 += gridDim.x * GELU_RU) {
            uint4 v[GELU_RU], g[GELU_RU];
#pragma unroll
            for (int u = 0; u < GELU_RU; ++u)
                if (row0 + u < M) { v[u] = x[(size_t)(row0 + u) * C8 + c]; g[u] = gy[(size_t)(row0 + u) * C8 + c]; }
#pragma unroll
            for (int u = 0; u < GELU_RU; ++u) {
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
        }
        if (g_bias) {
#pragma unroll
            for (int k = 0; k < 8; ++k) atomicAdd(g_bias + c * 8 + k, acc[k]);
        }
    }
}

// dq, dk, dv: each [M, C] bf16 dense (what the SDPA backward returns for q/k/v views of a packed
// [M, 3C] projection) -> dqkv [M, 3C].  One 16-byte vector per thread.
__global__ void pack_qkv_kernel(const uint4 *__restrict__ dq, const uint4 *__restrict__ dk, const uint4 *__restrict__ dv,
                                uint4 *__restrict__ out, size_t M, int C8) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t per_row = (size_t)3 * C8;
    if (i >= M * per_row) return;
    size_t row = i / per_row;
    int c = (int)(i - row * per_row);
    int which = c / C8, cc = c - which * C8;
    const uint4 *src = which == 0 ? dq : (which == 1 ? dk : dv);
    out[i] = src[row * C8 + cc];
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
    int grid = (M + WARPS - 1) / WARPS;
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
    const size_t smem = sizeof(float) * (size_t)WARPS * NACC_S * D;
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

int xq_vit_pack_qkv(const void *dq, const void *dk, const void *dv, void *dqkv, size_t M, int C, void *stream) {
    if (!dq || !dk || !dv || !dqkv || M == 0 || C <= 0 || (C & 7)) return XQ_ERR_ARG;
    size_t n = M * 3 * (size_t)(C / 8);
    pack_qkv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint4 *)dq, (const uint4 *)dk, (const uint4 *)dv, (uint4 *)dqkv, M, C / 8);
    return cudaGetLastError() == cudaSuccess ? XQ_OK : XQ_ERR_CUDA;
}

int xq_vit_gelu_fwd(const void *x, const float *bias, void *y, int M, int C, void *stream) {
    if (!x || !y || M <= 0 || C <= 0 || (C & 7)) return XQ_ERR_ARG;
    int C8 = C / 8;
    int threads = C8 >= 384 ? 384 : (C8 >= 192 ? 192 : 128);
    int grid = (M + GELU_RU - 1) / GELU_RU < 148 * 4 ? (M + GELU_RU - 1) / GELU_RU : 148 * 4;
    gelu_fwd_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>((const uint4 *)x, bias, (uint4 *)y, M, C8);
    return cudaGetLastError() == cudaSuccess ? XQ_OK : XQ_ERR_CUDA;
}

int xq_vit_gelu_bwd(const void *x, const float *bias, const void *gy, void *gx, float *g_bias, int M, int C, void *stream) {
    if (!x || !gy || !gx || M <= 0 || C <= 0 || (C & 7)) return XQ_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int C8 = C / 8;
    int threads = C8 >= 384 ? 384 : (C8 >= 192 ? 192 : 128);
    int grid = (M + GELU_RU - 1) / GELU_RU < 148 * 4 ? (M + GELU_RU - 1) / GELU_RU : 148 * 4;
    if (g_bias && cudaMemsetAsync(g_bias, 0, sizeof(float) * (size_t)C, st) != cudaSuccess) return XQ_ERR_CUDA;
    gelu_bwd_kernel<<<grid, threads, 0, st>>>((const uint4 *)x, bias, (const uint4 *)gy, (uint4 *)gx, g_bias, M, C8);
    return cudaGetLastError() == cudaSuccess ? XQ_OK : XQ_ERR_CUDA;
}

}  // extern "C"
