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
template <int NV>
__global__ void __launch_bounds__(THREADS)
residual_ln_fwd_kernel(const float *__restrict__ x, const __nv_bfloat16 *__restrict__ branch,
                       const float *__restrict__ ls_gamma, const float *__restrict__ rowscale, int rows_per_sample,
                       const float *__restrict__ ln_w, const float *__restrict__ ln_b, float eps, int M,
                       float *__restrict__ x_out, __nv_bfloat16 *__restrict__ y, float *__restrict__ mean_out,
                       float *__restrict__ rstd_out) {
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

// Backward.  Persistent grid; each warp walks rows with stride, keeping its column partial sums in
// registers; CTA partials -> part[blockIdx][3][D]; a second kernel reduces over blocks.
//   g_xout may be null (no later residual gradient), g_y may be null (the LN output was unused).
template <int NV>
__global__ void __launch_bounds__(THREADS)
residual_ln_bwd_kernel(const float *__restrict__ g_xout, const __nv_bfloat16 *__restrict__ g_y,
                       const float *__restrict__ x_out, const float *__restrict__ mean_in,
                       const float *__restrict__ rstd_in, const float *__restrict__ ln_w,
                       const __nv_bfloat16 *__restrict__ branch, const float *__restrict__ ls_gamma,
                       const float *__restrict__ rowscale, int rows_per_sample, int M, float *__restrict__ g_x,
                       __nv_bfloat16 *__restrict__ g_branch, float *__restrict__ part) {
    constexpr int D = NV * 128;
    __shared__ float red[WARPS][128];  // staging for the cross-warp column reduction (one chunk at a time)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 aw[NV], ab[NV], ag[NV];  // d ln_w, d ln_b, d ls_gamma partials
#pragma unroll
    for (int i = 0; i < NV; ++i) { aw[i] = ab[i] = ag[i] = make_float4(0.f, 0.f, 0.f, 0.f); }
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
                ab[i].x += g.x; ab[i].y += g.y; ab[i].z += g.z; ab[i].w += g.w;
                aw[i].x += g.x * xh[i].x; aw[i].y += g.y * xh[i].y; aw[i].z += g.z * xh[i].z; aw[i].w += g.w * xh[i].w;
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
                ag[i].x += G.x * s * b.x; ag[i].y += G.y * s * b.y; ag[i].z += G.z * s * b.z; ag[i].w += G.w * s * b.w;
                if (g_branch)
                    store_bf16x4(g_branch + base + col, make_float4(G.x * s * gm.x, G.y * s * gm.y, G.z * s * gm.z, G.w * s * gm.w));
            }
        }
    }
    // cross-warp reduction of the column partials, one 128-column chunk at a time
    float *outp = part + (size_t)blockIdx.x * 3 * D;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float4 a = q == 0 ? aw[i] : (q == 1 ? ab[i] : ag[i]);
            __syncthreads();
            *reinterpret_cast<float4 *>(&red[warp][lane * 4]) = a;
            __syncthreads();
            if (threadIdx.x < 128) {
                float acc = 0.f;
#pragma unroll
                for (int w = 0; w < WARPS; ++w) acc += red[w][threadIdx.x];
                outp[(size_t)q * D + i * 128 + threadIdx.x] = acc;
            }
        }
    }
}

__global__ void reduce_parts_kernel(const float *__restrict__ part, int nblocks, int n, float *__restrict__ o0,
                                    float *__restrict__ o1, float *__restrict__ o2, int D) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float acc = 0.f;
    for (int b = 0; b < nblocks; ++b) acc += part[(size_t)b * n + e];
    int q = e / D, d = e - q * D;
    float *o = q == 0 ? o0 : (q == 1 ? o1 : o2);
    if (o) o[d] = acc;
}

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_f(float x) {
    return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

__global__ void gelu_fwd_kernel(const uint4 *__restrict__ x, uint4 *__restrict__ y, size_t n8) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    uint4 v = x[i];
    __nv_bfloat162 *p = reinterpret_cast<__nv_bfloat162 *>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float2 f = __bfloat1622float2(p[k]);
        p[k] = __floats2bfloat162_rn(gelu_f(f.x), gelu_f(f.y));
    }
    y[i] = v;
}

__global__ void gelu_bwd_kernel(const uint4 *__restrict__ x, const uint4 *__restrict__ gy, uint4 *__restrict__ gx,
                                size_t n8) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    uint4 v = x[i], g = gy[i];
    __nv_bfloat162 *p = reinterpret_cast<__nv_bfloat162 *>(&v);
    __nv_bfloat162 *q = reinterpret_cast<__nv_bfloat162 *>(&g);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float2 f = __bfloat1622float2(p[k]), h = __bfloat1622float2(q[k]);
        p[k] = __floats2bfloat162_rn(h.x * dgelu_f(f.x), h.y * dgelu_f(f.y));
    }
    gx[i] = v;
}

static int bwd_grid() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms * 2;
}

}  // namespace xqv

using namespace xqv;

#define XQV_DISPATCH(D, CALL)                 \
    switch (D) {                               \
        case 384: { constexpr int NV = 3; CALL; break; }   \
        case 768: { constexpr int NV = 6; CALL; break; }   \
        case 1024: { constexpr int NV = 8; CALL; break; }  \
        default: return XQ_ERR_UNSUPPORTED;    \
    }

extern "C" {

size_t xq_vit_ln_bwd_workspace_bytes(int D) { return sizeof(float) * (size_t)bwd_grid() * 3 * D; }

int xq_vit_residual_ln_fwd(const float *x, const void *branch, const float *ls_gamma, const float *rowscale,
                           int rows_per_sample, const float *ln_w, const float *ln_b, float eps, int M, int D,
                           float *x_out, void *y, float *mean, float *rstd, void *stream) {
    if (!x || M <= 0 || (y && (!ln_w || !ln_b)) || (!x_out && !y)) return XQ_ERR_ARG;
    if (branch && rowscale && rows_per_sample <= 0) return XQ_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int grid = (M + WARPS - 1) / WARPS;
    XQV_DISPATCH(D, (residual_ln_fwd_kernel<NV><<<grid, THREADS, 0, st>>>(
                        x, (const __nv_bfloat16 *)branch, ls_gamma, rowscale, rows_per_sample, ln_w, ln_b, eps, M,
                        x_out, (__nv_bfloat16 *)y, mean, rstd)));
    return cudaGetLastError() == cudaSuccess ? XQ_OK : XQ_ERR_CUDA;
}

int xq_vit_residual_ln_bwd(const float *g_xout, const void *g_y, const float *x_out, const float *mean,
                           const float *rstd, const float *ln_w, const void *branch, const float *ls_gamma,
                           const float *rowscale, int rows_per_sample, int M, int D, float *g_x, void *g_branch,
                           float *g_ln_w, float *g_ln_b, float *g_ls_gamma, void *workspace, size_t workspace_bytes,
                           void *stream) {
    if (!x_out || !mean || !rstd || M <= 0 || !workspace) return XQ_ERR_ARG;
    if (g_y && !ln_w) return XQ_ERR_ARG;
    const int grid = bwd_grid();
    if (workspace_bytes < sizeof(float) * (size_t)grid * 3 * D) return XQ_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float *part = (float *)workspace;
    XQV_DISPATCH(D, (residual_ln_bwd_kernel<NV><<<grid, THREADS, 0, st>>>(
                        g_xout, (const __nv_bfloat16 *)g_y, x_out, mean, rstd, ln_w, (const __nv_bfloat16 *)branch,
                        ls_gamma, rowscale, rows_per_sample, M, g_x, (__nv_bfloat16 *)g_branch, part)));
    if (cudaGetLastError() != cudaSuccess) return XQ_ERR_CUDA;
    int n = 3 * D;
    reduce_parts_kernel<<<(n + 255) / 256, 256, 0, st>>>(part, grid, n, g_ln_w, g_ln_b, g_ls_gamma, D);
    return cudaGetLastError() == cudaSuccess ? XQ_OK : XQ_ERR_CUDA;
}

int xq_vit_gelu_fwd(const void *x, void *y, size_t n, void *stream) {
    if (!x || !y || (n & 7)) return XQ_ERR_ARG;
    size_t n8 = n / 8;
    gelu_fwd_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint4 *)x, (uint4 *)y, n8);
    return cudaGetLastError() == cudaSuccess ? XQ_OK : XQ_ERR_CUDA;
}

int xq_vit_gelu_bwd(const void *x, const void *gy, void *gx, size_t n, void *stream) {
    if (!x || !gy || !gx || (n & 7)) return XQ_ERR_ARG;
    size_t n8 = n / 8;
    gelu_bwd_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint4 *)x, (const uint4 *)gy,
                                                                                  (uint4 *)gx, n8);
    return cudaGetLastError() == cudaSuccess ? XQ_OK : XQ_ERR_CUDA;
}

}  // extern "C"
