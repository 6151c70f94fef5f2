// xq_gelu.cuh -- the exact-erf GELU and its derivative as ONE set of device functions, shared by the stand-alone bias + GELU
// kernels (vit_kernels.cu) and the fused GEMM epilogues (gemm_kernel.cu), so that both paths produce the same bits.
// Reference op: nn.GELU() (erf form) inside timm's Mlp, tokenizer/tokenizer_image/dino_enc/vision_transformer.py:336-339.
#pragma once
#include <cuda_runtime.h>

namespace xqv {

__device__ __forceinline__ float rcp_fast(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// gelu(x) = x/2 * (1 + erf(x/sqrt2)) with erf(|z|) = 1 - r, r = (1 + a1|z| + ... + a6|z|^6)^-16 (A&S 7.1.28):
//   gelu(x) = h + |h| * (1 - r) = (h + |h|) - |h| * r ,  h = x/2.
// The 1/sqrt2 of z is folded into the coefficients (b_k = a_k * 2^(-k/2)) and the sign handling into the abs/neg operand
// modifiers, which takes 3 instructions off the previous form (the kernel is instruction-issue bound, not HBM bound).
__device__ __forceinline__ float gelu_f(float x) {
    const float a = fabsf(x);
    float p = fmaf(a, 5.382975000e-06f, 4.889063564e-05f);
    p = fmaf(p, a, 3.800357500e-05f);
    p = fmaf(p, a, 3.277626324e-03f);
    p = fmaf(p, a, 2.114100615e-02f);
    p = fmaf(p, a, 4.986734697e-02f);
    p = fmaf(p, a, 1.0f);
    p = p * p; p = p * p; p = p * p; p = p * p;
    const float h = 0.5f * x;
    return fmaf(-fabsf(h), rcp_fast(p), h + fabsf(h));
}
// gelu'(x) = Phi(x) + x phi(x).  Both need exp(-x^2/2): with A&S 7.1.26,  erf(z) = 1 - (a1 t + ... + a5 t^5) exp(-z^2),
// t = 1/(1 + p z), the SAME exponential serves Phi and phi (z = |x|/sqrt2), so the derivative costs one ex2 + one rcp +
// ~14 FP32 instructions (the 7.1.28 erf + a separate exp used before: ~20).  |abs err| 3e-7.
__device__ __forceinline__ float dgelu_f(float x) {
    const float ax = fabsf(x);
    const float t = rcp_fast(fmaf(ax, 0.23164189f, 1.0f));             // p / sqrt2 = 0.3275911 / 1.41421356
    float q = fmaf(t, 1.061405429f, -1.453152027f);
    q = fmaf(q, t, 1.421413741f);
    q = fmaf(q, t, -0.284496736f);
    q = fmaf(q, t, 0.254829592f);
    const float e = ex2_fast(x * x * -0.72134752044448170f);            // exp(-x^2 / 2)
    const float pe = q * t * e;                                          // 1 - erf(|x| / sqrt2)
    const float half = fmaf(-0.5f, pe, 0.5f);                           // Phi(|x|) - 1/2
    return fmaf(x * 0.3989422804014327f, e, 0.5f + copysignf(half, x));
}


// ---- two elements per instruction: Blackwell executes packed fp32 pairs (FFMA2 / FMUL2 / FADD2).  The fused GEMM epilogues are
// bound by instruction issue, so they use these; every lane performs exactly the operation sequence of the scalar functions above
// (same fma / mul / add, same order, IEEE rounding per lane), i.e. the same bits.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void up2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 dup2(float c) { return pk2(c, c); }

// (gelu_f(x0), gelu_f(x1))
__device__ __forceinline__ f32x2 gelu_f2(float x0, float x1) {
    const f32x2 A = pk2(fabsf(x0), fabsf(x1));
    f32x2 P = fma2(A, dup2(5.382975000e-06f), dup2(4.889063564e-05f));
    P = fma2(P, A, dup2(3.800357500e-05f));
    P = fma2(P, A, dup2(3.277626324e-03f));
    P = fma2(P, A, dup2(2.114100615e-02f));
    P = fma2(P, A, dup2(4.986734697e-02f));
    P = fma2(P, A, dup2(1.0f));
    P = mul2(P, P); P = mul2(P, P); P = mul2(P, P); P = mul2(P, P);
    float p0, p1, h0, h1;
    up2(P, p0, p1);
    const f32x2 H = mul2(dup2(0.5f), pk2(x0, x1));
    up2(H, h0, h1);
    const f32x2 AH = pk2(fabsf(h0), fabsf(h1));
    // -|h| * r + (h + |h|)  ==  |h| * rcp(-p) + (h + |h|): the reciprocal is odd, the product's sign symmetric
    return fma2(AH, pk2(rcp_fast(-p0), rcp_fast(-p1)), add2(H, AH));
}
// (dgelu_f(x0), dgelu_f(x1))
__device__ __forceinline__ f32x2 dgelu_f2(float x0, float x1) {
    const f32x2 X = pk2(x0, x1);
    float u0, u1;
    up2(fma2(pk2(fabsf(x0), fabsf(x1)), dup2(0.23164189f), dup2(1.0f)), u0, u1);
    const f32x2 T = pk2(rcp_fast(u0), rcp_fast(u1));
    f32x2 Q = fma2(T, dup2(1.061405429f), dup2(-1.453152027f));
    Q = fma2(Q, T, dup2(1.421413741f));
    Q = fma2(Q, T, dup2(-0.284496736f));
    Q = fma2(Q, T, dup2(0.254829592f));
    float s0, s1;
    up2(mul2(mul2(X, X), dup2(-0.72134752044448170f)), s0, s1);
    const f32x2 E = pk2(ex2_fast(s0), ex2_fast(s1));
    const f32x2 PE = mul2(mul2(Q, T), E);
    float hf0, hf1;
    up2(fma2(dup2(-0.5f), PE, dup2(0.5f)), hf0, hf1);
    const f32x2 S = add2(dup2(0.5f), pk2(copysignf(hf0, x0), copysignf(hf1, x1)));
    return fma2(mul2(X, dup2(0.3989422804014327f)), E, S);
}

}  // namespace xqv
