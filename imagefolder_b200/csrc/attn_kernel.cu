// attn_kernel.cu -- tcgen05 / TMEM / TMA flash attention for the ViT blocks (head_dim 64, bf16, no mask), sm_100a.
//
// Replaces F.scaled_dot_product_attention in Attention.forward
//   (tokenizer/tokenizer_image/dino_enc/vision_transformer.py:173-197: q,k,v = qkv.reshape(B,N,3,H,hd).permute(2,0,3,1,4);
//    x = sdpa(q,k,v); x = x.transpose(1,2).reshape(B,N,C))
// reading q/k/v straight out of the packed projection [B,N,3,H,64] through ONE 3-D tensor map and writing the
// head-merged output [B,N,H*64], so that no permute / contiguous copy exists on either side.
//
// Forward, one CTA per (batch, head, 128-query tile), 2 CTAs per SM (256 TMEM columns each):
//   warp 0   TMA producer : Q tile once, K / V row tiles (128 keys x 64) through 2-stage rings
//   warp 1   MMA issuer   : S = Q K^T   (kind::f16, A,B K-major smem, N = 128 or the 16-rounded tail)  -> TMEM[0,128)
//                           O += P V    (A = P from TMEM[128,192), B = V MN-major smem, N = 64)        -> TMEM[192,256)
//   warps 4-7 softmax     : one query row per thread (TMEM lane): S -> registers, running max with LAZY rescaling of O
//                           (only when the row max grows by more than 2^8), P = exp2(S*c - m*c) -> bf16 -> TMEM,
//                           epilogue O / l -> bf16 -> swizzled smem -> TMA store (rows beyond N are clipped by the map)
//   sequence lengths need not be multiples of anything: the last key block is issued with N = ceil16(valid keys) and
//   its invalid columns are masked to -inf; rows of the last query tile beyond N are zero-filled by TMA and clipped on store.
// The statistics tensor holds L2[b,h,n] = m*c + log2(l) (base-2 log-sum-exp of the SCALED scores), what backward needs.
#include "xq_common.cuh"
#include "xq_tc.cuh"

#include <mutex>

namespace xq {
using namespace xqtc;

constexpr int AT_BM = 128;          // queries per CTA
constexpr int AT_BN = 128;          // keys per block
constexpr int AT_D = 64;            // head dim
constexpr int AT_NS = 2;            // K / V ring stages
constexpr int AT_TILE = AT_BM * AT_D * 2;   // bytes of one [128][64] bf16 row tile
constexpr int AT_THREADS = 256;
constexpr float AT_LAZY = 8.0f;     // rescale O only when the scaled max grows by more than this (log2 units)

struct AttnFwdSmem {
    // offsets from the 1024-aligned base
    static constexpr int Q = 0;
    static constexpr int K = AT_TILE;
    static constexpr int V = AT_TILE * (1 + AT_NS);
    static constexpr int BAR = AT_TILE * (1 + 2 * AT_NS);
    static constexpr int BYTES = BAR + 256;
};

template <int REGS>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO, float *__restrict__ lse2,
                int N, int H, int nQ, float c /* softmax scale * log2(e) */) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = (uint64_t *)(base + AttnFwdSmem::BAR);
    uint64_t *q_full = bars + 0;
    uint64_t *k_full = bars + 1;              // [AT_NS]
    uint64_t *k_empty = bars + 1 + AT_NS;     // [AT_NS]
    uint64_t *v_full = bars + 1 + 2 * AT_NS;
    uint64_t *v_empty = bars + 1 + 3 * AT_NS;
    uint64_t *s_full = bars + 1 + 4 * AT_NS;
    uint64_t *s_free = s_full + 1;
    uint64_t *p_full = s_full + 2;
    uint64_t *pv_done = s_full + 3;
    uint32_t *tmem_holder = (uint32_t *)(s_full + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bh = blockIdx.x / nQ, qt = blockIdx.x - bh * nQ;
    const int b = bh / H, h = bh - b * H;
    const int q0 = qt * AT_BM;
    const int nK = (N + AT_BN - 1) / AT_BN;
    const int colQ = h * AT_D, colK = (H + h) * AT_D, colV = (2 * H + h) * AT_D;

    if (tid == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < AT_NS; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
        mbar_init(s_full, 1);
        mbar_init(s_free, 4);
        mbar_init(p_full, 4);
        mbar_init(pv_done, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<256>(tmem_holder);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;
    const uint32_t tS = tmem, tP = tmem + 128, tO = tmem + 192;

    if (warp < 4) {
        reg_dec<40>();
        if (warp == 0 && lane == 0) {
            // ===== TMA producer =====
            tma_prefetch_desc(&tmQKV);
            mbar_expect_tx(q_full, AT_TILE);
            tma_load_3d(base + AttnFwdSmem::Q, &tmQKV, colQ, q0, b, q_full);
            for (int j = 0; j < nK; ++j) {
                const int st = j % AT_NS;
                const uint32_t ph = ((j / AT_NS) & 1) ^ 1;
                mbar_wait(&k_empty[st], ph);
                mbar_expect_tx(&k_full[st], AT_TILE);
                tma_load_3d(base + AttnFwdSmem::K + st * AT_TILE, &tmQKV, colK, j * AT_BN, b, &k_full[st]);
                mbar_wait(&v_empty[st], ph);
                mbar_expect_tx(&v_full[st], AT_TILE);
                tma_load_3d(base + AttnFwdSmem::V + st * AT_TILE, &tmQKV, colV, j * AT_BN, b, &v_full[st]);
            }
        } else if (warp == 1 && lane == 0) {
            // ===== MMA issuer =====
            const uint32_t qa = smem_u32(base + AttnFwdSmem::Q);
            auto issue_qk = [&](int j) {
                const int st = j % AT_NS;
                const int nv = min(AT_BN, N - j * AT_BN);
                const int nj = (nv + 15) & ~15;
                mbar_wait(&k_full[st], (j / AT_NS) & 1);
                tc_fence_after();
                const uint32_t ka = smem_u32(base + AttnFwdSmem::K + st * AT_TILE);
                const uint32_t id = idesc_bf16(AT_BM, nj, 0, 0);
#pragma unroll
                for (int k = 0; k < AT_D / 16; ++k) umma_ss(tS, desc_k_sw128(qa + k * 32), desc_k_sw128(ka + k * 32), id, k > 0);
                umma_commit(&k_empty[st]);
                umma_commit(s_full);
            };
            mbar_wait(q_full, 0);
            issue_qk(0);
            for (int j = 0; j < nK; ++j) {
                if (j + 1 < nK) {
                    mbar_wait(s_free, j & 1);          // softmax holds S_j in registers
                    tc_fence_after();
                    issue_qk(j + 1);
                }
                const int st = j % AT_NS;
                const int nv = min(AT_BN, N - j * AT_BN);
                const int nj = (nv + 15) & ~15;
                mbar_wait(&v_full[st], (j / AT_NS) & 1);
                mbar_wait(p_full, j & 1);              // P_j written (and O rescaled)
                tc_fence_after();
                const uint32_t va = smem_u32(base + AttnFwdSmem::V + st * AT_TILE);
                const uint32_t id = idesc_bf16(AT_BM, AT_D, 0, 1);
                for (int k = 0; k < nj / 16; ++k)
                    umma_ts(tO, tP + k * 8, desc_mn_sw128(va + k * 2048, 16384, 1024), id, (j | k) != 0);
                umma_commit(&v_empty[st]);
                umma_commit(pv_done);
            }
        }
    } else {
        reg_inc<216>();
        // ===== softmax: thread = query row =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        float m_used = -CUDART_INF_F;     // raw-score max the current P / O are expressed against
        float l = 0.f;
        for (int j = 0; j < nK; ++j) {
            const int nv = min(AT_BN, N - j * AT_BN);       // valid keys in this block
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            uint32_t s[128];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                if (ch * 32 < nv) tmem_ld32(tS + lane_addr + ch * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[ch * 32]));
            }
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free);
            float mx = -CUDART_INF_F;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                if (ch * 32 < nv) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float x = __uint_as_float(s[ch * 32 + i]);
                        if (ch * 32 + i >= nv) x = -CUDART_INF_F;
                        s[ch * 32 + i] = __float_as_uint(x);
                        mx = fmaxf(mx, x);
                    }
                }
            }
            if (j > 0) mbar_wait(pv_done, (j - 1) & 1);       // O and the P buffer are quiescent
            tc_fence_after();
            const bool grow = (mx - m_used) * c > AT_LAZY;     // j == 0: m_used = -inf -> true
            if (j == 0) {
                m_used = mx;
            } else if (__any_sync(0xffffffffu, grow)) {
                const float m_new = grow ? mx : m_used;
                const float alpha = ex2_approx((m_used - m_new) * c);
                l *= alpha;
                m_used = m_new;
#pragma unroll
                for (int c0 = 0; c0 < AT_D; c0 += 16) {
                    uint32_t o[16];
                    tmem_ld16(tO + lane_addr + c0, o);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st16(tO + lane_addr + c0, o);
                }
            }
            const float mc = m_used * c;
            float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                if (ch * 32 < nv) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float p0 = ex2_approx(fmaf(__uint_as_float(s[ch * 32 + i]), c, -mc));
                        float p1 = ex2_approx(fmaf(__uint_as_float(s[ch * 32 + i + 1]), c, -mc));
                        sum0 += p0;
                        sum1 += p1;
                        pk[i >> 1] = pack_bf16(p0, p1);
                    }
                    tmem_st16(tP + lane_addr + ch * 16, pk);
                }
            }
            l += sum0 + sum1;
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // ---- epilogue: O / l -> bf16 -> smem (the Q tile is dead: every S MMA has completed) -> TMA store
        mbar_wait(pv_done, (nK - 1) & 1);
        tc_fence_after();
        const float inv = 1.0f / l;
        uint8_t *so = base + AttnFwdSmem::Q;
#pragma unroll
        for (int c0 = 0; c0 < AT_D; c0 += 16) {
            uint32_t o[16];
            tmem_ld16(tO + lane_addr + c0, o);
            tmem_wait_ld();
            uint4 v0, v1;
            v0.x = pack_bf16(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
            v0.y = pack_bf16(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
            v0.z = pack_bf16(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
            v0.w = pack_bf16(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
            v1.x = pack_bf16(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
            v1.y = pack_bf16(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
            v1.z = pack_bf16(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
            v1.w = pack_bf16(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
            *reinterpret_cast<uint4 *>(so + rowtile_unit(row, c0 / 8)) = v0;
            *reinterpret_cast<uint4 *>(so + rowtile_unit(row, c0 / 8 + 1)) = v1;
        }
        if (q0 + row < N) lse2[(size_t)bh * N + q0 + row] = fmaf(m_used, c, log2f(l));
        fence_async_smem();
        tc_fence_before();
        named_bar_sync(1, 128);
        if (warp == 4 && lane == 0) {
            tma_store_3d(&tmO, so, h * AT_D, q0, b);
            bulk_commit();
            bulk_wait_read<0>();
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<256>(tmem);
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------
struct AttnMaps {
    const void *qkv, *out;
    int B, N, H;
    CUtensorMap tmQKV, tmO;
};

static bool get_fwd_maps(const void *qkv, void *out, int B, int N, int H, AttnMaps &m) {
    static std::mutex mu;
    static AttnMaps cache[16];
    static int n_cached = 0, next = 0;
    std::lock_guard<std::mutex> g(mu);
    for (int i = 0; i < n_cached; ++i)
        if (cache[i].qkv == qkv && cache[i].out == out && cache[i].B == B && cache[i].N == N && cache[i].H == H) { m = cache[i]; return true; }
    AttnMaps e;
    e.qkv = qkv; e.out = out; e.B = B; e.N = N; e.H = H;
    const uint64_t W = (uint64_t)3 * H * AT_D;
    if (!make_map_3d(&e.tmQKV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(qkv), W, N, B, W * 2, (uint64_t)N * W * 2, AT_D, AT_BM)) return false;
    const uint64_t Wo = (uint64_t)H * AT_D;
    if (!make_map_3d(&e.tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, Wo, N, B, Wo * 2, (uint64_t)N * Wo * 2, AT_D, AT_BM)) return false;
    cache[next] = e;
    next = (next + 1) % 16;
    if (n_cached < 16) ++n_cached;
    m = e;
    return true;
}

}  // namespace xq

extern "C" {

int xq_vit_attn_fwd(const void *qkv, void *out, float *lse2, int B, int N, int H, int head_dim, float scale, void *stream) {
    using namespace xq;
    if (!qkv || !out || !lse2 || B <= 0 || N <= 0 || H <= 0) return XQ_ERR_ARG;
    if (head_dim != AT_D) return XQ_ERR_UNSUPPORTED;
    if (((uintptr_t)qkv & 15) || ((uintptr_t)out & 15)) return XQ_ERR_ARG;
    AttnMaps m;
    if (!get_fwd_maps(qkv, out, B, N, H, m)) return XQ_ERR_UNSUPPORTED;
    const int nQ = (N + AT_BM - 1) / AT_BM;
    const size_t smem = AttnFwdSmem::BYTES + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        XQ_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const long long ctas = (long long)B * H * nQ;
    if (ctas > 0x7fffffffLL) return XQ_ERR_ARG;
    attn_fwd_kernel<<<(unsigned)ctas, AT_THREADS, smem, (cudaStream_t)stream>>>(m.tmQKV, m.tmO, lse2, N, H, nQ, scale * 1.4426950408889634f);
    XQ_LAUNCH_CHECK("attn_fwd_kernel");
    return XQ_OK;
}

}  // extern "C"
