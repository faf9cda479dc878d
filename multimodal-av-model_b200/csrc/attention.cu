// attention.cu — the cross attention of CrossAttentionFusion as ONE tcgen05 kernel per direction (sm_100a).
//
// Replaces, inside nn.MultiheadAttention(query=audio, key=value=visual) (/root/reference/model/fusion_module.py:61 ->
// torch/nn/functional.py:6630-6652), the chain  bmm(q/sqrt(d), k^T) -> softmax over ALL T keys (no padding mask) ->
// bmm(P, v)  and its autograd.  Round 1 ran it as three (forward) / six (backward) launches with the [B*H,T,T] fp32
// scores and bf16 probabilities round-tripping HBM; here scores live in tensor memory only:
//
//   forward   CTA = (128-query tile, batch, head):  S = Q.K^T  (tcgen05.mma, fp32 in TMEM) -> each of 128 threads owns
//             one TMEM lane = one query row: max / exp2 / sum straight from TMEM (tcgen05.ld), un-normalised P as bf16
//             into swizzled shared memory -> O = P.V (second tcgen05.mma, V read in place as the MN-major operand) ->
//             O / rowsum -> HBM.  Also writes lse2 = log2(sum exp) per row for the backward.
//   backward  recomputes S (flash-attention style; D = rowsum(dO * O) from the saved output), two kinds of CTA in one
//             launch:  "dQ" CTA (query tile):  S = Q_i.K^T, dP = dO_i.V^T -> dS = P*(dP - D)*alpha -> dQ_i = dS.K
//                      "dKV" CTA (key tile):   S^T = K_j.Q^T, dP^T = V_j.dO^T -> dS^T, P^T -> dK_j = dS^T.Q, dV_j = P^T.dO
//             Both are the same program with the roles of the operands swapped (X = the CTA's 128-row tile, Y = all T rows
//             of the other side), so every operand tile is used both K-major (first MMAs) and MN-major (second MMAs)
//             from a single copy in shared memory.
//
// Shared-memory operand format: "panels" of rows x 64 bf16 (128 B per row, SWIZZLE_128B, filled by TMA boxes of 32 rows);
// a [rows x 128] head slice = 2 panels.  T <= 192 and head_dim == 128 (the reference: T = lip frames ~ 75-150, 512/4);
// other shapes take the unfused route of fusion_path.cu.
#include <cuda.h>
#include <math.h>

#include "common.cuh"
#include "gemm_internal.h"
#include "tcgen05.cuh"

namespace avctc {

constexpr int kAttThreads = 160;          // warp 0: TMA + MMA issue + TMEM allocation; warps 1-4: one TMEM lane (row) per thread
constexpr int kAttMaxT = 192;
constexpr uint32_t kXPanel = 128 * 128;   // bytes of one X panel (128 rows x 128 B)

struct AttnParams {
    int B, T, H, E;
    int tiles, ny16, ny32, panels;
    int forward;
    float alpha, scale_log2;
    __nv_bfloat16* o;            // [B,T,E]   forward: output; backward: input (for D)
    float* lse2;                 // [B*H,T]   forward: output; backward: input
    const __nv_bfloat16* dout;   // [B,T,E]   backward
    __nv_bfloat16* dq;           // [B,T,E]
    __nv_bfloat16* dkv;          // [B,T,2E]
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// byte address of 8 consecutive bf16 (columns c0..c0+7, c0 % 8 == 0) of row `row` in a [128 x (64*panels)] G tile
__device__ __forceinline__ uint32_t g_addr(uint32_t base, int row, int c0) {
    return base + (uint32_t)(c0 >> 6) * kXPanel + (uint32_t)row * 128u + ((uint32_t)(((c0 & 63) >> 3) ^ (row & 7)) << 4);
}
// sum_c a[c] * b[c] over one head slice (128 bf16 = 16 x 16 bytes) of two rows in global memory
__device__ __forceinline__ float head_dot(const __nv_bfloat16* a, const __nv_bfloat16* b) {
    const uint4* pa = reinterpret_cast<const uint4*>(a);
    const uint4* pb = reinterpret_cast<const uint4*>(b);
    float acc = 0.f;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const uint4 x = pa[i], y = pb[i];
        const __nv_bfloat162* hx = reinterpret_cast<const __nv_bfloat162*>(&x);
        const __nv_bfloat162* hy = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 fx = __bfloat1622float2(hx[k]), fy = __bfloat1622float2(hy[k]);
            acc = fmaf(fx.x, fy.x, acc);
            acc = fmaf(fx.y, fy.y, acc);
        }
    }
    return acc;
}
// 128 fp32 accumulator columns of this thread's TMEM lane -> bf16 row in global memory
__device__ __forceinline__ void store_row(uint32_t taddr, __nv_bfloat16* dst, float scale, bool valid) {
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
        uint32_t v[32];
        tmem_ld32(taddr + ch * 32, v);
        tmem_ld_wait();
        if (valid) {
            uint4* d4 = reinterpret_cast<uint4*>(dst + ch * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 pk;
                pk.x = pack_bf16(__uint_as_float(v[8 * j + 0]) * scale, __uint_as_float(v[8 * j + 1]) * scale);
                pk.y = pack_bf16(__uint_as_float(v[8 * j + 2]) * scale, __uint_as_float(v[8 * j + 3]) * scale);
                pk.z = pack_bf16(__uint_as_float(v[8 * j + 4]) * scale, __uint_as_float(v[8 * j + 5]) * scale);
                pk.w = pack_bf16(__uint_as_float(v[8 * j + 6]) * scale, __uint_as_float(v[8 * j + 7]) * scale);
                d4[j] = pk;
            }
        }
    }
}

__global__ void __launch_bounds__(kAttThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                 const __grid_constant__ CUtensorMap map_do, const AttnParams p) {
    extern __shared__ uint8_t att_smem_raw[];
    const uint32_t raw = smem_u32(att_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gen = att_smem_raw + (base - raw);
    const uint32_t ypanel = (uint32_t)p.ny32 * 128u;            // bytes of one Y panel
    const uint32_t sX1 = base, sX2 = base + 2 * kXPanel;
    const uint32_t sY1 = base + 4 * kXPanel, sY2 = sY1 + 2 * ypanel;
    const uint32_t sG = base;                                    // aliases X1|X2 (free once the first MMAs retired)
    const uint32_t sG2 = sY2 + 2 * ypanel;                       // backward only
    const uint32_t tail = (sG2 - base) + (p.forward ? 0u : (uint32_t)p.panels * kXPanel);
    const uint32_t bars = base + tail;                           // ld1, ld2, mma1, mma2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + tail + 32);
    float* sL = reinterpret_cast<float*>(gen + tail + 64);       // [256] lse2 per query column (dKV CTAs)
    float* sD = sL + 256;                                        // [256] D per query column
    const uint32_t bar_ld1 = bars, bar_ld2 = bars + 8, bar_mma1 = bars + 16, bar_mma2 = bars + 24;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = p.forward ? blockIdx.x : (int)(blockIdx.x % p.tiles);
    const int mode = p.forward ? 0 : ((int)blockIdx.x < p.tiles ? 1 : 2);
    const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H;
    const int T = p.T, E = p.E;

    pdl_launch_dependents();
    if (warp == 0) {
        if (lane == 0) {
            mbar_init(bar_ld1, 1); mbar_init(bar_ld2, 1); mbar_init(bar_mma1, 1); mbar_init(bar_mma2, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        tmem_alloc(smem_u32(tmem_slot), 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            // ---- operand roles: X = this CTA's 128-row tile, Y = all T rows of the other side
            const CUtensorMap *mX1, *mY1, *mX2, *mY2;
            int cX1, cY1, cX2, cY2;
            if (mode == 2) { mX1 = &map_kv; cX1 = h * 128; mY1 = &map_q; cY1 = h * 128;
                             mX2 = &map_kv; cX2 = E + h * 128; mY2 = &map_do; cY2 = h * 128; }
            else           { mX1 = &map_q; cX1 = h * 128; mY1 = &map_kv; cY1 = h * 128;
                             mX2 = &map_do; cX2 = h * 128; mY2 = &map_kv; cY2 = E + h * 128; }
            const int row0 = tile * 128;
            const int xrb = min(4, (T - row0 + 31) / 32), yrb = p.ny32 / 32;     // 32-row TMA boxes that start inside T
            mbar_expect_tx(bar_ld1, (uint32_t)(2 * (xrb + yrb)) * 4096u);
            for (int pn = 0; pn < 2; ++pn) {
                for (int rb = 0; rb < xrb; ++rb)
                    tma_load_3d(sX1 + pn * kXPanel + rb * 4096, mX1, bar_ld1, cX1 + pn * 64, row0 + rb * 32, b);
                for (int rb = 0; rb < yrb; ++rb)
                    tma_load_3d(sY1 + pn * ypanel + rb * 4096, mY1, bar_ld1, cY1 + pn * 64, rb * 32, b);
            }
            mbar_expect_tx(bar_ld2, (uint32_t)(2 * ((mode ? xrb : 0) + yrb)) * 4096u);
            for (int pn = 0; pn < 2; ++pn) {
                if (mode)
                    for (int rb = 0; rb < xrb; ++rb)
                        tma_load_3d(sX2 + pn * kXPanel + rb * 4096, mX2, bar_ld2, cX2 + pn * 64, row0 + rb * 32, b);
                for (int rb = 0; rb < yrb; ++rb)
                    tma_load_3d(sY2 + pn * ypanel + rb * 4096, mY2, bar_ld2, cY2 + pn * 64, rb * 32, b);
            }
            // ---- first MMAs: S = X1.Y1^T  (and dP = X2.Y2^T), K = head_dim = 128 = 8 steps of 16
            const uint32_t idesc1 = umma_idesc(p.ny16, 0, 0);
            mbar_wait(bar_ld1, 0);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 8; ++k)
                umma_bf16(tmem_base, umma_desc(sX1 + (k >> 2) * kXPanel + (k & 3) * 32, 16),
                          umma_desc(sY1 + (k >> 2) * ypanel + (k & 3) * 32, 16), idesc1, k ? 1u : 0u);
            if (mode) {
                mbar_wait(bar_ld2, 0);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_bf16(tmem_base + 256, umma_desc(sX2 + (k >> 2) * kXPanel + (k & 3) * 32, 16),
                              umma_desc(sY2 + (k >> 2) * ypanel + (k & 3) * 32, 16), idesc1, k ? 1u : 0u);
            }
            umma_commit(bar_mma1);
        }
        __syncwarp();
        asm volatile("bar.sync 1, %0;" ::"n"(kAttThreads) : "memory");      // G (and G2) are in shared memory
        if (lane == 0) {
            tc_fence_after();
            if (mode == 0) { mbar_wait(bar_ld2, 0); tc_fence_after(); }
            // ---- second MMAs: out = G.Ym, K = keys/queries (ny16), Ym read MN-major from the same panels
            const uint32_t idesc2 = umma_idesc(128, 0, 1);
            const int nk = p.ny16 / 16;
            const uint32_t sYm = (mode == 0) ? sY2 : sY1;
            for (int k = 0; k < nk; ++k)
                umma_bf16(tmem_base, umma_desc(sG + (k >> 2) * kXPanel + (k & 3) * 32, 16),
                          umma_desc(sYm + k * 2048, ypanel), idesc2, k ? 1u : 0u);
            if (mode == 2)
                for (int k = 0; k < nk; ++k)
                    umma_bf16(tmem_base + 256, umma_desc(sG2 + (k >> 2) * kXPanel + (k & 3) * 32, 16),
                              umma_desc(sY2 + k * 2048, ypanel), idesc2, k ? 1u : 0u);
            umma_commit(bar_mma2);
        }
        __syncwarp();
    } else {
        const int q4 = warp & 3;                       // TMEM lane quarter this warp may touch
        const int r = q4 * 32 + lane;                  // row inside the tile = TMEM lane
        const int gi = tile * 128 + r;                 // query index (modes 0, 1) / key index (mode 2)
        const bool valid = gi < T;
        const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
        const size_t hoff = (size_t)h * 128;
        float rowL = 0.f, rowD = 0.f;
        if (mode == 1) {
            if (valid) {
                rowL = p.lse2[(size_t)bh * T + gi];
                const size_t ro = ((size_t)b * T + gi) * E + hoff;
                rowD = head_dot(p.o + ro, p.dout + ro);
            } else {
                rowL = CUDART_INF_F;
            }
        } else if (mode == 2) {
            for (int i = threadIdx.x - 32; i < p.panels * 64; i += 128) {
                float L = CUDART_INF_F, D = 0.f;
                if (i < T) {
                    L = p.lse2[(size_t)bh * T + i];
                    const size_t ro = ((size_t)b * T + i) * E + hoff;
                    D = head_dot(p.o + ro, p.dout + ro);
                }
                sL[i] = L; sD[i] = D;
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");
        }
        mbar_wait(bar_mma1, 0);
        tc_fence_after();
        const int nch = p.panels * 2;                  // 32-column chunks covering [0, 64*panels) >= ny16
        float inv = 1.f;
        if (mode == 0) {
            float m = AVCTC_NEG_INF;
#pragma unroll 1
            for (int ch = 0; ch < nch; ++ch) {
                uint32_t v[32];
                tmem_ld32(taddr + ch * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (ch * 32 + j < T) m = fmaxf(m, __uint_as_float(v[j]));
            }
            const float ms = m * p.scale_log2;
            float sum = 0.f;
#pragma unroll 1
            for (int ch = 0; ch < nch; ++ch) {
                uint32_t v[32];
                tmem_ld32(taddr + ch * 32, v);
                tmem_ld_wait();
                float e[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    e[j] = (ch * 32 + j < T) ? ex2_approx(fmaf(__uint_as_float(v[j]), p.scale_log2, -ms)) : 0.f;
                    sum += e[j];
                }
#pragma unroll
                for (int j = 0; j < 32; j += 8)
                    st_shared_v4(g_addr(sG, r, ch * 32 + j), pack_bf16(e[j], e[j + 1]), pack_bf16(e[j + 2], e[j + 3]),
                                 pack_bf16(e[j + 4], e[j + 5]), pack_bf16(e[j + 6], e[j + 7]));
            }
            inv = 1.f / sum;
            if (valid) p.lse2[(size_t)bh * T + gi] = ms + lg2_approx(sum);
        } else {
#pragma unroll 1
            for (int ch = 0; ch < nch; ++ch) {
                uint32_t v[32], w[32];
                tmem_ld32(taddr + ch * 32, v);
                tmem_ld32(taddr + 256 + ch * 32, w);
                tmem_ld_wait();
                float ds[32], pr[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int c = ch * 32 + j;
                    const float L = (mode == 1) ? rowL : sL[c];
                    const float D = (mode == 1) ? rowD : sD[c];
                    // columns >= T: exactly 0 (TMEM beyond the MMA's N columns is stale and may hold NaN patterns)
                    pr[j] = (c < T) ? ex2_approx(fmaf(__uint_as_float(v[j]), p.scale_log2, -L)) : 0.f;
                    ds[j] = (c < T) ? pr[j] * (__uint_as_float(w[j]) - D) * p.alpha : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    st_shared_v4(g_addr(sG, r, ch * 32 + j), pack_bf16(ds[j], ds[j + 1]), pack_bf16(ds[j + 2], ds[j + 3]),
                                 pack_bf16(ds[j + 4], ds[j + 5]), pack_bf16(ds[j + 6], ds[j + 7]));
                    if (mode == 2)
                        st_shared_v4(g_addr(sG2, r, ch * 32 + j), pack_bf16(pr[j], pr[j + 1]), pack_bf16(pr[j + 2], pr[j + 3]),
                                     pack_bf16(pr[j + 4], pr[j + 5]), pack_bf16(pr[j + 6], pr[j + 7]));
                }
            }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(kAttThreads) : "memory");
        mbar_wait(bar_mma2, 0);
        tc_fence_after();
        if (mode == 0) {
            store_row(taddr, p.o + ((size_t)b * T + gi) * E + hoff, inv, valid);
        } else if (mode == 1) {
            store_row(taddr, p.dq + ((size_t)b * T + gi) * E + hoff, 1.f, valid);
        } else {
            __nv_bfloat16* d = p.dkv + ((size_t)b * T + gi) * 2 * E + hoff;
            store_row(taddr, d, 1.f, valid);
            store_row(taddr + 256, d + E, 1.f, valid);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace avctc

using namespace avctc;

// 1 when the fused kernels cover (T, head_dim); the caller keeps the unfused GEMM + softmax route otherwise.
int avctc_attention_supported(int T, int E, int H) { return T >= 1 && T <= kAttMaxT && H > 0 && E == H * 128; }

// forward: q [B,T,E], kv [B,T,2E] (keys | values) -> o [B,T,E], lse2 [B*H,T]
// backward (forward == 0): also dout [B,T,E] and the saved o / lse2 -> dq [B,T,E], dkv [B,T,2E]
int avctc_attention_launch(int forward, const void* q, const void* kv, const void* dout, void* o, float* lse2, void* dq,
                           void* dkv, int B, int T, int H, int E, void* stream) {
    if (!avctc_attention_supported(T, E, H) || B <= 0) return AVCTC_ERR_UNSUPPORTED;
    if (!q || !kv || !o || !lse2 || (!forward && (!dout || !dq || !dkv))) return AVCTC_ERR_BAD_ARG;
    AttnParams p;
    p.B = B; p.T = T; p.H = H; p.E = E;
    p.tiles = (T + 127) / 128; p.ny16 = (T + 15) / 16 * 16; p.ny32 = (T + 31) / 32 * 32; p.panels = (T + 63) / 64;
    p.forward = forward ? 1 : 0;
    p.alpha = 1.f / sqrtf(128.f);
    p.scale_log2 = p.alpha * AVCTC_LOG2E;
    p.o = reinterpret_cast<__nv_bfloat16*>(o); p.lse2 = lse2;
    p.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
    p.dq = reinterpret_cast<__nv_bfloat16*>(dq); p.dkv = reinterpret_cast<__nv_bfloat16*>(dkv);
    CUtensorMap mq, mkv, mdo;
    int rc = avctc_tensor_map(&mq, q, E, T, B, E, (long long)T * E, 64, 32);
    if (rc) return rc;
    rc = avctc_tensor_map(&mkv, kv, 2 * E, T, B, 2 * E, (long long)T * 2 * E, 64, 32);
    if (rc) return rc;
    if (forward) mdo = mq;
    else if ((rc = avctc_tensor_map(&mdo, dout, E, T, B, E, (long long)T * E, 64, 32))) return rc;
    const size_t smem = 4 * (size_t)kXPanel + 4 * (size_t)p.ny32 * 128 + (forward ? 0 : (size_t)p.panels * kXPanel) + 64 +
                        2 * 256 * sizeof(float) + 1024;
    static size_t configured = 0;
    if (smem > configured) {
        AVCTC_CUDA_RETURN(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const dim3 grid(forward ? p.tiles : 2 * p.tiles, B * H);
    return (int)avctc_launch_pdl(attention_kernel, grid, dim3(kAttThreads), smem, reinterpret_cast<cudaStream_t>(stream),
                                 mq, mkv, mdo, p);
}

// See include/avctc_b200.h.
extern "C" int avctc_attention_forward(const void* q, const void* kv, void* o, float* lse2, int B, int T, int H, int E,
                                       void* stream) {
    return avctc_attention_launch(1, q, kv, nullptr, o, lse2, nullptr, nullptr, B, T, H, E, stream);
}
extern "C" int avctc_attention_backward(const void* q, const void* kv, const void* dout, const void* o, const float* lse2,
                                        void* dq, void* dkv, int B, int T, int H, int E, void* stream) {
    return avctc_attention_launch(0, q, kv, dout, const_cast<void*>(o), const_cast<float*>(lse2), dq, dkv, B, T, H, E, stream);
}
