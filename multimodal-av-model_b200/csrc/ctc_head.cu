// ctc_head.cu — CTCDecoder forward as ONE kernel: log_softmax(x . W^T + b) with the logits never leaving the SM.
//
// Replaces /root/reference/model/decoder.py:24-25 (nn.Linear(1024 -> V) then F.log_softmax) — and, with passes = 2, the
// second F.log_softmax that evaluate() applies to the decoder's output (/root/reference/model/trainer.py:212,221).
// Round 1 wrote the fp32 logits [M,V] to HBM, read them back in a log_softmax kernel and wrote the log-probs: three
// trips over the largest activation of the path.  Here a thread-block CLUSTER owns 128 rows x all V classes: CTA r of
// the cluster computes the 128 x 128 logits tile of classes [128 r, 128 r + 128) with tcgen05.mma (TMA-fed 3-stage
// ring, fp32 accumulator in tensor memory), reduces its tile to per-row (max, sum exp) straight from TMEM, publishes
// the pair into every peer's shared memory (DSMEM, st.shared::cluster), and after one cluster barrier every CTA knows
// the row's log-sum-exp over all V classes and writes  logit - lse  for its own tile.  V <= 1024 (cluster of <= 8).
#include <cuda.h>
#include <math.h>

#include "common.cuh"
#include "gemm_internal.h"
#include "tcgen05.cuh"

namespace avctc {

constexpr int kHeadThreads = 192;                  // warp 0 TMA, warp 1 MMA + TMEM, warps 2-5: one row per thread
constexpr int kHeadStages = 3;
constexpr int kHeadTile = 128 * 64 * 2;            // bytes of one operand stage (128 rows x 64 bf16)

struct HeadParams {
    int M, V, K, passes;
    const float* bias;
    float* out;                                    // [M,V] fp32 log-probs
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f2(uint32_t local_addr, uint32_t rank, float a, float b) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(remote), "f"(a), "f"(b) : "memory");
}

__global__ void __launch_bounds__(kHeadThreads, 2)
ctc_head_fwd_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                    const HeadParams p) {
    extern __shared__ uint8_t head_smem_raw[];
    const uint32_t raw = smem_u32(head_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* gen = head_smem_raw + (base - raw);
    const uint32_t sA = base, sB = base + kHeadStages * kHeadTile;
    const uint32_t bars = base + 2 * kHeadStages * kHeadTile;          // full[3], empty[3], tmem_full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + 2 * kHeadStages * kHeadTile + 64);
    float2* stats = reinterpret_cast<float2*>(gen + 2 * kHeadStages * kHeadTile + 128);    // [passes][8 ranks][128 rows]
    const uint32_t stats_s = bars + 128;
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (kHeadStages + s); };
    const uint32_t tmem_full = bars + 8u * 2 * kHeadStages;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int nt = gridDim.x;                        // cluster size = number of 128-class tiles
    const int n0 = (int)rank * 128, m0 = blockIdx.y * 128;
    const int num_kb = (p.K + 63) / 64;

    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        for (int s = 0; s < kHeadStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kHeadStages;
                mbar_wait(empty(s), ((kb / kHeadStages) & 1) ^ 1);
                mbar_expect_tx(full(s), 2 * kHeadTile);
                tma_load_3d(sA + s * kHeadTile, &map_x, full(s), kb * 64, m0, 0);
                tma_load_3d(sB + s * kHeadTile, &map_w, full(s), kb * 64, n0, 0);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(128, 0, 0);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kHeadStages;
                mbar_wait(full(s), (kb / kHeadStages) & 1);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base, umma_desc(sA + s * kHeadTile + k * 32, 16), umma_desc(sB + s * kHeadTile + k * 32, 16),
                              idesc, (kb | k) ? 1u : 0u);
                umma_commit(empty(s));
            }
            umma_commit(tmem_full);
        }
        __syncwarp();
    }
    // ---- epilogue: thread = one row of this CTA's tile (TMEM lane); 128 class columns in four 32-wide chunks
    const int q4 = warp & 3, r = q4 * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const int ncols = min(128, p.V - n0);            // valid classes of this tile (>= 1 by construction of the grid)
    float lse = 0.f, lse_b = 0.f;       // lse: log-sum-exp of the logits; lse_b (passes = 2): that of the log-probs themselves
    for (int round = 0; round < p.passes; ++round) {
        if (warp >= 2) {
            if (round == 0) { mbar_wait(tmem_full, 0); tc_fence_after(); }
            float m = AVCTC_NEG_INF;
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t v[32];
                tmem_ld32(taddr + ch * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int c = ch * 32 + j;
                    if (c < ncols) m = fmaxf(m, __uint_as_float(v[j]) + __ldg(p.bias + n0 + c) - lse);
                }
            }
            float s = 0.f;
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t v[32];
                tmem_ld32(taddr + ch * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int c = ch * 32 + j;
                    if (c < ncols) s += __expf(__uint_as_float(v[j]) + __ldg(p.bias + n0 + c) - lse - m);
                }
            }
            // publish (max, sum) of my tile's row r to every CTA of the cluster (my own copy included)
            const uint32_t slot = stats_s + (uint32_t)(((round * 8 + (int)rank) * 128 + r) * 8);
            for (int peer = 0; peer < nt; ++peer) st_cluster_f2(slot, (uint32_t)peer, m, s);
        }
        cluster_sync_all();                          // every thread of every CTA: the stats of this round are visible
        if (warp >= 2) {
            float mm = AVCTC_NEG_INF;
            for (int i = 0; i < nt; ++i) mm = fmaxf(mm, stats[(round * 8 + i) * 128 + r].x);
            float ss = 0.f;
            for (int i = 0; i < nt; ++i) {
                const float2 st = stats[(round * 8 + i) * 128 + r];
                ss += st.y * __expf(st.x - mm);
            }
            if (round == 0) lse = mm + logf(ss);
            else lse_b = mm + logf(ss);
        }
    }
    if (warp >= 2) {
        const int row = m0 + r;
        float* orow = p.out + (size_t)row * p.V + n0;
        const bool vec = ((p.V & 3) == 0);
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
            uint32_t v[32];
            tmem_ld32(taddr + ch * 32, v);
            tmem_ld_wait();
            if (row < p.M) {
                if (vec && ch * 32 + 32 <= ncols) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + ch * 32 + j));
                        float4 o;
                        o.x = (__uint_as_float(v[j]) + b4.x - lse) - lse_b; o.y = (__uint_as_float(v[j + 1]) + b4.y - lse) - lse_b;
                        o.z = (__uint_as_float(v[j + 2]) + b4.z - lse) - lse_b; o.w = (__uint_as_float(v[j + 3]) + b4.w - lse) - lse_b;
                        *reinterpret_cast<float4*>(orow + ch * 32 + j) = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int c = ch * 32 + j;
                        if (c < ncols) orow[c] = (__uint_as_float(v[j]) + __ldg(p.bias + n0 + c) - lse) - lse_b;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 128);
}

}  // namespace avctc

using namespace avctc;

// See include/avctc_b200.h.
extern "C" int avctc_ctc_head_forward(const void* x_bf16, const void* w_bf16, const float* bias, int M, int V, int K,
                                      float* log_probs, int passes, void* stream) {
    if (!x_bf16 || !w_bf16 || !bias || !log_probs || M <= 0 || V <= 0 || K <= 0) return AVCTC_ERR_BAD_ARG;
    if (passes != 1 && passes != 2) return AVCTC_ERR_BAD_ARG;
    if (V > 1024 || (K & 7)) return AVCTC_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(bias) & 15) || (reinterpret_cast<uintptr_t>(log_probs) & 15)) return AVCTC_ERR_ALIGNMENT;
    CUtensorMap mx, mw;
    int rc = avctc_tensor_map(&mx, x_bf16, K, M, 1, K, 0, 64, 128);
    if (rc) return rc;
    rc = avctc_tensor_map(&mw, w_bf16, K, V, 1, K, 0, 64, 128);
    if (rc) return rc;
    HeadParams p;
    p.M = M; p.V = V; p.K = K; p.passes = passes; p.bias = bias; p.out = log_probs;
    const int nt = (V + 127) / 128;
    // one stats round per pass: with passes = 1 (training) two CTAs fit an SM
    const size_t smem = 2 * kHeadStages * kHeadTile + 128 + (size_t)passes * 8 * 128 * sizeof(float2) + 1024;
    static size_t configured = 0;
    if (smem > configured) {
        AVCTC_CUDA_RETURN(cudaFuncSetAttribute(ctc_head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nt, (M + 127) / 128);
    cfg.blockDim = dim3(kHeadThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = reinterpret_cast<cudaStream_t>(stream);
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = nt; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = avctc_tuning_get("pdl", 1) ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 2;
    return (int)cudaLaunchKernelEx(&cfg, ctc_head_fwd_kernel, mx, mw, p);
}

// CTCDecoder backward (decoder.py:24-25 under autograd): dz = dlp - exp(lp) * rowsum(dlp) (bf16, row stride Vp = V
// rounded up to 8), then in ONE grouped tcgen05 launch g_w = dz^T . x (split-K) and dx = dz . W, then g_b = colsum(dz).
extern "C" int avctc_ctc_head_backward(const float* log_probs, const float* dlog_probs, const void* x_bf16,
                                       const void* w_bf16, int M, int V, int K, void* dz_bf16, float* g_w, float* g_b,
                                       void* dx_bf16, void* stream) {
    if (!log_probs || !dlog_probs || !x_bf16 || !w_bf16 || !dz_bf16 || !g_w || !g_b || M <= 0 || V <= 0 || K <= 0)
        return AVCTC_ERR_BAD_ARG;
    if (K & 7) return AVCTC_ERR_UNSUPPORTED;
    const int Vp = (V + 7) / 8 * 8;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (Vp != V) AVCTC_CUDA_RETURN(cudaMemsetAsync(dz_bf16, 0, (size_t)M * Vp * 2, st));    // pad columns stay zero
    int rc = avctc_log_softmax_backward(log_probs, dlog_probs, AVCTC_F32, dz_bf16, M, V, Vp, stream);
    if (rc) return rc;
    auto opnd = [](const void* ptr, long long rows, long long kdim, long long ld, int mn) {
        avctc_gemm_operand o;
        o.ptr = ptr; o.rows = rows; o.kdim = kdim; o.zdim = 1; o.ld = ld; o.zstride = 0;
        o.k_outer = o.k_inner = o.r_outer = o.r_inner = o.z_outer = o.z_inner = 0; o.mn_major = mn;
        return o;
    };
    AvctcGemmJob g[2];
    int n = 0;
    const int mt = (M + 127) / 128;
    int other = 0;
    if (dx_bf16) {      // dx[M,K] = dz[M,V] . W[V,K]
        AvctcGemmJob& j = g[n++];
        j.a = opnd(dz_bf16, M, V, Vp, 0); j.b = opnd(w_bf16, K, V, K, 1);
        j.M = M; j.N = K; j.K = V; j.batch = 1; j.inner_count = 1; j.C = dx_bf16; j.out_dtype = AVCTC_BF16; j.ldc = K;
        j.c_outer = j.c_inner = 0; j.bias = nullptr; j.bias_mode = 0; j.alpha = 1.f; j.accumulate = 0; j.splits = 1;
        other = mt * ((K + 127) / 128);
    }
    {                   // g_w[V,K] = dz[M,V]^T . x[M,K], split-K over M
        AvctcGemmJob& j = g[n++];
        j.a = opnd(dz_bf16, V, M, Vp, 1); j.b = opnd(x_bf16, K, M, K, 1);
        j.M = V; j.N = K; j.K = M; j.batch = 1; j.inner_count = 1; j.C = g_w; j.out_dtype = AVCTC_F32; j.ldc = K;
        j.c_outer = j.c_inner = 0; j.bias = nullptr; j.bias_mode = 0; j.alpha = 1.f; j.accumulate = 0;
        const int tiles = ((V + 127) / 128) * ((K + 127) / 128);
        int splits = (296 - other + tiles - 1) / tiles;
        j.splits = splits < 1 ? 1 : (splits > 8 ? 8 : splits);
    }
    rc = avctc_gemm_launch_group(g, n, stream);
    if (rc) return rc;
    return avctc_colsum(dz_bf16, AVCTC_BF16, M, V, Vp, g_b, 0, stream);
}
