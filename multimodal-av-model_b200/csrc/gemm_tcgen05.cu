// gemm_tcgen05.cu — the dense contraction of the fusion path on 5th-gen tensor cores (sm_100a).
//
// One kernel serves every GEMM-shaped step of CrossAttentionFusion / CTCDecoder forward and backward
// (/root/reference/model/fusion_module.py:57-63, model/decoder.py:24; nn.MultiheadAttention's projections,
// Q.K^T and P.V, torch/nn/functional.py:5848-5866,6630-6652):
//
//     C[z] = alpha * A[z] * B[z]^T (+ bias)          A: M x K,  B: N x K,  fp32 accumulation in TMEM
//
// Operands are bf16 in HBM, fetched by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a 4-stage
// shared-memory ring; one elected thread issues tcgen05.mma (UMMA 128x128x16, cta_group::1) with the
// accumulator in tensor memory; four epilogue warps read it back with tcgen05.ld, add the bias, convert
// and store.  Either operand may be "K-major" (row-major rows x K, e.g. activations / nn.Linear weights)
// or "MN-major" (row-major K x rows, i.e. the transposed view) so that forward, dX = dY.W and
// dW = dY^T.X all run WITHOUT materialising a transpose; batched problems (attention heads) address their
// slices through TMA coordinates (z -> (outer, inner) -> element offsets), never through copies.
// Warp roles: warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2..9 epilogue (two warps per TMEM lane
// quarter, each taking half of the columns: a lone warp per SM sub-partition issues too slowly to drain the tile).
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "gemm_internal.h"
#include "tcgen05.cuh"

namespace avctc {

constexpr int kBM = 128, kBN = 128, kBK = 64, kStages = 3, kUmmaK = 16;   // 3 stages = 96 KiB: two CTAs per SM
constexpr int kGemmThreads = 320;                  // TMA warp, MMA warp, eight epilogue warps
constexpr int kEpiThreads = kGemmThreads - 64;
constexpr int kTileBytes = kBM * kBK * 2;           // 16 KiB per operand per stage
constexpr int kTmemCols = 128;
constexpr int kStgLd = kBN + 4;                     // fp32 staging tile row stride (bank-conflict-free float4 rows)

struct OperandSpec {      // where batch z's slice starts, in elements of the TMA tensor
    int k_outer, k_inner;   // offset along the reduction dim
    int r_outer, r_inner;   // offset along the row (M or N) dim
    int z_outer, z_inner;   // third TMA coordinate
    int mn_major;           // 0: tensor is rows x K (K contiguous); 1: tensor is K x rows (rows contiguous)
};

struct GemmParams {
    int M, N, K, batch, inner_count;
    OperandSpec a, b;
    void* C; long long ldc, c_outer, c_inner; int out_dtype;
    const float* bias; int bias_mode;   // 0 none, 1 per output column (N), 2 per output row (M)
    float alpha;
    int accumulate;                      // C += result (fp32 output only)
    int splits;                          // split-K: blockIdx.z = z * splits + s; partial sums go to C with red.add.f32
    int dbg;                             // record phase timestamps of CTA (0,0,0) into g_gemm_dbg
};

// Several independent GEMMs in ONE launch ("grouped"): the flat grid is cut into per-job ranges of CTAs.  The steps of the
// fusion path that do not depend on each other (e.g. a layer's dgrad and its wgrad, or the two input projections) share
// a launch, so the SMs see 2-4x more tiles per wave and the chain has fewer launch + prologue + drain latencies.
constexpr int kMaxJobs = 6;
struct GemmGroup {
    int njobs;
    int cta_end[kMaxJobs];          // exclusive prefix of CTAs per job
    int gx[kMaxJobs], gy[kMaxJobs]; // M tiles, N tiles of the job (z = the rest)
    GemmParams p[kMaxJobs];
    CUtensorMap maps[2 * kMaxJobs]; // A, B of job j at [2j], [2j+1]
};

__device__ long long g_gemm_dbg[16];
__device__ long long g_gemm_dbg2[2 * 2048 + 2];   // per-CTA start/end timestamps (debug)
__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define GEMM_DBG(slot) do { if (p.dbg && blockIdx.x == 0) g_gemm_dbg[slot] = gtime(); } while (0)

// UMMA shared-memory descriptor, 128-byte swizzle (cute/arch/mma_sm100_desc.hpp bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2 (SWIZZLE_128B)
// K-major tile (rows x 64 bf16, 128 B per row):   SBO = 8 rows * 128 B = 1024, LBO unused (=1)
// MN-major tile (64 k-rows x 64 bf16 per 64-wide MN block): k-row stride 128 B, SBO = 1024 (8 k-rows),
//   LBO = 8192 (next 64-wide MN block)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int mn_major) {
    const uint64_t lbo = mn_major ? (8192u >> 4) : 1u;
    return (uint64_t)((saddr >> 4) & 0x3fffu) | (lbo << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}

__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_bf16_kernel(const __grid_constant__ GemmGroup grp) {
    int job = 0;
    while (job + 1 < grp.njobs && (int)blockIdx.x >= grp.cta_end[job]) ++job;
    const GemmParams p = grp.p[job];
    const CUtensorMap* map_a = &grp.maps[2 * job];
    const CUtensorMap* map_b = map_a + 1;
    const int cta_local = blockIdx.x - (job ? grp.cta_end[job - 1] : 0);
    const int bx = cta_local % grp.gx[job], by = (cta_local / grp.gx[job]) % grp.gy[job];
    const int bz = cta_local / (grp.gx[job] * grp.gy[job]);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;             // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* gen = smem_raw + (base - raw);
    const uint32_t sA = base, sB = base + kStages * kTileBytes;
    const uint32_t bars = base + 2 * kStages * kTileBytes;     // full[kStages], empty[kStages], tmem_full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + 2 * kStages * kTileBytes + (2 * kStages + 1) * 8);
    float* bias_s = reinterpret_cast<float*>(gen + ((2 * kStages * kTileBytes + (2 * kStages + 1) * 8 + 16 + 15) & ~15));   // [kBN]
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (kStages + s); };
    const uint32_t tmem_full = bars + 8u * 2 * kStages;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) GEMM_DBG(0);
    const int cta_lin = blockIdx.x;
    if (p.dbg && threadIdx.x == 0 && cta_lin < 2048) g_gemm_dbg2[2 * cta_lin] = gtime();
    const int m0 = bx * kBM, n0 = by * kBN;
    const int z = bz / p.splits, split = bz % p.splits;
    const int zo = z / p.inner_count, zi = z % p.inner_count;
    const int total_kb = (p.K + kBK - 1) / kBK;
    const int kb_per = (total_kb + p.splits - 1) / p.splits;
    const int kb0 = split * kb_per;
    const int num_kb = max(0, min(total_kb, kb0 + kb_per) - kb0);   // this CTA's share of the reduction

    pdl_launch_dependents();          // the next kernel of the stream may be scheduled; it waits for this grid itself
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    if (threadIdx.x == 0) GEMM_DBG(1);
    pdl_wait();                       // everything above overlapped the predecessor; global memory is touched from here on

    if (warp == 0) {
        if (lane == 0) {   // ===== TMA producer =====
            const int ak = zo * p.a.k_outer + zi * p.a.k_inner, ar = zo * p.a.r_outer + zi * p.a.r_inner + m0;
            const int az = zo * p.a.z_outer + zi * p.a.z_inner;
            const int bk = zo * p.b.k_outer + zi * p.b.k_inner, br = zo * p.b.r_outer + zi * p.b.r_inner + n0;
            const int bzc = zo * p.b.z_outer + zi * p.b.z_inner;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                mbar_wait(empty(s), ((kb / kStages) & 1) ^ 1);
                mbar_expect_tx(full(s), 2 * kTileBytes);
                const uint32_t da = sA + s * kTileBytes, db = sB + s * kTileBytes;
                const int ko = (kb0 + kb) * kBK;
                if (p.a.mn_major) {   // tensor dims {rows, K, z}: two 64-wide row blocks of 64 k-rows each
                    tma_load_3d(da, map_a, full(s), ar, ak + ko, az);
                    tma_load_3d(da + kTileBytes / 2, map_a, full(s), ar + 64, ak + ko, az);
                } else {              // tensor dims {K, rows, z}
                    tma_load_3d(da, map_a, full(s), ak + ko, ar, az);
                }
                if (p.b.mn_major) {
                    tma_load_3d(db, map_b, full(s), br, bk + ko, bzc);
                    tma_load_3d(db + kTileBytes / 2, map_b, full(s), br + 64, bk + ko, bzc);
                } else {
                    tma_load_3d(db, map_b, full(s), bk + ko, br, bzc);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ===== MMA issuer (one thread) =====
            // instruction descriptor (kind::f16): D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
            // a_major bit15, b_major bit16, N>>3 at [17,23), M>>4 at [24,29)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.a.mn_major & 1) << 15) |
                                   ((uint32_t)(p.b.mn_major & 1) << 16) | ((uint32_t)(kBN >> 3) << 17) |
                                   ((uint32_t)(kBM >> 4) << 24);
            const uint32_t a_step = p.a.mn_major ? (kUmmaK * 128u) >> 4 : (kUmmaK * 2u) >> 4;
            const uint32_t b_step = p.b.mn_major ? (kUmmaK * 128u) >> 4 : (kUmmaK * 2u) >> 4;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                mbar_wait(full(s), (kb / kStages) & 1);
                if (kb == 0) GEMM_DBG(2);
                if (kb == num_kb - 1) GEMM_DBG(3);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t adesc = make_desc(sA + s * kTileBytes, p.a.mn_major);
                const uint64_t bdesc = make_desc(sB + s * kTileBytes, p.b.mn_major);
#pragma unroll
                for (int k = 0; k < kBK / kUmmaK; ++k)
                    umma_bf16(tmem_base, adesc + (uint64_t)(k * a_step), bdesc + (uint64_t)(k * b_step), idesc,
                              (kb | k) ? 1u : 0u);
                umma_commit(empty(s));            // frees the smem stage when these MMAs retire
            }
            umma_commit(tmem_full);               // accumulator complete
        }
        __syncwarp();
    } else {               // ===== epilogue: TMEM -> registers -> shared staging -> coalesced global =====
        if (threadIdx.x < 64 + kBN) {         // per-column bias of this N tile (zero when absent / not the leading split)
            const int i = threadIdx.x - 64;
            bias_s[i] = (split == 0 && p.bias_mode == 1 && n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        mbar_wait(tmem_full, 0);
        if (threadIdx.x == 64) GEMM_DBG(4);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                   // a warp may only touch TMEM lanes [32q, 32q+32)
        const int chalf = (warp - 2) >> 2;        // which half of the tile's columns this warp drains
        const int r_loc = q * 32 + lane;
        const int row = m0 + r_loc;
        const long long coff = (long long)zo * p.c_outer + (long long)zi * p.c_inner;
        const bool lead = (split == 0);           // only one split adds the bias
        const float rb = (lead && p.bias_mode == 2 && row < p.M) ? p.bias[row] : 0.f;
        // every MMA has retired (tmem_full), so the operand stages are free: reuse them as a [128][kStgLd] fp32 tile
        float* stg = reinterpret_cast<float*>(gen);
        if (num_kb > 0) {
#pragma unroll 1
            for (int c = chalf * (kBN / 64); c < (chalf + 1) * (kBN / 64); ++c) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float* srow = stg + r_loc * kStgLd + c * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c * 32 + j);
                    float4 o;
                    o.x = fmaf(__uint_as_float(v[j]), p.alpha, rb + b4.x);
                    o.y = fmaf(__uint_as_float(v[j + 1]), p.alpha, rb + b4.y);
                    o.z = fmaf(__uint_as_float(v[j + 2]), p.alpha, rb + b4.z);
                    o.w = fmaf(__uint_as_float(v[j + 3]), p.alpha, rb + b4.w);
                    *reinterpret_cast<float4*>(srow + j) = o;
                }
            }
        }
        if (threadIdx.x == 64) GEMM_DBG(7);
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");       // the epilogue warps only
        if (threadIdx.x == 64) GEMM_DBG(8);
        if (num_kb > 0) {
            const int te = threadIdx.x - 64;                   // 0..kEpiThreads-1
            const int rows_valid = min(kBM, p.M - m0), cols_valid = min(kBN, p.N - n0);
            const size_t esz = (p.out_dtype == AVCTC_F32) ? 4 : 2;
            const long long tile_off = coff + (long long)m0 * p.ldc + n0;
            const bool vec_ok = (cols_valid == kBN) && ((p.ldc * (long long)esz) % 16 == 0) &&
                                (((reinterpret_cast<uintptr_t>(p.C) + (uintptr_t)tile_off * esz) & 15) == 0);
            if (vec_ok && p.out_dtype == AVCTC_BF16) {
                __nv_bfloat16* Cb = reinterpret_cast<__nv_bfloat16*>(p.C) + tile_off;
#pragma unroll 4
                for (int i = 0; i < (kBM * kBN / 8) / kEpiThreads; ++i) {   // 16 threads cover one 256-byte output row
                    const int idx = te + kEpiThreads * i, r = idx >> 4, c8 = (idx & 15) * 8;
                    if (r < rows_valid) {
                        const float4 x = *reinterpret_cast<const float4*>(stg + r * kStgLd + c8);
                        const float4 y = *reinterpret_cast<const float4*>(stg + r * kStgLd + c8 + 4);
                        uint4 pk;
                        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
                        h[0] = __floats2bfloat162_rn(x.x, x.y); h[1] = __floats2bfloat162_rn(x.z, x.w);
                        h[2] = __floats2bfloat162_rn(y.x, y.y); h[3] = __floats2bfloat162_rn(y.z, y.w);
                        *reinterpret_cast<uint4*>(Cb + (long long)r * p.ldc + c8) = pk;
                    }
                }
            } else if (vec_ok) {
                float* Cf = reinterpret_cast<float*>(p.C) + tile_off;
#pragma unroll 4
                for (int i = 0; i < (kBM * kBN / 4) / kEpiThreads; ++i) {   // one warp covers one 512-byte output row
                    const int idx = te + kEpiThreads * i, r = idx >> 5, c4 = (idx & 31) * 4;
                    if (r < rows_valid) {
                        float4 x = *reinterpret_cast<const float4*>(stg + r * kStgLd + c4);
                        float4* dst = reinterpret_cast<float4*>(Cf + (long long)r * p.ldc + c4);
                        if (p.splits > 1) atomicAdd(dst, x);   // split-K partial sum: C was zeroed by the host
                        else {
                            if (p.accumulate) { const float4 o = *dst; x.x += o.x; x.y += o.y; x.z += o.z; x.w += o.w; }
                            *dst = x;
                        }
                    }
                }
            } else {                                           // ragged tile / unaligned C: element-wise
                for (int idx = te; idx < kBM * kBN; idx += kEpiThreads) {
                    const int r = idx / kBN, c = idx % kBN;
                    if (r >= rows_valid || c >= cols_valid) continue;
                    const float x = stg[r * kStgLd + c];
                    const long long off = tile_off + (long long)r * p.ldc + c;
                    if (p.out_dtype == AVCTC_F32) {
                        float* dst = reinterpret_cast<float*>(p.C) + off;
                        if (p.splits > 1) atomicAdd(dst, x);
                        else *dst = p.accumulate ? *dst + x : x;
                    } else {
                        reinterpret_cast<__nv_bfloat16*>(p.C)[off] = __float2bfloat16(x);
                    }
                }
            }
        }
    }
    if (threadIdx.x == 64) GEMM_DBG(5);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) GEMM_DBG(6);
    if (p.dbg && threadIdx.x == 0 && cta_lin < 2048) g_gemm_dbg2[2 * cta_lin + 1] = gtime();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int get_encode() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess) return (int)e;
    if (!fn || q != cudaDriverEntryPointSuccess) return (int)cudaErrorNotSupported;
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return 0;
}

}  // namespace avctc

using namespace avctc;

namespace {
struct MapKey {
    const void* ptr; long long d0, d1, d2, ld, zs; int b0, b1;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && ld == o.ld && zs == o.zs && b0 == o.b0 && b1 == o.b1;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        for (long long v : {k.d0, k.d1, k.d2, k.ld, k.zs, (long long)k.b0, (long long)k.b1})
            h = h * 1099511628211ull ^ (size_t)v;
        return h;
    }
};
std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;
}  // namespace

int avctc_tensor_map(CUtensorMap* m, const void* ptr, long long dim0, long long dim1, long long dim2, long long ld,
                     long long zstride, int box0, int box1) {
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld * 2) % 16 || (dim2 > 1 && (zstride * 2) % 16))
        return AVCTC_ERR_ALIGNMENT;
    int rc = get_encode();
    if (rc) return rc;
    const MapKey key{ptr, dim0, dim1, dim2, ld, zstride, box0, box1};
    {
        std::lock_guard<std::mutex> lk(g_map_mu);
        auto it = g_map_cache.find(key);
        if (it != g_map_cache.end()) { *m = it->second; return 0; }
    }
    cuuint64_t dims[3] = {(cuuint64_t)dim0, (cuuint64_t)dim1, (cuuint64_t)(dim2 > 0 ? dim2 : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(dim2 > 1 ? zstride : ld * dim1) * 2};
    cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return AVCTC_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(g_map_mu);
    if (g_map_cache.size() > 8192) g_map_cache.clear();
    g_map_cache.emplace(key, *m);
    return 0;
}

// debug only (not part of the public header): phase timestamps (ns) of CTA (0,0,0) of the last launch with gemm_dbg=1
extern "C" __attribute__((visibility("default"))) int avctc_debug_gemm_timestamps(long long* host_out16) {
    return (int)cudaMemcpyFromSymbol(host_out16, g_gemm_dbg, sizeof(long long) * 16);
}
extern "C" __attribute__((visibility("default"))) int avctc_debug_gemm_cta_times(long long* host_out, int n_ctas) {
    return (int)cudaMemcpyFromSymbol(host_out, g_gemm_dbg2, sizeof(long long) * 2 * (n_ctas < 2048 ? n_ctas : 2048));
}

// See include/avctc_b200.h for the argument contract.
extern "C" int avctc_gemm_bf16(const avctc_gemm_operand* a, const avctc_gemm_operand* b, int M, int N, int K, int batch,
                               int inner_count, void* C, int out_dtype, long long ldc, long long c_outer,
                               long long c_inner, const float* bias, int bias_mode, float alpha, int accumulate,
                               void* stream) {
    return avctc_gemm_launch(a, b, M, N, K, batch, inner_count, C, out_dtype, ldc, c_outer, c_inner, bias, bias_mode,
                             alpha, accumulate, 1, stream);
}

int avctc_gemm_launch(const avctc_gemm_operand* a, const avctc_gemm_operand* b, int M, int N, int K, int batch,
                      int inner_count, void* C, int out_dtype, long long ldc, long long c_outer, long long c_inner,
                      const float* bias, int bias_mode, float alpha, int accumulate, int splits, void* stream) {
    if (!a || !b) return AVCTC_ERR_BAD_ARG;
    AvctcGemmJob j;
    j.a = *a; j.b = *b; j.M = M; j.N = N; j.K = K; j.batch = batch; j.inner_count = inner_count;
    j.C = C; j.out_dtype = out_dtype; j.ldc = ldc; j.c_outer = c_outer; j.c_inner = c_inner;
    j.bias = bias; j.bias_mode = bias_mode; j.alpha = alpha; j.accumulate = accumulate; j.splits = splits;
    return avctc_gemm_launch_group(&j, 1, stream);
}

int avctc_gemm_launch_group(const AvctcGemmJob* jobs, int njobs, void* stream) {
    if (!jobs || njobs < 1 || njobs > kMaxJobs) return AVCTC_ERR_BAD_ARG;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    GemmGroup g;
    g.njobs = njobs;
    int total = 0;
    const int dbg = avctc_tuning_get("gemm_dbg", 0);
    for (int ji = 0; ji < njobs; ++ji) {
        const AvctcGemmJob& J = jobs[ji];
        if (!J.C || J.M <= 0 || J.N <= 0 || J.K <= 0 || J.batch <= 0 || J.inner_count <= 0) return AVCTC_ERR_BAD_ARG;
        if (J.out_dtype != AVCTC_F32 && J.out_dtype != AVCTC_BF16) return AVCTC_ERR_BAD_ARG;
        if (J.accumulate && J.out_dtype != AVCTC_F32) return AVCTC_ERR_BAD_ARG;
        if (J.bias_mode < 0 || J.bias_mode > 2 || (J.bias_mode && !J.bias)) return AVCTC_ERR_BAD_ARG;
        GemmParams& p = g.p[ji];
        const avctc_gemm_operand* ops[2] = {&J.a, &J.b};
        OperandSpec* specs[2] = {&p.a, &p.b};
        for (int i = 0; i < 2; ++i) {
            const avctc_gemm_operand* o = ops[i];
            if (!o->ptr) return AVCTC_ERR_BAD_ARG;
            // K-major: tensor is rows x K -> TMA dims {K, rows, z}, box {64 k, 128 rows}
            // MN-major: tensor is K x rows -> TMA dims {rows, K, z}, box {64 rows, 64 k}
            const int rc = o->mn_major
                ? avctc_tensor_map(&g.maps[2 * ji + i], o->ptr, o->rows, o->kdim, o->zdim, o->ld, o->zstride, 64, kBK)
                : avctc_tensor_map(&g.maps[2 * ji + i], o->ptr, o->kdim, o->rows, o->zdim, o->ld, o->zstride, kBK, kBM);
            if (rc) return rc;
            specs[i]->k_outer = o->k_outer; specs[i]->k_inner = o->k_inner;
            specs[i]->r_outer = o->r_outer; specs[i]->r_inner = o->r_inner;
            specs[i]->z_outer = o->z_outer; specs[i]->z_inner = o->z_inner;
            specs[i]->mn_major = o->mn_major ? 1 : 0;
        }
        p.M = J.M; p.N = J.N; p.K = J.K; p.batch = J.batch; p.inner_count = J.inner_count;
        p.C = J.C; p.ldc = J.ldc; p.c_outer = J.c_outer; p.c_inner = J.c_inner; p.out_dtype = J.out_dtype;
        p.bias = J.bias; p.bias_mode = J.bias_mode; p.alpha = J.alpha; p.accumulate = J.accumulate;
        int splits = J.splits;
        const int total_kb = (J.K + kBK - 1) / kBK;
        const bool prezeroed = splits < 0;      // negative: |splits|-way split-K into a C the caller already zeroed
        if (prezeroed) splits = -splits;
        if (splits < 1) splits = 1;
        if (splits > total_kb) splits = total_kb;
        const int kb_per = (total_kb + splits - 1) / splits;
        splits = (total_kb + kb_per - 1) / kb_per;            // every split owns at least one k-block
        if (splits > 1) {
            if (J.out_dtype != AVCTC_F32 || J.accumulate || J.batch != 1) return AVCTC_ERR_BAD_ARG;
            if (!prezeroed)
                AVCTC_CUDA_RETURN(cudaMemset2DAsync(J.C, sizeof(float) * (size_t)J.ldc, 0, sizeof(float) * (size_t)J.N,
                                                    (size_t)J.M, st));
        }
        p.splits = splits;
        p.dbg = dbg;
        g.gx[ji] = (J.M + kBM - 1) / kBM;
        g.gy[ji] = (J.N + kBN - 1) / kBN;
        total += g.gx[ji] * g.gy[ji] * J.batch * splits;
        g.cta_end[ji] = total;
    }
    for (int ji = njobs; ji < kMaxJobs; ++ji) { g.cta_end[ji] = total; g.gx[ji] = g.gy[ji] = 1; }
    const size_t smem = 2 * kStages * kTileBytes + (2 * kStages + 1) * 8 + 32 + kBN * sizeof(float) + 1024;
    static bool configured = false;
    if (!configured) {
        AVCTC_CUDA_RETURN(cudaFuncSetAttribute(gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    return (int)avctc_launch_pdl(gemm_bf16_kernel, dim3(total), dim3(kGemmThreads), smem, st, g);
}
