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

#include "common.cuh"

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

__device__ long long g_gemm_dbg[16];
__device__ long long g_gemm_dbg2[2 * 2048 + 2];   // per-CTA start/end timestamps (debug)
__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define GEMM_DBG(slot) do { if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_gemm_dbg[slot] = gtime(); } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc),
        "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

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
gemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const GemmParams p) {
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
    const int cta_lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    if (p.dbg && threadIdx.x == 0 && cta_lin < 2048) g_gemm_dbg2[2 * cta_lin] = gtime();
    const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * kBN;
    const int z = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
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
            const int bz = zo * p.b.z_outer + zi * p.b.z_inner;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                mbar_wait(empty(s), ((kb / kStages) & 1) ^ 1);
                mbar_expect_tx(full(s), 2 * kTileBytes);
                const uint32_t da = sA + s * kTileBytes, db = sB + s * kTileBytes;
                const int ko = (kb0 + kb) * kBK;
                if (p.a.mn_major) {   // tensor dims {rows, K, z}: two 64-wide row blocks of 64 k-rows each
                    tma_load_3d(da, &map_a, full(s), ar, ak + ko, az);
                    tma_load_3d(da + kTileBytes / 2, &map_a, full(s), ar + 64, ak + ko, az);
                } else {              // tensor dims {K, rows, z}
                    tma_load_3d(da, &map_a, full(s), ak + ko, ar, az);
                }
                if (p.b.mn_major) {
                    tma_load_3d(db, &map_b, full(s), br, bk + ko, bz);
                    tma_load_3d(db + kTileBytes / 2, &map_b, full(s), br + 64, bk + ko, bz);
                } else {
                    tma_load_3d(db, &map_b, full(s), bk + ko, br, bz);
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

// 3-D bf16 tensor: dim0 (contiguous) x dim1 (stride ld elements) x dim2 (stride zstride elements)
static int make_map(CUtensorMap* m, const void* ptr, long long dim0, long long dim1, long long dim2, long long ld,
                    long long zstride, int box0, int box1) {
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld * 2) % 16 || (dim2 > 1 && (zstride * 2) % 16))
        return AVCTC_ERR_ALIGNMENT;
    cuuint64_t dims[3] = {(cuuint64_t)dim0, (cuuint64_t)dim1, (cuuint64_t)(dim2 > 0 ? dim2 : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(dim2 > 1 ? zstride : ld * dim1) * 2};
    cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : AVCTC_ERR_BAD_ARG;
}

}  // namespace avctc

using namespace avctc;

// Internal launcher shared with fusion_path.cu.  splits > 1: split-K over blockIdx.z with fp32 red.add into a C that
// this function zeroes first (fp32 output, no accumulate, batch C slices must be disjoint).
int avctc_gemm_launch(const avctc_gemm_operand* a, const avctc_gemm_operand* b, int M, int N, int K, int batch,
                      int inner_count, void* C, int out_dtype, long long ldc, long long c_outer, long long c_inner,
                      const float* bias, int bias_mode, float alpha, int accumulate, int splits, void* stream);

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
    if (!a || !b || !C || M <= 0 || N <= 0 || K <= 0 || batch <= 0 || inner_count <= 0) return AVCTC_ERR_BAD_ARG;
    if (out_dtype != AVCTC_F32 && out_dtype != AVCTC_BF16) return AVCTC_ERR_BAD_ARG;
    if (accumulate && out_dtype != AVCTC_F32) return AVCTC_ERR_BAD_ARG;
    if (bias_mode < 0 || bias_mode > 2 || (bias_mode && !bias)) return AVCTC_ERR_BAD_ARG;
    int rc = get_encode();
    if (rc) return rc;
    CUtensorMap ma, mb;
    const avctc_gemm_operand* ops[2] = {a, b};
    CUtensorMap* maps[2] = {&ma, &mb};
    GemmParams p;
    OperandSpec* specs[2] = {&p.a, &p.b};
    for (int i = 0; i < 2; ++i) {
        const avctc_gemm_operand* o = ops[i];
        if (!o->ptr) return AVCTC_ERR_BAD_ARG;
        // K-major: tensor is rows x K -> TMA dims {K, rows, z}, box {64 k, 128 rows}
        // MN-major: tensor is K x rows -> TMA dims {rows, K, z}, box {64 rows, 64 k}
        rc = o->mn_major ? make_map(maps[i], o->ptr, o->rows, o->kdim, o->zdim, o->ld, o->zstride, 64, kBK)
                         : make_map(maps[i], o->ptr, o->kdim, o->rows, o->zdim, o->ld, o->zstride, kBK, kBM);
        if (rc) return rc;
        specs[i]->k_outer = o->k_outer; specs[i]->k_inner = o->k_inner;
        specs[i]->r_outer = o->r_outer; specs[i]->r_inner = o->r_inner;
        specs[i]->z_outer = o->z_outer; specs[i]->z_inner = o->z_inner;
        specs[i]->mn_major = o->mn_major ? 1 : 0;
    }
    p.M = M; p.N = N; p.K = K; p.batch = batch; p.inner_count = inner_count;
    p.C = C; p.ldc = ldc; p.c_outer = c_outer; p.c_inner = c_inner; p.out_dtype = out_dtype;
    p.bias = bias; p.bias_mode = bias_mode; p.alpha = alpha; p.accumulate = accumulate;
    {
        const int total_kb = (K + kBK - 1) / kBK;
        const bool prezeroed = splits < 0;      // negative: |splits|-way split-K into a C the caller already zeroed
        if (prezeroed) splits = -splits;
        if (splits < 1) splits = 1;
        if (splits > total_kb) splits = total_kb;
        const int kb_per = (total_kb + splits - 1) / splits;
        splits = (total_kb + kb_per - 1) / kb_per;            // every split owns at least one k-block
        if (splits > 1) {
            if (out_dtype != AVCTC_F32 || accumulate || batch != 1) return AVCTC_ERR_BAD_ARG;
            if (!prezeroed) AVCTC_CUDA_RETURN(cudaMemset2DAsync(C, sizeof(float) * (size_t)ldc, 0, sizeof(float) * (size_t)N, (size_t)M,
                                                reinterpret_cast<cudaStream_t>(stream)));
        }
        p.splits = splits;
        p.dbg = avctc_tuning_get("gemm_dbg", 0);
    }
    const size_t smem = 2 * kStages * kTileBytes + (2 * kStages + 1) * 8 + 32 + kBN * sizeof(float) + 1024;
    static bool configured = false;
    if (!configured) {
        AVCTC_CUDA_RETURN(cudaFuncSetAttribute(gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid((M + kBM - 1) / kBM, (N + kBN - 1) / kBN, batch * p.splits);
    return (int)avctc_launch_pdl(gemm_bf16_kernel, grid, dim3(kGemmThreads), smem, reinterpret_cast<cudaStream_t>(stream), ma, mb, p);
}
