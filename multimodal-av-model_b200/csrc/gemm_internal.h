// gemm_internal.h — library-internal launch interface of the tcgen05 GEMM (gemm_tcgen05.cu), shared with the host
// orchestration of the fusion path (fusion_path.cu).  Not part of the public C ABI.
#pragma once
#include "../../include/avctc_b200.h"

struct AvctcGemmJob {
    avctc_gemm_operand a, b;
    int M, N, K, batch, inner_count;
    void* C; int out_dtype; long long ldc, c_outer, c_inner;
    const float* bias; int bias_mode;
    float alpha; int accumulate;
    int splits;        // > 1: split-K with fp32 red.add into a C this call zeroes; < -1: the caller zeroed C already
};

// One launch for up to 6 independent GEMMs (see GemmGroup in gemm_tcgen05.cu).
int avctc_gemm_launch_group(const AvctcGemmJob* jobs, int njobs, void* stream);

int avctc_gemm_launch(const avctc_gemm_operand* a, const avctc_gemm_operand* b, int M, int N, int K, int batch,
                      int inner_count, void* C, int out_dtype, long long ldc, long long c_outer, long long c_inner,
                      const float* bias, int bias_mode, float alpha, int accumulate, int splits, void* stream);

// 3-D bf16 TMA descriptor (dim0 contiguous; dim1 stride ld elements; dim2 stride zstride elements), 128-byte swizzle,
// memoised by (pointer, extents, strides, box): encoding costs ~1 us of host time per map and the fusion path needs ~40
// per step on pointers the caching allocator hands back every step.
#include <cuda.h>
int avctc_tensor_map(CUtensorMap* out, const void* ptr, long long dim0, long long dim1, long long dim2, long long ld,
                     long long zstride, int box0, int box1);
