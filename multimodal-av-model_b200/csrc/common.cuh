// common.cuh — shared device helpers for libavctc_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/avctc_b200.h"

#define AVCTC_LOG2E 1.4426950408889634f
#define AVCTC_LN2 0.6931471805599453f
#define AVCTC_LN2_D 0.69314718055994530942
#define AVCTC_NEG_INF (-CUDART_INF_F)

#define AVCTC_CUDA_RETURN(expr)                       \
    do {                                              \
        cudaError_t _e = (expr);                      \
        if (_e != cudaSuccess) return (int)_e;        \
    } while (0)

namespace avctc {

constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// log2-domain log-sum-exp of two / three terms; -inf safe (all -inf -> -inf), 1 or 2 ex2 + 1 lg2.
__device__ __forceinline__ float lse2_log2(float x, float y) {
    const float m = fmaxf(x, y);
    const float n = fminf(x, y);
    const float ms = (m == AVCTC_NEG_INF) ? 0.f : m;
    return m + lg2_approx(1.f + ex2_approx(n - ms));
}
__device__ __forceinline__ float lse3_log2(float x, float y, float z) {
    const float hi = fmaxf(x, y);
    const float lo = fminf(x, y);
    const float m = fmaxf(hi, z);
    const float mid = fminf(hi, z);
    const float ms = (m == AVCTC_NEG_INF) ? 0.f : m;
    return m + lg2_approx(1.f + ex2_approx(lo - ms) + ex2_approx(mid - ms));
}

// ---- sentinel-based variants for the CTC scan: "-inf" is the finite kNegBig (absorbing under every add the
// scan performs), so no NaN guards sit on the dependent chain; FMNMX3 gives the 3-way max in one op and the
// emission is pre-added to the max off the chain:  lse(x,y,z) + v = lg2(sum ex2(. - m)) + (m + v).
#define AVCTC_NEG_BIG (-1.0e30f)
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float lse2_plus(float x, float y, float v) {
    const float m = fmaxf(x, y);
    const float mv = m + v;
    return lg2_approx(ex2_approx(x - m) + ex2_approx(y - m)) + mv;
}
__device__ __forceinline__ float lse3_plus(float x, float y, float z, float v) {
    const float m = fmax3(x, y, z);
    const float mv = m + v;
    return lg2_approx(ex2_approx(x - m) + ex2_approx(y - m) + ex2_approx(z - m)) + mv;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }

__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// Programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start while its predecessor in the stream is still running; pdl_wait() blocks until the predecessor grid has
// completed and its memory is visible, pdl_launch_dependents() lets the successor be scheduled early.  Both are no-ops
// for a normally launched kernel.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ int ld_volatile_shared_s32(const int* p) {
    int v;
    unsigned s = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(s));
    return v;
}
__device__ __forceinline__ void st_volatile_shared_s32(int* p, int v) {
    unsigned s = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(s), "r"(v));
}

__device__ __forceinline__ int4 ld_volatile_shared_v4(const int4* p) {
    int4 v;
    unsigned s = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.volatile.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(s));
    return v;
}
__device__ __forceinline__ void st_volatile_shared_v4(int4* p, int4 v) {
    unsigned s = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("st.volatile.shared.v4.s32 [%0], {%1,%2,%3,%4};" ::"r"(s), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}

// stage one row of V elements into fp32 shared memory at buf[o + c], o = element misalignment of the
// global row start w.r.t. 16 bytes, so that 16-byte global chunks land on 16-byte shared chunks.
__device__ __forceinline__ int stage_row(const float* grow, int V, float* buf, int lane) {
    const int o = (int)((reinterpret_cast<uintptr_t>(grow) >> 2) & 3);
    const int vstart = (o + 3) & ~3, vend = (o + V) & ~3;
    if (vstart < vend) {
        for (int pidx = o + lane; pidx < vstart; pidx += 32) cp_async_4(buf + pidx, grow + (pidx - o));
        for (int q = vstart + 4 * lane; q < vend; q += 128) cp_async_16(buf + q, grow + (q - o));
        for (int pidx = vend + lane; pidx < o + V; pidx += 32) cp_async_4(buf + pidx, grow + (pidx - o));
    } else {
        for (int pidx = o + lane; pidx < o + V; pidx += 32) cp_async_4(buf + pidx, grow + (pidx - o));
    }
    cp_async_commit();
    return o;
}

}  // namespace avctc

// host: launch with the PDL attribute (falls back to a plain launch when the knob "pdl" is 0)
int avctc_tuning_get(const char* key, int dflt);
template <typename... KArgs, typename... Args>
static inline cudaError_t avctc_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                           Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = avctc_tuning_get("pdl", 1) ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// process-global tuning knobs (c_api.cu)
int avctc_tuning_get(const char* key, int dflt);
