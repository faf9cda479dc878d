// lstm.cu — 2-layer bidirectional LSTM (CrossAttentionFusion.temporal_model) as persistent sm_100a kernels.
//
// Replaces nn.LSTM(input_size=E, hidden_size=E, num_layers=2, batch_first=True, bidirectional=True)
// (/root/reference/model/fusion_module.py:21-27, called at :64 over ALL padded frames, zero initial state) and its
// autograd.  cuDNN runs this recurrence as ~600 tiny kernels per call (one GEMM + one pointwise kernel per
// step, layer and direction); here a layer is
//   * one tcgen05 GEMM for the input projections of all frames and both directions  (x W_ih^T + b_ih + b_hh),
//   * ONE persistent cooperative kernel for the recurrence: CTA (direction, c) owns 16 hidden units, keeps its
//     64 gate rows of W_hh as mma.sync A fragments in REGISTERS for the whole sequence, and per step
//       - reads h_{t-1} (all units, written by the other CTAs of its direction) from the output tensor itself,
//       - multiplies (mma.sync m16n8k16 bf16, fp32 accumulate; batch is the N dimension),
//       - applies the gates, keeps c in registers, writes h_t,
//       - passes a per-direction step barrier (one atomic + acquire poll in L2),
//   * backward: the mirrored persistent kernel (gate gradients exchanged through the dG tensor, W_hh^T fragments in
//     registers), then four tcgen05 GEMMs (dW_ih, dW_hh, dX) and a column sum (biases).
// Gate order i, f, g, o and every formula follow torch.nn.LSTM; operands are bf16 (what autocast feeds cuDNN as
// fp16 in the reference), accumulation and cell state fp32.
#include <cooperative_groups.h>

#include "common.cuh"

int avctc_gemm_launch(const avctc_gemm_operand* a, const avctc_gemm_operand* b, int M, int N, int K, int batch,
                      int inner_count, void* C, int out_dtype, long long ldc, long long c_outer, long long c_inner,
                      const float* bias, int bias_mode, float alpha, int accumulate, int splits, void* stream);

namespace avctc {

constexpr int kLstmThreads = 256;
constexpr int kLstmUnits = 16;        // hidden units per CTA (one m16 tile per gate)

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void step_barrier_wait(const unsigned* bar, unsigned need) {
    int spins = 0;
    while (ld_acquire_gpu(bar) < need) {
        if (++spins > (1 << 24)) __trap();      // never hang the GPU
    }
}
// Sentinel exchange: the host fills an exchange buffer with 0xFF bytes (bf16 0xFFFF, a NaN pattern no conversion
// produces: cvt.rn.bf16.f32 gives the canonical 0x7FFF); a reader polls the 16-byte chunk itself until none of its
// eight halves is the sentinel, so a step needs no fence + counter + counter poll (one L2 round trip instead of three).
__device__ __forceinline__ bool has_sentinel16(const uint4& v) {
    const uint32_t a = ~v.x, b = ~v.y, c = ~v.z, d = ~v.w;             // a half equal to 0xFFFF is zero in the complement
    const uint32_t z = ((a - 0x00010001u) & ~a) | ((b - 0x00010001u) & ~b) | ((c - 0x00010001u) & ~c) | ((d - 0x00010001u) & ~d);
    return (z & 0x80008000u) != 0u;
}
__device__ __forceinline__ uint4 ld_relaxed_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// loads n 16-byte chunks (chunk c of this thread: src(c) -> dst(c)), four in flight, re-polling the ones not yet written
template <typename SrcF, typename DstF>
__device__ __forceinline__ void poll_load_chunks(int first, int n, int stride, SrcF src, DstF dst) {
    for (int c0 = first; c0 < n; c0 += 4 * stride) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int c = c0 + u * stride; if (c < n) v[u] = ld_relaxed_v4(src(c)); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = c0 + u * stride;
            if (c < n) {
                int spins = 0;
                while (has_sentinel16(v[u])) {
                    v[u] = ld_relaxed_v4(src(c));
                    if (++spins > (1 << 22)) __trap();          // never hang the GPU
                }
                *dst(c) = v[u];
            }
        }
    }
}
// one MUFU each (tanh.approx.f32, |err| ~ 5e-4): well inside the bf16 rounding of h; sigmoid(x) = 0.5 tanh(x/2) + 0.5
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoidf_(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 lo, __nv_bfloat16 hi) {
    return (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
}

struct LstmFwdParams {
    const float* xproj;          // [B*T][8H] fp32: x W_ih^T + b_ih + b_hh; direction d at column d*4H, gate g at +g*H
    const __nv_bfloat16* whh;    // [2][4H][H]
    __nv_bfloat16* y;            // [B][T][2H]; also the h exchange buffer between CTAs
    __nv_bfloat16* hprev;        // [B*T][2H]: the h_{t-1} step t consumed (saved for dW_hh), or nullptr
    float* gates;                // [2][T][B][4][H] post-activation i,f,g,o, or nullptr
    float* cst;                  // [2][T][B][H] cell state, or nullptr
    unsigned* bar;               // [2] step counters, zeroed by the host
    int B, T;                    // B = sequences of the whole call (array strides)
    int Bg;                      // sequences per batch group: blockIdx.y handles sequences [Bg*blockIdx.y, +Bg) on its own
                                 // set of 2*H/16 CTAs with its own step counters (sequences are independent)
    int tagged;                  // 1: y was filled with the 0xFFFF sentinel by the host; exchange by polling the data
};

// NB = batch tiles of 8 (B <= 8*NB)
template <int H, int NB>
__global__ void __launch_bounds__(kLstmThreads, 1) lstm_fwd_kernel(const LstmFwdParams p) {
    constexpr int NC = H / kLstmUnits;               // CTAs per direction
    constexpr int KS = H / 16 / 2;                   // k-steps per warp (two warps share a gate, split over K)
    constexpr int HS = H + 8;                        // padded smem row (bank-conflict-free B fragments)
    constexpr int BP = 8 * NB;
    constexpr int PP = (kLstmUnits * BP + kLstmThreads - 1) / kLstmThreads;   // (unit,batch) pairs per thread
    extern __shared__ __align__(16) unsigned char lstm_smem[];
    __nv_bfloat16* hs = reinterpret_cast<__nv_bfloat16*>(lstm_smem);                        // [BP][HS]
    float* red = reinterpret_cast<float*>(lstm_smem + (size_t)BP * HS * 2);                  // [2][4][16][BP]
    const int d = blockIdx.x / NC, cta = blockIdx.x % NC;
    const int u0 = cta * kLstmUnits;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int gate = warp & 3, khalf = warp >> 2;
    const int T = p.T, BS = p.B;                         // BS: batch stride of the [.,T,B,.] arrays
    const int b0 = blockIdx.y * p.Bg;                     // my batch group
    const int B = min(p.Bg, p.B - b0);
    const unsigned* const bar = p.bar + 2 * blockIdx.y;

    // W_hh rows of my gate and units, my K half: A fragments for the whole sequence
    uint32_t afrag[KS][4];
    {
        const __nv_bfloat16* W = p.whh + ((size_t)d * 4 * H + (size_t)gate * H + u0) * H;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int k0 = (khalf * KS + ks) * 16 + tig * 2;
            afrag[ks][0] = *reinterpret_cast<const uint32_t*>(W + (size_t)gid * H + k0);
            afrag[ks][1] = *reinterpret_cast<const uint32_t*>(W + (size_t)(gid + 8) * H + k0);
            afrag[ks][2] = *reinterpret_cast<const uint32_t*>(W + (size_t)gid * H + k0 + 8);
            afrag[ks][3] = *reinterpret_cast<const uint32_t*>(W + (size_t)(gid + 8) * H + k0 + 8);
        }
    }
    float creg[PP];
    __nv_bfloat16 hcur_reg[PP], hprev_reg[PP];
#pragma unroll
    for (int i = 0; i < PP; ++i) { creg[i] = 0.f; hcur_reg[i] = __float2bfloat16(0.f); hprev_reg[i] = __float2bfloat16(0.f); }
    for (int i = tid; i < BP * HS; i += kLstmThreads) hs[i] = __float2bfloat16(0.f);
    __syncthreads();

    // input projections of my pairs, fetched ONE STEP AHEAD so their L2 latency never sits on the recurrence
    float xnext[PP][4];
    auto load_x = [&](int step) {
        const int tt = d ? (T - 1 - step) : step;
#pragma unroll
        for (int i = 0; i < PP; ++i) {
            const int e = tid + kLstmThreads * i, u = e & 15, b = e >> 4;
#pragma unroll
            for (int g = 0; g < 4; ++g)
                xnext[i][g] = (b < B && step < T) ? __ldg(p.xproj + ((size_t)(b0 + b) * T + tt) * 8 * H + (size_t)d * 4 * H + g * H + u0 + u) : 0.f;
        }
    };
    load_x(0);
    for (int s = 0; s < T; ++s) {
        const int t = d ? (T - 1 - s) : s;
        const int tprev = d ? t + 1 : t - 1;
        float xg[PP][4], sv[PP][4];
#pragma unroll
        for (int i = 0; i < PP; ++i)
#pragma unroll
            for (int g = 0; g < 4; ++g) xg[i][g] = xnext[i][g];
        load_x(s + 1);
        if (s > 0) {
            if (p.tagged) {                                              // h_{t-1}: poll the data itself
                poll_load_chunks(tid, B * (H / 8), kLstmThreads,
                    [&](int c) { const int b = c / (H / 8), q = c % (H / 8);
                                 return reinterpret_cast<const uint4*>(p.y + ((size_t)(b0 + b) * T + tprev) * 2 * H + (size_t)d * H) + q; },
                    [&](int c) { const int b = c / (H / 8), q = c % (H / 8); return reinterpret_cast<uint4*>(hs + b * HS + q * 8); });
            } else {
                if (tid == 0) step_barrier_wait(bar + d, (unsigned)NC * s);
                __syncthreads();
                for (int c = tid; c < B * (H / 8); c += kLstmThreads) {    // h_{t-1}: L2 loads (written by other SMs)
                    const int b = c / (H / 8), q = c % (H / 8);
                    const uint4 v = __ldcg(reinterpret_cast<const uint4*>(p.y + ((size_t)(b0 + b) * T + tprev) * 2 * H + (size_t)d * H) + q);
                    *reinterpret_cast<uint4*>(hs + b * HS + q * 8) = v;
                }
            }
            __syncthreads();
        }
        float accp[4][NB][4];          // four independent accumulator chains (mma.sync latency), summed below
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4)
#pragma unroll
            for (int n = 0; n < NB; ++n)
#pragma unroll
                for (int j = 0; j < 4; ++j) accp[c4][n][j] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int k0 = (khalf * KS + ks) * 16 + tig * 2;
#pragma unroll
            for (int n = 0; n < NB; ++n) {
                const __nv_bfloat16* hb = hs + (n * 8 + gid) * HS + k0;
                mma_bf16_16816(accp[ks & 3][n], afrag[ks], *reinterpret_cast<const uint32_t*>(hb),
                               *reinterpret_cast<const uint32_t*>(hb + 8));
            }
        }
        float acc[NB][4];
#pragma unroll
        for (int n = 0; n < NB; ++n)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[n][j] = (accp[0][n][j] + accp[1][n][j]) + (accp[2][n][j] + accp[3][n][j]);
        {   // partial sums -> red[khalf][gate][unit][batch]
            float* r = red + ((size_t)(khalf * 4 + gate) * 16) * BP;
#pragma unroll
            for (int n = 0; n < NB; ++n) {
                const int bc = n * 8 + tig * 2;
                r[gid * BP + bc] = acc[n][0]; r[gid * BP + bc + 1] = acc[n][1];
                r[(gid + 8) * BP + bc] = acc[n][2]; r[(gid + 8) * BP + bc + 1] = acc[n][3];
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < PP; ++i) {
            const int e = tid + kLstmThreads * i, u = e & 15, b = e >> 4;
            if (b < B) {
                float pre[4];
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    pre[g] = red[((size_t)g * 16 + u) * BP + b] + red[((size_t)(4 + g) * 16 + u) * BP + b] + xg[i][g];
                const float ig = sigmoidf_(pre[0]), fg = sigmoidf_(pre[1]), gg = tanh_fast(pre[2]), og = sigmoidf_(pre[3]);
                const float c = fg * creg[i] + ig * gg;
                creg[i] = c;
                const float h = og * tanh_fast(c);
                hprev_reg[i] = hcur_reg[i];                     // the h_{t-1} of my own unit that this step consumed
                hcur_reg[i] = __float2bfloat16(h);
                p.y[((size_t)(b0 + b) * T + t) * 2 * H + (size_t)d * H + u0 + u] = hcur_reg[i];
                sv[i][0] = ig; sv[i][1] = fg; sv[i][2] = gg; sv[i][3] = og;
            }
        }
        if (!p.tagged) {
            __syncthreads();
            if (tid == 0) { __threadfence(); atomicAdd(const_cast<unsigned*>(bar) + d, 1u); }
        }
        // everything only the backward pass needs is stored AFTER the arrive: the fence above must not wait for it
        if (p.gates) {
#pragma unroll
            for (int i = 0; i < PP; ++i) {
                const int e = tid + kLstmThreads * i, u = e & 15, b = e >> 4;
                if (b < B) {
                    float* gp = p.gates + (((size_t)d * T + t) * BS + b0 + b) * 4 * H + u0 + u;
                    gp[0] = sv[i][0]; gp[H] = sv[i][1]; gp[2 * H] = sv[i][2]; gp[3 * H] = sv[i][3];
                    p.cst[(((size_t)d * T + t) * BS + b0 + b) * H + u0 + u] = creg[i];
                    p.hprev[((size_t)(b0 + b) * T + t) * 2 * H + (size_t)d * H + u0 + u] = hprev_reg[i];
                }
            }
        }
    }
}

struct LstmBwdParams {
    const __nv_bfloat16* dy;     // [B][T][2H] gradient w.r.t. the layer output
    const __nv_bfloat16* whh;    // [2][4H][H]
    const float* gates;          // [2][T][B][4][H]
    const float* cst;            // [2][T][B][H]
    __nv_bfloat16* dG;           // [B*T][8H] gate pre-activation gradients (output; also the exchange buffer)
    unsigned* bar;               // [2]
    int B, T, Bg;                // as in LstmFwdParams
    int tagged;                  // 1: dG was filled with the 0xFFFF sentinel by the host; exchange by polling the data
};

template <int H, int NB>
__global__ void __launch_bounds__(kLstmThreads, 1) lstm_bwd_kernel(const LstmBwdParams p) {
    constexpr int NC = H / kLstmUnits;
    constexpr int KS = 4 * H / 16 / 8;               // k-steps per warp over K = 4H (8 warps split K)
    constexpr int GS = 4 * H + 8;                    // padded smem row of dG
    constexpr int BP = 8 * NB;
    constexpr int PP = (kLstmUnits * BP + kLstmThreads - 1) / kLstmThreads;
    extern __shared__ __align__(16) unsigned char lstm_smem[];
    __nv_bfloat16* dgs = reinterpret_cast<__nv_bfloat16*>(lstm_smem);                        // [BP][GS]
    float* red = reinterpret_cast<float*>(lstm_smem + (size_t)BP * GS * 2);                  // [8][16][BP]
    const int d = blockIdx.x / NC, cta = blockIdx.x % NC;
    const int u0 = cta * kLstmUnits;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int T = p.T, BS = p.B;
    const int b0 = blockIdx.y * p.Bg;
    const int B = min(p.Bg, p.B - b0);
    unsigned* const bar = p.bar + 2 * blockIdx.y;

    // A = W_hh^T rows (my 16 units) x K = 4H gate rows, my K slice
    uint32_t afrag[KS][4];
    {
        const __nv_bfloat16* W = p.whh + (size_t)d * 4 * H * H;         // W[k][u]
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int k0 = (warp * KS + ks) * 16 + tig * 2;
            afrag[ks][0] = pack_bf16(W[(size_t)k0 * H + u0 + gid], W[(size_t)(k0 + 1) * H + u0 + gid]);
            afrag[ks][1] = pack_bf16(W[(size_t)k0 * H + u0 + gid + 8], W[(size_t)(k0 + 1) * H + u0 + gid + 8]);
            afrag[ks][2] = pack_bf16(W[(size_t)(k0 + 8) * H + u0 + gid], W[(size_t)(k0 + 9) * H + u0 + gid]);
            afrag[ks][3] = pack_bf16(W[(size_t)(k0 + 8) * H + u0 + gid + 8], W[(size_t)(k0 + 9) * H + u0 + gid + 8]);
        }
    }
    float dh_rec[PP], dc_carry[PP];
#pragma unroll
    for (int i = 0; i < PP; ++i) { dh_rec[i] = 0.f; dc_carry[i] = 0.f; }
    for (int i = tid; i < BP * GS; i += kLstmThreads) dgs[i] = __float2bfloat16(0.f);
    __syncthreads();

    // saved gates / cell states / incoming gradient of a step, fetched ONE STEP AHEAD (L2 latency off the chain)
    float nx[PP][7];
    auto load_step = [&](int step) {
        const int tt = d ? step : (T - 1 - step);
        const int tp = d ? tt + 1 : tt - 1;
#pragma unroll
        for (int i = 0; i < PP; ++i) {
            const int e = tid + kLstmThreads * i, u = e & 15, b = e >> 4;
#pragma unroll
            for (int k = 0; k < 7; ++k) nx[i][k] = 0.f;
            if (b < B && step < T) {
                const size_t gi = (((size_t)d * T + tt) * BS + b0 + b) * 4 * H + u0 + u;
                nx[i][0] = p.gates[gi]; nx[i][1] = p.gates[gi + H]; nx[i][2] = p.gates[gi + 2 * H]; nx[i][3] = p.gates[gi + 3 * H];
                nx[i][4] = p.cst[(((size_t)d * T + tt) * BS + b0 + b) * H + u0 + u];
                nx[i][5] = (tp >= 0 && tp < T) ? p.cst[(((size_t)d * T + tp) * BS + b0 + b) * H + u0 + u] : 0.f;
                nx[i][6] = __bfloat162float(p.dy[((size_t)(b0 + b) * T + tt) * 2 * H + (size_t)d * H + u0 + u]);
            }
        }
    };
    load_step(0);
    for (int s = 0; s < T; ++s) {
        const int t = d ? s : (T - 1 - s);           // reverse of the forward order of this direction
        float cur[PP][7];
#pragma unroll
        for (int i = 0; i < PP; ++i)
#pragma unroll
            for (int k = 0; k < 7; ++k) cur[i][k] = nx[i][k];
        load_step(s + 1);
#pragma unroll
        for (int i = 0; i < PP; ++i) {
            const int e = tid + kLstmThreads * i, u = e & 15, b = e >> 4;
            if (b < B) {
                const float ig = cur[i][0], fg = cur[i][1], gg = cur[i][2], og = cur[i][3];
                const float c = cur[i][4];
                const float cp = cur[i][5];
                const float dh = cur[i][6] + dh_rec[i];
                const float tc = tanh_fast(c);
                const float dc = dc_carry[i] + dh * og * (1.f - tc * tc);
                const float d_o = dh * tc * og * (1.f - og);
                const float d_i = dc * gg * ig * (1.f - ig);
                const float d_f = dc * cp * fg * (1.f - fg);
                const float d_g = dc * ig * (1.f - gg * gg);
                dc_carry[i] = dc * fg;
                __nv_bfloat16* go = p.dG + ((size_t)(b0 + b) * T + t) * 8 * H + (size_t)d * 4 * H + u0 + u;
                go[0] = __float2bfloat16(d_i); go[H] = __float2bfloat16(d_f);
                go[2 * H] = __float2bfloat16(d_g); go[3 * H] = __float2bfloat16(d_o);
            }
        }
        if (s == T - 1) break;                       // nothing precedes the first forward step
        if (p.tagged) {                                                    // all gate gradients of step t: poll the data
            poll_load_chunks(tid, B * (4 * H / 8), kLstmThreads,
                [&](int c) { const int b = c / (4 * H / 8), q = c % (4 * H / 8);
                             return reinterpret_cast<const uint4*>(p.dG + ((size_t)(b0 + b) * T + t) * 8 * H + (size_t)d * 4 * H) + q; },
                [&](int c) { const int b = c / (4 * H / 8), q = c % (4 * H / 8); return reinterpret_cast<uint4*>(dgs + b * GS + q * 8); });
        } else {
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                atomicAdd(bar + d, 1u);
                step_barrier_wait(bar + d, (unsigned)NC * (s + 1));
            }
            __syncthreads();
            for (int c = tid; c < B * (4 * H / 8); c += kLstmThreads) {   // all gate gradients of step t, every unit
                const int b = c / (4 * H / 8), q = c % (4 * H / 8);
                const uint4 v = __ldcg(reinterpret_cast<const uint4*>(p.dG + ((size_t)(b0 + b) * T + t) * 8 * H + (size_t)d * 4 * H) + q);
                *reinterpret_cast<uint4*>(dgs + b * GS + q * 8) = v;
            }
        }
        __syncthreads();
        float accp[4][NB][4];          // four independent accumulator chains (mma.sync latency), summed below
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4)
#pragma unroll
            for (int n = 0; n < NB; ++n)
#pragma unroll
                for (int j = 0; j < 4; ++j) accp[c4][n][j] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int k0 = (warp * KS + ks) * 16 + tig * 2;
#pragma unroll
            for (int n = 0; n < NB; ++n) {
                const __nv_bfloat16* gb = dgs + (n * 8 + gid) * GS + k0;
                mma_bf16_16816(accp[ks & 3][n], afrag[ks], *reinterpret_cast<const uint32_t*>(gb),
                               *reinterpret_cast<const uint32_t*>(gb + 8));
            }
        }
        float acc[NB][4];
#pragma unroll
        for (int n = 0; n < NB; ++n)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[n][j] = (accp[0][n][j] + accp[1][n][j]) + (accp[2][n][j] + accp[3][n][j]);
        {
            float* r = red + (size_t)warp * 16 * BP;
#pragma unroll
            for (int n = 0; n < NB; ++n) {
                const int bc = n * 8 + tig * 2;
                r[gid * BP + bc] = acc[n][0]; r[gid * BP + bc + 1] = acc[n][1];
                r[(gid + 8) * BP + bc] = acc[n][2]; r[(gid + 8) * BP + bc + 1] = acc[n][3];
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < PP; ++i) {
            const int e = tid + kLstmThreads * i, u = e & 15, b = e >> 4;
            float v = 0.f;
            if (b < B) {
#pragma unroll
                for (int w = 0; w < 8; ++w) v += red[((size_t)w * 16 + u) * BP + b];
            }
            dh_rec[i] = v;
        }
        __syncthreads();
    }
}

// (A thread-block-cluster variant — 16 CTAs per direction, h_t / gate gradients broadcast through DSMEM with
// st.shared::cluster and barrier.cluster arrive/wait instead of the L2 atomic + poll — was built and measured in round 1:
// 655 us vs 656 us forward and 1140 us vs 956 us backward per 2-layer call at B=8, T=150.  The release of
// barrier.cluster.arrive waits ~1300 cycles for the DSMEM stores, as long as the L2 round trips it replaces, so the
// cooperative kernels above are the only path; see git history for the variant.)

// ---- small helpers
struct LCastJob { const float* src; __nv_bfloat16* dst; long long n; };
struct LCastJobs { LCastJob j[8]; int count; };
// fp32 -> bf16 of up to 8 weight tensors in one launch: one flat index space over all of them, 4 elements per step
// (16-byte loads, 8-byte stores; element counts are multiples of 4 and the tensors 16-byte aligned, else scalar)
__global__ void lstm_cast_kernel(const LCastJobs jobs) {
    long long total4 = 0;
    for (int t = 0; t < jobs.count; ++t) total4 += (jobs.j[t].n + 3) >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        int t = 0;
        while (r >= ((jobs.j[t].n + 3) >> 2)) { r -= (jobs.j[t].n + 3) >> 2; ++t; }
        const LCastJob jb = jobs.j[t];
        const long long e = r << 2;
        if (e + 4 <= jb.n && ((reinterpret_cast<uintptr_t>(jb.src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(jb.dst) & 7) == 0)) {
            const float4 v = *reinterpret_cast<const float4*>(jb.src + e);
            __nv_bfloat162* d = reinterpret_cast<__nv_bfloat162*>(jb.dst + e);
            d[0] = __floats2bfloat162_rn(v.x, v.y);
            d[1] = __floats2bfloat162_rn(v.z, v.w);
        } else {
            for (long long k = e; k < jb.n && k < e + 4; ++k) jb.dst[k] = __float2bfloat16(jb.src[k]);
        }
    }
}
// out[0:4H] = a0 + b0 ; out[4H:8H] = a1 + b1
__global__ void lstm_bias_kernel(const float* a0, const float* b0, const float* a1, const float* b1, int n4h, float* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4h) { out[i] = a0[i] + b0[i]; out[n4h + i] = a1[i] + b1[i]; }
}
__global__ void lstm_copy2_kernel(const float* src, int n, float* d0, float* d1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float v = src[i]; d0[i] = v; d1[i] = v; }
}

struct LDims { int B, T, In, H, NB, G, Bg; long long BT; };
// Batch groups: the sequences of a call are independent, and one group of 2*H/16 CTAs (64 for H = 512) leaves more than
// half of the 148 SMs idle while its per-step cost grows with the batch (h / gate-gradient exchange, MMA N tiles).  A call
// with more than 8 sequences is therefore split into G groups that run side by side in ONE cooperative launch, each on
// its own CTAs with its own step counters: 16 sequences run as 2 x 8 at the step time of 8 (knob "lstm_groups": 0 auto,
// 1 = one group as in round 1).
static bool ldims(int B, int T, int In, int H, LDims* d) {
    if (B <= 0 || T <= 0 || In <= 0) return false;
    if (H != 512 && H != 256) return false;
    if (In % 8) return false;
    const int max_groups = 128 / (2 * H / kLstmUnits);          // co-resident CTAs: one per SM, 148 SMs
    int G = 1;
    const int knob = avctc_tuning_get("lstm_groups", 0);
    if (knob > 0) G = knob;
    else if (B > 8) G = max_groups < 2 ? 1 : (B > 32 && max_groups >= 4 ? 4 : 2);
    if (G > max_groups) G = max_groups;
    while ((B + G - 1) / G > 32 && G < max_groups) ++G;        // a group holds at most 32 sequences (knob or not)
    if (G > B) G = B;
    const int Bg = (B + G - 1) / G;
    if (Bg > 32) return false;
    d->B = B; d->T = T; d->In = In; d->H = H; d->G = (B + Bg - 1) / Bg; d->Bg = Bg;
    d->NB = Bg <= 8 ? 1 : (Bg <= 16 ? 2 : 4); d->BT = (long long)B * T;
    return true;
}
struct LCarver {
    char* base; size_t off;
    template <typename T> T* take(size_t n) {
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off = (off + n * sizeof(T) + 255) / 256 * 256;
        return p;
    }
};
struct LLayerSaved { __nv_bfloat16 *wih, *whh, *y, *hprev; float *gates, *cst; };
struct LSaved { LLayerSaved l[2]; size_t total; };
static LSaved lcarve_saved(void* base, const LDims& d) {
    LCarver c{reinterpret_cast<char*>(base), 0};
    LSaved s;
    const size_t H = d.H, BT = d.BT;
    for (int l = 0; l < 2; ++l) {
        const size_t In = l == 0 ? d.In : 2 * H;
        s.l[l].wih = c.take<__nv_bfloat16>(8 * H * In);
        s.l[l].whh = c.take<__nv_bfloat16>(2 * 4 * H * H);
        s.l[l].y = c.take<__nv_bfloat16>(BT * 2 * H);      // layer 1's y is the op's output; kept here too (h exchange)
        s.l[l].hprev = c.take<__nv_bfloat16>(BT * 2 * H);
        s.l[l].gates = c.take<float>(2 * BT * 4 * H);
        s.l[l].cst = c.take<float>(2 * BT * H);
    }
    s.total = c.off;
    return s;
}
struct LScratch { float* xproj; float* bias; unsigned* bar; __nv_bfloat16 *dG, *dmid, *dyb; float* dbias; size_t total; };
static LScratch lcarve_scratch(void* base, const LDims& d, bool backward) {
    LCarver c{reinterpret_cast<char*>(base), 0};
    LScratch s{};
    const size_t H = d.H, BT = d.BT;
    s.bar = c.take<unsigned>(64);
    if (!backward) {
        s.xproj = c.take<float>(BT * 8 * H);
        s.bias = c.take<float>(8 * H);
    } else {
        s.dG = c.take<__nv_bfloat16>(BT * 8 * H);
        s.dmid = c.take<__nv_bfloat16>(BT * 2 * H);
        s.dbias = c.take<float>(8 * H);
    }
    s.total = c.off;
    return s;
}

static avctc_gemm_operand lop(const void* ptr, long long rows, long long kdim, long long ld, bool mn = false) {
    avctc_gemm_operand o;
    o.ptr = ptr; o.rows = rows; o.kdim = kdim; o.zdim = 1; o.ld = ld; o.zstride = 0;
    o.k_outer = o.k_inner = o.r_outer = o.r_inner = o.z_outer = o.z_inner = 0;
    o.mn_major = mn ? 1 : 0;
    return o;
}

template <int H, int NB>
static int launch_fwd(const LstmFwdParams& p, int groups, cudaStream_t st) {
    const size_t smem = (size_t)8 * NB * (H + 8) * 2 + (size_t)2 * 4 * 16 * 8 * NB * 4;
    static bool cfg = false;
    if (!cfg && smem > 48 * 1024) {
        AVCTC_CUDA_RETURN(cudaFuncSetAttribute(lstm_fwd_kernel<H, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cfg = true;
    }
    void* args[] = {const_cast<LstmFwdParams*>(&p)};
    return (int)cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(lstm_fwd_kernel<H, NB>),
                                            dim3(2 * H / kLstmUnits, groups), dim3(kLstmThreads), args, smem, st);
}
template <int H, int NB>
static int launch_bwd(const LstmBwdParams& p, int groups, cudaStream_t st) {
    const size_t smem = (size_t)8 * NB * (4 * H + 8) * 2 + (size_t)8 * 16 * 8 * NB * 4;
    static bool cfg = false;
    if (!cfg && smem > 48 * 1024) {
        AVCTC_CUDA_RETURN(cudaFuncSetAttribute(lstm_bwd_kernel<H, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cfg = true;
    }
    void* args[] = {const_cast<LstmBwdParams*>(&p)};
    return (int)cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(lstm_bwd_kernel<H, NB>),
                                            dim3(2 * H / kLstmUnits, groups), dim3(kLstmThreads), args, smem, st);
}
static int dispatch_fwd(int H, int NB, int G, const LstmFwdParams& p, cudaStream_t st) {
    if (H == 512) { if (NB == 1) return launch_fwd<512, 1>(p, G, st); if (NB == 2) return launch_fwd<512, 2>(p, G, st); return launch_fwd<512, 4>(p, G, st); }
    if (NB == 1) return launch_fwd<256, 1>(p, G, st); if (NB == 2) return launch_fwd<256, 2>(p, G, st); return launch_fwd<256, 4>(p, G, st);
}
static int dispatch_bwd(int H, int NB, int G, const LstmBwdParams& p, cudaStream_t st) {
    if (H == 512) { if (NB == 1) return launch_bwd<512, 1>(p, G, st); if (NB == 2) return launch_bwd<512, 2>(p, G, st); return launch_bwd<512, 4>(p, G, st); }
    if (NB == 1) return launch_bwd<256, 1>(p, G, st); if (NB == 2) return launch_bwd<256, 2>(p, G, st); return launch_bwd<256, 4>(p, G, st);
}

#define LSTM_TRY(expr) do { const int rc_ = (expr); if (rc_) return rc_; } while (0)

}  // namespace avctc

using namespace avctc;

// params / grads: 16 pointers in nn.LSTM's flat order
//   l0: weight_ih, weight_hh, bias_ih, bias_hh, then the same four with suffix _reverse; then l1 likewise.
extern "C" size_t avctc_bilstm_workspace_bytes(int B, int T, int In, int H, int which) {
    LDims d;
    if (!ldims(B, T, In, H, &d)) return 0;
    if (which == 0) return lcarve_saved(nullptr, d).total;
    return lcarve_scratch(nullptr, d, which == 2).total;
}

extern "C" int avctc_bilstm_forward(const void* x_bf16, int B, int T, int In, int H, const float* const* params,
                                    void* y_bf16, void* saved, size_t saved_bytes, void* scratch, size_t scratch_bytes,
                                    int need_grad, void* stream) {
    LDims d;
    if (!ldims(B, T, In, H, &d)) return AVCTC_ERR_UNSUPPORTED;
    if (!x_bf16 || !params || !y_bf16 || !saved || !scratch) return AVCTC_ERR_BAD_ARG;
    for (int i = 0; i < 16; ++i) if (!params[i]) return AVCTC_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(saved) | reinterpret_cast<uintptr_t>(scratch)) & 255) return AVCTC_ERR_ALIGNMENT;
    LSaved s = lcarve_saved(saved, d);
    LScratch w = lcarve_scratch(scratch, d, false);
    if (saved_bytes < s.total || scratch_bytes < w.total) return AVCTC_ERR_WORKSPACE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t Hh = H;
    {   // bf16 weights of both layers: [8H,In] (fwd rows then reverse rows) and [2][4H][H]
        LCastJobs cj;
        cj.count = 8;
        for (int l = 0; l < 2; ++l) {
            const size_t Inl = l == 0 ? (size_t)In : 2 * Hh;
            const float* const* pp = params + 8 * l;
            cj.j[4 * l + 0] = {pp[0], s.l[l].wih, (long long)(4 * Hh * Inl)};
            cj.j[4 * l + 1] = {pp[4], s.l[l].wih + 4 * Hh * Inl, (long long)(4 * Hh * Inl)};
            cj.j[4 * l + 2] = {pp[1], s.l[l].whh, (long long)(4 * Hh * Hh)};
            cj.j[4 * l + 3] = {pp[5], s.l[l].whh + 4 * Hh * Hh, (long long)(4 * Hh * Hh)};
        }
        lstm_cast_kernel<<<1184, 256, 0, st>>>(cj);
        AVCTC_CUDA_RETURN(cudaGetLastError());
    }
    const __nv_bfloat16* xin = reinterpret_cast<const __nv_bfloat16*>(x_bf16);
    for (int l = 0; l < 2; ++l) {
        const int Inl = l == 0 ? In : 2 * H;
        const float* const* pp = params + 8 * l;
        lstm_bias_kernel<<<(4 * H + 255) / 256, 256, 0, st>>>(pp[2], pp[3], pp[6], pp[7], 4 * H, w.bias);
        AVCTC_CUDA_RETURN(cudaGetLastError());
        const avctc_gemm_operand A = lop(xin, d.BT, Inl, Inl), Bo = lop(s.l[l].wih, 8 * H, Inl, Inl);
        LSTM_TRY(avctc_gemm_launch(&A, &Bo, (int)d.BT, 8 * H, Inl, 1, 1, w.xproj, AVCTC_F32, 8 * H, 0, 0, w.bias, 1, 1.f, 0, 1,
                                   stream));
        AVCTC_CUDA_RETURN(cudaMemsetAsync(w.bar, 0, 64 * sizeof(unsigned), st));
        LstmFwdParams fp;
        fp.xproj = w.xproj; fp.whh = s.l[l].whh;
        fp.y = (l == 1) ? reinterpret_cast<__nv_bfloat16*>(y_bf16) : s.l[l].y;
        fp.hprev = need_grad ? s.l[l].hprev : nullptr;
        fp.gates = need_grad ? s.l[l].gates : nullptr;
        fp.cst = need_grad ? s.l[l].cst : nullptr;
        fp.bar = w.bar; fp.B = B; fp.T = T; fp.Bg = d.Bg;
        // measured (tools/exp_lstm_tag.py, B200): forward 766 -> 546 us at B=8, no gain at B=16 (more chunks to poll per
        // step); backward is slower with it at every batch size (every CTA polls all B x 4H gate gradients).
        // knob lstm_tag: 0 off, 1 forward when B <= 8 (default), 2 forward always, 3 forward and backward.
        const int tagk = avctc_tuning_get("lstm_tag", 1);
        fp.tagged = (tagk >= 2 || (tagk == 1 && d.Bg <= 8)) ? 1 : 0;
        if (fp.tagged) AVCTC_CUDA_RETURN(cudaMemsetAsync(fp.y, 0xFF, (size_t)d.BT * 2 * H * sizeof(__nv_bfloat16), st));
        LSTM_TRY(dispatch_fwd(H, d.NB, d.G, fp, st));
        xin = fp.y;
    }
    return AVCTC_OK;
}

extern "C" int avctc_bilstm_backward(const void* dy_bf16, const void* x_bf16, int B, int T, int In, int H,
                                     float* const* grads, void* dx_bf16, const void* saved, size_t saved_bytes,
                                     void* scratch, size_t scratch_bytes, void* stream) {
    LDims d;
    if (!ldims(B, T, In, H, &d)) return AVCTC_ERR_UNSUPPORTED;
    if (!dy_bf16 || !x_bf16 || !grads || !saved || !scratch) return AVCTC_ERR_BAD_ARG;
    for (int i = 0; i < 16; ++i) if (!grads[i]) return AVCTC_ERR_BAD_ARG;
    LSaved s = lcarve_saved(const_cast<void*>(saved), d);
    LScratch w = lcarve_scratch(scratch, d, true);
    if (saved_bytes < s.total || scratch_bytes < w.total) return AVCTC_ERR_WORKSPACE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const __nv_bfloat16* dy = reinterpret_cast<const __nv_bfloat16*>(dy_bf16);
    for (int l = 1; l >= 0; --l) {
        const int Inl = l == 0 ? In : 2 * H;
        const __nv_bfloat16* xl = (l == 0) ? reinterpret_cast<const __nv_bfloat16*>(x_bf16) : s.l[0].y;
        float* const* gp = grads + 8 * l;
        AVCTC_CUDA_RETURN(cudaMemsetAsync(w.bar, 0, 64 * sizeof(unsigned), st));
        LstmBwdParams bp;
        bp.dy = dy; bp.whh = s.l[l].whh; bp.gates = s.l[l].gates; bp.cst = s.l[l].cst; bp.dG = w.dG; bp.bar = w.bar;
        bp.B = B; bp.T = T; bp.Bg = d.Bg;
        bp.tagged = avctc_tuning_get("lstm_tag", 1) >= 3 ? 1 : 0;
        if (bp.tagged) AVCTC_CUDA_RETURN(cudaMemsetAsync(w.dG, 0xFF, (size_t)d.BT * 8 * H * sizeof(__nv_bfloat16), st));
        LSTM_TRY(dispatch_bwd(H, d.NB, d.G, bp, st));
        for (int dir = 0; dir < 2; ++dir) {
            // dW_ih[dir] [4H,In] = dG[:, dir*4H:+4H]^T . x ; dW_hh[dir] [4H,H] = dG_dir^T . hprev[:, dir*H:+H]
            const avctc_gemm_operand A = lop(w.dG + (size_t)dir * 4 * H, 4 * H, d.BT, 8 * H, true);
            const avctc_gemm_operand Bx = lop(xl, Inl, d.BT, Inl, true);
            LSTM_TRY(avctc_gemm_launch(&A, &Bx, 4 * H, Inl, (int)d.BT, 1, 1, gp[4 * dir + 0], AVCTC_F32, Inl, 0, 0, nullptr, 0,
                                       1.f, 0, 2, stream));
            const avctc_gemm_operand Bh = lop(s.l[l].hprev + (size_t)dir * H, H, d.BT, 2 * H, true);
            LSTM_TRY(avctc_gemm_launch(&A, &Bh, 4 * H, H, (int)d.BT, 1, 1, gp[4 * dir + 1], AVCTC_F32, H, 0, 0, nullptr, 0, 1.f,
                                       0, 2, stream));
        }
        LSTM_TRY(avctc_colsum(w.dG, AVCTC_BF16, d.BT, 8 * H, 8 * H, w.dbias, 0, stream));
        lstm_copy2_kernel<<<(4 * H + 255) / 256, 256, 0, st>>>(w.dbias, 4 * H, gp[2], gp[3]);
        lstm_copy2_kernel<<<(4 * H + 255) / 256, 256, 0, st>>>(w.dbias + 4 * H, 4 * H, gp[6], gp[7]);
        AVCTC_CUDA_RETURN(cudaGetLastError());
        // dX [BT,In] = dG [BT,8H] . W_ih_cat [8H,In]
        __nv_bfloat16* dxo = (l == 0) ? reinterpret_cast<__nv_bfloat16*>(dx_bf16) : w.dmid;
        if (dxo) {
            const avctc_gemm_operand A = lop(w.dG, d.BT, 8 * H, 8 * H), Bw = lop(s.l[l].wih, Inl, 8 * H, Inl, true);
            LSTM_TRY(avctc_gemm_launch(&A, &Bw, (int)d.BT, Inl, 8 * H, 1, 1, dxo, AVCTC_BF16, Inl, 0, 0, nullptr, 0, 1.f, 0, 1,
                                       stream));
        }
        dy = w.dmid;
    }
    return AVCTC_OK;
}
