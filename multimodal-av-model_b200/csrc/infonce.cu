// infonce.cu — fused mask-driven InfoNCE (similarity + softmax + means) for sm_100a.
//
// Replaces the body of contrastive_loss_with_mask (/root/reference/contrastive.py:13-44) after the optional
// projection: drop mask==3 rows, F.normalize, index sets strong(2)/weak(1)/neg(0), and for the pairs
// (weak,strong; weight 1.0) and (weak,neg; weight 0.3):  mean_ij( -log_softmax_row(A.S^T / 0.07) ).
// The reference materialises boolean-indexed copies and both [Nw,Ns] similarity matrices (nonzero() host
// syncs, 2 GEMMs, 2 softmaxes).  Here nothing is materialised and nothing syncs:
//   index kernel   ordered compaction of the three row sets (device counts)
//   normalize      z = y / max(|y|, 1e-12), fp32
//   pair forward   one warp per anchor streams the other set, lane-per-column dot products, online
//                  log-sum-exp; per-anchor (lse, sum sim) -> deterministic final reduction
//   pair backward  same streaming loop, recomputes sim, p = exp(sim - lse); run once per side so every
//                  launch owns the rows it writes (no atomics, bit-reproducible)
//   normalize bwd  dy = (dz - z (z.dz)) / max(|y|, eps)
// mean over ALL entries = mean_i(lse_i) - mean_ij(sim_ij); there are no "diagonal positives" (SURVEY.md a10).
#include "common.cuh"

namespace avctc {

constexpr float kNceEps = 1e-12f;

struct NceWs {
    int* idx[3];      // row lists: [0] mask==0 (neg), [1] mask==1 (weak anchors), [2] mask==2 (strong)
    int* cnt;         // [3]
    float* z;         // [N][P]
    float* invn;      // [N] 1/max(|y|,eps)
    float* nrm;       // [N] |y|
    float* lse[2];    // per anchor slot, pair 0 = (weak,strong), pair 1 = (weak,neg)
    float* ssum[2];
    float* dz;        // [N][P]
    size_t total;
};

static NceWs carve(void* base, int N, int P) {
    NceWs w;
    size_t o = 0;
    char* b = reinterpret_cast<char*>(base);
    auto take = [&](size_t bytes) { char* p = b ? b + o : nullptr; o = (o + bytes + 255) / 256 * 256; return p; };
    for (int i = 0; i < 3; ++i) w.idx[i] = reinterpret_cast<int*>(take(sizeof(int) * (size_t)N));
    w.cnt = reinterpret_cast<int*>(take(sizeof(int) * 4));
    w.z = reinterpret_cast<float*>(take(sizeof(float) * (size_t)N * P));
    w.invn = reinterpret_cast<float*>(take(sizeof(float) * (size_t)N));
    w.nrm = reinterpret_cast<float*>(take(sizeof(float) * (size_t)N));
    for (int i = 0; i < 2; ++i) w.lse[i] = reinterpret_cast<float*>(take(sizeof(float) * (size_t)N));
    for (int i = 0; i < 2; ++i) w.ssum[i] = reinterpret_cast<float*>(take(sizeof(float) * (size_t)N));
    w.dz = reinterpret_cast<float*>(take(sizeof(float) * (size_t)N * P));
    w.total = o;
    return w;
}

// single CTA: ordered (ascending row) compaction of the three sets
__global__ void nce_index_kernel(const int64_t* __restrict__ mask, int N, int* i0, int* i1, int* i2, int* cnt) {
    __shared__ int wtot[3][32];
    __shared__ int base[3];
    if (threadIdx.x < 3) base[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int* outs[3] = {i0, i1, i2};
    for (int r0 = 0; r0 < N; r0 += blockDim.x) {
        const int r = r0 + threadIdx.x;
        const long long m = (r < N) ? mask[r] : 3;
        unsigned bal[3];
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            bal[s] = __ballot_sync(kFullMask, m == s);
            if (lane == 0) wtot[s][warp] = __popc(bal[s]);
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            if (m == s) {
                int pos = base[s] + __popc(bal[s] & ((1u << lane) - 1));
                for (int w = 0; w < warp; ++w) pos += wtot[s][w];
                outs[s][pos] = r;
            }
        }
        __syncthreads();
        if (threadIdx.x < 3) {
            int t = 0;
            for (int w = 0; w < nw; ++w) t += wtot[threadIdx.x][w];
            base[threadIdx.x] += t;
        }
        __syncthreads();
    }
    if (threadIdx.x < 3) cnt[threadIdx.x] = base[threadIdx.x];
}

template <typename TIn>
__global__ void nce_normalize_kernel(const TIn* __restrict__ y, long long ld, int N, int P, float* __restrict__ z,
                                     float* __restrict__ invn, float* __restrict__ nrm) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= N) return;
    const int lane = threadIdx.x & 31;
    const TIn* yr = y + (long long)row * ld;
    float ss = 0.f;
    for (int d = lane; d < P; d += 32) { const float v = to_float(yr[d]); ss += v * v; }
    ss = warp_sum(ss);
    const float n = sqrtf(ss);
    const float inv = 1.f / fmaxf(n, kNceEps);
    for (int d = lane; d < P; d += 32) z[(size_t)row * P + d] = to_float(yr[d]) * inv;
    if (lane == 0) { invn[row] = inv; nrm[row] = n; }
}

// one warp per anchor slot i; lane-per-column dot products over the other set
__global__ void nce_pair_fwd_kernel(const float* __restrict__ z, int P, const int* __restrict__ idxA,
                                    const int* __restrict__ idxO, const int* __restrict__ cnt, int setA, int setO,
                                    float inv_tau, float* __restrict__ lse, float* __restrict__ ssum) {
    extern __shared__ float za_s[];              // [warps][P]
    const int nA = cnt[setA], nO = cnt[setO];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * (blockDim.x >> 5) + warp;
    if (i >= nA || nO == 0) return;
    float* za = za_s + warp * P;
    const float* zi = z + (size_t)idxA[i] * P;
    for (int d = lane; d < P; d += 32) za[d] = zi[d];
    __syncwarp();
    float m = AVCTC_NEG_INF, s = 0.f, tot = 0.f;
    for (int j0 = 0; j0 < nO; j0 += 32) {
        const int j = j0 + lane;
        if (j < nO) {
            const float* zj = z + (size_t)idxO[j] * P;
            float dot = 0.f;
            for (int d = 0; d < P; ++d) dot = fmaf(za[d], zj[d], dot);
            const float sim = dot * inv_tau;
            tot += sim;
            const float mn = fmaxf(m, sim);
            s = s * __expf(m - mn) + __expf(sim - mn);
            m = mn;
        }
    }
    const float mw = warp_max(m);
    const float sw = warp_sum((m == AVCTC_NEG_INF) ? 0.f : s * __expf(m - mw));
    const float tw = warp_sum(tot);
    if (lane == 0) { lse[i] = mw + logf(sw); ssum[i] = tw; }
}

// loss[0] = w0 * pair0 + w1 * pair1, pair = mean_i(lse_i) - sum_i(ssum_i)/(nA*nO); single CTA, fixed order
__global__ void nce_finalize_kernel(const int* __restrict__ cnt, const float* lse0, const float* ss0, const float* lse1,
                                    const float* ss1, float w0, float w1, float* __restrict__ loss) {
    __shared__ double part[2][32];
    const int nA = cnt[1];
    const int nO[2] = {cnt[2], cnt[0]};
    const float* L[2] = {lse0, lse1};
    const float* S[2] = {ss0, ss1};
    double acc[2] = {0.0, 0.0};
    for (int pr = 0; pr < 2; ++pr) {
        if (nA == 0 || nO[pr] == 0) continue;
        const double invA = 1.0 / nA, invAO = 1.0 / ((double)nA * nO[pr]);
        for (int i = threadIdx.x; i < nA; i += blockDim.x) acc[pr] += (double)L[pr][i] * invA - (double)S[pr][i] * invAO;
    }
    for (int pr = 0; pr < 2; ++pr) {
        double v = acc[pr];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
        if ((threadIdx.x & 31) == 0) part[pr][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0.0, t1 = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t0 += part[0][w]; t1 += part[1][w]; }
        loss[0] = (float)((double)w0 * t0 + (double)w1 * t1);
    }
}

// gradient w.r.t. z for one side of one pair.  rows = the set this launch writes, cols = the set it sums
// over; anchor_is_row says whose lse normalises p_ij.  One warp per row, lane-per-column coefficients,
// lane-per-dimension accumulation.
__global__ void nce_pair_bwd_kernel(const float* __restrict__ z, int P, const int* __restrict__ idxR,
                                    const int* __restrict__ idxC, const int* __restrict__ cnt, int setR, int setC,
                                    int anchor_is_row, const float* __restrict__ lse, float inv_tau, float weight,
                                    const float* __restrict__ gout, float* __restrict__ dz, int accumulate) {
    extern __shared__ float zr_s[];
    const int nR = cnt[setR], nC = cnt[setC];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x * (blockDim.x >> 5) + warp;
    if (r >= nR) return;
    const int row = idxR[r];
    float* drow = dz + (size_t)row * P;
    if (nC == 0) {
        if (!accumulate) for (int d = lane; d < P; d += 32) drow[d] = 0.f;
        return;
    }
    float* zr = zr_s + warp * P;
    const float* zg = z + (size_t)row * P;
    for (int d = lane; d < P; d += 32) zr[d] = zg[d];
    __syncwarp();
    const int nA = anchor_is_row ? nR : nC, nO = anchor_is_row ? nC : nR;
    const float g = gout[0] * weight * inv_tau;
    const float invA = 1.f / (float)nA, invAO = 1.f / ((float)nA * (float)nO);
    const float lse_r = anchor_is_row ? lse[r] : 0.f;
    // lane owns dims d = lane + 32*k
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    const int nk = (P + 31) / 32;           // host guarantees P <= 256
    for (int c0 = 0; c0 < nC; c0 += 32) {
        const int c = c0 + lane;
        float coef = 0.f;
        int crow = 0;
        if (c < nC) {
            crow = idxC[c];
            const float* zc = z + (size_t)crow * P;
            float dot = 0.f;
            for (int d = 0; d < P; ++d) dot = fmaf(zr[d], zc[d], dot);
            const float sim = dot * inv_tau;
            const float l = anchor_is_row ? lse_r : lse[c];
            coef = g * (__expf(sim - l) * invA - invAO);
        }
        const int lim = min(32, nC - c0);
        for (int t = 0; t < lim; ++t) {
            const float cf = __shfl_sync(kFullMask, coef, t);
            const int cr = __shfl_sync(kFullMask, crow, t);
            const float* zc = z + (size_t)cr * P;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < nk) { const int d = lane + 32 * k; if (d < P) acc[k] = fmaf(cf, zc[d], acc[k]); }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k < nk) { const int d = lane + 32 * k; if (d < P) drow[d] = accumulate ? drow[d] + acc[k] : acc[k]; }
}

// rows that belong to no set (mask==3, or sets that never pair) must end with dz = 0
__global__ void nce_zero_rows_kernel(const int64_t* __restrict__ mask, const int* __restrict__ cnt, int N, int P,
                                     float* __restrict__ dz) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= N) return;
    const long long m = mask[row];
    const int nW = cnt[1], nS = cnt[2], nN = cnt[0];
    bool used = false;
    if (m == 1) used = (nS > 0 || nN > 0);
    else if (m == 2) used = (nW > 0);      // written by the (weak,strong) other-side launch
    else if (m == 0) used = (nW > 0);
    if (!used) for (int d = threadIdx.x & 31; d < P; d += 32) dz[(size_t)row * P + d] = 0.f;
}

template <typename TOut>
__global__ void nce_normalize_bwd_kernel(const float* __restrict__ z, const float* __restrict__ dz,
                                         const float* __restrict__ invn, const float* __restrict__ nrm, int N, int P,
                                         TOut* __restrict__ dy, long long ld) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= N) return;
    const int lane = threadIdx.x & 31;
    const float* zr = z + (size_t)row * P;
    const float* dr = dz + (size_t)row * P;
    float dot = 0.f;
    for (int d = lane; d < P; d += 32) dot += zr[d] * dr[d];
    dot = warp_sum(dot);
    const float inv = invn[row];
    const bool clamped = !(nrm[row] > kNceEps);   // F.normalize: y / clamp_min(|y|, eps); clamp has zero slope
    for (int d = lane; d < P; d += 32) {
        const float v = clamped ? dr[d] * inv : (dr[d] - zr[d] * dot) * inv;
        if constexpr (sizeof(TOut) == 4) dy[(long long)row * ld + d] = v;
        else dy[(long long)row * ld + d] = __float2bfloat16(v);
    }
}

}  // namespace avctc

using namespace avctc;

extern "C" size_t avctc_infonce_workspace_bytes(int N, int P) {
    if (N <= 0 || P <= 0) return 0;
    return carve(nullptr, N, P).total;
}

extern "C" int avctc_infonce_forward(const void* y, int dtype, long long ld, const int64_t* flat_mask, int N, int P,
                                     float temperature, float w_pos, float w_neg, float* loss, void* workspace,
                                     size_t workspace_bytes, void* stream) {
    if (!y || !flat_mask || !loss || !workspace || N <= 0 || P <= 0 || temperature <= 0.f) return AVCTC_ERR_BAD_ARG;
    if (P > 256) return AVCTC_ERR_UNSUPPORTED;
    if (dtype != AVCTC_F32 && dtype != AVCTC_BF16) return AVCTC_ERR_BAD_ARG;
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return AVCTC_ERR_ALIGNMENT;
    NceWs w = carve(workspace, N, P);
    if (workspace_bytes < w.total) return AVCTC_ERR_WORKSPACE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    nce_index_kernel<<<1, 1024, 0, st>>>(flat_mask, N, w.idx[0], w.idx[1], w.idx[2], w.cnt);
    const int wpb = 8;
    const unsigned rgrid = (N + wpb - 1) / wpb;
    if (dtype == AVCTC_F32)
        nce_normalize_kernel<float><<<rgrid, wpb * 32, 0, st>>>(reinterpret_cast<const float*>(y), ld, N, P, w.z, w.invn, w.nrm);
    else
        nce_normalize_kernel<__nv_bfloat16><<<rgrid, wpb * 32, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(y), ld, N, P, w.z, w.invn, w.nrm);
    const float inv_tau = 1.f / temperature;
    const size_t smem = (size_t)wpb * P * sizeof(float);
    nce_pair_fwd_kernel<<<rgrid, wpb * 32, smem, st>>>(w.z, P, w.idx[1], w.idx[2], w.cnt, 1, 2, inv_tau, w.lse[0], w.ssum[0]);
    nce_pair_fwd_kernel<<<rgrid, wpb * 32, smem, st>>>(w.z, P, w.idx[1], w.idx[0], w.cnt, 1, 0, inv_tau, w.lse[1], w.ssum[1]);
    nce_finalize_kernel<<<1, 256, 0, st>>>(w.cnt, w.lse[0], w.ssum[0], w.lse[1], w.ssum[1], w_pos, w_neg, loss);
    return (int)cudaGetLastError();
}

extern "C" int avctc_infonce_backward(const int64_t* flat_mask, int N, int P, float temperature, float w_pos,
                                      float w_neg, const float* grad_out, void* dy, int dtype, long long ld,
                                      void* workspace, size_t workspace_bytes, void* stream) {
    if (!flat_mask || !grad_out || !dy || !workspace || N <= 0 || P <= 0 || temperature <= 0.f) return AVCTC_ERR_BAD_ARG;
    if (P > 256) return AVCTC_ERR_UNSUPPORTED;
    if (dtype != AVCTC_F32 && dtype != AVCTC_BF16) return AVCTC_ERR_BAD_ARG;
    NceWs w = carve(workspace, N, P);
    if (workspace_bytes < w.total) return AVCTC_ERR_WORKSPACE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int wpb = 8;
    const unsigned rgrid = (N + wpb - 1) / wpb;
    const float inv_tau = 1.f / temperature;
    const size_t smem = (size_t)wpb * P * sizeof(float);
    // anchors (weak rows): pair 0 writes, pair 1 accumulates; others: each set written by exactly one launch
    nce_pair_bwd_kernel<<<rgrid, wpb * 32, smem, st>>>(w.z, P, w.idx[1], w.idx[2], w.cnt, 1, 2, 1, w.lse[0], inv_tau, w_pos, grad_out, w.dz, 0);
    nce_pair_bwd_kernel<<<rgrid, wpb * 32, smem, st>>>(w.z, P, w.idx[1], w.idx[0], w.cnt, 1, 0, 1, w.lse[1], inv_tau, w_neg, grad_out, w.dz, 1);
    nce_pair_bwd_kernel<<<rgrid, wpb * 32, smem, st>>>(w.z, P, w.idx[2], w.idx[1], w.cnt, 2, 1, 0, w.lse[0], inv_tau, w_pos, grad_out, w.dz, 0);
    nce_pair_bwd_kernel<<<rgrid, wpb * 32, smem, st>>>(w.z, P, w.idx[0], w.idx[1], w.cnt, 0, 1, 0, w.lse[1], inv_tau, w_neg, grad_out, w.dz, 0);
    nce_zero_rows_kernel<<<rgrid, wpb * 32, 0, st>>>(flat_mask, w.cnt, N, P, w.dz);
    if (dtype == AVCTC_F32)
        nce_normalize_bwd_kernel<float><<<rgrid, wpb * 32, 0, st>>>(w.z, w.dz, w.invn, w.nrm, N, P, reinterpret_cast<float*>(dy), ld);
    else
        nce_normalize_bwd_kernel<__nv_bfloat16><<<rgrid, wpb * 32, 0, st>>>(w.z, w.dz, w.invn, w.nrm, N, P, reinterpret_cast<__nv_bfloat16*>(dy), ld);
    return (int)cudaGetLastError();
}
