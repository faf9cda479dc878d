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
constexpr int kTile = 64;
constexpr int kNceThreads = 256;
constexpr int kSlots = 8;      // partial-gradient slots (fixed-order reduction keeps the result bit-reproducible)

struct NceWs {
    int* idx[3];      // row lists: [0] mask==0 (neg), [1] mask==1 (weak anchors), [2] mask==2 (strong)
    int* cnt;         // [3]
    float* z;         // [N][P]
    float* invn;      // [N] 1/max(|y|,eps)
    float* nrm;       // [N] |y|
    float* lse[2];    // per anchor slot, pair 0 = (weak,strong), pair 1 = (weak,neg)
    float* ssum[2];
    float* dzp;       // [kSlots][N][P] partial gradients w.r.t. z
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
    w.dzp = reinterpret_cast<float*>(take(sizeof(float) * (size_t)kSlots * N * P));
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

// ---- tiled pair kernels ------------------------------------------------------------------------------
// A CTA owns a 64-row tile of one row set ("rows", kept in shared memory) and streams 64-row tiles of the other
// set ("cols").  256 threads as 16 x 16: thread (ty,tx) holds the 4x4 similarities of rows ty+16i, cols tx+16j
// (interleaved so that the float4 shared reads of 16 different cols hit 16 different bank groups).
__device__ __forceinline__ void nce_load_tile(float* dst, int PS, int P4, const float* __restrict__ z, int P,
                                              const int* __restrict__ idx, int first, int count) {
    // rows first..first+63 of the compacted list -> dst[r][0..P4), zero rows / zero pad columns beyond
    const int per_row = P4 >> 2;
    for (int e = threadIdx.x; e < kTile * per_row; e += kNceThreads) {
        const int r = e / per_row, q = (e - r * per_row) << 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (first + r < count) {
            const float* src = z + (size_t)idx[first + r] * P + q;
            if (((P & 3) == 0)) v = *reinterpret_cast<const float4*>(src);
            else {
                v.x = src[0];
                if (q + 1 < P) v.y = src[1];
                if (q + 2 < P) v.z = src[2];
                if (q + 3 < P) v.w = src[3];
            }
        }
        *reinterpret_cast<float4*>(dst + r * PS + q) = v;
    }
}

__device__ __forceinline__ void nce_sim_tile(const float* __restrict__ zR, const float* __restrict__ zC, int PS, int P4,
                                             int ty, int tx, float (&sim)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) sim[i][j] = 0.f;
    for (int d = 0; d < P4; d += 4) {
        float4 a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(zR + (ty + 16 * i) * PS + d);
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(zC + (tx + 16 * j) * PS + d);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sim[i][j] = fmaf(a[i].x, b[j].x, sim[i][j]);
                sim[i][j] = fmaf(a[i].y, b[j].y, sim[i][j]);
                sim[i][j] = fmaf(a[i].z, b[j].z, sim[i][j]);
                sim[i][j] = fmaf(a[i].w, b[j].w, sim[i][j]);
            }
    }
}

// forward: per anchor i of the weak set, lse_i = log sum_j exp(sim_ij) and ssum_i = sum_j sim_ij over the other set.
// grid (anchor tiles, 2 pairs).
__global__ void __launch_bounds__(kNceThreads)
nce_pair_fwd_kernel(const float* __restrict__ z, int P, const int* __restrict__ idxW, const int* __restrict__ idxS,
                    const int* __restrict__ idxN, const int* __restrict__ cnt, float inv_tau,
                    float* __restrict__ lse0, float* __restrict__ ss0, float* __restrict__ lse1, float* __restrict__ ss1) {
    extern __shared__ __align__(16) float nce_smem[];
    const int pair = blockIdx.y;
    const int nA = cnt[1], nO = pair ? cnt[0] : cnt[2];
    const int first = blockIdx.x * kTile;
    if (first >= nA || nO == 0) return;
    const int* idxO = pair ? idxN : idxS;
    const int P4 = (P + 3) & ~3, PS = P4 + 4;
    float* zR = nce_smem;
    float* zC = zR + kTile * PS;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    nce_load_tile(zR, PS, P4, z, P, idxW, first, nA);
    float m[4], s[4], tot[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { m[i] = AVCTC_NEG_INF; s[i] = 0.f; tot[i] = 0.f; }
    for (int c0 = 0; c0 < nO; c0 += kTile) {
        __syncthreads();
        nce_load_tile(zC, PS, P4, z, P, idxO, c0, nO);
        __syncthreads();
        float sim[4][4];
        nce_sim_tile(zR, zC, PS, P4, ty, tx, sim);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float tm = AVCTC_NEG_INF, tt = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool ok = (c0 + tx + 16 * j) < nO;
                sim[i][j] = ok ? sim[i][j] * inv_tau : AVCTC_NEG_INF;
                tm = fmaxf(tm, sim[i][j]);
                tt += ok ? sim[i][j] : 0.f;
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) tm = fmaxf(tm, __shfl_xor_sync(kFullMask, tm, o));   // 16 lanes share a row
            const float mn = fmaxf(m[i], tm);          // finite: every tile has at least one valid column
            float ts = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) ts += __expf(sim[i][j] - mn);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                ts += __shfl_xor_sync(kFullMask, ts, o);
                tt += __shfl_xor_sync(kFullMask, tt, o);
            }
            s[i] = s[i] * __expf(m[i] - mn) + ts;
            m[i] = mn;
            tot[i] += tt;
        }
    }
    if (tx == 0) {
        float* lse = pair ? lse1 : lse0;
        float* ss = pair ? ss1 : ss0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = first + ty + 16 * i;
            if (r < nA) { lse[r] = m[i] + logf(s[i]); ss[r] = tot[i]; }
        }
    }
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

// backward w.r.t. z.  grid (row tiles, kSlots, 3 sets).  A CTA owns 64 rows of `set` and a slot:
//   set 1 (weak anchors): slots [0,4) stream the strong set (pair 0), slots [4,8) the neg set (pair 1)
//   set 2 / 0 (strong / neg): every slot streams a strided share of the weak anchors
// and writes  dzp[slot][row][:] = sum_cols coef(row,col) * z[col],  coef = g (p_ij/nA - 1/(nA nO)),
// p_ij = exp(sim_ij - lse_anchor).  Every (slot,row) is written by exactly one CTA -> no atomics.
template <int NE>    // second product: thread owns dims tx*4 + 64*e .. +3, e < NE  (P <= 64*NE)
__global__ void __launch_bounds__(kNceThreads)
nce_pair_bwd_kernel(const float* __restrict__ z, int P, int N, const int* __restrict__ idxN, const int* __restrict__ idxW,
                    const int* __restrict__ idxS, const int* __restrict__ cnt, const float* __restrict__ lse0,
                    const float* __restrict__ lse1, float inv_tau, float w_pos, float w_neg,
                    const float* __restrict__ gout, float* __restrict__ dzp) {
    extern __shared__ __align__(16) float nce_smem[];
    const int set = blockIdx.z, slot = blockIdx.y;
    const int nR = cnt[set];
    const int first = blockIdx.x * kTile;
    if (first >= nR) return;
    int setC, pair, sub, nsub;
    if (set == 1) { pair = slot / (kSlots / 2); sub = slot % (kSlots / 2); nsub = kSlots / 2; setC = pair ? 0 : 2; }
    else { pair = (set == 2) ? 0 : 1; sub = slot; nsub = kSlots; setC = 1; }
    const bool anchor_is_row = (set == 1);
    const int nC = cnt[setC];
    const int* idxR = (set == 0) ? idxN : (set == 1) ? idxW : idxS;
    const int* idxC = (setC == 0) ? idxN : (setC == 1) ? idxW : idxS;
    const float* lse = pair ? lse1 : lse0;
    const int P4 = (P + 3) & ~3, PS = P4 + 4;
    float* zR = nce_smem;
    float* zC = zR + kTile * PS;
    float* coef = zC + kTile * PS;            // [64][65]
    float* lse_c = coef + kTile * (kTile + 1);  // [64]
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    nce_load_tile(zR, PS, P4, z, P, idxR, first, nR);
    const int nA = anchor_is_row ? nR : nC, nO = anchor_is_row ? nC : nR;
    const float g = gout[0] * (pair ? w_neg : w_pos) * inv_tau;
    const float invA = nA > 0 ? 1.f / (float)nA : 0.f;
    const float invAO = (nA > 0 && nO > 0) ? 1.f / ((float)nA * (float)nO) : 0.f;
    float lse_r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = first + ty + 16 * i;
        lse_r[i] = (anchor_is_row && r < nR) ? lse[r] : 0.f;
    }
    float acc[4][NE][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < NE; ++e)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][e][k] = 0.f;

    for (int c0 = sub * kTile; c0 < nC; c0 += nsub * kTile) {
        __syncthreads();
        nce_load_tile(zC, PS, P4, z, P, idxC, c0, nC);
        if (!anchor_is_row && threadIdx.x < kTile) lse_c[threadIdx.x] = (c0 + threadIdx.x < nC) ? lse[c0 + threadIdx.x] : 0.f;
        __syncthreads();
        float sim[4][4];
        nce_sim_tile(zR, zC, PS, P4, ty, tx, sim);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int rr = ty + 16 * i, cc = tx + 16 * j;
                const bool ok = (first + rr < nR) && (c0 + cc < nC);
                const float l = anchor_is_row ? lse_r[i] : lse_c[cc];
                coef[rr * (kTile + 1) + cc] = ok ? g * (__expf(sim[i][j] * inv_tau - l) * invA - invAO) : 0.f;
            }
        __syncthreads();
#pragma unroll 4
        for (int c = 0; c < kTile; ++c) {
            float cf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) cf[i] = coef[(ty + 16 * i) * (kTile + 1) + c];
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                const int d = tx * 4 + 64 * e;
                if (d < P4) {
                    const float4 v = *reinterpret_cast<const float4*>(zC + c * PS + d);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[i][e][0] = fmaf(cf[i], v.x, acc[i][e][0]);
                        acc[i][e][1] = fmaf(cf[i], v.y, acc[i][e][1]);
                        acc[i][e][2] = fmaf(cf[i], v.z, acc[i][e][2]);
                        acc[i][e][3] = fmaf(cf[i], v.w, acc[i][e][3]);
                    }
                }
            }
        }
    }
    float* out = dzp + (size_t)slot * N * P;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = first + ty + 16 * i;
        if (r >= nR) continue;
        float* drow = out + (size_t)idxR[r] * P;
#pragma unroll
        for (int e = 0; e < NE; ++e)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int d = tx * 4 + 64 * e + k;
                if (d < P) drow[d] = acc[i][e][k];
            }
    }
}

// dz[row] = sum over the kSlots partials (fixed order; rows with mask 3 get 0), then the F.normalize backward.
template <typename TOut>
__global__ void nce_normalize_bwd_kernel(const float* __restrict__ z, const float* __restrict__ dzp,
                                         const int64_t* __restrict__ mask, const float* __restrict__ invn,
                                         const float* __restrict__ nrm, int N, int P, TOut* __restrict__ dy, long long ld) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= N) return;
    const int lane = threadIdx.x & 31;
    const float* zr = z + (size_t)row * P;
    const long long mk = mask[row];
    const bool used = (mk >= 0 && mk <= 2);
    float dr[8];                                   // P <= 256
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int d = lane + 32 * k;
        float v = 0.f;
        if (used && d < P) {
#pragma unroll
            for (int s = 0; s < kSlots; ++s) v += dzp[((size_t)s * N + row) * P + d];
            dot += zr[d] * v;
        }
        dr[k] = v;
    }
    dot = warp_sum(dot);
    const float inv = invn[row];
    const bool clamped = !(nrm[row] > kNceEps);   // F.normalize: y / clamp_min(|y|, eps); clamp has zero slope
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int d = lane + 32 * k;
        if (d >= P) continue;
        const float v = clamped ? dr[k] * inv : (dr[k] - zr[d] * dot) * inv;
        if constexpr (sizeof(TOut) == 4) dy[(long long)row * ld + d] = v;
        else dy[(long long)row * ld + d] = __float2bfloat16(v);
    }
}

static size_t nce_tile_smem(int P) {
    const int P4 = (P + 3) & ~3, PS = P4 + 4;
    return sizeof(float) * ((size_t)2 * kTile * PS + (size_t)kTile * (kTile + 1) + kTile);
}
static int nce_configure_smem() {
    static bool done = false;
    if (done) return 0;
    const int maxb = (int)nce_tile_smem(256);
    AVCTC_CUDA_RETURN(cudaFuncSetAttribute(nce_pair_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxb));
    AVCTC_CUDA_RETURN(cudaFuncSetAttribute(nce_pair_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxb));
    AVCTC_CUDA_RETURN(cudaFuncSetAttribute(nce_pair_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxb));
    AVCTC_CUDA_RETURN(cudaFuncSetAttribute(nce_pair_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxb));
    AVCTC_CUDA_RETURN(cudaFuncSetAttribute(nce_pair_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxb));
    done = true;
    return 0;
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
    const size_t smem = nce_tile_smem(P);
    { const int rc_ = nce_configure_smem(); if (rc_) return rc_; }
    dim3 fgrid((N + kTile - 1) / kTile, 2);
    nce_pair_fwd_kernel<<<fgrid, kNceThreads, smem, st>>>(w.z, P, w.idx[1], w.idx[2], w.idx[0], w.cnt, inv_tau,
                                                         w.lse[0], w.ssum[0], w.lse[1], w.ssum[1]);
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
    const size_t smem = nce_tile_smem(P);
    { const int rc_ = nce_configure_smem(); if (rc_) return rc_; }
    dim3 bgrid((N + kTile - 1) / kTile, kSlots, 3);
#define AVCTC_NCE_BWD(NE)                                                                                            \
    nce_pair_bwd_kernel<NE><<<bgrid, kNceThreads, smem, st>>>(w.z, P, N, w.idx[0], w.idx[1], w.idx[2], w.cnt, w.lse[0], \
                                                             w.lse[1], inv_tau, w_pos, w_neg, grad_out, w.dzp)
    if (P <= 64) AVCTC_NCE_BWD(1);
    else if (P <= 128) AVCTC_NCE_BWD(2);
    else if (P <= 192) AVCTC_NCE_BWD(3);
    else AVCTC_NCE_BWD(4);
#undef AVCTC_NCE_BWD
    if (dtype == AVCTC_F32)
        nce_normalize_bwd_kernel<float><<<rgrid, wpb * 32, 0, st>>>(w.z, w.dzp, flat_mask, w.invn, w.nrm, N, P, reinterpret_cast<float*>(dy), ld);
    else
        nce_normalize_bwd_kernel<__nv_bfloat16><<<rgrid, wpb * 32, 0, st>>>(w.z, w.dzp, flat_mask, w.invn, w.nrm, N, P, reinterpret_cast<__nv_bfloat16*>(dy), ld);
    return (int)cudaGetLastError();
}
