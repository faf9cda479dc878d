// ctc_loss.cu — CTC loss forward/backward for sm_100a.
//
// Replaces nn.CTCLoss(blank, zero_infinity=True) as the reference calls it
// (/root/reference/model/trainer.py:25,116-117,224-225; model/decoder.py:12,28-33), i.e. ATen's
// _ctc_loss / _ctc_loss_backward, with two kernels:
//
//  ctc_scan_kernel  one CTA per (sample, direction).  The label-extended lattice (S = 2L+1 states) is
//                   held in REGISTERS, K consecutive states per lane; the alpha recurrence and the beta
//                   recurrence (run as the same recurrence on the time- and label-mirrored problem) are
//                   two CTAs that run concurrently.  Neighbour states come from __shfl_up; when the
//                   lattice needs more than one warp the warps form a systolic chain through a
//                   shared-memory ring (warp w works on frame t while warp w-1 is already ahead), so the
//                   t-loop has no __syncthreads.  Emissions are gathered D frames ahead by cp.async
//                   into a shared-memory ring.  Arithmetic is log2-domain fp32 with MUFU ex2/lg2, and every warp keeps its
//                   lattice states relative to a lane-local running integer offset (renormalised every 4 frames) so
//                   |alpha| stays O(100) instead of O(T*10): that is what lets fp32 meet 1e-4 against
//                   the fp64 reference at T=1000.
//  ctc_grad_kernel  one WARP per (t,b) row, all SMs: streams the log-prob row in (cp.async), combines
//                   alpha/beta into per-class posteriors (deterministic chains for repeated labels),
//                   writes (exp(lp) - posterior) * g as aligned 16-byte stores.  HBM-bound:
//                   algorithmic bytes = one read of log_probs + one write of grad.
#include <algorithm>

#include "common.cuh"

namespace avctc {

constexpr int kScanPrefetch = 16;  // D: emission prefetch depth (frames)
constexpr int kRenorm = 4;         // renormalisation period (frames); lag = period/2
constexpr int kRing = 32;          // handoff ring slots between neighbouring warps
constexpr int kMaxWarps = 16;
constexpr int kChainNone = 0x3fffffff;
constexpr int kChainFirst = 0x40000000;
// Largest spread (natural-log units) between the emissions of one lane's classes in one frame for which the
// probability-domain scan is used: p = exp(lp - max) >= e^-40 = 2^-58, so two frames between renormalisations stay
// inside fp32's exponent range.  Beyond it the batch is recomputed by the log-domain kernels (device-side flag).
constexpr float kLinSafeRange = 40.f;
// Workspace tail: the guard flag (first int of kFlagBytes) followed by one completion counter per sample.  Every
// probability-domain scan CTA adds 2 to done[b] (release) when all of its alpha/beta rows, nll and repeat chains are
// in memory; kDoneTarget = both directions.  The gradient pass, when it is launched right behind the scan (PDL), starts
// on a sample as soon as its counter is complete instead of waiting for the whole scan grid (see ctc_grad_lin_kernel).
constexpr size_t kFlagBytes = 256;
constexpr int kDoneTarget = 4;

struct CtcPlan {
    int K, W, S_pad, Lpad, linear;
    int Klog, Wlog, CW;          // log-domain plan of the same problem (fallback of the linear path); coff row stride
    size_t off_alpha, off_beta, off_coff_a, off_coff_b, off_nll2, off_chain, off_flag, total;
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static bool make_plan(int T, int B, int Lmax, CtcPlan* pl) {
    const int S = 2 * Lmax + 1;
    int K = 0, W = 0;
    pl->linear = 0;
    // probability-domain single-warp scan (ctc_scan_lin_kernel): K states per lane, S <= 32*K <= 512
    if (S <= 512 && avctc_tuning_get("ctc_lin", 1) != 0) {
        const int need = (S + 31) / 32;
        static const int ks[6] = {2, 4, 6, 8, 12, 16};
        for (int i = 0; i < 6; ++i) if (ks[i] >= need) { K = ks[i]; break; }
        W = 1;
        pl->linear = 1;
    }
    const int forced = avctc_tuning_get("ctc_k", 0);
    int Kl = 0, Wl = 0;
    if (forced == 2 || forced == 4 || forced == 8 || forced == 16) {
        int w = (S + 32 * forced - 1) / (32 * forced);
        if (w <= kMaxWarps) { Kl = forced; Wl = w; }
    }
    if (Kl == 0) {
        if (S <= 64) { Kl = 2; Wl = 1; }
        else if (S <= 128) { Kl = 4; Wl = 1; }
        else if (S <= 512) { Kl = 2; Wl = (S + 63) / 64; }
        else if (S <= 2048) { Kl = 4; Wl = (S + 127) / 128; }
        else if (S <= 4096) { Kl = 8; Wl = (S + 255) / 256; }
        else if (S <= 8192) { Kl = 16; Wl = (S + 511) / 512; }
        else return false;
    }
    pl->Klog = Kl; pl->Wlog = Wl;
    if (!pl->linear) { K = Kl; W = Wl; }
    pl->K = K; pl->W = W; pl->Lpad = Lmax > 0 ? Lmax : 1;
    // one set of buffers serves both layouts: rows are as long as the longer of the two, one exponent per lane
    pl->S_pad = std::max(W * 32 * K, Wl * 32 * Kl);
    pl->CW = 32 * std::max(W, Wl);
    const size_t TB = (size_t)(T > 0 ? T : 1) * (size_t)B;
    size_t o = 0;
    pl->off_alpha = o; o = align_up(o + TB * pl->S_pad * sizeof(float), 256);
    pl->off_beta = o; o = align_up(o + TB * pl->S_pad * sizeof(float), 256);
    pl->off_coff_a = o; o = align_up(o + TB * pl->CW * sizeof(int), 256);
    pl->off_coff_b = o; o = align_up(o + TB * pl->CW * sizeof(int), 256);
    pl->off_nll2 = o; o = align_up(o + (size_t)B * sizeof(double), 256);
    pl->off_chain = o; o = align_up(o + (size_t)B * pl->Lpad * sizeof(int), 256);
    pl->off_flag = o; o = align_up(o + kFlagBytes + (size_t)B * sizeof(int), 256);   // guard flag, then done[B]
    pl->total = o;
    return true;
}

struct ScanParams {
    const void* lp; int64_t stride_t, stride_b;
    int T, B, V;
    const int64_t* targets; int64_t target_stride; const int64_t* target_offsets;
    const int64_t* input_lengths; const int64_t* target_lengths;
    int Lmax, blank, store;
    float* nll;
    float* alpha; float* beta; int* coff_a; int* coff_b; double* nll2; int* chain;
    int W, S_pad, Lpad;
    int cw;            // coff row stride (ints)
    int* flag;         // device int: set to 1 by the probability-domain scan when an emission ratio is out of its safe
                       // range; the log-domain kernels run only when run_if == *flag (flag == nullptr: always)
    int run_if;
    int* done;         // per-sample completion counters (nullptr: not signalled)
    int stamp;         // debug timestamps into the flag block
};

// all lanes of a warp: everything this warp stored is visible device-wide before the counter moves
__device__ __forceinline__ void signal_done(int* done, int b, int lane, int amount) {
    if (!done) return;
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicAdd(done + b, amount);
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Debug stamps (tuning knob "ctc_stamp"): four u64 at byte 64 of the flag block = first scan CTA start (stored as ~t),
// last scan CTA end, last gradient warp end, first gradient chunk started in early mode (~t; 0 = none) -- globaltimer ns.  The flag block is zeroed by avctc_ctc_forward.
__device__ __forceinline__ void stamp_min(int* flag, int slot) {      // stored as max(~t): the block starts as zeros
    unsigned long long* s = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(flag) + 64) + slot;
    atomicMax(s, ~global_timer_ns());
}
__device__ __forceinline__ void stamp_max(int* flag, int slot) {
    unsigned long long* s = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(flag) + 64) + slot;
    atomicMax(s, global_timer_ns());
}
// debug: maximum over CTAs of a duration (ns); slots 4..7 and 10..13 of the same block (8, 9 are the gradient pass's
// control words)
__device__ __forceinline__ void stamp_dur(int* flag, int slot, unsigned long long ns) {
    unsigned long long* s = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(flag) + 64) + slot;
    atomicMax(s, ns);
}
__device__ __forceinline__ int ld_acquire_gpu_s32(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int K>
__device__ __forceinline__ void store_states(float* dst, const float (&a)[K]) {
    if constexpr (K == 2) {
        *reinterpret_cast<float2*>(dst) = make_float2(a[0], a[1]);
    } else {
#pragma unroll
        for (int i = 0; i < K / 4; ++i)
            reinterpret_cast<float4*>(dst)[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
    }
}

template <int K, typename TIn, bool MULTI>
__global__ void __launch_bounds__(32 * kMaxWarps) ctc_scan_kernel(const ScanParams p) {
    constexpr int KL = K / 2;
    constexpr int D = (K >= 16) ? kScanPrefetch / 4 : (K >= 8) ? kScanPrefetch / 2 : kScanPrefetch;
    constexpr int RN = (D < kRenorm) ? D : kRenorm;
    pdl_launch_dependents();
    pdl_wait();             // PDL launch: the probability-domain scan (and its guard flag) is complete from here on
    if (p.flag && *reinterpret_cast<volatile int*>(p.flag) != p.run_if) return;   // fallback launch, not needed
    const int b = blockIdx.x;
    const int dir = blockIdx.y;  // 0: alpha (forward in time), 1: beta (mirrored problem)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // handoff ring: one 16-byte slot {value bits, offset lo, offset hi, frame tag} per frame, written and read
    // with single 128-bit shared accesses, so no fence is needed between payload and tag (a fence in the
    // frame loop would also wait for the emission prefetches in flight).
    __shared__ __align__(16) int4 ring[MULTI ? kMaxWarps : 1][kRing];
    __shared__ int progress[kMaxWarps];
    __shared__ double fin[2];

    long long tbl = p.input_lengths[b];
    long long tll = p.target_lengths[b];
    const int Tb = (int)(tbl < 0 ? 0 : (tbl > p.T ? p.T : tbl));
    const int L = (int)(tll < 0 ? 0 : (tll > p.Lmax ? p.Lmax : tll));
    const int S = 2 * L + 1;
    const int64_t* tgt = p.targets + (p.target_offsets ? p.target_offsets[b] : (int64_t)b * p.target_stride);

    if (Tb == 0) {  // ATen: input_length 0 -> nll = 0 if L == 0 else inf
        if (dir == 0 && tid == 0) {
            p.nll[b] = (L == 0) ? 0.f : CUDART_INF_F;
            if (p.nll2) p.nll2[b] = (L == 0) ? 0.0 : (double)CUDART_INF_F;
        }
        if (dir == 0 && p.chain)
            for (int j = tid; j < L; j += blockDim.x) p.chain[(size_t)b * p.Lpad + j] = kChainFirst | kChainNone;
        return;
    }
    if (tid < kMaxWarps) progress[tid] = 0;
    if (tid < 2) fin[tid] = -(double)CUDART_INF_F;
    if (MULTI)
        for (int i = tid; i < kMaxWarps * kRing; i += blockDim.x) ring[i / kRing][i % kRing] = make_int4(0, 0, 0, -1);
    __syncthreads();

    const int Wact = (S + 32 * K - 1) / (32 * K);  // warps that own at least one real state

    if (warp < Wact) {
        const int g = tid;             // global lane: owns (mirrored) states g*K .. g*K+K-1
        const int nvalid = min(max(S - g * K, 0), K);
        int lab[KL];
        bool skip[KL];
#pragma unroll
        for (int i = 0; i < KL; ++i) {
            const int j = g * KL + i;  // label position in (mirrored) order
            lab[i] = p.blank; skip[i] = false;
            if (j < L) {
                long long c = tgt[dir ? (L - 1 - j) : j];
                c = c < 0 ? 0 : (c >= p.V ? p.V - 1 : c);
                lab[i] = (int)c;
                if (j >= 1) {
                    long long cp = tgt[dir ? (L - j) : (j - 1)];
                    cp = cp < 0 ? 0 : (cp >= p.V ? p.V - 1 : cp);
                    skip[i] = (cp != c);
                }
            }
        }
        // everything that depends on the frame is a running pointer / offset: the frame loop is issue-bound
        // for a single warp, so no multiplies or 64-bit index math inside it.
        long long fstep = dir ? -p.stride_t : p.stride_t;                      // elements per scan frame
        asm volatile("" : "+l"(fstep));
        const TIn* row0 = reinterpret_cast<const TIn*>(p.lp) + (int64_t)b * p.stride_b +
                          (int64_t)(dir ? Tb - 1 : 0) * p.stride_t;           // scan frame 0
        const TIn* gp_b = row0 + p.blank + fstep;                              // gather pointers, scan frame 1
        const TIn* gp_l[KL];
#pragma unroll
        for (int i = 0; i < KL; ++i) gp_l[i] = row0 + lab[i] + fstep;

        // ---- frame 0
        float a[K];
#pragma unroll
        for (int j = 0; j < K; ++j) a[j] = AVCTC_NEG_BIG;   // finite "-inf" (see common.cuh)
        if (g == 0) {
            a[0] = to_float(__ldg(row0 + p.blank)) * AVCTC_LOG2E;
            if (L > 0) a[1] = to_float(__ldg(row0 + lab[0])) * AVCTC_LOG2E;
        }
        // Every LANE keeps its K states relative to its own running INTEGER offset C (log2 units): the states
        // that carry posterior mass are often 2^-200 below the lattice maximum, so a warp-wide offset leaves
        // them at |a| ~ 200 (ulp 1.5e-5) and the rounding random-walk over T=1000 frames reaches 2e-4; with
        // lane-local offsets |a| stays O(10) and the error drops ~10x (emulated: 2e-5).
        int C = 0;
        float dC = 0.f;              // float(C of lane-1) - C, refreshed whenever offsets change
        int cons = 0;
        const size_t rowi0 = (size_t)b * p.T + (dir ? Tb - 1 : 0);
        float* wsp = (dir ? p.beta : p.alpha) + (p.store ? rowi0 * p.S_pad + (size_t)g * K : 0);
        int* cfp = (dir ? p.coff_b : p.coff_a) + (p.store ? rowi0 * p.cw + g : 0);
        long long wstep = dir ? -(long long)p.S_pad : (long long)p.S_pad;
        long long cstep = dir ? -(long long)p.cw : (long long)p.cw;
        asm volatile("" : "+l"(wstep), "+l"(cstep));   // keep the steps in registers (no per-frame recompute)
        const bool do_store = p.store != 0;
        const bool is_lane0 = (lane == 0);

        auto publish = [&](int tau) {
            if (do_store) {
                store_states<K>(wsp, a);
                *cfp = C;
                wsp += wstep; cfp += cstep;
            }
            if (MULTI) {
                if (warp < Wact - 1) {
                    if (lane == 31) {
                        const int need = tau - kRing + 2;   // consumer must have finished frame tau-kRing+1
                        while (cons < need) cons = ld_volatile_shared_s32(&progress[warp + 1]);
                        st_volatile_shared_v4(&ring[warp][tau & (kRing - 1)],
                                              make_int4(__float_as_int(a[K - 1]), C, 0, tau));
                    }
                }
                if (warp > 0 && is_lane0) st_volatile_shared_s32(&progress[warp], tau + 1);
            }
        };
        publish(0);

        // ---- emission staging: cp.async (LDGSTS) gathers D frames ahead into a shared-memory ring.
        // (A register ring does not work here: 6 scoreboard slots cannot track D loads individually, so
        // every frame ends up waiting for the newest load = one full memory latency per frame.)
        extern __shared__ __align__(16) unsigned em_raw[];   // [D][blockDim.x][KL+1] 32-bit words
        const int slot_words = blockDim.x * (KL + 1);
        unsigned* em_mine = em_raw + tid * (KL + 1);
        int par_b = 0, par_l[KL];                            // bf16: which half of the staged word
#pragma unroll
        for (int i = 0; i < KL; ++i) par_l[i] = 0;
        const int par_step = (int)(fstep & 1);
        auto issue = [&](unsigned* dst, bool live) {          // gathers the frame gp_* point at, then advances
            if (live) {
                if constexpr (sizeof(TIn) == 4) {
                    cp_async_4(dst, gp_b);
#pragma unroll
                    for (int i = 0; i < KL; ++i) cp_async_4(dst + 1 + i, gp_l[i]);
                } else {   // 2-byte elements: copy the aligned 4-byte word holding the element
                    cp_async_4(dst, reinterpret_cast<const void*>(reinterpret_cast<uintptr_t>(gp_b) & ~(uintptr_t)3));
#pragma unroll
                    for (int i = 0; i < KL; ++i)
                        cp_async_4(dst + 1 + i,
                                   reinterpret_cast<const void*>(reinterpret_cast<uintptr_t>(gp_l[i]) & ~(uintptr_t)3));
                }
            }
            cp_async_commit();
            gp_b += fstep;
#pragma unroll
            for (int i = 0; i < KL; ++i) gp_l[i] += fstep;
        };
        if constexpr (sizeof(TIn) == 2) {   // parity of the element address at scan frame 1
            par_b = (int)((reinterpret_cast<uintptr_t>(gp_b) >> 1) & 1);
#pragma unroll
            for (int i = 0; i < KL; ++i) par_l[i] = (int)((reinterpret_cast<uintptr_t>(gp_l[i]) >> 1) & 1);
        }
        for (int d = 1; d <= D; ++d) issue(em_mine + (d & (D - 1)) * slot_words, d < Tb);

        int slot = 1 & (D - 1);
        int4 nxt_slot = make_int4(0, 0, 0, -1);
        if (MULTI) {
            if (warp > 0) nxt_slot = ld_volatile_shared_v4(&ring[warp - 1][0]);
        }
        int live_left = Tb - 1 - D;           // frames still to be gathered after the prologue
#pragma unroll 1
        for (int tau = 1; tau < Tb; ++tau) {
            cp_async_wait<D - 1>();           // frames <= tau have landed (own copies only: no barrier needed)
            unsigned* sl_ptr = em_mine + slot * slot_words;
            float vb, vl[KL];
            if constexpr (sizeof(TIn) == 4) {
                if constexpr (KL == 1) {
                    const uint2 w = *reinterpret_cast<const uint2*>(sl_ptr);
                    vb = __uint_as_float(w.x) * AVCTC_LOG2E;
                    vl[0] = __uint_as_float(w.y) * AVCTC_LOG2E;
                } else {
                    vb = __uint_as_float(sl_ptr[0]) * AVCTC_LOG2E;
#pragma unroll
                    for (int i = 0; i < KL; ++i) vl[i] = __uint_as_float(sl_ptr[1 + i]) * AVCTC_LOG2E;
                }
            } else {
                auto pick = [](unsigned w, int par) { return __uint_as_float(par ? (w & 0xffff0000u) : (w << 16)); };
                vb = pick(sl_ptr[0], par_b) * AVCTC_LOG2E;
                par_b ^= par_step;
#pragma unroll
                for (int i = 0; i < KL; ++i) {
                    vl[i] = pick(sl_ptr[1 + i], par_l[i]) * AVCTC_LOG2E;
                    par_l[i] ^= par_step;
                }
            }
            issue(sl_ptr, live_left > 0);     // refills the slot just read (frame tau + D)
            --live_left;
            slot = (slot + 1) & (D - 1);

            float prev_in = AVCTC_NEG_BIG;
            int pc = 0;
            if (MULTI) {
                if (warp > 0) {
                    int4 sl = nxt_slot;               // read at the end of the previous frame
                    while (sl.w != tau - 1) sl = ld_volatile_shared_v4(&ring[warp - 1][(tau - 1) & (kRing - 1)]);
                    pc = sl.y;
                    prev_in = __int_as_float(sl.x) + (float)(pc - C);
                }
            }
            const float up = __shfl_up_sync(kFullMask, a[K - 1], 1);
            const float prev = is_lane0 ? prev_in : up + dC;
            float nw[K];
#pragma unroll
            for (int i = 0; i < KL; ++i) {
                const float below = (i == 0) ? prev : a[2 * i - 1];
                // invalid states get the sentinel as emission, which absorbs: no select on the chain
                const float eb = (2 * i < nvalid) ? vb : AVCTC_NEG_BIG;
                const float el = (2 * i + 1 < nvalid) ? vl[i] : AVCTC_NEG_BIG;
                nw[2 * i] = lse2_plus(a[2 * i], below, eb);
                nw[2 * i + 1] = lse3_plus(a[2 * i + 1], a[2 * i], skip[i] ? below : AVCTC_NEG_BIG, el);
            }
#pragma unroll
            for (int j = 0; j < K; ++j) a[j] = nw[j];
            if ((tau & (RN - 1)) == 0) {              // lane-local renormalisation, no reduction needed
                float ml = a[0];
#pragma unroll
                for (int j = 1; j < K; ++j) ml = fmaxf(ml, a[j]);
                const bool has_mass = ml > 0.5f * AVCTC_NEG_BIG;
                const float mq = has_mass ? floorf(ml) : 0.f;   // integer shift: offsets stay exact in int32
#pragma unroll
                for (int j = 0; j < K; ++j) a[j] = fmaxf(a[j] - mq, AVCTC_NEG_BIG);
                C += (int)mq;
                int upC = __shfl_up_sync(kFullMask, C, 1);
                if (is_lane0) upC = (MULTI && warp > 0) ? pc : C;
                if (!has_mass) C = upC;               // empty lane: adopt the frame mass will arrive in
                upC = __shfl_up_sync(kFullMask, C, 1);   // neighbours may have adopted too: re-read
                if (is_lane0) upC = C;                // lane 0 converts explicitly from the ring slot
                dC = (float)(upC - C);
            }
            publish(tau);
            if (MULTI) {   // early read of the next frame's boundary slot (usually already published)
                if (warp > 0) nxt_slot = ld_volatile_shared_v4(&ring[warp - 1][tau & (kRing - 1)]);
            }
        }
        cp_async_wait<0>();
        // ---- the two terminal states (mirrored index S-1 / S-2 are terminal for alpha)
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int s = g * K + j;
            const double tv = (a[j] > 0.5f * AVCTC_NEG_BIG) ? (double)a[j] + (double)C : -(double)CUDART_INF_F;
            if (s == S - 1) fin[0] = tv;
            if (s == S - 2) fin[1] = tv;
        }
    }
    __syncthreads();
    if (dir == 0) {
        if (tid == 0) {
            const double x = fin[0], y = fin[1];
            const double m = fmax(x, y), n = fmin(x, y);
            double ll2;
            if (m == -(double)CUDART_INF_F) ll2 = m;
            else ll2 = m + log2(1.0 + exp2(n - m));
            p.nll[b] = (float)(-ll2 * AVCTC_LN2_D);
            if (p.nll2) p.nll2[b] = -ll2;
        }
        if (p.chain) {
            // repeated-label chains for the deterministic per-class posterior sum in ctc_grad_kernel
            for (int j = tid; j < L; j += blockDim.x) {
                const long long c = tgt[j];
                bool first = true;
                for (int k = 0; k < j; ++k) if (tgt[k] == c) { first = false; break; }
                int nxt = kChainNone;
                for (int k = j + 1; k < L; ++k) if (tgt[k] == c) { nxt = k; break; }
                p.chain[(size_t)b * p.Lpad + j] = nxt | (first ? kChainFirst : 0);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ctc_scan_lin_kernel — the same recurrence in the PROBABILITY domain, one warp per (sample, direction).
//
// A lane holds K consecutive lattice states as fp32 mantissas a[j] relative to a lane-local integer exponent C:
// alpha(s) = a[j] * 2^C.  A frame is then   a' = (a[s] + a[s-1] + skip * a[s-2]) * p   — two adds and one multiply
// per state instead of a 3-way log-sum-exp (3 MUFU ex2 + 1 lg2 on the dependent chain), with
//   * emissions p = 2^(lp*log2e - E), E = floor of the lane's largest emission exponent this frame (computed one
//     frame ahead, off the chain), so the best class of the lane has p in (0.5, 1];
//   * an exact power-of-two renormalisation of the lane every frame (lane maximum back into [1,2)), exponents
//     accumulated in C as integers — no rounding is introduced by scaling, and fp32 products carry a relative
//     error of 2^-24 per frame instead of an absolute error in the log domain;
//   * the neighbour lane's last state arrives by __shfl_up together with its exponent and is converted with an
//     exact 2^(C_up - C) factor; a lane that is empty adopts the upstream exponent, a lane whose upstream is more
//     than 2^40 larger is rescaled down first (its own mass is then below fp32 resolution of the incoming mass).
// States that fall more than 2^-126 below their lane's maximum flush to zero; they are re-fed by their lower
// neighbours on the next frame, and a flushed state cannot carry posterior mass unless two classes of the label
// set differ by more than e^-87 in the same frame.
template <int K, typename TIn>
__global__ void __launch_bounds__(32) ctc_scan_lin_kernel(const ScanParams p) {
    constexpr int KL = K / 2;
    constexpr int D = (K >= 12) ? 8 : 16;     // emission prefetch depth (frames)
    constexpr int EW = KL + 1;                // staged words per lane per frame: blank + KL labels
    const int b = blockIdx.x;
    const int dir = blockIdx.y;               // 0: alpha, 1: beta (time- and label-mirrored problem)
    const int lane = threadIdx.x;
    __shared__ double fin[2];

    long long tbl = p.input_lengths[b];
    long long tll = p.target_lengths[b];
    const int Tb = (int)(tbl < 0 ? 0 : (tbl > p.T ? p.T : tbl));
    const int L = (int)(tll < 0 ? 0 : (tll > p.Lmax ? p.Lmax : tll));
    const int S = 2 * L + 1;
    const int64_t* tgt = p.targets + (p.target_offsets ? p.target_offsets[b] : (int64_t)b * p.target_stride);

    if (Tb == 0) {  // ATen: input_length 0 -> nll = 0 if L == 0 else inf
        if (dir == 0 && lane == 0) {
            p.nll[b] = (L == 0) ? 0.f : CUDART_INF_F;
            if (p.nll2) p.nll2[b] = (L == 0) ? 0.0 : (double)CUDART_INF_F;
        }
        if (dir == 0 && p.chain)
            for (int j = lane; j < L; j += 32) p.chain[(size_t)b * p.Lpad + j] = kChainFirst | kChainNone;
        signal_done(p.done, b, lane, 2);
        return;
    }
    if (lane < 2) fin[lane] = -(double)CUDART_INF_F;
    __syncwarp();

    const int g = lane;                        // owns (mirrored) states g*K .. g*K+K-1
    const int nvalid = min(max(S - g * K, 0), K);
    int lab[KL];
    bool skip[KL];
#pragma unroll
    for (int i = 0; i < KL; ++i) {
        const int j = g * KL + i;
        lab[i] = p.blank; skip[i] = false;
        if (j < L) {
            long long c = tgt[dir ? (L - 1 - j) : j];
            c = c < 0 ? 0 : (c >= p.V ? p.V - 1 : c);
            lab[i] = (int)c;
            if (j >= 1) {
                long long cp = tgt[dir ? (L - j) : (j - 1)];
                cp = cp < 0 ? 0 : (cp >= p.V ? p.V - 1 : cp);
                skip[i] = (cp != c);
            }
        }
    }
    long long fstep = dir ? -p.stride_t : p.stride_t;
    asm volatile("" : "+l"(fstep));
    const TIn* row0 = reinterpret_cast<const TIn*>(p.lp) + (int64_t)b * p.stride_b +
                      (int64_t)(dir ? Tb - 1 : 0) * p.stride_t;
    const TIn* gp_b = row0 + p.blank + fstep;
    const TIn* gp_l[KL];
#pragma unroll
    for (int i = 0; i < KL; ++i) gp_l[i] = row0 + lab[i] + fstep;

    // ---- frame 0
    float a[K];
#pragma unroll
    for (int j = 0; j < K; ++j) a[j] = 0.f;
    int C = 0;
    bool has_mass = false;
    if (g == 0) {
        const float xb = to_float(__ldg(row0 + p.blank)) * AVCTC_LOG2E;
        const float xl = (L > 0) ? to_float(__ldg(row0 + lab[0])) * AVCTC_LOG2E : AVCTC_NEG_INF;
        const float mx = fmaxf(xb, xl);
        const float Ef = (mx > -1.0e29f) ? floorf(mx) : 0.f;
        a[0] = ex2_approx(xb - Ef);
        a[1] = ex2_approx(xl - Ef);
        C = (int)Ef;
        has_mass = fmaxf(a[0], a[1]) > 0.f;
        if (L > 0 && p.flag && fabsf(xb - xl) > kLinSafeRange * AVCTC_LOG2E) *p.flag = 1;
    }
    const size_t rowi0 = (size_t)b * p.T + (dir ? Tb - 1 : 0);
    float* wsp = (dir ? p.beta : p.alpha) + (p.store ? rowi0 * p.S_pad + (size_t)g * K : 0);
    int* cfp = (dir ? p.coff_b : p.coff_a) + (p.store ? rowi0 * p.cw + g : 0);
    long long wstep = dir ? -(long long)p.S_pad : (long long)p.S_pad;
    long long cstep = dir ? -(long long)p.cw : (long long)p.cw;
    asm volatile("" : "+l"(wstep), "+l"(cstep));
    const bool do_store = p.store != 0;
    auto publish = [&]() {
        if (do_store) {
            if constexpr (K % 4 == 0) {
#pragma unroll
                for (int i = 0; i < K / 4; ++i)
                    reinterpret_cast<float4*>(wsp)[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < K / 2; ++i) reinterpret_cast<float2*>(wsp)[i] = make_float2(a[2 * i], a[2 * i + 1]);
            }
            *cfp = C;
            wsp += wstep; cfp += cstep;
        }
    };
    publish();

    // ---- emission staging (cp.async gathers, D frames ahead)
    extern __shared__ __align__(16) unsigned em_raw[];   // [D][32][EW]
    constexpr int slot_words = 32 * EW;
    unsigned* em_mine = em_raw + lane * EW;
    int par_b = 0, par_l[KL];
#pragma unroll
    for (int i = 0; i < KL; ++i) par_l[i] = 0;
    const int par_step = (int)(fstep & 1);
    const bool gathers = nvalid > 0;
    auto issue = [&](unsigned* dst, bool live) {
        if (live && gathers) {
            if constexpr (sizeof(TIn) == 4) {
                cp_async_4(dst, gp_b);
#pragma unroll
                for (int i = 0; i < KL; ++i) cp_async_4(dst + 1 + i, gp_l[i]);
            } else {
                cp_async_4(dst, reinterpret_cast<const void*>(reinterpret_cast<uintptr_t>(gp_b) & ~(uintptr_t)3));
#pragma unroll
                for (int i = 0; i < KL; ++i)
                    cp_async_4(dst + 1 + i,
                               reinterpret_cast<const void*>(reinterpret_cast<uintptr_t>(gp_l[i]) & ~(uintptr_t)3));
            }
        }
        cp_async_commit();
        gp_b += fstep;
#pragma unroll
        for (int i = 0; i < KL; ++i) gp_l[i] += fstep;
    };
    if constexpr (sizeof(TIn) == 2) {
        par_b = (int)((reinterpret_cast<uintptr_t>(gp_b) >> 1) & 1);
#pragma unroll
        for (int i = 0; i < KL; ++i) par_l[i] = (int)((reinterpret_cast<uintptr_t>(gp_l[i]) >> 1) & 1);
    }
    for (int d = 1; d <= D; ++d) issue(em_mine + (d & (D - 1)) * slot_words, d < Tb);

    // emissions of one frame -> per-state probabilities relative to the lane exponent E
    float pb[KL], pl[KL];
    int E = 0;
    int slot = 1 & (D - 1);
    int live_left = Tb - 1 - D;
    auto prepare = [&]() {          // consumes the staged frame in `slot`, refills it, advances
        cp_async_wait<D - 1>();
        unsigned* sl_ptr = em_mine + slot * slot_words;
        float xb, xl[KL];
        if constexpr (sizeof(TIn) == 4) {
            xb = __uint_as_float(sl_ptr[0]);
#pragma unroll
            for (int i = 0; i < KL; ++i) xl[i] = __uint_as_float(sl_ptr[1 + i]);
        } else {
            auto pick = [](unsigned w, int par) { return __uint_as_float(par ? (w & 0xffff0000u) : (w << 16)); };
            xb = pick(sl_ptr[0], par_b);
            par_b ^= par_step;
#pragma unroll
            for (int i = 0; i < KL; ++i) { xl[i] = pick(sl_ptr[1 + i], par_l[i]); par_l[i] ^= par_step; }
        }
        issue(sl_ptr, live_left > 0);
        --live_left;
        slot = (slot + 1) & (D - 1);
        xb = (nvalid > 0) ? xb * AVCTC_LOG2E : AVCTC_NEG_INF;
        float mx = xb, mn = (nvalid > 0) ? xb : CUDART_INF_F;
#pragma unroll
        for (int i = 0; i < KL; ++i) {
            xl[i] = (2 * i + 1 < nvalid) ? xl[i] * AVCTC_LOG2E : AVCTC_NEG_INF;
            mx = fmaxf(mx, xl[i]);
            mn = fminf(mn, (2 * i + 1 < nvalid) ? xl[i] : CUDART_INF_F);
        }
        if (mx - mn > kLinSafeRange * AVCTC_LOG2E && p.flag) *p.flag = 1;   // batch is redone by the log-domain kernels
        const float Ef = (mx > -1.0e29f) ? floorf(mx) : 0.f;
        E = (int)Ef;
        const float pblank = ex2_approx(xb - Ef);
#pragma unroll
        for (int i = 0; i < KL; ++i) {
            pb[i] = (2 * i < nvalid) ? pblank : 0.f;
            pl[i] = ex2_approx(xl[i] - Ef);
        }
    };
    if (Tb > 1) prepare();

#pragma unroll 1
    for (int tau = 1; tau < Tb; ++tau) {
        const float up_a = __shfl_up_sync(kFullMask, a[K - 1], 1);
        const int up_C = __shfl_up_sync(kFullMask, C, 1);
        // this frame's probabilities (prepared one frame ahead)
        float cb_[KL], cl_[KL];
#pragma unroll
        for (int i = 0; i < KL; ++i) { cb_[i] = pb[i]; cl_[i] = pl[i]; }
        const int Ecur = E;
        if (tau + 1 < Tb) prepare();                     // next frame's emissions, off the dependent chain
        // incoming boundary state: exact power-of-two conversion between lane exponents
        const bool has_up = (lane > 0) && (up_a > 0.f);
        int d = has_up ? up_C - C : 0;
        if (!has_mass && has_up) { C = up_C; d = 0; }
        if (d > 40) {
            const int sh = d - 40;
            const float k = (sh < 126) ? __int_as_float((127 - sh) << 23) : 0.f;
#pragma unroll
            for (int j = 0; j < K; ++j) a[j] *= k;
            C += sh; d = 40;
        }
        const float sc = (has_up && d > -126) ? __int_as_float((127 + d) << 23) : 0.f;
        const float prev = up_a * sc;
        float nw[K];
#pragma unroll
        for (int i = 0; i < KL; ++i) {
            const float below = (i == 0) ? prev : a[2 * i - 1];
            nw[2 * i] = (a[2 * i] + below) * cb_[i];
            nw[2 * i + 1] = (a[2 * i + 1] + a[2 * i] + (skip[i] ? below : 0.f)) * cl_[i];
        }
        float mm = nw[0];
#pragma unroll
        for (int j = 1; j < K; ++j) mm = fmaxf(mm, nw[j]);
        has_mass = mm > 0.f;
        const int e = has_mass ? ((__float_as_int(mm) >> 23) - 127) : 0;
        const float sn = __int_as_float((127 - e) << 23);
#pragma unroll
        for (int j = 0; j < K; ++j) a[j] = nw[j] * sn;
        C += Ecur + e;
        publish();
    }
    cp_async_wait<0>();
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const int st = g * K + j;
        const double tv = (a[j] > 0.f) ? log2((double)a[j]) + (double)C : -(double)CUDART_INF_F;
        if (st == S - 1) fin[0] = tv;
        if (st == S - 2) fin[1] = tv;
    }
    __syncwarp();
    if (dir == 0) {
        if (lane == 0) {
            const double x = fin[0], y = fin[1];
            const double m = fmax(x, y), n = fmin(x, y);
            double ll2;
            if (m == -(double)CUDART_INF_F) ll2 = m;
            else ll2 = m + log2(1.0 + exp2(n - m));
            p.nll[b] = (float)(-ll2 * AVCTC_LN2_D);
            if (p.nll2) p.nll2[b] = -ll2;
        }
        if (p.chain) {
            for (int j = lane; j < L; j += 32) {
                const long long c = tgt[j];
                bool first = true;
                for (int k = 0; k < j; ++k) if (tgt[k] == c) { first = false; break; }
                int nxt = kChainNone;
                for (int k = j + 1; k < L; ++k) if (tgt[k] == c) { nxt = k; break; }
                p.chain[(size_t)b * p.Lpad + j] = nxt | (first ? kChainFirst : 0);
            }
        }
    }
    signal_done(p.done, b, lane, 2);
}

// ------------------------------------------------------------------------------------------------
struct GradParams {
    const void* lp; int64_t stride_t, stride_b;
    int T, B, V;
    const int64_t* targets; int64_t target_stride; const int64_t* target_offsets;
    const int64_t* input_lengths; const int64_t* target_lengths;
    int Lmax, blank, reduction, zero_infinity;
    const float* nll; const float* grad_out; int64_t grad_out_stride;
    void* grad;
    const float* alpha; const float* beta; const int* coff_a; const int* coff_b;
    const double* nll2; const int* chain;
    int K, W, S_pad, Lpad, linear;
    int cw; const int* flag; int run_if;
    int stamp;        // debug timestamps into the flag block
    const int* done;  // grad_lin: per-sample completion counters of the scan (nullptr: wait for the whole grid)
    int prefetch;     // grad_lin: L2-prefetch the warp's next log-prob row (tuning knob "ctc_pf")
    int row_floats;   // per-warp smem floats for one staged row (>= V + 8, multiple of 4)
    int w_floats;     // per-warp smem floats for state weights (>= 2*Lmax+1, multiple of 4)
    int stage;        // grad_lin, fp32: the whole log-prob row goes to per-warp shared memory with cp.async first
};

template <typename T>
struct VecTraits;
template <>
struct VecTraits<float> { static constexpr int kVec = 4; };
template <>
struct VecTraits<__nv_bfloat16> { static constexpr int kVec = 8; };

__device__ __forceinline__ int stage_row(const __nv_bfloat16* grow, int V, float* buf, int lane) {
    const int o = (int)((reinterpret_cast<uintptr_t>(grow) >> 1) & 7);
    const int vstart = (o + 7) & ~7, vend = (o + V) & ~7;
    auto scalar = [&](int lo, int hi) {
        for (int pidx = lo + lane; pidx < hi; pidx += 32) buf[pidx] = __bfloat162float(grow[pidx - o]);
    };
    if (vstart < vend) {
        scalar(o, vstart);
        for (int q = vstart + 8 * lane; q < vend; q += 256) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(grow + (q - o)));
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
            const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
            const float2 f2 = __bfloat1622float2(h[2]), f3 = __bfloat1622float2(h[3]);
            *reinterpret_cast<float4*>(buf + q) = make_float4(f0.x, f0.y, f1.x, f1.y);
            *reinterpret_cast<float4*>(buf + q + 4) = make_float4(f2.x, f2.y, f3.x, f3.y);
        }
        scalar(vend, o + V);
    } else {
        scalar(o, o + V);
    }
    return o;
}

// write one output row from fp32 shared memory (buf[o + c], o = misalignment of the global row) or
// zeros (buf == nullptr) with aligned 16-byte stores plus scalar head/tail.
__device__ __forceinline__ void write_row(float* grow, int V, const float* buf, int o, int lane) {
    const int vstart = (o + 3) & ~3, vend = (o + V) & ~3;
    auto scalar = [&](int lo, int hi) {
        for (int pidx = lo + lane; pidx < hi; pidx += 32) grow[pidx - o] = buf ? buf[pidx] : 0.f;
    };
    if (vstart < vend) {
        scalar(o, vstart);
        for (int q = vstart + 4 * lane; q < vend; q += 128) {
            const float4 v = buf ? *reinterpret_cast<const float4*>(buf + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            __stcs(reinterpret_cast<float4*>(grow + (q - o)), v);
        }
        scalar(vend, o + V);
    } else {
        scalar(o, o + V);
    }
}
__device__ __forceinline__ void write_row(__nv_bfloat16* grow, int V, const float* buf, int o, int lane) {
    const int vstart = (o + 7) & ~7, vend = (o + V) & ~7;
    auto scalar = [&](int lo, int hi) {
        for (int pidx = lo + lane; pidx < hi; pidx += 32) grow[pidx - o] = __float2bfloat16(buf ? buf[pidx] : 0.f);
    };
    if (vstart < vend) {
        scalar(o, vstart);
        for (int q = vstart + 8 * lane; q < vend; q += 256) {
            uint4 raw = make_uint4(0, 0, 0, 0);
            if (buf) {
                const float4 x = *reinterpret_cast<const float4*>(buf + q);
                const float4 y = *reinterpret_cast<const float4*>(buf + q + 4);
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
                h[0] = __floats2bfloat162_rn(x.x, x.y); h[1] = __floats2bfloat162_rn(x.z, x.w);
                h[2] = __floats2bfloat162_rn(y.x, y.y); h[3] = __floats2bfloat162_rn(y.z, y.w);
            }
            __stcs(reinterpret_cast<uint4*>(grow + (q - o)), raw);
        }
        scalar(vend, o + V);
    } else {
        scalar(o, o + V);
    }
}

template <typename TIn>
__global__ void __launch_bounds__(256) ctc_grad_kernel(const GradParams p) {
    extern __shared__ __align__(16) float smem[];
    constexpr int kVec = VecTraits<TIn>::kVec;
    constexpr int NS = 6;   // states per lane whose alpha/beta loads are issued before the row lands
    pdl_launch_dependents();
    pdl_wait();             // PDL launch: everything before it on the stream (scan, guard flag) is complete from here on
    if (p.flag && *p.flag != p.run_if) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    float* rowbuf = smem + (size_t)warp * (2 * p.row_floats + p.w_floats);
    float* outbuf = rowbuf + p.row_floats;
    float* wbuf = outbuf + p.row_floats;
    const long long rows = (long long)p.T * p.B;
    TIn* grad = reinterpret_cast<TIn*>(p.grad);
    const TIn* lpbase = reinterpret_cast<const TIn*>(p.lp);

    for (long long row = (long long)blockIdx.x * nwarp + warp; row < rows; row += (long long)gridDim.x * nwarp) {
        const int t = (int)(row / p.B), b = (int)(row % p.B);
        TIn* grow = grad + (size_t)row * p.V;
        const int o_out = (int)((reinterpret_cast<uintptr_t>(grow) / sizeof(TIn)) & (kVec - 1));
        long long tbl = p.input_lengths[b], tll = p.target_lengths[b];
        const int Tb = (int)(tbl < 0 ? 0 : (tbl > p.T ? p.T : tbl));
        const int L = (int)(tll < 0 ? 0 : (tll > p.Lmax ? p.Lmax : tll));
        const int S = 2 * L + 1;
        const float nllb = p.nll[b];
        if (t >= Tb || (p.zero_infinity && nllb == CUDART_INF_F)) {
            write_row(grow, p.V, nullptr, o_out, lane);
            continue;
        }
        const TIn* lrow = lpbase + (int64_t)t * p.stride_t + (int64_t)b * p.stride_b;
        const int o_in = stage_row(lrow, p.V, rowbuf, lane);

        const int64_t* tgt = p.targets + (p.target_offsets ? p.target_offsets[b] : (int64_t)b * p.target_stride);
        float scale = p.grad_out[(int64_t)b * p.grad_out_stride];
        if (p.reduction == AVCTC_REDUCE_MEAN) scale /= ((float)p.B * (float)(L > 1 ? L : 1));
        const size_t rowi = (size_t)b * p.T + t;
        const float* arow = p.alpha + rowi * p.S_pad;
        const float* brow = p.beta + rowi * p.S_pad;
        const int* ca = p.coff_a + rowi * p.cw;
        const int* cb = p.coff_b + rowi * p.cw;
        const double nll2 = p.nll2[b];
        const int perw = p.K;   // states per lane (offsets are per lane)

        // log2 of alpha_t(s)*beta_t(s)/P, offsets folded in fp64; loads issued before the row is waited on
        // state weight = alpha*beta/(P*p): (mantissa product, log2 exponent) pairs; linear workspaces hold fp32
        // mantissas relative to integer lane exponents, log workspaces hold log2 values (mantissa 1).
        auto state_term = [&](int s, float& mant) -> double {
            const int sm = S - 1 - s;
            const int cexp = ca[s / perw] + cb[sm / perw];
            if (p.linear) {
                const float av = arow[s], bv = brow[sm];
                if (!(av >= 1.17549435e-38f) || !(bv >= 1.17549435e-38f)) { mant = 0.f; return 0.0; }
                const int ia = __float_as_int(av), ib = __float_as_int(bv);
                mant = __int_as_float((ia & 0x007fffff) | 0x3f800000) * __int_as_float((ib & 0x007fffff) | 0x3f800000);
                return (double)((ia >> 23) + (ib >> 23) - 254 + cexp) + nll2;
            }
            mant = 1.f;
            return (double)arow[s] + (double)brow[sm] + (double)cexp + nll2;
        };
        double e0[NS];
        float mt[NS];
        int cls[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const int s = lane + 32 * i;
            e0[i] = 0.0; mt[i] = 0.f; cls[i] = p.blank;
            if (s < S) {
                e0[i] = state_term(s, mt[i]);
                if (s & 1) {
                    long long c = tgt[s >> 1];
                    cls[i] = (int)(c < 0 ? 0 : (c >= p.V ? p.V - 1 : c));
                }
            }
        }
        cp_async_wait<0>();
        __syncwarp();
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const int s = lane + 32 * i;
            if (s < S) wbuf[s] = mt[i] * ex2_approx((float)(e0[i] - (double)(rowbuf[o_in + cls[i]] * AVCTC_LOG2E)));
        }
        for (int s = lane + 32 * NS; s < S; s += 32) {
            int c = p.blank;
            if (s & 1) {
                long long cc = tgt[s >> 1];
                c = (int)(cc < 0 ? 0 : (cc >= p.V ? p.V - 1 : cc));
            }
            float mant;
            const double e = state_term(s, mant) - (double)(rowbuf[o_in + c] * AVCTC_LOG2E);
            wbuf[s] = mant * ex2_approx((float)e);
        }
        __syncwarp();
        // class posteriors: blank = sum of even states; label c = sum along its repeat chain
        float pb = 0.f;
        for (int s = 2 * lane; s < S; s += 64) pb += wbuf[s];
        pb = warp_sum(pb);
        const int* chain = p.chain + (size_t)b * p.Lpad;
        for (int j = lane; j < L; j += 32) {
            const int ch = chain[j];
            if (ch & kChainFirst) {
                float acc = wbuf[2 * j + 1];
                int k = ch & kChainNone;
                while (k != kChainNone) {
                    acc += wbuf[2 * k + 1];
                    k = chain[k] & kChainNone;
                }
                wbuf[2 * j + 1] = acc;   // only this lane ever reads this chain's slots
            }
        }
        // exp(lp) * g
        for (int c = lane; c < p.V; c += 32)
            outbuf[o_out + c] = ex2_approx(rowbuf[o_in + c] * AVCTC_LOG2E) * scale;
        __syncwarp();
        if (lane == 0) outbuf[o_out + p.blank] -= pb * scale;
        for (int j = lane; j < L; j += 32) {
            if (chain[j] & kChainFirst) {
                long long c = tgt[j];
                c = c < 0 ? 0 : (c >= p.V ? p.V - 1 : c);
                outbuf[o_out + (int)c] -= wbuf[2 * j + 1] * scale;
            }
        }
        __syncwarp();
        write_row(grow, p.V, outbuf, o_out, lane);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// ctc_scan_ws_kernel — the probability-domain scan, warp-specialised.  The frame loop of one (sample, direction)
// is a dependent chain on a single in-order warp (~0.25-0.5 instructions/cycle), so everything that is not the
// recurrence itself runs on the other three SM sub-partitions of the CTA:
//   warps 0,1  producers  (even / odd frame groups) wait for their group's rows (mbarrier "full"), each lane picks its
//                         lattice states' emissions out of shared memory and publishes 2^(lp*log2e - E) per state plus
//                         the lane exponent E into the ring "pring"; a consumed slot group is handed back through an
//                         "empty" mbarrier
//   warp 2     recurrence the dependent chain only: exponent conversion, adds/multiplies, exact power-of-two
//                         renormalisation every second frame -> ring "sring" (lane states + exponent); the (top state,
//                         exponent) pair for the neighbour lane is shuffled BEFORE the renormalisation so that the
//                         shuffle latency overlaps it; ring polls once per group of kWsGroup frames
//   warp 3     writer     sring -> alpha/beta workspace in HBM
//   warps 4,5  DMA        issue the TMA bulk copies (cp.async.bulk, whole log-prob rows, kWsGroup rows per mbarrier
//                         phase, kWsGroups groups ahead per producer) as soon as a slot group is empty
// Lane l of a role talks only to lane l of its neighbour role through per-lane progress words (volatile shared
// accesses, payload before progress in program order), so neither fences nor block barriers sit in the frame loop.
// Per-lane 4-byte gathers straight from HBM (ctc_scan_lin_kernel) are bound by DRAM random-access efficiency
// (measured ~1 TB/s of 32-byte sectors); full rows stream at HBM speed and are also what the hardware prefetches.
constexpr int kWsRing = 8;   // frames per ring (power of two)

constexpr int kWsGroup = 4;  // frames per TMA mbarrier phase
constexpr int kWsGroups = 4; // groups in flight per producer warp
constexpr int kWsThreads = 192;   // warps: 0,1 producers, 2 recurrence, 3 writer, 4,5 DMA (bulk-copy issue)

__device__ __forceinline__ int ws_wait_ge(const int* word, int need) {
    int v = ld_volatile_shared_s32(word);
    int spins = 0;
    while (v < need) {
        v = ld_volatile_shared_s32(word);
        if (++spins > (1 << 26)) __trap();      // never hang the GPU on a protocol bug
    }
    asm volatile("" ::: "memory");              // payload accesses stay after the poll
    return v;
}
__device__ __forceinline__ uint32_t ws_smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void ws_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ws_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ws_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ws_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    int spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (++spins > (1 << 24)) __trap();
    }
}
__device__ __forceinline__ void ws_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar) : "memory");
}

template <int K, typename TIn, bool STORE>
__global__ void __launch_bounds__(kWsThreads) ctc_scan_ws_kernel(const ScanParams p, const int row_stride_bytes) {
    constexpr int KL = K / 2;
    constexpr int G = kWsGroup, NG = kWsGroups;
    constexpr int PW = (K + 1 + 3) & ~3;      // payload words per lane per frame: K floats + one int
    constexpr int RD = kWsRing;
    constexpr int ES = (int)sizeof(TIn);
    const int b = blockIdx.x, dir = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();     // the small kernels behind this one (guarded log-domain scan, reduce) wait for it themselves
    if (p.stamp && p.flag && threadIdx.x == 0) stamp_min(p.flag, 0);
    extern __shared__ __align__(128) unsigned char ws_smem[];
    unsigned char* rowbuf = ws_smem;                                               // [2][NG][G][row_stride_bytes]
    float* pring = reinterpret_cast<float*>(ws_smem + (size_t)2 * NG * G * row_stride_bytes);   // [RD][32][PW]
    float* sring = pring + RD * 32 * PW;                                           // [RD][32][PW]
    int* prog = reinterpret_cast<int*>(sring + RD * 32 * PW);                      // [4][32]: P0, P1, R, W next frame
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(prog + 128);  // [2][NG] rows landed ("full")
    unsigned long long* mbar_empty = mbar + 2 * NG;                                // [2][NG] rows consumed ("empty")
    __shared__ double fin[2];
    __shared__ int tgt_s[256];            // clamped labels of the sample (the probability-domain plan has 2L+1 <= 512)

    long long tbl = p.input_lengths[b];
    long long tll = p.target_lengths[b];
    const int Tb = (int)(tbl < 0 ? 0 : (tbl > p.T ? p.T : tbl));
    const int L = (int)(tll < 0 ? 0 : (tll > p.Lmax ? p.Lmax : tll));
    const int S = 2 * L + 1;
    const int64_t* tgt = p.targets + (p.target_offsets ? p.target_offsets[b] : (int64_t)b * p.target_stride);

    if (Tb == 0) {
        if (dir == 0 && threadIdx.x == 0) {
            p.nll[b] = (L == 0) ? 0.f : CUDART_INF_F;
            if (p.nll2) p.nll2[b] = (L == 0) ? 0.0 : (double)CUDART_INF_F;
        }
        if (dir == 0 && p.chain)
            for (int j = threadIdx.x; j < L; j += blockDim.x) p.chain[(size_t)b * p.Lpad + j] = kChainFirst | kChainNone;
        if (p.done) {
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) atomicAdd(p.done + b, 2);
        }
        return;
    }
    if (threadIdx.x < 128) prog[threadIdx.x] = 1;
    if (threadIdx.x < 2) fin[threadIdx.x] = -(double)CUDART_INF_F;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4 * NG; ++i) ws_mbar_init(ws_smem_u32(&mbar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int g = lane;
    const int nvalid = min(max(S - g * K, 0), K);
    constexpr bool do_store = STORE;
    const TIn* row0 = reinterpret_cast<const TIn*>(p.lp) + (int64_t)b * p.stride_b +
                      (int64_t)(dir ? Tb - 1 : 0) * p.stride_t;
    const long long fstep = dir ? -p.stride_t : p.stride_t;
    float* pmine = pring + lane * PW;
    float* smine = sring + lane * PW;

    if (warp >= 4) {
        // ================================================================ DMA warps (warp 4+pw feeds producer pw)
        // Issuing the bulk copies of a group (address arithmetic, expect_tx, the rare software row) was ~40 % of a
        // producer's instruction stream and the producers were what the recurrence warp waited for; it now runs here,
        // NG groups ahead, gated by the "empty" mbarrier the producer arrives on when it has consumed a slot group.
        const int pw = warp - 4;
        // A row's 16-byte aligned window may read up to 15 bytes of its neighbours.  Below the tensor that is always
        // inside the allocation (an unaligned base is an interior pointer); above it only the row with the highest
        // address can leave the tensor: that single frame (if any) is copied by plain loads instead.
        const uintptr_t hi = reinterpret_cast<uintptr_t>(p.lp) +
                             (uintptr_t)(((long long)(p.T - 1) * p.stride_t + (long long)(p.B - 1) * p.stride_b + p.V) * ES);
        const long long fstep_b = fstep * ES;
        const uint32_t VB = (uint32_t)p.V * ES;
        int bad_f = -1;
        if (fstep >= 0 && Tb > 1) {
            const uintptr_t sl_ = reinterpret_cast<uintptr_t>(row0) + (uintptr_t)((long long)(Tb - 1) * fstep_b);
            if ((sl_ & ~(uintptr_t)15) + (((sl_ & 15) + VB + 15) & ~(uintptr_t)15) > hi) bad_f = Tb - 1;
        }
        unsigned char* myrows = rowbuf + (size_t)pw * NG * G * row_stride_bytes;
        const uint32_t myrows_s = ws_smem_u32(myrows);
        unsigned long long* mybar = mbar + pw * NG;
        // frames of my k-th group: 1 + (2k + pw)*G + j
        uintptr_t src_arm = reinterpret_cast<uintptr_t>(row0) + (uintptr_t)(fstep_b * (1 + pw * G));
        int f_arm = 1 + pw * G;
        auto arm = [&](int sl) {             // warp-wide: start the copies of frames f_arm .. f_arm+G-1 into slot group sl
            const uint32_t bar = ws_smem_u32(&mybar[sl]);
            if (bad_f >= f_arm && bad_f < f_arm + G) {
                const int j = bad_f - f_arm;
                const uintptr_t src = src_arm + (uintptr_t)((long long)j * fstep_b);
                unsigned char* dst = myrows + (size_t)(sl * G + j) * row_stride_bytes + (src & 15);
                const TIn* srow = reinterpret_cast<const TIn*>(src);
                for (int c = lane; c < p.V; c += 32) reinterpret_cast<TIn*>(dst)[c] = srow[c];
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
            }
            if (lane == 0) {
                uint32_t total = 0;
                uintptr_t src = src_arm;
                uint32_t dst = myrows_s + (uint32_t)(sl * G * row_stride_bytes);
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    const int f = f_arm + j;
                    if (f < Tb && f != bad_f) {
                        const uint32_t nb = (uint32_t)(((uint32_t)(src & 15) + VB + 15u) & ~15u);
                        ws_bulk_g2s(dst, reinterpret_cast<const void*>(src & ~(uintptr_t)15), nb, bar);
                        total += nb;
                    }
                    src += (uintptr_t)fstep_b;
                    dst += (uint32_t)row_stride_bytes;
                }
                ws_mbar_expect_tx(bar, total);       // the one arrival of this phase; tx may complete before it
            }
            src_arm += (uintptr_t)(fstep_b * 2 * G);
            f_arm += 2 * G;
        };
        const int ngroups_all = (Tb - 1 + G - 1) / G;
        const int mygroups = (ngroups_all - pw + 1) / 2;       // groups pw, pw+2, ... < ngroups_all
#pragma unroll 1
        for (int k = 0; k < mygroups; ++k) {
            const int sl = k % NG;
            if (k >= NG) ws_mbar_wait(ws_smem_u32(&mbar_empty[pw * NG + sl]), (uint32_t)((k / NG) - 1) & 1u);
            arm(sl);
        }
        return;
    }

    if (warp < 2) {
        // ================================================================ producers (warp pw: groups pw, pw+2, ...)
        const int pw = warp;
        int lab[KL];
#pragma unroll
        for (int i = 0; i < KL; ++i) {
            const int j = g * KL + i;
            lab[i] = p.blank;
            if (j < L) {
                long long c = tgt[dir ? (L - 1 - j) : j];
                lab[i] = (int)(c < 0 ? 0 : (c >= p.V ? p.V - 1 : c));
            }
        }
        const long long fstep_b = fstep * ES;
        unsigned char* myrows = rowbuf + (size_t)pw * NG * G * row_stride_bytes;
        const uint32_t myrows_s = ws_smem_u32(myrows);
        unsigned long long* mybar = mbar + pw * NG;
        const int ngroups_all = (Tb - 1 + G - 1) / G;
        const int mygroups = (ngroups_all - pw + 1) / 2;       // groups pw, pw+2, ... < ngroups_all
        const uint32_t off_b = (uint32_t)p.blank * ES;
        uint32_t off_l[KL];
        float ninf_l[KL], one_b[KL];                            // validity as arithmetic (no predicate re-materialisation)
#pragma unroll
        for (int i = 0; i < KL; ++i) {
            off_l[i] = (uint32_t)lab[i] * ES;
            ninf_l[i] = (2 * i + 1 < nvalid) ? 0.f : AVCTC_NEG_INF;
            one_b[i] = (2 * i < nvalid) ? 1.f : 0.f;
        }
        const float ninf_b = (nvalid > 0) ? 0.f : AVCTC_NEG_INF;
        auto lds_val = [&](uint32_t addr) -> float {
            if constexpr (ES == 4) {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
                return v;
            } else {
                unsigned short h;
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(addr));
                return __uint_as_float((unsigned)h << 16);
            }
        };
        int cons = 1;
        int sl = 0;
        uint32_t par = 0;
        bool risky = false;
        uintptr_t src_c = reinterpret_cast<uintptr_t>(row0) + (uintptr_t)(fstep_b * (1 + pw * G));
        int f = 1 + pw * G;
#pragma unroll 1
        for (int k = 0; k < mygroups; ++k) {
            ws_mbar_wait(ws_smem_u32(&mybar[sl]), par);
            uint32_t rbase = myrows_s + (uint32_t)(sl * G * row_stride_bytes);
#pragma unroll
            for (int j = 0; j < G; ++j) {
                if (f < Tb) {
                    const uint32_t rb = rbase + (uint32_t)(src_c & 15);
                    float xb = lds_val(rb + off_b) + ninf_b;
                    float xl[KL];
                    float mx = xb, mn = (nvalid > 0) ? xb : CUDART_INF_F;
#pragma unroll
                    for (int i = 0; i < KL; ++i) {
                        const float raw = lds_val(rb + off_l[i]);
                        xl[i] = raw + ninf_l[i];
                        mx = fmaxf(mx, xl[i]);
                        mn = fminf(mn, raw - ninf_l[i]);          // invalid states contribute +inf
                    }
                    // p of some class would leave fp32's safe range: raise the guard flag NOW, before this frame is handed to
                    // the recurrence -- the sample cannot be reported complete (done[b]) with the store still pending, and
                    // the gradient pass / the guarded twins read the flag long after that
                    if (mx - mn > kLinSafeRange && !risky) { risky = true; if (p.flag) *p.flag = 1; }
                    const float ms = mx * AVCTC_LOG2E;
                    const float Ef = (ms > -1.0e29f) ? floorf(ms) : 0.f;
                    const float pblank = ex2_approx(fmaf(xb, AVCTC_LOG2E, -Ef));
                    float pay[PW];
#pragma unroll
                    for (int i = 0; i < KL; ++i) {
                        pay[2 * i] = pblank * one_b[i];
                        pay[2 * i + 1] = ex2_approx(fmaf(xl[i], AVCTC_LOG2E, -Ef));
                    }
                    pay[K] = __int_as_float((int)Ef);
#pragma unroll
                    for (int i = K + 1; i < PW; ++i) pay[i] = 0.f;
                    if (f - cons >= RD) cons = ws_wait_ge(&prog[64 + lane], f - RD + 1);
                    float4* dst = reinterpret_cast<float4*>(pmine + (f & (RD - 1)) * 32 * PW);
#pragma unroll
                    for (int i = 0; i < PW / 4; ++i)
                        dst[i] = make_float4(pay[4 * i], pay[4 * i + 1], pay[4 * i + 2], pay[4 * i + 3]);
                    asm volatile("" ::: "memory");
                    st_volatile_shared_s32(&prog[pw * 32 + lane], f + 1);
                }
                ++f;
                src_c += (uintptr_t)fstep_b;
                rbase += (uint32_t)row_stride_bytes;
            }
            f += G;                                   // skip the other producer's group
            src_c += (uintptr_t)(fstep_b * G);
            __syncwarp();                             // every lane is done reading this group's rows
            if (lane == 0) ws_mbar_arrive(ws_smem_u32(&mbar_empty[pw * NG + sl]));   // the DMA warp may refill the slot
            if (++sl == NG) { sl = 0; par ^= 1u; }
        }
        return;
    }

    if (warp == 3) {
        // ================================================================ writer
        // Repeated-label chains for the gradient pass (which label positions share a class: deterministic per-class
        // posterior sums).  They depend on the targets only, so they are built here, while the row pipeline fills and
        // this warp has nothing to write yet — not by the recurrence warp behind its last frame, where every sample
        // paid for them (~L^2/32 global loads) on its critical path.
        if (dir == 0 && p.chain) {
            for (int j = lane; j < L; j += 32) {
                const long long c = tgt[j];
                tgt_s[j] = (int)(c < 0 ? 0 : (c >= p.V ? p.V - 1 : c));
            }
            __syncwarp();
            for (int j = lane; j < L; j += 32) {
                const int c = tgt_s[j];
                bool first = true;
                for (int k = 0; k < j; ++k) if (tgt_s[k] == c) { first = false; break; }
                int nxt = kChainNone;
                for (int k = j + 1; k < L; ++k) if (tgt_s[k] == c) { nxt = k; break; }
                p.chain[(size_t)b * p.Lpad + j] = nxt | (first ? kChainFirst : 0);
            }
        }
        if (!do_store || Tb < 2) { signal_done(p.done, b, lane, 1); return; }
        const size_t rowi1 = (size_t)b * p.T + (dir ? Tb - 2 : 1);       // scan frame 1
        float* wsp = (dir ? p.beta : p.alpha) + rowi1 * p.S_pad + (size_t)g * K;
        int* cfp = (dir ? p.coff_b : p.coff_a) + rowi1 * p.cw + g;
        long long wstep = dir ? -(long long)p.S_pad : (long long)p.S_pad;
        long long cstep = dir ? -(long long)p.cw : (long long)p.cw;
        asm volatile("" : "+l"(wstep), "+l"(cstep));
        int avail = 1;
#pragma unroll 1
        for (int f = 1; f < Tb; ++f) {
            if (avail <= f) avail = ws_wait_ge(&prog[64 + lane], f + 1);
            const float4* src = reinterpret_cast<const float4*>(smine + (f & (RD - 1)) * 32 * PW);
            float pay[PW];
#pragma unroll
            for (int i = 0; i < PW / 4; ++i) {
                const float4 v = src[i];
                pay[4 * i] = v.x; pay[4 * i + 1] = v.y; pay[4 * i + 2] = v.z; pay[4 * i + 3] = v.w;
            }
            asm volatile("" ::: "memory");
            st_volatile_shared_s32(&prog[96 + lane], f + 1);
            if constexpr (K % 4 == 0) {
#pragma unroll
                for (int i = 0; i < K / 4; ++i)
                    reinterpret_cast<float4*>(wsp)[i] = make_float4(pay[4 * i], pay[4 * i + 1], pay[4 * i + 2], pay[4 * i + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < K / 2; ++i) reinterpret_cast<float2*>(wsp)[i] = make_float2(pay[2 * i], pay[2 * i + 1]);
            }
            *cfp = __float_as_int(pay[K]);
            wsp += wstep; cfp += cstep;
        }
        signal_done(p.done, b, lane, 1);          // this direction's rows 1 .. Tb-1 are in memory
        if (p.stamp && p.flag && lane == 0) stamp_max(p.flag, 1);
        return;
    }

    // ==================================================================== recurrence (warp 2)
    const unsigned long long tm0 = p.stamp ? global_timer_ns() : 0ull;
    float skipf[KL];
#pragma unroll
    for (int i = 0; i < KL; ++i) {
        const int j = g * KL + i;
        skipf[i] = 0.f;
        if (j < L && j >= 1) {
            long long c = tgt[dir ? (L - 1 - j) : j];
            c = c < 0 ? 0 : (c >= p.V ? p.V - 1 : c);
            long long cp = tgt[dir ? (L - j) : (j - 1)];
            cp = cp < 0 ? 0 : (cp >= p.V ? p.V - 1 : cp);
            skipf[i] = (cp != c) ? 1.f : 0.f;
        }
    }
    float a[K];
#pragma unroll
    for (int j = 0; j < K; ++j) a[j] = 0.f;
    int C = 0;
    bool has_mass = false;
    if (g == 0) {
        int lab0 = p.blank;
        if (L > 0) {
            long long c = tgt[dir ? (L - 1) : 0];
            lab0 = (int)(c < 0 ? 0 : (c >= p.V ? p.V - 1 : c));
        }
        const float xb = to_float(__ldg(row0 + p.blank)) * AVCTC_LOG2E;
        const float xl = (L > 0) ? to_float(__ldg(row0 + lab0)) * AVCTC_LOG2E : AVCTC_NEG_INF;
        const float mx = fmaxf(xb, xl);
        const float Ef = (mx > -1.0e29f) ? floorf(mx) : 0.f;
        a[0] = ex2_approx(xb - Ef);
        a[1] = ex2_approx(xl - Ef);
        C = (int)Ef;
        has_mass = fmaxf(a[0], a[1]) > 0.f;
        if (L > 0 && p.flag && fabsf(xb - xl) > kLinSafeRange * AVCTC_LOG2E) *p.flag = 1;
    }
    if (do_store) {     // frame 0 goes straight to the workspace
        const size_t rowi0 = (size_t)b * p.T + (dir ? Tb - 1 : 0);
        float* w0 = (dir ? p.beta : p.alpha) + rowi0 * p.S_pad + (size_t)g * K;
#pragma unroll
        for (int j = 0; j < K; ++j) w0[j] = a[j];
        ((dir ? p.coff_b : p.coff_a) + rowi0 * p.cw)[g] = C;
    }
    const bool not_lane0 = lane > 0;
    int availP[2] = {1, 1};
    int doneW = 1;
    // one frame of the chain; RENORM frames bring the lane maximum back into [1,2) (exact power of two), the frames
    // in between only accumulate the emission exponent — two frames cannot move a lane by more than fp32's range
    // unless a class the mass sits on is > e^-40 below the best class of its lane twice in a row.
    float nx_up_a = __shfl_up_sync(kFullMask, a[K - 1], 1);
    int nx_up_C = __shfl_up_sync(kFullMask, C, 1);
    // `ahead` = frames (this one included) whose ring slots are checked by this call: the full-group loop polls the
    // producer and the writer once per group of G frames (a branch costs ~8 ALU slots on this lone warp), the tail
    // polls every frame.  ahead = 0: no poll.  Neither wait can deadlock: the producer of group [tau, tau+G) needs the
    // recurrence at tau+G-RD <= tau, the writer can reach tau.
    auto frame = [&](const int tau, const int slot, const int owner, const bool renorm, const int ahead) {   // slot = tau & (RD-1), static
        const float up_a = nx_up_a;          // shuffled by the previous frame as soon as its top state was known
        const int up_C = nx_up_C;
        if (ahead > 0) {
            if (availP[owner] < tau + ahead) availP[owner] = ws_wait_ge(&prog[owner * 32 + lane], tau + ahead);
            if (do_store && tau + ahead - doneW > RD) doneW = ws_wait_ge(&prog[96 + lane], tau + ahead - RD);
        }
        const int roff = slot * 32 * PW;
        float pr[PW];
        {
            const float4* src = reinterpret_cast<const float4*>(pmine + roff);
#pragma unroll
            for (int i = 0; i < PW / 4; ++i) {
                const float4 v = src[i];
                pr[4 * i] = v.x; pr[4 * i + 1] = v.y; pr[4 * i + 2] = v.z; pr[4 * i + 3] = v.w;
            }
        }
        const int Ecur = __float_as_int(pr[K]);
        const bool has_up = not_lane0 && (up_a > 0.f);
        int d = has_up ? up_C - C : 0;
        if (!has_mass && has_up) { C = up_C; d = 0; }
        if (d > 40) {
            const int sh = d - 40;
            const float k = (sh < 126) ? __int_as_float((127 - sh) << 23) : 0.f;
#pragma unroll
            for (int j = 0; j < K; ++j) a[j] *= k;
            C += sh; d = 40;
        }
        const float sc = (has_up && d > -126) ? __int_as_float((127 + d) << 23) : 0.f;
        const float prev = up_a * sc;
        float nw[K];
#pragma unroll
        for (int i = 0; i < KL; ++i) {
            const float below = (i == 0) ? prev : a[2 * i - 1];
            nw[2 * i] = (a[2 * i] + below) * pr[2 * i];
            nw[2 * i + 1] = fmaf(skipf[i], below, a[2 * i + 1] + a[2 * i]) * pr[2 * i + 1];
        }
        // the neighbour lane needs (top state, exponent) as a consistent pair, not the renormalised one: send the raw
        // pair now, so that the shuffle latency overlaps the renormalisation and the ring store below
        nx_up_a = __shfl_up_sync(kFullMask, nw[K - 1], 1);
        nx_up_C = __shfl_up_sync(kFullMask, C + Ecur, 1);
        if (renorm) {
            float mm = nw[0];
#pragma unroll
            for (int j = 1; j < K; ++j) mm = fmaxf(mm, nw[j]);
            has_mass = mm > 0.f;
            const int e = has_mass ? ((__float_as_int(mm) >> 23) - 127) : 0;
            const float sn = __int_as_float((127 - e) << 23);
#pragma unroll
            for (int j = 0; j < K; ++j) a[j] = nw[j] * sn;
            C += Ecur + e;
        } else {
#pragma unroll
            for (int j = 0; j < K; ++j) a[j] = nw[j];
            has_mass = has_mass || has_up;
            C += Ecur;
        }
        if (do_store) {
            float pay[PW];
#pragma unroll
            for (int j = 0; j < K; ++j) pay[j] = a[j];
            pay[K] = __int_as_float(C);
#pragma unroll
            for (int i = K + 1; i < PW; ++i) pay[i] = 0.f;
            float4* dst = reinterpret_cast<float4*>(smine + roff);
#pragma unroll
            for (int i = 0; i < PW / 4; ++i) dst[i] = make_float4(pay[4 * i], pay[4 * i + 1], pay[4 * i + 2], pay[4 * i + 3]);
        }
        asm volatile("" ::: "memory");
        st_volatile_shared_s32(&prog[64 + lane], tau + 1);
    };
    const unsigned long long tm1 = p.stamp ? global_timer_ns() : 0ull;
    int base = 1;
#pragma unroll 1
    for (; base + 2 * G <= Tb; base += 2 * G) {          // full groups: no bounds checks on the chain
#pragma unroll
        for (int u = 0; u < 2 * G; ++u) frame(base + u, (1 + u) & (RD - 1), (u / G) & 1, (u & 1) == 1, (u % G == 0) ? G : 0);
    }
#pragma unroll
    for (int u = 0; u < 2 * G; ++u) {
        const int tau = base + u;
        if (tau < Tb) frame(tau, (1 + u) & (RD - 1), (u / G) & 1, (u & 1) == 1, 1);
    }
    const unsigned long long tm2 = p.stamp ? global_timer_ns() : 0ull;
    if (dir == 0) {      // only the last two lattice states enter the likelihood: the fp64 logarithm runs for them alone
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int st = g * K + j;
            if (st == S - 1 || st == S - 2) {
                const double tv = (a[j] > 0.f) ? log2((double)a[j]) + (double)C : -(double)CUDART_INF_F;
                fin[st == S - 1 ? 0 : 1] = tv;
            }
        }
        __syncwarp();
        if (lane == 0) {
            const double x = fin[0], y = fin[1];
            const double m = fmax(x, y), n = fmin(x, y);
            double ll2;
            if (m == -(double)CUDART_INF_F) ll2 = m;
            else ll2 = m + log2(1.0 + exp2(n - m));
            p.nll[b] = (float)(-ll2 * AVCTC_LN2_D);
            if (p.nll2) p.nll2[b] = -ll2;
        }
    }
    signal_done(p.done, b, lane, 1);              // frame 0 and nll are in memory (the writer warp signals rows + chains)
    if (p.stamp && p.flag && lane == 0) {
        stamp_max(p.flag, 1);
        const unsigned long long tm3 = global_timer_ns();
        stamp_dur(p.flag, 4, tm1 - tm0);                       // prologue of the recurrence warp
        stamp_dur(p.flag, 5, tm3 - tm2);                       // epilogue (likelihood, fence, counter)
        stamp_dur(p.flag, 6, (tm2 - tm1) * 1000ull / (unsigned long long)(Tb > 1 ? Tb - 1 : 1));   // ps per frame
        stamp_dur(p.flag, 7, tm2 - tm1);                       // longest frame loop
        atomicMax(p.flag + 60, *reinterpret_cast<volatile int*>(p.flag + 33));      // gradient tickets drawn when this scan CTA ended
        atomicMax(p.flag + 61, *reinterpret_cast<volatile int*>(p.flag + 35));      // zero-row tickets drawn
    }
}

// ------------------------------------------------------------------------------------------------
// ctc_grad_lin_kernel — gradient pass for the probability-domain workspaces (ctc_scan_lin_kernel).
//
// A warp owns ONE sample and a strided set of its frames, so everything that depends only on the sample (labels,
// repeat chains, lengths, scale, -log2 P split into integer + fraction) is set up once and lives in registers.
// Lane l owns the same K lattice states as scan lane l (alpha is read as K contiguous floats + one exponent).
// Per frame: state weights alpha*beta/(P*p) = mantissa product * 2^(integer exponents + frac - lp*log2e) with the
// exponents summed as integers; class posteriors go into a per-warp shared "delta" row that is all-zero between
// rows; then the log-prob row streams through registers once:  grad = exp(lp)*g - delta*g  with 16-byte loads and
// 16-byte streaming stores (no shared-memory staging of the row).  Algorithmic traffic: one read of log_probs,
// one write of grad (+ alpha/beta once).
template <typename TIn, bool POS>
__device__ __forceinline__ void grad_stream_row(const TIn* __restrict__ lrow, TIn* __restrict__ grow, int V,
                                                const float* __restrict__ delta, int o_out, float scale, float lscale,
                                                int lane) {
    constexpr int kVec = VecTraits<TIn>::kVec;
    const int o_in = (int)((reinterpret_cast<uintptr_t>(lrow) / sizeof(TIn)) & (kVec - 1));
    const int nq = (o_out + V + kVec - 1) / kVec;
    const bool same = (o_in == o_out);
    for (int q = lane; q < nq; q += 32) {
        const int c0 = q * kVec - o_out;
        const bool full = (c0 >= 0) && (c0 + kVec <= V);
        float x[kVec];
        if (full && same) {
            if constexpr (sizeof(TIn) == 4) {
                const float4 v = __ldcs(reinterpret_cast<const float4*>(lrow + c0));
                x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
            } else {
                const uint4 raw = __ldcs(reinterpret_cast<const uint4*>(lrow + c0));
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h[k]); x[2 * k] = f.x; x[2 * k + 1] = f.y; }
            }
        } else {
#pragma unroll
            for (int k = 0; k < kVec; ++k) {
                const int c = c0 + k;
                x[k] = (c >= 0 && c < V) ? to_float(lrow[c]) : 0.f;
            }
        }
        float o[kVec];
#pragma unroll
        for (int k4 = 0; k4 < kVec / 4; ++k4) {
            const float4 d = *reinterpret_cast<const float4*>(delta + q * kVec + 4 * k4);
            const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float e = ex2_approx(fmaf(x[4 * k4 + k], AVCTC_LOG2E, lscale));   // |scale| * exp(lp)
                o[4 * k4 + k] = POS ? fmaf(dd[k], -scale, e) : -fmaf(dd[k], scale, e);  // scale < 0: -( |s|e + d s )
            }
        }
        if (full) {
            if constexpr (sizeof(TIn) == 4) {
                __stcs(reinterpret_cast<float4*>(grow + c0), make_float4(o[0], o[1], o[2], o[3]));
            } else {
                uint4 raw;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
                for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(o[2 * k], o[2 * k + 1]);
                __stcs(reinterpret_cast<uint4*>(grow + c0), raw);
            }
        } else {
#pragma unroll
            for (int k = 0; k < kVec; ++k) {
                const int c = c0 + k;
                if (c >= 0 && c < V) {
                    if constexpr (sizeof(TIn) == 4) grow[c] = o[k];
                    else grow[c] = __float2bfloat16(o[k]);
                }
            }
        }
    }
}

// Stage one fp32 row for grad_stream_row_smem: the 16-byte chunks that COVER the row are copied whole (cp.async.cg), so
// the loop is one predicate + one copy per 512 bytes; buf[o + c] = row[c] with o = the row's misalignment in elements.
// The up to 12 bytes before / after the row that come along lie in the same 16-byte aligned chunk as row bytes, hence in
// the same page: always readable (they are never used).
__device__ __forceinline__ int stage_row_chunks(const float* __restrict__ grow, int V, float* __restrict__ buf, int lane) {
    const int o = (int)((reinterpret_cast<uintptr_t>(grow) >> 2) & 3);
    const int nq = (o + V + 3) >> 2;
    const char* g = reinterpret_cast<const char*>(grow - o) + lane * 16;
    unsigned sa = (unsigned)__cvta_generic_to_shared(buf) + lane * 16;
    for (int q = lane; q < nq; q += 32, g += 512, sa += 512)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g));
    cp_async_commit();
    return o;
}

// Same arithmetic with the log-prob row already in shared memory (row_s[c] = class c): the global loads of the row
// were issued all at once by cp.async (stage_row) instead of one 16-byte load per lane per loop iteration, so a warp
// has its whole 3.2 KB row in flight, not 512 bytes of it.
template <bool POS>
__device__ __forceinline__ void grad_stream_row_smem(const float* __restrict__ row_s, int o_in, float* __restrict__ grow,
                                                     int V, const float* __restrict__ delta, int o_out, float scale,
                                                     float lscale, int lane) {
    // 16-byte chunk q of the OUTPUT row covers classes c0 = 4q - o_out .. c0 + 3.  Chunks [q_lo, q_hi) are complete
    // (no bounds predicates in the loop); the at most two ragged ones at the ends are left to one lane each.
    const int q_lo = (o_out + 3) >> 2, q_hi = (o_out + V) >> 2;
    auto one = [&](float x, float d) -> float {
        const float e = ex2_approx(fmaf(x, AVCTC_LOG2E, lscale));          // |scale| * exp(lp)
        return POS ? fmaf(d, -scale, e) : -fmaf(d, scale, e);
    };
    if (o_in == o_out) {                         // row_s + c0 is 16-byte aligned whenever grow + c0 is
        for (int q = q_lo + lane; q < q_hi; q += 32) {
            const int c0 = q * 4 - o_out;
            const float4 v = *reinterpret_cast<const float4*>(row_s + c0);
            const float4 d = *reinterpret_cast<const float4*>(delta + q * 4);
            __stcs(reinterpret_cast<float4*>(grow + c0), make_float4(one(v.x, d.x), one(v.y, d.y), one(v.z, d.z), one(v.w, d.w)));
        }
    } else {
        for (int q = q_lo + lane; q < q_hi; q += 32) {
            const int c0 = q * 4 - o_out;
            const float4 d = *reinterpret_cast<const float4*>(delta + q * 4);
            __stcs(reinterpret_cast<float4*>(grow + c0),
                   make_float4(one(row_s[c0], d.x), one(row_s[c0 + 1], d.y), one(row_s[c0 + 2], d.z), one(row_s[c0 + 3], d.w)));
        }
    }
    if (lane == 0 && o_out > 0) {                // ragged head: classes 0 .. 3 - o_out of chunk 0
        for (int c = 0; c < 4 - o_out && c < V; ++c) grow[c] = one(row_s[c], delta[o_out + c]);
    }
    if (lane == 1 && ((o_out + V) & 3) && q_hi >= q_lo) {      // ragged tail: chunk q_hi
        for (int c = max(q_hi * 4 - o_out, (o_out > 0) ? 4 - o_out : 0); c < V; ++c) grow[c] = one(row_s[c], delta[o_out + c]);
    }
}

constexpr int kZeroChunk = 4;          // early mode: rows beyond the input length are handed out four at a time
constexpr int kSpreadMaxB = 256;      // batch sizes whose completion order is sorted inside the gradient kernel

// (fp32: the staged row makes shared memory, not registers, the occupancy limit: 3 CTAs per SM, 85 registers)
template <int K, typename TIn>
__global__ void __launch_bounds__(256, (K <= 8) ? (sizeof(TIn) == 4 ? 3 : 4) : 1) ctc_grad_lin_kernel(const GradParams p) {
    constexpr int KL = K / 2;
    constexpr int kVec = VecTraits<TIn>::kVec;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    pdl_launch_dependents();
    const unsigned long long tg0 = p.stamp ? global_timer_ns() : 0ull;
    if (p.stamp && p.flag && threadIdx.x == 0) {      // debug: histogram of CTA entry times, 16 us buckets from the first scan CTA
        const unsigned long long t0 = ~*reinterpret_cast<const volatile unsigned long long*>(reinterpret_cast<const char*>(p.flag) + 64);
        const long long q = (long long)(tg0 - t0) / 16000;     // no scan stamp (single-warp scan): t0 = ~0, q is huge
        const int bk = q < 0 ? 0 : (q > 15 ? 15 : (int)q);
        atomicAdd(const_cast<int*>(p.flag) + 44 + bk, 1);
    }
    // Launched with the PDL attribute, this grid can be resident while the scan before it on the stream is still
    // running (backward enqueued right behind forward).  If every sample's completion counter is already full the
    // kernel takes the ordinary route: griddepcontrol.wait, after which everything before it on the stream (scan,
    // guard flag, whatever produced grad_out) is complete, then a fixed set of warps per sample.  Otherwise the scan
    // is still in flight -- which implies that nothing but this library's own PDL-chained kernels sits between it and
    // this launch, so grad_out predates the scan -- and the grid works "early": samples are taken in the order their
    // scans finish (= order of input length), cut into chunks of consecutive frames that the resident warps draw from
    // a ticket counter, each chunk waiting only for ITS sample's counter.  The bandwidth-bound gradient rows of the
    // short utterances then stream while the latency-bound recurrences of the long ones still run.  (Tickets, not a
    // fixed chunk -> warp map: while the scan holds its shared memory only part of this grid is resident, and a
    // chunk owned by a CTA that cannot start yet would wait for the whole scan.)
    // Control words in the flag block (zeroed by avctc_ctc_forward, re-zeroed by the last warp of this grid):
    // int[32] mode (0 undecided, 1 ordinary, 2 early; the first CTA decides for the grid), int[33] next ticket,
    // int[34] CTAs finished, int[35] next ticket of the zero rows.
    __shared__ int early_s, chunk_s, warps_left_s, zdrawn_s;
    __shared__ int tb_s[kSpreadMaxB], order_s[kSpreadMaxB], start_s[kSpreadMaxB + 1], zstart_s[kSpreadMaxB + 1];
    int* const ctrl = p.done ? const_cast<int*>(p.flag) + 32 : nullptr;
    // A CTA that becomes resident only when the scan CTAs leave starts on a saturated memory system, where every
    // dependent global round trip costs microseconds of the tail: its independent loads (input lengths, the mode word,
    // the zero-row ticket counter) are issued together here, and a mode that is already decided is taken as it is.
    const bool spread_ok = ctrl && p.B <= kSpreadMaxB && p.B <= (int)blockDim.x;      // one thread per sample builds the tables
    int mine = 0;
    if (spread_ok && (int)threadIdx.x < p.B) {
        const long long v = p.input_lengths[threadIdx.x];
        mine = (int)(v < 0 ? 0 : (v > p.T ? p.T : v));
    }
    if (threadIdx.x < 32) {
        int mode = 1;
        if (spread_ok) {
            int seen = 0;
            if (lane < 2) seen = *reinterpret_cast<const volatile int*>(ctrl + (lane ? 3 : 0));
            const int decided = __shfl_sync(kFullMask, seen, 0);
            const int zdrawn = __shfl_sync(kFullMask, seen, 1);
            if (decided) {
                mode = decided;
            } else {
                bool all = true;
                for (int i = lane; i < p.B; i += 32) all = all && (ld_acquire_gpu_s32(p.done + i) >= kDoneTarget);
                all = __all_sync(kFullMask, all);
                if (lane == 0) {
                    const int want = all ? 1 : 2;
                    const int old = atomicCAS(ctrl, 0, want);
                    mode = old ? old : want;
                }
            }
            if (lane == 0) zdrawn_s = zdrawn;
        }
        if (lane == 0) { early_s = (mode == 2) ? 1 : 0; warps_left_s = nwarp; }
    }
    if (spread_ok && (int)threadIdx.x < p.B) tb_s[threadIdx.x] = mine;
    __syncthreads();
    const bool early = early_s != 0;
    const int gw = blockIdx.x * nwarp + warp, Wtot = gridDim.x * nwarp;
    // ---- balanced mapping (set up before any waiting; input lengths are inputs of the forward pass too): samples in
    // order of input length (= the order in which their scans finish), each cut into chunks of consecutive frames;
    // chunk i of that list goes to warp i mod Wtot.  Every warp gets the same number of chunks (utterances of
    // different lengths no longer decide which warps finish last) and meets the samples in the order they complete.
    const bool use_spread = early;
    auto finish = [&]() {            // every warp, exactly once, on its way out
        if (p.stamp && p.flag && lane == 0) stamp_max(const_cast<int*>(p.flag), 2);
        // Last warp of the CTA reports the CTA; the last CTA of the grid re-arms the control words for the next launch.
        // No fence needed: a warp's ticket draws have returned (it compared them) before it gets here.
        if (ctrl && lane == 0 && atomicSub(&warps_left_s, 1) == 1) {
            if (atomicAdd(ctrl + 2, 1) == (int)gridDim.x - 1) {
                ctrl[1] = 0; ctrl[2] = 0; ctrl[3] = 0;
                __threadfence();
                ctrl[0] = 0;
            }
        }
    };
    if (use_spread) {
        const int tid = threadIdx.x;
        if (tid < p.B) {
            int R = 0;
            for (int j = 0; j < p.B; ++j) R += tb_s[j];
            int c = R / (Wtot * 4);          // (measured: 3-frame chunks at config 2; 6 -> 219 us, 1 -> 208 us, 3 -> 188 us)
            c = c < 1 ? 1 : (c > 16 ? 16 : c);
            order_s[tid] = (mine + c - 1) / c;               // chunks of this sample (order_s is scratch until below)
            if (tid == 0) chunk_s = c;
        }
        __syncthreads();
        int rank = 0, start = 0, nch = 0, zs = 0;
        if (tid < p.B) {
            nch = order_s[tid];
            for (int j = 0; j < p.B; ++j) {
                const int o = tb_s[j];
                const bool before = (o < mine) || (o == mine && j < tid);
                rank += before ? 1 : 0;
                start += before ? order_s[j] : 0;
                zs += (j < tid) ? (p.T - o + kZeroChunk - 1) / kZeroChunk : 0;
            }
            zstart_s[tid] = zs;
            if (tid == p.B - 1) zstart_s[p.B] = zs + (p.T - mine + kZeroChunk - 1) / kZeroChunk;
        }
        __syncthreads();
        if (tid < p.B) {
            order_s[rank] = tid;
            start_s[rank] = start;
            if (rank == p.B - 1) start_s[p.B] = start + nch;
        }
        __syncthreads();
    }
    if (!early) {
        pdl_wait();
        if (p.flag && *p.flag != p.run_if) { finish(); return; }
    }
    float* delta = smem + (size_t)warp * (p.row_floats + p.w_floats);   // class posteriors; all-zero between rows
    float* wbuf = delta + p.row_floats;                                  // label-state weights of the current row
    const bool stage = (sizeof(TIn) == 4) && p.stage;
    if (stage) {         // per-warp layout with a staged row: delta | wbuf | rowbuf
        delta = smem + (size_t)warp * (2 * p.row_floats + p.w_floats);
        wbuf = delta + p.row_floats;
    }
    float* const rowbuf = wbuf + p.w_floats;                             // [row_floats] the current log-prob row (stage)
    for (int i = lane; i < p.row_floats; i += 32) delta[i] = 0.f;
    __syncwarp();
    const int WS = max(1, Wtot / p.B);                                   // warps per sample (static mapping)
    const long long nitems = (long long)p.B * WS;
    TIn* grad = reinterpret_cast<TIn*>(p.grad);
    const TIn* lpbase = reinterpret_cast<const TIn*>(p.lp);
    // lane l prefetches the line holding byte 128*l of the row: every address stays inside the row (and the tensor)
    const int pf_lines = p.prefetch ? min(32, (int)((p.V * sizeof(TIn) + 127) / 128)) : 0;
    const int pf_dist = max(1, p.prefetch & 3);
    const bool pf_ab = (p.prefetch & 4) != 0;
    const int ab_lines = min(16, (p.S_pad * 4 + 127) / 128);

    auto clamp_tb = [&](int bb) -> int {
        const long long v = p.input_lengths[bb];
        return (int)(v < 0 ? 0 : (v > p.T ? p.T : v));
    };
    auto zero_row = [&](TIn* grow) {
        const int o_out = (int)((reinterpret_cast<uintptr_t>(grow) / sizeof(TIn)) & (kVec - 1));
        const int nq = (o_out + p.V + kVec - 1) / kVec;
        for (int q = lane; q < nq; q += 32) {
            const int c0 = q * kVec - o_out;
            if (c0 >= 0 && c0 + kVec <= p.V) __stcs(reinterpret_cast<uint4*>(grow + c0), make_uint4(0, 0, 0, 0));
            else
                for (int k = 0; k < kVec; ++k)
                    if (c0 + k >= 0 && c0 + k < p.V) {
                        if constexpr (sizeof(TIn) == 4) grow[c0 + k] = 0.f;
                        else grow[c0 + k] = __float2bfloat16(0.f);
                    }
        }
    };

    // ---- rows beyond the input length are zero whatever the scan finds: written first, before any waiting (early
    // mode: drawn from their own ticket counter, so the CTAs that are resident from the start write all of them while
    // no sample is complete yet)
    if (early) {
        const int nz = zstart_s[p.B];
        while (zdrawn_s < nz) {                     // (every zero chunk already drawn when this CTA arrived: nothing to do)
            int z = 0;
            if (lane == 0) z = atomicAdd(ctrl + 3, 1);
            z = __shfl_sync(kFullMask, z, 0);
            if (z >= nz) break;
            int lo = 0, hi = p.B - 1;                       // last sample whose first zero chunk is <= z
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (zstart_s[mid] <= z) lo = mid; else hi = mid - 1;
            }
            const int t0 = tb_s[lo] + (z - zstart_s[lo]) * kZeroChunk;
            for (int t = t0; t < min(t0 + kZeroChunk, p.T); ++t) zero_row(grad + ((size_t)t * p.B + lo) * p.V);
        }
    } else {
        for (long long item = gw; item < nitems; item += Wtot) {
            const int b = (int)(item % p.B), r = (int)(item / p.B);
            const int Tb = clamp_tb(b);
            int t = r + ((max(Tb - r, 0) + WS - 1) / WS) * WS;           // first row of this warp with t >= Tb
            for (; t < p.T; t += WS) zero_row(grad + ((size_t)t * p.B + b) * p.V);
        }
    }

    // ---- rows t0, t0+step, ... < t1 of sample b
    auto process = [&](const int b, const int t0, const int t1, const int step) {
        if (t0 >= t1) return;
        if (early) {
            if (lane == 0) {
                int spins = 0;
                while (ld_acquire_gpu_s32(p.done + b) < kDoneTarget) {
                    __nanosleep(500);
                    if (++spins > (1 << 23)) __trap();      // never hang the GPU on a protocol bug
                }
            }
            __syncwarp();
            __threadfence();
            if (p.stamp && lane == 0) stamp_min(const_cast<int*>(p.flag), 3);
        }
        long long tll = p.target_lengths[b];
        const int L = (int)(tll < 0 ? 0 : (tll > p.Lmax ? p.Lmax : tll));
        const int S = 2 * L + 1;
        const float nllb = p.nll[b];
        const bool zero_all = (p.zero_infinity && nllb == CUDART_INF_F);
        if (zero_all) {
            for (int t = t0; t < t1; t += step) zero_row(grad + ((size_t)t * p.B + b) * p.V);
            return;
        }
        const double nll2 = p.nll2[b];                                   // -log2 P
        const bool bad = !(fabs(nll2) < 1.0e300);                        // infeasible without zero_infinity: NaN rows
        const double nfl = bad ? 0.0 : floor(nll2);
        const int nll2_i = (int)nfl;
        const float nll2_f = bad ? 0.f : (float)(nll2 - nfl);
        float scale = p.grad_out[(int64_t)b * p.grad_out_stride];
        if (p.reduction == AVCTC_REDUCE_MEAN) scale /= ((float)p.B * (float)(L > 1 ? L : 1));
        if (bad) scale = __int_as_float(0x7fc00000);
        const float lscale = lg2_approx(fabsf(scale));
        const int64_t* tgt = p.targets + (p.target_offsets ? p.target_offsets[b] : (int64_t)b * p.target_stride);
        const int* chain = p.chain + (size_t)b * p.Lpad;
        int cls[KL], chn[KL];
#pragma unroll
        for (int i = 0; i < KL; ++i) {
            const int pos = lane * KL + i;
            cls[i] = p.blank; chn[i] = 0;
            if (pos < L) {
                long long c = tgt[pos];
                cls[i] = (int)(c < 0 ? 0 : (c >= p.V ? p.V - 1 : c));
                chn[i] = chain[pos];
            }
        }
        const int s0 = lane * K;
        const int sm0 = S - 1 - s0;              // beta (mirrored) index of state s0; state s0+j -> sm0 - j
        // Per-sample set-up of the row loop (an instruction diet: the loop below used to spend ~45 % of its 1 330
        // instructions per row on predicated loads with 64-bit index arithmetic and a branch per state).  States beyond
        // S have alpha == 0 in the workspace (the scan multiplies them by a zero emission), so their weight is zero
        // whatever beta / exponent they are paired with: every load is unconditional, with the beta index clamped
        // into the row.  The K mirrored states of a lane fall into at most two lanes of the beta scan: two exponent
        // loads and a per-state select.
        // (beta index: sm0 - j with NO clamp for a lane that holds at least one state -- at most K - 1 floats before the
        // row, i.e. the previous row or, for the first row, the tail of the alpha array that precedes beta in the
        // workspace; a lane without states reads the first K floats of the row)
        const int sm0u = sm0 >= 0 ? sm0 : K - 1;
        unsigned lo_mask = 0;
        const int cb_hi_i = sm0u / K, cb_lo_i = max(sm0u - (K - 1), 0) / K;
#pragma unroll
        for (int j = 0; j < K; ++j)
            if (max(sm0u - j, 0) / K != cb_hi_i) lo_mask |= 1u << j;
        // running pointers of the rows t0, t0 + step, ...
        const TIn* lrow = lpbase + (int64_t)t0 * p.stride_t + (int64_t)b * p.stride_b;
        TIn* grow = grad + ((size_t)t0 * p.B + b) * p.V;
        const size_t rowi0 = (size_t)b * p.T + t0;
        const float* ap = p.alpha + rowi0 * p.S_pad + s0;
        const float* bp = p.beta + rowi0 * p.S_pad + sm0u;                 // state s0 + j  ->  bp[-j]
        const int* cap = p.coff_a + rowi0 * p.cw + lane;
        const int* cbp = p.coff_b + rowi0 * p.cw;
        const int64_t lstep = (int64_t)step * p.stride_t;
        const size_t gstep = (size_t)step * p.B * p.V;
        const int abstep = step * p.S_pad, cstep = step * p.cw;
        const int base_e = nll2_i - 254;

        for (int t = t0; t < t1; t += step, lrow += lstep, grow += gstep, ap += abstep, bp += abstep, cap += cstep, cbp += cstep) {
            const int o_out = (int)((reinterpret_cast<uintptr_t>(grow) / sizeof(TIn)) & (kVec - 1));
            int o_in_s = 0;
            if constexpr (sizeof(TIn) == 4) {
                if (stage) o_in_s = stage_row_chunks(reinterpret_cast<const float*>(lrow), p.V, rowbuf, lane);   // cp.async, one group
            }
            if (pf_lines > 0 && t + pf_dist * step < t1) {    // a later row of this warp: pull its lines into L2 now
                if (lane < pf_lines) {
                    const char* nx = reinterpret_cast<const char*>(lrow + (int64_t)pf_dist * lstep) + lane * 128;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
                }
                if (pf_ab && lane < 2 * ab_lines) {
                    const char* nx = reinterpret_cast<const char*>((lane < ab_lines ? ap - s0 : bp - sm0u) + pf_dist * abstep) +
                                     (lane < ab_lines ? lane : lane - ab_lines) * 128;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
                }
            }
            float av[K], bv[K];
            int e_hi = 0, e_lo = 0;
#pragma unroll
            for (int j = 0; j < K; ++j) { av[j] = 0.f; bv[j] = 0.f; }
            if (s0 < S) {                // lanes beyond the lattice load nothing (on average half of every alpha / beta row)
#pragma unroll
                for (int j = 0; j < K / 2; ++j) {
                    const float2 v = reinterpret_cast<const float2*>(ap)[j];
                    av[2 * j] = v.x; av[2 * j + 1] = v.y;
                }
#pragma unroll
                for (int j = 0; j < K; ++j) bv[j] = bp[-j];
                const int ca = *cap;
                e_hi = cbp[cb_hi_i] + ca + base_e; e_lo = cbp[cb_lo_i] + ca + base_e;
            }
            float xb, xl[KL];
            if (stage) {                 // the row has landed (its latency overlapped the alpha/beta loads above)
                cp_async_wait<0>();
                __syncwarp();
                const float* row_s = rowbuf + o_in_s;
                xb = row_s[p.blank] * AVCTC_LOG2E;
#pragma unroll
                for (int i = 0; i < KL; ++i) xl[i] = row_s[cls[i]] * AVCTC_LOG2E;
            } else {
                xb = to_float(__ldg(lrow + p.blank)) * AVCTC_LOG2E;
#pragma unroll
                for (int i = 0; i < KL; ++i) xl[i] = to_float(__ldg(lrow + cls[i])) * AVCTC_LOG2E;
            }
            const float fb = nll2_f - xb;
            float w[K];
#pragma unroll
            for (int j = 0; j < K; ++j) {      // alpha*beta/(P*p): mantissa product, integer exponents, no branch
                const int ia = __float_as_int(av[j]), ib = __float_as_int(bv[j]);
                const bool ok = (av[j] >= 1.17549435e-38f) && (bv[j] >= 1.17549435e-38f);
                const float mant = __int_as_float((ia & 0x007fffff) | 0x3f800000) *
                                   __int_as_float((ib & 0x007fffff) | 0x3f800000);
                const int ei = (ia >> 23) + (ib >> 23) + (((lo_mask >> j) & 1u) ? e_lo : e_hi);
                const float x = (float)ei + ((j & 1) ? (nll2_f - xl[j >> 1]) : fb);
                const float v = mant * ex2_approx(x);
                w[j] = ok ? v : 0.f;
            }
            float pbs = 0.f;
#pragma unroll
            for (int i = 0; i < KL; ++i) pbs += w[2 * i];
            pbs = warp_sum(pbs);
#pragma unroll
            for (int i = 0; i < KL; ++i) {
                const int pos = lane * KL + i;
                if (pos < L) wbuf[pos] = w[2 * i + 1];
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < KL; ++i) {
                if (chn[i] & kChainFirst) {
                    float acc = w[2 * i + 1];
                    int k = chn[i] & kChainNone;
                    while (k != kChainNone) {
                        acc += wbuf[k];
                        k = chain[k] & kChainNone;
                    }
                    delta[o_out + cls[i]] = acc;
                }
            }
            if (lane == 0) delta[o_out + p.blank] = pbs;
            __syncwarp();
            bool streamed = false;
            if constexpr (sizeof(TIn) == 4) {
                if (stage) {
                    float* g32 = reinterpret_cast<float*>(grow);
                    if (scale > 0.f) grad_stream_row_smem<true>(rowbuf + o_in_s, o_in_s, g32, p.V, delta, o_out, scale, lscale, lane);
                    else grad_stream_row_smem<false>(rowbuf + o_in_s, o_in_s, g32, p.V, delta, o_out, scale, lscale, lane);
                    streamed = true;
                }
            }
            if (!streamed) {
                if (scale > 0.f) grad_stream_row<TIn, true>(lrow, grow, p.V, delta, o_out, scale, lscale, lane);
                else grad_stream_row<TIn, false>(lrow, grow, p.V, delta, o_out, scale, lscale, lane);
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < KL; ++i)
                if (chn[i] & kChainFirst) delta[o_out + cls[i]] = 0.f;
            if (lane == 0) delta[o_out + p.blank] = 0.f;
            __syncwarp();
        }
    };

    if (p.stamp && p.flag && lane == 0) {       // debug: set-up + zero rows of this warp, when the last warp started on its
        int* const f = const_cast<int*>(p.flag);                                        // rows, when the last CTA entered
        stamp_dur(f, 10, global_timer_ns() - tg0);
        stamp_max(f, 11);
        atomicMax(reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(f) + 64) + 12, tg0);
    }
    if (use_spread) {
        const int c = chunk_s, ntasks = start_s[p.B];
        auto draw = [&]() -> int {
            int v = 0;
            if (lane == 0) v = atomicAdd(ctrl + 1, 1);
            return __shfl_sync(kFullMask, v, 0);
        };
        int task = draw();
        while (task < ntasks) {
            const int nxt = draw();                         // its round trip hides behind this chunk
            auto locate = [&](const int tk, int& bb, int& tt) {
                int lo = 0, hi = p.B - 1;                   // last rank whose first chunk is <= tk
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (start_s[mid] <= tk) lo = mid; else hi = mid - 1;
                }
                bb = order_s[lo];
                tt = (tk - start_s[lo]) * c;
            };
            int b, t0;
            locate(task, b, t0);
            if (nxt < ntasks && pf_lines > 0) {             // first row of the next chunk: log-probs (and alpha/beta) into L2
                int bn, tn;
                locate(nxt, bn, tn);
                if (lane < pf_lines)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(
                        lpbase + (int64_t)tn * p.stride_t + (int64_t)bn * p.stride_b) + lane * 128));
                if (lane < 2 * ab_lines) {
                    const size_t rn = ((size_t)bn * p.T + tn) * p.S_pad;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(
                        (lane < ab_lines ? p.alpha : p.beta) + rn) + (lane < ab_lines ? lane : lane - ab_lines) * 128));
                }
            }
            // guard tripped (so far): the log-domain kernels behind this one redo the whole batch, also when the guard
            // trips later and rows computed here from half-rewritten workspaces are garbage
            if (p.flag && *reinterpret_cast<const volatile int*>(p.flag) != p.run_if) break;
            process(b, t0, min(t0 + c, tb_s[b]), 1);
            task = nxt;
        }
    } else {
        for (long long item = gw; item < nitems; item += Wtot) {
            const int b = (int)(item % p.B), r = (int)(item / p.B);
            process(b, r, clamp_tb(b), WS);
        }
    }
    finish();
}

__global__ void ctc_reduce_kernel(const float* __restrict__ nll, const int64_t* __restrict__ tl, int B,
                                  int reduction, int zero_infinity, float* __restrict__ loss) {
    __shared__ double part[32];
    pdl_launch_dependents();
    pdl_wait();                  // launched with the PDL attribute: the scan kernels before it must have finished
    double acc = 0.0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float v = nll[b];
        if (zero_infinity && v == CUDART_INF_F) v = 0.f;
        if (reduction == AVCTC_REDUCE_NONE) { loss[b] = v; continue; }
        if (reduction == AVCTC_REDUCE_MEAN) {
            long long l = tl[b];
            // ATen divides in fp32: (nll / clamp_min(L,1)).mean()
            acc += (double)(v / (float)(l > 1 ? l : 1));
        } else acc += (double)v;
    }
    if (reduction == AVCTC_REDUCE_NONE) return;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFullMask, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w];
        loss[0] = (float)(reduction == AVCTC_REDUCE_MEAN ? s / (double)B : s);
    }
}

template <int K, typename TIn, bool MULTI>
static int launch_scan_impl(const ScanParams& sp, int ndir, cudaStream_t st) {
    constexpr int D = (K >= 16) ? kScanPrefetch / 4 : (K >= 8) ? kScanPrefetch / 2 : kScanPrefetch;
    dim3 grid(sp.B, ndir), block(32 * sp.W);
    const size_t smem = (size_t)D * (K / 2 + 1) * block.x * sizeof(unsigned);
    static bool configured = false;
    if (smem > 32 * 1024 && !configured) {
        AVCTC_CUDA_RETURN(cudaFuncSetAttribute(ctc_scan_kernel<K, TIn, MULTI>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        configured = true;
    }
    AVCTC_CUDA_RETURN(avctc_launch_pdl(ctc_scan_kernel<K, TIn, MULTI>, grid, block, smem, st, sp));
    return (int)cudaGetLastError();
}
template <int K, typename TIn>
static int launch_scan(const ScanParams& sp, int ndir, cudaStream_t st) {
    if (sp.W > 1) return launch_scan_impl<K, TIn, true>(sp, ndir, st);
    return launch_scan_impl<K, TIn, false>(sp, ndir, st);
}
template <typename TIn>
static int dispatch_scan(const ScanParams& sp, int K, int ndir, cudaStream_t st) {
    switch (K) {
        case 2: return launch_scan<2, TIn>(sp, ndir, st);
        case 4: return launch_scan<4, TIn>(sp, ndir, st);
        case 8: return launch_scan<8, TIn>(sp, ndir, st);
        case 16: return launch_scan<16, TIn>(sp, ndir, st);
    }
    return AVCTC_ERR_UNSUPPORTED;
}

template <int K, typename TIn>
static int launch_scan_lin(const ScanParams& sp, int ndir, cudaStream_t st) {
    constexpr int D = (K >= 12) ? 8 : 16;
    dim3 grid(sp.B, ndir), block(32);
    const size_t smem = (size_t)D * 32 * (K / 2 + 1) * sizeof(unsigned);
    ctc_scan_lin_kernel<K, TIn><<<grid, block, smem, st>>>(sp);
    return (int)cudaGetLastError();
}
template <int K, typename TIn>
static int launch_scan_ws(const ScanParams& sp, int ndir, cudaStream_t st) {
    constexpr int PW = (K + 1 + 3) & ~3;
    constexpr int D = 2 * kWsGroup * kWsGroups;
    dim3 grid(sp.B, ndir), block(kWsThreads);
    const int row_stride = (int)(((size_t)sp.V * sizeof(TIn) + 32 + 127) & ~(size_t)127);
    const size_t smem = (size_t)D * row_stride + sizeof(float) * 2 * kWsRing * 32 * PW + sizeof(int) * 128 +
                        sizeof(unsigned long long) * 4 * kWsGroups;
    if (smem > 200 * 1024) return -1000;     // rows too long for the shared-memory ring: caller falls back
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        AVCTC_CUDA_RETURN(cudaFuncSetAttribute(ctc_scan_ws_kernel<K, TIn, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               200 * 1024));
        AVCTC_CUDA_RETURN(cudaFuncSetAttribute(ctc_scan_ws_kernel<K, TIn, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               200 * 1024));
        configured = 200 * 1024;
    }
    if (sp.store) ctc_scan_ws_kernel<K, TIn, true><<<grid, block, smem, st>>>(sp, row_stride);
    else ctc_scan_ws_kernel<K, TIn, false><<<grid, block, smem, st>>>(sp, row_stride);
    return (int)cudaGetLastError();
}
template <typename TIn>
static int dispatch_scan_lin(const ScanParams& sp, int K, int ndir, cudaStream_t st) {
    if (avctc_tuning_get("ctc_ws", 1) != 0 && sp.stride_t >= 0 && sp.stride_b >= 0) {
        int rc = -1000;
        switch (K) {
            case 2: rc = launch_scan_ws<2, TIn>(sp, ndir, st); break;
            case 4: rc = launch_scan_ws<4, TIn>(sp, ndir, st); break;
            case 6: rc = launch_scan_ws<6, TIn>(sp, ndir, st); break;
            case 8: rc = launch_scan_ws<8, TIn>(sp, ndir, st); break;
            case 12: rc = launch_scan_ws<12, TIn>(sp, ndir, st); break;
            case 16: rc = launch_scan_ws<16, TIn>(sp, ndir, st); break;
        }
        if (rc != -1000) return rc;
    }
    switch (K) {
        case 2: return launch_scan_lin<2, TIn>(sp, ndir, st);
        case 4: return launch_scan_lin<4, TIn>(sp, ndir, st);
        case 6: return launch_scan_lin<6, TIn>(sp, ndir, st);
        case 8: return launch_scan_lin<8, TIn>(sp, ndir, st);
        case 12: return launch_scan_lin<12, TIn>(sp, ndir, st);
        case 16: return launch_scan_lin<16, TIn>(sp, ndir, st);
    }
    return AVCTC_ERR_UNSUPPORTED;
}

static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

template <typename TIn>
static int launch_grad(const GradParams& gp_in, cudaStream_t st) {
    GradParams gp = gp_in;
    const size_t per_warp = (size_t)(2 * gp.row_floats + gp.w_floats) * sizeof(float);
    int warps = 8;
    while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
    if (per_warp * warps > 200 * 1024) return AVCTC_ERR_UNSUPPORTED;
    const size_t smem = per_warp * warps;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        AVCTC_CUDA_RETURN(cudaFuncSetAttribute(ctc_grad_kernel<TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               200 * 1024));
        configured = 200 * 1024;
    }
    int occ = 1;
    AVCTC_CUDA_RETURN(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ctc_grad_kernel<TIn>, warps * 32, smem));
    if (occ < 1) occ = 1;
    const long long rows = (long long)gp.T * gp.B;
    long long blocks = (rows + warps - 1) / warps;
    const long long cap = (long long)num_sms() * occ;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) return AVCTC_OK;
    AVCTC_CUDA_RETURN(avctc_launch_pdl(ctc_grad_kernel<TIn>, dim3((unsigned)blocks), dim3(warps * 32), smem, st, gp));
    return (int)cudaGetLastError();
}

template <int K, typename TIn>
static int launch_grad_lin(const GradParams& gp, cudaStream_t st) {
    const int warps = 8;
    const bool stage = (sizeof(TIn) == 4) && gp.stage;
    const size_t smem = (size_t)((stage ? 2 : 1) * gp.row_floats + gp.w_floats) * sizeof(float) * warps;
    if (smem > 200 * 1024) return AVCTC_ERR_UNSUPPORTED;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        AVCTC_CUDA_RETURN(cudaFuncSetAttribute(ctc_grad_lin_kernel<K, TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               200 * 1024));
        configured = 200 * 1024;
    }
    int occ = 1;
    AVCTC_CUDA_RETURN(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ctc_grad_lin_kernel<K, TIn>, warps * 32, smem));
    if (occ < 1) occ = 1;
    const long long rows = (long long)gp.T * gp.B;
    long long blocks = (rows + warps - 1) / warps;
    const long long cap = (long long)num_sms() * occ;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) return AVCTC_OK;
    AVCTC_CUDA_RETURN(avctc_launch_pdl(ctc_grad_lin_kernel<K, TIn>, dim3((unsigned)blocks), dim3(warps * 32), smem, st, gp));
    return (int)cudaGetLastError();
}
template <typename TIn>
static int dispatch_grad_lin(const GradParams& gp, cudaStream_t st) {
    switch (gp.K) {
        case 2: return launch_grad_lin<2, TIn>(gp, st);
        case 4: return launch_grad_lin<4, TIn>(gp, st);
        case 6: return launch_grad_lin<6, TIn>(gp, st);
        case 8: return launch_grad_lin<8, TIn>(gp, st);
        case 12: return launch_grad_lin<12, TIn>(gp, st);
        case 16: return launch_grad_lin<16, TIn>(gp, st);
    }
    return AVCTC_ERR_UNSUPPORTED;
}

}  // namespace avctc

using namespace avctc;

// grad[t][b][:] *= g[b * gstride] in place; rows whose factor is exactly 1 are not touched (a scalar factor of 1 — the
// gradient of the loss with respect to itself — makes the whole launch a no-op after one 4-byte read per CTA)
template <typename T>
__global__ void __launch_bounds__(256) ctc_scale_grad_kernel(T* __restrict__ grad, long long rows, int B, int V,
                                                             const float* __restrict__ g, long long gstride) {
    if (gstride == 0 && g[0] == 1.f) return;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const float f = g[(row % B) * gstride];
        if (f == 1.f) continue;
        T* r = grad + row * V;
        for (int c = threadIdx.x; c < V; c += blockDim.x) {
            if constexpr (sizeof(T) == 4) r[c] = r[c] * f;
            else r[c] = __float2bfloat16(__bfloat162float(r[c]) * f);
        }
    }
}

extern "C" size_t avctc_ctc_workspace_bytes(int T, int B, int max_target_len) {
    CtcPlan pl;
    if (T < 0 || B < 0 || max_target_len < 0) return 0;
    if (!make_plan(T, B, max_target_len, &pl)) return 0;
    return pl.total;
}

extern "C" int avctc_ctc_forward(const void* log_probs, int dtype, int64_t stride_t, int64_t stride_b,
                                 int T, int B, int V, const int64_t* targets, int64_t target_stride,
                                 const int64_t* target_offsets, const int64_t* input_lengths,
                                 const int64_t* target_lengths, int max_target_len, int blank, int need_grad,
                                 float* nll, void* workspace, size_t workspace_bytes, void* stream) {
    if (T < 0 || B < 0 || V <= 0 || max_target_len < 0 || blank < 0 || blank >= V) return AVCTC_ERR_BAD_ARG;
    if (B == 0) return AVCTC_OK;
    if (!log_probs || !targets || !input_lengths || !target_lengths || !nll) return AVCTC_ERR_BAD_ARG;
    if (dtype != AVCTC_F32 && dtype != AVCTC_BF16) return AVCTC_ERR_BAD_ARG;
    CtcPlan pl;
    if (!make_plan(T, B, max_target_len, &pl)) return AVCTC_ERR_UNSUPPORTED;
    if (need_grad && (!workspace || workspace_bytes < pl.total)) return AVCTC_ERR_WORKSPACE;
    if (need_grad && (reinterpret_cast<uintptr_t>(workspace) & 255)) return AVCTC_ERR_ALIGNMENT;
    ScanParams sp;
    sp.lp = log_probs; sp.stride_t = stride_t; sp.stride_b = stride_b;
    sp.T = T; sp.B = B; sp.V = V;
    sp.targets = targets; sp.target_stride = target_stride; sp.target_offsets = target_offsets;
    sp.input_lengths = input_lengths; sp.target_lengths = target_lengths;
    sp.Lmax = max_target_len; sp.blank = blank; sp.store = need_grad ? 1 : 0;
    sp.nll = nll;
    char* w = reinterpret_cast<char*>(workspace);
    sp.alpha = need_grad ? reinterpret_cast<float*>(w + pl.off_alpha) : nullptr;
    sp.beta = need_grad ? reinterpret_cast<float*>(w + pl.off_beta) : nullptr;
    sp.coff_a = need_grad ? reinterpret_cast<int*>(w + pl.off_coff_a) : nullptr;
    sp.coff_b = need_grad ? reinterpret_cast<int*>(w + pl.off_coff_b) : nullptr;
    sp.nll2 = need_grad ? reinterpret_cast<double*>(w + pl.off_nll2) : nullptr;
    sp.chain = need_grad ? reinterpret_cast<int*>(w + pl.off_chain) : nullptr;
    sp.W = pl.W; sp.S_pad = pl.S_pad; sp.Lpad = pl.Lpad;
    sp.cw = pl.CW; sp.flag = nullptr; sp.run_if = 0; sp.done = nullptr; sp.stamp = avctc_tuning_get("ctc_stamp", 0);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int ndir = need_grad ? 2 : 1;
    // The probability-domain scan needs the device flag of the workspace for its range guard; a forward-only call
    // without workspace (evaluation) goes straight to the log-domain kernel.
    if (pl.linear && need_grad) {
        sp.flag = reinterpret_cast<int*>(w + pl.off_flag);
        sp.done = reinterpret_cast<int*>(w + pl.off_flag + kFlagBytes);
        AVCTC_CUDA_RETURN(cudaMemsetAsync(sp.flag, 0, kFlagBytes + (size_t)B * sizeof(int), st));
        int rc = (dtype == AVCTC_F32) ? dispatch_scan_lin<float>(sp, pl.K, ndir, st)
                                      : dispatch_scan_lin<__nv_bfloat16>(sp, pl.K, ndir, st);
        if (rc) return rc;
        ScanParams sl = sp;                       // conditional fallback: runs only if the guard tripped
        sl.W = pl.Wlog; sl.run_if = 1;
        if (dtype == AVCTC_F32) return dispatch_scan<float>(sl, pl.Klog, ndir, st);
        return dispatch_scan<__nv_bfloat16>(sl, pl.Klog, ndir, st);
    }
    sp.W = pl.Wlog;
    if (dtype == AVCTC_F32) return dispatch_scan<float>(sp, pl.Klog, ndir, st);
    return dispatch_scan<__nv_bfloat16>(sp, pl.Klog, ndir, st);
}

extern "C" int avctc_ctc_reduce(const float* nll, const int64_t* target_lengths, int B, int reduction,
                                int zero_infinity, float* loss, void* stream) {
    if (B < 0 || !loss) return AVCTC_ERR_BAD_ARG;
    if (reduction < AVCTC_REDUCE_NONE || reduction > AVCTC_REDUCE_SUM) return AVCTC_ERR_BAD_ARG;
    if (B > 0 && (!nll || !target_lengths)) return AVCTC_ERR_BAD_ARG;
    return (int)avctc_launch_pdl(ctc_reduce_kernel, dim3(1), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), nll,
                                 target_lengths, B, reduction, zero_infinity, loss);
}

extern "C" int avctc_ctc_backward(const void* log_probs, int dtype, int64_t stride_t, int64_t stride_b,
                                  int T, int B, int V, const int64_t* targets, int64_t target_stride,
                                  const int64_t* target_offsets, const int64_t* input_lengths,
                                  const int64_t* target_lengths, int max_target_len, int blank, int reduction,
                                  int zero_infinity, const float* nll, const float* grad_out,
                                  int64_t grad_out_stride, void* grad, const void* workspace,
                                  size_t workspace_bytes, void* stream) {
    if (T < 0 || B < 0 || V <= 0 || max_target_len < 0 || blank < 0 || blank >= V) return AVCTC_ERR_BAD_ARG;
    if (B == 0 || T == 0) return AVCTC_OK;
    if (!log_probs || !targets || !input_lengths || !target_lengths || !nll || !grad_out || !grad || !workspace)
        return AVCTC_ERR_BAD_ARG;
    if (dtype != AVCTC_F32 && dtype != AVCTC_BF16) return AVCTC_ERR_BAD_ARG;
    if (reduction < AVCTC_REDUCE_NONE || reduction > AVCTC_REDUCE_SUM) return AVCTC_ERR_BAD_ARG;
    CtcPlan pl;
    if (!make_plan(T, B, max_target_len, &pl)) return AVCTC_ERR_UNSUPPORTED;
    if (workspace_bytes < pl.total) return AVCTC_ERR_WORKSPACE;
    GradParams gp;
    gp.lp = log_probs; gp.stride_t = stride_t; gp.stride_b = stride_b;
    gp.T = T; gp.B = B; gp.V = V;
    gp.targets = targets; gp.target_stride = target_stride; gp.target_offsets = target_offsets;
    gp.input_lengths = input_lengths; gp.target_lengths = target_lengths;
    gp.Lmax = max_target_len; gp.blank = blank; gp.reduction = reduction; gp.zero_infinity = zero_infinity;
    gp.nll = nll; gp.grad_out = grad_out; gp.grad_out_stride = grad_out_stride; gp.grad = grad;
    const char* w = reinterpret_cast<const char*>(workspace);
    gp.alpha = reinterpret_cast<const float*>(w + pl.off_alpha);
    gp.beta = reinterpret_cast<const float*>(w + pl.off_beta);
    gp.coff_a = reinterpret_cast<const int*>(w + pl.off_coff_a);
    gp.coff_b = reinterpret_cast<const int*>(w + pl.off_coff_b);
    gp.nll2 = reinterpret_cast<const double*>(w + pl.off_nll2);
    gp.chain = reinterpret_cast<const int*>(w + pl.off_chain);
    gp.K = pl.K; gp.W = pl.W; gp.S_pad = pl.S_pad; gp.Lpad = pl.Lpad; gp.linear = pl.linear;
    gp.cw = pl.CW; gp.flag = nullptr; gp.run_if = 0; gp.done = nullptr; gp.stamp = avctc_tuning_get("ctc_stamp", 0);
    gp.prefetch = avctc_tuning_get("ctc_pf", 1);
    gp.stage = avctc_tuning_get("ctc_stage", 1);
    gp.row_floats = (V + 8 + 3) & ~3;
    gp.w_floats = (2 * max_target_len + 1 + 3) & ~3;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (pl.linear) {
        gp.flag = reinterpret_cast<const int*>(w + pl.off_flag);
        gp.run_if = 0;                            // guard not tripped: probability-domain workspaces
        if (avctc_tuning_get("ctc_overlap", 1) != 0)
            gp.done = reinterpret_cast<const int*>(w + pl.off_flag + kFlagBytes);
        int rc = (dtype == AVCTC_F32) ? dispatch_grad_lin<float>(gp, st) : dispatch_grad_lin<__nv_bfloat16>(gp, st);
        if (rc) return rc;
        GradParams gl = gp;                       // guard tripped: the log-domain scan rewrote the workspaces
        gl.K = pl.Klog; gl.W = pl.Wlog; gl.linear = 0; gl.run_if = 1;
        if (dtype == AVCTC_F32) return launch_grad<float>(gl, st);
        return launch_grad<__nv_bfloat16>(gl, st);
    }
    gp.K = pl.Klog; gp.W = pl.Wlog;
    if (dtype == AVCTC_F32) return launch_grad<float>(gp, st);
    return launch_grad<__nv_bfloat16>(gp, st);
}

extern "C" int avctc_ctc_scale_grad(void* grad, int dtype, int T, int B, int V, const float* grad_out,
                                    int64_t grad_out_stride, void* stream) {
    if (T < 0 || B < 0 || V <= 0) return AVCTC_ERR_BAD_ARG;
    if (T == 0 || B == 0) return AVCTC_OK;
    if (!grad || !grad_out) return AVCTC_ERR_BAD_ARG;
    if (dtype != AVCTC_F32 && dtype != AVCTC_BF16) return AVCTC_ERR_BAD_ARG;
    const long long rows = (long long)T * B;
    const int grid = (int)(rows < 148 * 8 ? rows : 148 * 8);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == AVCTC_F32)
        ctc_scale_grad_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<float*>(grad), rows, B, V, grad_out,
                                                           grad_out_stride);
    else
        ctc_scale_grad_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(grad), rows, B, V,
                                                                   grad_out, grad_out_stride);
    return (int)cudaGetLastError();
}

// forward + reduce + backward(unit or given grad_out) as ONE host call: what the autograd host enqueues at forward time.
// (Measured: issuing the gradient pass directly behind the scan, ahead of the guarded log-domain scan and the reduction,
// lets its late CTAs enter 2-4 us instead of 12-23 us after the last scan CTA, but the guard scan and the reduction then
// trail the gradient pass one after the other instead of running under it: 189 us either way at T=1000.)
extern "C" int avctc_ctc_forward_backward(const void* log_probs, int dtype, int64_t stride_t, int64_t stride_b, int T, int B,
                                          int V, const int64_t* targets, int64_t target_stride,
                                          const int64_t* target_offsets, const int64_t* input_lengths,
                                          const int64_t* target_lengths, int max_target_len, int blank, int reduction,
                                          int zero_infinity, float* nll, float* loss, const float* grad_out,
                                          int64_t grad_out_stride, void* grad, void* workspace, size_t workspace_bytes,
                                          void* stream) {
    int rc = avctc_ctc_forward(log_probs, dtype, stride_t, stride_b, T, B, V, targets, target_stride, target_offsets,
                               input_lengths, target_lengths, max_target_len, blank, 1, nll, workspace, workspace_bytes,
                               stream);
    if (rc) return rc;
    rc = avctc_ctc_reduce(nll, target_lengths, B, reduction, zero_infinity, loss, stream);
    if (rc) return rc;
    return avctc_ctc_backward(log_probs, dtype, stride_t, stride_b, T, B, V, targets, target_stride, target_offsets,
                              input_lengths, target_lengths, max_target_len, blank, reduction, zero_infinity, nll, grad_out,
                              grad_out_stride, grad, workspace, workspace_bytes, stream);
}
