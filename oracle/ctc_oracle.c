/*
 * oracle/ctc_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C, double-precision restatement of the CTC loss the reference trains with:
 *   nn.CTCLoss(blank=tokenizer.blank_id, zero_infinity=True)
 *     constructed  /root/reference/model/trainer.py:25   (dup. model/decoder.py:12)
 *     called       /root/reference/model/trainer.py:116-117, 224-225 (model/decoder.py:28-33)
 *
 * The arithmetic itself lives in an un-vendored third-party dependency: PyTorch (the reference
 * pins no version; this container has torch 2.11.0+cu128), ATen native op `_ctc_loss` /
 * `_ctc_loss_backward` (upstream aten/src/ATen/native/LossCTC.cpp, not present under
 * /root/reference). This file restates the published algorithm (Graves et al. 2006, eq. 6-16)
 * with ATen's conventions, which the golden fixtures in tests/golden/ctc_*.npz
 * (produced by oracle/gen_golden.py from torch.nn.CTCLoss on CPU in float64) pin:
 *   - label-extended lattice l' = (blank, l0, blank, l1, ..., blank), S = 2L+1 states
 *   - alpha_t(s) and beta_t(s) BOTH include the emission at t, so the posterior of state s at t
 *     is alpha+beta-lp
 *   - input_length == 0: nll = 0 when L == 0 else +inf
 *   - gradient returned is the softmax-folded one:  (exp(lp) - exp(lcab + nll - lp)) * g_b,
 *     zero for t >= input_length, and zero for the whole sample when nll==inf && zero_infinity
 *   - reduction 'mean' = mean_b( nll_b / max(L_b,1) ) with inf -> 0 first when zero_infinity
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static double lse2(double a, double b) {
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    double m = a > b ? a : b;
    return m + log(exp(a - m) + exp(b - m));
}

static inline int64_t ext_label(const int64_t* tg, int s, int blank) {
    return (s & 1) ? tg[s >> 1] : blank;
}

/* One sample. lp: [T][V] with element (t,c) at lp[t*stride_t + c]. alpha/beta: [T][S] scratch.
 * grad (may be NULL): [T][V] at grad[t*gstride_t + c], receives the UNSCALED-by-reduction
 * gradient times g. Returns nll. */
static double ctc_one(const double* lp, int64_t stride_t, int T_total, int V,
                      const int64_t* tg, int Tb, int L, int blank,
                      double* alpha, double* beta,
                      double* grad, int64_t gstride_t, double g, int zero_infinity) {
    const int S = 2 * L + 1;
    double nll;
    if (Tb == 0) {
        nll = (L == 0) ? 0.0 : INFINITY;
    } else {
        for (int s = 0; s < S; ++s) alpha[s] = -INFINITY;
        alpha[0] = lp[blank];
        if (L > 0) alpha[1] = lp[tg[0]];
        for (int t = 1; t < Tb; ++t) {
            const double* row = lp + (int64_t)t * stride_t;
            const double* ap = alpha + (int64_t)(t - 1) * S;
            double* an = alpha + (int64_t)t * S;
            for (int s = 0; s < S; ++s) {
                int64_t c = ext_label(tg, s, blank);
                double a = ap[s];
                if (s > 0) a = lse2(a, ap[s - 1]);
                if (s > 1 && ext_label(tg, s - 2, blank) != c) a = lse2(a, ap[s - 2]);
                an[s] = a + row[c];
            }
        }
        const double* al = alpha + (int64_t)(Tb - 1) * S;
        double ll = al[S - 1];
        if (L > 0) ll = lse2(ll, al[S - 2]);
        nll = -ll;
    }
    if (!grad) return nll;

    for (int t = 0; t < T_total; ++t)
        for (int c = 0; c < V; ++c) grad[(int64_t)t * gstride_t + c] = 0.0;
    if (Tb == 0) return nll;
    if (zero_infinity && nll == INFINITY) return nll;

    double* lcab = (double*)malloc(sizeof(double) * (size_t)V);
    for (int t = Tb - 1; t >= 0; --t) {
        const double* row = lp + (int64_t)t * stride_t;
        double* bn = beta + (int64_t)t * S;
        if (t == Tb - 1) {
            for (int s = 0; s < S; ++s) bn[s] = -INFINITY;
            bn[S - 1] = row[blank];
            if (L > 0) bn[S - 2] = row[tg[L - 1]];
        } else {
            const double* bp = beta + (int64_t)(t + 1) * S;
            for (int s = 0; s < S; ++s) {
                int64_t c = ext_label(tg, s, blank);
                double b = bp[s];
                if (s + 1 < S) b = lse2(b, bp[s + 1]);
                if (s + 2 < S && ext_label(tg, s + 2, blank) != c) b = lse2(b, bp[s + 2]);
                bn[s] = b + row[c];
            }
        }
        for (int c = 0; c < V; ++c) lcab[c] = -INFINITY;
        const double* at = alpha + (int64_t)t * S;
        for (int s = 0; s < S; ++s) {
            int64_t c = ext_label(tg, s, blank);
            lcab[c] = lse2(lcab[c], at[s] + bn[s]);
        }
        for (int c = 0; c < V; ++c) {
            double l = row[c];
            grad[(int64_t)t * gstride_t + c] = (exp(l) - exp(lcab[c] + nll - l)) * g;
        }
    }
    free(lcab);
    return nll;
}

/* Batched entry used by tests/ and bench.py's cpu_baseline leg through ctypes.
 * lp: element (t,b,c) at lp[t*st + b*sb + c]; targets: [B][Lmax] int64, row stride tstride.
 * nll_out[B]; loss_out[1] (reduction: 0 none(not written) / 1 mean / 2 sum);
 * grad (may be NULL): contiguous [T][B][V], already multiplied by d(loss)/d(nll_b) * grad_out. */
int ctc_oracle(const double* lp, int64_t st, int64_t sb, int T, int B, int V,
               const int64_t* targets, int64_t tstride,
               const int64_t* input_lengths, const int64_t* target_lengths,
               int blank, int reduction, int zero_infinity, double grad_out,
               double* nll_out, double* loss_out, double* grad) {
    int Lmax = 0;
    for (int b = 0; b < B; ++b) if (target_lengths[b] > Lmax) Lmax = (int)target_lengths[b];
    const int Smax = 2 * Lmax + 1;
    double* alpha = (double*)malloc(sizeof(double) * (size_t)(T > 0 ? T : 1) * Smax);
    double* beta = (double*)malloc(sizeof(double) * (size_t)(T > 0 ? T : 1) * Smax);
    double acc = 0.0;
    for (int b = 0; b < B; ++b) {
        const int L = (int)target_lengths[b];
        const int Tb = (int)input_lengths[b];
        if (Tb > T || L > Lmax || Tb < 0 || L < 0) { free(alpha); free(beta); return -1; }
        double denom = 1.0;
        if (reduction == 1) denom = (double)B * (double)(L > 1 ? L : 1);
        double g = grad_out / denom;
        double nll = ctc_one(lp + (int64_t)b * sb, st, T, V, targets + (int64_t)b * tstride, Tb, L,
                             blank, alpha, beta, grad ? grad + (int64_t)b * V : 0,
                             (int64_t)B * V, g, zero_infinity);
        nll_out[b] = nll;
        double v = (zero_infinity && nll == INFINITY) ? 0.0 : nll;
        acc += (reduction == 1) ? v / (double)(L > 1 ? L : 1) : v;
    }
    if (loss_out) loss_out[0] = (reduction == 1) ? acc / (double)B : acc;
    free(alpha); free(beta);
    return 0;
}
