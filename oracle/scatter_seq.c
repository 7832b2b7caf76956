/*
 * TEST INFRASTRUCTURE ONLY -- never linked, imported or called by the product
 * path (mma_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.
 *
 * Sequential CPU restatement of the third-party segment reductions the
 * reference calls on its hot path:
 *
 *   torch_scatter.scatter(src, index, dim=0, out=None, dim_size, reduce=...)
 *     called at /root/reference/graph_regression/mma_conv.py:166,168,169
 *   torch_geometric.utils.degree(index, dim_size)
 *     called at /root/reference/graph_regression/mma_conv.py:178
 *
 * torch_scatter is NOT vendored in /root/reference and not installable here
 * (no requirements file; README.md:34-38 implies torch-scatter 2.0.8/2.0.9).
 * Published algorithm restated (csrc/cpu/scatter_cpu.cpp + reducer.h of that
 * era): one sequential loop over e = 0..E-1 in edge order;
 *   sum : out[idx[e]] += src[e]
 *   mean: sum, then out /= max(count, 1)
 *   min : out initialised to +FLT_MAX, `if (src[e] < out) { out = src[e]; arg = e; }`
 *   max : out initialised to -FLT_MAX, `if (src[e] > out) { out = src[e]; arg = e; }`
 *         => first occurrence wins, -0.0 == +0.0 never replaces, NaN never wins;
 *         afterwards entries still equal to the init value (empty rows) are
 *         set to 0 and keep arg = E.
 * The loops are sequential in e on purpose: this file is the ground truth for
 * tie-breaking order, against which both the vectorised torch restatement
 * (oracle/restate.py) and the CUDA kernels are checked bit-for-bit.
 *
 * PARITY UNPINNED: the reference ships no golden vectors or tests for this
 * path; see DESIGN.md.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

/* src [E,F] row-major, index [E] in [0,N), out [N,F], arg [N,F] (may be NULL) */

void seq_scatter_sum(const float *src, const int64_t *index, int64_t E, int64_t F,
                     int64_t N, float *out)
{
    memset(out, 0, sizeof(float) * (size_t)(N * F));
    for (int64_t e = 0; e < E; ++e) {
        float *o = out + index[e] * F;
        const float *s = src + e * F;
        for (int64_t f = 0; f < F; ++f) o[f] += s[f];
    }
}

void seq_degree(const int64_t *index, int64_t E, int64_t N, float *deg)
{
    memset(deg, 0, sizeof(float) * (size_t)N);
    for (int64_t e = 0; e < E; ++e) deg[index[e]] += 1.0f;
}

void seq_scatter_mean(const float *src, const int64_t *index, int64_t E, int64_t F,
                      int64_t N, float *out)
{
    seq_scatter_sum(src, index, E, F, N, out);
    /* count via the same scatter of ones, clamp(min=1), true division */
    for (int64_t n = 0; n < N; ++n) { (void)n; }
    float *cnt = (float *)__builtin_malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
    seq_degree(index, E, N, cnt);
    for (int64_t n = 0; n < N; ++n) {
        float c = cnt[n] < 1.0f ? 1.0f : cnt[n];
        for (int64_t f = 0; f < F; ++f) out[n * F + f] = out[n * F + f] / c;
    }
    __builtin_free(cnt);
}

static void seq_scatter_minmax(const float *src, const int64_t *index, int64_t E,
                               int64_t F, int64_t N, float *out, int64_t *arg,
                               int is_max)
{
    const float init = is_max ? -FLT_MAX : FLT_MAX;
    for (int64_t k = 0; k < N * F; ++k) { out[k] = init; if (arg) arg[k] = E; }
    for (int64_t e = 0; e < E; ++e) {
        float *o = out + index[e] * F;
        int64_t *a = arg ? arg + index[e] * F : 0;
        const float *s = src + e * F;
        for (int64_t f = 0; f < F; ++f) {
            int better = is_max ? (s[f] > o[f]) : (s[f] < o[f]);
            if (better) { o[f] = s[f]; if (a) a[f] = e; }
        }
    }
    for (int64_t k = 0; k < N * F; ++k) if (out[k] == init) out[k] = 0.0f;
}

void seq_scatter_min(const float *src, const int64_t *index, int64_t E, int64_t F,
                     int64_t N, float *out, int64_t *arg)
{ seq_scatter_minmax(src, index, E, F, N, out, arg, 0); }

void seq_scatter_max(const float *src, const int64_t *index, int64_t E, int64_t F,
                     int64_t N, float *out, int64_t *arg)
{ seq_scatter_minmax(src, index, E, F, N, out, arg, 1); }

/*
 * Fused-op restatement used to pin the kernels bit-for-bit on min/max:
 * message m_e = ((P[dst_e] + Q[src_e]) + R[e]) * keep[e]   (each term optional),
 * restating /root/reference/graph_regression/mma_conv.py:146-157 with the
 * separable mask linear, then the reductions of
 * mma_conv.py:163-172 sequentially in edge order.  Outputs are the five raw
 * aggregates [N,F] each (any may be NULL): sum, mean, min, max, std, plus args.
 */
void seq_mmconv_aggregate(const float *P, const float *Q, const float *R,
                          const float *keep, const int64_t *src_idx,
                          const int64_t *dst_idx, int64_t E, int64_t F, int64_t N,
                          float *o_sum, float *o_mean, float *o_min, float *o_max,
                          float *o_std, float *o_var, int64_t *arg_min, int64_t *arg_max)
{
    float *m = (float *)__builtin_malloc(sizeof(float) * (size_t)(E * F > 0 ? E * F : 1));
    float *m2 = (float *)__builtin_malloc(sizeof(float) * (size_t)(E * F > 0 ? E * F : 1));
    for (int64_t e = 0; e < E; ++e)
        for (int64_t f = 0; f < F; ++f) {
            volatile float v = 0.0f;
            if (P && Q) v = P[dst_idx[e] * F + f] + Q[src_idx[e] * F + f];
            else if (P) v = P[dst_idx[e] * F + f];
            else if (Q) v = Q[src_idx[e] * F + f];
            if (R) { if (P || Q) v = v + R[e * F + f]; else v = R[e * F + f]; }
            if (keep) v = v * keep[e * F + f];
            m[e * F + f] = v;
            volatile float sq = v * v;
            m2[e * F + f] = sq;
        }
    if (o_sum) seq_scatter_sum(m, dst_idx, E, F, N, o_sum);
    if (o_mean) seq_scatter_mean(m, dst_idx, E, F, N, o_mean);
    if (o_min) seq_scatter_min(m, dst_idx, E, F, N, o_min, arg_min);
    if (o_max) seq_scatter_max(m, dst_idx, E, F, N, o_max, arg_max);
    if (o_std || o_var) {
        float *mean = (float *)__builtin_malloc(sizeof(float) * (size_t)(N * F > 0 ? N * F : 1));
        float *msq = (float *)__builtin_malloc(sizeof(float) * (size_t)(N * F > 0 ? N * F : 1));
        seq_scatter_mean(m, dst_idx, E, F, N, mean);
        seq_scatter_mean(m2, dst_idx, E, F, N, msq);
        for (int64_t k = 0; k < N * F; ++k) {
            volatile float mm = mean[k] * mean[k];      /* no FMA contraction */
            volatile float var = msq[k] - mm;
            if (o_var) o_var[k] = var;
            if (o_std) {
                volatile float r = var > 0.0f ? var : 0.0f;
                volatile float t = r + 1e-5f;
                o_std[k] = sqrtf(t);
            }
        }
        __builtin_free(mean); __builtin_free(msq);
    }
    __builtin_free(m); __builtin_free(m2);
}
