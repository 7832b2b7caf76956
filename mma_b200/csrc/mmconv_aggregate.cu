// K1: fused MultiMaskConv aggregate, forward + destination pass of the backward.
//
// Replaces (reference paths relative to /root/reference/graph_regression):
//   mma_conv.py:130      PyG propagate: x_j = x[src], x_i = x[dst] materialised as [E,T,F_in]
//   mma_conv.py:146-157  message: mask linear over cat([x_i,x_j,e]) (separable -> P[dst]+Q[src]+R[e]),
//                        always-on dropout
//   mma_conv.py:159-196  aggregate: A torch_scatter passes (+2 for var/std), degree, cumulative scalers, cats
// and the autograd backward of all of it, without atomics.
//
// Mapping: a group of LANES (<=32, power of two) threads owns one destination row x one
// chunk of LANES*VEC feature columns; each lane keeps its VEC columns' running
// sum / sum-of-squares / (min,pos) / (max,pos) in registers and walks the row's in-edges in
// CSR order == original edge order (the CSR is a STABLE sort by destination).  Because a
// column is owned by one lane and scanned sequentially, "first strict improvement wins"
// (torch_scatter CPU) and the sequential fp32 summation order are reproduced exactly, with
// no cross-lane combine.  Loads of the gathered rows are 128-bit, coalesced along F, issued
// U at a time before use for memory-level parallelism.
#include "common.cuh"

namespace mma {

struct MMConvParams {
    const int32_t *rowptr, *col, *perm, *gid;
    int64_t n_rows, E, E_total;
    const float *P, *Q, *R, *keep;
    int64_t ldp, ldq, ldr, ldk;
    Dropout drop;
    int use_philox;
    int T, F_in, F, A, S;
    int akind[MMA_MAX_AGGR];
    int skind[MMA_MAX_SCALER];
    const float *scale_tab;
    int64_t tab_stride;
    // forward outputs
    float *Y;
    int64_t ldy;
    int32_t *arg_min, *arg_max;
    float *stat_mean, *stat_var;
    // backward inputs / outputs
    const float *dY;
    const int32_t *c_arg_min, *c_arg_max;
    const float *c_mean, *c_var;
    const int32_t *gslot;
    float *G;
    int64_t ldg;
    float *dP;
    int64_t lddp;
    // launch geometry
    int lanes_log2, chunks;
    int col0, ncols;        // column window [col0, col0 + ncols) processed by this launch
    int64_t n_groups;
};

template <int VEC, bool MINMAX, bool SQ>
struct Acc {
    float sum[VEC];
    float sq[SQ ? VEC : 1];
    float mn[MINMAX ? VEC : 1], mx[MINMAX ? VEC : 1];
    int amn[MINMAX ? VEC : 1], amx[MINMAX ? VEC : 1];
};

// message of one edge for this lane's columns, in the reference's arithmetic order
template <int VEC>
__device__ __forceinline__ Vec<VEC> message(const MMConvParams &p, const Vec<VEC> &pv, const Vec<VEC> &q,
                                            const Vec<VEC> &r, const Vec<VEC> &ks, bool has_scale) {
    Vec<VEC> m;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        float x;
        if (p.P && p.Q) x = __fadd_rn(pv.v[v], q.v[v]);
        else if (p.P) x = pv.v[v];
        else if (p.Q) x = q.v[v];
        else x = r.v[v];
        if (p.R && (p.P || p.Q)) x = __fadd_rn(x, r.v[v]);
        if (has_scale) x = __fmul_rn(x, ks.v[v]);     // x * 0 keeps the sign of x, like F.dropout
        m.v[v] = x;
    }
    return m;
}

// loads everything U edges need (all loads issued before any use), then consumes in order
template <int VEC, int U, typename Consume>
__device__ __forceinline__ void visit_edges(const MMConvParams &p, int pos, int c, const Vec<VEC> &pv,
                                            bool need_m, bool need_eid, Consume &&consume) {
    int j[U], eid[U], ge[U];
    Vec<VEC> q[U], r[U], ks[U];
    if (need_m && p.Q) {
#pragma unroll
        for (int u = 0; u < U; ++u) j[u] = __ldg(p.col + pos + u);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) eid[u] = (need_eid && p.perm) ? __ldg(p.perm + pos + u) : pos + u;
#pragma unroll
    for (int u = 0; u < U; ++u) ge[u] = (need_eid && p.gid) ? __ldg(p.gid + eid[u]) : eid[u];   // global edge id
    if (need_m && p.Q) {
#pragma unroll
        for (int u = 0; u < U; ++u) q[u] = ld_vec_stream<VEC>(p.Q + (int64_t)j[u] * p.ldq + c);
    }
    if (need_m && p.R) {
#pragma unroll
        for (int u = 0; u < U; ++u) r[u] = ld_vec_stream<VEC>(p.R + (int64_t)eid[u] * p.ldr + c);
    }
    const bool has_scale = p.keep != nullptr || p.use_philox;
    if (p.keep) {
#pragma unroll
        for (int u = 0; u < U; ++u) ks[u] = ld_vec_stream<VEC>(p.keep + (int64_t)eid[u] * p.ldk + c);
    } else if (p.use_philox) {
#pragma unroll
        for (int u = 0; u < U; ++u) ks[u] = dropout_keep<VEC>(p.drop, (uint32_t)ge[u], c, 0u);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        Vec<VEC> m{};
        if (need_m) m = message<VEC>(p, pv, q[u], r[u], ks[u], has_scale);
        consume(pos + u, ge[u], m, ks[u], has_scale);
    }
}

template <int VEC, typename Consume>
__device__ __forceinline__ void for_each_edge(const MMConvParams &p, int beg, int end, int c,
                                              const Vec<VEC> &pv, bool need_m, bool need_eid,
                                              Consume &&consume) {
    constexpr int U = (VEC == 4) ? 4 : 8;
    int pos = beg;
    for (; pos + U <= end; pos += U) visit_edges<VEC, U>(p, pos, c, pv, need_m, need_eid, consume);
    for (; pos < end; ++pos) visit_edges<VEC, 1>(p, pos, c, pv, need_m, need_eid, consume);
}

__device__ __forceinline__ bool locate(const MMConvParams &p, int64_t &row, int &c, int vec) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t group = tid >> p.lanes_log2;
    if (group >= p.n_groups) return false;
    const int sub = (int)(tid & ((1 << p.lanes_log2) - 1));
    row = group / p.chunks;
    const int chunk = (int)(group - row * p.chunks);
    c = ((chunk << p.lanes_log2) + sub) * vec;
    return c < p.F;
}

// CSR position -> (global) original edge id
__device__ __forceinline__ int32_t orig_edge_id(const MMConvParams &p, int pos) {
    const int32_t e = p.perm ? __ldg(p.perm + pos) : pos;
    return p.gid ? __ldg(p.gid + e) : e;
}

__device__ __forceinline__ void scaler_factors(const MMConvParams &p, int degc, float *fac) {
    const int64_t d = degc < p.tab_stride ? degc : p.tab_stride - 1;
#pragma unroll
    for (int s = 0; s < MMA_MAX_SCALER; ++s) {
        if (s < p.S) {
            const int k = p.skind[s];
            fac[s] = (k == MMA_SCALE_IDENTITY) ? 1.0f : __ldg(p.scale_tab + (int64_t)(k - 1) * p.tab_stride + d);
        }
    }
}

// ----------------------------------------------------------------------------------------
// forward
// ----------------------------------------------------------------------------------------
template <int VEC, bool MINMAX, bool SQ>
__global__ void __launch_bounds__(256) mmconv_fwd_kernel(const __grid_constant__ MMConvParams p) {
    int64_t row;
    int c;
    if (!locate(p, row, c, VEC)) return;
    const int beg = __ldg(p.rowptr + row), end = __ldg(p.rowptr + row + 1);

    Vec<VEC> pv{};
    if (p.P) pv = ld_vec<VEC>(p.P + row * p.ldp + c);

    Acc<VEC, MINMAX, SQ> acc;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        acc.sum[v] = 0.0f;
        if constexpr (SQ) acc.sq[v] = 0.0f;
        if constexpr (MINMAX) { acc.mn[v] = FLT_MAX; acc.mx[v] = -FLT_MAX; acc.amn[v] = -1; acc.amx[v] = -1; }
    }

    const bool need_eid = p.R != nullptr || p.keep != nullptr || p.use_philox;
    for_each_edge<VEC>(p, beg, end, c, pv, true, need_eid,
        [&](int pos, int, const Vec<VEC> &m, const Vec<VEC> &, bool) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float x = m.v[v];
                acc.sum[v] = __fadd_rn(acc.sum[v], x);
                if constexpr (SQ) acc.sq[v] = __fadd_rn(acc.sq[v], __fmul_rn(x, x));
                if constexpr (MINMAX) {
                    if (x < acc.mn[v]) { acc.mn[v] = x; acc.amn[v] = pos; }   // strict: first occurrence wins,
                    if (x > acc.mx[v]) { acc.mx[v] = x; acc.amx[v] = pos; }   // -0.0 == +0.0, NaN never wins
                }
            }
        });

    // ---- epilogue: aggregates -> cumulative scalers -> Y[row, t, (s*A+a)*F_in + f] ----
    const int cnt = end - beg;
    const int degc = cnt > 1 ? cnt : 1;                    // deg.clamp_(1), mma_conv.py:179
    const float degf = (float)degc;
    float fac[MMA_MAX_SCALER];
    scaler_factors(p, degc, fac);

    Vec<VEC> mean, var;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        mean.v[v] = __fdiv_rn(acc.sum[v], degf);           // sum / count.clamp(min=1)
        if constexpr (SQ) {
            const float msq = __fdiv_rn(acc.sq[v], degf);
            var.v[v] = __fsub_rn(msq, __fmul_rn(mean.v[v], mean.v[v]));   // mma_conv.py:170, no FMA
        } else {
            var.v[v] = 0.0f;
        }
    }

    const int t = c / p.F_in, f = c - t * p.F_in;
    float *yrow = p.Y + row * p.ldy + (int64_t)t * ((int64_t)p.S * p.A * p.F_in) + f;
    for (int a = 0; a < p.A; ++a) {
        Vec<VEC> val;
        const int kind = p.akind[a];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            float x;
            switch (kind) {
                case MMA_AGGR_SUM: x = acc.sum[v]; break;
                case MMA_AGGR_MEAN: x = mean.v[v]; break;
                case MMA_AGGR_MIN: x = MINMAX ? (acc.amn[v] >= 0 ? acc.mn[v] : 0.0f) : 0.0f; break;
                case MMA_AGGR_MAX: x = MINMAX ? (acc.amx[v] >= 0 ? acc.mx[v] : 0.0f) : 0.0f; break;
                case MMA_AGGR_VAR: x = var.v[v]; break;
                default: x = sqrtf(__fadd_rn(fmaxf(var.v[v], 0.0f), 1e-5f)); break;   // STD, mma_conv.py:172
            }
            val.v[v] = x;
        }
        for (int s = 0; s < p.S; ++s) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) val.v[v] = __fmul_rn(val.v[v], fac[s]);     // cumulative (Q4)
            st_vec_stream<VEC>(yrow + (int64_t)(s * p.A + a) * p.F_in, val);
        }
    }

    if constexpr (MINMAX) {
        int32_t amn[VEC], amx[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            amn[v] = acc.amn[v] < 0 ? (int32_t)p.E_total : orig_edge_id(p, acc.amn[v]);
            amx[v] = acc.amx[v] < 0 ? (int32_t)p.E_total : orig_edge_id(p, acc.amx[v]);
        }
        if (p.arg_min) st_vec_i32_stream<VEC>(p.arg_min + row * p.F + c, amn);
        if (p.arg_max) st_vec_i32_stream<VEC>(p.arg_max + row * p.F + c, amx);
    }
    if (p.stat_mean) st_vec_stream<VEC>(p.stat_mean + row * p.F + c, mean);
    if constexpr (SQ) {
        if (p.stat_var) st_vec_stream<VEC>(p.stat_var + row * p.F + c, var);
    }
}

// ----------------------------------------------------------------------------------------
// backward, destination pass: per-edge gradient rows G and dP
// ----------------------------------------------------------------------------------------
template <int VEC, bool NEEDM>
__global__ void __launch_bounds__(256) mmconv_bwd_dst_kernel(const __grid_constant__ MMConvParams p) {
    int64_t row;
    int c;
    if (!locate(p, row, c, VEC)) return;
    const int beg = __ldg(p.rowptr + row), end = __ldg(p.rowptr + row + 1);
    const int cnt = end - beg;
    const int degc = cnt > 1 ? cnt : 1;
    const float degf = (float)degc;

    // cumulative scaler factors: block s of Y carries prod_{s'<=s} f_{s'}
    float fac[MMA_MAX_SCALER];
    scaler_factors(p, degc, fac);
    for (int s = 1; s < p.S; ++s) fac[s] *= fac[s - 1];

    Vec<VEC> mean{}, var{};
    if constexpr (NEEDM) {
        mean = ld_vec<VEC>(p.c_mean + row * p.F + c);
        var = ld_vec<VEC>(p.c_var + row * p.F + c);
    }

    // fold dY over scalers and aggregators into: base (same for every in-edge), gmin / gmax
    // (routed to the arg edge only) and alpha (coefficient of m_e, from var/std)
    Vec<VEC> base{}, gmin{}, gmax{}, alpha{};
    bool has_min = false, has_max = false;
    const int t = c / p.F_in, f = c - t * p.F_in;
    const float *dyrow = p.dY + row * p.ldy + (int64_t)t * ((int64_t)p.S * p.A * p.F_in) + f;
    for (int a = 0; a < p.A; ++a) {
        Vec<VEC> dz{};
        for (int s = 0; s < p.S; ++s) {
            const Vec<VEC> d = ld_vec_stream<VEC>(dyrow + (int64_t)(s * p.A + a) * p.F_in);
#pragma unroll
            for (int v = 0; v < VEC; ++v) dz.v[v] += d.v[v] * fac[s];
        }
        const int kind = p.akind[a];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const float g = dz.v[v];
            switch (kind) {
                case MMA_AGGR_SUM: base.v[v] += g; break;
                case MMA_AGGR_MEAN: base.v[v] += g / degf; break;
                case MMA_AGGR_MIN: gmin.v[v] += g; break;
                case MMA_AGGR_MAX: gmax.v[v] += g; break;
                case MMA_AGGR_VAR: {            // var = E[m^2] - E[m]^2 -> d/dm_e = 2 (m_e - mean) / cnt
                    const float k = 2.0f * g / degf;
                    alpha.v[v] += k; base.v[v] -= k * mean.v[v];
                } break;
                default: {                      // std = sqrt(relu(var) + 1e-5); relu'(0) = 0
                    if (var.v[v] > 0.0f) {
                        const float sd = sqrtf(var.v[v] + 1e-5f);
                        const float k = g / (sd * degf);
                        alpha.v[v] += k; base.v[v] -= k * mean.v[v];
                    }
                } break;
            }
        }
        has_min |= kind == MMA_AGGR_MIN;
        has_max |= kind == MMA_AGGR_MAX;
    }

    int32_t amn[VEC], amx[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { amn[v] = -1; amx[v] = -1; }
    if (has_min) ld_vec_i32_as<VEC>(p.c_arg_min + row * p.F + c, amn);
    if (has_max) ld_vec_i32_as<VEC>(p.c_arg_max + row * p.F + c, amx);

    Vec<VEC> pv{};
    if (NEEDM && p.P) pv = ld_vec<VEC>(p.P + row * p.ldp + c);

    Vec<VEC> dp{};
    for_each_edge<VEC>(p, beg, end, c, pv, NEEDM, true,
        [&](int pos, int eid, const Vec<VEC> &m, const Vec<VEC> &ks, bool has_scale) {
            Vec<VEC> g;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float x = base.v[v];
                if (eid == amn[v]) x += gmin.v[v];
                if (eid == amx[v]) x += gmax.v[v];
                if constexpr (NEEDM) x += alpha.v[v] * m.v[v];
                if (has_scale) x *= ks.v[v];           // dL/dm_pre = dL/dm * keepscale
                g.v[v] = x;
                dp.v[v] += x;
            }
            if (p.G) {
                const int64_t slot = p.gslot ? (int64_t)__ldg(p.gslot + pos) : (int64_t)pos;
                st_vec<VEC>(p.G + slot * p.ldg + c, g);
            }
        });
    if (p.dP) st_vec_stream<VEC>(p.dP + row * p.lddp + c, dp);
}

// ----------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------
static int fill_params(MMConvParams &p, const int32_t *rowptr, const int32_t *col, const int32_t *perm,
                       const int32_t *gid, int64_t E_total, int64_t n_rows, int64_t E, const float *P, int64_t ldp, const float *Q, int64_t ldq,
                       const float *R, int64_t ldr, const float *keep, int64_t ldk, float p_drop, uint64_t seed,
                       int T, int F_in, int A, const int32_t *aggr_kinds, int S, const int32_t *scaler_kinds,
                       const float *scale_tab, int64_t tab_stride) {
    if (!rowptr || n_rows < 0 || E < 0 || T < 1 || F_in < 1 || A < 1 || S < 1 || !aggr_kinds || !scaler_kinds)
        return MMA_ERR_INVALID;
    if (E >= INT32_MAX || n_rows >= INT32_MAX) return MMA_ERR_UNSUPPORTED;
    if (A > MMA_MAX_AGGR || S > MMA_MAX_SCALER) return MMA_ERR_UNSUPPORTED;
    if (!P && !Q && !R && E > 0) return MMA_ERR_INVALID;
    if (Q && !col && E > 0) return MMA_ERR_INVALID;
    if (p_drop < 0.0f || p_drop > 1.0f) return MMA_ERR_INVALID;
    p = MMConvParams{};
    p.rowptr = rowptr; p.col = col; p.perm = perm; p.gid = gid; p.n_rows = n_rows; p.E = E;
    p.E_total = gid ? E_total : E;
    p.P = P; p.Q = Q; p.R = R; p.keep = keep; p.ldp = ldp; p.ldq = ldq; p.ldr = ldr; p.ldk = ldk;
    p.drop = make_dropout(p_drop, seed);
    p.use_philox = (keep == nullptr && p_drop > 0.0f) ? 1 : 0;
    p.T = T; p.F_in = F_in; p.F = T * F_in; p.A = A; p.S = S;
    bool any_scaled = false;
    for (int a = 0; a < A; ++a) {
        if (aggr_kinds[a] < MMA_AGGR_SUM || aggr_kinds[a] > MMA_AGGR_STD) return MMA_ERR_INVALID;
        p.akind[a] = aggr_kinds[a];
    }
    for (int s = 0; s < S; ++s) {
        if (scaler_kinds[s] < MMA_SCALE_IDENTITY || scaler_kinds[s] > MMA_SCALE_INVERSE_LINEAR) return MMA_ERR_INVALID;
        p.skind[s] = scaler_kinds[s];
        any_scaled |= scaler_kinds[s] != MMA_SCALE_IDENTITY;
    }
    if (any_scaled && (!scale_tab || tab_stride < 2)) return MMA_ERR_INVALID;
    p.scale_tab = scale_tab; p.tab_stride = tab_stride > 0 ? tab_stride : 1;
    return MMA_OK;
}

static int choose_geometry(MMConvParams &p, bool vec4_ok, int col0, int ncols) {
    if (ncols <= 0) { col0 = 0; ncols = p.F; }
    p.col0 = col0; p.ncols = ncols;
    vec4_ok = vec4_ok && (col0 % 4 == 0) && (ncols % 4 == 0);
    const int vec = vec4_ok ? 4 : 1;
    const int per_row = (ncols + vec - 1) / vec;     // lanes needed for one row of the window
    int lg = 0;
    while ((1 << lg) < per_row && lg < 5) ++lg;
    p.lanes_log2 = lg;
    const int lanes = 1 << lg;
    p.chunks = (per_row + lanes - 1) / lanes;
    p.n_groups = p.n_rows * p.chunks;
    return vec;
}

static inline bool ok4(const void *ptr, int64_t ld) { return ptr == nullptr || (aligned16(ptr) && (ld % 4) == 0); }

}  // namespace mma

using namespace mma;

extern "C" int mmconv_aggregate_fwd(const int32_t *rowptr, const int32_t *col, const int32_t *perm,
                                    const int32_t *edge_gid, int64_t E_total, int64_t n_rows, int64_t E, const float *P, int64_t ldp,
                                    const float *Q, int64_t ldq, const float *R, int64_t ldr,
                                    const float *keep, int64_t ldk, float p_drop, uint64_t seed,
                                    int T, int F_in, int A, const int32_t *aggr_kinds, int S,
                                    const int32_t *scaler_kinds, const float *scale_tab, int64_t tab_stride,
                                    float *Y, int64_t ldy, int32_t *arg_min, int32_t *arg_max,
                                    float *stat_mean, float *stat_var, int col0, int ncols,
                                    mma_stream_t stream) {
    MMConvParams p;
    int rc = fill_params(p, rowptr, col, perm, edge_gid, E_total, n_rows, E, P, ldp, Q, ldq, R, ldr, keep, ldk, p_drop, seed,
                         T, F_in, A, aggr_kinds, S, scaler_kinds, scale_tab, tab_stride);
    if (rc != MMA_OK) return rc;
    if (!Y) return MMA_ERR_INVALID;
    p.Y = Y; p.ldy = ldy; p.arg_min = arg_min; p.arg_max = arg_max; p.stat_mean = stat_mean; p.stat_var = stat_var;
    if (n_rows == 0) return MMA_OK;
    bool minmax = false, sq = false;
    for (int a = 0; a < A; ++a) {
        minmax |= (p.akind[a] == MMA_AGGR_MIN || p.akind[a] == MMA_AGGR_MAX);
        sq |= (p.akind[a] == MMA_AGGR_VAR || p.akind[a] == MMA_AGGR_STD);
    }
    const bool v4 = (F_in % 4 == 0) && ok4(P, ldp) && ok4(Q, ldq) && ok4(R, ldr) && ok4(keep, ldk) &&
                    ok4(Y, ldy) && ok4(arg_min, 4) && ok4(arg_max, 4) && ok4(stat_mean, 4) && ok4(stat_var, 4);
    if (col0 < 0 || col0 + (ncols > 0 ? ncols : 0) > p.F) return MMA_ERR_INVALID;
    const int vec = choose_geometry(p, v4, col0, ncols);
    const int64_t threads = p.n_groups << p.lanes_log2;
    const int block = 256;
    const int64_t grid = (threads + block - 1) / block;
    if (grid > INT32_MAX) return MMA_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define LAUNCH_FWD(V, MM, SQ) mmconv_fwd_kernel<V, MM, SQ><<<(unsigned)grid, block, 0, st>>>(p)
    if (vec == 4) {
        if (minmax && sq) LAUNCH_FWD(4, true, true);
        else if (minmax) LAUNCH_FWD(4, true, false);
        else if (sq) LAUNCH_FWD(4, false, true);
        else LAUNCH_FWD(4, false, false);
    } else {
        if (minmax && sq) LAUNCH_FWD(1, true, true);
        else if (minmax) LAUNCH_FWD(1, true, false);
        else if (sq) LAUNCH_FWD(1, false, true);
        else LAUNCH_FWD(1, false, false);
    }
#undef LAUNCH_FWD
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}

extern "C" int mmconv_aggregate_bwd_dst(const int32_t *rowptr, const int32_t *col, const int32_t *perm,
                                        const int32_t *edge_gid, int64_t E_total, int64_t n_rows, int64_t E, const float *P, int64_t ldp,
                                        const float *Q, int64_t ldq, const float *R, int64_t ldr,
                                        const float *keep, int64_t ldk, float p_drop, uint64_t seed,
                                        int T, int F_in, int A, const int32_t *aggr_kinds, int S,
                                        const int32_t *scaler_kinds, const float *scale_tab, int64_t tab_stride,
                                        const float *dY, int64_t ldy, const int32_t *arg_min,
                                        const int32_t *arg_max, const float *stat_mean, const float *stat_var,
                                        const int32_t *gslot, float *G, int64_t ldg, float *dP, int64_t lddp,
                                        int col0, int ncols, mma_stream_t stream) {
    MMConvParams p;
    int rc = fill_params(p, rowptr, col, perm, edge_gid, E_total, n_rows, E, P, ldp, Q, ldq, R, ldr, keep, ldk, p_drop, seed,
                         T, F_in, A, aggr_kinds, S, scaler_kinds, scale_tab, tab_stride);
    if (rc != MMA_OK) return rc;
    if (!dY || (!G && !dP && E > 0)) return MMA_ERR_INVALID;
    bool needm = false;
    for (int a = 0; a < A; ++a) {
        const int k = p.akind[a];
        if ((k == MMA_AGGR_MIN && !arg_min) || (k == MMA_AGGR_MAX && !arg_max)) return MMA_ERR_INVALID;
        if (k == MMA_AGGR_VAR || k == MMA_AGGR_STD) needm = true;
    }
    if (needm && (!stat_mean || !stat_var)) return MMA_ERR_INVALID;
    p.dY = dY; p.ldy = ldy; p.c_arg_min = arg_min; p.c_arg_max = arg_max; p.c_mean = stat_mean; p.c_var = stat_var;
    p.gslot = gslot; p.G = G; p.ldg = ldg; p.dP = dP; p.lddp = lddp;
    if (n_rows == 0) return MMA_OK;
    const bool v4 = (F_in % 4 == 0) && ok4(P, ldp) && ok4(Q, ldq) && ok4(R, ldr) && ok4(keep, ldk) &&
                    ok4(dY, ldy) && ok4(arg_min, 4) && ok4(arg_max, 4) && ok4(stat_mean, 4) &&
                    ok4(stat_var, 4) && ok4(G, ldg) && ok4(dP, lddp);
    if (col0 < 0 || col0 + (ncols > 0 ? ncols : 0) > p.F) return MMA_ERR_INVALID;
    const int vec = choose_geometry(p, v4, col0, ncols);
    const int64_t threads = p.n_groups << p.lanes_log2;
    const int block = 256;
    const int64_t grid = (threads + block - 1) / block;
    if (grid > INT32_MAX) return MMA_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (vec == 4) {
        if (needm) mmconv_bwd_dst_kernel<4, true><<<(unsigned)grid, block, 0, st>>>(p);
        else mmconv_bwd_dst_kernel<4, false><<<(unsigned)grid, block, 0, st>>>(p);
    } else {
        if (needm) mmconv_bwd_dst_kernel<1, true><<<(unsigned)grid, block, 0, st>>>(p);
        else mmconv_bwd_dst_kernel<1, false><<<(unsigned)grid, block, 0, st>>>(p);
    }
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}
