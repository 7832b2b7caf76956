// K1: fused MultiMaskConv aggregate, forward + destination pass of the backward.
//
// Replaces (reference paths relative to /root/reference/graph_regression):
//   mma_conv.py:130      PyG propagate: x_j = x[src], x_i = x[dst] materialised as [E,T,F_in]
//   mma_conv.py:146-157  message: mask linear over cat([x_i,x_j,e]) (separable -> P[dst]+Q[src]+R[e]),
//                        always-on dropout
//   mma_conv.py:159-196  aggregate: A torch_scatter passes (+2 for var/std), degree, cumulative scalers, cats
// and the autograd backward of all of it, without atomics.
//
// Mapping: a group of L (<=32, power of two) threads owns one destination row x one window of
// L*VEC feature columns; each lane keeps its VEC columns' running sum / sum-of-squares /
// (min,pos) / (max,pos) in registers and walks the row's in-edges in CSR order == original edge
// order (the CSR is a STABLE sort by destination).  A column is owned by one lane and scanned
// sequentially, so "first strict improvement wins" (torch_scatter CPU) and the sequential fp32
// summation order are reproduced exactly, with no cross-lane combine of values.
//
// Per row the group cooperates on everything that is NOT per-column:
//   * edge indices: lane s loads col[base+s] / perm[base+s] (one coalesced load per L edges) and
//     the group broadcasts them with __shfl_sync -> no dependent uniform loads in the hot loop;
//   * dropout bits: one Philox4x32-10 call yields 16 columns x 8 bits; for a batch of 4 edges the
//     L lanes generate the L calls the group needs (one per lane) and exchange the words by
//     shuffle, instead of every lane running Philox for every edge;
//   * the gathered rows Q[src] (and R[e], keep[e]) are fetched with 128-bit loads, 4 edges
//     (4 x 512 B per warp at F=128) in flight per group before any is consumed.
#include "common.cuh"

namespace mma {

enum { DROP_NONE = 0, DROP_KEEP = 1, DROP_PHILOX_SHARED = 2, DROP_PHILOX_LANE = 3 };
// which terms make up the message (compile time, so the hot loop carries no null checks)
enum { MSG_PQ = 0, MSG_PQR = 1, MSG_R = 2, MSG_GENERIC = 3 };

struct MMConvParams {
    const int32_t *rowptr, *col, *perm, *gid, *row_map;
    int64_t n_rows, E, E_total;
    const float *P, *Q, *R, *keep;
    int64_t ldp, ldq, ldr, ldk;
    Dropout drop;
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

struct GroupCtx {
    int64_t row;            // CSR row handled by this group
    int c;                  // first (global) column of this lane
    int s;                  // lane index inside the group
    int L;                  // group size
    bool live;              // lane owns real columns of a real row (ghost lanes only help with
                            // index loads, shuffles and the RNG: every warp runs warp-uniform loops
                            // so all shuffles use the constant full mask)
    int beg, deg, wdeg;     // row start, row length, max row length over the warp
};

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ void locate(const MMConvParams &p, GroupCtx &g, int vec) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t group = tid >> p.lanes_log2;
    const bool real = group < p.n_groups;
    g.L = 1 << p.lanes_log2;
    g.s = (int)(tid & (g.L - 1));
    g.row = real ? group / p.chunks : 0;
    const int chunk = real ? (int)(group - g.row * p.chunks) : 0;
    g.c = p.col0 + ((chunk << p.lanes_log2) + g.s) * vec;
    g.live = real && g.c < p.col0 + p.ncols;
    g.beg = 0; g.deg = 0;
    if (real) {
        g.beg = __ldg(p.rowptr + g.row);
        g.deg = __ldg(p.rowptr + g.row + 1) - g.beg;
    }
    g.wdeg = __reduce_max_sync(kFull, g.deg);
}

// message of one edge for this lane's columns, in the reference's arithmetic order
// (before the keep-scale): ((P[dst] + Q[src]) + R[e])
template <int VEC, int MSG>
__device__ __forceinline__ Vec<VEC> message(const MMConvParams &p, const Vec<VEC> &pv, const Vec<VEC> &q,
                                            const Vec<VEC> &r) {
    Vec<VEC> m;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        float x;
        if constexpr (MSG == MSG_PQ) x = __fadd_rn(pv.v[v], q.v[v]);
        else if constexpr (MSG == MSG_PQR) x = __fadd_rn(__fadd_rn(pv.v[v], q.v[v]), r.v[v]);
        else if constexpr (MSG == MSG_R) x = r.v[v];
        else {
            if (p.P && p.Q) x = __fadd_rn(pv.v[v], q.v[v]);
            else if (p.P) x = pv.v[v];
            else if (p.Q) x = q.v[v];
            else x = r.v[v];
            if (p.R && (p.P || p.Q)) x = __fadd_rn(x, r.v[v]);
        }
        m.v[v] = x;
    }
    return m;
}

// Dropout bits are generated cooperatively by the group: for a batch of 4 edges, lane s runs
// Philox for (edge s / (L/4), 16-column block s % (L/4)); word k of that call belongs to the lane
// 4*block + k.  Requires VEC == 4 and L >= 4 with every group starting on a 16-column boundary.
__device__ __forceinline__ uint4 philox_shared_generate(const MMConvParams &p, const GroupCtx &g,
                                                        const int (&ge)[4]) {
    const int per_edge = g.L >> 2;                      // calls (16-column blocks) per edge
    const int u_mine = g.s / per_edge, b_mine = g.s - u_mine * per_edge;
    const int e_mine = u_mine == 0 ? ge[0] : (u_mine == 1 ? ge[1] : (u_mine == 2 ? ge[2] : ge[3]));
    const int c_blk = ((g.c - g.s * 4) >> 4) + b_mine;  // global 16-column block of my call
    return philox4x32_10(make_uint4((uint32_t)e_mine, (uint32_t)c_blk, 0u, 0u), make_uint2(p.drop.k0, p.drop.k1));
}

__device__ __forceinline__ Vec<4> philox_shared_fetch(const MMConvParams &p, const GroupCtx &g, const uint4 &bits,
                                                      int u) {
    const int src = u * (g.L >> 2) + (g.s >> 2);
    const uint32_t w0 = __shfl_sync(kFull, bits.x, src, g.L);
    const uint32_t w1 = __shfl_sync(kFull, bits.y, src, g.L);
    const uint32_t w2 = __shfl_sync(kFull, bits.z, src, g.L);
    const uint32_t w3 = __shfl_sync(kFull, bits.w, src, g.L);
    const int wsel = g.s & 3;
    const uint32_t w = wsel == 0 ? w0 : (wsel == 1 ? w1 : (wsel == 2 ? w2 : w3));
    return keep_from_word<4>(p.drop, w, 0);
}

// Visits the in-edges of the group's row in order, 4 at a time: indices by shuffle, then all
// gathers of the batch issued, then the edges consumed one by one.  Loops run to the warp-wide
// maximum row length (rows are degree-sorted, so this costs nothing in practice); `valid` masks
// the edges beyond the group's own row.
//   consume(valid, pos, global_edge_id, m, ks): m is the message BEFORE the keep-scale.
template <int VEC, int MSG, int DROP, bool NEED_M, bool NEED_EID, typename Consume>
__device__ __forceinline__ void for_each_edge(const MMConvParams &p, const GroupCtx &g, const Vec<VEC> &pv,
                                              Consume &&consume) {
    constexpr bool kShared = DROP == DROP_PHILOX_SHARED && VEC == 4;
    constexpr bool kNeedQ = NEED_M && MSG != MSG_R;
    constexpr bool kNeedR = NEED_M && MSG != MSG_PQ;
    constexpr bool kNeedE = NEED_EID || DROP != DROP_NONE || kNeedR;
    const bool has_q = kNeedQ && (MSG != MSG_GENERIC || p.Q != nullptr);
    const bool has_r = kNeedR && (MSG != MSG_GENERIC || p.R != nullptr);
    const char *Qc = reinterpret_cast<const char *>(p.Q + g.c);
    const char *Rc = reinterpret_cast<const char *>(p.R + g.c);
    const char *Kc = reinterpret_cast<const char *>(p.keep + g.c);
    const uint32_t ldq_b = (uint32_t)p.ldq * 4u, ldr_b = (uint32_t)p.ldr * 4u, ldk_b = (uint32_t)p.ldk * 4u;
    for (int base = 0; base < g.wdeg; base += g.L) {
        int my_j = 0, my_e = g.beg + base + g.s, my_g;
        const bool mine = base + g.s < g.deg;
        if (mine) {
            if (has_q) my_j = __ldg(p.col + g.beg + base + g.s);
            if (kNeedE && p.perm) my_e = __ldg(p.perm + g.beg + base + g.s);
        }
        my_g = my_e;
        if (kNeedE && p.gid && mine) my_g = __ldg(p.gid + my_e);
        const int nbw = min(g.L, g.wdeg - base);
        for (int k = 0; k < nbw; k += 4) {
            int ge[4];
            bool ok[4];
            Vec<VEC> q[kNeedQ ? 4 : 1], r[kNeedR ? 4 : 1], kp[DROP == DROP_KEEP ? 4 : 1];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ok[u] = g.live && (base + k + u < g.deg);
                const int j = has_q ? __shfl_sync(kFull, my_j, k + u, g.L) : 0;
                const int e = kNeedE ? __shfl_sync(kFull, my_e, k + u, g.L) : 0;
                ge[u] = (kNeedE && p.gid) ? __shfl_sync(kFull, my_g, k + u, g.L) : e;
                if constexpr (kNeedQ) {
                    if (ok[u] && has_q) q[u] = ld_vec_stream<VEC>(reinterpret_cast<const float *>(Qc + (uint64_t)(uint32_t)j * ldq_b));
                }
                if constexpr (kNeedR) {
                    if (ok[u] && has_r) r[u] = ld_vec_stream<VEC>(reinterpret_cast<const float *>(Rc + (uint64_t)(uint32_t)e * ldr_b));
                }
                if constexpr (DROP == DROP_KEEP) {
                    if (ok[u]) kp[u] = ld_vec_stream<VEC>(reinterpret_cast<const float *>(Kc + (uint64_t)(uint32_t)e * ldk_b));
                }
            }
            uint4 bits = make_uint4(0u, 0u, 0u, 0u);
            if constexpr (kShared) bits = philox_shared_generate(p, g, ge);       // one Philox call per lane per batch
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                Vec<VEC> ks{};
                if constexpr (kShared) {
                    if constexpr (VEC == 4) ks = philox_shared_fetch(p, g, bits, u);   // every lane of the warp takes part
                } else if constexpr (DROP == DROP_PHILOX_LANE || DROP == DROP_PHILOX_SHARED) {
                    ks = dropout_keep<VEC>(p.drop, (uint32_t)ge[u], g.c, 0u);
                } else if constexpr (DROP == DROP_KEEP) {
                    ks = kp[u];
                }
                Vec<VEC> m{};
                if constexpr (NEED_M) m = message<VEC, MSG>(p, pv, q[kNeedQ ? u : 0], r[kNeedR ? u : 0]);
                consume(ok[u], g.beg + base + k + u, ge[u], m, ks);
            }
        }
    }
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
template <int VEC, int MSG, int DROP, bool MINMAX, bool SQ>
__global__ void __launch_bounds__(256, 3) mmconv_fwd_kernel(const __grid_constant__ MMConvParams p) {
    GroupCtx g;
    locate(p, g, VEC);
    const int64_t row = g.row;
    const int c = g.c;

    Vec<VEC> pv{};
    if (MSG != MSG_R && g.live && p.P) {
        const int64_t prow = p.row_map ? (int64_t)__ldg(p.row_map + row) : row;
        pv = ld_vec<VEC>(p.P + prow * p.ldp + c);
    }

    float sum[VEC], sq[VEC], mn[VEC], mx[VEC];
    int amn[VEC], amx[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        sum[v] = 0.0f; sq[v] = 0.0f; mn[v] = FLT_MAX; mx[v] = -FLT_MAX; amn[v] = -1; amx[v] = -1;
    }

    for_each_edge<VEC, MSG, DROP, true, false>(p, g, pv,
        [&](bool ok, int pos, int, const Vec<VEC> &m, const Vec<VEC> &ks) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float x = m.v[v];
                if constexpr (DROP != DROP_NONE) x = __fmul_rn(x, ks.v[v]);   // x * 0 keeps the sign of x, like F.dropout
                if (ok) sum[v] = __fadd_rn(sum[v], x);
                if constexpr (SQ) { if (ok) sq[v] = __fadd_rn(sq[v], __fmul_rn(x, x)); }
                if constexpr (MINMAX) {
                    if (ok && x < mn[v]) { mn[v] = x; amn[v] = pos; }   // strict: first occurrence wins,
                    if (ok && x > mx[v]) { mx[v] = x; amx[v] = pos; }   // -0.0 == +0.0, NaN never wins
                }
            }
        });
    if (!g.live) return;

    // ---- epilogue: aggregates -> cumulative scalers -> Y[row, t, (s*A+a)*F_in + f] ----
    const int degc = g.deg > 1 ? g.deg : 1;                // deg.clamp_(1), mma_conv.py:179
    const float degf = (float)degc;
    float fac[MMA_MAX_SCALER];
    scaler_factors(p, degc, fac);

    Vec<VEC> mean, var, sd, vmin, vmax, vsum;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        vsum.v[v] = sum[v];
        mean.v[v] = __fdiv_rn(sum[v], degf);               // sum / count.clamp(min=1)
        if constexpr (SQ) {
            const float msq = __fdiv_rn(sq[v], degf);
            var.v[v] = __fsub_rn(msq, __fmul_rn(mean.v[v], mean.v[v]));         // mma_conv.py:170, no FMA
            sd.v[v] = sqrtf(__fadd_rn(fmaxf(var.v[v], 0.0f), 1e-5f));           // mma_conv.py:172
        } else {
            var.v[v] = 0.0f; sd.v[v] = 0.0f;
        }
        vmin.v[v] = (MINMAX && amn[v] >= 0) ? mn[v] : 0.0f;                      // empty row -> 0
        vmax.v[v] = (MINMAX && amx[v] >= 0) ? mx[v] : 0.0f;
    }

    const int t = c / p.F_in, f = c - t * p.F_in;
    float *yrow = p.Y + row * p.ldy + (int64_t)t * ((int64_t)p.S * p.A * p.F_in) + f;
    for (int a = 0; a < p.A; ++a) {
        const int kind = p.akind[a];
        Vec<VEC> val = kind == MMA_AGGR_SUM ? vsum : kind == MMA_AGGR_MEAN ? mean : kind == MMA_AGGR_MIN ? vmin
                     : kind == MMA_AGGR_MAX ? vmax : kind == MMA_AGGR_VAR ? var : sd;
        for (int s = 0; s < p.S; ++s) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) val.v[v] = __fmul_rn(val.v[v], fac[s]);     // cumulative (Q4)
            st_vec_stream<VEC>(yrow + (int64_t)(s * p.A + a) * p.F_in, val);
        }
    }

    if constexpr (MINMAX) {
        int32_t o_mn[VEC], o_mx[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            o_mn[v] = amn[v] < 0 ? (int32_t)p.E_total : orig_edge_id(p, amn[v]);
            o_mx[v] = amx[v] < 0 ? (int32_t)p.E_total : orig_edge_id(p, amx[v]);
        }
        if (p.arg_min) st_vec_i32_stream<VEC>(p.arg_min + row * p.F + c, o_mn);
        if (p.arg_max) st_vec_i32_stream<VEC>(p.arg_max + row * p.F + c, o_mx);
    }
    if (p.stat_mean) st_vec_stream<VEC>(p.stat_mean + row * p.F + c, mean);
    if constexpr (SQ) {
        if (p.stat_var) st_vec_stream<VEC>(p.stat_var + row * p.F + c, var);
    }
}

// ----------------------------------------------------------------------------------------
// backward, destination pass: per-edge gradient rows G and dP
// ----------------------------------------------------------------------------------------
template <int VEC, int MSG, int DROP, bool NEEDM>
__global__ void __launch_bounds__(256, 3) mmconv_bwd_dst_kernel(const __grid_constant__ MMConvParams p) {
    GroupCtx g;
    locate(p, g, VEC);
    const int64_t row = g.row;
    const int c = g.c;
    const int64_t prow = (g.live && p.row_map) ? (int64_t)__ldg(p.row_map + row) : row;
    const int degc = g.deg > 1 ? g.deg : 1;
    const float degf = (float)degc;

    Vec<VEC> base{}, gmin{}, gmax{}, alpha{}, pv{};
    int32_t amn[VEC], amx[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { amn[v] = -1; amx[v] = -1; }

    if (g.live) {
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
                const float gg = dz.v[v];
                if (kind == MMA_AGGR_SUM) base.v[v] += gg;
                else if (kind == MMA_AGGR_MEAN) base.v[v] += gg / degf;
                else if (kind == MMA_AGGR_MIN) gmin.v[v] += gg;
                else if (kind == MMA_AGGR_MAX) gmax.v[v] += gg;
                else if (kind == MMA_AGGR_VAR) {    // var = E[m^2] - E[m]^2 -> d/dm_e = 2 (m_e - mean) / cnt
                    const float k = 2.0f * gg / degf;
                    alpha.v[v] += k; base.v[v] -= k * mean.v[v];
                } else if (var.v[v] > 0.0f) {       // std = sqrt(relu(var) + 1e-5); relu'(0) = 0
                    const float k = gg / (sqrtf(var.v[v] + 1e-5f) * degf);
                    alpha.v[v] += k; base.v[v] -= k * mean.v[v];
                }
            }
            has_min |= kind == MMA_AGGR_MIN;
            has_max |= kind == MMA_AGGR_MAX;
        }
        if (has_min) ld_vec_i32_as<VEC>(p.c_arg_min + row * p.F + c, amn);
        if (has_max) ld_vec_i32_as<VEC>(p.c_arg_max + row * p.F + c, amx);
        if (NEEDM && MSG != MSG_R && p.P) pv = ld_vec<VEC>(p.P + prow * p.ldp + c);
    }

    Vec<VEC> dp{};
    float *Gc = p.G ? p.G + c : nullptr;
    for_each_edge<VEC, MSG, DROP, NEEDM, true>(p, g, pv,
        [&](bool ok, int pos, int eid, const Vec<VEC> &m, const Vec<VEC> &ks) {
            Vec<VEC> gr;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float x = base.v[v];
                if (eid == amn[v]) x += gmin.v[v];
                if (eid == amx[v]) x += gmax.v[v];
                if constexpr (NEEDM) {
                    float mm = m.v[v];
                    if constexpr (DROP != DROP_NONE) mm = __fmul_rn(mm, ks.v[v]);
                    x += alpha.v[v] * mm;
                }
                if constexpr (DROP != DROP_NONE) x *= ks.v[v];     // dL/dm_pre = dL/dm * keepscale
                gr.v[v] = x;
                if (ok) dp.v[v] += x;
            }
            if (ok && Gc) {
                const int64_t slot = p.gslot ? (int64_t)__ldg(p.gslot + pos) : (int64_t)pos;
                st_vec<VEC>(Gc + slot * p.ldg, gr);
            }
        });
    if (p.dP && g.live) st_vec_stream<VEC>(p.dP + prow * p.lddp + c, dp);
}

// ----------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------
static int fill_params(MMConvParams &p, const int32_t *rowptr, const int32_t *col, const int32_t *perm,
                       const int32_t *gid, int64_t E_total, const int32_t *row_map, int64_t n_rows, int64_t E,
                       const float *P, int64_t ldp, const float *Q, int64_t ldq,
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
    p.rowptr = rowptr; p.col = col; p.perm = perm; p.gid = gid; p.row_map = row_map; p.n_rows = n_rows; p.E = E;
    p.E_total = gid ? E_total : E;
    p.P = P; p.Q = Q; p.R = R; p.keep = keep; p.ldp = ldp; p.ldq = ldq; p.ldr = ldr; p.ldk = ldk;
    p.drop = make_dropout(p_drop, seed);
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
    int lg = 2;                                      // groups of >= 4 lanes (shared index loads / RNG)
    while ((1 << lg) < per_row && lg < 5) ++lg;
    p.lanes_log2 = lg;
    const int lanes = 1 << lg;
    p.chunks = (per_row + lanes - 1) / lanes;
    p.n_groups = p.n_rows * p.chunks;
    return vec;
}

static inline bool ok4(const void *ptr, int64_t ld) { return ptr == nullptr || (aligned16(ptr) && (ld % 4) == 0); }

static int drop_mode(const MMConvParams &p, const float *keep, float p_drop, int vec) {
    if (keep) return DROP_KEEP;
    if (p_drop <= 0.0f) return DROP_NONE;
    // the shared generator needs each group's first column on a 16-column block boundary
    const bool blocks_aligned = (p.col0 % 16 == 0) && (((1 << p.lanes_log2) * 4) % 16 == 0);
    return (vec == 4 && blocks_aligned) ? DROP_PHILOX_SHARED : DROP_PHILOX_LANE;
}

}  // namespace mma

using namespace mma;

static int msg_mode(const float *P, const float *Q, const float *R) {
    if (P && Q && !R) return MSG_PQ;
    if (P && Q && R) return MSG_PQR;
    if (!P && !Q && R) return MSG_R;
    return MSG_GENERIC;
}

extern "C" int mmconv_aggregate_fwd(const int32_t *rowptr, const int32_t *col, const int32_t *perm,
                                    const int32_t *edge_gid, int64_t E_total, const int32_t *row_map,
                                    int64_t n_rows, int64_t E, const float *P, int64_t ldp,
                                    const float *Q, int64_t ldq, const float *R, int64_t ldr,
                                    const float *keep, int64_t ldk, float p_drop, uint64_t seed,
                                    int T, int F_in, int A, const int32_t *aggr_kinds, int S,
                                    const int32_t *scaler_kinds, const float *scale_tab, int64_t tab_stride,
                                    float *Y, int64_t ldy, int32_t *arg_min, int32_t *arg_max,
                                    float *stat_mean, float *stat_var, int col0, int ncols,
                                    mma_stream_t stream) {
    MMConvParams p;
    int rc = fill_params(p, rowptr, col, perm, edge_gid, E_total, row_map, n_rows, E, P, ldp, Q, ldq, R, ldr,
                         keep, ldk, p_drop, seed, T, F_in, A, aggr_kinds, S, scaler_kinds, scale_tab, tab_stride);
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
    const int drop = drop_mode(p, keep, p_drop, vec);
    const int64_t threads = p.n_groups << p.lanes_log2;
    const int block = 256;
    const int64_t grid = (threads + block - 1) / block;
    if (grid > INT32_MAX) return MMA_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int msg = msg_mode(P, Q, R);
    bool launched = false;
#define FWD(V, M, D, MM, SQ) (mmconv_fwd_kernel<V, M, D, MM, SQ><<<(unsigned)grid, block, 0, st>>>(p), launched = true)
    // fast kernels: 128-bit path, message mode and dropout mode fixed at compile time
#define FWD_FAST(M, D)                                       \
    do {                                                     \
        if (minmax && sq) FWD(4, M, D, true, true);          \
        else if (minmax) FWD(4, M, D, true, false);          \
        else if (sq) FWD(4, M, D, false, true);              \
        else FWD(4, M, D, false, false);                     \
    } while (0)
    if (vec == 4 && msg != MSG_GENERIC && (drop == DROP_NONE || drop == DROP_PHILOX_SHARED)) {
        if (msg == MSG_PQ) { if (drop == DROP_NONE) FWD_FAST(MSG_PQ, DROP_NONE); else FWD_FAST(MSG_PQ, DROP_PHILOX_SHARED); }
        else if (msg == MSG_PQR) { if (drop == DROP_NONE) FWD_FAST(MSG_PQR, DROP_NONE); else FWD_FAST(MSG_PQR, DROP_PHILOX_SHARED); }
        else { if (drop == DROP_NONE) FWD_FAST(MSG_R, DROP_NONE); else FWD_FAST(MSG_R, DROP_PHILOX_SHARED); }
    }
    if (!launched) {    // generic kernels (runtime null checks, all accumulators): test / odd-shape paths
        if (vec == 4) {
            switch (drop) {
                case DROP_NONE: FWD(4, MSG_GENERIC, DROP_NONE, true, true); break;
                case DROP_KEEP: FWD(4, MSG_GENERIC, DROP_KEEP, true, true); break;
                case DROP_PHILOX_SHARED: FWD(4, MSG_GENERIC, DROP_PHILOX_SHARED, true, true); break;
                default: FWD(4, MSG_GENERIC, DROP_PHILOX_LANE, true, true); break;
            }
        } else {
            switch (drop) {
                case DROP_NONE: FWD(1, MSG_GENERIC, DROP_NONE, true, true); break;
                case DROP_KEEP: FWD(1, MSG_GENERIC, DROP_KEEP, true, true); break;
                default: FWD(1, MSG_GENERIC, DROP_PHILOX_LANE, true, true); break;
            }
        }
    }
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}

extern "C" int mmconv_aggregate_bwd_dst(const int32_t *rowptr, const int32_t *col, const int32_t *perm,
                                        const int32_t *edge_gid, int64_t E_total, const int32_t *row_map,
                                        int64_t n_rows, int64_t E, const float *P, int64_t ldp,
                                        const float *Q, int64_t ldq, const float *R, int64_t ldr,
                                        const float *keep, int64_t ldk, float p_drop, uint64_t seed,
                                        int T, int F_in, int A, const int32_t *aggr_kinds, int S,
                                        const int32_t *scaler_kinds, const float *scale_tab, int64_t tab_stride,
                                        const float *dY, int64_t ldy, const int32_t *arg_min,
                                        const int32_t *arg_max, const float *stat_mean, const float *stat_var,
                                        const int32_t *gslot, float *G, int64_t ldg, float *dP, int64_t lddp,
                                        int col0, int ncols, mma_stream_t stream) {
    MMConvParams p;
    int rc = fill_params(p, rowptr, col, perm, edge_gid, E_total, row_map, n_rows, E, P, ldp, Q, ldq, R, ldr,
                         keep, ldk, p_drop, seed, T, F_in, A, aggr_kinds, S, scaler_kinds, scale_tab, tab_stride);
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
    const int drop = drop_mode(p, keep, p_drop, vec);
    const int64_t threads = p.n_groups << p.lanes_log2;
    const int block = 256;
    const int64_t grid = (threads + block - 1) / block;
    if (grid > INT32_MAX) return MMA_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int msg = msg_mode(P, Q, R);
    bool launched = false;
#define BWD(V, M, D, NM) (mmconv_bwd_dst_kernel<V, M, D, NM><<<(unsigned)grid, block, 0, st>>>(p), launched = true)
    if (vec == 4 && (drop == DROP_NONE || drop == DROP_PHILOX_SHARED)) {
        if (!needm) {       // the message itself is not needed: no gather at all
            if (drop == DROP_NONE) BWD(4, MSG_PQ, DROP_NONE, false); else BWD(4, MSG_PQ, DROP_PHILOX_SHARED, false);
        } else if (msg == MSG_PQ) {
            if (drop == DROP_NONE) BWD(4, MSG_PQ, DROP_NONE, true); else BWD(4, MSG_PQ, DROP_PHILOX_SHARED, true);
        } else if (msg == MSG_PQR) {
            if (drop == DROP_NONE) BWD(4, MSG_PQR, DROP_NONE, true); else BWD(4, MSG_PQR, DROP_PHILOX_SHARED, true);
        } else if (msg == MSG_R) {
            if (drop == DROP_NONE) BWD(4, MSG_R, DROP_NONE, true); else BWD(4, MSG_R, DROP_PHILOX_SHARED, true);
        }
    }
    if (!launched) {
        if (vec == 4) {
            switch (drop) {
                case DROP_NONE: if (needm) BWD(4, MSG_GENERIC, DROP_NONE, true); else BWD(4, MSG_GENERIC, DROP_NONE, false); break;
                case DROP_KEEP: if (needm) BWD(4, MSG_GENERIC, DROP_KEEP, true); else BWD(4, MSG_GENERIC, DROP_KEEP, false); break;
                case DROP_PHILOX_SHARED: if (needm) BWD(4, MSG_GENERIC, DROP_PHILOX_SHARED, true); else BWD(4, MSG_GENERIC, DROP_PHILOX_SHARED, false); break;
                default: if (needm) BWD(4, MSG_GENERIC, DROP_PHILOX_LANE, true); else BWD(4, MSG_GENERIC, DROP_PHILOX_LANE, false); break;
            }
        } else {
            switch (drop) {
                case DROP_NONE: if (needm) BWD(1, MSG_GENERIC, DROP_NONE, true); else BWD(1, MSG_GENERIC, DROP_NONE, false); break;
                case DROP_KEEP: if (needm) BWD(1, MSG_GENERIC, DROP_KEEP, true); else BWD(1, MSG_GENERIC, DROP_KEEP, false); break;
                default: if (needm) BWD(1, MSG_GENERIC, DROP_PHILOX_LANE, true); else BWD(1, MSG_GENERIC, DROP_PHILOX_LANE, false); break;
            }
        }
    }
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}
