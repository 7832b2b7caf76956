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
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace mma {

enum { DROP_NONE = 0, DROP_KEEP = 1, DROP_PHILOX = 2 };     // generic kernels
enum { FD_NONE = 0, FD_BIT = 1, FD_BYTE = 2 };              // fast kernels: no dropout / 1 bit (p = 0.5) / 1 byte per element
// which terms make up the message (compile time, so the hot loop carries no null checks)
enum { MSG_PQ = 0, MSG_PQR = 1, MSG_R = 2, MSG_GENERIC = 3 };

struct MMConvParams {
    const int32_t *rowptr, *col, *perm, *gid, *row_map, *rng_row, *row_chunks;
    int64_t n_rows, E, E_total, rng_row0, n_chunks;
    // long rows cut into segments (stream kernels only, see mmconv_stream.cuh)
    const int32_t *vrowptr, *seg_tab, *split_tab;
    int64_t n_vrows, n_split;
    float *seg_ws;
    int use_rng, args_local;
    int simple_out;          // S == 1 and every aggregator kind at most once: zoff[kind] = its column offset in a Y row (-1: absent)
    int zoff[6];
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
    int wmin;               // min row length over the warp's real groups: batches below it need no masking
    uint32_t rid;           // id of the row in the dropout stream (global destination node id)
};

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ void locate(const MMConvParams &p, GroupCtx &g, int vec) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t group = tid >> p.lanes_log2;
    const bool real = group < p.n_groups;
    g.L = 1 << p.lanes_log2;
    g.s = (int)(tid & (g.L - 1));
    g.row = !real ? 0 : (p.chunks == 1 ? group : group / p.chunks);
    const int chunk = (real && p.chunks > 1) ? (int)(group - g.row * p.chunks) : 0;
    g.c = p.col0 + ((chunk << p.lanes_log2) + g.s) * vec;
    g.live = real && g.c < p.col0 + p.ncols;
    g.beg = 0; g.deg = 0;
    if (real) {
        g.beg = __ldg(p.rowptr + g.row);
        g.deg = __ldg(p.rowptr + g.row + 1) - g.beg;
    }
    g.wdeg = __reduce_max_sync(kFull, g.deg);
    g.wmin = __reduce_min_sync(kFull, real ? g.deg : 0x7fffffff);
    g.rid = 0;
    if (real && p.use_rng) g.rid = (uint32_t)(p.rng_row0 + (p.rng_row ? (int64_t)__ldg(p.rng_row + g.row) : g.row));
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

// Visits the in-edges of the group's row in order, 4 at a time: indices by shuffle, then all
// gathers of the batch issued, then the edges consumed one by one.  Loops run to the warp-wide
// maximum row length (rows are degree-sorted, so this costs nothing in practice); `valid` masks
// the edges beyond the group's own row.
//   consume(valid, pos, global_edge_id, m, ks): m is the message BEFORE the keep-scale.
template <int VEC, int MSG, int DROP, bool NEED_M, bool NEED_EID, typename Consume>
__device__ __forceinline__ void for_each_edge(const MMConvParams &p, const GroupCtx &g, const Vec<VEC> &pv,
                                              Consume &&consume) {
    constexpr bool kNeedQ = NEED_M && MSG != MSG_R;
    constexpr bool kNeedR = NEED_M && MSG != MSG_PQ;
    constexpr bool kNeedE = NEED_EID || DROP == DROP_KEEP || kNeedR;
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
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                Vec<VEC> ks{};
                if constexpr (DROP == DROP_PHILOX) {
                    ks = dropout_keep_row<VEC>(p.drop, g.rid, (uint32_t)(base + k + u), g.c);
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
            o_mn[v] = amn[v] < 0 ? (int32_t)p.E_total : (p.args_local ? amn[v] : orig_edge_id(p, amn[v]));
            o_mx[v] = amx[v] < 0 ? (int32_t)p.E_total : (p.args_local ? amx[v] : orig_edge_id(p, amx[v]));
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
        [&](bool ok, int pos, int geid, const Vec<VEC> &m, const Vec<VEC> &ks) {
            const int eid = p.args_local ? pos : geid;
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
// Fast path: 128-bit columns (VEC = 4), messages P+Q or P+Q+R, dropout none / in-kernel Philox.
//
// The hot loop is branch-free and kept to ~13 instructions per element: predicated 128-bit
// gathers, 8 rows in flight per group (two register batches of 4 edges, the next batch issued
// before the current one is consumed), dropout bits generated by the lane that consumes them
// (one Philox call per 32 edges in the 1-bit mode), and for p = 0.5 the keep-scale 2 folded into
// the message: fma(q, 2, 2p) == 2 * (p + q) exactly (scaling by a power of two commutes with
// rounding), so a kept element costs one FFMA.  A dropped element is x * 0 = +-0 in the
// reference: it cannot change a running sum (which is never -0) and enters min/max as a signed
// zero, exactly as F.dropout leaves it.
// ----------------------------------------------------------------------------------------
template <bool NQ, bool NR, bool NSLOT, bool NEID>
struct Batch {                       // 4 consecutive in-edges of one row, in registers
    Vec<4> q[NQ ? 4 : 1];            // Q[src]
    Vec<4> r[NR ? 4 : 1];            // R[original edge id]
    int slot[NSLOT ? 4 : 1];         // row of G (backward)
    int eid[NEID ? 4 : 1];           // global original edge id (backward with public arg indices)
};

__device__ __forceinline__ void ld_row4_pred(Vec<4> &r, const void *ptr, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3])
                 : "l"(ptr), "r"((int)pred));
}

// Walks the group's row 4 edges at a time, software-pipelined: consume(pos, batch, checked) sees
// the edges [pos, pos+4) of the row (pos relative to the row start).  checked = false_type: every
// group of the warp has all 4 edges (pos + 4 <= g.wmin), nothing to mask; true_type: lanes beyond
// their row see stale registers and must mask with pos + u < g.deg.  All loops are warp-uniform
// (trip counts from g.wdeg / g.wmin), so the index shuffles use the full mask.
struct Unchecked { static constexpr bool value = false; };
struct Checked { static constexpr bool value = true; };

template <bool NQ, bool NR, bool NSLOT, bool NEID, typename Consume>
__device__ __forceinline__ void walk_row(const MMConvParams &p, const GroupCtx &g, Consume &&consume) {
    using B = Batch<NQ, NR, NSLOT, NEID>;
    constexpr bool kNeedE = NR || NEID;
    if (g.wdeg <= 0) return;
    const char *Qc = reinterpret_cast<const char *>(p.Q + g.c);
    const char *Rc = reinterpret_cast<const char *>(p.R + g.c);
    const uint32_t ldq_b = (uint32_t)p.ldq * 4u, ldr_b = (uint32_t)p.ldr * 4u;
    const int Lm = g.L - 1;
    int my_j = 0, my_e = 0, my_g = 0, my_s = 0;

    auto load_idx = [&](int pos) {                 // lane s takes in-row edge pos + s (coalesced)
        const int i = pos + g.s;
        if (i < g.deg) {
            const int at = g.beg + i;
            if constexpr (NQ) my_j = __ldg(p.col + at);
            if constexpr (kNeedE) my_e = p.perm ? __ldg(p.perm + at) : at;
            if constexpr (NSLOT) my_s = p.gslot ? __ldg(p.gslot + at) : at;
            if constexpr (NEID) my_g = p.gid ? __ldg(p.gid + my_e) : my_e;
        }
    };
    auto issue = [&](int pos, B &b) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int l = (pos + u) & Lm;
            const bool ok = g.live & (pos + u < g.deg);
            if constexpr (NQ) {
                const int j = __shfl_sync(kFull, my_j, l, g.L);
                ld_row4_pred(b.q[u], Qc + (uint64_t)(uint32_t)j * ldq_b, ok);
            }
            if constexpr (NR) {
                const int e = __shfl_sync(kFull, my_e, l, g.L);
                ld_row4_pred(b.r[u], Rc + (uint64_t)(uint32_t)e * ldr_b, ok);
            }
            if constexpr (NSLOT) b.slot[u] = __shfl_sync(kFull, my_s, l, g.L);
            if constexpr (NEID) b.eid[u] = __shfl_sync(kFull, my_g, l, g.L);
        }
    };
    auto eat = [&](int pos, const B &b) {
        if (pos + 4 <= g.wmin) consume(pos, b, Unchecked{});
        else consume(pos, b, Checked{});
    };

    B b0, b1;
    load_idx(0);
    issue(0, b0);
    for (int pos = 0; pos < g.wdeg; pos += 8) {
        const bool more = pos + 4 < g.wdeg;
        if (more) {
            if (((pos + 4) & Lm) == 0) load_idx(pos + 4);
            issue(pos + 4, b1);
        }
        eat(pos, b0);
        if (pos + 8 < g.wdeg) {
            if (((pos + 8) & Lm) == 0) load_idx(pos + 8);
            issue(pos + 8, b0);
        }
        if (more) eat(pos + 4, b1);
    }
}

// dropout words of a batch: w[v] holds, for column c + v, bit u (FD_BIT) or byte u (FD_BYTE) of edge pos + u
template <int DROP>
__device__ __forceinline__ void batch_rng(const MMConvParams &p, const GroupCtx &g, int pos, uint32_t (&bits)[4],
                                          uint32_t (&w)[4]) {
    if constexpr (DROP == FD_BIT) {
        if ((pos & 31) == 0) {
            const uint4 t = row_rng_bits1(p.drop, g.rid, (uint32_t)pos, (uint32_t)g.c);
            bits[0] = t.x; bits[1] = t.y; bits[2] = t.z; bits[3] = t.w;
        }
        const int sh = pos & 31;
#pragma unroll
        for (int v = 0; v < 4; ++v) w[v] = bits[v] >> sh;
    } else if constexpr (DROP == FD_BYTE) {
        const uint4 t = row_rng_bits8(p.drop, g.rid, (uint32_t)pos, (uint32_t)g.c);
        w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
    } else {
        w[0] = w[1] = w[2] = w[3] = 0u;
    }
}

// non-zero iff the element (edge u of the batch) is kept
template <int DROP>
__device__ __forceinline__ uint32_t keep_word(uint32_t w, int u, uint32_t thr) {
    if constexpr (DROP == FD_BIT) return w & (1u << u);
    else if constexpr (DROP == FD_BYTE) return ((w >> (8 * u)) & 0xFFu) >= thr ? 1u : 0u;
    else return 1u;
}

// message of the fast path AFTER the keep-scale (for a kept element), reference rounding order.
// pv is pre-doubled in the 1-bit mode.
template <int MSG, int DROP>
__device__ __forceinline__ float fast_message(float pvv, float q, float r, float scale) {
    if constexpr (DROP == FD_BIT) {
        float x = __fmaf_rn(q, 2.0f, pvv);                       // == 2 * (P + Q)
        if constexpr (MSG == MSG_PQR) x = __fmaf_rn(r, 2.0f, x); // == 2 * ((P + Q) + R)
        return x;
    } else {
        float x = __fadd_rn(pvv, q);
        if constexpr (MSG == MSG_PQR) x = __fadd_rn(x, r);
        if constexpr (DROP == FD_BYTE) x = __fmul_rn(x, scale);
        return x;
    }
}

// One element into the running aggregates, as predicated straight-line code (the compiler turns
// the equivalent C++ selects into a branch per element).  x: message of a KEPT element; kbi != 0:
// kept; oki != 0: the edge exists (CHECK only); at: CSR slot of the edge.
//   sum/sq: a dropped element is +-0 and cannot change a running sum -> the add is predicated off.
//   min/max: a dropped element takes part as sign(x) * 0; strict compare: first occurrence wins,
//   -0.0 == +0.0, NaN never wins.
#define MMA_ACC_OPERANDS                                                                                         \
    : "+f"(sum), "+f"(sq), "+f"(mn), "+f"(mx), "+r"(amn), "+r"(amx) : "f"(x), "r"(kbi), "r"(oki), "r"(at)
#define MMA_ACC_HEAD "{\n\t.reg .pred kb, kp, ok, lt, gt;\n\t.reg .f32 xx, xm;\n\tsetp.ne.b32 kb, %7, 0;\n\t"
#define MMA_ACC_CHK "setp.ne.b32 ok, %8, 0;\n\tand.pred kp, kb, ok;\n\t"
#define MMA_ACC_SUM(KP) "@" KP " add.rn.f32 %0, %0, %6;\n\t"
// x*x is rounded BEFORE it is added, like the reference's scatter(inputs * inputs): var = E[x^2] - E[x]^2 cancels
// to ~0 on rows of (nearly) equal messages and std = sqrt(relu(var) + 1e-5) magnifies a 1-ulp difference in the sum
// of squares by ~10^2 there -- an FFMA here (tried) breaks the 1e-5 bar on std and its gradient.
#define MMA_ACC_SQ(KP) "mul.rn.f32 xx, %6, %6;\n\t@" KP " add.rn.f32 %1, %1, xx;\n\t"
#define MMA_ACC_XM_DROP "and.b32 xm, %6, 0x80000000;\n\tselp.f32 xm, %6, xm, kb;\n\t"
#define MMA_ACC_XM_NODROP "mov.f32 xm, %6;\n\t"
#define MMA_ACC_CMP_U "setp.lt.f32 lt, xm, %2;\n\tsetp.gt.f32 gt, xm, %3;\n\t"
#define MMA_ACC_CMP_C "setp.lt.and.f32 lt, xm, %2, ok;\n\tsetp.gt.and.f32 gt, xm, %3, ok;\n\t"
#define MMA_ACC_SEL                                                                                              \
    "selp.f32 %2, xm, %2, lt;\n\tselp.b32 %4, %9, %4, lt;\n\tselp.f32 %3, xm, %3, gt;\n\tselp.b32 %5, %9, %5, gt;\n\t"

template <bool CHECK, bool DROPZ, bool MINMAX, bool SQ>
__device__ __forceinline__ void accumulate(float x, uint32_t kbi, uint32_t oki, int at, float &sum, float &sq,
                                           float &mn, float &mx, int &amn, int &amx) {
#define MMA_ACC_VARIANT(SQS, MMS_U, MMS_C)                                                                        \
    do {                                                                                                         \
        if constexpr (CHECK) asm(MMA_ACC_HEAD MMA_ACC_CHK MMA_ACC_SUM("kp") SQS("kp") MMS_C "}" MMA_ACC_OPERANDS); \
        else asm(MMA_ACC_HEAD MMA_ACC_SUM("kb") SQS("kb") MMS_U "}" MMA_ACC_OPERANDS);                             \
    } while (0)
#define MMA_ACC_NOSQ(KP) ""
    if constexpr (MINMAX && SQ) {
        if constexpr (DROPZ) MMA_ACC_VARIANT(MMA_ACC_SQ, MMA_ACC_XM_DROP MMA_ACC_CMP_U MMA_ACC_SEL, MMA_ACC_XM_DROP MMA_ACC_CMP_C MMA_ACC_SEL);
        else MMA_ACC_VARIANT(MMA_ACC_SQ, MMA_ACC_XM_NODROP MMA_ACC_CMP_U MMA_ACC_SEL, MMA_ACC_XM_NODROP MMA_ACC_CMP_C MMA_ACC_SEL);
    } else if constexpr (MINMAX) {
        if constexpr (DROPZ) MMA_ACC_VARIANT(MMA_ACC_NOSQ, MMA_ACC_XM_DROP MMA_ACC_CMP_U MMA_ACC_SEL, MMA_ACC_XM_DROP MMA_ACC_CMP_C MMA_ACC_SEL);
        else MMA_ACC_VARIANT(MMA_ACC_NOSQ, MMA_ACC_XM_NODROP MMA_ACC_CMP_U MMA_ACC_SEL, MMA_ACC_XM_NODROP MMA_ACC_CMP_C MMA_ACC_SEL);
    } else if constexpr (SQ) {
        MMA_ACC_VARIANT(MMA_ACC_SQ, "", "");
    } else {
        MMA_ACC_VARIANT(MMA_ACC_NOSQ, "", "");
    }
#undef MMA_ACC_VARIANT
#undef MMA_ACC_NOSQ
}

template <int MSG, int DROP, bool MINMAX, bool SQ>
__global__ void __launch_bounds__(256, 2) mmconv_fwd_fast(const __grid_constant__ MMConvParams p) {
    GroupCtx g;
    locate(p, g, 4);
    const int64_t row = g.row;
    const int c = g.c;
    constexpr bool NR = MSG == MSG_PQR;

    Vec<4> pv{};
    if (g.live) {
        const int64_t prow = p.row_map ? (int64_t)__ldg(p.row_map + row) : row;
        pv = ld_vec<4>(p.P + prow * p.ldp + c);
    }
    if constexpr (DROP == FD_BIT) {
#pragma unroll
        for (int v = 0; v < 4; ++v) pv.v[v] *= 2.0f;
    }
    const float scale = p.drop.scale;
    const uint32_t thr = p.drop.thr;

    float sum[4], sq[4], mn[4], mx[4];
    int amn[4], amx[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        sum[v] = 0.0f; sq[v] = 0.0f; mn[v] = FLT_MAX; mx[v] = -FLT_MAX; amn[v] = -1; amx[v] = -1;
    }
    uint32_t bits[4] = {0u, 0u, 0u, 0u};

    walk_row<true, NR, false, false>(p, g, [&](int pos, const Batch<true, NR, false, false> &b, auto checked) {
        constexpr bool CHECK = decltype(checked)::value;
        uint32_t w[4];
        batch_rng<DROP>(p, g, pos, bits, w);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t oki = CHECK ? (uint32_t)(g.live & (pos + u < g.deg)) : 1u;
            const int at = g.beg + pos + u;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float x = fast_message<MSG, DROP>(pv.v[v], b.q[u].v[v], b.r[NR ? u : 0].v[v], scale);
                accumulate<CHECK, DROP != FD_NONE, MINMAX, SQ>(x, keep_word<DROP>(w[v], u, thr), oki, at, sum[v], sq[v],
                                                               mn[v], mx[v], amn[v], amx[v]);
            }
        }
    });
    if (!g.live) return;

    // ---- epilogue: aggregates -> cumulative scalers -> Y[row, t, (s*A+a)*F_in + f] ----
    const int degc = g.deg > 1 ? g.deg : 1;                // deg.clamp_(1), mma_conv.py:179
    const float degf = (float)degc;
    float fac[MMA_MAX_SCALER];
    scaler_factors(p, degc, fac);

    Vec<4> mean, var, sd, vmin, vmax, vsum;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
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

    const int t = p.T == 1 ? 0 : c / p.F_in, f = c - t * p.F_in;
    float *yrow = p.Y + row * p.ldy + (int64_t)t * ((int64_t)p.S * p.A * p.F_in) + f;
    for (int a = 0; a < p.A; ++a) {
        const int kind = p.akind[a];
        Vec<4> val = kind == MMA_AGGR_SUM ? vsum : kind == MMA_AGGR_MEAN ? mean : kind == MMA_AGGR_MIN ? vmin
                   : kind == MMA_AGGR_MAX ? vmax : kind == MMA_AGGR_VAR ? var : sd;
        for (int s = 0; s < p.S; ++s) {
#pragma unroll
            for (int v = 0; v < 4; ++v) val.v[v] = __fmul_rn(val.v[v], fac[s]);       // cumulative (Q4)
            st_vec_stream<4>(yrow + (int64_t)(s * p.A + a) * p.F_in, val);
        }
    }
    if constexpr (MINMAX) {
        int32_t o_mn[4], o_mx[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            o_mn[v] = amn[v] < 0 ? (int32_t)p.E_total : (p.args_local ? amn[v] : orig_edge_id(p, amn[v]));
            o_mx[v] = amx[v] < 0 ? (int32_t)p.E_total : (p.args_local ? amx[v] : orig_edge_id(p, amx[v]));
        }
        if (p.arg_min) st_vec_i32_stream<4>(p.arg_min + row * p.F + c, o_mn);
        if (p.arg_max) st_vec_i32_stream<4>(p.arg_max + row * p.F + c, o_mx);
    }
    if (p.stat_mean) st_vec_stream<4>(p.stat_mean + row * p.F + c, mean);
    if constexpr (SQ) {
        if (p.stat_var) st_vec_stream<4>(p.stat_var + row * p.F + c, var);
    }
}

// Backward, destination pass, fast path.  The keep-scale s (2 / 1/(1-p) / 1) is folded into the
// per-row coefficients:  dL/dh_e = s * (base + [e = argmin] gmin + [e = argmax] gmax + alpha * (s h_e))
// for a kept element, 0 for a dropped one; h_e = P + Q (+ R) is needed only for var/std (NEEDM).
//   t: s * (base + alpha * x) on entry; gv: the element's gradient; dp: running dP.
template <bool CHECK>
__device__ __forceinline__ void route_grad(float t, int id, int amn, int amx, float gmin, float gmax, uint32_t kbi,
                                           uint32_t oki, float &gv, float &dp) {
    if constexpr (CHECK)
        asm("{\n\t.reg .pred a, b, k, ok;\n\t.reg .f32 t;\n\tmov.f32 t, %2;\n\t"
            "setp.eq.s32 a, %3, %4;\n\tsetp.eq.s32 b, %3, %5;\n\t@a add.f32 t, t, %6;\n\t@b add.f32 t, t, %7;\n\t"
            "setp.ne.b32 k, %8, 0;\n\tsetp.ne.b32 ok, %9, 0;\n\tand.pred k, k, ok;\n\t"
            "selp.f32 %0, t, 0f00000000, k;\n\tadd.f32 %1, %1, %0;\n\t}"
            : "=f"(gv), "+f"(dp) : "f"(t), "r"(id), "r"(amn), "r"(amx), "f"(gmin), "f"(gmax), "r"(kbi), "r"(oki));
    else
        asm("{\n\t.reg .pred a, b, k;\n\t.reg .f32 t;\n\tmov.f32 t, %2;\n\t"
            "setp.eq.s32 a, %3, %4;\n\tsetp.eq.s32 b, %3, %5;\n\t@a add.f32 t, t, %6;\n\t@b add.f32 t, t, %7;\n\t"
            "setp.ne.b32 k, %8, 0;\n\t"
            "selp.f32 %0, t, 0f00000000, k;\n\tadd.f32 %1, %1, %0;\n\t}"
            : "=f"(gv), "+f"(dp) : "f"(t), "r"(id), "r"(amn), "r"(amx), "f"(gmin), "f"(gmax), "r"(kbi), "r"(oki));
}

template <int MSG, int DROP, bool NEEDM, bool LOCAL>
__global__ void __launch_bounds__(256, 2) mmconv_bwd_fast(const __grid_constant__ MMConvParams p) {
    GroupCtx g;
    locate(p, g, 4);
    const int64_t row = g.row;
    const int c = g.c;
    constexpr bool NR = NEEDM && MSG == MSG_PQR;
    const int64_t prow = (g.live && p.row_map) ? (int64_t)__ldg(p.row_map + row) : row;
    const int degc = g.deg > 1 ? g.deg : 1;
    const float degf = (float)degc;
    const float scale = DROP == FD_BIT ? 2.0f : (DROP == FD_BYTE ? p.drop.scale : 1.0f);
    const uint32_t thr = p.drop.thr;

    Vec<4> base{}, gmin{}, gmax{}, alpha{}, pv{};
    int32_t amn[4], amx[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) { amn[v] = -1; amx[v] = -1; }

    if (g.live) {
        float fac[MMA_MAX_SCALER];
        scaler_factors(p, degc, fac);
        for (int s = 1; s < p.S; ++s) fac[s] *= fac[s - 1];
        Vec<4> mean{}, var{};
        if constexpr (NEEDM) {
            mean = ld_vec<4>(p.c_mean + row * p.F + c);
            var = ld_vec<4>(p.c_var + row * p.F + c);
        }
        bool has_min = false, has_max = false;
        const int t = p.T == 1 ? 0 : c / p.F_in, f = c - t * p.F_in;
        const float *dyrow = p.dY + row * p.ldy + (int64_t)t * ((int64_t)p.S * p.A * p.F_in) + f;
        for (int a = 0; a < p.A; ++a) {
            Vec<4> dz{};
            for (int s = 0; s < p.S; ++s) {
                const Vec<4> d = ld_vec_stream<4>(dyrow + (int64_t)(s * p.A + a) * p.F_in);
#pragma unroll
                for (int v = 0; v < 4; ++v) dz.v[v] += d.v[v] * fac[s];
            }
            const int kind = p.akind[a];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
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
        if (has_min) ld_vec_i32_as<4>(p.c_arg_min + row * p.F + c, amn);
        if (has_max) ld_vec_i32_as<4>(p.c_arg_max + row * p.F + c, amx);
        if (NEEDM && p.P) pv = ld_vec<4>(p.P + prow * p.ldp + c);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            base.v[v] *= scale; gmin.v[v] *= scale; gmax.v[v] *= scale; alpha.v[v] *= scale;
            if constexpr (DROP == FD_BIT) pv.v[v] *= 2.0f;
        }
    }

    Vec<4> dp{};
    char *Gc = p.G ? reinterpret_cast<char *>(p.G + c) : nullptr;
    const uint32_t ldg_b = (uint32_t)p.ldg * 4u;
    uint32_t bits[4] = {0u, 0u, 0u, 0u};
    using B = Batch<NEEDM, NR, true, !LOCAL>;
    walk_row<NEEDM, NR, true, !LOCAL>(p, g, [&](int pos, const B &b, auto checked) {
        constexpr bool CHECK = decltype(checked)::value;
        uint32_t w[4];
        batch_rng<DROP>(p, g, pos, bits, w);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool ok = CHECK ? (g.live & (pos + u < g.deg)) : g.live;
            const int id = LOCAL ? g.beg + pos + u : b.eid[LOCAL ? 0 : u];
            Vec<4> gr;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                float t = base.v[v];
                if constexpr (NEEDM) {
                    const float x = fast_message<MSG, DROP>(pv.v[v], b.q[NEEDM ? u : 0].v[v], b.r[NR ? u : 0].v[v], scale);
                    t = __fmaf_rn(alpha.v[v], x, t);
                }
                route_grad<CHECK>(t, id, amn[v], amx[v], gmin.v[v], gmax.v[v], keep_word<DROP>(w[v], u, thr),
                                  (uint32_t)ok, gr.v[v], dp.v[v]);
            }
            if (ok && Gc) st_vec_stream<4>(reinterpret_cast<float *>(Gc + (uint64_t)(uint32_t)b.slot[u] * ldg_b), gr);
        }
    });
    if (p.dP && g.live) st_vec_stream<4>(p.dP + prow * p.lddp + c, dp);
}

#include "mmconv_stream.cuh"

// ----------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------
static int fill_params(MMConvParams &p, const int32_t *rowptr, const int32_t *col, const int32_t *perm,
                       const int32_t *gid, int64_t E_total, const int32_t *row_map, const int32_t *rng_row,
                       int64_t rng_row0, const int32_t *row_chunks, int64_t n_chunks,
                       const int32_t *vrowptr, int64_t n_vrows, const int32_t *seg_tab, const int32_t *split_tab,
                       int64_t n_split, float *seg_ws, int64_t n_rows, int64_t E,
                       const float *P, int64_t ldp, const float *Q, int64_t ldq,
                       const float *R, int64_t ldr, const float *keep, int64_t ldk, float p_drop, uint64_t seed,
                       const uint64_t *seed_dev,
                       int T, int F_in, int A, const int32_t *aggr_kinds, int S, const int32_t *scaler_kinds,
                       const float *scale_tab, int64_t tab_stride, int flags) {
    if (!rowptr || n_rows < 0 || E < 0 || T < 1 || F_in < 1 || A < 1 || S < 1 || !aggr_kinds || !scaler_kinds)
        return MMA_ERR_INVALID;
    if (E >= INT32_MAX || n_rows >= INT32_MAX) return MMA_ERR_UNSUPPORTED;
    if (A > MMA_MAX_AGGR || S > MMA_MAX_SCALER) return MMA_ERR_UNSUPPORTED;
    if (!P && !Q && !R && E > 0) return MMA_ERR_INVALID;
    if (Q && !col && E > 0) return MMA_ERR_INVALID;
    if (p_drop < 0.0f || p_drop > 1.0f) return MMA_ERR_INVALID;
    if (flags & ~MMA_K1_ARGS_LOCAL) return MMA_ERR_INVALID;
    p = MMConvParams{};
    p.rowptr = rowptr; p.col = col; p.perm = perm; p.gid = gid; p.row_map = row_map; p.n_rows = n_rows; p.E = E;
    p.rng_row = rng_row; p.rng_row0 = rng_row0;
    if (row_chunks && n_chunks < 1) return MMA_ERR_INVALID;
    p.row_chunks = row_chunks; p.n_chunks = row_chunks ? n_chunks : 0;
    if (vrowptr && (n_vrows < n_rows || !seg_tab || n_split < 0 || (n_split > 0 && (!split_tab || !seg_ws))))
        return MMA_ERR_INVALID;
    p.vrowptr = vrowptr; p.n_vrows = vrowptr ? n_vrows : 0; p.seg_tab = vrowptr ? seg_tab : nullptr;
    p.split_tab = vrowptr ? split_tab : nullptr; p.n_split = vrowptr ? n_split : 0; p.seg_ws = vrowptr ? seg_ws : nullptr;
    p.E_total = gid ? E_total : E;
    p.P = P; p.Q = Q; p.R = R; p.keep = keep; p.ldp = ldp; p.ldq = ldq; p.ldr = ldr; p.ldk = ldk;
    p.drop = make_dropout(p_drop, seed, seed_dev);
    p.use_rng = (!keep && p.drop.thr > 0u) ? 1 : 0;
    p.args_local = (flags & MMA_K1_ARGS_LOCAL) ? 1 : 0;
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
    p.simple_out = (S == 1 && !any_scaled) ? 1 : 0;
    for (int k = 0; k < 6; ++k) p.zoff[k] = -1;
    for (int a = 0; a < A; ++a) {
        if (p.zoff[p.akind[a]] >= 0) p.simple_out = 0;
        p.zoff[p.akind[a]] = a * F_in;
    }
    return MMA_OK;
}

static int choose_geometry(MMConvParams &p, bool vec4_ok, int col0, int ncols) {
    if (ncols <= 0) { col0 = 0; ncols = p.F; }
    p.col0 = col0; p.ncols = ncols;
    vec4_ok = vec4_ok && (col0 % 4 == 0) && (ncols % 4 == 0);
    const int vec = vec4_ok ? 4 : 1;
    const int per_row = (ncols + vec - 1) / vec;     // lanes needed for one row of the window
    int lg = 2;                                      // groups of >= 4 lanes (shared index loads)
    while ((1 << lg) < per_row && lg < 5) ++lg;
    p.lanes_log2 = lg;
    const int lanes = 1 << lg;
    p.chunks = (per_row + lanes - 1) / lanes;
    p.n_groups = p.n_rows * p.chunks;
    return vec;
}

static inline bool ok4(const void *ptr, int64_t ld) { return ptr == nullptr || (aligned16(ptr) && (ld % 4) == 0); }

static int drop_mode(const float *keep, const Dropout &d) {
    if (keep) return DROP_KEEP;
    return d.thr == 0u ? DROP_NONE : DROP_PHILOX;
}
static int fast_drop_mode(const Dropout &d) { return d.thr == 0u ? FD_NONE : (d.thr == 128u ? FD_BIT : FD_BYTE); }

}  // namespace mma

using namespace mma;

static int msg_mode(const float *P, const float *Q, const float *R) {
    if (P && Q && !R) return MSG_PQ;
    if (P && Q && R) return MSG_PQR;
    if (!P && !Q && R) return MSG_R;
    return MSG_GENERIC;
}

// ---- stream kernels: launch geometry ----
struct StreamCfg { int nst, warps; };
static StreamCfg stream_cfg() {                 // ring stages x warps per CTA; MMA_K1_STREAM = "0" (off) | "8x12" | "4x16": tuning aid
    static StreamCfg cfg = [] {
        StreamCfg c{6, 16};
        const char *e = getenv("MMA_K1_STREAM");
        if (e) {
            if (e[0] == '0') c = StreamCfg{0, 0};
            else if (!strcmp(e, "8x12")) c = StreamCfg{8, 12};
            else if (!strcmp(e, "4x16")) c = StreamCfg{4, 16};
        }
        return c;
    }();
    return cfg;
}
static int sm_count() {
    static int n = [] { int dev = 0, v = kSMs; cudaGetDevice(&dev); cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); return v; }();
    return n;
}
template <typename K>
static cudaError_t launch_stream(K kernel, const MMConvParams &p, int nst, int warps, cudaStream_t st) {
    const int smem = warps * nst * 4 * p.ncols * 4 + warps * 512;     // rings + the forward kernel's index windows (128 ints per warp)
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    kernel<<<sm_count(), warps * 32, smem, st>>>(p);
    return cudaSuccess;
}
template <int NST, int WARPS, int VEC, int DROP>
static cudaError_t launch_fwd_stream(const MMConvParams &p, bool minmax, bool sq, cudaStream_t st) {
    if (minmax && sq) return launch_stream(stream::mmconv_fwd_stream<NST, WARPS, VEC, DROP, true, true>, p, NST, WARPS, st);
    if (minmax) return launch_stream(stream::mmconv_fwd_stream<NST, WARPS, VEC, DROP, true, false>, p, NST, WARPS, st);
    if (sq) return launch_stream(stream::mmconv_fwd_stream<NST, WARPS, VEC, DROP, false, true>, p, NST, WARPS, st);
    return launch_stream(stream::mmconv_fwd_stream<NST, WARPS, VEC, DROP, false, false>, p, NST, WARPS, st);
}
template <int NST, int WARPS, int VEC, int DROP>
static cudaError_t launch_bwd_stream(const MMConvParams &p, bool needm, cudaStream_t st) {
    if (needm) {
        if (p.args_local) return launch_stream(stream::mmconv_bwd_stream<NST, WARPS, VEC, DROP, true, true>, p, NST, WARPS, st);
        return launch_stream(stream::mmconv_bwd_stream<NST, WARPS, VEC, DROP, true, false>, p, NST, WARPS, st);
    }
    if (p.args_local) return launch_stream(stream::mmconv_bwd_stream<NST, WARPS, VEC, DROP, false, true>, p, NST, WARPS, st);
    return launch_stream(stream::mmconv_bwd_stream<NST, WARPS, VEC, DROP, false, false>, p, NST, WARPS, st);
}
// columns per lane of the stream kernels: 4 (128-bit) for windows of up to 128 columns; windows of <= 64 / <= 32
// columns use 2 / 1 so that all 32 lanes of the warp stay busy (per-byte dropout modes keep 4)
static int stream_vec(const MMConvParams &p, int fd) {
    if (fd == FD_BYTE) return 4;
    return p.ncols > 64 ? 4 : (p.ncols > 32 ? 2 : 1);
}
// the stream kernels take: 128-bit columns, one warp per row window (<= 128 columns), message P + Q
static bool stream_ok(const MMConvParams &p, int vec, const float *keep) {
    return stream_cfg().nst > 0 && vec == 4 && !keep && p.chunks == 1 && p.ncols <= 128 && p.E > 0;
}

template <int MSG, int DROP>
static void launch_fwd_fast(const MMConvParams &p, bool minmax, bool sq, unsigned grid, int block, cudaStream_t st) {
    if (minmax && sq) mmconv_fwd_fast<MSG, DROP, true, true><<<grid, block, 0, st>>>(p);
    else if (minmax) mmconv_fwd_fast<MSG, DROP, true, false><<<grid, block, 0, st>>>(p);
    else if (sq) mmconv_fwd_fast<MSG, DROP, false, true><<<grid, block, 0, st>>>(p);
    else mmconv_fwd_fast<MSG, DROP, false, false><<<grid, block, 0, st>>>(p);
}

template <int MSG, int DROP>
static void launch_bwd_fast(const MMConvParams &p, bool needm, unsigned grid, int block, cudaStream_t st) {
    if (needm) {
        if (p.args_local) mmconv_bwd_fast<MSG, DROP, true, true><<<grid, block, 0, st>>>(p);
        else mmconv_bwd_fast<MSG, DROP, true, false><<<grid, block, 0, st>>>(p);
    } else {
        if (p.args_local) mmconv_bwd_fast<MSG_PQ, DROP, false, true><<<grid, block, 0, st>>>(p);
        else mmconv_bwd_fast<MSG_PQ, DROP, false, false><<<grid, block, 0, st>>>(p);
    }
}

extern "C" int mmconv_aggregate_fwd(const int32_t *rowptr, const int32_t *col, const int32_t *perm,
                                    const int32_t *edge_gid, int64_t E_total, const int32_t *row_map,
                                    const int32_t *rng_row, int64_t rng_row0,
                                    const int32_t *row_chunks, int64_t n_chunks,
                                    const int32_t *vrowptr, int64_t n_vrows, const int32_t *seg_tab,
                                    const int32_t *split_tab, int64_t n_split, float *seg_ws,
                                    int64_t n_rows, int64_t E, const float *P, int64_t ldp,
                                    const float *Q, int64_t ldq, const float *R, int64_t ldr,
                                    const float *keep, int64_t ldk, float p_drop, uint64_t seed,
                                    const uint64_t *seed_dev,
                                    int T, int F_in, int A, const int32_t *aggr_kinds, int S,
                                    const int32_t *scaler_kinds, const float *scale_tab, int64_t tab_stride,
                                    float *Y, int64_t ldy, int32_t *arg_min, int32_t *arg_max,
                                    float *stat_mean, float *stat_var, int col0, int ncols, int flags,
                                    mma_stream_t stream) {
    MMConvParams p;
    int rc = fill_params(p, rowptr, col, perm, edge_gid, E_total, row_map, rng_row, rng_row0, row_chunks, n_chunks,
                         vrowptr, n_vrows, seg_tab, split_tab, n_split, seg_ws, n_rows, E, P, ldp,
                         Q, ldq, R, ldr, keep, ldk, p_drop, seed, seed_dev, T, F_in, A, aggr_kinds, S, scaler_kinds,
                         scale_tab, tab_stride, flags);
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
    const int64_t grid64 = (threads + block - 1) / block;
    if (grid64 > INT32_MAX) return MMA_ERR_UNSUPPORTED;
    const unsigned grid = (unsigned)grid64;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int msg = msg_mode(P, Q, R);
    if (msg == MSG_PQ && stream_ok(p, vec, keep)) {
        // persistent cp.async-ring kernels (mmconv_stream.cuh)
        const int fd = fast_drop_mode(p.drop);
        const StreamCfg sc = stream_cfg();
        cudaError_t e;
        const int vw = stream_vec(p, fd);
        if (vw == 4) {
            if (fd == FD_BIT && minmax && sq && sc.nst == 8) e = launch_fwd_stream<8, 12, 4, FD_BIT>(p, minmax, sq, st);
            else if (fd == FD_BIT && minmax && sq && sc.nst == 4 && sc.warps == 16) e = launch_fwd_stream<4, 16, 4, FD_BIT>(p, minmax, sq, st);
            else if (fd == FD_NONE) e = launch_fwd_stream<6, 16, 4, FD_NONE>(p, minmax, sq, st);
            else if (fd == FD_BIT) e = launch_fwd_stream<6, 16, 4, FD_BIT>(p, minmax, sq, st);
            else e = launch_fwd_stream<6, 16, 4, FD_BYTE>(p, minmax, sq, st);
        } else if (vw == 2) {
            if (fd == FD_NONE) e = launch_fwd_stream<6, 16, 2, FD_NONE>(p, minmax, sq, st);
            else e = launch_fwd_stream<6, 16, 2, FD_BIT>(p, minmax, sq, st);
        } else {
            if (fd == FD_NONE) e = launch_fwd_stream<6, 16, 1, FD_NONE>(p, minmax, sq, st);
            else e = launch_fwd_stream<6, 16, 1, FD_BIT>(p, minmax, sq, st);
        }
        MMA_CUDA_CHECK(e);
        if (p.n_split > 0) {                            // merge the segments of the split rows, emit those rows
            const unsigned g = (unsigned)p.n_split;
            if (minmax && sq) stream::mmconv_fwd_merge<true, true><<<g, 128, 0, st>>>(p);
            else if (minmax) stream::mmconv_fwd_merge<true, false><<<g, 128, 0, st>>>(p);
            else if (sq) stream::mmconv_fwd_merge<false, true><<<g, 128, 0, st>>>(p);
            else stream::mmconv_fwd_merge<false, false><<<g, 128, 0, st>>>(p);
        }
    } else if (vec == 4 && !keep && (msg == MSG_PQ || msg == MSG_PQR)) {
        // fast kernels: 128-bit columns, message and dropout mode fixed at compile time
        const int fd = fast_drop_mode(p.drop);
        if (msg == MSG_PQ) {
            if (fd == FD_NONE) launch_fwd_fast<MSG_PQ, FD_NONE>(p, minmax, sq, grid, block, st);
            else if (fd == FD_BIT) launch_fwd_fast<MSG_PQ, FD_BIT>(p, minmax, sq, grid, block, st);
            else launch_fwd_fast<MSG_PQ, FD_BYTE>(p, minmax, sq, grid, block, st);
        } else {
            if (fd == FD_NONE) launch_fwd_fast<MSG_PQR, FD_NONE>(p, minmax, sq, grid, block, st);
            else if (fd == FD_BIT) launch_fwd_fast<MSG_PQR, FD_BIT>(p, minmax, sq, grid, block, st);
            else launch_fwd_fast<MSG_PQR, FD_BYTE>(p, minmax, sq, grid, block, st);
        }
    } else {
        // generic kernels (runtime null checks, all accumulators): explicit keep masks, materialised
        // messages, widths that are not a multiple of 4
        const int drop = drop_mode(keep, p.drop);
#define FWD(V, D) mmconv_fwd_kernel<V, MSG_GENERIC, D, true, true><<<grid, block, 0, st>>>(p)
        if (vec == 4) {
            if (drop == DROP_NONE) FWD(4, DROP_NONE); else if (drop == DROP_KEEP) FWD(4, DROP_KEEP); else FWD(4, DROP_PHILOX);
        } else {
            if (drop == DROP_NONE) FWD(1, DROP_NONE); else if (drop == DROP_KEEP) FWD(1, DROP_KEEP); else FWD(1, DROP_PHILOX);
        }
#undef FWD
    }
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}

extern "C" int mmconv_aggregate_bwd_dst(const int32_t *rowptr, const int32_t *col, const int32_t *perm,
                                        const int32_t *edge_gid, int64_t E_total, const int32_t *row_map,
                                        const int32_t *rng_row, int64_t rng_row0,
                                        const int32_t *row_chunks, int64_t n_chunks,
                                        const int32_t *vrowptr, int64_t n_vrows, const int32_t *seg_tab,
                                        const int32_t *split_tab, int64_t n_split, float *seg_ws,
                                        int64_t n_rows, int64_t E, const float *P, int64_t ldp,
                                        const float *Q, int64_t ldq, const float *R, int64_t ldr,
                                        const float *keep, int64_t ldk, float p_drop, uint64_t seed,
                                        const uint64_t *seed_dev,
                                        int T, int F_in, int A, const int32_t *aggr_kinds, int S,
                                        const int32_t *scaler_kinds, const float *scale_tab, int64_t tab_stride,
                                        const float *dY, int64_t ldy, const int32_t *arg_min,
                                        const int32_t *arg_max, const float *stat_mean, const float *stat_var,
                                        const int32_t *gslot, float *G, int64_t ldg, float *dP, int64_t lddp,
                                        int col0, int ncols, int flags, mma_stream_t stream) {
    MMConvParams p;
    int rc = fill_params(p, rowptr, col, perm, edge_gid, E_total, row_map, rng_row, rng_row0, row_chunks, n_chunks,
                         vrowptr, n_vrows, seg_tab, split_tab, n_split, seg_ws, n_rows, E, P, ldp,
                         Q, ldq, R, ldr, keep, ldk, p_drop, seed, seed_dev, T, F_in, A, aggr_kinds, S, scaler_kinds,
                         scale_tab, tab_stride, flags);
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
    const int64_t grid64 = (threads + block - 1) / block;
    if (grid64 > INT32_MAX) return MMA_ERR_UNSUPPORTED;
    const unsigned grid = (unsigned)grid64;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int msg = msg_mode(P, Q, R);
    if ((!needm || msg == MSG_PQ) && stream_ok(p, vec, keep)) {
        const int fd = fast_drop_mode(p.drop);
        const StreamCfg sc = stream_cfg();
        cudaError_t e;
        const int vw = stream_vec(p, fd);
        if (vw == 4) {
            if (fd == FD_BIT && needm && sc.nst == 8) e = launch_bwd_stream<8, 12, 4, FD_BIT>(p, needm, st);
            else if (fd == FD_BIT && needm && sc.nst == 4 && sc.warps == 16) e = launch_bwd_stream<4, 16, 4, FD_BIT>(p, needm, st);
            else if (fd == FD_NONE) e = launch_bwd_stream<6, 16, 4, FD_NONE>(p, needm, st);
            else if (fd == FD_BIT) e = launch_bwd_stream<6, 16, 4, FD_BIT>(p, needm, st);
            else e = launch_bwd_stream<6, 16, 4, FD_BYTE>(p, needm, st);
        } else if (vw == 2) {
            if (fd == FD_NONE) e = launch_bwd_stream<6, 16, 2, FD_NONE>(p, needm, st);
            else e = launch_bwd_stream<6, 16, 2, FD_BIT>(p, needm, st);
        } else {
            if (fd == FD_NONE) e = launch_bwd_stream<6, 16, 1, FD_NONE>(p, needm, st);
            else e = launch_bwd_stream<6, 16, 1, FD_BIT>(p, needm, st);
        }
        MMA_CUDA_CHECK(e);
        if (p.n_split > 0) stream::mmconv_bwd_merge<<<(unsigned)p.n_split, 128, 0, st>>>(p);
    } else if (vec == 4 && !keep && (!needm || msg == MSG_PQ || msg == MSG_PQR)) {
        const int fd = fast_drop_mode(p.drop);
        if (needm && msg == MSG_PQR) {
            if (fd == FD_NONE) launch_bwd_fast<MSG_PQR, FD_NONE>(p, needm, grid, block, st);
            else if (fd == FD_BIT) launch_bwd_fast<MSG_PQR, FD_BIT>(p, needm, grid, block, st);
            else launch_bwd_fast<MSG_PQR, FD_BYTE>(p, needm, grid, block, st);
        } else {
            if (fd == FD_NONE) launch_bwd_fast<MSG_PQ, FD_NONE>(p, needm, grid, block, st);
            else if (fd == FD_BIT) launch_bwd_fast<MSG_PQ, FD_BIT>(p, needm, grid, block, st);
            else launch_bwd_fast<MSG_PQ, FD_BYTE>(p, needm, grid, block, st);
        }
    } else {
        const int drop = drop_mode(keep, p.drop);
#define BWD(V, D) do { if (needm) mmconv_bwd_dst_kernel<V, MSG_GENERIC, D, true><<<grid, block, 0, st>>>(p); \
                       else mmconv_bwd_dst_kernel<V, MSG_GENERIC, D, false><<<grid, block, 0, st>>>(p); } while (0)
        if (vec == 4) {
            if (drop == DROP_NONE) BWD(4, DROP_NONE); else if (drop == DROP_KEEP) BWD(4, DROP_KEEP); else BWD(4, DROP_PHILOX);
        } else {
            if (drop == DROP_NONE) BWD(1, DROP_NONE); else if (drop == DROP_KEEP) BWD(1, DROP_KEEP); else BWD(1, DROP_PHILOX);
        }
#undef BWD
    }
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}

// ---- versioned argument block (include/mma_b200.h: mma_k1_args_t) ----
static int load_k1_args(const mma_k1_args_t *in, mma_k1_args_t &a) {
    constexpr size_t kV1 = offsetof(mma_k1_args_t, ncols) + sizeof(int32_t);              // the first published layout
    if (!in || in->struct_size < kV1) return MMA_ERR_INVALID;
    memset(&a, 0, sizeof(a));
    const size_t have = in->struct_size, known = sizeof(mma_k1_args_t);
    memcpy(&a, in, have < known ? have : known);
    if (have > known) {                          // a newer caller: fields this library does not know must be unset
        const unsigned char *extra = reinterpret_cast<const unsigned char *>(in) + known;
        for (size_t i = 0; i < have - known; ++i)
            if (extra[i]) return MMA_ERR_UNSUPPORTED;
    }
    return MMA_OK;
}

extern "C" int mmconv_aggregate_fwd_args(const mma_k1_args_t *args, mma_stream_t stream) {
    mma_k1_args_t a;
    const int rc = load_k1_args(args, a);
    if (rc != MMA_OK) return rc;
    return mmconv_aggregate_fwd(a.rowptr, a.col, a.perm, a.edge_gid, a.E_total, a.row_map, a.rng_row, a.rng_row0,
                                a.row_chunks, a.n_chunks, a.vrowptr, a.n_vrows, a.seg_tab, a.split_tab, a.n_split, a.seg_ws,
                                a.n_rows, a.E, a.P, a.ldp, a.Q, a.ldq, a.R, a.ldr, a.keep, a.ldk, a.p_drop, a.seed,
                                a.seed_dev, a.T, a.F_in, a.A, a.aggr_kinds, a.S, a.scaler_kinds, a.scale_tab, a.tab_stride,
                                a.Y, a.ldy, a.arg_min, a.arg_max, a.stat_mean, a.stat_var, a.col0, a.ncols, a.flags, stream);
}

extern "C" int mmconv_aggregate_bwd_dst_args(const mma_k1_args_t *args, mma_stream_t stream) {
    mma_k1_args_t a;
    const int rc = load_k1_args(args, a);
    if (rc != MMA_OK) return rc;
    return mmconv_aggregate_bwd_dst(a.rowptr, a.col, a.perm, a.edge_gid, a.E_total, a.row_map, a.rng_row, a.rng_row0,
                                    a.row_chunks, a.n_chunks, a.vrowptr, a.n_vrows, a.seg_tab, a.split_tab, a.n_split,
                                    a.seg_ws, a.n_rows, a.E, a.P, a.ldp, a.Q, a.ldq, a.R, a.ldr, a.keep, a.ldk, a.p_drop,
                                    a.seed, a.seed_dev, a.T, a.F_in, a.A, a.aggr_kinds, a.S, a.scaler_kinds, a.scale_tab,
                                    a.tab_stride, a.Y, a.ldy, a.arg_min, a.arg_max, a.stat_mean, a.stat_var, a.gslot, a.G,
                                    a.ldg, a.dP, a.lddp, a.col0, a.ncols, a.flags, stream);
}
