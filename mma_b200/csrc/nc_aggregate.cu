// K2: masked multi-aggregator layer of node_classification, all A aggregators in one pass.
//
// Replaces the per-node Python loops of /root/reference/node_classification/layers.py
// (learnable_sum*/mean*/max*/min*/softmax/softmin, :201-728): for every node i the reference
// gathers x_i and its neighbours (:209-214), multiplies [x_i || x_j] by the 2F x F mask
// matrix (:215-216), applies sigmoid (or nothing, Q8) and ALWAYS-ON dropout (:217-219),
// sums mask * x_j (:221) and combines with x_i (:221 / :328-329 / :452 / :562).
// The mask matmul is separable: logits = x_i M[:F] + x_j M[F:] = PA[i] + QA[j]; PA/QA are
// node-level GEMMs done by the caller, so the per-edge work is one add + activation.
//
// Same thread mapping as K1: LANES threads own (node, chunk of LANES*VEC columns), walk the
// neighbour list sequentially, accumulate A masked sums in registers.  Backward = dst pass
// (dPA, gS, direct dX) + transpose pass over the CSC (dQA, dX through x_j); both recompute the
// mask from PA/QA and the Philox stream instead of storing [A,E,F]; no atomics.
#include <math_constants.h>

#include "common.cuh"

namespace mma {

struct NcParams {
    const int32_t *ptr, *idx, *eid;     // CSR (ptr,idx) or CSC (ptr,idx=row,eid)
    int64_t N, E;
    const float *X, *PA, *QA;
    int64_t ldx, ldpa, ldqa;
    int F, A;
    int act[MMA_MAX_AGGR], comb[MMA_MAX_AGGR];
    const float *keep;
    Dropout drop;
    int use_philox;
    float *OUT, *S_out;                 // fwd: [A,N,F]
    const float *S_saved, *dOUT;        // bwd dst: [A,N,F]
    float *gS_w;                        // bwd dst out: [N, A*F]
    const float *gS_r;                  // bwd src in
    float *dXdir, *dPA, *dQA, *dXnbr;
    int64_t lddpa, lddqa, lddx;
    int lanes_log2, chunks;
    int64_t n_groups;
};

// One WARP per (row, chunk of columns).  The L = 2^lanes_log2 lanes that cover the chunk's columns form a
// sub-group; the 32 / L sub-groups of the warp take the row's neighbours round-robin (sub-group s: neighbours
// s, s + nsub, ...), and their partial sums are combined by a fixed xor-butterfly at the end -- a fixed
// summation tree, so results stay bit-reproducible.  These graphs are small and L2 resident: the kernel lasts as
// long as its longest neighbour list, which this cuts by the number of sub-groups (8x at F = 16).
struct NcLane { int64_t row; int c, sub, nsub, L; bool live; };
__device__ __forceinline__ NcLane nc_locate(const NcParams &p, int vec) {
    NcLane t;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    t.L = 1 << p.lanes_log2;
    t.nsub = 32 >> p.lanes_log2;
    t.sub = lane >> p.lanes_log2;
    const bool real = warp < p.n_groups;
    t.row = real ? warp / p.chunks : 0;
    const int chunk = real ? (int)(warp - t.row * p.chunks) : 0;
    t.c = ((chunk << p.lanes_log2) + (lane & (t.L - 1))) * vec;
    t.live = real && t.c < p.F;
    if (!t.live) t.c = 0;               // idle lanes run along on valid addresses; they never store
    return t;
}
template <int VEC>
__device__ __forceinline__ void nc_combine(Vec<VEC> &a, int L) {      // sum over the warp's sub-groups
    for (int off = L; off < 32; off <<= 1) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) a.v[v] += __shfl_xor_sync(0xffffffffu, a.v[v], off);
    }
}

// neighbours in flight per batch, bounded by the registers the gathered rows need (rows = rows per neighbour)
__host__ __device__ constexpr int nc_batch(int rows) { return rows <= 2 ? 8 : (rows <= 5 ? 4 : 2); }

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int VEC>
__device__ __forceinline__ Vec<VEC> nc_keep(const NcParams &p, int a, int e, int c, bool &has) {
    Vec<VEC> k;
    has = true;
    if (p.keep) k = ld_vec_stream<VEC>(p.keep + ((int64_t)a * p.E + e) * p.F + c);
    else if (p.use_philox) k = dropout_keep<VEC>(p.drop, (uint32_t)e, c, (uint32_t)a);
    else { has = false;
#pragma unroll
        for (int v = 0; v < VEC; ++v) k.v[v] = 1.0f; }
    return k;
}

// ---------------------------------------------------------------------------- forward
template <int VEC, int A>
__global__ void __launch_bounds__(256) nc_fwd_kernel(const __grid_constant__ NcParams p) {
    const NcLane t = nc_locate(p, VEC);
    const int64_t row = t.row;
    const int c = t.c;
    const int beg = __ldg(p.ptr + row), end = __ldg(p.ptr + row + 1);
    const Vec<VEC> xi = ld_vec<VEC>(p.X + row * p.ldx + c);
    Vec<VEC> pa[A], S[A];
#pragma unroll
    for (int a = 0; a < A; ++a) {
        pa[a] = ld_vec<VEC>(p.PA + row * p.ldpa + a * p.F + c);
#pragma unroll
        for (int v = 0; v < VEC; ++v) S[a].v[v] = 0.0f;
    }
    // The graphs of this path are small and L2 resident: the kernel's duration is the latency chain of its
    // longest neighbour list.  Neighbours are therefore taken U at a time -- all indices, then all gathered rows
    // of the batch in flight together -- and consumed in list order (same summation order as before).
    auto edge = [&](int pos, const Vec<VEC> &xj, const Vec<VEC> (&qa)[A]) {
#pragma unroll
        for (int a = 0; a < A; ++a) {
            bool has;
            const Vec<VEC> ks = nc_keep<VEC>(p, a, pos, c, has);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float l = __fadd_rn(pa[a].v[v], qa[a].v[v]);
                float mk = p.act[a] == MMA_ACT_SIGMOID ? sigmoidf(l) : l;
                if (has) mk = __fmul_rn(mk, ks.v[v]);
                S[a].v[v] = __fadd_rn(S[a].v[v], __fmul_rn(mk, xj.v[v]));
            }
        }
    };
    constexpr int U = nc_batch(A);
    const int st = t.nsub;
    int pos = beg + t.sub;
    for (; pos + (U - 1) * st < end; pos += U * st) {
        int64_t j[U];
        Vec<VEC> xj[U], qa[U][A];
#pragma unroll
        for (int u = 0; u < U; ++u) j[u] = __ldg(p.idx + pos + u * st);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            xj[u] = ld_vec_stream<VEC>(p.X + j[u] * p.ldx + c);
#pragma unroll
            for (int a = 0; a < A; ++a) qa[u][a] = ld_vec_stream<VEC>(p.QA + j[u] * p.ldqa + a * p.F + c);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) edge(pos + u * st, xj[u], qa[u]);
    }
    for (; pos < end; pos += st) {
        const int64_t j = __ldg(p.idx + pos);
        const Vec<VEC> xj = ld_vec_stream<VEC>(p.X + j * p.ldx + c);
        Vec<VEC> qa[A];
#pragma unroll
        for (int a = 0; a < A; ++a) qa[a] = ld_vec_stream<VEC>(p.QA + j * p.ldqa + a * p.F + c);
        edge(pos, xj, qa);
    }
#pragma unroll
    for (int a = 0; a < A; ++a) nc_combine<VEC>(S[a], t.L);
    if (!t.live || t.sub != 0) return;
    const float D = (float)(end - beg);     // len(add_all[i]); D = 0 divides by zero like the reference (Q9)
#pragma unroll
    for (int a = 0; a < A; ++a) {
        Vec<VEC> o;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const float x = xi.v[v], s = S[a].v[v];
            float r;
            switch (p.comb[a]) {
                case MMA_NC_SUM: r = __fadd_rn(x, s); break;
                case MMA_NC_MEAN: r = __fdiv_rn(__fadd_rn(x, s), D); break;
                case MMA_NC_MAX: r = (x != x || s != s) ? CUDART_NAN_F : (x > s ? x : s); break;
                case MMA_NC_MIN: r = (x != x || s != s) ? CUDART_NAN_F : (x < s ? x : s); break;
                default: r = s; break;
            }
            o.v[v] = r;
        }
        st_vec<VEC>(p.OUT + ((int64_t)a * p.N + row) * p.F + c, o);
        if (p.S_out) st_vec<VEC>(p.S_out + ((int64_t)a * p.N + row) * p.F + c, S[a]);
    }
}

// ---------------------------------------------------------------------------- backward, dst pass
template <int VEC, int A>
__global__ void __launch_bounds__(256) nc_bwd_dst_kernel(const __grid_constant__ NcParams p) {
    const NcLane t = nc_locate(p, VEC);
    const int64_t row = t.row;
    const int c = t.c;
    const bool writer = t.live && t.sub == 0;
    const int beg = __ldg(p.ptr + row), end = __ldg(p.ptr + row + 1);
    const float D = (float)(end - beg);
    const Vec<VEC> xi = ld_vec<VEC>(p.X + row * p.ldx + c);
    Vec<VEC> pa[A], gS[A], dpa[A], dx{};
#pragma unroll
    for (int a = 0; a < A; ++a) {
        pa[a] = ld_vec<VEC>(p.PA + row * p.ldpa + a * p.F + c);
        const Vec<VEC> s = ld_vec<VEC>(p.S_saved + ((int64_t)a * p.N + row) * p.F + c);
        const Vec<VEC> g = ld_vec<VEC>(p.dOUT + ((int64_t)a * p.N + row) * p.F + c);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const float x = xi.v[v], go = g.v[v];
            float gs, gx;
            switch (p.comb[a]) {
                case MMA_NC_SUM: gs = go; gx = go; break;
                case MMA_NC_MEAN: gs = go / D; gx = go / D; break;
                case MMA_NC_MAX:     // torch.max(a,b) backward: ties split the gradient evenly
                    if (x > s.v[v]) { gx = go; gs = 0.0f; } else if (x < s.v[v]) { gx = 0.0f; gs = go; }
                    else { gx = 0.5f * go; gs = 0.5f * go; }
                    break;
                case MMA_NC_MIN:
                    if (x < s.v[v]) { gx = go; gs = 0.0f; } else if (x > s.v[v]) { gx = 0.0f; gs = go; }
                    else { gx = 0.5f * go; gs = 0.5f * go; }
                    break;
                default: gs = go; gx = 0.0f; break;
            }
            gS[a].v[v] = gs;
            dx.v[v] += gx;
            dpa[a].v[v] = 0.0f;
        }
        if (writer) st_vec<VEC>(p.gS_w + row * ((int64_t)A * p.F) + a * p.F + c, gS[a]);
    }
    if (writer) st_vec<VEC>(p.dXdir + row * p.lddx + c, dx);
    auto edge = [&](int pos, const Vec<VEC> &xj, const Vec<VEC> (&qa)[A]) {
#pragma unroll
        for (int a = 0; a < A; ++a) {
            bool has;
            const Vec<VEC> ks = nc_keep<VEC>(p, a, pos, c, has);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float l = pa[a].v[v] + qa[a].v[v];
                float d = gS[a].v[v] * xj.v[v] * ks.v[v];           // dL/dmask_pre-dropout
                if (p.act[a] == MMA_ACT_SIGMOID) { const float sg = sigmoidf(l); d *= sg * (1.0f - sg); }
                dpa[a].v[v] += d;
            }
        }
    };
    constexpr int U = nc_batch(A);
    const int st = t.nsub;
    int pos = beg + t.sub;
    for (; pos + (U - 1) * st < end; pos += U * st) {
        int64_t j[U];
        Vec<VEC> xj[U], qa[U][A];
#pragma unroll
        for (int u = 0; u < U; ++u) j[u] = __ldg(p.idx + pos + u * st);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            xj[u] = ld_vec_stream<VEC>(p.X + j[u] * p.ldx + c);
#pragma unroll
            for (int a = 0; a < A; ++a) qa[u][a] = ld_vec_stream<VEC>(p.QA + j[u] * p.ldqa + a * p.F + c);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) edge(pos + u * st, xj[u], qa[u]);
    }
    for (; pos < end; pos += st) {
        const int64_t j = __ldg(p.idx + pos);
        const Vec<VEC> xj = ld_vec_stream<VEC>(p.X + j * p.ldx + c);
        Vec<VEC> qa[A];
#pragma unroll
        for (int a = 0; a < A; ++a) qa[a] = ld_vec_stream<VEC>(p.QA + j * p.ldqa + a * p.F + c);
        edge(pos, xj, qa);
    }
#pragma unroll
    for (int a = 0; a < A; ++a) nc_combine<VEC>(dpa[a], t.L);
    if (!writer) return;
#pragma unroll
    for (int a = 0; a < A; ++a) st_vec<VEC>(p.dPA + row * p.lddpa + a * p.F + c, dpa[a]);
}

// ---------------------------------------------------------------------------- backward, src pass
template <int VEC, int A>
__global__ void __launch_bounds__(256) nc_bwd_src_kernel(const __grid_constant__ NcParams p) {
    const NcLane t = nc_locate(p, VEC);
    const int64_t j = t.row;
    const int c = t.c;
    const int beg = __ldg(p.ptr + j), end = __ldg(p.ptr + j + 1);
    const Vec<VEC> xj = ld_vec<VEC>(p.X + j * p.ldx + c);
    Vec<VEC> qa[A], dqa[A], dx{};
#pragma unroll
    for (int a = 0; a < A; ++a) {
        qa[a] = ld_vec<VEC>(p.QA + j * p.ldqa + a * p.F + c);
#pragma unroll
        for (int v = 0; v < VEC; ++v) dqa[a].v[v] = 0.0f;
    }
    auto edge = [&](int e, const Vec<VEC> (&pa)[A], const Vec<VEC> (&gs)[A]) {
#pragma unroll
        for (int a = 0; a < A; ++a) {
            bool has;
            const Vec<VEC> ks = nc_keep<VEC>(p, a, e, c, has);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float l = pa[a].v[v] + qa[a].v[v];
                const bool sig = p.act[a] == MMA_ACT_SIGMOID;
                const float sg = sig ? sigmoidf(l) : l;
                const float mk = sg * ks.v[v];
                dx.v[v] += gs[a].v[v] * mk;                          // through x_j in mask * x_j
                float d = gs[a].v[v] * xj.v[v] * ks.v[v];
                if (sig) d *= sg * (1.0f - sg);
                dqa[a].v[v] += d;
            }
        }
    };
    constexpr int U = nc_batch(2 * A);
    const int st = t.nsub;
    int k = beg + t.sub;
    for (; k + (U - 1) * st < end; k += U * st) {
        int64_t i[U];
        int e[U];
        Vec<VEC> pa[U][A], gs[U][A];
#pragma unroll
        for (int u = 0; u < U; ++u) { i[u] = __ldg(p.idx + k + u * st); e[u] = __ldg(p.eid + k + u * st); }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int a = 0; a < A; ++a) {
                pa[u][a] = ld_vec_stream<VEC>(p.PA + i[u] * p.ldpa + a * p.F + c);
                gs[u][a] = ld_vec_stream<VEC>(p.gS_r + i[u] * ((int64_t)A * p.F) + a * p.F + c);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) edge(e[u], pa[u], gs[u]);
    }
    for (; k < end; k += st) {
        const int64_t i = __ldg(p.idx + k);
        const int e = __ldg(p.eid + k);
        Vec<VEC> pa[A], gs[A];
#pragma unroll
        for (int a = 0; a < A; ++a) {
            pa[a] = ld_vec_stream<VEC>(p.PA + i * p.ldpa + a * p.F + c);
            gs[a] = ld_vec_stream<VEC>(p.gS_r + i * ((int64_t)A * p.F) + a * p.F + c);
        }
        edge(e, pa, gs);
    }
#pragma unroll
    for (int a = 0; a < A; ++a) nc_combine<VEC>(dqa[a], t.L);
    nc_combine<VEC>(dx, t.L);
    if (!t.live || t.sub != 0) return;
#pragma unroll
    for (int a = 0; a < A; ++a) st_vec<VEC>(p.dQA + j * p.lddqa + a * p.F + c, dqa[a]);
    st_vec<VEC>(p.dXnbr + j * p.lddx + c, dx);
}

static int nc_fill(NcParams &p, const int32_t *ptr, const int32_t *idx, int64_t N, int64_t E, const float *X,
                   int64_t ldx, const float *PA, int64_t ldpa, const float *QA, int64_t ldqa, int F, int A,
                   const int32_t *act, const int32_t *comb, const float *keep, float p_drop, uint64_t seed,
                   const uint64_t *seed_dev) {
    if (!ptr || (E > 0 && !idx) || !X || !PA || !QA || N < 0 || E < 0 || F < 1 || A < 1 || !act) return MMA_ERR_INVALID;
    if (A > MMA_MAX_AGGR) return MMA_ERR_UNSUPPORTED;
    if (N >= INT32_MAX || E >= INT32_MAX) return MMA_ERR_UNSUPPORTED;
    if (p_drop < 0.0f || p_drop > 1.0f) return MMA_ERR_INVALID;
    p = NcParams{};
    p.ptr = ptr; p.idx = idx; p.N = N; p.E = E; p.X = X; p.PA = PA; p.QA = QA;
    p.ldx = ldx; p.ldpa = ldpa; p.ldqa = ldqa; p.F = F; p.A = A;
    for (int a = 0; a < A; ++a) {
        if (act[a] != MMA_ACT_SIGMOID && act[a] != MMA_ACT_RAW) return MMA_ERR_INVALID;
        p.act[a] = act[a];
        if (comb) {
            if (comb[a] < MMA_NC_SUM || comb[a] > MMA_NC_NONE) return MMA_ERR_INVALID;
            p.comb[a] = comb[a];
        }
    }
    p.keep = keep;
    p.drop = make_dropout(p_drop, seed, seed_dev);
    p.use_philox = (!keep && p_drop > 0.0f) ? 1 : 0;
    return MMA_OK;
}

static int nc_geometry(NcParams &p, bool v4) {
    const int vec = v4 ? 4 : 1;
    const int per_row = (p.F + vec - 1) / vec;
    int lg = 0;
    while ((1 << lg) < per_row && lg < 5) ++lg;
    p.lanes_log2 = lg;
    p.chunks = (per_row + (1 << lg) - 1) >> lg;
    p.n_groups = p.N * p.chunks;
    return vec;
}

static inline bool ok4(const void *ptr, int64_t ld) { return ptr == nullptr || (aligned16(ptr) && (ld % 4) == 0); }

#define NC_DISPATCH(KERNEL)                                                              \
    do {                                                                                 \
        const int block = 256;                                                           \
        const int64_t grid = (p.n_groups * 32 + block - 1) / block;  /* one warp per group */ \
        if (grid > INT32_MAX) return MMA_ERR_UNSUPPORTED;                                \
        cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);                        \
        switch (vec * 16 + p.A) {                                                        \
            case 4 * 16 + 1: KERNEL<4, 1><<<(unsigned)grid, block, 0, st>>>(p); break;   \
            case 4 * 16 + 2: KERNEL<4, 2><<<(unsigned)grid, block, 0, st>>>(p); break;   \
            case 4 * 16 + 3: KERNEL<4, 3><<<(unsigned)grid, block, 0, st>>>(p); break;   \
            case 4 * 16 + 4: KERNEL<4, 4><<<(unsigned)grid, block, 0, st>>>(p); break;   \
            case 4 * 16 + 5: KERNEL<4, 5><<<(unsigned)grid, block, 0, st>>>(p); break;   \
            case 4 * 16 + 6: KERNEL<4, 6><<<(unsigned)grid, block, 0, st>>>(p); break;   \
            case 4 * 16 + 7: KERNEL<4, 7><<<(unsigned)grid, block, 0, st>>>(p); break;   \
            case 4 * 16 + 8: KERNEL<4, 8><<<(unsigned)grid, block, 0, st>>>(p); break;   \
            case 1 * 16 + 1: KERNEL<1, 1><<<(unsigned)grid, block, 0, st>>>(p); break;   \
            case 1 * 16 + 2: KERNEL<1, 2><<<(unsigned)grid, block, 0, st>>>(p); break;   \
            case 1 * 16 + 3: KERNEL<1, 3><<<(unsigned)grid, block, 0, st>>>(p); break;   \
            case 1 * 16 + 4: KERNEL<1, 4><<<(unsigned)grid, block, 0, st>>>(p); break;   \
            case 1 * 16 + 5: KERNEL<1, 5><<<(unsigned)grid, block, 0, st>>>(p); break;   \
            case 1 * 16 + 6: KERNEL<1, 6><<<(unsigned)grid, block, 0, st>>>(p); break;   \
            case 1 * 16 + 7: KERNEL<1, 7><<<(unsigned)grid, block, 0, st>>>(p); break;   \
            case 1 * 16 + 8: KERNEL<1, 8><<<(unsigned)grid, block, 0, st>>>(p); break;   \
            default: return MMA_ERR_UNSUPPORTED;                                         \
        }                                                                                \
        MMA_LAUNCH_CHECK();                                                              \
    } while (0)

}  // namespace mma

using namespace mma;

extern "C" int mma_nc_aggregate_fwd(const int32_t *rowptr, const int32_t *col, int64_t N, int64_t E,
                                    const float *X, int64_t ldx, const float *PA, int64_t ldpa,
                                    const float *QA, int64_t ldqa, int F, int A, const int32_t *act_kinds,
                                    const int32_t *comb_kinds, const float *keep, float p_drop, uint64_t seed, const uint64_t *seed_dev,
                                    float *OUT, float *S_out, mma_stream_t stream) {
    NcParams p;
    int rc = nc_fill(p, rowptr, col, N, E, X, ldx, PA, ldpa, QA, ldqa, F, A, act_kinds, comb_kinds, keep, p_drop, seed,
                     seed_dev);
    if (rc != MMA_OK) return rc;
    if (!OUT || !comb_kinds) return MMA_ERR_INVALID;
    p.OUT = OUT; p.S_out = S_out;
    if (N == 0) return MMA_OK;
    const bool v4 = F % 4 == 0 && ok4(X, ldx) && ok4(PA, ldpa) && ok4(QA, ldqa) && ok4(keep, 4) && ok4(OUT, 4) && ok4(S_out, 4);
    const int vec = nc_geometry(p, v4);
    NC_DISPATCH(nc_fwd_kernel);
    return MMA_OK;
}

extern "C" int mma_nc_aggregate_bwd_dst(const int32_t *rowptr, const int32_t *col, int64_t N, int64_t E,
                                        const float *X, int64_t ldx, const float *PA, int64_t ldpa,
                                        const float *QA, int64_t ldqa, int F, int A, const int32_t *act_kinds,
                                        const int32_t *comb_kinds, const float *keep, float p_drop, uint64_t seed, const uint64_t *seed_dev,
                                        const float *S_saved, const float *dOUT, float *gS, float *dXdir,
                                        float *dPA, int64_t lddpa, mma_stream_t stream) {
    NcParams p;
    int rc = nc_fill(p, rowptr, col, N, E, X, ldx, PA, ldpa, QA, ldqa, F, A, act_kinds, comb_kinds, keep, p_drop, seed,
                     seed_dev);
    if (rc != MMA_OK) return rc;
    if (!comb_kinds || !S_saved || !dOUT || !gS || !dXdir || !dPA) return MMA_ERR_INVALID;
    p.S_saved = S_saved; p.dOUT = dOUT; p.gS_w = gS; p.dXdir = dXdir; p.dPA = dPA; p.lddpa = lddpa; p.lddx = F;
    if (N == 0) return MMA_OK;
    const bool v4 = F % 4 == 0 && ok4(X, ldx) && ok4(PA, ldpa) && ok4(QA, ldqa) && ok4(keep, 4) &&
                    ok4(S_saved, 4) && ok4(dOUT, 4) && ok4(gS, 4) && ok4(dXdir, 4) && ok4(dPA, lddpa);
    const int vec = nc_geometry(p, v4);
    NC_DISPATCH(nc_bwd_dst_kernel);
    return MMA_OK;
}

extern "C" int mma_nc_aggregate_bwd_src(const int32_t *colptr, const int32_t *row, const int32_t *eid,
                                        int64_t N, int64_t E, const float *X, int64_t ldx, const float *PA,
                                        int64_t ldpa, const float *QA, int64_t ldqa, int F, int A,
                                        const int32_t *act_kinds, const float *keep, float p_drop, uint64_t seed, const uint64_t *seed_dev,
                                        const float *gS, float *dQA, int64_t lddqa, float *dXnbr, int64_t lddx,
                                        mma_stream_t stream) {
    NcParams p;
    int rc = nc_fill(p, colptr, row, N, E, X, ldx, PA, ldpa, QA, ldqa, F, A, act_kinds, nullptr, keep, p_drop, seed,
                     seed_dev);
    if (rc != MMA_OK) return rc;
    if ((E > 0 && !eid) || !gS || !dQA || !dXnbr) return MMA_ERR_INVALID;
    p.eid = eid; p.gS_r = gS; p.dQA = dQA; p.lddqa = lddqa; p.dXnbr = dXnbr; p.lddx = lddx;
    if (N == 0) return MMA_OK;
    const bool v4 = F % 4 == 0 && ok4(X, ldx) && ok4(PA, ldpa) && ok4(QA, ldqa) && ok4(keep, 4) &&
                    ok4(gS, 4) && ok4(dQA, lddqa) && ok4(dXnbr, lddx);
    const int vec = nc_geometry(p, v4);
    NC_DISPATCH(nc_bwd_src_kernel);
    return MMA_OK;
}
