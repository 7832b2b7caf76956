/*
 * mma_b200 -- C ABI of the B200-native multi-mask aggregation hot path.
 *
 * The reference (asarigun/mma) is pure Python and exposes no FFI; its boundary
 * for this path is the nn.Module API.  The entry points below
 * are what a maintainer would bind from those modules (ctypes stub in
 * INTEGRATION.md); each cites the reference code it replaces (paths relative
 * to /root/reference).
 *
 * Conventions
 *   - every function returns 0 (MMA_OK) or a negative MMA_ERR_* code; nothing throws;
 *   - all pointers are DEVICE pointers unless the parameter is documented "host";
 *   - the library never allocates or frees: the caller owns inputs, outputs and
 *     workspaces (query sizes with the *_workspace_bytes functions);
 *   - no global mutable state; work is enqueued on `stream` of the CURRENT device;
 *   - fp32 data, int32 indices inside (E, N < 2^31), row-major, leading
 *     dimensions (ld*) in elements;
 *   - no atomics on any data path: every output element has exactly one
 *     owner thread and a fixed summation order => bit-reproducible runs.
 */
#ifndef MMA_B200_H
#define MMA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *mma_stream_t; /* == cudaStream_t */

#define MMA_OK 0
#define MMA_ERR_INVALID (-1)     /* bad argument (null, size, alignment, unknown kind) */
#define MMA_ERR_CUDA (-2)        /* a CUDA runtime call failed (see mma_last_cuda_error) */
#define MMA_ERR_UNSUPPORTED (-3) /* valid but outside the compiled limits (A > 8, S > 8 ...) */
#define MMA_ERR_WORKSPACE (-4)   /* workspace too small */

#define MMA_MAX_AGGR 8
#define MMA_MAX_SCALER 8

/* aggregator kinds: reduce names accepted by MMAConv.aggregate, mma_conv.py:163-174 */
enum { MMA_AGGR_SUM = 0, MMA_AGGR_MEAN = 1, MMA_AGGR_MIN = 2, MMA_AGGR_MAX = 3,
       MMA_AGGR_VAR = 4, MMA_AGGR_STD = 5 };
/* scaler kinds: mma_conv.py:181-195 (applied cumulatively, in list order) */
enum { MMA_SCALE_IDENTITY = 0, MMA_SCALE_AMPLIFICATION = 1, MMA_SCALE_ATTENUATION = 2,
       MMA_SCALE_LINEAR = 3, MMA_SCALE_INVERSE_LINEAR = 4 };
/* node-classification combine kinds: layers.py:221 / 328-329 / 452 / 562 / 676-682 */
enum { MMA_NC_SUM = 0, MMA_NC_MEAN = 1, MMA_NC_MAX = 2, MMA_NC_MIN = 3, MMA_NC_NONE = 4 };
/* node-classification mask activation: layers.py:217 (sigmoid) or the raw logit (Q8) */
enum { MMA_ACT_SIGMOID = 0, MMA_ACT_RAW = 1 };

/* library version / build info (host). */
int mma_b200_version(void);
/* cudaGetLastError() text of the last failing call on this thread (host string). */
const char *mma_last_cuda_error(void);

/* ------------------------------------------------------------------------
 * Graph preprocessing.  Replaces the implicit index plumbing of
 * MessagePassing.propagate (mma_conv.py:130) / torch_scatter's scatter-by-index
 * (mma_conv.py:166) and utils.py:98-100 (neighbour lists): a STABLE sort of the
 * edges by `key` gives a CSR whose in-row order is the original edge order, so
 * "first occurrence wins" (torch_scatter CPU) is preserved.
 *   key   [E] int64  segment id of each edge (dst for CSR-by-destination, src for the transpose)
 *   other [E] int64  the other endpoint (may be NULL -> col not written)
 *   rowptr [n_keys+1], col [E] = other[perm], perm [E] = original edge id of each slot.
 * ---------------------------------------------------------------------- */
int mma_csr_build_workspace_bytes(int64_t E, int64_t n_keys, size_t *bytes);
int mma_csr_build(const int64_t *key, const int64_t *other, int64_t E, int64_t n_keys,
                  int32_t *rowptr, int32_t *col, int32_t *perm,
                  void *workspace, size_t workspace_bytes, mma_stream_t stream);
/* inverse of a permutation: inv[perm[k]] = k  (maps original edge id -> slot). */
int mma_invert_perm(const int32_t *perm, int64_t E, int32_t *inv, mma_stream_t stream);

/* ------------------------------------------------------------------------
 * K1 forward: fused MultiMaskConv aggregate.  Replaces, in ONE pass over the
 * destination-CSR, the x_j gather of MessagePassing.propagate (mma_conv.py:130),
 * the (separable) mask linear + dropout of MMAConv.message (mma_conv.py:146-157)
 * and MMAConv.aggregate (mma_conv.py:159-196: A scatter passes, degree, the
 * cumulative scalers and both cats).
 *
 *   m[e,c] = ((P[dst(e),c] + Q[src(e),c]) + R[e,c]) * keepscale[e,c]       c in [0, T*F_in)
 *   Y[i, t, (s*A + a)*F_in + f] = aggr_a({m[e, t*F_in+f] : dst(e)=i}) * prod_{s'<=s} scale_{s'}(deg_i)
 *
 *   rowptr/col/perm : CSR by destination from mma_csr_build (perm may be NULL = identity)
 *   edge_gid [E], E_total : for one shard of a partitioned graph, the GLOBAL id of each local
 *                       edge (local original order) and the global edge count; dropout and
 *                       arg_min/arg_max then use global ids so results are shard-invariant.
 *                       NULL = the local ids are the global ids.
 *   row_map [n_rows]  : node id of each CSR row, used to address P (and dP in the backward) when
 *                       the CSR rows are a permutation of the nodes (degree-sorted rows for the
 *                       bucketed post-transform); all other per-row tensors are in CSR row order.
 *                       NULL = identity.
 *   rng_row [n_rows], rng_row0 : id of CSR row r in the dropout stream = rng_row0 + (rng_row ? rng_row[r] : r);
 *                       pass the GLOBAL destination node id (row permutation / relabelling / shard
 *                       offset undone) and the stream does not depend on how the rows are laid out.
 *   row_chunks [n_chunks+1], n_chunks : optional work partition for the persistent kernels: chunk i
 *                       is the CSR rows [row_chunks[i], row_chunks[i+1]) (ascending, row_chunks[0] = 0,
 *                       row_chunks[n_chunks] = n_rows), chunks of about equal cost (edges + ~6 per row),
 *                       many more chunks than SMs x 16 warps; they are dealt round-robin to the warps.
 *                       NULL: one chunk per warp found by a binary search over rowptr (poorer balance).
 *   vrowptr [n_vrows+1], seg_tab [n_vrows][4], split_tab [n_split][4], seg_ws : optional handling of very long
 *                       rows (skewed degree distributions) in the persistent kernels.  vrowptr is a VIRTUAL row
 *                       pointer over the same CSR slots in which every row longer than a segment length L
 *                       (a multiple of 32) is cut into ceil(deg / L) segments; seg_tab[v] = {real row, in-row
 *                       position of the segment's first edge, partial slot (numbered 0.. over all segments of
 *                       split rows, -1 for an unsplit row), 0}; split_tab[i] = {real row, first slot, number of
 *                       segments, 0}; seg_ws: workspace of 6*F floats (forward) / F floats (backward) per slot.
 *                       Segments are walked by different warps and merged in segment order by a second kernel
 *                       (min/max and their arg indices stay bit-exact).  row_chunks then partition the VIRTUAL
 *                       rows.  NULL: every row is walked by one warp.
 *   P [n_rows, F] ld ldp | Q [n_src, F] ld ldq | R [E, F] ld ldr, ORIGINAL edge order | any may be NULL
 *   keep [E, F] ld ldk : explicit keep-scale (0 or 1/(1-p)), original edge order, or NULL
 *   p_drop, seed      : if keep == NULL and p_drop > 0: in-kernel Philox4x32-10 dropout keyed by
 *                       (seed, row id, position of the edge in the row's in-edge list, column) --
 *                       identical in fwd/bwd and across shards (the CSR sort is stable, so the in-row
 *                       position is layout-invariant); p is quantised to round(256 p)/256 (exact for
 *                       the reference's 0.5 / 0.75); p = 0.5 consumes one random bit per element,
 *                       any other p one byte (mma_dropout_keep_scale_rows materialises the stream)
 *   seed_dev          : optional DEVICE pointer to the 64-bit seed; when set it overrides `seed` and is read
 *                       by the kernel at run time, so a launch captured in a CUDA graph draws a fresh mask
 *                       on every replay (advance the value between replays, e.g. inside the graph)
 *   aggr_kinds (host) [A], scaler_kinds (host) [S]
 *   scale_tab [4, tab_stride] : factor of scaler kind k (1..4) at clamped degree d is
 *                       scale_tab[(k-1)*tab_stride + d], d <= tab_stride-1 (built by the
 *                       caller with the reference's own expression, mma_conv.py:185-191);
 *                       may be NULL when all scalers are identity
 *   Y [n_rows, T, S*A*F_in] ld ldy
 *   arg_min/arg_max [n_rows, F] int32 : ORIGINAL (global) edge id of the selected edge (E_total if
 *                       none); with MMA_K1_ARGS_LOCAL in `flags`: its CSR slot instead (what the
 *                       backward of the same graph needs; saves the perm lookups).  NULL ok
 *   stat_mean/stat_var [n_rows, F] : saved for the var/std backward; NULL ok
 *   col0, ncols : process only the column window [col0, col0+ncols) of F (ncols <= 0: all columns).
 *                       Every tensor is still addressed with GLOBAL column indices, so windows of one
 *                       problem can be launched independently (feature-sliced comm/compute pipeline)
 *                       and produce exactly the unsliced result (the dropout stream is keyed by the
 *                       global column).
 * ---------------------------------------------------------------------- */
#define MMA_K1_ARGS_LOCAL 1   /* flags: arg_min/arg_max hold CSR slots, not original edge ids */

int mmconv_aggregate_fwd(const int32_t *rowptr, const int32_t *col, const int32_t *perm,
                         const int32_t *edge_gid, int64_t E_total, const int32_t *row_map,
                         const int32_t *rng_row, int64_t rng_row0,
                         const int32_t *row_chunks, int64_t n_chunks,
                         const int32_t *vrowptr, int64_t n_vrows, const int32_t *seg_tab,
                         const int32_t *split_tab, int64_t n_split, float *seg_ws,
                         int64_t n_rows, int64_t E,
                         const float *P, int64_t ldp, const float *Q, int64_t ldq,
                         const float *R, int64_t ldr, const float *keep, int64_t ldk,
                         float p_drop, uint64_t seed, const uint64_t *seed_dev,
                         int T, int F_in, int A, const int32_t *aggr_kinds,
                         int S, const int32_t *scaler_kinds,
                         const float *scale_tab, int64_t tab_stride,
                         float *Y, int64_t ldy, int32_t *arg_min, int32_t *arg_max,
                         float *stat_mean, float *stat_var, int col0, int ncols, int flags,
                         mma_stream_t stream);

/* K1 backward, destination pass (replaces autograd of mma_conv.py:157-196):
 *   G[gslot(pos), c] = dL/dm_pre[e, c]   for every edge (row of the per-edge gradient),
 *   dP[i, c]        = sum over in-edges of i of that row.
 * gslot [E] maps CSR position -> row of G (NULL = CSR position; pass `perm` to get G in
 * original edge order == dL/dR).  arg_min/arg_max/stat_* are the forward's outputs. */
int mmconv_aggregate_bwd_dst(const int32_t *rowptr, const int32_t *col, const int32_t *perm,
                             const int32_t *edge_gid, int64_t E_total, const int32_t *row_map,
                             const int32_t *rng_row, int64_t rng_row0,
                             const int32_t *row_chunks, int64_t n_chunks,
                             const int32_t *vrowptr, int64_t n_vrows, const int32_t *seg_tab,
                             const int32_t *split_tab, int64_t n_split, float *seg_ws,
                             int64_t n_rows, int64_t E,
                             const float *P, int64_t ldp, const float *Q, int64_t ldq,
                             const float *R, int64_t ldr, const float *keep, int64_t ldk,
                             float p_drop, uint64_t seed, const uint64_t *seed_dev,
                             int T, int F_in, int A, const int32_t *aggr_kinds,
                             int S, const int32_t *scaler_kinds,
                             const float *scale_tab, int64_t tab_stride,
                             const float *dY, int64_t ldy,
                             const int32_t *arg_min, const int32_t *arg_max,
                             const float *stat_mean, const float *stat_var,
                             const int32_t *gslot, float *G, int64_t ldg,
                             float *dP, int64_t lddp, int col0, int ncols, int flags, mma_stream_t stream);

/* ------------------------------------------------------------------------
 * The same two calls through ONE versioned argument block -- what a binding that is written by hand should use
 * (the positional forms above are kept for existing callers).  Set struct_size = sizeof(mma_k1_args_t) and zero
 * the rest before filling it in; fields are only ever APPENDED, a library that knows more fields than the caller's
 * struct_size covers treats them as 0 / NULL, one that knows fewer rejects a block with non-zero bytes beyond its own
 * size (MMA_ERR_UNSUPPORTED).  Field meanings are those of the positional parameters of the same name.
 * Forward reads everything up to `ncols` except gslot / G / ldg / dP / lddp and writes Y, arg_min, arg_max, stat_mean,
 * stat_var; the backward destination pass reads Y as dY and arg_min .. stat_var as the saved tensors of the forward,
 * and writes G (at gslot) and dP.
 * ---------------------------------------------------------------------- */
typedef struct mma_k1_args {
    uint32_t struct_size;           /* sizeof(mma_k1_args_t) of the CALLER */
    int32_t flags;                  /* MMA_K1_ARGS_LOCAL */
    /* graph: destination CSR and its optional acceleration tables */
    const int32_t *rowptr, *col, *perm, *edge_gid;
    int64_t E_total;
    const int32_t *row_map, *rng_row;
    int64_t rng_row0;
    const int32_t *row_chunks;
    int64_t n_chunks;
    const int32_t *vrowptr;
    int64_t n_vrows;
    const int32_t *seg_tab, *split_tab;
    int64_t n_split;
    float *seg_ws;
    int64_t n_rows, E;
    /* message operands */
    const float *P;
    int64_t ldp;
    const float *Q;
    int64_t ldq;
    const float *R;
    int64_t ldr;
    const float *keep;
    int64_t ldk;
    /* dropout */
    float p_drop;
    int32_t reserved0;
    uint64_t seed;
    const uint64_t *seed_dev;
    /* aggregators / scalers */
    int32_t T, F_in, A, S;
    const int32_t *aggr_kinds, *scaler_kinds;
    const float *scale_tab;
    int64_t tab_stride;
    /* aggregates (forward: out; backward: dY in) and what the forward saves for the backward */
    float *Y;
    int64_t ldy;
    int32_t *arg_min, *arg_max;
    float *stat_mean, *stat_var;
    /* backward destination pass only */
    const int32_t *gslot;
    float *G;
    int64_t ldg;
    float *dP;
    int64_t lddp;
    /* column window */
    int32_t col0, ncols;
} mma_k1_args_t;

int mmconv_aggregate_fwd_args(const mma_k1_args_t *args, mma_stream_t stream);
int mmconv_aggregate_bwd_dst_args(const mma_k1_args_t *args, mma_stream_t stream);

/* ------------------------------------------------------------------------
 * K3 / transpose pass: deterministic segmented row sum (CSR SpMM)
 *   out[i, c] = sum_{k in [ptr[i], ptr[i+1])} val[k] * src[idx[k], c]      (val NULL = 1)
 * Used as (a) K1 backward source pass: dQ[j] = sum of G rows of j's out-edges
 * (replaces index_select-backward / index_add_ atomics), and
 * (b) torch.spmm(adj, support) of layers.py:41 and layers.py:862, and its backward
 * on the transposed CSR. */
int mma_segment_sum_rows(const int32_t *ptr, const int32_t *idx, const float *val,
                         int64_t n_rows, const float *src, int64_t lds, int F,
                         float *out, int64_t ldo, mma_stream_t stream);

/* Row gather with fused column sums (the row permutation x[node_perm] of the degree-sorted layer
 * interior and the bias gradient sum_r dOut[r] in one pass):
 *   out[r, :] = src[idx[r], :]   r < n_out (idx NULL = identity), F % 4 == 0; out NULL = column sums only;
 *   colsum_part [n_parts, F] (optional): part p holds the column sums of rows [p*per, (p+1)*per),
 *   per = ceil(n_out / n_parts); reduce them in order with mma_reduce_slabs (no atomics). */
int mma_gather_rows(const float *src, int64_t lds, const int32_t *idx, int64_t n_out, int F,
                    float *out, int64_t ldo, float *colsum_part, int64_t n_parts, mma_stream_t stream);

/* ------------------------------------------------------------------------
 * K2: masked multi-aggregator layer of node_classification (layers.py:201-728),
 * all A aggregators in one pass over the CSR of neighbour lists (add_all):
 *   logit[a] = PA[i, a, c] + QA[j, a, c]        PA = X @ M_a[:F], QA = X @ M_a[F:]  (layers.py:215-216)
 *   mask     = act_a(logit) * keepscale_a[e, c]   (layers.py:217-219, dropout always on)
 *   S[a,i,c] = sum_j mask * X[j, c]               (layers.py:221)
 *   OUT[a,i,c] = combine_a(X[i,c], S[a,i,c], D_i) (sum/mean/max/min/none)
 *   PA, QA [N, A*F] ld ldpa/ldqa; X [N,F] ld ldx; keep [A, E, F] contiguous or NULL; OUT, S_out [A, N, F]
 *   edge id for dropout/keep = CSR position (neighbour-list order, as the reference consumes it).
 *   seed_dev (optional, device pointer): when set, the Philox key is read from device memory at run time
 *   instead of `seed`, so that a launch captured in a CUDA graph draws a fresh mask on every replay
 *   (same convention as mmconv_aggregate_fwd).
 * ---------------------------------------------------------------------- */
int mma_nc_aggregate_fwd(const int32_t *rowptr, const int32_t *col, int64_t N, int64_t E,
                         const float *X, int64_t ldx, const float *PA, int64_t ldpa,
                         const float *QA, int64_t ldqa, int F, int A,
                         const int32_t *act_kinds, const int32_t *comb_kinds,
                         const float *keep, float p_drop, uint64_t seed, const uint64_t *seed_dev,
                         float *OUT, float *S_out, mma_stream_t stream);

/* K2 backward, destination pass: from dOUT [A,N,F] computes
 *   gS [N, A*F] = dL/dS, node-major so the transpose pass gathers one contiguous row per edge
 *                 (max/min ties split 1/2 like torch.max/min backward),
 *   dXdir [N,F] = direct gradient through x_i in combine (contiguous, ld = F),
 *   dPA [N, A*F] = sum_j dL/dlogit. */
int mma_nc_aggregate_bwd_dst(const int32_t *rowptr, const int32_t *col, int64_t N, int64_t E,
                             const float *X, int64_t ldx, const float *PA, int64_t ldpa,
                             const float *QA, int64_t ldqa, int F, int A,
                             const int32_t *act_kinds, const int32_t *comb_kinds,
                             const float *keep, float p_drop, uint64_t seed, const uint64_t *seed_dev,
                             const float *S_saved, const float *dOUT,
                             float *gS, float *dXdir, float *dPA, int64_t lddpa,
                             mma_stream_t stream);

/* K2 backward, source (transpose) pass over the CSC (colptr/row/eid from
 * mma_csr_build keyed by the neighbour id; eid = CSR position of the edge):
 *   dQA[j, a, c] = sum_i dL/dlogit,   dXnbr[j, c] = sum_i sum_a gS[a,i,c] * mask. */
int mma_nc_aggregate_bwd_src(const int32_t *colptr, const int32_t *row, const int32_t *eid,
                             int64_t N, int64_t E,
                             const float *X, int64_t ldx, const float *PA, int64_t ldpa,
                             const float *QA, int64_t ldqa, int F, int A,
                             const int32_t *act_kinds,
                             const float *keep, float p_drop, uint64_t seed, const uint64_t *seed_dev,
                             const float *gS, float *dQA, int64_t lddqa, float *dXnbr, int64_t lddx,
                             mma_stream_t stream);

/* Materialises K2's in-kernel Philox dropout keep-scale (0 or 1/(1-p)), keyed by (edge id,
 * column, stream_id = aggregator slot a), as [E, F] floats, so that tests can inject the
 * identical mask into the CPU oracle. */
int mma_dropout_keep_scale(float p_drop, uint64_t seed, uint32_t stream_id,
                           int64_t E, int F, float *out, int64_t ldo, mma_stream_t stream);

/* The same for K1's row-keyed stream (see mmconv_aggregate_fwd): out[perm[k], c] for every CSR slot k
 * of the destination CSR (rowptr, perm; perm NULL = identity), i.e. [E, F] in ORIGINAL edge order. */
int mma_dropout_keep_scale_rows(const int32_t *rowptr, const int32_t *perm, const int32_t *rng_row,
                                int64_t rng_row0, int64_t n_rows, int64_t E, float p_drop, uint64_t seed,
                                int F, float *out, int64_t ldo, mma_stream_t stream);

/* ------------------------------------------------------------------------
 * G1-G4: the dense projections on the tcgen05 tensor cores, fp32-accurate by
 * 3xTF32 operand splitting (x = hi + lo; hi*hi + hi*lo + lo*hi, fp32 accumulate
 * in TMEM).  They replace the reference's fp32 Linears: the mask projection
 * (mask_aggr.py:68), the post-transform over cat([x, out]) and `lin`
 * (mma_conv.py:132-136), GraphConvolution's x @ W (layers.py:40) -- and their
 * autograd backward.  All operands fp32 row-major, 16-byte aligned, leading
 * dimensions multiples of 4.
 * ------------------------------------------------------------------------ */

/* hi = w with the low 13 mantissa bits cleared (exact TF32), lo = w - hi. */
int mma_tf32_split(const float *w, float *hi, float *lo, int64_t n, mma_stream_t stream);

/* C[out_row(r), 0:N] = [A0 | A1][r, :] . B[b_off + 0:N, :]^T (+ bias) (+ add[out_row(r), 0:N])
 *   A0 [M, K0], A1 [M, K1] (optional second source concatenated along K; K0 % 32 == 0 then);
 *   Bhi/Blo [b_rows, K0+K1] pre-split weight(s) (mma_tf32_split);
 *   tile_tab (optional) int32 [n_tiles_m][4] = {row0, row_end, b_off, 0}: 128-row tiles of a
 *     GROUPED GEMM -- rows [row0, row_end) use the weight rows b_off .. b_off+N (degree ranges
 *     with scaler-folded weights); null = plain GEMM over M rows, b_off = 0;
 *   out_map (optional) int32 [M]: output row of input row r (row scatter fused in the epilogue);
 *   bias [N], add (indexed like C, i.e. by OUTPUT row, may alias C; with MMA_GEMM_ADD_BY_INPUT_ROW
 *     or-ed into `mode`: indexed by the input row r) optional.
 *   mode 0 = 3xTF32 (hi written back), 1 = 3xTF32 (raw operand as hi), 2 = plain TF32 (not fp32-accurate).
 *   max_ctas <= 0: one persistent CTA per SM.
 *   Kernel choice: mode 1 with one source, K0 <= 128 and N >= 256 runs the variant that keeps the [128 x K]
 *   activation tile resident in tensor memory (one accumulator for the three terms; full tiles of a plain
 *   row-major output leave through TMA tile stores); everything else the streaming kernel.  Environment
 *   switches for A/B runs, read once per process: MMA_GEMM_ARES=0 (always stream), MMA_GEMM_TMA_STORE=0. */
#define MMA_GEMM_ADD_BY_INPUT_ROW 4
#define MMA_GEMM_RELU 8           /* or-ed into `mode`: C = max(C, 0) after bias / addend (BatchNorm-in-eval + ReLU
                                     folded into the layer's `lin`, graph_regression/mma.py:120-121) */

int mma_linear_tf32x3(const float *A0, int64_t lda0, int K0, const float *A1, int64_t lda1, int K1,
                      const float *Bhi, const float *Blo, int64_t ldb, int64_t b_rows,
                      int64_t M, int N, const int32_t *tile_tab, int64_t n_tiles_m,
                      float *C, int64_t ldc, const int32_t *out_map, const float *bias,
                      const float *add, int64_t ldadd, int mode, int max_ctas, mma_stream_t stream);

/* Weight gradient: part[slot][n][k] = sum_{m in slab} [G0 | G1][m, n] * A[m, k].
 *   slab_tab int32 [n_slabs][4] = {row0, row_end, slot, 0} (row0 % 32 == 0 is not required;
 *   slabs may end anywhere: rows >= row_end are masked); part [slots][N0+N1][K].
 *   Sum the slots in a fixed order with mma_reduce_slabs (no atomics). */
int mma_wgrad_tf32x3(const float *G0, int64_t ldg0, int N0, const float *G1, int64_t ldg1, int N1,
                     const float *A, int64_t lda, int K, int64_t M, const int32_t *slab_tab,
                     int64_t n_slabs, float *part, int mode, int max_ctas, mma_stream_t stream);

/* out[i] = sum_s coef[s] * part[s][i], s ascending (coef optional), i < n, n % 4 == 0. */
int mma_reduce_slabs(const float *part, const float *coef, int64_t n_slots, int64_t n, float *out,
                     mma_stream_t stream);

/* out[g][i] = sum of part[s][i] for s in [seg_ptr[g], seg_ptr[g+1]), ascending s; out [n_segs][n]. */
int mma_reduce_slabs_segmented(const float *part, const int32_t *seg_ptr, int64_t n_segs, int64_t n,
                               float *out, mma_stream_t stream);

/* ------------------------------------------------------------------------
 * Weight-space algebra of the fused layer (mma_b200/fused_layer.py): the post Linear (reference
 * graph_regression/mma_conv.py:132-133), `lin` (:136) and the cumulative scalers (:181-196) are composed into one
 * effective weight per in-degree range; these three calls replace ~60 small torch launches per step.
 *
 * mma_small_gemm: C[z] [M, N] (ldc; slab z at C + z * M * ldc) = op(A) * op(B) over the K-range of split z, plain fp32
 *   FFMA in ascending k.  A is [M, K] (lda), or [K, M] with trans_a; B is [K, N] (ldb), or [N, K] with trans_b.
 *   k_splits must equal mma_small_gemm_splits(K, wanted) (K-ranges are multiples of 16); add the slabs in order with
 *   mma_reduce_slabs.
 * mma_compose_post_weight: W_c[b][c][m*F+f] = sum_{s,a} coef[b][s][a][m] * WlW[c][col0 + (s*A+a)*F + f], written split
 *   for the 3xTF32 GEMMs (hi = tf32(x), lo = tf32(x - hi)): hi / lo [n_ranges][Co][Am*F] and, optionally, the
 *   transposes hiT / loT [n_ranges][Am*F][Co].  coef [n_ranges][S][A][Am], WlW [Co][>= col0 + S*A*F] (ldw).
 * mma_compose_post_wgrad: D[c][col0 + (s*A+a)*F + f] = sum_b coef[b][s][a][block_of[a]] * dWc[b][c][block_of[a]*F + f]
 *   (b ascending); D[c][0 .. col0) = dX[c][.] (dX may be NULL: zeros).  dWc [n_ranges][Co][Am*F], D [Co][ldd].
 * ---------------------------------------------------------------------- */
int mma_small_gemm(const float *A, int64_t lda, int trans_a, const float *B, int64_t ldb, int trans_b, float *C,
                   int64_t ldc, int M, int N, int K, int k_splits, mma_stream_t stream);
int mma_small_gemm_splits(int K, int wanted);
int mma_compose_post_weight(const float *coef, int n_ranges, int S, int A, int Am, const float *WlW, int64_t ldw,
                            int col0, int Co, int F, float *hi, float *lo, float *hiT, float *loT,
                            mma_stream_t stream);
int mma_compose_post_wgrad(const float *coef, int n_ranges, int S, int A, int Am, const int32_t *block_of,
                           const float *dWc, int Co, int F, const float *dX, int64_t lddx, int col0, float *D,
                           int64_t ldd, mma_stream_t stream);

/* ------------------------------------------------------------------------
 * BatchNorm + ReLU over the rows of x [n, F] in ONE kernel per direction: the step that follows the layer in
 * the reference's Net (graph_regression/mma.py:120-121, `F.relu(batch_norm(conv(...)))`).
 *   mean_in / var_in NULL : batch statistics (training) -- two-pass mean / centred second moment, biased variance
 *                           for the normalisation; running_mean / running_var (optional) are updated with
 *                           `momentum` and the unbiased variance, as torch.nn.BatchNorm1d does;
 *   mean_in / var_in given: those statistics (eval mode).
 *   save_mean / save_rstd [F] (optional out): what the backward needs.  gamma / beta NULL = 1 / 0.
 * Backward: dz = dy * (y > 0); dx (optional), dgamma, dbeta (optional) -- batch_stats != 0: the statistics
 * depended on x.  Fixed-order reductions, no atomics.  One CTA per 32 columns: sized for batches of small graphs.
 * ---------------------------------------------------------------------- */
int mma_bn_relu_fwd(const float *x, int64_t ldx, int64_t n, int F, const float *gamma, const float *beta, float eps,
                    const float *mean_in, const float *var_in, float *y, int64_t ldy, float *save_mean,
                    float *save_rstd, float *running_mean, float *running_var, float momentum, mma_stream_t stream);
int mma_bn_relu_bwd(const float *x, int64_t ldx, const float *y, int64_t ldy, const float *dy, int64_t lddy,
                    int64_t n, int F, const float *gamma, const float *save_mean, const float *save_rstd,
                    int batch_stats, float *dx, int64_t lddx, float *dgamma, float *dbeta, mma_stream_t stream);

/* ------------------------------------------------------------------------
 * C1: the two exchanges of the destination-range sharded layer (new -- the reference
 * is single-device) over NVLink peer memory.  The payload (forward: every rank's Q rows
 * into every peer's gathered buffer; backward: every rank's partial dQ slice into its owner's slice
 * buffer) is moved by the COPY ENGINES through peer-to-peer copies the host enqueues; these entry
 * points are the device-side handshake around them and the owner's fixed-order reduction.
 *   epoch        device uint64, advanced once per layer call (a step captured in a CUDA graph then hands
 *                out a fresh value at every replay);
 *   phase        0..15: which exchange of the call (forward window k, backward window k, buffers free);
 *   a flag holds epoch * 16 + phase of the last announcement; flags only grow.
 * ---------------------------------------------------------------------- */
/* Host.  The exchange buffers are the one thing the library allocates itself (a CUDA IPC handle names a
 * whole cudaMalloc allocation): mma_peer_alloc = cudaMalloc + zero fill + cudaIpcGetMemHandle (handle64:
 * 64 bytes out, to be sent to the peer processes); mma_peer_open maps a peer process's buffer into the
 * CURRENT device's address space (cudaIpcOpenMemHandle with lazy peer access -- the importer's kernels
 * and copy engines then reach it over NVLink); mma_peer_close / mma_peer_free undo them. */
int mma_peer_alloc(size_t bytes, void **ptr, void *handle64);
int mma_peer_free(void *ptr);
int mma_peer_open(const void *handle64, void **ptr);
int mma_peer_close(void *ptr);
/* Host: lets the CURRENT device address `peer_device`'s memory directly (cudaDeviceEnablePeerAccess;
 * already-enabled is fine).  MMA_ERR_UNSUPPORTED when the two devices have no peer path. */
int mma_peer_enable_access(int peer_device);
/* epoch += 1; vals[p] = epoch * 16 + p for p < 16 (vals: device uint64 [16], optional). */
int mma_peer_epoch_advance(uint64_t *epoch, uint64_t *vals, mma_stream_t stream);
/* cudaMemcpyAsync / cudaMemcpy2DAsync between (peer-mapped) device pointers on `stream`: the copy engines
 * move the payload, and an 8-byte copy of vals[phase] into the peer's flag, enqueued right after the
 * payload on the same stream, announces it -- neither needs an SM, so both run under persistent kernels. */
int mma_peer_copy(void *dst, const void *src, size_t bytes, mma_stream_t stream);
int mma_peer_copy_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes,
                     size_t height, mma_stream_t stream);
/* Returns (on the stream) once flags[t] >= epoch * 16 + phase for all t < n_flags <= 32 (acquire, system
 * scope).  After timeout_ns without progress *err = 1 + t (device int32, optional) and the wait gives up:
 * a lost peer must never hang the GPU. */
int mma_peer_wait(const uint64_t *epoch, const uint64_t *flags, int n_flags, int phase,
                  uint64_t timeout_ns, int32_t *err, mma_stream_t stream);
/* out[r, 0:w] = sum_k slices[k][r, 0:w], k ascending (deterministic, no atomics); slices: HOST array of
 * n_slices <= 16 device pointers to contiguous [rows, w] fp32 blocks, w % 4 == 0. */
int mma_sum_slices(const float *const *slices, int n_slices, int64_t rows, int w, float *out, int64_t ldo,
                   mma_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MMA_B200_H */
