// K1 "stream" kernels: the HBM-bound configuration of the MultiMaskConv aggregate (128-bit column
// groups, message P[dst] + Q[src], in-kernel dropout) as a persistent, latency-tolerant pipeline.
//
// The register-batch kernels in mmconv_aggregate.cu give every destination row to a fresh warp, which
// walks a chain of dependent DRAM round trips (rowptr -> col -> Q rows -> next batch) with at most 8
// gathered rows in flight, and spends ~1000 instructions per row on generic addressing: ncu showed them
// latency / issue bound at ~50 % of the HBM roofline.  Here
//   * the rows are cut into cost-balanced CHUNKS (row_chunks, built once per graph; ~50 rows each) that
//     are dealt round-robin to persistent warps; a chunk's in-edges are one contiguous stream of CSR
//     slots, so col[] is read with coalesced, prefetched loads and the row boundaries come from rowptr
//     one row ahead -- no data-dependent index chain;
//   * the gathered rows Q[src] are copied global -> shared with cp.async (LDGSTS, 16 B per lane -- 8 / 4 B for
//     column windows of <= 64 / 32 columns, where a lane owns 2 / 1 columns: a lane copies exactly the
//     columns it later consumes, so completion needs no cross-lane barrier) into a
//     per-warp ring of NST stages x 4 edges that runs NST-1 stages AHEAD of consumption and does not
//     stop at row boundaries;
//   * a stage that lies inside one row is consumed by a 4-edge unrolled, branch-free body (12
//     instructions per element); stages that straddle a row boundary go edge by edge; the row epilogue
//     keeps running output pointers instead of re-deriving addresses.
// Same arithmetic, same summation order, same dropout stream as the other K1 kernels (bit-identical
// results); included by mmconv_aggregate.cu.
#pragma once

namespace stream {

constexpr int kRowCost = 6;        // a row costs about this many edges (used when no row_chunks are given)

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// BYTES = 16: L2-only (.cg); 8 / 4 (narrow column windows: 2 / 1 columns per lane): .ca is the only form
template <int BYTES>
__device__ __forceinline__ void cp_async_pred(uint32_t dst, const void *src, bool pred) {
    if constexpr (BYTES == 16)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                     :: "r"(dst), "l"(src), "r"((int)pred) : "memory");
    else if constexpr (BYTES == 8)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.ca.shared.global [%0], [%1], 8;\n\t}"
                     :: "r"(dst), "l"(src), "r"((int)pred) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.ca.shared.global [%0], [%1], 4;\n\t}"
                     :: "r"(dst), "l"(src), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// copy iff cnt > U (U a compile-time slot number): the predicate is formed inside, one SETP per copy
template <int BYTES, int U>
__device__ __forceinline__ void cp_async_if_gt(uint32_t dst, const void *src, int cnt) {
    if constexpr (BYTES == 16)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.gt.s32 p, %2, %3;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                     :: "r"(dst), "l"(src), "r"(cnt), "n"(U) : "memory");
    else if constexpr (BYTES == 8)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.gt.s32 p, %2, %3;\n\t@p cp.async.ca.shared.global [%0], [%1], 8;\n\t}"
                     :: "r"(dst), "l"(src), "r"(cnt), "n"(U) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.gt.s32 p, %2, %3;\n\t@p cp.async.ca.shared.global [%0], [%1], 4;\n\t}"
                     :: "r"(dst), "l"(src), "r"(cnt), "n"(U) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// dst = *ptr if pred, else unchanged -- a predicated load IN PLACE.  The merge with the old value pins the destination
// to the register that holds it, so ptxas can neither turn `if (c) x = load` into a load into a temporary + select
// (which consumes the loaded value on the spot) nor hoist the load above the last read of the old value.
__device__ __forceinline__ void ldg_i32_if(int &dst, const int32_t *ptr, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p ld.global.nc.b32 %0, [%1];\n\t}"
                 : "+r"(dst) : "l"(ptr), "r"((int)pred) : "memory");
}

template <int VEC>
__device__ __forceinline__ Vec<VEC> lds_vec(uint32_t addr) {
    Vec<VEC> r;
    if constexpr (VEC == 4)
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "r"(addr));
    else if constexpr (VEC == 2)
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(r.v[0]), "=f"(r.v[1]) : "r"(addr));
    else
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r.v[0]) : "r"(addr));
    return r;
}

// smallest r in [0, n_rows] with rowptr[r] + kRowCost * r >= target
__device__ __forceinline__ int first_row_at(const int32_t *rowptr, int64_t n_rows, int64_t target) {
    int64_t lo = 0, hi = n_rows;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(rowptr + mid) + kRowCost * mid >= target) hi = mid; else lo = mid + 1;
    }
    return (int)lo;
}

// IEEE-correct a / y for the row's degree y (integer valued, 1 <= y <= 2^24) with the reciprocal shared by
// all columns: r = 1/y refined once; q = a*r corrected by one residual step is the correctly rounded
// quotient whenever nothing under/overflows, which the exponent test guarantees; otherwise __fdiv_rn.
struct DivByDeg {
    float y, r;
    __device__ __forceinline__ explicit DivByDeg(float y_) : y(y_) {
        float r0;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(y_));
        const float e = __fmaf_rn(-y_, r0, 1.0f);
        r = __fmaf_rn(r0, e, r0);
    }
    __device__ __forceinline__ float operator()(float a) const {
        const uint32_t ex = (__float_as_uint(a) >> 23) & 0xFFu;
        if (ex - 40u < 176u) {                       // 2^-87 <= |a| < 2^89: no intermediate can leave the normal range
            const float q0 = __fmul_rn(a, r);
            const float rem = __fmaf_rn(-y, q0, a);
            return __fmaf_rn(r, rem, q0);
        }
        return __fdiv_rn(a, y);
    }
};

// (virtual) rows [r0, r1) of chunk i; rp / n_rows: the row pointer the stream walks
__device__ __forceinline__ void chunk_rows(const MMConvParams &p, const int32_t *rp, int64_t n_rows, int64_t i,
                                           int64_t n_chunks, int &r0, int &r1) {
    if (p.row_chunks) {
        r0 = __ldg(p.row_chunks + i); r1 = __ldg(p.row_chunks + i + 1);
    } else {
        const int64_t total = p.E + kRowCost * n_rows;
        r0 = first_row_at(rp, n_rows, (total * i) / n_chunks);
        r1 = (i + 1 == n_chunks) ? (int)n_rows : first_row_at(rp, n_rows, (total * (i + 1)) / n_chunks);
    }
}

// Long rows.  A row of a skewed graph can hold a large share of all edges; walked by one warp it would set
// the kernel's duration.  The caller may therefore pass a VIRTUAL row pointer in which every row longer than
// a segment length (a multiple of 32) is cut into segments: seg_tab[v] = {real row, in-row position of the
// segment's first edge, partial slot or -1 for an unsplit row, 0}.  A segment is walked like a row (same
// dropout positions, same edge order) but ends by writing its partial state to seg_ws[slot]; a small second
// kernel merges the partials of each split row IN SEGMENT ORDER (strict comparisons, so the first occurrence
// still wins and min/max stay bit-exact; sums differ from the sequential order only by rounding) and emits
// the row.
struct RowDesc { int real, pos0, slot; };
__device__ __forceinline__ RowDesc row_desc(const MMConvParams &p, int v) {
    RowDesc d{v, 0, -1};
    if (p.seg_tab) {
        const int4 t = __ldg(reinterpret_cast<const int4 *>(p.seg_tab) + v);
        d.real = t.x; d.pos0 = t.y; d.slot = t.z;
    }
    return d;
}

// aggregates of one finished row -> Y / arg indices / saved statistics (this lane's VEC columns)
template <int VEC, bool MINMAX, bool SQ>
__device__ __forceinline__ void emit_row(const MMConvParams &p, bool simple_out, int deg, const float (&sum)[VEC],
                                         const float (&sq)[VEC], const float (&mn)[VEC], const float (&mx)[VEC],
                                         const int (&amn)[VEC], const int (&amx)[VEC], float *yp, int32_t *amn_p,
                                         int32_t *amx_p, float *mean_p, float *var_p) {
    const int degc = deg > 1 ? deg : 1;                     // deg.clamp_(1), mma_conv.py:179
    const DivByDeg div((float)degc);
    Vec<VEC> mean, var, sd, vmin, vmax, vsum;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        vsum.v[v] = sum[v];
        mean.v[v] = div(sum[v]);                            // sum / count.clamp(min=1)
        if constexpr (SQ) {
            const float msq = div(sq[v]);
            var.v[v] = __fsub_rn(msq, __fmul_rn(mean.v[v], mean.v[v]));         // mma_conv.py:170, no FMA
            sd.v[v] = sqrtf(__fadd_rn(fmaxf(var.v[v], 0.0f), 1e-5f));           // mma_conv.py:172
        } else {
            var.v[v] = 0.0f; sd.v[v] = 0.0f;
        }
        vmin.v[v] = (MINMAX && amn[v] >= 0) ? mn[v] : 0.0f;                      // empty row -> 0
        vmax.v[v] = (MINMAX && amx[v] >= 0) ? mx[v] : 0.0f;
    }
    if (simple_out) {
        if (p.zoff[MMA_AGGR_SUM] >= 0) st_vec_stream<VEC>(yp + p.zoff[MMA_AGGR_SUM], vsum);
        if (p.zoff[MMA_AGGR_MEAN] >= 0) st_vec_stream<VEC>(yp + p.zoff[MMA_AGGR_MEAN], mean);
        if constexpr (MINMAX) {
            if (p.zoff[MMA_AGGR_MIN] >= 0) st_vec_stream<VEC>(yp + p.zoff[MMA_AGGR_MIN], vmin);
            if (p.zoff[MMA_AGGR_MAX] >= 0) st_vec_stream<VEC>(yp + p.zoff[MMA_AGGR_MAX], vmax);
        }
        if constexpr (SQ) {
            if (p.zoff[MMA_AGGR_VAR] >= 0) st_vec_stream<VEC>(yp + p.zoff[MMA_AGGR_VAR], var);
            if (p.zoff[MMA_AGGR_STD] >= 0) st_vec_stream<VEC>(yp + p.zoff[MMA_AGGR_STD], sd);
        }
    } else {
        float fac[MMA_MAX_SCALER];
        scaler_factors(p, degc, fac);
        for (int a = 0; a < p.A; ++a) {
            const int kind = p.akind[a];
            Vec<VEC> val = kind == MMA_AGGR_SUM ? vsum : kind == MMA_AGGR_MEAN ? mean : kind == MMA_AGGR_MIN ? vmin
                       : kind == MMA_AGGR_MAX ? vmax : kind == MMA_AGGR_VAR ? var : sd;
            for (int s = 0; s < p.S; ++s) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) val.v[v] = __fmul_rn(val.v[v], fac[s]);       // cumulative (Q4)
                st_vec_stream<VEC>(yp + (int64_t)(s * p.A + a) * p.F_in, val);
            }
        }
    }
    if constexpr (MINMAX) {
        int32_t o_mn[VEC], o_mx[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            o_mn[v] = amn[v] < 0 ? (int32_t)p.E_total : (p.args_local ? amn[v] : orig_edge_id(p, amn[v]));
            o_mx[v] = amx[v] < 0 ? (int32_t)p.E_total : (p.args_local ? amx[v] : orig_edge_id(p, amx[v]));
        }
        if (amn_p) st_vec_i32_stream<VEC>(amn_p, o_mn);
        if (amx_p) st_vec_i32_stream<VEC>(amx_p, o_mx);
    }
    if (mean_p) st_vec_stream<VEC>(mean_p, mean);
    if constexpr (SQ) {
        if (var_p) st_vec_stream<VEC>(var_p, var);
    }
}

// Issue side of the pipeline: index chunks + the cp.async ring of one warp.
template <int NST, bool NQ, int VEC>
struct Ring {
    uint32_t base;             // shared-memory address of this lane's 16 bytes in ring slot 0
    int rowb;                  // bytes of one gathered row window
    const char *Qc;            // Q + this lane's column
    uint32_t ldq_b;
    const int32_t *colp;       // col + first CSR slot of the stream
    int len, lane;
    bool live;
    int idx_cur, idx_nxt;      // col[] of the 32-edge chunk being issued / the one after it
    int p_issue;               // stream position of the next stage to issue
    int ring_i;                // its ring slot

    __device__ __forceinline__ int load_chunk(int pbase) const {
        const int i = pbase + lane;
        return (NQ && i < len) ? __ldg(colp + i) : 0;
    }
    // issues one stage (4 stream edges) and commits it as one cp.async group (an empty group past
    // the end of the stream keeps the wait_group arithmetic uniform)
    __device__ __forceinline__ void issue() {
        if constexpr (NQ) {
            if (p_issue < len) {
                if ((p_issue & 31) == 0) { idx_cur = idx_nxt; idx_nxt = load_chunk(p_issue + 32); }
                const uint32_t dst = base + (uint32_t)(ring_i * 4 * rowb);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = __shfl_sync(0xffffffffu, idx_cur, (p_issue & 31) + u);
                    cp_async_pred<VEC * 4>(dst + (uint32_t)(u * rowb), Qc + (uint64_t)(uint32_t)j * ldq_b, live && (p_issue + u < len));
                }
            }
        }
        cp_async_commit();
        p_issue += 4;
        ring_i = ring_i + 1 == NST ? 0 : ring_i + 1;
    }
    __device__ __forceinline__ void begin(const int32_t *col_at_s0, int len_) {
        colp = col_at_s0; len = len_;
        p_issue = 0; ring_i = 0;
        idx_cur = 0; idx_nxt = load_chunk(0);
#pragma unroll 1
        for (int k = 0; k < NST; ++k) issue();
        cp_async_wait<NST - 1>();                   // stage 0 has landed
    }
    // the consumer is done with its current stage: refill that slot with the stage NST ahead, then make
    // sure the next stage has landed
    __device__ __forceinline__ void advance() {
        issue();
        cp_async_wait<NST - 1>();
    }
    __device__ __forceinline__ void end() { cp_async_wait<0>(); }
};

// Forward kernel: the same ring with ROW-ALIGNED stages.  A stage holds the next (up to) 4 edges OF ONE ROW: the
// last stage of a row is short (its unused slots are predicated off), so every full stage lies inside one row and is
// consumed by the unrolled 4-edge body, and an edge's in-row position is a multiple of 4 at every stage start -- a
// stage never straddles a 32-position dropout word either.  Only the deg mod 4 edges of a row's last stage go edge by
// edge (with position-aligned stages three row boundaries in four fell INSIDE a stage at Poisson(16) in-degrees and
// sent a quarter of all edges down the single-edge path).  The issue side walks the row boundaries itself, NST - 1
// stages ahead of the consumer; both derive the same stage sequence from the row pointer.
template <int NST, int VEC>
struct RowRing {
    uint32_t base;             // shared-memory address of this lane's bytes in ring slot 0
    int rowb;                  // bytes of one gathered row window
    const char *Qc;            // Q + this lane's column
    uint32_t ldq_b;
    int lane;
    bool live;
    const int32_t *colp;       // col + first CSR slot of the stream
    const int32_t *rp;         // (virtual) row pointer
    int s0, len;
    int64_t n_rp;              // rows of the (virtual) row pointer: rp[0 .. n_rp] are readable
    int i_row, i_pos, i_end;   // issue side: row being issued, stream position of its next edge, end of that row
    int i_end_nxt;             // rp[i_row + 2], fetched one row ahead and left RAW until it is consumed (an add on a
                               // freshly loaded value would stall the issue side for a full memory round trip per row)
    // col[] of the stream travels through a 128-entry shared-memory window of this warp (slot = position & 127), filled
    // 32 positions at a time with 4-byte cp.async copies that ride in the ring's own commit groups: the chunk
    // [idx_base + 64, +32) is requested when the issue position enters [idx_base, +32), i.e. >= 15 stages before its first
    // entry is read, and wait_group<NST - 1> after every stage has retired it long before (NST <= 8).  Registers would
    // not do: `cur = nxt; nxt = load` makes ptxas load into a temporary and copy (or select) it into the loop-carried
    // register in the same basic block -- a full memory round trip per chunk spent waiting (ncu: 6-7 % of all stall
    // samples, whichever way the source was arranged).
    uint32_t idx_smem;         // shared-memory address of the window
    int idx_base;
    int ring_i;

    __device__ __forceinline__ void request_chunk(int pbase) {          // positions [pbase, pbase + 32) -> their slots
        const int i = pbase + lane;
        cp_async_pred<4>(idx_smem + (uint32_t)((i & 127) * 4), colp + i, i < len);
    }
    __device__ __forceinline__ int col_at(int pos) const {             // broadcast read
        int j;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(j) : "r"(idx_smem + (uint32_t)((pos & 127) * 4)));
        return j;
    }
    __device__ __forceinline__ void issue() {
        __syncwarp();                                       // the other lanes' retired index copies become visible
        while (i_pos == i_end && i_pos < len) {                                                    // skip finished / empty rows
            ++i_row;
            i_end = i_end_nxt - s0;
            const bool in_range = (int64_t)i_row + 2 <= n_rp;
            if (!in_range) i_end_nxt = 0x7fffffff;
            ldg_i32_if(i_end_nxt, rp + i_row + 2, in_range);
        }
        int n = i_end - i_pos;
        n = n < 4 ? n : 4;
        if (i_pos >= len) n = 0;
        if (n > 0) {
            if (i_pos >= idx_base + 32) { idx_base += 32; request_chunk(idx_base + 64); }
            const uint32_t dst = base + (uint32_t)(ring_i * 4 * rowb);
            const int cnt = live ? n : 0;                   // slots u >= cnt: a stale index, the copy is predicated off
            const int sl = i_pos & 127;
            int j0, j1, j2, j3;
            if (sl <= 124) {                                // the four entries are consecutive slots of the window
                const uint32_t a = idx_smem + (uint32_t)(sl * 4);
                asm volatile("ld.shared.b32 %0, [%4];\n\tld.shared.b32 %1, [%4+4];\n\tld.shared.b32 %2, [%4+8];\n\t"
                             "ld.shared.b32 %3, [%4+12];" : "=r"(j0), "=r"(j1), "=r"(j2), "=r"(j3) : "r"(a));
            } else {
                j0 = col_at(i_pos); j1 = col_at(i_pos + 1); j2 = col_at(i_pos + 2); j3 = col_at(i_pos + 3);
            }
            cp_async_if_gt<VEC * 4, 0>(dst, Qc + (uint64_t)(uint32_t)j0 * ldq_b, cnt);
            cp_async_if_gt<VEC * 4, 1>(dst + (uint32_t)rowb, Qc + (uint64_t)(uint32_t)j1 * ldq_b, cnt);
            cp_async_if_gt<VEC * 4, 2>(dst + (uint32_t)(2 * rowb), Qc + (uint64_t)(uint32_t)j2 * ldq_b, cnt);
            cp_async_if_gt<VEC * 4, 3>(dst + (uint32_t)(3 * rowb), Qc + (uint64_t)(uint32_t)j3 * ldq_b, cnt);
        }
        cp_async_commit();
        i_pos += n;
        ring_i = ring_i + 1 == NST ? 0 : ring_i + 1;
    }
    __device__ __forceinline__ void begin(const int32_t *col_at_s0, const int32_t *rp_, int64_t n_rp_, int r0, int s0_, int len_) {
        colp = col_at_s0; rp = rp_; n_rp = n_rp_; s0 = s0_; len = len_;
        i_row = r0; i_pos = 0; i_end = __ldg(rp + r0 + 1) - s0;
        i_end_nxt = (int64_t)r0 + 2 <= n_rp ? __ldg(rp + r0 + 2) : 0x7fffffff;
        ring_i = 0;
        idx_base = 0;
        __syncwarp();                                       // every lane is done reading the previous stream's window
        request_chunk(0); request_chunk(32); request_chunk(64);
        cp_async_commit();
        cp_async_wait<0>();
#pragma unroll 1
        for (int k = 0; k < NST; ++k) issue();
        cp_async_wait<NST - 1>();                   // stage 0 has landed
    }
    __device__ __forceinline__ void advance() {
        issue();
        cp_async_wait<NST - 1>();
    }
    __device__ __forceinline__ void end() { cp_async_wait<0>(); }
};

// One int per stream edge (G-row slot, original edge id), read by the CONSUMER in coalesced chunks of
// 32 edges, one chunk ahead.  load(pbase) returns the value of stream edge pbase + lane (0 past the end).
struct EdgeAttr {
    int cur, nxt;
    template <typename Load>
    __device__ __forceinline__ void init(Load &&load) { cur = load(0); nxt = load(32); }
    template <typename Load>
    __device__ __forceinline__ void at_stage(int s, Load &&load) {      // s: first edge of the stage (multiple of 4)
        if ((s & 31) == 0 && s > 0) { cur = nxt; nxt = load(s + 32); }
    }
    __device__ __forceinline__ int get(int s) const { return __shfl_sync(0xffffffffu, cur, s & 31); }
};

// Drives one chunk: walks the stages of the stream; calls
//   row_next()            when the stream passes the end of the current row (finish it, start the next),
//   edge4(at, q[4], s)    for a stage of 4 edges inside the current row (at = CSR slot of the first),
//   edge1(at, q, s)       for single edges,
//   fast_ok()             whether the 4-edge body may be used now (dropout word boundaries),
//   stage_begin(s)        at the first edge of every stage.
// row_end is read through a reference: row_next() updates it.
template <int NST, bool NQ, int VEC, typename RowNext, typename Edge4, typename Edge1, typename FastOk, typename StageBegin>
__device__ __forceinline__ void run_stream(Ring<NST, NQ, VEC> &ring, int s0, int len, const int &row_end, RowNext &&row_next,
                                           Edge4 &&edge4, Edge1 &&edge1, FastOk &&fast_ok, StageBegin &&stage_begin) {
    int cons_i = 0;
#pragma unroll 1
    for (int s = 0; s < len; s += 4) {
        const int n = len - s < 4 ? len - s : 4;
        const uint32_t sb = ring.base + (uint32_t)(cons_i * 4 * ring.rowb);
        stage_begin(s);
        int u = 0;
#pragma unroll 1
        while (u < n) {
            const int at = s0 + s + u;
            if (at == row_end) { row_next(); continue; }
            if (u == 0 && n == 4 && row_end - at >= 4 && fast_ok()) {
                Vec<VEC> q[4];
                if constexpr (NQ) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) q[i] = lds_vec<VEC>(sb + (uint32_t)(i * ring.rowb));
                }
                edge4(at, q, s);
                u = 4;
            } else {
                Vec<VEC> q{};
                if constexpr (NQ) q = lds_vec<VEC>(sb + (uint32_t)(u * ring.rowb));
                edge1(at, q, s + u);
                ++u;
            }
        }
        ring.advance();
        cons_i = cons_i + 1 == NST ? 0 : cons_i + 1;
    }
    ring.end();
}

// the words of a Philox call (4 columns of one column group) that belong to this lane's VEC columns
template <int VEC>
__device__ __forceinline__ void pick_words(const uint4 &t, int c, uint32_t (&bits)[VEC]) {
    if constexpr (VEC == 4) {
        bits[0] = t.x; bits[1] = t.y; bits[2] = t.z; bits[3] = t.w;
    } else if constexpr (VEC == 2) {
        bits[0] = (c & 2) ? t.z : t.x; bits[1] = (c & 2) ? t.w : t.y;
    } else {
        const int k = c & 3;
        bits[0] = k == 0 ? t.x : (k == 1 ? t.y : (k == 2 ? t.z : t.w));
    }
}
// The dropout key of this launch, resolved ONCE: with a device-resident seed (CUDA-graph replays) dropout_key() is a
// global load, and repeated inside every Philox call it shared a scoreboard with the row prefetches issued just before
// it -- the first Philox of every row then waited for THEIR memory round trip (ncu: 9 % of all stall samples).
__device__ __forceinline__ Dropout resolve_key(const Dropout &d) {
    Dropout r = d;
    const uint2 k = dropout_key(d);
    r.k0 = k.x; r.k1 = k.y; r.seed_dev = nullptr;
    return r;
}

// dropout words for in-row position pos: refreshes `bits` at the word boundaries of the row's stream
template <int DROP, int VEC>
__device__ __forceinline__ void rng_refresh(const Dropout &drop, uint32_t rid, int pos, int c, uint32_t (&bits)[VEC]) {
    if constexpr (DROP == FD_BIT) {
        if ((pos & 31) == 0) pick_words<VEC>(row_rng_bits1(drop, rid, (uint32_t)pos, (uint32_t)c), c, bits);
    } else if constexpr (DROP == FD_BYTE) {
        if ((pos & 3) == 0) pick_words<VEC>(row_rng_bits8(drop, rid, (uint32_t)pos, (uint32_t)c), c, bits);
    }
}
// non-zero iff column v of the edge at in-row position pos is kept
template <int DROP>
__device__ __forceinline__ uint32_t keep_at(uint32_t word, int pos, uint32_t thr) {
    if constexpr (DROP == FD_BIT) return word & (1u << (pos & 31));
    else if constexpr (DROP == FD_BYTE) return ((word >> (8 * (pos & 3))) & 0xFFu) >= thr ? 1u : 0u;
    else return 1u;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int NST, int WARPS, int VEC, int DROP, bool MINMAX, bool SQ>
__global__ void __launch_bounds__(WARPS * 32, 1) mmconv_fwd_stream(const __grid_constant__ MMConvParams p) {
    extern __shared__ __align__(16) uint8_t smem_ring[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t gw = (int64_t)blockIdx.x * WARPS + warp, tw = (int64_t)gridDim.x * WARPS;
    const int64_t n_chunks = p.row_chunks ? p.n_chunks : tw;
    const bool live = lane * VEC < p.ncols;
    const int c = p.col0 + lane * VEC;
    const Dropout drop = resolve_key(p.drop);
    const float scale = p.drop.scale;
    const uint32_t thr = p.drop.thr;
    // outputs of the common case (S == 1, every aggregator kind at most once): column offset of each kind in a
    // row of Y, or -1 (fill_params); anything else goes through the generic loop
    const bool simple_out = p.simple_out != 0;

    RowRing<NST, VEC> ring;
    ring.rowb = p.ncols * 4;
    ring.base = smem_addr(smem_ring) + (uint32_t)warp * (uint32_t)(NST * 4 * ring.rowb) + (live ? (uint32_t)(lane * VEC * 4) : 0u);   // idle lanes stay inside the ring
    ring.Qc = reinterpret_cast<const char *>(p.Q + c);
    ring.ldq_b = (uint32_t)p.ldq * 4u;
    ring.lane = lane; ring.live = live;
    ring.idx_smem = smem_addr(smem_ring) + (uint32_t)WARPS * (uint32_t)(NST * 4 * ring.rowb) + (uint32_t)warp * 512u;   // after the rings
    const int32_t *rp = p.vrowptr ? p.vrowptr : p.rowptr;      // the (virtual) row boundaries walked by the stream
    const int64_t n_rows = p.vrowptr ? p.n_vrows : p.n_rows;
    const int ycol = p.T == 1 ? c : (c / p.F_in) * (p.S * p.A * p.F_in) + c % p.F_in;

#pragma unroll 1
    for (int64_t chunk = gw; chunk < n_chunks; chunk += tw) {
        int r0, r1;
        chunk_rows(p, rp, n_rows, chunk, n_chunks, r0, r1);
        if (r0 >= r1) continue;
        const int s0 = __ldg(rp + r0);
        const int len = __ldg(rp + r1) - s0;

        // ---- per-row state; the next row's inputs are fetched one row ahead ----
        int row = r0;
        int row_beg = s0, row_end = __ldg(rp + r0 + 1);
        int next_end = r0 + 2 <= n_rows ? __ldg(rp + r0 + 2) : 0x7fffffff;
        // Per-row inputs arrive through a two-level software pipeline so that no load address and no arithmetic ever
        // waits on a load issued in the same step: the INDICES of row r + 2 (segment descriptor, P row = row_map[.],
        // dropout row id = rng_row[.]) are fetched while row r is walked and kept raw; the P row of r + 1 is fetched
        // with the index that arrived one row earlier; the doubling of P (1-bit dropout mode) and the rng id's offset
        // are applied when the row becomes current.  (ncu, round 2: with a single level the row_map -> P and
        // rng_row -> id chains stalled every warp for a memory round trip per row -- 13 % of all stall samples.)
        // The loads are issued from the FIRST STAGE of a row (`pending`), not from row_next(): in one basic block with
        // the moves that consume the previous values ptxas hoists the loads to the top, lands them in temporaries and
        // copies the temporaries into the loop-carried registers at the end of the block -- the stall all over again.
        struct RowIn { RowDesc d; int map_raw, rng_raw; };
        Vec<VEC> pv{}, pv_next{};
        uint32_t rid = 0;
        RowDesc rd{};
        RowIn in1{}, in2{};                                 // indices of row + 1 / row + 2
        bool pending = false;                               // the loads for the rows after `row` are still to be issued
        float sum[VEC], sq[VEC], mn[VEC], mx[VEC];
        int amn[VEC], amx[VEC];
        uint32_t bits[VEC] = {};
        int pos = 0;                                        // in-row position of the next edge
        float *yp = nullptr;                                // running output pointers (this lane's columns of the row)
        int32_t *amn_p = nullptr, *amx_p = nullptr;
        float *mean_p = nullptr, *var_p = nullptr;

        auto fetch_idx = [&](int r, RowIn &x) {
            x.d = RowDesc{r, 0, -1}; x.map_raw = 0; x.rng_raw = 0;
            if (r < r1) {
                x.d = row_desc(p, r);
                x.map_raw = p.row_map ? __ldg(p.row_map + x.d.real) : x.d.real;
                x.rng_raw = (p.use_rng && p.rng_row) ? __ldg(p.rng_row + x.d.real) : x.d.real;
            }
        };
        auto fetch_P = [&](int r, const RowIn &x, Vec<VEC> &pvv) {             // raw P row (this lane's columns)
            pvv = Vec<VEC>{};
            if (r < r1 && live) pvv = ld_vec_stream<VEC>(p.P + (int64_t)x.map_raw * p.ldp + c);
        };
        auto make_current = [&](const RowIn &x, const Vec<VEC> &raw) {        // the fetched inputs become the current row's
            rd = x.d;
            rid = p.use_rng ? (uint32_t)(p.rng_row0 + (int64_t)x.rng_raw) : 0u;
#pragma unroll
            for (int v = 0; v < VEC; ++v) pv.v[v] = DROP == FD_BIT ? raw.v[v] * 2.0f : raw.v[v];
        };
        auto reset_acc = [&]() {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                sum[v] = 0.0f; sq[v] = 0.0f; mn[v] = FLT_MAX; mx[v] = -FLT_MAX; amn[v] = -1; amx[v] = -1;
            }
            pos = rd.pos0;
        };
        auto point_at = [&](int real) {
            yp = p.Y + (int64_t)real * p.ldy + ycol;
            const int64_t rowF = (int64_t)real * p.F + c;
            amn_p = p.arg_min ? p.arg_min + rowF : nullptr;
            amx_p = p.arg_max ? p.arg_max + rowF : nullptr;
            mean_p = p.stat_mean ? p.stat_mean + rowF : nullptr;
            var_p = p.stat_var ? p.stat_var + rowF : nullptr;
        };
        auto finish_row = [&]() {
            if (!live) return;
            if (rd.slot >= 0) {                             // a segment of a split row: partial state -> workspace
                float *w = p.seg_ws + (int64_t)rd.slot * 6 * p.F + c;
                Vec<VEC> t;
#pragma unroll
                for (int v = 0; v < VEC; ++v) t.v[v] = sum[v];
                st_vec<VEC>(w, t);
#pragma unroll
                for (int v = 0; v < VEC; ++v) t.v[v] = sq[v];
                st_vec<VEC>(w + p.F, t);
#pragma unroll
                for (int v = 0; v < VEC; ++v) t.v[v] = mn[v];
                st_vec<VEC>(w + 2 * p.F, t);
#pragma unroll
                for (int v = 0; v < VEC; ++v) t.v[v] = mx[v];
                st_vec<VEC>(w + 3 * p.F, t);
#pragma unroll
                for (int v = 0; v < VEC; ++v) t.v[v] = __int_as_float(amn[v]);
                st_vec<VEC>(w + 4 * p.F, t);
#pragma unroll
                for (int v = 0; v < VEC; ++v) t.v[v] = __int_as_float(amx[v]);
                st_vec<VEC>(w + 5 * p.F, t);
                return;
            }
            emit_row<VEC, MINMAX, SQ>(p, simple_out, row_end - row_beg, sum, sq, mn, mx, amn, amx, yp, amn_p, amx_p,
                                      mean_p, var_p);
        };
        auto issue_row_loads = [&]() {                      // next_end = rp[row + 2], P of row + 1, indices of row + 2
            next_end = row + 2 <= n_rows ? __ldg(rp + row + 2) : 0x7fffffff;
            fetch_P(row + 1, in1, pv_next);
            fetch_idx(row + 2, in2);
            pending = false;
        };
        auto row_next = [&]() {
            finish_row();
            if (pending) issue_row_loads();                 // the row had no stage (an empty row): fetch now, and wait
            ++row;
            if (in1.d.real != rd.real) {                    // segments of one row share the output row
                yp += p.ldy;
                if (amn_p) amn_p += p.F;
                if (amx_p) amx_p += p.F;
                if (mean_p) mean_p += p.F;
                if (var_p) var_p += p.F;
            }
            row_beg = row_end;
            row_end = next_end;
            make_current(in1, pv_next);
            in1 = in2;
            pending = true;
            reset_acc();
        };

        {
            RowIn in0;
            fetch_idx(row, in0);
            fetch_idx(row + 1, in1);
            fetch_idx(row + 2, in2);
            Vec<VEC> raw;
            fetch_P(row, in0, raw);
            fetch_P(row + 1, in1, pv_next);
            make_current(in0, raw);
        }
        point_at(rd.real);
        reset_acc();

        if (len > 0) {
            ring.begin(p.col + s0, rp, n_rows, r0, s0, len);
            int c_pos = 0, cons_i = 0;
#pragma unroll 1
            while (c_pos < len) {
                const int at = s0 + c_pos;
                if (at == row_end) { row_next(); continue; }
                if (pending) issue_row_loads();
                const int n = row_end - at < 4 ? row_end - at : 4;
                const uint32_t sb = ring.base + (uint32_t)(cons_i * 4 * ring.rowb);
                if (n == 4) {                                                 // a full stage: 4 edges of this row
                    Vec<VEC> q[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) q[i] = lds_vec<VEC>(sb + (uint32_t)(i * ring.rowb));
                    rng_refresh<DROP, VEC>(drop, rid, pos, c, bits);
                    uint32_t w[VEC];
                    const int sh = DROP == FD_BIT ? (pos & 31) : 0;
#pragma unroll
                    for (int v = 0; v < VEC; ++v) w[v] = bits[v] >> sh;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            const float x = fast_message<MSG_PQ, DROP>(pv.v[v], q[u].v[v], 0.0f, scale);
                            accumulate<false, DROP != FD_NONE, MINMAX, SQ>(x, keep_word<DROP>(w[v], u, thr), 1u, at + u,
                                                                           sum[v], sq[v], mn[v], mx[v], amn[v], amx[v]);
                        }
                    }
                    pos += 4;
                } else {                                                      // the row's last, short stage
#pragma unroll 1
                    for (int u = 0; u < n; ++u) {
                        const Vec<VEC> q = lds_vec<VEC>(sb + (uint32_t)(u * ring.rowb));
                        rng_refresh<DROP, VEC>(drop, rid, pos, c, bits);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            const float x = fast_message<MSG_PQ, DROP>(pv.v[v], q.v[v], 0.0f, scale);
                            accumulate<false, DROP != FD_NONE, MINMAX, SQ>(x, keep_at<DROP>(bits[v], pos, thr), 1u, at + u,
                                                                           sum[v], sq[v], mn[v], mx[v], amn[v], amx[v]);
                        }
                        ++pos;
                    }
                }
                c_pos += n;
                ring.advance();
                cons_i = cons_i + 1 == NST ? 0 : cons_i + 1;
            }
            ring.end();
        }
        // the last row with edges, then trailing empty rows
        for (;;) {
            if (row + 1 >= r1) { finish_row(); break; }
            row_next();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward, destination pass
// ---------------------------------------------------------------------------------------------
template <int NST, int WARPS, int VEC, int DROP, bool NEEDM, bool LOCAL>
__global__ void __launch_bounds__(WARPS * 32, 1) mmconv_bwd_stream(const __grid_constant__ MMConvParams p) {
    extern __shared__ __align__(16) uint8_t smem_ring[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t gw = (int64_t)blockIdx.x * WARPS + warp, tw = (int64_t)gridDim.x * WARPS;
    const int64_t n_chunks = p.row_chunks ? p.n_chunks : tw;
    const bool live = lane * VEC < p.ncols;
    const int c = p.col0 + lane * VEC;
    const Dropout drop = resolve_key(p.drop);
    const float scale = DROP == FD_BIT ? 2.0f : (DROP == FD_BYTE ? p.drop.scale : 1.0f);
    const uint32_t thr = p.drop.thr;
    const bool simple_out = p.simple_out != 0;

    Ring<NST, NEEDM, VEC> ring;
    ring.rowb = p.ncols * 4;
    ring.base = smem_addr(smem_ring) + (uint32_t)warp * (uint32_t)(NST * 4 * ring.rowb) + (live ? (uint32_t)(lane * VEC * 4) : 0u);   // idle lanes stay inside the ring
    ring.Qc = reinterpret_cast<const char *>(p.Q + c);
    ring.ldq_b = (uint32_t)p.ldq * 4u;
    ring.lane = lane; ring.live = live;
    char *Gc = p.G ? reinterpret_cast<char *>(p.G + c) : nullptr;
    const uint32_t ldg_b = (uint32_t)p.ldg * 4u;
    const int32_t *rp = p.vrowptr ? p.vrowptr : p.rowptr;      // the (virtual) row boundaries walked by the stream
    const int64_t n_rows = p.vrowptr ? p.n_vrows : p.n_rows;
    const int ycol = p.T == 1 ? c : (c / p.F_in) * (p.S * p.A * p.F_in) + c % p.F_in;

#pragma unroll 1
    for (int64_t chunk = gw; chunk < n_chunks; chunk += tw) {
        int r0, r1;
        chunk_rows(p, rp, n_rows, chunk, n_chunks, r0, r1);
        if (r0 >= r1) continue;
        const int s0 = __ldg(rp + r0);
        const int len = __ldg(rp + r1) - s0;

        int row = r0;
        int row_beg = s0, row_end = __ldg(rp + r0 + 1);
        Vec<VEC> base{}, gmin{}, gmax{}, alpha{}, pv{}, dp{};
        int32_t amn[VEC], amx[VEC];
        uint32_t rid = 0;
        uint32_t bits[VEC] = {};
        int pos = 0;
        int64_t prow = 0;
        RowDesc rd{};
        const float *dyp = nullptr;
        int64_t rowF = 0;

        auto start_row = [&]() {          // folds dY of `row` into base / gmin / gmax / alpha (times the keep-scale)
            rd = row_desc(p, row);
            pos = rd.pos0;
            base = Vec<VEC>{}; gmin = Vec<VEC>{}; gmax = Vec<VEC>{}; alpha = Vec<VEC>{}; pv = Vec<VEC>{}; dp = Vec<VEC>{};
#pragma unroll
            for (int v = 0; v < VEC; ++v) { amn[v] = -1; amx[v] = -1; }
            prow = p.row_map ? (int64_t)__ldg(p.row_map + rd.real) : (int64_t)rd.real;
            if (p.use_rng) rid = (uint32_t)(p.rng_row0 + (p.rng_row ? (int64_t)__ldg(p.rng_row + rd.real) : (int64_t)rd.real));
            if (!live) return;
            dyp = p.dY + (int64_t)rd.real * p.ldy + ycol;
            rowF = (int64_t)rd.real * p.F + c;
            // a segment of a split row needs the degree of the whole row
            const int deg = rd.slot >= 0 ? __ldg(p.rowptr + rd.real + 1) - __ldg(p.rowptr + rd.real) : row_end - row_beg;
            const int degc = deg > 1 ? deg : 1;
            const float degf = (float)degc;
            const float rdeg = 1.0f / degf;
            Vec<VEC> mean{}, var{};
            if constexpr (NEEDM) {
                mean = ld_vec_stream<VEC>(p.c_mean + rowF);
                var = ld_vec_stream<VEC>(p.c_var + rowF);
            }
            if (simple_out) {
                // every kind at most once, S == 1: straight-line fold
                if (p.zoff[MMA_AGGR_SUM] >= 0) {
                    const Vec<VEC> d = ld_vec_stream<VEC>(dyp + p.zoff[MMA_AGGR_SUM]);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) base.v[v] += d.v[v];
                }
                if (p.zoff[MMA_AGGR_MEAN] >= 0) {
                    const Vec<VEC> d = ld_vec_stream<VEC>(dyp + p.zoff[MMA_AGGR_MEAN]);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) base.v[v] += d.v[v] * rdeg;
                }
                if (p.zoff[MMA_AGGR_MIN] >= 0) {
                    gmin = ld_vec_stream<VEC>(dyp + p.zoff[MMA_AGGR_MIN]);
                    ld_vec_i32_as<VEC>(p.c_arg_min + rowF, amn);
                }
                if (p.zoff[MMA_AGGR_MAX] >= 0) {
                    gmax = ld_vec_stream<VEC>(dyp + p.zoff[MMA_AGGR_MAX]);
                    ld_vec_i32_as<VEC>(p.c_arg_max + rowF, amx);
                }
                if constexpr (NEEDM) {
                    if (p.zoff[MMA_AGGR_VAR] >= 0) {    // var = E[m^2] - E[m]^2 -> d/dm_e = 2 (m_e - mean) / cnt
                        const Vec<VEC> d = ld_vec_stream<VEC>(dyp + p.zoff[MMA_AGGR_VAR]);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            const float k = 2.0f * d.v[v] * rdeg;
                            alpha.v[v] += k; base.v[v] -= k * mean.v[v];
                        }
                    }
                    if (p.zoff[MMA_AGGR_STD] >= 0) {    // std = sqrt(relu(var) + 1e-5); relu'(0) = 0
                        const Vec<VEC> d = ld_vec_stream<VEC>(dyp + p.zoff[MMA_AGGR_STD]);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            if (var.v[v] > 0.0f) {
                                const float k = d.v[v] * rdeg * rsqrtf(var.v[v] + 1e-5f);
                                alpha.v[v] += k; base.v[v] -= k * mean.v[v];
                            }
                        }
                    }
                }
            } else {
                float fac[MMA_MAX_SCALER];
                scaler_factors(p, degc, fac);
                for (int s = 1; s < p.S; ++s) fac[s] *= fac[s - 1];
                bool has_min = false, has_max = false;
                for (int a = 0; a < p.A; ++a) {
                    Vec<VEC> dz{};
                    for (int s = 0; s < p.S; ++s) {
                        const Vec<VEC> d = ld_vec_stream<VEC>(dyp + (int64_t)(s * p.A + a) * p.F_in);
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
                        else if (kind == MMA_AGGR_VAR) {
                            const float k = 2.0f * gg / degf;
                            alpha.v[v] += k; base.v[v] -= k * mean.v[v];
                        } else if (var.v[v] > 0.0f) {
                            const float k = gg / (sqrtf(var.v[v] + 1e-5f) * degf);
                            alpha.v[v] += k; base.v[v] -= k * mean.v[v];
                        }
                    }
                    has_min |= kind == MMA_AGGR_MIN;
                    has_max |= kind == MMA_AGGR_MAX;
                }
                if (has_min) ld_vec_i32_as<VEC>(p.c_arg_min + rowF, amn);
                if (has_max) ld_vec_i32_as<VEC>(p.c_arg_max + rowF, amx);
            }
            if (NEEDM && p.P) pv = ld_vec_stream<VEC>(p.P + prow * p.ldp + c);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                base.v[v] *= scale; gmin.v[v] *= scale; gmax.v[v] *= scale; alpha.v[v] *= scale;
                if constexpr (DROP == FD_BIT) pv.v[v] *= 2.0f;
            }
        };
        auto finish_row = [&]() {
            if (!live) return;
            if (rd.slot >= 0) st_vec<VEC>(p.seg_ws + (int64_t)rd.slot * p.F + c, dp);      // partial dP of a segment
            else if (p.dP) st_vec_stream<VEC>(p.dP + prow * p.lddp + c, dp);
        };
        auto row_next = [&]() {
            finish_row();
            ++row;
            row_beg = row_end;
            row_end = __ldg(rp + row + 1);
            start_row();
        };

        start_row();
        if (len > 0) {
            auto load_slot = [&](int pbase) {
                const int i = pbase + lane;
                return i < len ? (p.gslot ? __ldg(p.gslot + s0 + i) : s0 + i) : 0;
            };
            auto load_eid = [&](int pbase) {               // global original edge id (public arg indices)
                const int i = pbase + lane;
                if (i >= len) return 0;
                const int e = p.perm ? __ldg(p.perm + s0 + i) : s0 + i;
                return p.gid ? __ldg(p.gid + e) : e;
            };
            EdgeAttr gslot, eid;
            gslot.init(load_slot);
            if constexpr (!LOCAL) eid.init(load_eid);
            else { eid.cur = 0; eid.nxt = 0; }

            auto one_edge = [&](int at, int s, const Vec<VEC> &q, const uint32_t (&kw)[VEC]) {
                const int gs = gslot.get(s);
                int id = at;
                if constexpr (!LOCAL) id = eid.get(s);
                Vec<VEC> gr;
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    float t = base.v[v];
                    if constexpr (NEEDM) {
                        const float x = fast_message<MSG_PQ, DROP>(pv.v[v], q.v[v], 0.0f, scale);
                        t = __fmaf_rn(alpha.v[v], x, t);
                    }
                    route_grad<false>(t, id, amn[v], amx[v], gmin.v[v], gmax.v[v], kw[v], 1u, gr.v[v], dp.v[v]);
                }
                if (live && Gc) st_vec_stream<VEC>(reinterpret_cast<float *>(Gc + (uint64_t)(uint32_t)gs * ldg_b), gr);
            };

            ring.begin(p.col + s0, len);
            run_stream<NST, NEEDM, VEC>(ring, s0, len, row_end, row_next,
                [&](int at, const Vec<VEC> (&q)[4], int s) {
                    rng_refresh<DROP, VEC>(drop, rid, pos, c, bits);
                    uint32_t w[VEC];
                    const int sh = DROP == FD_BIT ? (pos & 31) : 0;
#pragma unroll
                    for (int v = 0; v < VEC; ++v) w[v] = bits[v] >> sh;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        uint32_t kw[VEC];
#pragma unroll
                        for (int v = 0; v < VEC; ++v) kw[v] = keep_word<DROP>(w[v], u, thr);
                        one_edge(at + u, s + u, q[u], kw);
                    }
                    pos += 4;
                },
                [&](int at, const Vec<VEC> &q, int s) {
                    rng_refresh<DROP, VEC>(drop, rid, pos, c, bits);
                    uint32_t kw[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; ++v) kw[v] = keep_at<DROP>(bits[v], pos, thr);
                    one_edge(at, s, q, kw);
                    ++pos;
                },
                [&]() { return DROP == FD_BIT ? (pos & 31) <= 28 : (DROP == FD_BYTE ? (pos & 3) == 0 : true); },
                [&](int s) {
                    gslot.at_stage(s, load_slot);
                    if constexpr (!LOCAL) eid.at_stage(s, load_eid);
                });
        }
        for (;;) {
            if (row + 1 >= r1) { finish_row(); break; }
            row_next();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// split rows: merge the segments' partial states in segment order
// ---------------------------------------------------------------------------------------------
// forward: one block per split row, a thread per 4 columns.  split_tab[i] = {real row, first slot, segments, 0}
template <bool MINMAX, bool SQ>
__global__ void __launch_bounds__(128) mmconv_fwd_merge(const __grid_constant__ MMConvParams p) {
    const int4 t = __ldg(reinterpret_cast<const int4 *>(p.split_tab) + blockIdx.x);
    const int real = t.x, slot0 = t.y, nseg = t.z;
    const int deg = __ldg(p.rowptr + real + 1) - __ldg(p.rowptr + real);
    for (int c = p.col0 + threadIdx.x * 4; c < p.col0 + p.ncols; c += blockDim.x * 4) {
        float sum[4], sq[4], mn[4], mx[4];
        int amn[4], amx[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) { sum[v] = 0.f; sq[v] = 0.f; mn[v] = FLT_MAX; mx[v] = -FLT_MAX; amn[v] = -1; amx[v] = -1; }
        for (int k = 0; k < nseg; ++k) {
            const float *w = p.seg_ws + (int64_t)(slot0 + k) * 6 * p.F + c;
            const Vec<4> ps = ld_vec<4>(w), pq = ld_vec<4>(w + p.F), pmn = ld_vec<4>(w + 2 * p.F), pmx = ld_vec<4>(w + 3 * p.F);
            const Vec<4> pan = ld_vec<4>(w + 4 * p.F), pax = ld_vec<4>(w + 5 * p.F);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                sum[v] = __fadd_rn(sum[v], ps.v[v]);
                sq[v] = __fadd_rn(sq[v], pq.v[v]);
                const int an = __float_as_int(pan.v[v]), ax = __float_as_int(pax.v[v]);
                if (an >= 0 && pmn.v[v] < mn[v]) { mn[v] = pmn.v[v]; amn[v] = an; }     // strict: the earlier segment wins ties
                if (ax >= 0 && pmx.v[v] > mx[v]) { mx[v] = pmx.v[v]; amx[v] = ax; }
            }
        }
        const int ycol = p.T == 1 ? c : (c / p.F_in) * (p.S * p.A * p.F_in) + c % p.F_in;
        const int64_t rowF = (int64_t)real * p.F + c;
        emit_row<4, MINMAX, SQ>(p, p.simple_out != 0, deg, sum, sq, mn, mx, amn, amx, p.Y + (int64_t)real * p.ldy + ycol,
                                p.arg_min ? p.arg_min + rowF : nullptr, p.arg_max ? p.arg_max + rowF : nullptr,
                                p.stat_mean ? p.stat_mean + rowF : nullptr, p.stat_var ? p.stat_var + rowF : nullptr);
    }
}

// backward: dP of a split row = sum of its segments' partial dP, in segment order
__global__ void __launch_bounds__(128) mmconv_bwd_merge(const __grid_constant__ MMConvParams p) {
    const int4 t = __ldg(reinterpret_cast<const int4 *>(p.split_tab) + blockIdx.x);
    const int real = t.x, slot0 = t.y, nseg = t.z;
    if (!p.dP) return;
    const int64_t prow = p.row_map ? (int64_t)__ldg(p.row_map + real) : (int64_t)real;
    for (int c = p.col0 + threadIdx.x * 4; c < p.col0 + p.ncols; c += blockDim.x * 4) {
        Vec<4> acc{};
        for (int k = 0; k < nseg; ++k) {
            const Vec<4> d = ld_vec<4>(p.seg_ws + (int64_t)(slot0 + k) * p.F + c);
#pragma unroll
            for (int v = 0; v < 4; ++v) acc.v[v] += d.v[v];
        }
        st_vec_stream<4>(p.dP + prow * p.lddp + c, acc);
    }
}

}  // namespace stream
