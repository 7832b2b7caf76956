// Shared device helpers for the mma_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "../../include/mma_b200.h"

namespace mma {

void set_last_error(cudaError_t e);

#define MMA_CUDA_CHECK(expr)                                   \
    do {                                                       \
        cudaError_t _e = (expr);                               \
        if (_e != cudaSuccess) { ::mma::set_last_error(_e); return MMA_ERR_CUDA; } \
    } while (0)

#define MMA_LAUNCH_CHECK() MMA_CUDA_CHECK(cudaGetLastError())

constexpr int kSMs = 148;   // B200: 2 dies x 74 SMs

// ---------------------------------------------------------------------------
// VEC-wide register tiles.  VEC = 4 -> one 128-bit access per lane; VEC = 1 is
// the fallback for widths/alignments that do not allow it (e.g. F_in = 75).
// ---------------------------------------------------------------------------
template <int VEC> struct Vec { float v[VEC]; };

template <int VEC>
__device__ __forceinline__ Vec<VEC> ld_vec(const float *p) {
    Vec<VEC> r;
    if constexpr (VEC == 4) {
        float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else if constexpr (VEC == 2) {
        float2 t = __ldg(reinterpret_cast<const float2 *>(p));
        r.v[0] = t.x; r.v[1] = t.y;
    } else {
        r.v[0] = __ldg(p);
    }
    return r;
}

// gather of a feature row that is touched ~deg times over the whole kernel but far apart in
// time: ld.global.cg (L2 only) keeps it out of L1 so index lines stay resident.  An intrinsic, not
// inline asm, so the compiler can predicate it instead of branching around it.
template <int VEC>
__device__ __forceinline__ Vec<VEC> ld_vec_stream(const float *p) {
    Vec<VEC> r;
    if constexpr (VEC == 4) {
        const float4 t = __ldcg(reinterpret_cast<const float4 *>(p));
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else if constexpr (VEC == 2) {
        const float2 t = __ldcg(reinterpret_cast<const float2 *>(p));
        r.v[0] = t.x; r.v[1] = t.y;
    } else {
        r.v[0] = __ldcg(p);
    }
    return r;
}

template <int VEC>
__device__ __forceinline__ Vec<VEC> ld_vec_i32_as(const int32_t *p, int32_t *out) {
    Vec<VEC> dummy{};
    if constexpr (VEC == 4) {
        int4 t = __ldg(reinterpret_cast<const int4 *>(p));
        out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
    } else if constexpr (VEC == 2) {
        int2 t = __ldg(reinterpret_cast<const int2 *>(p));
        out[0] = t.x; out[1] = t.y;
    } else {
        out[0] = __ldg(p);
    }
    return dummy;
}

template <int VEC>
__device__ __forceinline__ void st_vec(float *p, const Vec<VEC> &r) {
    if constexpr (VEC == 4) {
        *reinterpret_cast<float4 *>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    } else if constexpr (VEC == 2) {
        *reinterpret_cast<float2 *>(p) = make_float2(r.v[0], r.v[1]);
    } else {
        *p = r.v[0];
    }
}

// write-once outputs: streaming store (evict-first) so they do not displace
// the gathered feature rows from the 126 MB L2.
template <int VEC>
__device__ __forceinline__ void st_vec_stream(float *p, const Vec<VEC> &r) {
    if constexpr (VEC == 4) {
        __stcs(reinterpret_cast<float4 *>(p), make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
    } else if constexpr (VEC == 2) {
        __stcs(reinterpret_cast<float2 *>(p), make_float2(r.v[0], r.v[1]));
    } else {
        __stcs(p, r.v[0]);
    }
}

template <int VEC>
__device__ __forceinline__ void st_vec_i32_stream(int32_t *p, const int32_t *r) {
    if constexpr (VEC == 4) {
        __stcs(reinterpret_cast<int4 *>(p), make_int4(r[0], r[1], r[2], r[3]));
    } else if constexpr (VEC == 2) {
        __stcs(reinterpret_cast<int2 *>(p), make_int2(r[0], r[1]));
    } else {
        __stcs(p, r[0]);
    }
}

// ---------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al., SC'11).  The dropout decision
// of element (edge id e, column c) of mask stream s depends only on
// (seed, e, c/4, s): forward, both backward passes and every shard of a
// partitioned graph regenerate the same bits without storing a mask.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}

// Dropout resolution is 8 bits per element: p is quantised to thr/256 (exact for the
// reference's 0.5 and 0.75) and the keep-scale is 1/(1 - thr/256), so the estimator stays
// unbiased.  One Philox call covers 16 consecutive columns of one edge: call (eid, c >> 4,
// stream); column c uses byte (c & 3) of word ((c >> 2) & 3).
struct Dropout {
    uint32_t thr;     // drop iff random byte < thr   (thr = round(p * 256), 0..256)
    float scale;      // 1 / (1 - thr/256)
    uint32_t k0, k1;  // seed
    const unsigned long long *seed_dev;   // if set, the seed is read from device memory at run time instead: a
                                          // launch captured in a CUDA graph then draws a fresh mask on every replay
};

__device__ __forceinline__ uint2 dropout_key(const Dropout &d) {
    if (d.seed_dev) {
        const unsigned long long s = __ldg(d.seed_dev);
        return make_uint2((uint32_t)(s & 0xFFFFFFFFull), (uint32_t)(s >> 32));
    }
    return make_uint2(d.k0, d.k1);
}

__host__ inline Dropout make_dropout(float p, uint64_t seed, const uint64_t *seed_dev = nullptr) {
    Dropout d;
    d.seed_dev = reinterpret_cast<const unsigned long long *>(seed_dev);
    int t = (int)(p * 256.0f + 0.5f);
    d.thr = t < 0 ? 0u : (t > 256 ? 256u : (uint32_t)t);
    d.scale = d.thr < 256u ? 256.0f / (float)(256u - d.thr) : 0.0f;
    d.k0 = (uint32_t)(seed & 0xFFFFFFFFull);
    d.k1 = (uint32_t)(seed >> 32);
    return d;
}

// keep-scale of VEC consecutive columns whose random bytes start at byte `b0` of word `w`
template <int VEC>
__device__ __forceinline__ Vec<VEC> keep_from_word(const Dropout &d, uint32_t w, int b0) {
    Vec<VEC> r;
#pragma unroll
    for (int v = 0; v < VEC; ++v) r.v[v] = (((w >> (8 * ((b0 + v) & 3))) & 0xFFu) < d.thr) ? 0.0f : d.scale;
    return r;
}

// keep-scale (0 or 1/(1-p)) for VEC consecutive columns starting at c (c % VEC == 0, VEC in {1,2,4}).
template <int VEC>
__device__ __forceinline__ Vec<VEC> dropout_keep(const Dropout &d, uint32_t eid, int c, uint32_t stream_id) {
    const uint4 bits = philox4x32_10(make_uint4(eid, (uint32_t)(c >> 4), stream_id, 0u), dropout_key(d));
    const int ws = (c >> 2) & 3;
    const uint32_t w = ws == 0 ? bits.x : (ws == 1 ? bits.y : (ws == 2 ? bits.z : bits.w));
    return keep_from_word<VEC>(d, w, c & 3);
}

// ---------------------------------------------------------------------------
// K1's dropout stream is keyed by (destination node id, position of the edge inside that node's
// in-edge list, column): the CSR is a STABLE sort by destination, so the in-row position of an edge
// does not depend on how rows are permuted, relabelled or sharded -- forward, backward and every
// shard regenerate the same bits, and a lane that owns 4 columns of a row generates exactly the bits
// it consumes (no exchange between lanes).
//   p == 0.5 (thr == 128, the reference's constant): ONE BIT per element.  Philox counter
//     (row id, pos >> 5, column >> 2, 0): word v <-> column 4*(c>>2) + v, bit (pos & 31) <-> edge.
//     keep iff the bit is set.  One call covers 32 edges x 4 columns.
//   any other p: one BYTE per element.  Counter (row id, pos >> 2, column >> 2, 1): word v <-> column,
//     byte (pos & 3) <-> edge; drop iff byte < thr.  One call covers 4 edges x 4 columns.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool dropout_one_bit(const Dropout &d) { return d.thr == 128u; }

__device__ __forceinline__ uint4 row_rng_bits1(const Dropout &d, uint32_t row_id, uint32_t pos, uint32_t c) {
    return philox4x32_10(make_uint4(row_id, pos >> 5, c >> 2, 0u), dropout_key(d));
}
__device__ __forceinline__ uint4 row_rng_bits8(const Dropout &d, uint32_t row_id, uint32_t pos, uint32_t c) {
    return philox4x32_10(make_uint4(row_id, pos >> 2, c >> 2, 1u), dropout_key(d));
}

// keep-scale of VEC consecutive columns starting at c for in-row edge `pos` of row `row_id`
// (slow per-edge form for the generic kernels; the fast kernels amortise the Philox call).
template <int VEC>
__device__ __forceinline__ Vec<VEC> dropout_keep_row(const Dropout &d, uint32_t row_id, uint32_t pos, int c) {
    Vec<VEC> r;
    const bool one = dropout_one_bit(d);
    const uint4 b = one ? row_rng_bits1(d, row_id, pos, (uint32_t)c) : row_rng_bits8(d, row_id, pos, (uint32_t)c);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const int ws = (c + v) & 3;                     // VEC > 1 implies c % VEC == 0: same Philox call
        const uint32_t w = ws == 0 ? b.x : (ws == 1 ? b.y : (ws == 2 ? b.z : b.w));
        bool keep;
        if (one) keep = (w >> (pos & 31u)) & 1u;
        else keep = ((w >> (8u * (pos & 3u))) & 0xFFu) >= d.thr;
        r.v[v] = keep ? d.scale : 0.0f;
    }
    return r;
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace mma
