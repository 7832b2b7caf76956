// K3 / transpose pass: deterministic segmented row sum  out[i] = sum_k val[k] * src[idx[k]].
//
// Replaces (paths relative to /root/reference):
//   node_classification/layers.py:41, :862   torch.spmm(adj, support)       (cuSPARSE COO SpMM)
//   autograd of graph_regression/mma_conv.py:130  index_select-backward == index_add_ with atomics
//     -> here: dQ[j] = sum of the per-edge gradient rows of j's out-edges, in CSC order, no atomics.
// Same mapping as K1: a group of LANES threads owns one output row x one chunk of LANES*VEC
// columns and walks its segment sequentially (fixed summation order => bit-reproducible).
#include "common.cuh"

namespace mma {

struct SegParams {
    const int32_t *ptr, *idx;
    const float *val, *src;
    float *out;
    int64_t n_rows, lds, ldo;
    int F, lanes_log2, chunks;
    int64_t n_groups;
};

template <int VEC>
__global__ void __launch_bounds__(256) segment_sum_rows_kernel(const __grid_constant__ SegParams p) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t group = tid >> p.lanes_log2;
    if (group >= p.n_groups) return;
    const int sub = (int)(tid & ((1 << p.lanes_log2) - 1));
    const int64_t row = group / p.chunks;
    const int chunk = (int)(group - row * p.chunks);
    const int c = ((chunk << p.lanes_log2) + sub) * VEC;
    if (c >= p.F) return;
    const int beg = __ldg(p.ptr + row), end = __ldg(p.ptr + row + 1);
    Vec<VEC> acc{};
    constexpr int U = (VEC == 4) ? 4 : 8;
    int pos = beg;
    for (; pos + U <= end; pos += U) {
        int64_t j[U];
        float w[U];
        Vec<VEC> x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) j[u] = p.idx ? (int64_t)__ldg(p.idx + pos + u) : (int64_t)(pos + u);
#pragma unroll
        for (int u = 0; u < U; ++u) w[u] = p.val ? __ldg(p.val + pos + u) : 1.0f;
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = ld_vec_stream<VEC>(p.src + j[u] * p.lds + c);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc.v[v] += w[u] * x[u].v[v];
    }
    for (; pos < end; ++pos) {
        const int64_t j = p.idx ? (int64_t)__ldg(p.idx + pos) : (int64_t)pos;
        const float w = p.val ? __ldg(p.val + pos) : 1.0f;
        const Vec<VEC> x = ld_vec_stream<VEC>(p.src + j * p.lds + c);
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc.v[v] += w * x.v[v];
    }
    st_vec_stream<VEC>(p.out + row * p.ldo + c, acc);
}


// out[r] = src[idx[r]] for r < n_out, plus (optionally) the column sums of the gathered rows as
// per-part partial sums: part p owns the contiguous rows [p * per, (p + 1) * per) and writes
// colsum_part[p][c]; summing the parts in ascending p (mma_reduce_slabs) is a fixed-order reduction.
// One warp per part; a lane owns 4-column groups lane, lane + 32, ... (F % 4 == 0).
__global__ void __launch_bounds__(256) gather_rows_kernel(const float *__restrict__ src, int64_t lds,
                                                          const int32_t *__restrict__ idx, int64_t n_out, int F,
                                                          float *__restrict__ out, int64_t ldo,
                                                          float *__restrict__ colsum_part, int64_t n_parts) {
    const int64_t part = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (part >= n_parts) return;
    const int64_t per = (n_out + n_parts - 1) / n_parts;
    const int64_t r0 = part * per, r1 = r0 + per < n_out ? r0 + per : n_out;
    for (int c = lane * 4; c < F; c += 128) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int64_t r = r0;
        for (; r + 4 <= r1; r += 4) {          // 4 independent gathers in flight
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t j = idx ? (int64_t)__ldg(idx + r + u) : r + u;
                v[u] = __ldcs(reinterpret_cast<const float4 *>(src + j * lds + c));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (out) __stcs(reinterpret_cast<float4 *>(out + (r + u) * ldo + c), v[u]);
                acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
            }
        }
        for (; r < r1; ++r) {
            const int64_t j = idx ? (int64_t)__ldg(idx + r) : r;
            const float4 v = __ldcs(reinterpret_cast<const float4 *>(src + j * lds + c));
            if (out) __stcs(reinterpret_cast<float4 *>(out + r * ldo + c), v);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        if (colsum_part) *reinterpret_cast<float4 *>(colsum_part + part * F + c) = acc;
    }
}

}  // namespace mma

using namespace mma;

extern "C" int mma_gather_rows(const float *src, int64_t lds, const int32_t *idx, int64_t n_out, int F,
                               float *out, int64_t ldo, float *colsum_part, int64_t n_parts, mma_stream_t stream) {
    if (!src || (!out && !colsum_part) || n_out < 0 || F < 1 || n_parts < 1) return MMA_ERR_INVALID;
    if ((F % 4) != 0 || (lds % 4) != 0 || (out && (ldo % 4) != 0) || !aligned16(src) || !aligned16(out) || !aligned16(colsum_part))
        return MMA_ERR_UNSUPPORTED;
    if (n_parts * 32 > (int64_t)INT32_MAX * 256) return MMA_ERR_UNSUPPORTED;
    const int64_t threads = n_parts * 32;
    gather_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        src, lds, idx, n_out, F, out, ldo, colsum_part, n_parts);
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}

extern "C" int mma_segment_sum_rows(const int32_t *ptr, const int32_t *idx, const float *val,
                                    int64_t n_rows, const float *src, int64_t lds, int F,
                                    float *out, int64_t ldo, mma_stream_t stream) {
    if (!ptr || !src || !out || n_rows < 0 || F < 1) return MMA_ERR_INVALID;
    if (n_rows == 0) return MMA_OK;
    if (n_rows >= INT32_MAX) return MMA_ERR_UNSUPPORTED;
    SegParams p{};
    p.ptr = ptr; p.idx = idx; p.val = val; p.src = src; p.out = out;
    p.n_rows = n_rows; p.lds = lds; p.ldo = ldo; p.F = F;
    const bool v4 = (F % 4 == 0) && aligned16(src) && aligned16(out) && lds % 4 == 0 && ldo % 4 == 0;
    const int vec = v4 ? 4 : 1;
    const int per_row = (F + vec - 1) / vec;
    int lg = 0;
    while ((1 << lg) < per_row && lg < 5) ++lg;
    p.lanes_log2 = lg;
    p.chunks = (per_row + (1 << lg) - 1) >> lg;
    p.n_groups = n_rows * p.chunks;
    const int block = 256;
    const int64_t grid = ((p.n_groups << lg) + block - 1) / block;
    if (grid > INT32_MAX) return MMA_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (vec == 4) segment_sum_rows_kernel<4><<<(unsigned)grid, block, 0, st>>>(p);
    else segment_sum_rows_kernel<1><<<(unsigned)grid, block, 0, st>>>(p);
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}
