// Materialises the in-kernel Philox keep-scale so tests can inject the identical dropout
// mask into the CPU oracle (the reference's F.dropout, mma_conv.py:157 / layers.py:219,
// uses torch's CPU mt19937 stream, which no GPU kernel can reproduce: SURVEY.md "Hard parts").
#include "common.cuh"

namespace mma {
__global__ void keep_scale_kernel(Dropout d, uint32_t stream_id, int64_t E, int F, float *out, int64_t ldo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E * F) return;
    const int64_t e = i / F;
    const int c = (int)(i - e * F);
    out[e * ldo + c] = dropout_keep<1>(d, (uint32_t)e, c, stream_id).v[0];
}

// one warp per CSR row: slot k of row i, column c -> out[perm[k], c]
__global__ void keep_scale_rows_kernel(Dropout d, const int32_t *rowptr, const int32_t *perm, const int32_t *rng_row,
                                       int64_t rng_row0, int64_t n_rows, int F, float *out, int64_t ldo) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const uint32_t rid = (uint32_t)(rng_row0 + (rng_row ? (int64_t)rng_row[row] : row));
    const int beg = rowptr[row], end = rowptr[row + 1];
    for (int k = beg; k < end; ++k) {
        const int64_t e = perm ? perm[k] : k;
        for (int c = lane; c < F; c += 32)
            out[e * ldo + c] = dropout_keep_row<1>(d, rid, (uint32_t)(k - beg), c).v[0];
    }
}
}  // namespace mma

extern "C" int mma_dropout_keep_scale_rows(const int32_t *rowptr, const int32_t *perm, const int32_t *rng_row,
                                           int64_t rng_row0, int64_t n_rows, int64_t E, float p_drop,
                                           uint64_t seed, int F, float *out, int64_t ldo, mma_stream_t stream) {
    if (!rowptr || !out || n_rows < 0 || E < 0 || F < 1 || p_drop < 0.0f || p_drop > 1.0f) return MMA_ERR_INVALID;
    if (E == 0 || n_rows == 0) return MMA_OK;
    const int64_t threads = n_rows * 32;
    mma::keep_scale_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        mma::make_dropout(p_drop, seed), rowptr, perm, rng_row, rng_row0, n_rows, F, out, ldo);
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}

extern "C" int mma_dropout_keep_scale(float p_drop, uint64_t seed, uint32_t stream_id, int64_t E, int F,
                                      float *out, int64_t ldo, mma_stream_t stream) {
    if (!out || E < 0 || F < 1 || p_drop < 0.0f || p_drop > 1.0f) return MMA_ERR_INVALID;
    if (E == 0) return MMA_OK;
    const int64_t n = E * F;
    mma::keep_scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        mma::make_dropout(p_drop, seed), stream_id, E, F, out, ldo);
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}
