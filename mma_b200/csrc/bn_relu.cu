// BatchNorm (batch or given statistics) + ReLU over the rows of a node-feature matrix, one kernel per direction.
//
// Replaces `x = F.relu(batch_norm(conv(x, edge_index, edge_attr)))` of the reference's Net
// (/root/reference/graph_regression/mma.py:120-121; torch_geometric.nn.BatchNorm wraps BatchNorm1d): the step right
// after the MultiMaskConv layer.  In eval mode the normalisation is a per-channel affine map
// and is folded into the layer's `lin` weight with the ReLU in the GEMM epilogue (MMA_GEMM_RELU, gemm_tf32x3.cu); in
// training mode the statistics are over all rows of the GEMM's OUTPUT, so they cannot live in its epilogue: this kernel
// is the whole BatchNorm + ReLU in one launch instead of PyTorch's stats / normalise / relu (and their three backward
// kernels).
//
// One CTA per 32 columns, 32 x 32 threads (x: column, y: row lane).  Rows are reduced lane by lane in a fixed order
// and the 32 lanes are combined through shared memory in ascending lane order: no atomics, bit-reproducible.  The
// statistics are two-pass (mean, then the centred second moment) like torch's CPU kernel, so cancellation cannot bite.
// Sized for batches of small graphs (ZINC: ~3 K rows, L2 resident); it is correct for any n but a 2 M-row matrix would
// want the rows split over more CTAs.
#include "common.cuh"

namespace mma {

constexpr int BN_COLS = 32, BN_LANES = 32;

__device__ __forceinline__ float bn_col_reduce(float v, float (*sm)[BN_COLS + 1], int tx, int ty) {
    sm[ty][tx] = v;
    __syncthreads();
    float t = 0.0f;
    if (ty == 0) {
#pragma unroll 8
        for (int k = 0; k < BN_LANES; ++k) t += sm[k][tx];          // ascending lane order
        sm[0][tx] = t;
    }
    __syncthreads();
    t = sm[0][tx];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(BN_COLS * BN_LANES) bn_relu_fwd_kernel(
    const float *__restrict__ x, int64_t ldx, int64_t n, int F, const float *__restrict__ gamma, const float *__restrict__ beta,
    float eps, const float *__restrict__ mean_in, const float *__restrict__ var_in, float *__restrict__ y, int64_t ldy,
    float *__restrict__ save_mean, float *__restrict__ save_rstd, float *running_mean, float *running_var, float momentum) {
    __shared__ float sm[BN_LANES][BN_COLS + 1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = blockIdx.x * BN_COLS + tx;
    const bool live = c < F;
    float mean, rstd;
    if (mean_in) {                                   // given statistics (eval mode)
        mean = live ? mean_in[c] : 0.0f;
        rstd = live ? rsqrtf(var_in[c] + eps) : 0.0f;
    } else {
        float s = 0.0f;
        if (live) for (int64_t r = ty; r < n; r += BN_LANES) s += x[r * ldx + c];
        mean = bn_col_reduce(s, sm, tx, ty) / (float)n;
        float q = 0.0f;
        if (live) for (int64_t r = ty; r < n; r += BN_LANES) { const float d = x[r * ldx + c] - mean; q += d * d; }
        const float m2 = bn_col_reduce(q, sm, tx, ty);
        const float var = m2 / (float)n;             // biased: what normalises the batch
        rstd = 1.0f / sqrtf(var + eps);
        if (live && ty == 0 && running_mean) {       // BatchNorm1d's running statistics (unbiased variance)
            running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mean;
            running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (n > 1 ? m2 / (float)(n - 1) : var);
        }
    }
    if (!live) return;
    if (ty == 0) {
        if (save_mean) save_mean[c] = mean;
        if (save_rstd) save_rstd[c] = rstd;
    }
    const float g = gamma ? gamma[c] : 1.0f, b = beta ? beta[c] : 0.0f;
    for (int64_t r = ty; r < n; r += BN_LANES) {
        const float v = (x[r * ldx + c] - mean) * rstd * g + b;
        y[r * ldy + c] = v > 0.0f ? v : 0.0f;
    }
}

// dz = dy * (y > 0); batch statistics: dx = gamma rstd (dz - mean(dz) - xhat mean(dz xhat)); given statistics:
// dx = gamma rstd dz.  dgamma = sum dz xhat, dbeta = sum dz.
__global__ void __launch_bounds__(BN_COLS * BN_LANES) bn_relu_bwd_kernel(
    const float *__restrict__ x, int64_t ldx, const float *__restrict__ y, int64_t ldy, const float *__restrict__ dy,
    int64_t lddy, int64_t n, int F, const float *__restrict__ gamma, const float *__restrict__ save_mean,
    const float *__restrict__ save_rstd, int batch_stats, float *__restrict__ dx, int64_t lddx, float *__restrict__ dgamma,
    float *__restrict__ dbeta) {
    __shared__ float sm[BN_LANES][BN_COLS + 1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = blockIdx.x * BN_COLS + tx;
    const bool live = c < F;
    const float mean = live ? save_mean[c] : 0.0f, rstd = live ? save_rstd[c] : 0.0f;
    float s1 = 0.0f, s2 = 0.0f;
    if (live) {
        for (int64_t r = ty; r < n; r += BN_LANES) {
            const float dz = y[r * ldy + c] > 0.0f ? dy[r * lddy + c] : 0.0f;
            s1 += dz;
            s2 += dz * ((x[r * ldx + c] - mean) * rstd);
        }
    }
    const float sum_dz = bn_col_reduce(s1, sm, tx, ty);
    const float sum_dzx = bn_col_reduce(s2, sm, tx, ty);
    if (!live) return;
    if (ty == 0) {
        if (dbeta) dbeta[c] = sum_dz;
        if (dgamma) dgamma[c] = sum_dzx;
    }
    if (!dx) return;
    const float g = (gamma ? gamma[c] : 1.0f) * rstd;
    const float m1 = batch_stats ? sum_dz / (float)n : 0.0f, m2 = batch_stats ? sum_dzx / (float)n : 0.0f;
    for (int64_t r = ty; r < n; r += BN_LANES) {
        const float dz = y[r * ldy + c] > 0.0f ? dy[r * lddy + c] : 0.0f;
        const float xh = (x[r * ldx + c] - mean) * rstd;
        dx[r * lddx + c] = g * (dz - m1 - xh * m2);
    }
}

}  // namespace mma

using namespace mma;

extern "C" int mma_bn_relu_fwd(const float *x, int64_t ldx, int64_t n, int F, const float *gamma, const float *beta,
                               float eps, const float *mean_in, const float *var_in, float *y, int64_t ldy,
                               float *save_mean, float *save_rstd, float *running_mean, float *running_var,
                               float momentum, mma_stream_t stream) {
    if (!x || !y || n < 0 || F < 1 || (mean_in == nullptr) != (var_in == nullptr) ||
        (running_mean == nullptr) != (running_var == nullptr))
        return MMA_ERR_INVALID;
    if (n == 0) return MMA_OK;
    const dim3 block(BN_COLS, BN_LANES);
    bn_relu_fwd_kernel<<<(unsigned)((F + BN_COLS - 1) / BN_COLS), block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        x, ldx, n, F, gamma, beta, eps, mean_in, var_in, y, ldy, save_mean, save_rstd, running_mean, running_var, momentum);
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}

extern "C" int mma_bn_relu_bwd(const float *x, int64_t ldx, const float *y, int64_t ldy, const float *dy, int64_t lddy,
                               int64_t n, int F, const float *gamma, const float *save_mean, const float *save_rstd,
                               int batch_stats, float *dx, int64_t lddx, float *dgamma, float *dbeta, mma_stream_t stream) {
    if (!x || !y || !dy || !save_mean || !save_rstd || n < 0 || F < 1) return MMA_ERR_INVALID;
    if (n == 0) return MMA_OK;
    const dim3 block(BN_COLS, BN_LANES);
    bn_relu_bwd_kernel<<<(unsigned)((F + BN_COLS - 1) / BN_COLS), block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        x, ldx, y, ldy, dy, lddy, n, F, gamma, save_mean, save_rstd, batch_stats, dx, lddx, dgamma, dbeta);
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}
