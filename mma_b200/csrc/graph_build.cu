// Graph preprocessing on the device: STABLE sort of the edge list by a key -> CSR.
//
// Replaces the implicit index plumbing of the reference: PyG's propagate / torch_scatter
// take an unsorted `index` ([E] destination ids, graph_regression/mma_conv.py:130,166) and
// scatter with atomics; node_classification/utils.py:98-100 builds Python neighbour lists.
// A least-significant-digit radix sort (cub::DeviceRadixSort, stable) of (key, edge id)
// keeps, inside every row, the ORIGINAL edge order, which is what makes the sequential
// "first occurrence wins" tie-break of torch_scatter's CPU kernels reproducible.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace mma {

static inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

__global__ void prep_keys_kernel(const int64_t *key, int64_t E, int32_t *keys32, int32_t *iota) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < E) { keys32[i] = (int32_t)key[i]; iota[i] = (int32_t)i; }
}

// rowptr[k] = first slot whose sorted key is >= k   (k = 0 .. n_keys)
__global__ void rowptr_kernel(const int32_t *sorted, int64_t E, int64_t n_keys, int32_t *rowptr) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > n_keys) return;
    int64_t lo = 0, hi = E;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)sorted[mid] < k) lo = mid + 1; else hi = mid;
    }
    rowptr[k] = (int32_t)lo;
}

__global__ void gather_col_kernel(const int64_t *other, const int32_t *perm, int64_t E, int32_t *col) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < E) col[i] = (int32_t)other[perm[i]];
}

__global__ void invert_perm_kernel(const int32_t *perm, int64_t E, int32_t *inv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < E) inv[perm[i]] = (int32_t)i;
}

static int key_bits(int64_t n_keys) {
    int b = 1;
    while (b < 31 && ((int64_t)1 << b) < n_keys) ++b;
    return b;
}

static int cub_temp_bytes(int64_t E, int64_t n_keys, size_t *bytes) {
    size_t t = 0;
    MMA_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, t, (const int32_t *)nullptr, (int32_t *)nullptr,
                                                   (const int32_t *)nullptr, (int32_t *)nullptr, (int)E, 0,
                                                   key_bits(n_keys)));
    *bytes = t;
    return MMA_OK;
}

}  // namespace mma

using namespace mma;

extern "C" int mma_csr_build_workspace_bytes(int64_t E, int64_t n_keys, size_t *bytes) {
    if (!bytes || E < 0 || n_keys < 0) return MMA_ERR_INVALID;
    if (E >= INT32_MAX || n_keys >= INT32_MAX) return MMA_ERR_UNSUPPORTED;
    size_t t = 0;
    int rc = cub_temp_bytes(E > 0 ? E : 1, n_keys, &t);
    if (rc != MMA_OK) return rc;
    *bytes = 3 * align_up(sizeof(int32_t) * (size_t)(E > 0 ? E : 1)) + align_up(t);
    return MMA_OK;
}

extern "C" int mma_csr_build(const int64_t *key, const int64_t *other, int64_t E, int64_t n_keys,
                             int32_t *rowptr, int32_t *col, int32_t *perm, void *workspace,
                             size_t workspace_bytes, mma_stream_t stream) {
    if (!rowptr || E < 0 || n_keys < 0 || (E > 0 && (!key || !perm))) return MMA_ERR_INVALID;
    if (E >= INT32_MAX || n_keys >= INT32_MAX) return MMA_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int block = 256;
    if (E == 0) {
        MMA_CUDA_CHECK(cudaMemsetAsync(rowptr, 0, sizeof(int32_t) * (size_t)(n_keys + 1), st));
        return MMA_OK;
    }
    size_t need = 0, temp = 0;
    int rc = mma_csr_build_workspace_bytes(E, n_keys, &need);
    if (rc != MMA_OK) return rc;
    if (!workspace || workspace_bytes < need) return MMA_ERR_WORKSPACE;
    rc = cub_temp_bytes(E, n_keys, &temp);
    if (rc != MMA_OK) return rc;
    char *w = static_cast<char *>(workspace);
    const size_t seg = align_up(sizeof(int32_t) * (size_t)E);
    int32_t *keys_in = reinterpret_cast<int32_t *>(w);
    int32_t *keys_out = reinterpret_cast<int32_t *>(w + seg);
    int32_t *vals_in = reinterpret_cast<int32_t *>(w + 2 * seg);
    void *cub_tmp = w + 3 * seg;
    const unsigned grid_e = (unsigned)((E + block - 1) / block);
    prep_keys_kernel<<<grid_e, block, 0, st>>>(key, E, keys_in, vals_in);
    MMA_LAUNCH_CHECK();
    MMA_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(cub_tmp, temp, keys_in, keys_out, vals_in, perm, (int)E, 0,
                                                   key_bits(n_keys), st));
    rowptr_kernel<<<(unsigned)((n_keys + 1 + block - 1) / block), block, 0, st>>>(keys_out, E, n_keys, rowptr);
    MMA_LAUNCH_CHECK();
    if (col && other) {
        gather_col_kernel<<<grid_e, block, 0, st>>>(other, perm, E, col);
        MMA_LAUNCH_CHECK();
    }
    return MMA_OK;
}

extern "C" int mma_invert_perm(const int32_t *perm, int64_t E, int32_t *inv, mma_stream_t stream) {
    if (E < 0 || (E > 0 && (!perm || !inv))) return MMA_ERR_INVALID;
    if (E == 0) return MMA_OK;
    invert_perm_kernel<<<(unsigned)((E + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(perm, E, inv);
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}
