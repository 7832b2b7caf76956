// Cross-GPU plumbing of the destination-range sharded layer over NVLink 5 / NVSwitch PEER MEMORY
// (the reference has no distributed code at all).
//
// The bytes of the two exchanges (forward: every rank's Q rows -> every peer; backward: every rank's partial dQ
// slice -> its owner) are moved by the COPY ENGINES with plain peer-to-peer copies issued by the host side
// (mma_b200/peer.py); nothing here touches the payload except the fixed-order slice sum.  What the copy engines
// cannot do is tell the consumer that a slice has landed, so this file provides the device-side handshake:
//
//   mma_peer_epoch_advance   epoch += 1; vals[p] = epoch * 16 + p   (once per layer call; device memory, so a step
//                                                   captured in a CUDA graph hands out new values every replay)
//   mma_peer_copy / _2d      cudaMemcpyAsync / cudaMemcpy2DAsync between peer-mapped pointers: the payload, and --
//                                                   as an 8-byte copy of vals[p] into the peer's flag enqueued right
//                                                   after it on the same stream -- its announcement.  Neither needs an
//                                                   SM, so they run under the persistent one-CTA-per-SM kernels
//   mma_peer_wait            spins until every flag[t] >= epoch * 16 + phase       (acquire, system scope), with a
//                                                   wall-clock bound: a lost peer sets *err instead of hanging
//                                                   the GPU
//   mma_sum_slices           out[r, col0 + c] = sum_k slice_k[r, c] in ascending k: the owner's deterministic
//                                                   reduction of the partial dQ slices (no atomics)
#include "common.cuh"
#include <cstring>

namespace mma {

constexpr int kMaxPeers = 16;
struct SliceList { const float *p[kMaxPeers]; };

__global__ void peer_epoch_advance_kernel(unsigned long long *epoch, unsigned long long *vals) {
    const unsigned long long e = *epoch + 1ull;
    __syncthreads();
    if (threadIdx.x == 0) *epoch = e;
    if (vals && threadIdx.x < 16) vals[threadIdx.x] = e * 16ull + threadIdx.x;
}

__global__ void peer_wait_kernel(const unsigned long long *epoch, const unsigned long long *flags, int n, unsigned phase,
                                 unsigned long long timeout_ns, int *err) {
    const int t = threadIdx.x;
    if (t < n) {
        const unsigned long long want = *epoch * 16ull + phase;
        unsigned long long t0 = 0, now = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + t) : "memory");
            if (v >= want) break;
            __nanosleep(200);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > timeout_ns) { if (err) *err = 1 + t; break; }
        }
    }
    __syncthreads();
    __threadfence_system();
}

// out = sum of the slices in ascending order.  Slices are contiguous [rows, w]; a thread owns UNR 16-byte pieces of
// the output (a block-stride apart, so every load instruction of the warp covers whole lines) and issues the loads of
// ALL slices for all of them before the first add: n_slices x UNR x 16 B in flight per thread (the first version had
// one piece per thread and a load -> add chain: 2.6 TB/s at 2 slices, 0.6 ms of a 14 ms step).
template <int NS>
__global__ void __launch_bounds__(256) sum_slices_kernel(SliceList sl, int n_slices, int64_t rows, int w,
                                                         float *__restrict__ out, int64_t ldo) {
    constexpr int UNR = NS <= 2 ? 4 : 2;
    const int per_row = w >> 2;
    const int64_t total = rows * per_row;
    const int ns = NS > 0 ? NS : n_slices;
    for (int64_t i0 = (int64_t)blockIdx.x * (256 * UNR) + threadIdx.x; i0 < total; i0 += (int64_t)gridDim.x * (256 * UNR)) {
        if constexpr (NS > 0) {
            float4 v[UNR][NS];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int64_t i = i0 + (int64_t)u * 256;
                if (i < total) {
                    const int64_t r = i / per_row;
                    const int64_t off = r * w + (i - r * per_row) * 4;
#pragma unroll
                    for (int k = 0; k < NS; ++k) v[u][k] = __ldcs(reinterpret_cast<const float4 *>(sl.p[k] + off));
                }
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int64_t i = i0 + (int64_t)u * 256;
                if (i < total) {
                    float4 acc = v[u][0];
#pragma unroll
                    for (int k = 1; k < NS; ++k) { acc.x += v[u][k].x; acc.y += v[u][k].y; acc.z += v[u][k].z; acc.w += v[u][k].w; }
                    const int64_t r = i / per_row;
                    *reinterpret_cast<float4 *>(out + r * ldo + (i - r * per_row) * 4) = acc;
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int64_t i = i0 + (int64_t)u * 256;
                if (i >= total) continue;
                const int64_t r = i / per_row;
                const int64_t off = r * w + (i - r * per_row) * 4;
                float4 acc = __ldcs(reinterpret_cast<const float4 *>(sl.p[0] + off));
                for (int k = 1; k < ns; ++k) {
                    const float4 t = __ldcs(reinterpret_cast<const float4 *>(sl.p[k] + off));
                    acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
                }
                *reinterpret_cast<float4 *>(out + r * ldo + (i - r * per_row) * 4) = acc;
            }
        }
    }
}

}  // namespace mma

using namespace mma;

extern "C" int mma_peer_epoch_advance(uint64_t *epoch, uint64_t *vals, mma_stream_t stream) {
    if (!epoch) return MMA_ERR_INVALID;
    peer_epoch_advance_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<unsigned long long *>(epoch), reinterpret_cast<unsigned long long *>(vals));
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}

// The exchange buffers are the one thing this library allocates itself: a CUDA IPC handle names a whole cudaMalloc
// allocation, so the buffer a peer process maps must BE one (a slice of a framework allocator's segment would drag
// the rest of the segment along).  The importer opens the handle on ITS OWN device: the mapping then is peer memory
// of that device, reachable by its kernels and its copy engines over NVLink.  (Opening it on the exporter's device
// index instead -- what torch's tensor sharing does -- gives a mapping the importer's device cannot touch: measured
// 37 GB/s through host staging for copies and an illegal address for kernel stores, against 766 GB/s.)
extern "C" int mma_peer_alloc(size_t bytes, void **ptr, void *handle64) {
    if (!ptr || !handle64 || bytes == 0) return MMA_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    void *p = nullptr;
    MMA_CUDA_CHECK(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handle64), p);
    if (e != cudaSuccess) { (void)cudaFree(p); ::mma::set_last_error(e); return MMA_ERR_CUDA; }
    *ptr = p;
    return MMA_OK;
}
extern "C" int mma_peer_free(void *ptr) {
    if (!ptr) return MMA_OK;
    MMA_CUDA_CHECK(cudaFree(ptr));
    return MMA_OK;
}
extern "C" int mma_peer_open(const void *handle64, void **ptr) {
    if (!handle64 || !ptr) return MMA_ERR_INVALID;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    MMA_CUDA_CHECK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return MMA_OK;
}
extern "C" int mma_peer_close(void *ptr) {
    if (!ptr) return MMA_OK;
    MMA_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
    return MMA_OK;
}

extern "C" int mma_peer_enable_access(int peer_device) {
    int cur = -1;
    MMA_CUDA_CHECK(cudaGetDevice(&cur));
    if (peer_device == cur) return MMA_OK;
    int can = 0;
    MMA_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, cur, peer_device));
    if (!can) return MMA_ERR_UNSUPPORTED;
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); return MMA_OK; }
    MMA_CUDA_CHECK(e);
    return MMA_OK;
}

extern "C" int mma_peer_copy(void *dst, const void *src, size_t bytes, mma_stream_t stream) {
    if (!dst || !src) return MMA_ERR_INVALID;
    if (bytes == 0) return MMA_OK;
    MMA_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, reinterpret_cast<cudaStream_t>(stream)));
    return MMA_OK;
}

extern "C" int mma_peer_copy_2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes,
                                size_t height, mma_stream_t stream) {
    if (!dst || !src || dpitch < width_bytes || spitch < width_bytes) return MMA_ERR_INVALID;
    if (width_bytes == 0 || height == 0) return MMA_OK;
    MMA_CUDA_CHECK(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, height, cudaMemcpyDefault,
                                     reinterpret_cast<cudaStream_t>(stream)));
    return MMA_OK;
}

extern "C" int mma_peer_wait(const uint64_t *epoch, const uint64_t *flags, int n_flags, int phase, uint64_t timeout_ns,
                             int32_t *err, mma_stream_t stream) {
    if (!epoch || !flags || n_flags < 0 || phase < 0 || phase > 15) return MMA_ERR_INVALID;
    if (n_flags > 32) return MMA_ERR_UNSUPPORTED;
    if (n_flags == 0) return MMA_OK;
    peer_wait_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const unsigned long long *>(epoch), reinterpret_cast<const unsigned long long *>(flags), n_flags,
        (unsigned)phase, (unsigned long long)timeout_ns, err);
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}

extern "C" int mma_sum_slices(const float *const *slices_host, int n_slices, int64_t rows, int w, float *out, int64_t ldo,
                              mma_stream_t stream) {
    if (!slices_host || !out || n_slices < 1 || rows < 0 || w < 1) return MMA_ERR_INVALID;
    if (n_slices > kMaxPeers || (w % 4) != 0 || (ldo % 4) != 0 || !aligned16(out)) return MMA_ERR_UNSUPPORTED;
    if (rows == 0) return MMA_OK;
    SliceList sl{};
    for (int i = 0; i < n_slices; ++i) {
        if (!slices_host[i] || !aligned16(slices_host[i])) return MMA_ERR_INVALID;
        sl.p[i] = slices_host[i];
    }
    const int64_t pieces = rows * (w / 4);
    const int64_t want = (pieces + 511) / 512;
    const unsigned grid = (unsigned)(want < (int64_t)kSMs * 16 ? want : (int64_t)kSMs * 16);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (n_slices) {
        case 2: sum_slices_kernel<2><<<grid, 256, 0, st>>>(sl, n_slices, rows, w, out, ldo); break;
        case 4: sum_slices_kernel<4><<<grid, 256, 0, st>>>(sl, n_slices, rows, w, out, ldo); break;
        case 8: sum_slices_kernel<8><<<grid, 256, 0, st>>>(sl, n_slices, rows, w, out, ldo); break;
        default: sum_slices_kernel<0><<<grid, 256, 0, st>>>(sl, n_slices, rows, w, out, ldo); break;
    }
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}
