// Dev tool: cycles per tcgen05.mma instruction on sm_100a, by shape and operand source.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I mma_b200/csrc scripts/mma_probe.cu -o scripts/mma_probe
//   ./scripts/mma_probe            (one CTA per SM, every CTA runs every variant; prints cycles / instruction)
//
//   (-DPROBE_LANE0: the issuing thread is chosen by `lane == 0` instead of elect.sync)
//
// The GEMMs of the layer retire one tcgen05.mma (M = 128, N = 128, K = 8, kind::tf32) every ~115 cycles.  Is that the
// tensor core?  The probe issues R back-to-back MMAs from one thread on operand tiles that stay in shared / tensor
// memory (contents irrelevant), commits to an mbarrier and reads clock64 around the whole chain.  Answer (B200,
// profiles/r2K_mma_probe_*.log): no -- 64 cycles for N = 128 and 128 for N = 256, from shared or tensor memory, K- or
// MN-major, in any mix; the kernels lose the rest to shared-memory and L2 bandwidth.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"

using namespace mma::tc;

constexpr int R = 1536;                 // instructions per measurement
constexpr int SMEM_BYTES = 176 * 1024;  // A region [0, 32 KB), B region [32 KB, 128 KB), LSU traffic of the contention test [128 KB, 160 KB)

__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}

// One measurement = one template instantiation: descriptors are built before the clock starts and the issue loop is
// fully unrolled over the 4 K steps of a 32-wide K block (setp + tcgen05.mma per instruction, as in the real kernels),
// so the chain is limited by the tensor core and not by the issuing thread's address arithmetic.
//   KIND 0: tf32 SS K-major   1: tf32 TS (A in tensor memory)   2: tf32 SS MN-major (wgrad layout)   3: bf16 SS K-major
//   N0, N1, N2: N of the up to three instructions of one K step (0 = unused); their accumulators lie side by side, a
//   128-wide instruction after a 256-wide one lands on its upper half (the N256 pattern of the streaming kernel)
template <int KIND, int N0, int N1, int N2>
__device__ __forceinline__ long long run_chain(uint32_t tmem, uint32_t pa, uint32_t pb, uint32_t bar, uint32_t phase, int *n_issued) {
    constexpr int NS[3] = {N0, N1, N2};
    constexpr int CNT = (N0 ? 1 : 0) + (N1 ? 1 : 0) + (N2 ? 1 : 0);
    uint64_t ad[4], bd[4][3];
    uint32_t at[4][3], dd[3];
    uint32_t dcol = 0;
#pragma unroll
    for (int j = 0; j < CNT; ++j) {
        dd[j] = tmem + ((dcol + NS[j] > 256) ? 128u : dcol);
        dcol += NS[j];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ad[k] = KIND == 2 ? umma_desc(pa + k * 1024, 4096, 512, 1) : umma_desc_sw128(pa + k * 32, 16, 1024);
#pragma unroll
        for (int j = 0; j < CNT; ++j) {
            bd[k][j] = KIND == 2 ? umma_desc(pb + j * 32768 + k * 1024, 4096, 512, 1)
                                 : umma_desc_sw128(pb + j * 16384 + k * 32, 16, 1024);
            at[k][j] = tmem + 384u + k * 8 + (j & 1) * 32;
        }
    }
    constexpr int PER_ITER = 4 * CNT;
    constexpr int ITERS = R / PER_ITER;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int j = 0; j < CNT; ++j) {
                if (KIND == 0 || KIND == 2) umma_tf32_ss(dd[j], ad[k], bd[k][j], umma_idesc_tf32(128, NS[j], KIND == 2, KIND == 2), 1u);
                else if (KIND == 1) umma_tf32_ts(dd[j], at[k][j], bd[k][j], umma_idesc_tf32(128, NS[j], 0, 0), 1u);
                else umma_bf16_ss(dd[j], ad[k], bd[k][j], idesc_bf16(128, NS[j]), 1u);
            }
        }
    }
    tc_commit(bar);
    mbar_wait(bar, phase);
    const long long t1 = clock64();
    *n_issued = ITERS * PER_ITER;
    return t1 - t0;
}

#define VARIANTS(X)                                                                                 \
    X(0, 64, 0, 0, "tf32 SS  N=64")                                                               \
    X(0, 128, 0, 0, "tf32 SS  N=128")                                                             \
    X(0, 256, 0, 0, "tf32 SS  N=256")                                                             \
    X(0, 128, 128, 0, "tf32 SS  K step = 128 + 128 (two accumulators)")                           \
    X(0, 128, 128, 128, "tf32 SS  K step = 3 x 128")                                              \
    X(0, 256, 128, 0, "tf32 SS  K step = 256 + 128 (N256 pattern of the streaming kernel)")       \
    X(0, 256, 256, 0, "tf32 SS  K step = 256 + 256")                                              \
    X(1, 64, 0, 0, "tf32 TS  N=64")                                                               \
    X(1, 128, 0, 0, "tf32 TS  N=128")                                                             \
    X(1, 256, 0, 0, "tf32 TS  N=256")                                                             \
    X(1, 128, 128, 128, "tf32 TS  K step = 3 x 128")                                              \
    X(1, 256, 128, 0, "tf32 TS  K step = 256 + 128")                                              \
    X(2, 128, 0, 0, "tf32 SS MN-major  N=128 (wgrad layout)")                                     \
    X(2, 256, 0, 0, "tf32 SS MN-major  N=256")                                                    \
    X(2, 128, 128, 128, "tf32 SS MN-major  K step = 3 x 128")                                     \
    X(3, 128, 0, 0, "bf16 SS  N=128 (K=16)")                                                      \
    X(3, 256, 0, 0, "bf16 SS  N=256 (K=16)")

__global__ void __launch_bounds__(128, 1) probe_kernel(long long *cycles, int n_variants) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar_mem;
    __shared__ uint32_t tmem_slot;
    const uint32_t bar = smem_u32(&bar_mem);
    const int warp = threadIdx.x >> 5;
    // zero the operand tiles (denormal / NaN patterns could change the data path's behaviour)
    for (uint32_t i = threadIdx.x; i < (SMEM_BYTES - 1024) / 16; i += blockDim.x)
        asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(base + i * 16), "f"(0.0f) : "memory");
    fence_proxy_async();
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t pa = base, pb = base + 32 * 1024;
#ifdef PROBE_LANE0
    if (warp == 0 && (threadIdx.x & 31) == 0) {     // as the kernels were written before: ptxas wraps every MMA in an ELECT loop
#else
    if (warp == 0 && elect_one()) {
#endif
        uint32_t phase = 0;
        int v = 0, n = 0;
#define X(KIND, N0, N1, N2, NAME)                                                          \
        {                                                                                    \
            const long long c = run_chain<KIND, N0, N1, N2>(tmem, pa, pb, bar, phase, &n);   \
            cycles[(size_t)blockIdx.x * n_variants + v] = c * 1000 / n;                      \
            phase ^= 1u; ++v;                                                                \
        }
        VARIANTS(X)
#undef X
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- the same chains while four other warps (128 threads, like the converter warpgroup of the GEMM kernels) keep the
// shared-memory pipe busy with 128-bit loads (LOAD = 1) or loads + stores (LOAD = 2) on tiles of their own: how the
// tensor core's operand reads and LSU traffic share shared-memory bandwidth.
#define CONTENDED(X)                                                                                \
    X(0, 128, 128, 128, 1, "tf32 SS  3 x 128         | 4 warps LDS.128")                            \
    X(0, 128, 128, 128, 2, "tf32 SS  3 x 128         | 4 warps LDS.128 + STS.128")                  \
    X(0, 256, 128, 0, 1, "tf32 SS  256 + 128       | 4 warps LDS.128")                              \
    X(0, 256, 128, 0, 2, "tf32 SS  256 + 128       | 4 warps LDS.128 + STS.128")                    \
    X(1, 128, 128, 128, 1, "tf32 TS  3 x 128         | 4 warps LDS.128")                            \
    X(1, 128, 128, 128, 2, "tf32 TS  3 x 128         | 4 warps LDS.128 + STS.128")                  \
    X(2, 128, 128, 128, 2, "tf32 SS MN-major 3 x 128 | 4 warps LDS.128 + STS.128")                  \
    X(2, 256, 128, 0, 2, "tf32 SS MN-major 256+128 | 4 warps LDS.128 + STS.128")

__global__ void __launch_bounds__(160, 1) contend_kernel(long long *cycles, long long *lsu_bytes, int n_variants) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar_mem;
    __shared__ uint32_t tmem_slot;
    __shared__ int done;
    __shared__ unsigned long long moved;
    __shared__ long long chain_cycles;
    __shared__ int chain_n;
    const uint32_t bar = smem_u32(&bar_mem);
    const int warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < (SMEM_BYTES - 1024) / 16; i += blockDim.x)
        asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(base + i * 16), "f"(0.0f) : "memory");
    fence_proxy_async();
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t pa = base, pb = base + 32 * 1024;
    const uint32_t rd = base + 128 * 1024, wr = base + 144 * 1024;       // one 16 KB tile each
    const int ct = threadIdx.x - 32;
    uint32_t phase = 0;
    int v = 0;
    float sink = 0.f;
#define X(KIND, N0, N1, N2, LOAD, NAME)                                                                     \
    {                                                                                                        \
        if (threadIdx.x == 0) { done = 0; moved = 0ull; }                                                    \
        __syncthreads();                                                                                     \
        if (warp == 0) {                                                                                     \
            if (elect_one()) {                                                                               \
                int n = 0;                                                                                   \
                chain_cycles = run_chain<KIND, N0, N1, N2>(tmem, pa, pb, bar, phase, &n);                    \
                chain_n = n;                                                                                 \
                *(volatile int *)&done = 1;                                                                  \
            }                                                                                                \
        } else {                                                                                             \
            unsigned long long it = 0;                                                                       \
            while (*(volatile int *)&done == 0) {                                                            \
                float4 x[8];                                                                                 \
                _Pragma("unroll") for (int i = 0; i < 8; ++i)                                                \
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"                                  \
                                 : "=f"(x[i].x), "=f"(x[i].y), "=f"(x[i].z), "=f"(x[i].w)                    \
                                 : "r"(rd + (i * 128 + ct) * 16));                                           \
                if (LOAD == 2) {                                                                             \
                    _Pragma("unroll") for (int i = 0; i < 8; ++i)                                            \
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(wr + (i * 128 + ct) * 16), \
                                     "f"(x[i].x), "f"(x[i].y), "f"(x[i].z), "f"(x[i].w) : "memory");         \
                } else {                                                                                     \
                    _Pragma("unroll") for (int i = 0; i < 8; ++i) sink += x[i].x;                            \
                }                                                                                            \
                ++it;                                                                                        \
            }                                                                                                \
            atomicAdd(&moved, it * (LOAD == 2 ? 256ull : 128ull));                                           \
        }                                                                                                    \
        __syncthreads();                                                                                     \
        if (threadIdx.x == 0) {                                                                              \
            cycles[(size_t)blockIdx.x * n_variants + v] = chain_cycles * 1000 / chain_n;                     \
            lsu_bytes[(size_t)blockIdx.x * n_variants + v] = (long long)(moved * 1000ull / (unsigned long long)chain_cycles); \
        }                                                                                                    \
        phase ^= 1u; ++v;                                                                                    \
    }
    CONTENDED(X)
#undef X
    if (sink == 12345.678f) cycles[0] = 0;        // keeps the loads of LOAD = 1 alive
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    std::vector<const char *> names;
#define X(KIND, N0, N1, N2, NAME) names.push_back(NAME);
    VARIANTS(X)
#undef X
    const int nv = (int)names.size();
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    long long *dc;
    cudaMalloc(&dc, (size_t)sms * nv * sizeof(long long));
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    for (int rep = 0; rep < 2; ++rep) {
        probe_kernel<<<sms, 128, SMEM_BYTES>>>(dc, nv);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    }
    std::vector<long long> hc((size_t)sms * nv);
    cudaMemcpy(hc.data(), dc, hc.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    printf("tcgen05.mma issue probe: %d SMs, max clock %d MHz, ~%d instructions per chain, M = 128\n", sms, khz / 1000, R);
    printf("%-72s %10s %10s %10s\n", "variant", "cyc/instr", "min", "max");
    for (int v = 0; v < nv; ++v) {
        double sum = 0; long long mn = 1ll << 60, mx = 0;
        for (int b = 0; b < sms; ++b) {
            const long long c = hc[(size_t)b * nv + v];
            sum += c; if (c < mn) mn = c; if (c > mx) mx = c;
        }
        printf("%-72s %10.1f %10.1f %10.1f\n", names[v], sum / sms / 1000.0, mn / 1000.0, mx / 1000.0);
    }
    // ---- under shared-memory contention
    std::vector<const char *> cn;
#define X(KIND, N0, N1, N2, LOAD, NAME) cn.push_back(NAME);
    CONTENDED(X)
#undef X
    const int nc = (int)cn.size();
    long long *dcc, *dlb;
    cudaMalloc(&dcc, (size_t)sms * nc * sizeof(long long));
    cudaMalloc(&dlb, (size_t)sms * nc * sizeof(long long));
    cudaFuncSetAttribute(contend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    for (int rep = 0; rep < 2; ++rep) {
        contend_kernel<<<sms, 160, SMEM_BYTES>>>(dcc, dlb, nc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error (contention test): %s\n", cudaGetErrorString(e)); return 1; }
    }
    std::vector<long long> hcc((size_t)sms * nc), hlb((size_t)sms * nc);
    cudaMemcpy(hcc.data(), dcc, hcc.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(hlb.data(), dlb, hlb.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    printf("\nwith LSU traffic beside the chain (operand bytes per instruction: SS N=128 8 KB, SS N=256 12 KB, TS N=128 4 KB):\n");
    printf("%-72s %10s %14s\n", "variant", "cyc/instr", "LSU B/cycle");
    for (int v = 0; v < nc; ++v) {
        double sc = 0, sb = 0;
        for (int b = 0; b < sms; ++b) { sc += hcc[(size_t)b * nc + v]; sb += hlb[(size_t)b * nc + v]; }
        printf("%-72s %10.1f %14.1f\n", cn[v], sc / sms / 1000.0, sb / sms / 1000.0);
    }
    return 0;
}
