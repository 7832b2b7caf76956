// G1-G4: the dense projections of the MultiMaskConv layer on the 5th-gen tensor cores.
//
// Replaces (reference paths relative to /root/reference/graph_regression):
//   mask_aggr.py:68        the mask projection Linear          -> P, Q = X W_i^T + b, X W_j^T
//   mma_conv.py:132-136    cat([x, out]) -> post_nns -> lin    -> grouped GEMM over degree ranges + lin
// and their autograd backward (dgrad: same kernel with the transposed weight; wgrad: gemm_wgrad below).
//
// The reference computes these in fp32 and parity is judged at 1e-5, so plain TF32 (10-bit mantissa)
// is not acceptable.  Every fp32 operand is split as x = hi + lo with hi exactly representable in
// TF32 and lo = x - hi; D += A_lo B_hi + A_hi B_lo + A_hi B_hi (3xTF32) recovers the fp32 product to
// ~2^-21 with fp32 accumulation in TMEM.  Weights (small) are pre-split on the host side of the
// call; activations are split INSIDE the kernel by a converter warpgroup working on the TMA-filled
// shared-memory tile, so no extra pass over HBM is needed.
//
// One persistent CTA per SM, warp-specialised (linear: 352 threads; wgrad: 320):
//   warp 0      TMA producer     raw activation tiles [128 x 32] fp32 (from HBM) into a deep ring
//   warp 10     TMA producer     (linear only) pre-split weight tiles B_hi/B_lo [128 x 32] (from L2), shallow ring
//   warp 1      MMA issuer       one thread issues 12 tcgen05.mma.kind::tf32 (M=128,N=128,K=8) per 32-wide K block
//   warps 2-5   converter        A -> A_lo tile (and, in mode 0, the rounded hi in place), fence.proxy.async
//   warps 6-9   epilogue         TMEM -> registers -> 256-bit global stores (a thread owns one row of the tile:
//                                32 consecutive columns = one full 128-byte line), with optional bias,
//                                row-indexed addend and output-row scatter
// Pipelines: raw / lo / weight rings with full+empty mbarriers each, and TMEM full/empty (2 accumulators of
// 128 + 128 columns), so the epilogue of tile i overlaps the main loop of tile i+1.
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace mma {
namespace gemm {

using namespace tc;

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int A_BYTES = BM * BK * 4;                      // 16 KB
constexpr int B_BYTES = BN * BK * 4;                      // 16 KB
constexpr int EPI_BYTES = 4 * 32 * 32 * 4;                // one 32x32 fp32 transpose buffer per epilogue warp
constexpr int BAR_BYTES = 256;
// The operands that come from HBM need ~90 KB in flight per SM to cover the DRAM latency at full
// bandwidth, so the RAW tiles get a deep ring of their own; the split (lo) tiles live only from the
// converter to the MMA and the weights come from L2, so their rings are shallow.
//   linear: raw A ring (hi written in place) | A_lo ring | B ring (B_hi + B_lo per slot)
constexpr int NT_NA = 5, NT_NL = 2, NT_NB = 3;
constexpr int NT_SMEM_BYTES = (NT_NA + NT_NL) * A_BYTES + NT_NB * 2 * B_BYTES + EPI_BYTES + BAR_BYTES + 1024;
constexpr int NT_THREADS = 352;                           // + 1 warp: the B (weight) producer
//   wgrad: raw [G | A | A_lo] ring | G_lo ring.  A_lo sits right behind the raw A tile of its stage so that ONE MMA of
//   N = 256 runs over [A_hi ; A_lo] (see the kernel)
constexpr int WG_NR = 4, WG_NL = 2;
constexpr int WG_SLOT_BYTES = A_BYTES + 2 * B_BYTES;     // 48 KB
constexpr int WG_SMEM_BYTES = WG_NR * WG_SLOT_BYTES + WG_NL * A_BYTES + BAR_BYTES + 1024;
static_assert(NT_SMEM_BYTES <= 232448 && WG_SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory per CTA");
constexpr int THREADS = 320;
constexpr int TMEM_COLS = 512;   // 2 tiles in flight x (main + correction) accumulators of 128 columns

struct alignas(64) NtParams {
    CUtensorMap map_a0, map_a1, map_bhi, map_blo;
    CUtensorMap map_c;          // output as [32 rows x 32 columns] boxes (resident-A variant, plain row-major output)
    int tma_store;              // 1: full tiles leave through shared memory + TMA tile stores
    int kb_split, num_kb;       // k-blocks [0,kb_split) read map_a0, [kb_split,num_kb) read map_a1
    int n_tiles_n;
    int mode;                   // 0: 3xTF32, hi written back explicitly; 1: 3xTF32, raw A as hi; 2: 1xTF32
    int64_t M, n_tiles_m;
    int N;
    const int32_t *tile_tab;    // [n_tiles_m][4] = row0, row_end, b_row_off, -   (null: plain GEMM)
    float *C;
    int64_t ldc;
    const int32_t *out_map;     // output row of tile row r (null: r)
    const float *bias;          // [N]
    const float *add;           // indexed like C (output rows) or, add_in != 0, by the input row
    int64_t ldadd;
    int add_in;
    int relu;                   // epilogue: C = max(C, 0) after bias / addend (MMA_GEMM_RELU)
    int n256;                   // streaming kernel: A_hi B_hi and A_hi B_lo as one N = 256 MMA (MMA_GEMM_N256)
};

struct Tile { int64_t row0, row_end; int b_off, n0; };

__device__ __forceinline__ Tile locate_tile(const NtParams &p, int64_t t) {
    Tile tl;
    const int64_t m = t / p.n_tiles_n;
    tl.n0 = (int)(t - m * p.n_tiles_n) * BN;
    if (p.tile_tab) {
        const int4 e = __ldg(reinterpret_cast<const int4 *>(p.tile_tab) + m);
        tl.row0 = e.x; tl.row_end = e.y; tl.b_off = e.z;
    } else {
        tl.row0 = m * BM; tl.row_end = p.M; tl.b_off = 0;
    }
    return tl;
}

// main + correction accumulator chunk (32 lanes x 32 columns) -> registers
__device__ __forceinline__ void load_acc(uint32_t taddr, bool with_corr, uint32_t (&r)[32]) {
    tmem_ld_32x32(taddr, r);
    if (with_corr) {
        uint32_t c[32];
        tmem_ld_32x32(taddr + BN, c);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(c[i]));
    } else {
        tmem_ld_wait();
    }
}

// One 32-row x 32-column accumulator chunk held row-per-thread (thread = row) -> global memory.
// A thread owns 32 consecutive columns of its row = one full 128-byte line, so it stores straight from
// registers (8 x 16 B; consecutive instructions complete each 32-byte sector): no shared-memory
// transpose, a quarter of the instructions of the staged version, which had made the epilogue warps
// the slowest stage of the pipeline.  `stg` is unused (kept for the call sites).
__device__ __forceinline__ void store_chunk(float *stg, const uint32_t (&r)[32], int lane, int64_t row0, int64_t row_end,
                                            int col0, int n_cols, float *C, int64_t ldc, const int32_t *out_map,
                                            const float *bias, const float *add, int64_t ldadd, int add_in) {
    (void)stg;
    const bool relu = (add_in & 2) != 0;            // bit 1 of the flag word: ReLU after bias / addend
    add_in &= 1;
    const int64_t grow = row0 + lane;
    if (grow >= row_end || col0 >= n_cols) return;
    const int64_t orow = out_map ? (int64_t)__ldg(out_map + grow) : grow;
    float *crow = C + orow * ldc + col0;
    const float *arow = add ? add + (add_in ? grow : orow) * ldadd + col0 : nullptr;
    const bool wide = ((reinterpret_cast<uintptr_t>(crow) | reinterpret_cast<uintptr_t>(arow)) & 31u) == 0 && col0 + 32 <= n_cols;
    if (wide) {
        // 256-bit accesses: every store instruction writes whole 32-byte sectors
        float av[32];
        if (arow) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=f"(av[8 * j]), "=f"(av[8 * j + 1]), "=f"(av[8 * j + 2]), "=f"(av[8 * j + 3]),
                               "=f"(av[8 * j + 4]), "=f"(av[8 * j + 5]), "=f"(av[8 * j + 6]), "=f"(av[8 * j + 7])
                             : "l"(arow + 8 * j));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[8 * j + i]);
            if (bias) {
                const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bias + col0) + 2 * j);
                const float4 b1 = __ldg(reinterpret_cast<const float4 *>(bias + col0) + 2 * j + 1);
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if (arow) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] += av[8 * j + i];
            }
            if (relu) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.0f);
            }
            asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(crow + 8 * j), "f"(v[0]), "f"(v[1]),
                         "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                         : "memory");
        }
        return;
    }
    float4 av[8];
    if (arow) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (col0 + 4 * j < n_cols) av[j] = __ldcs(reinterpret_cast<const float4 *>(arow) + j);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (col0 + 4 * j < n_cols) {
            float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                   __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            if (bias) {
                const float4 bb = __ldg(reinterpret_cast<const float4 *>(bias + col0) + j);
                v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
            }
            if (arow) { v.x += av[j].x; v.y += av[j].y; v.z += av[j].z; v.w += av[j].w; }
            if (relu) { v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f); }
            __stcs(reinterpret_cast<float4 *>(crow) + j, v);
        }
    }
}

__global__ void __launch_bounds__(NT_THREADS, 1) gemm_nt_kernel(const __grid_constant__ NtParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t lo_base = smem_base + NT_NA * A_BYTES;
    const uint32_t b_base = lo_base + NT_NL * A_BYTES;
    const uint32_t epi_base = b_base + NT_NB * 2 * B_BYTES;
    const uint32_t bar_base = epi_base + EPI_BYTES;
    // barriers: a_full[NA], a_empty[NA], lo_full[NL], lo_empty[NL], b_full[NB], b_empty[NB], tfull[2], tempty[2],
    // then the TMEM base address word
    auto a_full = [&](int i) { return bar_base + 8u * i; };
    auto a_empty = [&](int i) { return bar_base + 8u * (NT_NA + i); };
    auto lo_full = [&](int i) { return bar_base + 8u * (2 * NT_NA + i); };
    auto lo_empty = [&](int i) { return bar_base + 8u * (2 * NT_NA + NT_NL + i); };
    auto b_full = [&](int i) { return bar_base + 8u * (2 * NT_NA + 2 * NT_NL + i); };
    auto b_empty = [&](int i) { return bar_base + 8u * (2 * NT_NA + 2 * NT_NL + NT_NB + i); };
    constexpr int kTBar = 2 * NT_NA + 2 * NT_NL + 2 * NT_NB;
    auto tfull_bar = [&](int b) { return bar_base + 8u * (kTBar + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (kTBar + 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (kTBar + 4);
    static_assert(8 * (kTBar + 4) + 4 <= BAR_BYTES, "barrier block too small");
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = p.n_tiles_m * p.n_tiles_n;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.map_a0); tma_prefetch_desc(&p.map_a1);
        tma_prefetch_desc(&p.map_bhi); tma_prefetch_desc(&p.map_blo);
        for (int i = 0; i < NT_NA; ++i) { mbar_init(a_full(i), 1); mbar_init(a_empty(i), 1); }
        for (int i = 0; i < NT_NL; ++i) { mbar_init(lo_full(i), 4); mbar_init(lo_empty(i), 1); }
        for (int i = 0; i < NT_NB; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===================================================================== TMA producer: activations (HBM)
        if (elect_one()) {
            uint32_t it = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const Tile tl = locate_tile(p, t);
                for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
                    const int sa = it % NT_NA;
                    mbar_wait(a_empty(sa), ((it / NT_NA) & 1u) ^ 1u);
                    mbar_expect_tx(a_full(sa), A_BYTES);
                    if (kb < p.kb_split) tma_load_2d(smem_base + sa * A_BYTES, &p.map_a0, a_full(sa), kb * BK, (int)tl.row0);
                    else tma_load_2d(smem_base + sa * A_BYTES, &p.map_a1, a_full(sa), (kb - p.kb_split) * BK, (int)tl.row0);
                }
            }
        }
    } else if (warp == 10) {
        // ===================================================================== TMA producer: weights (L2)
        if (elect_one()) {
            uint32_t it = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const Tile tl = locate_tile(p, t);
                for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
                    const int sb = it % NT_NB;
                    mbar_wait(b_empty(sb), ((it / NT_NB) & 1u) ^ 1u);
                    const uint32_t dst = b_base + sb * 2 * B_BYTES;
                    mbar_expect_tx(b_full(sb), (p.mode == 2 ? 1 : 2) * B_BYTES);
                    tma_load_2d(dst, &p.map_bhi, b_full(sb), kb * BK, tl.b_off + tl.n0);
                    if (p.mode != 2) tma_load_2d(dst + B_BYTES, &p.map_blo, b_full(sb), kb * BK, tl.b_off + tl.n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_tf32(BM, BN, 0, 0);
            constexpr uint32_t idesc2 = umma_idesc_tf32(BM, 2 * BN, 0, 0);
            uint32_t it = 0, ti = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
                const uint32_t buf = ti & 1u;
                mbar_wait(tempty_bar(buf), ((ti >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 2 * BN;      // main accumulator; correction at +BN
                for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
                    const int sa = it % NT_NA, sl = it % NT_NL, sb = it % NT_NB;
                    mbar_wait(a_full(sa), (it / NT_NA) & 1u);
                    if (p.mode != 2) mbar_wait(lo_full(sl), (it / NT_NL) & 1u);
                    mbar_wait(b_full(sb), (it / NT_NB) & 1u);
                    tc_fence_after();
                    const uint32_t pa = smem_base + sa * A_BYTES, pl = lo_base + sl * A_BYTES, pb = b_base + sb * 2 * B_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        const uint64_t a_hi = umma_desc_sw128(pa + k * 32, 16, 1024);
                        const uint64_t a_lo = umma_desc_sw128(pl + k * 32, 16, 1024);
                        const uint64_t b_hi = umma_desc_sw128(pb + k * 32, 16, 1024);
                        const uint32_t first = (kb | k) != 0;
                        // The tensor core truncates when it adds into the fp32 accumulator, a bias that grows
                        // with the length of the accumulation chain.  The two small cross terms go to their
                        // own accumulator (2^-11 of the magnitude, so their truncation is negligible) and the
                        // main chain is 1 MMA per K step instead of 3; the epilogue adds the two.
                        //
                        // A_hi B_hi (main) and A_hi B_lo (correction) are ONE instruction of N = 256: the weight slot
                        // holds B_hi and B_lo back to back (128 rows of 128 B each, 8-row groups 1024 B apart), so a
                        // descriptor at B_hi with N = 256 runs straight on into B_lo, and the two accumulators are
                        // adjacent in tensor memory -- A_hi crosses shared memory once instead of twice (20 instead
                        // of 24 KB of operand reads per K step), 2 MMAs instead of 3: measured 8 % on the long-K GEMMs.
                        if (p.mode != 2 && p.n256) {
                            umma_tf32_ss(d_tmem, a_hi, b_hi, idesc2, first);             // [main | corr] (+)= A_hi [B_hi ; B_lo]^T
                            umma_tf32_ss(d_tmem + BN, a_lo, b_hi, idesc, 1u);            // corr += A_lo B_hi^T
                        } else if (p.mode != 2) {
                            const uint64_t b_lo = umma_desc_sw128(pb + B_BYTES + k * 32, 16, 1024);
                            umma_tf32_ss(d_tmem + BN, a_lo, b_hi, idesc, first);
                            umma_tf32_ss(d_tmem + BN, a_hi, b_lo, idesc, 1u);
                            umma_tf32_ss(d_tmem, a_hi, b_hi, idesc, first);
                        } else {
                            umma_tf32_ss(d_tmem, a_hi, b_hi, idesc, first);
                        }
                    }
                    tc_commit(a_empty(sa));             // slots reusable once these MMAs retire
                    if (p.mode != 2) tc_commit(lo_empty(sl));
                    tc_commit(b_empty(sb));
                }
                tc_commit(tfull_bar(buf));              // accumulator complete
            }
        }
    } else if (warp < 6) {
        // ===================================================================== converter (A -> hi, lo)
        if (p.mode != 2) {
            const int ct = threadIdx.x - 64;            // 0..127
            uint32_t it = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
                    const int sa = it % NT_NA, sl = it % NT_NL;
                    mbar_wait(a_full(sa), (it / NT_NA) & 1u);
                    mbar_wait(lo_empty(sl), ((it / NT_NL) & 1u) ^ 1u);
                    float4 *a = reinterpret_cast<float4 *>(smem + sa * A_BYTES);
                    float4 *alo = reinterpret_cast<float4 *>(smem + (lo_base - smem_base) + sl * A_BYTES);
                    // all loads first (8 independent LDS.128 in flight), then the split: a serial
                    // load -> split -> store chain made the converter the slowest stage of the pipeline
                    float4 v[A_BYTES / 16 / 128];
#pragma unroll
                    for (int i = 0; i < A_BYTES / 16 / 128; ++i) v[i] = a[i * 128 + ct];
                    if (p.mode == 0) {
#pragma unroll
                        for (int i = 0; i < A_BYTES / 16 / 128; ++i) {
                            const float4 h = make_float4(tf32_rna(v[i].x), tf32_rna(v[i].y), tf32_rna(v[i].z), tf32_rna(v[i].w));
                            alo[i * 128 + ct] = make_float4(tf32_rna(v[i].x - h.x), tf32_rna(v[i].y - h.y),
                                                            tf32_rna(v[i].z - h.z), tf32_rna(v[i].w - h.w));
                            a[i * 128 + ct] = h;
                        }
                    } else {            // the tensor core ignores the low 13 mantissa bits: raw A acts as trunc(A)
#pragma unroll
                        for (int i = 0; i < A_BYTES / 16 / 128; ++i) {
                            const float4 h = make_float4(tf32_hi(v[i].x), tf32_hi(v[i].y), tf32_hi(v[i].z), tf32_hi(v[i].w));
                            alo[i * 128 + ct] = make_float4(v[i].x - h.x, v[i].y - h.y, v[i].z - h.z, v[i].w - h.w);
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(lo_full(sl));
                }
            }
        }
    } else {
        // ===================================================================== epilogue
        const int wq = warp & 3;                        // TMEM lane quarter this warp may access
        float *stg = reinterpret_cast<float *>(smem + (epi_base - smem_base) + (warp - 6) * 4096);
        uint32_t ti = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++ti) {
            const Tile tl = locate_tile(p, t);
            const uint32_t buf = ti & 1u;
            mbar_wait(tfull_bar(buf), (ti >> 1) & 1u);
            __syncwarp();            // reconverge before the .sync.aligned tcgen05.ld
            tc_fence_after();
            const int n_chunks = min(4, (p.N - tl.n0 + 31) / 32);
            for (int c = 0; c < n_chunks; ++c) {
                uint32_t r[32];
                load_acc(tmem_base + buf * 2 * BN + c * 32 + ((uint32_t)(wq * 32) << 16), p.mode != 2, r);
                if (c == n_chunks - 1) {                // accumulator fully read: hand the buffer back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(buf));
                }
                store_chunk(stg, r, lane, tl.row0 + wq * 32, tl.row_end, tl.n0 + c * 32, p.N, p.C, p.ldc, p.out_map,
                            p.bias, p.add, p.ldadd, p.add_in | (p.relu << 1));
                __syncwarp();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------
// Short-K variant (K <= 128, several 128-column tiles per row block): the activation tile stays RESIDENT IN
// TENSOR MEMORY.
//
// The kernel above retires an MMA every ~105-135 cycles where the tensor core needs 64: per tcgen05.mma ~16 KB move
// through shared memory (operand reads, the converter's read + write, the TMA fills -- a third of it weight tiles
// re-streamed from L2), and it repeats the conversion of A for every 128-column tile of the same rows.  For
// the two GEMMs with K = 128 and 384 / 640 output columns (mask projection, dgrad of the post transform) this
// variant
//   * converts a [128 x K] activation tile ONCE per row block: TMA -> shared memory -> registers (a thread owns
//     a row) -> hi / lo written with tcgen05.st into tensor-memory columns [0,128) / [128,256);
//   * issues the MMAs with the A operand read FROM TENSOR MEMORY (tcgen05.mma [d], [a], b_desc): only the
//     weight tiles cross shared memory;
//   * walks the output columns in 128-wide sub-tiles that all reuse the resident A.  K <= 128 means at most 48
//     accumulations per element, so ONE accumulator serves the three 3xTF32 terms (the streaming kernel keeps the
//     cross terms apart because its chains reach 240); the two 128-column accumulators at columns [256,512)
//     belong to one epilogue warpgroup each, so the epilogue of sub-tile j overlaps the MMAs of sub-tile j + 1.
//     (64-column sub-tiles with a separate correction accumulator were measured first: 92 cycles per N = 64 MMA,
//     slower than the streaming kernel; N = 128 MMAs take ~115 cycles.)
// The K-block slots of A are released one by one while the LAST sub-tile of a row block is being issued, so
// the converter refills them for the next row block under the tail of the current one.  Full tiles of a plain
// row-major output leave through swizzled shared-memory staging + TMA tile stores.
//   warp 0 TMA (activations) | warp 1 MMA issuer | warps 2-5 converter | warps 6-9, 10-13 epilogue | warp 14 TMA (weights)
// Measured at M = 2M: K=128 -> 384 columns 1.30 -> 1.00 ms, K=128 -> 640 columns 2.10 -> 1.61 ms.
// ------------------------------------------------------------------------------------------
constexpr int AR_BN = 128;                                 // output columns per sub-tile
constexpr int AR_B_BYTES = AR_BN * BK * 4;                 // 16 KB per half (hi or lo)
constexpr int AR_NA = 4, AR_NB = 4, AR_MAX_KB = 4;
constexpr int AR_BAR_BYTES = 512;
constexpr int AR_STG_BYTES = 8 * 4096;                     // per epilogue warp: one [32 x 32] fp32 staging tile
constexpr int AR_SMEM_BYTES = AR_NA * A_BYTES + AR_NB * 2 * AR_B_BYTES + AR_STG_BYTES + AR_BAR_BYTES + 1024;
constexpr int AR_THREADS = 480;
static_assert(AR_SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory per CTA");

__global__ void __launch_bounds__(AR_THREADS, 1) gemm_nt_ares_kernel(const __grid_constant__ NtParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t b_base = smem_base + AR_NA * A_BYTES;
    const uint32_t stg_base = b_base + AR_NB * 2 * AR_B_BYTES;
    const uint32_t bar_base = stg_base + AR_STG_BYTES;
    auto a_full = [&](int i) { return bar_base + 8u * i; };
    auto a_empty = [&](int i) { return bar_base + 8u * (AR_NA + i); };
    auto atm_full = [&](int i) { return bar_base + 8u * (2 * AR_NA + i); };
    auto atm_empty = [&](int i) { return bar_base + 8u * (2 * AR_NA + AR_MAX_KB + i); };
    auto b_full = [&](int i) { return bar_base + 8u * (2 * AR_NA + 2 * AR_MAX_KB + i); };
    auto b_empty = [&](int i) { return bar_base + 8u * (2 * AR_NA + 2 * AR_MAX_KB + AR_NB + i); };
    constexpr int kTBar = 2 * AR_NA + 2 * AR_MAX_KB + 2 * AR_NB;
    auto tfull_bar = [&](int b) { return bar_base + 8u * (kTBar + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (kTBar + 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (kTBar + 4);
    static_assert(8 * (kTBar + 4) + 4 <= AR_BAR_BYTES, "barrier block too small");
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = p.num_kb;                            // <= AR_MAX_KB
    const int n_sub = (p.N + AR_BN - 1) / AR_BN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.map_a0); tma_prefetch_desc(&p.map_bhi); tma_prefetch_desc(&p.map_blo);
        if (p.tma_store) tma_prefetch_desc(&p.map_c);
        for (int i = 0; i < AR_NA; ++i) { mbar_init(a_full(i), 1); mbar_init(a_empty(i), 4); }
        for (int i = 0; i < AR_MAX_KB; ++i) { mbar_init(atm_full(i), 4); mbar_init(atm_empty(i), 1); }
        for (int i = 0; i < AR_NB; ++i) { mbar_init(b_full(i), 1); mbar_init(b_empty(i), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    constexpr uint32_t A_HI_COL = 0, A_LO_COL = 128, ACC_COL = 256;     // accumulator of buffer b at ACC_COL + 128 b

    if (warp == 0) {
        // ===================================================================== TMA producer: activations (HBM)
        if (elect_one()) {
            uint32_t it = 0;
            for (int64_t t = blockIdx.x; t < p.n_tiles_m; t += gridDim.x) {
                const Tile tl = locate_tile(p, t * p.n_tiles_n);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int sa = it % AR_NA;
                    mbar_wait(a_empty(sa), ((it / AR_NA) & 1u) ^ 1u);
                    mbar_expect_tx(a_full(sa), A_BYTES);
                    tma_load_2d(smem_base + sa * A_BYTES, &p.map_a0, a_full(sa), kb * BK, (int)tl.row0);
                }
            }
        }
    } else if (warp == 14) {
        // ===================================================================== TMA producer: weights (L2)
        if (elect_one()) {
            uint32_t it = 0;
            for (int64_t t = blockIdx.x; t < p.n_tiles_m; t += gridDim.x) {
                const Tile tl = locate_tile(p, t * p.n_tiles_n);
                for (int j = 0; j < n_sub; ++j) {
                    for (int kb = 0; kb < num_kb; ++kb, ++it) {
                        const int sb = it % AR_NB;
                        mbar_wait(b_empty(sb), ((it / AR_NB) & 1u) ^ 1u);
                        const uint32_t dst = b_base + sb * 2 * AR_B_BYTES;
                        mbar_expect_tx(b_full(sb), 2 * AR_B_BYTES);
                        tma_load_2d(dst, &p.map_bhi, b_full(sb), kb * BK, tl.b_off + j * AR_BN);
                        tma_load_2d(dst + AR_B_BYTES, &p.map_blo, b_full(sb), kb * BK, tl.b_off + j * AR_BN);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (elect_one()) {
            constexpr int AR_SPLIT = 1;                      // column slices per sub-tile issued as separate MMAs (see below)
            constexpr int AR_HN = AR_BN / AR_SPLIT;
            constexpr uint32_t idesc = umma_idesc_tf32(BM, AR_HN, 0, 0);
            uint32_t itb = 0, si = 0, mt = 0;
            for (int64_t t = blockIdx.x; t < p.n_tiles_m; t += gridDim.x, ++mt) {
                for (int j = 0; j < n_sub; ++j, ++si) {
                    const uint32_t buf = si & 1u;
                    mbar_wait(tempty_bar(buf), ((si >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t d_main = tmem_base + ACC_COL + buf * 128u;
                    for (int kb = 0; kb < num_kb; ++kb, ++itb) {
                        const int sb = itb % AR_NB;
                        if (j == 0) mbar_wait(atm_full(kb), mt & 1u);
                        mbar_wait(b_full(sb), (itb / AR_NB) & 1u);
                        tc_fence_after();
                        const uint32_t pb = b_base + sb * 2 * AR_B_BYTES;
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k) {
                            const uint32_t a_hi = tmem_base + A_HI_COL + kb * BK + k * 8;
                            const uint32_t a_lo = tmem_base + A_LO_COL + kb * BK + k * 8;
                            const uint32_t first = (kb | k) != 0;
                            // K <= 128: at most 48 accumulations per element, so one accumulator serves all three
                            // terms (the streaming kernel keeps a second one for the cross terms because its chains
                            // reach 240); that leaves room for 128-column sub-tiles next to the resident A
                            // AR_SPLIT = 2 (two alternating N = 64 slices, i.e. independent accumulator chains) was
                            // measured SLOWER (1.00 -> 1.26 ms): the cost is per instruction (~70 cycles + 0.36 per
                            // column), not the dependence between consecutive MMAs, so wide MMAs win.
#pragma unroll
                            for (int term = 0; term < 3; ++term) {
#pragma unroll
                                for (int h = 0; h < AR_SPLIT; ++h) {
                                    const uint32_t boff = (uint32_t)h * AR_HN * 128u;           // AR_HN weight rows of 128 B
                                    const uint64_t bd = umma_desc_sw128((term == 1 ? pb + AR_B_BYTES : pb) + boff + k * 32, 16, 1024);
                                    umma_tf32_ts(d_main + h * AR_HN, term == 0 ? a_lo : a_hi, bd, idesc, term == 0 ? first : 1u);
                                }
                            }
                        }
                        tc_commit(b_empty(sb));
                        if (j == n_sub - 1) tc_commit(atm_empty(kb));      // the converter may refill this K block
                    }
                    tc_commit(tfull_bar(buf));
                }
            }
        }
    } else if (warp < 6) {
        // ===================================================================== converter: smem row -> (hi, lo) in TMEM
        const int wq = warp & 3;                             // TMEM lane quarter this warp may access
        const int row = wq * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
        uint32_t it = 0, mt = 0;
        for (int64_t t = blockIdx.x; t < p.n_tiles_m; t += gridDim.x, ++mt) {
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const int sa = it % AR_NA;
                mbar_wait(a_full(sa), (it / AR_NA) & 1u);
                const uint8_t *arow = smem + sa * A_BYTES + row * 128;
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {                // 128-byte swizzle: 16-byte chunk c of row r sits at c ^ (r & 7)
                    const float4 v = *reinterpret_cast<const float4 *>(arow + ((c ^ (row & 7)) << 4));
                    const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float h = tf32_hi(x[e]);
                        hi[4 * c + e] = __float_as_uint(h);
                        lo[4 * c + e] = __float_as_uint(x[e] - h);
                    }
                }
                mbar_wait(atm_empty(kb), (mt & 1u) ^ 1u);
                __syncwarp();        // lanes leave the spin loop one by one; tcgen05.st is .sync.aligned
                tc_fence_after();
                tmem_st_32x32(tmem_base + lane_addr + A_HI_COL + kb * BK, hi);
                tmem_st_32x32(tmem_base + lane_addr + A_LO_COL + kb * BK, lo);
                // Release the raw tile only now: the tcgen05.st instructions READ the registers the LDS.128 filled,
                // so the loads have completed.  An arrive issued right after the loads can overtake them (the
                // mbarrier unit does not queue behind LDS traffic that waits for shared-memory bandwidth) and let
                // the next TMA write land on rows that are still being read.
                __syncwarp();
                if (lane == 0) mbar_arrive(a_empty(sa));
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(atm_full(kb));
            }
        }
    } else if (warp < 14) {
        // ===================================================================== epilogue (two warpgroups, one per buffer)
        const int wq = warp & 3;
        const uint32_t g = (uint32_t)(warp - 6) >> 2;        // accumulator buffer this warpgroup drains
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
        // Full tiles of a plain row-major output leave through shared memory and the TMA unit: a thread owns a ROW of
        // the accumulator, so direct global stores cost one LSU wavefront per lane and instruction (the l1tex data pipe
        // was ~80 % busy with them); staging the 32 x 32 chunk in the 128-byte-swizzled layout (4 wavefronts per
        // instruction) and one TMA tile store per chunk takes the global stores off the LSU altogether.
        const uint32_t my_stg = stg_base + (uint32_t)(warp - 6) * 4096u;
        uint32_t si = 0, n_staged = 0;
        for (int64_t t = blockIdx.x; t < p.n_tiles_m; t += gridDim.x) {
            const Tile tl = locate_tile(p, t * p.n_tiles_n);
            const bool via_tma = p.tma_store && tl.row0 + BM <= tl.row_end;
            for (int j = 0; j < n_sub; ++j, ++si) {
                if ((si & 1u) != g) continue;
                mbar_wait(tfull_bar(g), (si >> 1) & 1u);
                __syncwarp();        // reconverge before the .sync.aligned tcgen05.ld
                tc_fence_after();
                const uint32_t acc = tmem_base + lane_addr + ACC_COL + g * 128u;
#pragma unroll 1
                for (int c = 0; c < AR_BN / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld_32x32(acc + c * 32, r);
                    tmem_ld_wait();
                    if (c == AR_BN / 32 - 1) {               // accumulator fully read: hand the buffer back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty_bar(g));
                    }
                    const int col0 = j * AR_BN + c * 32;
                    if (!via_tma) {
                        store_chunk(nullptr, r, lane, tl.row0 + wq * 32, tl.row_end, col0, p.N, p.C, p.ldc,
                                    p.out_map, p.bias, p.add, p.ldadd, p.add_in | (p.relu << 1));
                        __syncwarp();
                        continue;
                    }
                    if (col0 >= p.N) continue;
                    const uint32_t sbuf = my_stg;
                    if (lane == 0) tma_store_wait_read<0>();      // the previous tile store has read the buffer
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float4 v = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                               __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                        if (p.bias && col0 + 4 * q + 4 <= p.N) {
                            const float4 bb = __ldg(reinterpret_cast<const float4 *>(p.bias + col0) + q);
                            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
                        }
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sbuf + lane * 128u +
                                                                                       ((uint32_t)(q ^ (lane & 7)) << 4)),
                                     "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                                     : "memory");
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&p.map_c, sbuf, col0, (int)(tl.row0 + wq * 32));   // columns >= N, rows >= M: clipped
                        tma_store_commit();
                    }
                    ++n_staged;
                }
                __syncwarp();
            }
        }
        if (lane == 0) tma_store_wait_read<0>();             // shared memory must outlive the last tile stores
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------
// wgrad: dW[n, k] = sum_m G[m, n] * A[m, k]   (G = dL/dY [M, N], A = layer input [M, K])
//
// Both operands are "MN-major" for the tensor core (the reduction index m is the ROW of both
// row-major matrices): a stage holds 32 reduction rows as 4 + 4 TMA boxes of [32 rows x 32 columns]
// (TMA swizzle 128B_ATOM_32B), i.e. canonical MN-major SWIZZLE_128B_BASE32B atoms -- the only MN-major
// layout the tensor core accepts for 32-bit operands -- with LBO = 4096 B between 32-column groups and
// SBO = 512 B between groups of 4 reduction rows; one tcgen05.mma (K = 8) consumes 8 rows = 1024 B.  Both operands are activations, so the converter warps split both tiles; rows beyond the
// slab's end (slabs may end inside a 32-row block when they follow degree ranges) are zeroed there.
// The reduction over M is cut into slabs; a work unit = (slab, 128x128 output tile) writes its partial
// tile to part[slab_slot][N][K]; a fixed-order second pass sums the slabs (no atomics).
// ------------------------------------------------------------------------------------------
struct alignas(64) WgParams {
    CUtensorMap map_g0, map_g1, map_a;
    int n_split;                // output rows n < n_split come from map_g0, the rest from map_g1 (at n - n_split)
    int tiles_n, tiles_k, mode;
    int N, K;
    int n256;                   // G_hi A_hi and G_hi A_lo as one N = 256 MMA (MMA_GEMM_WG_N256)
    int64_t n_slabs;
    const int32_t *slab_tab;    // [n_slabs][4] = row0, row_end, out_slot, -
    float *part;                // [slots][N][K]
};

__global__ void __launch_bounds__(THREADS, 1) gemm_wgrad_kernel(const __grid_constant__ WgParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
    constexpr int RAW_BYTES = A_BYTES + B_BYTES;           // bytes TMA brings per stage: G [4 x 4 KB] | A [4 x 4 KB]
    constexpr int SLOT = WG_SLOT_BYTES;                    // stage: G | A | A_lo (written by the converter)
    const uint32_t lo_base = smem_base + WG_NR * SLOT;     // G_lo ring
    const uint32_t bar_base = lo_base + WG_NL * A_BYTES;
    // barriers: full[NR] (TMA landed), empty[NR] (MMAs retired), lo_full[NL] (converted), lo_empty[NL]
    auto full_bar = [&](int i) { return bar_base + 8u * i; };
    auto empty_bar = [&](int i) { return bar_base + 8u * (WG_NR + i); };
    auto lo_full = [&](int i) { return bar_base + 8u * (2 * WG_NR + i); };
    auto lo_empty = [&](int i) { return bar_base + 8u * (2 * WG_NR + WG_NL + i); };
    constexpr int kTBar = 2 * WG_NR + 2 * WG_NL;
    auto tfull_bar = [&](int b) { return bar_base + 8u * (kTBar + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (kTBar + 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (kTBar + 4);
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles = p.tiles_n * p.tiles_k;
    const int64_t n_units = p.n_slabs * tiles;
    struct Unit { int row0, row_end, slot, n0, k0; };
    auto locate = [&](int64_t u) {
        Unit x;
        const int64_t sl = u / tiles;
        const int tile = (int)(u - sl * tiles);
        const int4 e = __ldg(reinterpret_cast<const int4 *>(p.slab_tab) + sl);
        x.row0 = e.x; x.row_end = e.y; x.slot = e.z;
        x.n0 = (tile / p.tiles_k) * BM;
        x.k0 = (tile % p.tiles_k) * BN;
        return x;
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.map_g0); tma_prefetch_desc(&p.map_g1); tma_prefetch_desc(&p.map_a);
        for (int i = 0; i < WG_NR; ++i) { mbar_init(full_bar(i), 1); mbar_init(empty_bar(i), 1); }
        for (int i = 0; i < WG_NL; ++i) { mbar_init(lo_full(i), 4); mbar_init(lo_empty(i), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (elect_one()) {
            uint32_t it = 0;
            for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
                const Unit x = locate(u);
                const CUtensorMap *mg = x.n0 < p.n_split ? &p.map_g0 : &p.map_g1;
                const int gc0 = x.n0 < p.n_split ? x.n0 : x.n0 - p.n_split;
                for (int r = x.row0; r < x.row_end; r += BK, ++it) {
                    const int s = it % WG_NR;
                    mbar_wait(empty_bar(s), ((it / WG_NR) & 1u) ^ 1u);
                    const uint32_t sa = smem_base + s * SLOT;
                    mbar_expect_tx(full_bar(s), RAW_BYTES);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        tma_load_2d(sa + g * 4096, mg, full_bar(s), gc0 + g * 32, r);
                        tma_load_2d(sa + A_BYTES + g * 4096, &p.map_a, full_bar(s), x.k0 + g * 32, r);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_tf32(BM, BN, 1, 1);
            constexpr uint32_t idesc2 = umma_idesc_tf32(BM, 2 * BN, 1, 1);
            uint32_t it = 0, ti = 0;
            for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x, ++ti) {
                const Unit x = locate(u);
                const uint32_t buf = ti & 1u;
                mbar_wait(tempty_bar(buf), ((ti >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 2 * BN;
                uint32_t acc = 0;
                for (int r = x.row0; r < x.row_end; r += BK, ++it) {
                    const int s = it % WG_NR, sl = it % WG_NL;
                    mbar_wait(full_bar(s), (it / WG_NR) & 1u);
                    mbar_wait(lo_full(sl), (it / WG_NL) & 1u);
                    tc_fence_after();
                    const uint32_t sa = smem_base + s * SLOT, sq = lo_base + sl * A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        const uint64_t g_hi = umma_desc(sa + k * 1024, 4096, 512, 1);
                        const uint64_t g_lo = umma_desc(sq + k * 1024, 4096, 512, 1);
                        const uint64_t a_hi = umma_desc(sa + A_BYTES + k * 1024, 4096, 512, 1);
                        if (p.mode != 2 && p.n256) {
                            // G_hi A_hi (main) and G_hi A_lo (correction) as ONE instruction of N = 256: A_lo lies right
                            // behind the raw A tile (8 column groups 4096 B apart), the two accumulators are adjacent
                            // in tensor memory.  G_hi crosses shared memory once instead of twice, 2 MMAs per K step
                            // instead of 3: measured 5-10 % (profiles/r2M_wgrad_n256_ab.log)
                            umma_tf32_ss(d_tmem, g_hi, a_hi, idesc2, acc);              // [main | corr] (+)= G_hi^T [A_hi | A_lo]
                            umma_tf32_ss(d_tmem + BN, g_lo, a_hi, idesc, 1u);           // corr += G_lo^T A_hi
                        } else {
                            if (p.mode != 2) {
                                const uint64_t a_lo = umma_desc(sa + A_BYTES + B_BYTES + k * 1024, 4096, 512, 1);
                                umma_tf32_ss(d_tmem + BN, g_lo, a_hi, idesc, acc);
                                umma_tf32_ss(d_tmem + BN, g_hi, a_lo, idesc, 1u);
                            }
                            umma_tf32_ss(d_tmem, g_hi, a_hi, idesc, acc);
                        }
                        acc = 1u;
                    }
                    tc_commit(empty_bar(s));
                    tc_commit(lo_empty(sl));
                }
                tc_commit(tfull_bar(buf));
            }
        }
    } else if (warp < 6) {
        const int ct = threadIdx.x - 64;
        uint32_t it = 0;
        for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
            const Unit x = locate(u);
            for (int r = x.row0; r < x.row_end; r += BK, ++it) {
                const int s = it % WG_NR, sl = it % WG_NL;
                const int valid = x.row_end - r;            // rows of this block inside the slab (>= 32: all)
                mbar_wait(full_bar(s), (it / WG_NR) & 1u);
                mbar_wait(lo_empty(sl), ((it / WG_NL) & 1u) ^ 1u);
                if (p.mode == 1 && valid >= BK) {
                    // the common case, kept lean (the converter warps set the pace of this kernel): a full block of
                    // 32 rows, raw operand as hi (the tensor core ignores the low 13 mantissa bits), lo = x - trunc(x)
#pragma unroll
                    for (int op = 0; op < 2; ++op) {
                        const float4 *a = reinterpret_cast<const float4 *>(smem + s * SLOT + op * A_BYTES);
                        float4 *alo = reinterpret_cast<float4 *>(op == 0 ? smem + (lo_base - smem_base) + sl * A_BYTES
                                                                         : smem + s * SLOT + A_BYTES + B_BYTES);
                        float4 v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = a[i * 128 + ct];
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            alo[i * 128 + ct] = make_float4(v[i].x - tf32_hi(v[i].x), v[i].y - tf32_hi(v[i].y),
                                                            v[i].z - tf32_hi(v[i].z), v[i].w - tf32_hi(v[i].w));
                    }
                } else {
#pragma unroll
                    for (int op = 0; op < 2; ++op) {
                        float4 *a = reinterpret_cast<float4 *>(smem + s * SLOT + op * A_BYTES);
                        float4 *alo = reinterpret_cast<float4 *>(op == 0 ? smem + (lo_base - smem_base) + sl * A_BYTES
                                                                         : smem + s * SLOT + A_BYTES + B_BYTES);
                        float4 v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int idx = i * 128 + ct;       // 16-byte chunk; box g = idx/256, row = (idx/8) % 32
                            v[i] = a[idx];
                            if (((idx >> 3) & 31) >= valid) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int idx = i * 128 + ct;
                            if (p.mode == 0) {
                                const float4 h = make_float4(tf32_rna(v[i].x), tf32_rna(v[i].y), tf32_rna(v[i].z), tf32_rna(v[i].w));
                                alo[idx] = make_float4(tf32_rna(v[i].x - h.x), tf32_rna(v[i].y - h.y), tf32_rna(v[i].z - h.z),
                                                       tf32_rna(v[i].w - h.w));
                                a[idx] = h;
                            } else {
                                const float4 h = make_float4(tf32_hi(v[i].x), tf32_hi(v[i].y), tf32_hi(v[i].z), tf32_hi(v[i].w));
                                if (p.mode != 2) alo[idx] = make_float4(v[i].x - h.x, v[i].y - h.y, v[i].z - h.z, v[i].w - h.w);
                                if (valid < BK) a[idx] = h;
                            }
                        }
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(lo_full(sl));
            }
        }
    } else {
        const int wq = warp & 3;
        float *stg = nullptr;                    // store_chunk writes straight from registers
        uint32_t ti = 0;
        for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x, ++ti) {
            const Unit x = locate(u);
            const uint32_t buf = ti & 1u;
            mbar_wait(tfull_bar(buf), (ti >> 1) & 1u);
            __syncwarp();            // reconverge before the .sync.aligned tcgen05.ld
            tc_fence_after();
            float *out = p.part + (int64_t)x.slot * p.N * p.K;
            const int n_chunks = min(4, (p.K - x.k0 + 31) / 32);
            for (int c = 0; c < n_chunks; ++c) {
                uint32_t r[32];
                load_acc(tmem_base + buf * 2 * BN + c * 32 + ((uint32_t)(wq * 32) << 16), p.mode != 2, r);
                if (c == n_chunks - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(buf));
                }
                store_chunk(stg, r, lane, x.n0 + wq * 32, p.N, x.k0 + c * 32, p.K, out, p.K, nullptr, nullptr, nullptr, 0, 0);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// out[i] = sum_s coef[s] * part[s][i] (coef null: 1).  Block = 32 float4 columns x 32 slab partitions:
// partition ty sums its contiguous range of slabs in ascending order, then the 32 partial sums are added
// in ascending ty -- a fixed summation tree, so the result is bit-reproducible (no atomics).
__global__ void __launch_bounds__(1024) reduce_slabs_kernel(const float *__restrict__ part, const float *__restrict__ coef,
                                                            int64_t n_slots, int64_t n4, float *__restrict__ out) {
    __shared__ float4 sm[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t i = (int64_t)blockIdx.x * 32 + tx;
    const int64_t per = (n_slots + 31) / 32;
    const int64_t s0 = ty * per, s1 = s0 + per < n_slots ? s0 + per : n_slots;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n4) {
        // 8 loads in flight per thread, added in ascending s: with few outputs (a column sum: n4 = 32, thousands of
        // slots) ONE block does all the work and the kernel lasts as long as its chain of round trips (ncu: 75 us at
        // 4736 slots with 4 in flight)
        int64_t s = s0;
        for (; s + 8 <= s1; s += 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcs(reinterpret_cast<const float4 *>(part) + (s + u) * n4 + i);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float c = coef ? __ldg(coef + s + u) : 1.0f;
                acc.x += c * v[u].x; acc.y += c * v[u].y; acc.z += c * v[u].z; acc.w += c * v[u].w;
            }
        }
        for (; s < s1; ++s) {
            const float4 v = __ldcs(reinterpret_cast<const float4 *>(part) + s * n4 + i);
            const float c = coef ? __ldg(coef + s) : 1.0f;
            acc.x += c * v.x; acc.y += c * v.y; acc.z += c * v.z; acc.w += c * v.w;
        }
    }
    sm[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && i < n4) {
        float4 t = sm[0][tx];
        for (int k = 1; k < 32; ++k) { const float4 v = sm[k][tx]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
        reinterpret_cast<float4 *>(out)[i] = t;
    }
}

// out[g][i] = sum over slots s in [seg_ptr[g], seg_ptr[g+1]) of part[s][i], ascending s
__global__ void reduce_slabs_seg_kernel(const float *__restrict__ part, const int32_t *__restrict__ seg_ptr,
                                        int64_t n4, float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const int g = blockIdx.y;
    const int s0 = __ldg(seg_ptr + g), s1 = __ldg(seg_ptr + g + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = s0; s < s1; ++s) {
        const float4 v = __ldcs(reinterpret_cast<const float4 *>(part) + (int64_t)s * n4 + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4 *>(out)[(int64_t)g * n4 + i] = acc;
}

// ------------------------------------------------------------------------------------------
// weight pre-split
// ------------------------------------------------------------------------------------------
__global__ void split_tf32_kernel(const float *__restrict__ w, float *__restrict__ hi, float *__restrict__ lo, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float x = w[i];
        const float h = tf32_rna(x);
        hi[i] = h;
        lo[i] = tf32_rna(x - h);
    }
}

// ------------------------------------------------------------------------------------------
// host side: tensor maps
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            ptr = nullptr;
        return reinterpret_cast<EncodeTiledFn>(ptr);
    }();
    return fn;
}

// 2-D fp32 row-major tensor [rows, cols] with leading dimension ld (elements); box = [box_rows x box_cols],
// 128-byte swizzle (box_cols * 4 must be 128).  Out-of-bounds elements are zero-filled.
int make_map_2d(CUtensorMap *m, const float *base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols,
                CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return MMA_ERR_CUDA;
    if (!aligned16(base) || (ld % 4) != 0 || rows < 1 || cols < 1) return MMA_ERR_INVALID;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? MMA_OK : MMA_ERR_INVALID;
}

}  // namespace gemm
}  // namespace mma

using namespace mma;
using namespace mma::gemm;

extern "C" int mma_tf32_split(const float *w, float *hi, float *lo, int64_t n, mma_stream_t stream) {
    if (n < 0 || (n > 0 && (!w || !hi || !lo))) return MMA_ERR_INVALID;
    if (n == 0) return MMA_OK;
    split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(w, hi, lo, n);
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}

extern "C" int mma_linear_tf32x3(const float *A0, int64_t lda0, int K0, const float *A1, int64_t lda1, int K1,
                                 const float *Bhi, const float *Blo, int64_t ldb, int64_t b_rows,
                                 int64_t M, int N, const int32_t *tile_tab, int64_t n_tiles_m,
                                 float *C, int64_t ldc, const int32_t *out_map, const float *bias,
                                 const float *add, int64_t ldadd, int mode_flags, int max_ctas, mma_stream_t stream) {
    if (!A0 || !Bhi || !Blo || !C || M < 0 || N < 1 || K0 < 1 || K1 < 0 || b_rows < 1) return MMA_ERR_INVALID;
    const int mode = mode_flags & 3;
    if (mode_flags < 0 || mode > 2 || (mode_flags & ~(3 | MMA_GEMM_ADD_BY_INPUT_ROW | MMA_GEMM_RELU))) return MMA_ERR_INVALID;
    if ((N % 4) != 0 || (ldc % 4) != 0 || !aligned16(C) || !aligned16(bias) || !aligned16(add) || (ldadd % 4) != 0)
        return MMA_ERR_UNSUPPORTED;
    if (K1 > 0 && (!A1 || (K0 % BK) != 0)) return MMA_ERR_INVALID;
    if (M >= INT32_MAX || b_rows >= INT32_MAX) return MMA_ERR_UNSUPPORTED;
    if (M == 0) return MMA_OK;
    NtParams p;
    memset(&p, 0, sizeof(p));
    int rc;
    if ((rc = make_map_2d(&p.map_a0, A0, M, K0, lda0, BM, BK)) != MMA_OK) return rc;
    if (K1 > 0) { if ((rc = make_map_2d(&p.map_a1, A1, M, K1, lda1, BM, BK)) != MMA_OK) return rc; }
    else p.map_a1 = p.map_a0;
    if ((rc = make_map_2d(&p.map_bhi, Bhi, b_rows, K0 + K1, ldb, BN, BK)) != MMA_OK) return rc;
    if ((rc = make_map_2d(&p.map_blo, Blo, b_rows, K0 + K1, ldb, BN, BK)) != MMA_OK) return rc;
    p.kb_split = K1 > 0 ? K0 / BK : (K0 + BK - 1) / BK;
    p.num_kb = p.kb_split + (K1 + BK - 1) / BK;
    p.n_tiles_n = (N + BN - 1) / BN;
    p.mode = mode;
    p.M = M; p.N = N;
    p.tile_tab = tile_tab;
    p.n_tiles_m = tile_tab ? n_tiles_m : (M + BM - 1) / BM;
    if (p.n_tiles_m < 1) return tile_tab ? MMA_OK : MMA_ERR_INVALID;
    p.C = C; p.ldc = ldc; p.out_map = out_map; p.bias = bias; p.add = add; p.ldadd = ldadd;
    p.add_in = (mode_flags & MMA_GEMM_ADD_BY_INPUT_ROW) ? 1 : 0;
    p.relu = (mode_flags & MMA_GEMM_RELU) ? 1 : 0;
    int sms = 0, dev = 0;
    MMA_CUDA_CHECK(cudaGetDevice(&dev));
    MMA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;
    // short K, several 128-column tiles per row block: keep the activation tile resident in tensor memory
    static const bool ares_on = [] { const char *e = getenv("MMA_GEMM_ARES"); return !(e && e[0] == '0'); }();
    if (ares_on && mode == 1 && K1 == 0 && p.num_kb <= AR_MAX_KB && N >= 2 * BN) {
        if ((rc = make_map_2d(&p.map_bhi, Bhi, b_rows, K0, ldb, AR_BN, BK)) != MMA_OK) return rc;
        if ((rc = make_map_2d(&p.map_blo, Blo, b_rows, K0, ldb, AR_BN, BK)) != MMA_OK) return rc;
        static const bool tma_st = [] { const char *e = getenv("MMA_GEMM_TMA_STORE"); return !(e && e[0] == '0'); }();
        p.tma_store = (tma_st && !out_map && !add && !p.relu && make_map_2d(&p.map_c, C, M, N, ldc, 32, 32) == MMA_OK) ? 1 : 0;
        MMA_CUDA_CHECK(cudaFuncSetAttribute(gemm_nt_ares_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AR_SMEM_BYTES));
        const unsigned grid = (unsigned)(p.n_tiles_m < sms ? p.n_tiles_m : sms);
        gemm_nt_ares_kernel<<<grid, AR_THREADS, AR_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(p);
        MMA_LAUNCH_CHECK();
        return MMA_OK;
    }
    static const bool n256_on = [] { const char *e = getenv("MMA_GEMM_N256"); return !(e && e[0] == '0'); }();
    p.n256 = n256_on ? 1 : 0;
    MMA_CUDA_CHECK(cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NT_SMEM_BYTES));
    const int64_t n_tiles = p.n_tiles_m * p.n_tiles_n;
    const unsigned grid = (unsigned)(n_tiles < sms ? n_tiles : sms);
    gemm_nt_kernel<<<grid, NT_THREADS, NT_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}

extern "C" int mma_wgrad_tf32x3(const float *G0, int64_t ldg0, int N0, const float *G1, int64_t ldg1, int N1,
                                const float *A, int64_t lda, int K, int64_t M, const int32_t *slab_tab,
                                int64_t n_slabs, float *part, int mode, int max_ctas, mma_stream_t stream) {
    if (!G0 || !A || !part || !slab_tab || M < 1 || N0 < 1 || N1 < 0 || K < 1 || n_slabs < 0) return MMA_ERR_INVALID;
    if (mode < 0 || mode > 2) return MMA_ERR_INVALID;
    if (N1 > 0 && (!G1 || (N0 % BM) != 0)) return MMA_ERR_INVALID;
    if ((K % 4) != 0 || !aligned16(part) || M >= INT32_MAX) return MMA_ERR_UNSUPPORTED;
    if (n_slabs == 0) return MMA_OK;
    WgParams p;
    memset(&p, 0, sizeof(p));
    int rc;
    if ((rc = make_map_2d(&p.map_g0, G0, M, N0, ldg0, BK, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) != MMA_OK) return rc;
    if (N1 > 0) { if ((rc = make_map_2d(&p.map_g1, G1, M, N1, ldg1, BK, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) != MMA_OK) return rc; }
    else p.map_g1 = p.map_g0;
    if ((rc = make_map_2d(&p.map_a, A, M, K, lda, BK, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) != MMA_OK) return rc;
    p.N = N0 + N1; p.K = K;
    p.n_split = N1 > 0 ? N0 : INT32_MAX;
    p.tiles_n = (p.N + BM - 1) / BM;
    p.tiles_k = (K + BN - 1) / BN;
    p.mode = mode;
    p.n_slabs = n_slabs; p.slab_tab = slab_tab; p.part = part;
    static const bool wg_n256 = [] { const char *e = getenv("MMA_GEMM_WG_N256"); return !(e && e[0] == '0'); }();
    p.n256 = wg_n256 ? 1 : 0;
    MMA_CUDA_CHECK(cudaFuncSetAttribute(gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES));
    int sms = 0, dev = 0;
    MMA_CUDA_CHECK(cudaGetDevice(&dev));
    MMA_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;
    const int64_t n_units = n_slabs * p.tiles_n * p.tiles_k;
    const unsigned grid = (unsigned)(n_units < sms ? n_units : sms);
    gemm_wgrad_kernel<<<grid, THREADS, WG_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}

extern "C" int mma_reduce_slabs(const float *part, const float *coef, int64_t n_slots, int64_t n, float *out,
                                mma_stream_t stream) {
    if (!part || !out || n_slots < 1 || n < 0 || (n % 4) != 0 || !aligned16(part) || !aligned16(out))
        return MMA_ERR_INVALID;
    if (n == 0) return MMA_OK;
    const int64_t n4 = n / 4;
    reduce_slabs_kernel<<<(unsigned)((n4 + 31) / 32), dim3(32, 32), 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        part, coef, n_slots, n4, out);
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}

extern "C" int mma_reduce_slabs_segmented(const float *part, const int32_t *seg_ptr, int64_t n_segs, int64_t n,
                                          float *out, mma_stream_t stream) {
    if (!part || !seg_ptr || !out || n_segs < 0 || n < 0 || (n % 4) != 0 || !aligned16(part) || !aligned16(out))
        return MMA_ERR_INVALID;
    if (n == 0 || n_segs == 0) return MMA_OK;
    if (n_segs > 65535) return MMA_ERR_UNSUPPORTED;
    const int64_t n4 = n / 4;
    dim3 grid((unsigned)((n4 + 255) / 256), (unsigned)n_segs);
    reduce_slabs_seg_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(part, seg_ptr, n4, out);
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}
