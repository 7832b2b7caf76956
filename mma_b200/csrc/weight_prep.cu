// Weight-space algebra of the fused MultiMaskConv layer (mma_b200/fused_layer.py) in a handful of launches.
//
// The layer composes the reference's post Linear (graph_regression/mma_conv.py:132-133), its `lin` (:136) and the S
// cumulative scalers (:181-196) into ONE effective weight per in-degree, W_c(d) = sum_s sum_a c_s(d) coef_a(d) W_lin W_{s,a},
// and recovers the gradients of W_lin / W_post from the gradient of W_c(d).  These products are tiny ([128 x 128 x 2688]
// at the benchmark configuration) but were ~60 torch launches per step (batched matmuls, einsum, cat, split, transpose),
// 0.5-0.9 ms that every rank of a sharded run repeats.  Here: one small fp32 GEMM kernel (FFMA, fixed summation order,
// optional split-K into slabs that mma_reduce_slabs adds in order -- no atomics) and two composition kernels that
// write their results already split for the 3xTF32 GEMMs (hi = tf32(x), lo = tf32(x - hi), as mma_tf32_split).
#include "common.cuh"
#include "tc_common.cuh"

namespace mma {
namespace wprep {

using tc::tf32_rna;

constexpr int TM = 64, TN = 64, TK = 16;

// C[z] [M, N] (ldc) = op(A) [M, K-range of split z] * op(B) [K-range, N];  op = transpose when the flag is set:
// A is [M, K] (lda) or, transposed, [K, M]; B is [K, N] (ldb) or, transposed, [N, K].  256 threads, 4 x 4 outputs per
// thread; a thread fetches ONE 128-bit piece of each operand tile per k-step (VEC: bases and leading dimensions are
// 16-byte aligned and the contiguous extent is a multiple of 4; otherwise four scalar loads), and the fetch of tile
// k + 1 is in flight while tile k is multiplied.
struct TileLoad {
    const float *p; int64_t ld; int trans, rows, k_hi;      // rows: extent of the non-k dimension
    // element (r, k) of the operand tile at (r0, k0): thread t owns 4 elements that are contiguous in memory
    template <bool VEC>
    __device__ __forceinline__ float4 fetch(int t, int r0, int k0, int &r, int &k, bool k_contig) const {
        // k_contig: memory runs along k (A untransposed / B transposed), else along r
        if (k_contig) { r = t >> 2; k = (t & 3) * 4; } else { k = t >> 4; r = (t & 15) * 4; }
        const int gr = r0 + r, gk = k0 + k;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k_contig) {
            if (gr < rows) {
                const float *q = p + (int64_t)gr * ld + gk;
                if (VEC && gk + 3 < k_hi) v = __ldg(reinterpret_cast<const float4 *>(q));
                else {
                    if (gk < k_hi) v.x = __ldg(q);
                    if (gk + 1 < k_hi) v.y = __ldg(q + 1);
                    if (gk + 2 < k_hi) v.z = __ldg(q + 2);
                    if (gk + 3 < k_hi) v.w = __ldg(q + 3);
                }
            }
        } else {
            if (gk < k_hi) {
                const float *q = p + (int64_t)gk * ld + gr;
                if (VEC && gr + 3 < rows) v = __ldg(reinterpret_cast<const float4 *>(q));
                else {
                    if (gr < rows) v.x = __ldg(q);
                    if (gr + 1 < rows) v.y = __ldg(q + 1);
                    if (gr + 2 < rows) v.z = __ldg(q + 2);
                    if (gr + 3 < rows) v.w = __ldg(q + 3);
                }
            }
        }
        return v;
    }
};

template <bool VEC>
__global__ void __launch_bounds__(256) small_gemm_kernel(const float *__restrict__ A, int64_t lda, int ta,
                                                          const float *__restrict__ B, int64_t ldb, int tb,
                                                          float *__restrict__ C, int64_t ldc, int M, int N, int K,
                                                          int k_per_split) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int k_lo = blockIdx.z * k_per_split;
    const int k_hi = min(K, k_lo + k_per_split);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const TileLoad la{A, lda, ta, M, k_hi}, lb{B, ldb, tb, N, k_hi};
    const bool a_kc = !ta, b_kc = tb != 0;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    int ar, ak, br, bk;
    float4 av = la.fetch<VEC>(tid, m0, k_lo, ar, ak, a_kc);
    float4 bv = lb.fetch<VEC>(tid, n0, k_lo, br, bk, b_kc);
    for (int k0 = k_lo; k0 < k_hi; k0 += TK) {
        if (a_kc) { As[ak][ar] = av.x; As[ak + 1][ar] = av.y; As[ak + 2][ar] = av.z; As[ak + 3][ar] = av.w; }
        else *reinterpret_cast<float4 *>(&As[ak][ar]) = av;
        if (b_kc) { Bs[bk][br] = bv.x; Bs[bk + 1][br] = bv.y; Bs[bk + 2][br] = bv.z; Bs[bk + 3][br] = bv.w; }
        else *reinterpret_cast<float4 *>(&Bs[bk][br]) = bv;
        __syncthreads();
        if (k0 + TK < k_hi) {                               // next tile: in flight during the multiply
            av = la.fetch<VEC>(tid, m0, k0 + TK, ar, ak, a_kc);
            bv = lb.fetch<VEC>(tid, n0, k0 + TK, br, bk, b_kc);
        }
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *Cz = C + (int64_t)blockIdx.z * M * ldc;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn < N) Cz[(int64_t)gm * ldc + gn] = acc[i][j];
        }
    }
}

// W_c[b][c][m*F + f] = sum_s sum_a coef[b][s][a][m] * WlW[c][col0 + (s*A + a)*F + f]   (coef is 0 where aggregator a does not
// feed block m), written split: hi / lo [B][Co][K] and, when asked for, the transposes hiT / loT [B][K][Co] (the
// dgrad GEMM's weight).  K = Am * F.  One block per (range b, 32 rows c, 32 columns k): the tile goes through shared
// memory so that both layouts are written with full 128-byte lines.
__global__ void __launch_bounds__(256) compose_fwd_kernel(const float *__restrict__ coef, int nb, int S, int A, int Am,
                                                           const float *__restrict__ WlW, int64_t ldw, int col0, int Co,
                                                           int F, float *__restrict__ hi, float *__restrict__ lo,
                                                           float *__restrict__ hiT, float *__restrict__ loT) {
    __shared__ float th[32][33], tl[32][33];
    __shared__ float cf[64];                                // coef[b][.][.][m] of this tile's block m (S * A <= 64)
    const int K = Am * F;
    const int b = blockIdx.z, c0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    // a 32-column tile may straddle two blocks m only when F % 32 != 0: then every thread reads its own coefficients
    const bool one_m = (k0 / F) == ((min(k0 + 31, K - 1)) / F);
    if (one_m && threadIdx.x < S * A) cf[threadIdx.x] = __ldg(coef + ((int64_t)b * S * A + threadIdx.x) * Am + k0 / F);
    __syncthreads();
    const int k = k0 + tx;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int cl = ty + 8 * r, c = c0 + cl;
        float acc = 0.0f;
        if (k < K && c < Co) {
            const int m = k / F, f = k - m * F;
            const float *w = WlW + (int64_t)c * ldw + col0 + f;
            for (int sa = 0; sa < S * A; ++sa) {
                const float q = one_m ? cf[sa] : __ldg(coef + ((int64_t)b * S * A + sa) * Am + m);
                if (q != 0.0f) acc = __fmaf_rn(q, __ldg(w + (int64_t)sa * F), acc);
            }
            const float h = tf32_rna(acc), l = tf32_rna(acc - h);
            const int64_t i = ((int64_t)b * Co + c) * K + k;
            hi[i] = h; lo[i] = l;
            th[cl][tx] = h; tl[cl][tx] = l;
        }
    }
    if (!hiT) return;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int kl = ty + 8 * r, kk = k0 + kl, c = c0 + tx;
        if (kk < K && c < Co) {
            const int64_t t = ((int64_t)b * K + kk) * Co + c;
            hiT[t] = th[tx][kl]; loT[t] = tl[tx][kl];
        }
    }
}

// Gradient of the composition: D[c][col0 + (s*A + a)*F + f] = sum_b coef[b][s][a][m(a)] * dWc[b][c][m(a)*F + f], b ascending
// (m(a) = block_of[a]); columns [0, col0) of D are filled from dX [Co][col0] (the gradient of the composed x-part).
__global__ void __launch_bounds__(256) compose_bwd_kernel(const float *__restrict__ coef, int nb, int S, int A, int Am,
                                                           const int32_t *__restrict__ block_of,
                                                           const float *__restrict__ dWc, int Co, int F,
                                                           const float *__restrict__ dX, int64_t lddx, int col0,
                                                           float *__restrict__ D, int64_t ldd) {
    const int K = Am * F;
    const int W = col0 + S * A * F;
    const int64_t n = (int64_t)Co * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(i % W);
        const int c = (int)(i / W);
        float acc = 0.0f;
        if (col < col0) {
            acc = dX ? __ldg(dX + (int64_t)c * lddx + col) : 0.0f;
        } else {
            const int r = col - col0;
            const int sa = r / F, f = r - sa * F;
            const int a = sa % A;
            const int m = __ldg(block_of + a);
            const float *cf = coef + (int64_t)sa * Am + m;
            const float *g = dWc + (int64_t)c * K + m * F + f;
            for (int b = 0; b < nb; ++b)
                acc = __fmaf_rn(__ldg(cf + (int64_t)b * S * A * Am), __ldg(g + (int64_t)b * Co * K), acc);
        }
        D[(int64_t)c * ldd + col] = acc;
    }
}

}  // namespace wprep
}  // namespace mma

using namespace mma;

extern "C" int mma_small_gemm(const float *A, int64_t lda, int trans_a, const float *B, int64_t ldb, int trans_b,
                              float *C, int64_t ldc, int M, int N, int K, int k_splits, mma_stream_t stream) {
    if (!A || !B || !C || M < 1 || N < 1 || K < 1 || k_splits < 1 || lda < 1 || ldb < 1 || ldc < N) return MMA_ERR_INVALID;
    int per = (K + k_splits - 1) / k_splits;
    per = (per + wprep::TK - 1) / wprep::TK * wprep::TK;
    const int splits = (K + per - 1) / per;
    if (splits != k_splits) return MMA_ERR_INVALID;          // the caller sized C for k_splits slabs: see mma_small_gemm_splits
    dim3 grid((N + wprep::TN - 1) / wprep::TN, (M + wprep::TM - 1) / wprep::TM, splits);
    auto al = [](const float *p, int64_t ld) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && (ld % 4) == 0; };
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (al(A, lda) && al(B, ldb))
        wprep::small_gemm_kernel<true><<<grid, 256, 0, st>>>(A, lda, trans_a, B, ldb, trans_b, C, ldc, M, N, K, per);
    else
        wprep::small_gemm_kernel<false><<<grid, 256, 0, st>>>(A, lda, trans_a, B, ldb, trans_b, C, ldc, M, N, K, per);
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}

extern "C" int mma_small_gemm_splits(int K, int wanted) {
    if (K < 1 || wanted < 1) return 1;
    int per = (K + wanted - 1) / wanted;
    per = (per + wprep::TK - 1) / wprep::TK * wprep::TK;
    return (K + per - 1) / per;
}

extern "C" int mma_compose_post_weight(const float *coef, int n_ranges, int S, int A, int Am, const float *WlW,
                                       int64_t ldw, int col0, int Co, int F, float *hi, float *lo, float *hiT,
                                       float *loT, mma_stream_t stream) {
    if (!coef || !WlW || !hi || !lo || (hiT == nullptr) != (loT == nullptr)) return MMA_ERR_INVALID;
    if (n_ranges < 1 || S < 1 || A < 1 || Am < 1 || Co < 1 || F < 1 || col0 < 0) return MMA_ERR_INVALID;
    if (S * A > 64) return MMA_ERR_UNSUPPORTED;
    dim3 grid((Am * F + 31) / 32, (Co + 31) / 32, n_ranges);
    wprep::compose_fwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(coef, n_ranges, S, A, Am, WlW, ldw,
                                                                                       col0, Co, F, hi, lo, hiT, loT);
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}

extern "C" int mma_compose_post_wgrad(const float *coef, int n_ranges, int S, int A, int Am, const int32_t *block_of,
                                      const float *dWc, int Co, int F, const float *dX, int64_t lddx, int col0, float *D,
                                      int64_t ldd, mma_stream_t stream) {
    if (!coef || !block_of || !dWc || !D) return MMA_ERR_INVALID;
    if (n_ranges < 1 || S < 1 || A < 1 || Am < 1 || Co < 1 || F < 1 || col0 < 0 || ldd < col0 + (int64_t)S * A * F)
        return MMA_ERR_INVALID;
    const int64_t n = (int64_t)Co * (col0 + (int64_t)S * A * F);
    const unsigned grid = (unsigned)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    wprep::compose_bwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(coef, n_ranges, S, A, Am, block_of,
                                                                                       dWc, Co, F, dX, lddx, col0, D, ldd);
    MMA_LAUNCH_CHECK();
    return MMA_OK;
}
