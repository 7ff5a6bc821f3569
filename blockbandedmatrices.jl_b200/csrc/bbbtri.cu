// Triangular BandedBlockBandedMatrix solve:  B <- T^-1 B,  T = Upper/LowerTriangular (unit or not) of A.
//
// Reference behaviour: `ldiv!(LowerTriangular(A), b)` / `Ldiv(U, b)` with the layouts
// TriangularLayout{UPLO,UNIT,BandedBlockBandedColumnMajor} (test/test_triblockbanded.jl:80-99; the Gauss-Seidel
// sweep of examples/finitedifference_2d.jl:17-47): [upstream] BlockArrays block substitution with a banded
// tbsv on each diagonal block and one gbmv per off-diagonal block -- strictly sequential on the CPU.
//
// Here: a systolic level schedule.  Unknown (K,k) of a lower solve depends on (K, k-lam..k-1) and on (J, j) with
// K-l <= J < K, k-lam <= j <= k+mu.  With the skew delta = mu+1, unknown (K,k) is computed at step
// s = k + delta*K: every dependency was computed at a step < s, all unknowns of one step are independent, and
// block row K simply walks down its rows one per step.  One thread owns one block row (or a few, round-robin);
// the last D solution values of every block row live in a shared-memory ring, read by the owners of the next
// l block rows -- through distributed shared memory when those sit in another CTA of the 8-CTA thread-block
// cluster -- and the steps are separated by the hardware cluster barrier: no global-memory round trip, no
// spin-waits, 4096 block rows in flight on 8 SMs.  Upper solves run the mirror image (block rows and rows
// descending, delta = lam+1).  n + delta*N steps in total (12 285 for the 16.7M-unknown cfg2 Laplacian).
#include <cooperative_groups.h>

#include <algorithm>

#include "bbm_internal.h"

namespace cg = cooperative_groups;

namespace {

constexpr int kTriCtas = 8;        // portable cluster size
constexpr int kTriThreads = 512;
constexpr int kTriLanes = kTriCtas * kTriThreads;

template <typename T>
struct TriSolveParams {
    const T* data;
    T* x;              // right-hand side in, solution out
    const i64* R;      // prefix sums of the (square) block sizes
    i64 N, maxr, iters;
    int l, u, lam, mu, w, P;
    int upper, unit, passes, D;
};

template <typename T>
__global__ void __cluster_dims__(kTriCtas, 1, 1) __launch_bounds__(kTriThreads) bbb_trisolve_kernel(const TriSolveParams<T> p) {
    extern __shared__ __align__(16) unsigned char tri_ring_smem[];
    T* ring = reinterpret_cast<T*>(tri_ring_smem);  // ring[(pass*kTriThreads + tid)*D + (k & (D-1))]
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x;
    const int g = (int)cluster.block_rank() * kTriThreads + tid;
    const int D = p.D, w = p.w, P = p.P;
    const int delta = p.upper ? p.lam + 1 : p.mu + 1;
    const int nd = p.upper ? p.u : p.l;  // how many block rows back the off-diagonal blocks reach
    const T* __restrict__ data = p.data;

    for (i64 s = 0; s < p.iters; ++s) {
        for (int pass = 0; pass < p.passes; ++pass) {
            const i64 Kl = (i64)g + (i64)kTriLanes * pass;  // position in solve order
            if (Kl >= p.N) break;
            const i64 kk = s - (i64)delta * Kl;
            if (kk < 0 || kk >= p.maxr) continue;
            const i64 K = p.upper ? p.N - 1 - Kl : Kl;
            const i64 k = p.upper ? p.maxr - 1 - kk : kk;
            const i64 r0 = p.R[K], rK = p.R[K + 1] - r0;
            if (k >= rK) continue;
            T acc = p.x[r0 + k];
            // off-diagonal blocks: block rows solved earlier, values from their rings (maybe in another CTA)
            for (int d = 1; d <= nd; ++d) {
                const i64 Jl = Kl - d;
                if (Jl < 0) break;
                const i64 J = p.upper ? K + d : K - d;
                const i64 c0 = p.R[J], rJ = p.R[J + 1] - c0;
                const int gj = (int)(Jl % kTriLanes), pj = (int)(Jl / kTriLanes);
                const T* rem = cluster.map_shared_rank(ring, gj / kTriThreads) + ((size_t)pj * kTriThreads + (gj % kTriThreads)) * D;
                const T* slab = data + (i64)(p.u + K - J) * w;
                for (int ii = 0; ii < w; ++ii) {
                    const i64 j = k + p.mu - ii;
                    if (j < 0 || j >= rJ) continue;
                    acc = fma(-slab[ii + (i64)P * (c0 + j)], rem[j & (D - 1)], acc);
                }
            }
            // diagonal block: the part of the row already solved (j < k for lower, j > k for upper)
            T* mine = ring + ((size_t)pass * kTriThreads + tid) * D;
            const T* dslab = data + (i64)p.u * w;
            for (int ii = 0; ii < w; ++ii) {
                const i64 j = k + p.mu - ii;
                if (j < 0 || j >= rK) continue;
                if (p.upper ? j <= k : j >= k) continue;
                acc = fma(-dslab[ii + (i64)P * (r0 + j)], mine[j & (D - 1)], acc);
            }
            const T xk = p.unit ? acc : acc / dslab[p.mu + (i64)P * (r0 + k)];
            mine[k & (D - 1)] = xk;
            p.x[r0 + k] = xk;
        }
        cluster.sync();  // step boundary: this step's ring writes become visible cluster-wide
    }
}

template <typename T>
int32_t bbb_trsv_impl(bbm_handle_t h, bbm_bbb_desc_t d, int32_t uplo, int32_t unit, const T* d_data, T* d_B, i64 ldb, i64 nrhs) {
    BBM_REQUIRE_HANDLE(h);
    BbbDescView v;
    if (!bbm_bbb_desc_view(d, &v)) return bbm_fail(h, BBM_ERR_HANDLE, "invalid BandedBlockBandedMatrix descriptor");
    if (uplo != 'U' && uplo != 'L' && uplo != 'u' && uplo != 'l') return bbm_fail(h, BBM_ERR_ARG, "uplo must be 'U' or 'L'");
    const int upper = (uplo == 'U' || uplo == 'u') ? 1 : 0;
    if (nrhs < 0) return bbm_fail(h, BBM_ERR_ARG, "nrhs < 0");
    // hasmatchingblocks (src/triblockbanded.jl:12,20)
    bool match = v.Nr == v.Nc;
    for (i64 K = 0; match && K < v.Nr; ++K) match = v.r[K] == v.c[K];
    if (!match) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "triangular solve needs matching square diagonal blocks (the reference falls back to a generic solve)");
    if (nrhs > 0 && ldb < std::max<i64>(1, v.m)) return bbm_fail(h, BBM_ERR_DIM, "DimensionMismatch: T is %lld x %lld, ldb=%lld", (long long)v.m, (long long)v.m, (long long)ldb);
    if (v.m == 0 || nrhs == 0) return BBM_OK;
    if (v.P == 0 || v.l < 0 || v.u < 0 || v.lam < 0 || v.mu < 0) return bbm_fail(h, BBM_ERR_BAND, "BandError: the diagonal is outside the bands");
    if (!d_data || !d_B) return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
    BBM_CUDA(h, cudaSetDevice(h->device));
    TriSolveParams<T> p{};
    p.data = d_data; p.R = v.dR; p.N = v.Nr; p.maxr = v.maxr;
    p.l = (int)v.l; p.u = (int)v.u; p.lam = (int)v.lam; p.mu = (int)v.mu; p.w = (int)(v.lam + v.mu + 1); p.P = (int)v.P;
    p.upper = upper; p.unit = unit ? 1 : 0;
    const i64 delta = upper ? v.lam + 1 : v.mu + 1;
    const i64 nd = upper ? v.u : v.l;
    const i64 need = nd * delta + (upper ? v.mu : v.lam) + 2;  // ring depth: the oldest value a reader still needs
    int D = 4;
    while (D < need) D *= 2;
    p.D = D;
    p.passes = (int)((v.Nr + kTriLanes - 1) / kTriLanes);
    p.iters = v.maxr + delta * (v.Nr - 1);
    const size_t smem = (size_t)p.passes * kTriThreads * D * sizeof(T);
    if (smem > std::min<size_t>(h->smem_optin, 200 * 1024))
        return bbm_fail(h, BBM_ERR_UNSUPPORTED, "triangular solve: %lld block rows with bandwidths (%lld,%lld),(%lld,%lld) exceed the shared-memory rings",
                        (long long)v.Nr, (long long)v.l, (long long)v.u, (long long)v.lam, (long long)v.mu);
    cudaError_t e = cudaFuncSetAttribute(bbb_trisolve_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    for (i64 q = 0; q < nrhs; ++q) {
        p.x = d_B + q * ldb;
        bbb_trisolve_kernel<T><<<kTriCtas, kTriThreads, smem, h->stream>>>(p);
        e = cudaGetLastError();
        if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "bbb_trisolve_kernel launch: %s", cudaGetErrorString(e));
        h->launches += 1;
    }
    return BBM_OK;
}

}  // namespace

extern "C" {
int32_t bbm_bbb_trsv_f64(bbm_handle_t h, bbm_bbb_desc_t d, int32_t uplo, int32_t unit, const double* d_data, double* d_B, int64_t ldb,
                         int64_t nrhs) {
    return bbb_trsv_impl<double>(h, d, uplo, unit, d_data, d_B, ldb, nrhs);
}
int32_t bbm_bbb_trsv_f32(bbm_handle_t h, bbm_bbb_desc_t d, int32_t uplo, int32_t unit, const float* d_data, float* d_B, int64_t ldb,
                         int64_t nrhs) {
    return bbb_trsv_impl<float>(h, d, uplo, unit, d_data, d_B, ldb, nrhs);
}
}
