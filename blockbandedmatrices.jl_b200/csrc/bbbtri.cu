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
#include <vector>

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
    int upper, unit, passes, D, nc, relaxed, dbg;
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

// ---- staged variant: coefficients prefetched PF steps ahead ------------------------------------------------
// The direct kernel above pays a DRAM round trip inside every step (the matrix entries of a step are touched
// exactly once).  None of those loads depends on the solution, so this variant runs them PF-1 steps ahead of
// the substitution with 4/8-byte cp.async copies into a per-thread shared-memory ring; block offsets come from
// a shared-memory table instead of global memory.  What is left on the critical path of a step is shared /
// distributed-shared memory reads, a handful of FMAs, one division and the cluster barrier.
//
// shared memory (all lane-contiguous, i.e. [..][tid], so that a warp's accesses are conflict-free):
//   ring [passes][D][kTriThreads]        last D solution values of the thread's block row
//   coef [passes][PF][NC][kTriThreads]   staged step inputs: 0 = right-hand side, 1 = diagonal entry,
//                                        2+(d-1)w+ii = off-diagonal block d, 2+nd*w+t = diagonal block part
//   rs   [passes][kTriThreads+nd] i64    first row of block row Kl-d at index tid+nd-d
//   rn   [passes][kTriThreads+nd] int    its size (0 for Kl-d outside the matrix)
__device__ __forceinline__ void tri_cp_async(void* smem_dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void tri_cp_async(void* smem_dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(src) : "memory");
}

template <typename T, int PF>
__global__ void __cluster_dims__(kTriCtas, 1, 1) __launch_bounds__(kTriThreads) bbb_trisolve_staged_kernel(const TriSolveParams<T> p) {
    extern __shared__ __align__(16) unsigned char tri_ring_smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x;
    const int rank = (int)cluster.block_rank();
    const int g = rank * kTriThreads + tid;
    const int D = p.D, w = p.w, P = p.P, NC = p.nc, mu = p.mu;
    const int delta = p.upper ? p.lam + 1 : p.mu + 1;
    const int nd = p.upper ? p.u : p.l;
    const int ndiag = p.upper ? p.mu : p.lam;        // entries of the diagonal block left/right of the diagonal
    const int ii0 = p.upper ? 0 : p.mu + 1;           // first band row of that part
    const int mw = kTriThreads + nd;
    i64* rs = reinterpret_cast<i64*>(tri_ring_smem);
    T* ring = reinterpret_cast<T*>(rs + (size_t)p.passes * mw);
    T* coef = ring + (size_t)p.passes * D * kTriThreads;
    int* rn = reinterpret_cast<int*>(coef + (size_t)p.passes * PF * NC * kTriThreads);
    const T* __restrict__ data = p.data;

    for (int pass = 0; pass < p.passes; ++pass)
        for (int t = tid; t < mw; t += kTriThreads) {
            const i64 Kl = (i64)rank * kTriThreads + (i64)kTriLanes * pass + t - nd;
            i64 st = 0, sz = 0;
            if (Kl >= 0 && Kl < p.N) {
                const i64 K = p.upper ? p.N - 1 - Kl : Kl;
                st = p.R[K];
                sz = p.R[K + 1] - st;
            }
            rs[pass * mw + t] = st;
            rn[pass * mw + t] = (int)sz;
        }
    __syncthreads();

    // enqueue the loads of step `st` (all passes); one commit group per step
    auto issue = [&](i64 st) {
        for (int pass = 0; pass < p.passes; ++pass) {
            const i64 Kl = (i64)g + (i64)kTriLanes * pass;
            const i64 kk = st - (i64)delta * Kl;
            if (kk < 0 || kk >= p.maxr) continue;
            const int rK = rn[pass * mw + tid + nd];
            const i64 k = p.upper ? p.maxr - 1 - kk : kk;
            if (k >= rK) continue;
            const i64 r0 = rs[pass * mw + tid + nd];
            T* cb = coef + ((size_t)(pass * PF + (int)(st % PF)) * NC) * kTriThreads + tid;
            tri_cp_async(cb, p.x + r0 + k);
            const T* dcol = data + (i64)p.u * w;
            if (!p.unit) tri_cp_async(cb + kTriThreads, dcol + mu + (i64)P * (r0 + k));
            for (int d = 1; d <= nd; ++d) {
                const int rJ = rn[pass * mw + tid + nd - d];
                const i64 c0 = rs[pass * mw + tid + nd - d];
                const T* slab = data + (i64)(p.upper ? p.u - d : p.u + d) * w;
                for (int ii = 0; ii < w; ++ii) {
                    const i64 j = k + mu - ii;
                    if (j < 0 || j >= rJ) continue;
                    tri_cp_async(cb + (size_t)(2 + (d - 1) * w + ii) * kTriThreads, slab + ii + (i64)P * (c0 + j));
                }
            }
            for (int t = 0; t < ndiag; ++t) {
                const int ii = ii0 + t;
                const i64 j = k + mu - ii;
                if (j < 0 || j >= rK) continue;
                tri_cp_async(cb + (size_t)(2 + nd * w + t) * kTriThreads, dcol + ii + (i64)P * (r0 + j));
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    for (int t = 0; t < PF - 1; ++t) issue(t);
    for (i64 s = 0; s < p.iters; ++s) {
        if (!(p.dbg & 2)) issue(s + PF - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(PF - 1) : "memory");
        for (int pass = 0; pass < ((p.dbg & 1) ? 0 : p.passes); ++pass) {
            const i64 Kl = (i64)g + (i64)kTriLanes * pass;
            const i64 kk = s - (i64)delta * Kl;
            if (kk < 0 || kk >= p.maxr) continue;
            const int rK = rn[pass * mw + tid + nd];
            const i64 k = p.upper ? p.maxr - 1 - kk : kk;
            if (k >= rK) continue;
            const T* cb = coef + ((size_t)(pass * PF + (int)(s % PF)) * NC) * kTriThreads + tid;
            T acc = cb[0];
            for (int d = 1; d <= nd; ++d) {
                const int rJ = rn[pass * mw + tid + nd - d];
                if (rJ == 0) continue;
                int gj = g - d, pj = pass;          // owner of block row Kl-d
                if (gj < 0) { gj += kTriLanes; pj -= 1; }
                const T* rem = cluster.map_shared_rank(ring, gj / kTriThreads) + (size_t)pj * D * kTriThreads + (gj % kTriThreads);
                for (int ii = 0; ii < w; ++ii) {
                    const i64 j = k + mu - ii;
                    if (j < 0 || j >= rJ) continue;
                    acc = fma(-cb[(size_t)(2 + (d - 1) * w + ii) * kTriThreads], rem[(size_t)(j & (D - 1)) * kTriThreads], acc);
                }
            }
            T* mine = ring + (size_t)pass * D * kTriThreads + tid;
            for (int t = 0; t < ndiag; ++t) {
                const i64 j = k + mu - (ii0 + t);
                if (j < 0 || j >= rK) continue;
                acc = fma(-cb[(size_t)(2 + nd * w + t) * kTriThreads], mine[(size_t)(j & (D - 1)) * kTriThreads], acc);
            }
            const T xk = p.unit ? acc : acc / cb[kTriThreads];
            mine[(size_t)(k & (D - 1)) * kTriThreads] = xk;
            p.x[rs[pass * mw + tid + nd] + k] = xk;
        }
        if (p.relaxed) {
            // Step boundary without the cluster-scope release: that fence (MEMBAR + ERRBAR in SASS) also waits for
            // the global store of x and for every prefetch copy still in flight, i.e. it puts the DRAM round trip
            // back into each step.  Only the shared-memory ring has to be ordered: a CTA-scope fence completes this
            // thread's ring store in its own SM's shared memory -- the single copy that local and DSMEM readers
            // both access -- before the arrive; readers pass the wait (acquire) only after every thread arrived.
            asm volatile("fence.acq_rel.cta;" ::: "memory");
            asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        } else {
            cluster.sync();
        }
    }
    // nothing may leave while a neighbour could still read its ring
    cluster.sync();
}

template <typename T, int PF>
cudaError_t launch_staged(const TriSolveParams<T>& p, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(bbb_trisolve_staged_kernel<T, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    bbb_trisolve_staged_kernel<T, PF><<<kTriCtas, kTriThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

// ---- wave-ordered variant -------------------------------------------------------------------------------------
// In the staged kernel a warp's 32 threads own 32 different block rows, so each of its prefetch copies touches
// 32 different cache lines: ~6 x 512 line requests per SM and step, which is what bounds it (2 us per step, the
// L1 tag stage handles about one line per cycle).  Here a pre-pass running on all SMs gathers the right-hand
// side and every matrix entry the substitution will use into WAVE ORDER -- plane c, position
// base[s] + Kl - Klmin(s) for the unknown that block row Kl (in solve order) computes at step s -- so that the
// solve kernel's warps read (and write the solution) with unit stride; a post-pass scatters the solution back.
//   Klmin(s) = first block row still active at step s = max(0, floor((s - maxr)/delta) + 1)
//   planes: 0 right-hand side / solution, 1 diagonal entry, 2+(d-1)w+ii off-diagonal block d band row ii,
//           2+nd*w+t diagonal-block band row ii0+t; entries outside their block are stored as 0.
template <typename T>
struct TriWaveParams {
    const T* data;
    T* x;
    T* packed;         // [nc][total]
    const i64* base;   // [iters+1] wave offsets
    const i64* R;
    i64 N, maxr, iters, total;
    int l, u, lam, mu, w, P;
    int upper, unit, passes, D, nc, relaxed, what;
};

__device__ __forceinline__ i64 tri_klmin(i64 s, i64 maxr, int delta) { return s < maxr ? 0 : (s - maxr) / delta + 1; }

// what: bit 0 = right-hand side plane, bit 1 = matrix planes
template <typename T>
__global__ void __launch_bounds__(256) bbb_tri_pack_kernel(const TriWaveParams<T> p) {
    const i64 idx = (i64)blockIdx.x * 256 + threadIdx.x;
    if (idx >= p.N * p.maxr) return;
    const i64 Kl = idx / p.maxr, kk = idx % p.maxr;
    const i64 K = p.upper ? p.N - 1 - Kl : Kl;
    const i64 k = p.upper ? p.maxr - 1 - kk : kk;
    const i64 r0 = p.R[K], rK = p.R[K + 1] - r0;
    if (k >= rK) return;
    const int delta = p.upper ? p.lam + 1 : p.mu + 1;
    const int nd = p.upper ? p.u : p.l;
    const i64 s = kk + (i64)delta * Kl;
    T* out = p.packed + p.base[s] + Kl - tri_klmin(s, p.maxr, delta);
    if (p.what & 1) out[0] = p.x[r0 + k];
    if (!(p.what & 2)) return;
    const int w = p.w, mu = p.mu;
    const T* dcol = p.data + (i64)p.u * w;
    out[p.total] = p.unit ? T(1) : dcol[mu + (i64)p.P * (r0 + k)];
    for (int d = 1; d <= nd; ++d) {
        const i64 J = p.upper ? K + d : K - d;
        const bool have = J >= 0 && J < p.N;
        const i64 c0 = have ? p.R[J] : 0, rJ = have ? p.R[J + 1] - c0 : 0;
        const T* slab = p.data + (i64)(p.upper ? p.u - d : p.u + d) * w;
        for (int ii = 0; ii < w; ++ii) {
            const i64 j = k + mu - ii;
            out[(i64)(2 + (d - 1) * w + ii) * p.total] = (j >= 0 && j < rJ) ? slab[ii + (i64)p.P * (c0 + j)] : T(0);
        }
    }
    const int ndiag = p.upper ? p.mu : p.lam, ii0 = p.upper ? 0 : p.mu + 1;
    for (int t = 0; t < ndiag; ++t) {
        const i64 j = k + mu - (ii0 + t);
        out[(i64)(2 + nd * w + t) * p.total] = (j >= 0 && j < rK) ? dcol[ii0 + t + (i64)p.P * (r0 + j)] : T(0);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) bbb_tri_unpack_kernel(const TriWaveParams<T> p) {
    const i64 idx = (i64)blockIdx.x * 256 + threadIdx.x;
    if (idx >= p.N * p.maxr) return;
    const i64 Kl = idx / p.maxr, kk = idx % p.maxr;
    const i64 K = p.upper ? p.N - 1 - Kl : Kl;
    const i64 k = p.upper ? p.maxr - 1 - kk : kk;
    const i64 r0 = p.R[K], rK = p.R[K + 1] - r0;
    if (k >= rK) return;
    const int delta = p.upper ? p.lam + 1 : p.mu + 1;
    const i64 s = kk + (i64)delta * Kl;
    p.x[r0 + k] = p.packed[p.base[s] + Kl - tri_klmin(s, p.maxr, delta)];
}

// shared memory: pos [passes][PF][kTriThreads] i64 | ring [passes][D][kTriThreads] | coef [passes][PF][NC][kTriThreads]
//                | rn [passes][kTriThreads+nd] int (block sizes of block rows Kl-nd .. Kl)
template <typename T, int PF>
__global__ void __cluster_dims__(kTriCtas, 1, 1) __launch_bounds__(kTriThreads) bbb_trisolve_wave_kernel(const TriWaveParams<T> p) {
    extern __shared__ __align__(16) unsigned char tri_ring_smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x;
    const int rank = (int)cluster.block_rank();
    const int g = rank * kTriThreads + tid;
    const int D = p.D, w = p.w, NC = p.nc, mu = p.mu;
    const int delta = p.upper ? p.lam + 1 : p.mu + 1;
    const int nd = p.upper ? p.u : p.l;
    const int ndiag = p.upper ? p.mu : p.lam;
    const int ii0 = p.upper ? 0 : p.mu + 1;
    const int mw = kTriThreads + nd;
    i64* posr = reinterpret_cast<i64*>(tri_ring_smem);
    T* ring = reinterpret_cast<T*>(posr + (size_t)p.passes * PF * kTriThreads);
    T* coef = ring + (size_t)p.passes * D * kTriThreads;
    int* rn = reinterpret_cast<int*>(coef + (size_t)p.passes * PF * NC * kTriThreads);

    for (int pass = 0; pass < p.passes; ++pass) {
        for (int t = tid; t < mw; t += kTriThreads) {
            const i64 Kl = (i64)rank * kTriThreads + (i64)kTriLanes * pass + t - nd;
            i64 sz = 0;
            if (Kl >= 0 && Kl < p.N) {
                const i64 K = p.upper ? p.N - 1 - Kl : Kl;
                sz = p.R[K + 1] - p.R[K];
            }
            rn[pass * mw + t] = (int)sz;
        }
        for (int t = tid; t < D * kTriThreads; t += kTriThreads) ring[(size_t)pass * D * kTriThreads + t] = T(0);
    }
    cluster.sync();  // every ring is zeroed before a neighbour CTA may read it

    auto issue = [&](i64 st, i64 bs) {  // bs = base[st]
        if (st < p.iters) {
            const i64 kmin = tri_klmin(st, p.maxr, delta);
            for (int pass = 0; pass < p.passes; ++pass) {
                const i64 Kl = (i64)g + (i64)kTriLanes * pass;
                const i64 kk = st - (i64)delta * Kl;
                if (kk < 0 || kk >= p.maxr) continue;
                const int rK = rn[pass * mw + tid + nd];
                const i64 k = p.upper ? p.maxr - 1 - kk : kk;
                if (k >= rK) continue;
                const i64 pos = bs + Kl - kmin;
                const int slot = pass * PF + (int)(st % PF);
                posr[(size_t)slot * kTriThreads + tid] = pos;
                T* cb = coef + (size_t)slot * NC * kTriThreads + tid;
                const T* src = p.packed + pos;
                for (int c = 0; c < NC; ++c) tri_cp_async(cb + (size_t)c * kTriThreads, src + (i64)c * p.total);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    for (int t = 0; t < PF - 1; ++t) issue(t, t < p.iters ? p.base[t] : 0);
    i64 bnext = (PF - 1 < p.iters) ? p.base[PF - 1] : 0;  // offset of the step issued next, loaded one step early
    for (i64 s = 0; s < p.iters; ++s) {
        const i64 bcur = bnext;
        if (s + PF < p.iters) bnext = p.base[s + PF];
        issue(s + PF - 1, bcur);
        asm volatile("cp.async.wait_group %0;" ::"n"(PF - 1) : "memory");
        for (int pass = 0; pass < p.passes; ++pass) {
            const i64 Kl = (i64)g + (i64)kTriLanes * pass;
            const i64 kk = s - (i64)delta * Kl;
            if (kk < 0 || kk >= p.maxr) continue;
            const int rK = rn[pass * mw + tid + nd];
            const int k = (int)(p.upper ? p.maxr - 1 - kk : kk);
            if (k >= rK) continue;
            const int slot = pass * PF + (int)(s % PF);
            const T* cb = coef + (size_t)slot * NC * kTriThreads + tid;
            T acc = cb[0];
            for (int d = 1; d <= nd; ++d) {
                if (rn[pass * mw + tid + nd - d] == 0) continue;   // no such block row, or an empty one
                int gj = g - d, pj = pass;                          // owner of block row Kl-d
                if (gj < 0) { gj += kTriLanes; pj -= 1; }
                const T* rem = cluster.map_shared_rank(ring, gj / kTriThreads) + (size_t)pj * D * kTriThreads + (gj % kTriThreads);
                const T* cd = cb + (size_t)(2 + (d - 1) * w) * kTriThreads;
                for (int ii = 0; ii < w; ++ii)   // entries outside the block were packed as zeros
                    acc = fma(-cd[(size_t)ii * kTriThreads], rem[(size_t)((k + mu - ii) & (D - 1)) * kTriThreads], acc);
            }
            T* mine = ring + (size_t)pass * D * kTriThreads + tid;
            const T* cd = cb + (size_t)(2 + nd * w) * kTriThreads;
            for (int t = 0; t < ndiag; ++t)
                acc = fma(-cd[(size_t)t * kTriThreads], mine[(size_t)((k + mu - ii0 - t) & (D - 1)) * kTriThreads], acc);
            const T xk = p.unit ? acc : acc / cb[kTriThreads];
            mine[(size_t)(k & (D - 1)) * kTriThreads] = xk;
            p.packed[posr[(size_t)slot * kTriThreads + tid]] = xk;
        }
        if (p.relaxed) {
            asm volatile("fence.acq_rel.cta;" ::: "memory");
            asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        } else {
            cluster.sync();
        }
    }
    cluster.sync();
}

template <typename T, int PF>
cudaError_t launch_wave(const TriWaveParams<T>& p, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(bbb_trisolve_wave_kernel<T, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    bbb_trisolve_wave_kernel<T, PF><<<kTriCtas, kTriThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

// Returns BBM_OK when the solve was enqueued, 1 when this variant does not apply (caller falls back).
template <typename T>
int32_t bbb_trsv_wave(bbm_handle_t h, const BbbDescView& v, const TriSolveParams<T>& sp, T* d_B, i64 ldb, i64 nrhs) {
    const int upper = sp.upper;
    const i64 delta = upper ? v.lam + 1 : v.mu + 1;
    const i64 nd = upper ? v.u : v.l;
    TriWaveParams<T> p{};
    p.data = sp.data; p.R = sp.R; p.N = sp.N; p.maxr = sp.maxr; p.iters = sp.iters;
    p.l = sp.l; p.u = sp.u; p.lam = sp.lam; p.mu = sp.mu; p.w = sp.w; p.P = sp.P;
    p.upper = upper; p.unit = sp.unit; p.passes = sp.passes; p.D = sp.D; p.nc = sp.nc; p.relaxed = sp.relaxed;
    const size_t budget = std::min<size_t>(h->smem_optin, 200 * 1024);
    const size_t fixed = (size_t)p.passes * (kTriThreads * p.D * sizeof(T) + (kTriThreads + nd) * sizeof(int)) + 16;
    int pf = 0;
    const int want_pf = (int)bbm_opt(h, "tri.pf", 4);
    for (int cand : {8, 6, 4, 2})
        if (cand <= want_pf && pf == 0 && fixed + (size_t)p.passes * cand * kTriThreads * (p.nc * sizeof(T) + sizeof(i64)) <= budget) pf = cand;
    if (pf == 0) return 1;
    // wave offsets (host, cached in the handle)
    const size_t tbytes = (size_t)(p.iters + 1) * sizeof(i64);
    i64 total = 0;
    {
        std::vector<i64> base((size_t)p.iters + 1);
        for (i64 s = 0; s < p.iters; ++s) {
            base[(size_t)s] = total;
            const i64 kmax = std::min<i64>(p.N - 1, s / delta);
            const i64 kmin = s < p.maxr ? 0 : (s - p.maxr) / delta + 1;
            total += kmax >= kmin ? kmax - kmin + 1 : 0;
        }
        base[(size_t)p.iters] = total;
        const size_t need = (size_t)total * p.nc * sizeof(T);
        if (need > (size_t)bbm_opt(h, "tri.ws_max_mb", 16384) * 1048576) return 1;
        if (h->tri_base_key[0] != p.N || h->tri_base_key[1] != p.maxr || h->tri_base_key[2] != delta) {
            h->tri_base_key[0] = -1;
            if (bbm_ws_reserve(h, &h->tri_base, &h->tri_base_bytes, tbytes) != BBM_OK) return 1;
            BBM_CUDA(h, cudaMemcpyAsync(h->tri_base, base.data(), tbytes, cudaMemcpyHostToDevice, h->stream));
            BBM_CUDA(h, cudaStreamSynchronize(h->stream));  // `base` is pageable and dies with this scope
            h->tri_base_key[0] = p.N; h->tri_base_key[1] = p.maxr; h->tri_base_key[2] = delta;
        }
        if (bbm_ws_reserve(h, &h->ws_w, &h->ws_w_bytes, need) != BBM_OK) return 1;
    }
    p.base = static_cast<const i64*>(h->tri_base);
    p.packed = static_cast<T*>(h->ws_w);
    p.total = total;
    const size_t smem = fixed + (size_t)p.passes * pf * kTriThreads * (p.nc * sizeof(T) + sizeof(i64));
    const i64 cells = p.N * p.maxr;
    const unsigned grid = (unsigned)((cells + 255) / 256);
    for (i64 q = 0; q < nrhs; ++q) {
        p.x = d_B + q * ldb;
        p.what = q == 0 ? 3 : 1;
        bbb_tri_pack_kernel<T><<<grid, 256, 0, h->stream>>>(p);
        cudaError_t e = pf == 8 ? launch_wave<T, 8>(p, smem, h->stream) : pf == 6 ? launch_wave<T, 6>(p, smem, h->stream)
                      : pf == 4 ? launch_wave<T, 4>(p, smem, h->stream) : launch_wave<T, 2>(p, smem, h->stream);
        if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "bbb_trisolve_wave_kernel launch: %s", cudaGetErrorString(e));
        bbb_tri_unpack_kernel<T><<<grid, 256, 0, h->stream>>>(p);
        e = cudaGetLastError();
        if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "bbb_tri_unpack_kernel launch: %s", cudaGetErrorString(e));
        h->launches += 3;
    }
    return BBM_OK;
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
    const size_t budget = std::min<size_t>(h->smem_optin, 200 * 1024);
    // staged variant: deepest prefetch ring (6, 4 or 2 steps) that fits next to the solution rings
    p.nc = 2 + (int)(nd * (v.lam + v.mu + 1)) + (int)(upper ? v.mu : v.lam);
    const size_t meta = (size_t)p.passes * (kTriThreads + nd) * (sizeof(i64) + sizeof(int)) + 16;
    p.relaxed = bbm_opt(h, "tri.sync", 1) != 0 ? 1 : 0;
    p.dbg = (int)bbm_opt(h, "tri.dbg", 0);  // timing experiments only (bit 0: no substitution, bit 1: no prefetch)
    const int which = (int)bbm_opt(h, "tri.kernel", 0);  // 0 = wave-ordered, 1 = staged, 2 = direct
    if (which == 0) {
        const int32_t rc = bbb_trsv_wave<T>(h, v, p, d_B, ldb, nrhs);
        if (rc != 1) return rc;
    }
    int pf = 0;
    const int want_pf = which == 2 ? 0 : (int)bbm_opt(h, "tri.pf", 4);

    for (int cand : {6, 4, 2})
        if (cand <= want_pf && pf == 0 && smem + meta + (size_t)p.passes * cand * p.nc * kTriThreads * sizeof(T) <= budget) pf = cand;
    if (pf > 0) {
        const size_t total = smem + meta + (size_t)p.passes * pf * p.nc * kTriThreads * sizeof(T);
        for (i64 q = 0; q < nrhs; ++q) {
            p.x = d_B + q * ldb;
            cudaError_t e = pf == 6 ? launch_staged<T, 6>(p, total, h->stream) : pf == 4 ? launch_staged<T, 4>(p, total, h->stream) : launch_staged<T, 2>(p, total, h->stream);
            if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "bbb_trisolve_staged_kernel launch: %s", cudaGetErrorString(e));
            h->launches += 1;
        }
        return BBM_OK;
    }
    if (smem > budget)
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
