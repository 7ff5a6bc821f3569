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
// descending, delta = lam+1).  n + delta*N steps in total (12 286 for the 16.7M-unknown cfg2 Laplacian).
//
// Four kernels, fastest first; the entry point takes the first one that applies:
//   lean   wave-ordered inputs, one block row per thread, compile-time bandwidths ((1,1),(1,1) and (2,2),(2,2)),
//          group-boundary values pushed with st.async                      0.40 us/step   6.75 ms on cfg2 (CPU 154 ms)
//   wave   wave-ordered inputs, any bandwidths / up to 6 block rows per thread          1.3 us/step
//   staged band storage read in place, prefetched with cp.async (no workspace)          3.0 us/step
//   direct band storage read in place inside the step                                   4.4 us/step
#include <cooperative_groups.h>

#include <algorithm>

#include "bbm_internal.h"
#include "ptx.cuh"

namespace cg = cooperative_groups;
using namespace ptx;

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
    int upper, unit, passes, D, nc, relaxed;
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
        issue(s + PF - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(PF - 1) : "memory");
        for (int pass = 0; pass < p.passes; ++pass) {
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
// side and every matrix entry the substitution will use into WAVE ORDER -- plane c, position s*Wd + Kl mod Wd
// for the unknown that block row Kl (in solve order) computes at step s; the block rows active in one step are
// fewer than Wd consecutive ones, so the positions are unique -- and the solve kernel's warps read their inputs
// and write the solution with unit stride; a post-pass scatters the solution back.
//   planes: 0 right-hand side / solution, 1 diagonal entry, 2+(d-1)w+ii off-diagonal block d band row ii,
//           2+nd*w+t diagonal-block band row ii0+t; entries outside their block are stored as 0.
// Only maxr/delta block rows are active at a time (a window sliding over the lanes), so consecutive groups of
// 128 lanes are dealt round-robin to the 8 CTAs: all 8 SMs stay busy instead of the window's 4.
template <typename T>
struct TriWaveParams {
    const T* data;
    T* x;
    T* packed;         // [nc][total]
    const i64* R;
    i64 total, Wd;     // plane stride = iters*Wd
    int N, maxr, iters;
    int l, u, lam, mu, w, P;
    int upper, unit, passes, D, nc, relaxed, what, push;
    int* status;       // [right-hand side] set to 1 if a halo value never arrived (lean kernel; the result is then invalid)
    // Right-hand-side planes live apart from the matrix planes (plane 0 of `packed` is unused): the plane of right-hand
    // side q is packed + xoff + q*total.  Pack / unpack take q from blockIdx.z, the solve kernels from their cluster's
    // index -- several right-hand sides are solved concurrently, one 8-CTA cluster each -- and a prepared plan keeps
    // its matrix planes in a buffer of its own across calls (bbm_bbb_trsv_prepare_*).
    i64 xoff, ldb;
};

constexpr int kTriGroup = 128;                      // lanes dealt to one CTA at a time
constexpr int kTriGroups = kTriThreads / kTriGroup;

// Pack / unpack run on all SMs as tiled transposes between the two orders: a tile is 32 block rows (Kl) x 32
// steps (s).  On the band-storage side a warp handles one block row and its lanes 32 consecutive steps =
// 32 consecutive columns of `data` / entries of x (unit stride); on the wave-ordered side a warp handles one step
// and its lanes 32 consecutive block rows (unit stride again); the tile turns in shared memory.
// what: bit 0 = right-hand side plane, bit 1 = matrix planes.  Planes are staged kTriPackPlanes at a time.
constexpr int kTriPackPlanes = 4;

template <typename T>
__global__ void __launch_bounds__(256) bbb_tri_pack_kernel(const TriWaveParams<T> p) {
    __shared__ T stage[kTriPackPlanes][32][33];
    __shared__ unsigned char ok[32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int delta = p.upper ? p.lam + 1 : p.mu + 1;
    const int nd = p.upper ? p.u : p.l;
    const int w = p.w, mu = p.mu;
    const int ndiag = p.upper ? p.mu : p.lam, ii0 = p.upper ? 0 : p.mu + 1;
    const int Kl0 = blockIdx.y * 32;
    const i64 s0 = (i64)delta * Kl0 + (i64)blockIdx.x * 32;
    const int c_lo = (p.what & 1) ? 0 : 1, c_hi = (p.what & 2) ? p.nc : 1;
    for (int cg0 = c_lo; cg0 < c_hi; cg0 += kTriPackPlanes) {
        const int ncg = min(kTriPackPlanes, c_hi - cg0);
        // gather: warp = block row, lanes = consecutive steps
        for (int a = warp; a < 32; a += 8) {
            const int Kl = Kl0 + a;
            const i64 s = s0 + lane;
            const i64 kk64 = s - (i64)delta * Kl;
            bool valid = Kl < p.N && kk64 >= 0 && kk64 < p.maxr;
            int K = 0, k = 0;
            i64 r0 = 0, rK = 0;
            if (valid) {
                K = p.upper ? p.N - 1 - Kl : Kl;
                k = p.upper ? p.maxr - 1 - (int)kk64 : (int)kk64;
                r0 = p.R[K]; rK = p.R[K + 1] - r0;
                valid = k < rK;
            }
            if (cg0 == c_lo) ok[a][lane] = valid ? 1 : 0;
            if (!valid) continue;
            const T* dcol = p.data + (i64)p.u * w;
            for (int cc = 0; cc < ncg; ++cc) {
                const int c = cg0 + cc;
                T v;
                if (c == 0) {
                    v = p.x[(i64)blockIdx.z * p.ldb + r0 + k];
                } else if (c == 1) {
                    v = p.unit ? T(1) : dcol[mu + (i64)p.P * (r0 + k)];
                } else if (c < 2 + nd * w) {
                    const int d = 1 + (c - 2) / w, ii = (c - 2) % w;
                    const int J = p.upper ? K + d : K - d;
                    v = T(0);
                    if (J >= 0 && J < p.N) {
                        const i64 c0 = p.R[J], rJ = p.R[J + 1] - c0;
                        const i64 j = (i64)k + mu - ii;
                        if (j >= 0 && j < rJ) v = p.data[(i64)(p.upper ? p.u - d : p.u + d) * w + ii + (i64)p.P * (c0 + j)];
                    }
                } else {
                    const int t = c - 2 - nd * w;
                    const i64 j = (i64)k + mu - (ii0 + t);
                    v = (t < ndiag && j >= 0 && j < rK) ? dcol[ii0 + t + (i64)p.P * (r0 + j)] : T(0);
                }
                stage[cc][a][lane] = v;
            }
        }
        __syncthreads();
        // scatter: warp = step, lanes = consecutive block rows
        for (int b = warp; b < 32; b += 8) {
            const int Kl = Kl0 + lane;
            if (ok[lane][b]) {
                T* out = p.packed + (s0 + b) * p.Wd + Kl % p.Wd;
                for (int cc = 0; cc < ncg; ++cc) {
                    const int c = cg0 + cc;
                    out[c == 0 ? p.xoff + (i64)blockIdx.z * p.total : (i64)c * p.total] = stage[cc][lane][b];
                }
            }
        }
        __syncthreads();
    }
}

template <typename T>
__global__ void __launch_bounds__(256) bbb_tri_unpack_kernel(const TriWaveParams<T> p) {
    __shared__ T stage[32][33];
    __shared__ unsigned char ok[32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int delta = p.upper ? p.lam + 1 : p.mu + 1;
    const int Kl0 = blockIdx.y * 32;
    const i64 s0 = (i64)delta * Kl0 + (i64)blockIdx.x * 32;
    const bool poison = p.status[blockIdx.z] != 0;   // a halo value never arrived inside the solve kernel: NaN instead of a wrong answer
    const T* __restrict__ xplane = p.packed + p.xoff + (i64)blockIdx.z * p.total;
    // wave-ordered side: warp = step, lanes = consecutive block rows
    for (int b = warp; b < 32; b += 8) {
        const int Kl = Kl0 + lane;
        const i64 s = s0 + b;
        const i64 kk = s - (i64)delta * Kl;
        const bool in = Kl < p.N && kk >= 0 && kk < p.maxr && s < p.iters;
        ok[lane][b] = in ? 1 : 0;
        if (in) stage[lane][b] = xplane[s * p.Wd + Kl % p.Wd];
    }
    __syncthreads();
    // x side: warp = block row, lanes = consecutive steps = consecutive entries of x
    for (int a = warp; a < 32; a += 8) {
        if (!ok[a][lane]) continue;
        const int Kl = Kl0 + a;
        const i64 kk = s0 + lane - (i64)delta * Kl;
        const int K = p.upper ? p.N - 1 - Kl : Kl;
        const int k = p.upper ? p.maxr - 1 - (int)kk : (int)kk;
        const i64 r0 = p.R[K], rK = p.R[K + 1] - r0;
        if (k >= rK) continue;
        p.x[(i64)blockIdx.z * p.ldb + r0 + k] = poison ? T(NAN) : stage[a][lane];
    }
}

// shared memory: ring [passes][D][kTriThreads] | coef [passes][PF][NC][kTriThreads]
//                | rn [passes][kTriGroups][kTriGroup+nd] int: block sizes of block rows Kl-nd..Kl of each lane group
//                | klm [passes][kTriThreads] int: Kl mod Wd
template <typename T, int PF>
__global__ void __cluster_dims__(kTriCtas, 1, 1) __launch_bounds__(kTriThreads) bbb_trisolve_wave_kernel(const TriWaveParams<T> p) {
    static_assert((PF & (PF - 1)) == 0, "PF must be a power of two");
    extern __shared__ __align__(16) unsigned char tri_ring_smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x;
    const int rank = (int)cluster.block_rank();
    const int grp = tid / kTriGroup, lid = tid % kTriGroup;
    const int g = (grp * kTriCtas + rank) * kTriGroup + lid;   // lane in solve order (block row g + kTriLanes*pass)
    const i64 xq = p.xoff + (i64)(blockIdx.x / kTriCtas) * p.total;   // this cluster's right-hand-side plane
    const int D = p.D, w = p.w, NC = p.nc, mu = p.mu, maxr = p.maxr, iters = p.iters;
    const int delta = p.upper ? p.lam + 1 : p.mu + 1;
    const int nd = p.upper ? p.u : p.l;
    const int ndiag = p.upper ? p.mu : p.lam;
    const int ii0 = p.upper ? 0 : p.mu + 1;
    const int gw = kTriGroup + nd;
    T* ring = reinterpret_cast<T*>(tri_ring_smem);
    T* coef = ring + (size_t)p.passes * D * kTriThreads;
    int* rn = reinterpret_cast<int*>(coef + (size_t)p.passes * PF * NC * kTriThreads);
    int* klm = rn + p.passes * kTriGroups * gw;

    for (int pass = 0; pass < p.passes; ++pass) {
        for (int t = tid; t < kTriGroups * gw; t += kTriThreads) {
            const int gq = t / gw, o = t % gw;
            const i64 Kl = (i64)(gq * kTriCtas + rank) * kTriGroup + o - nd + (i64)kTriLanes * pass;
            int sz = 0;
            if (Kl >= 0 && Kl < p.N) {
                const i64 K = p.upper ? p.N - 1 - Kl : Kl;
                sz = (int)(p.R[K + 1] - p.R[K]);
            }
            rn[pass * kTriGroups * gw + t] = sz;
        }
        klm[pass * kTriThreads + tid] = (int)(((i64)g + (i64)kTriLanes * pass) % p.Wd);
        for (int t = tid; t < D * kTriThreads; t += kTriThreads) ring[(size_t)pass * D * kTriThreads + t] = T(0);
    }
    cluster.sync();  // every ring is zeroed before a neighbour CTA may read it

    const int* rnme = rn + grp * gw + lid + nd;   // + pass*kTriGroups*gw; [-d] = block row Kl-d
    // ring of the lane d block rows back (it may live in another CTA: distributed shared memory)
    auto ring_of = [&](int pass, int d) -> const T* {
        int gj = g - d, pj = pass;
        if (gj < 0) { gj += kTriLanes; pj -= 1; }
        if (pj < 0) return ring;  // no such block row (never dereferenced: its size is 0)
        const int G = gj / kTriGroup;
        return cluster.map_shared_rank(ring, G % kTriCtas) + (size_t)pj * D * kTriThreads + (G / kTriCtas) * kTriGroup + gj % kTriGroup;
    };
    const T* const rem1 = ring_of(0, 1);

    auto issue = [&](int st) {
        if (st < iters) {
            for (int pass = 0; pass < p.passes; ++pass) {
                const int kk = st - delta * (g + kTriLanes * pass);
                if ((unsigned)kk >= (unsigned)maxr) continue;
                const int k = p.upper ? maxr - 1 - kk : kk;
                if (k >= rnme[pass * kTriGroups * gw]) continue;
                const T* src = p.packed + ((i64)st * p.Wd + klm[pass * kTriThreads + tid]);
                T* cb = coef + (size_t)((pass * PF + (st & (PF - 1))) * NC) * kTriThreads + tid;
                tri_cp_async(cb, src + xq);
                if (!p.unit) tri_cp_async(cb + kTriThreads, src + p.total);
                for (int c = 2; c < NC; ++c) tri_cp_async(cb + (size_t)c * kTriThreads, src + (i64)c * p.total);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    for (int t = 0; t < PF - 1; ++t) issue(t);
    for (int s = 0; s < iters; ++s) {
        issue(s + PF - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(PF - 1) : "memory");
        for (int pass = 0; pass < p.passes; ++pass) {
            const int kk = s - delta * (g + kTriLanes * pass);
            if ((unsigned)kk >= (unsigned)maxr) continue;
            const int k = p.upper ? maxr - 1 - kk : kk;
            const int* rnp = rnme + pass * kTriGroups * gw;
            if (k >= rnp[0]) continue;
            const T* cb = coef + (size_t)((pass * PF + (s & (PF - 1))) * NC) * kTriThreads + tid;
            T acc = cb[0];
            const int km = k + mu;
            for (int d = 1; d <= nd; ++d) {
                if (rnp[-d] == 0) continue;   // no such block row, or an empty one
                const T* rem = (d == 1 && pass == 0) ? rem1 : ring_of(pass, d);
                const T* cd = cb + (size_t)(2 + (d - 1) * w) * kTriThreads;
                for (int ii = 0; ii < w; ++ii)   // entries outside the block were packed as zeros
                    acc = fma(-cd[ii * kTriThreads], rem[((km - ii) & (D - 1)) * kTriThreads], acc);
            }
            T* mine = ring + (size_t)pass * D * kTriThreads + tid;
            const T* cd = cb + (size_t)(2 + nd * w) * kTriThreads;
            for (int t = 0; t < ndiag; ++t)
                acc = fma(-cd[t * kTriThreads], mine[((km - ii0 - t) & (D - 1)) * kTriThreads], acc);
            const T xk = p.unit ? acc : acc / cb[kTriThreads];
            mine[(k & (D - 1)) * kTriThreads] = xk;
            p.packed[xq + (i64)s * p.Wd + klm[pass * kTriThreads + tid]] = xk;
        }
        if (p.relaxed) {
            // Step boundary without the cluster-scope release: that fence (MEMBAR.ALL.GPU + ERRBAR in SASS) also
            // waits for the global store of x and for every prefetch copy still in flight.  Only the shared-memory
            // ring has to be ordered: a CTA-scope fence completes this thread's ring store in its own SM's shared
            // memory -- the single copy that local and DSMEM readers both access -- before the arrive; readers
            // pass the wait (acquire) only after every thread of the cluster arrived.
            asm volatile("fence.acq_rel.cta;" ::: "memory");
            asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        } else {
            cluster.sync();
        }
    }
    cluster.sync();  // nothing may leave while a neighbour could still read its ring
}

// Lean instantiation of the wave kernel for at most kTriLanes block rows with compile-time block bandwidth ND
// (towards the already solved side), sub-block band width W and NDIAG diagonal-block entries beside the diagonal.
// A step is latency-bound (every warp runs the same short dependent instruction stream, then the cluster
// barrier), so everything that does not need the neighbours' newest values -- prefetch issue, coefficient loads
// into registers, the reciprocal of the diagonal entry, the ring addresses -- sits between the barrier's arrive
// and its wait, where it overlaps the barrier's own latency; the dependent chain of a step is then: barrier
// release -> ND*W + NDIAG ring loads -> as many FMAs -> quotient correction -> ring store -> arrive.  The
// solution's global store is issued after the arrive.
//
// Lane groups of neighbouring CTAs meet at every 128th block row.  Reading the neighbour's ring there through
// distributed shared memory costs ~600 cycles inside the dependent chain of EVERY step (the slowest warp sets
// the pace: measured 0.87 -> 0.55 us per step with those reads made local).  So the producer pushes instead:
// the last ND lanes of a group send each new value with st.async into a halo ring in the consumer's CTA, which
// completes a transaction on an mbarrier there; the first warp of the consumer group arms the mbarrier with
// expect_tx before the cluster barrier and checks it after -- by then the value has normally landed -- and every
// ring load of every lane is a local shared-memory load.
__device__ __forceinline__ uint32_t tri_mapa(uint32_t smem_addr, int cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void tri_st_async(uint32_t remote_addr, double v, uint32_t remote_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr),
                 "l"(__double_as_longlong(v)), "r"(remote_mbar)
                 : "memory");
}
__device__ __forceinline__ void tri_st_async(uint32_t remote_addr, float v, uint32_t remote_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
                 "r"(__float_as_uint(v)), "r"(remote_mbar)
                 : "memory");
}

template <typename T, int PF, int ND, int W, int NDIAG>
__global__ void __cluster_dims__(kTriCtas, 1, 1) __launch_bounds__(kTriThreads) bbb_trisolve_lean_kernel(const TriWaveParams<T> p) {
    static_assert((PF & (PF - 1)) == 0, "PF must be a power of two");
    static_assert(ND >= 1 && ND <= 32, "halo bookkeeping is done by the first warp of a lane group");
    constexpr int NC = 2 + ND * W + NDIAG;
    extern __shared__ __align__(16) unsigned char tri_ring_smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x;
    const int rank = (int)cluster.block_rank();
    const int grp = tid / kTriGroup, lid = tid % kTriGroup;
    const int G = grp * kTriCtas + rank;          // lane group in solve order
    const int g = G * kTriGroup + lid;            // block row in solve order
    const int Dm = p.D - 1, mu = p.mu, iters = p.iters;
    const bool unit = p.unit != 0, push = p.push != 0;
    const int delta = p.upper ? p.lam + 1 : p.mu + 1;
    const int ii0 = p.upper ? 0 : p.mu + 1;
    const i64 tot = p.total, Wd = p.Wd;
    const int rhs = blockIdx.x / kTriCtas;                    // this cluster's right-hand side
    const i64 xq = p.xoff + (i64)rhs * tot;
    T* ring = reinterpret_cast<T*>(tri_ring_smem);            // [D][kTriThreads]
    T* coef = ring + (size_t)p.D * kTriThreads;               // [PF][NC][kTriThreads]
    T* halo = coef + (size_t)PF * NC * kTriThreads;           // [kTriGroups][ND][D]: rings of the previous group's last ND lanes
    uint64_t* hbar = reinterpret_cast<uint64_t*>(halo + kTriGroups * ND * p.D);   // [kTriGroups][ND][2]: steps alternate

    auto block_size = [&](int lane) -> int {
        if (lane < 0 || lane >= p.N) return 0;
        const i64 K = p.upper ? p.N - 1 - lane : lane;
        return (int)(p.R[K + 1] - p.R[K]);
    };
    const int rK = block_size(g);
    for (int t = tid; t < p.D * kTriThreads; t += kTriThreads) ring[t] = T(0);
    for (int t = tid; t < kTriGroups * ND * p.D; t += kTriThreads) halo[t] = T(0);
    if (tid < kTriGroups * ND * 2) mbar_init(&hbar[tid], 1);
    fence_mbar_init();
    cluster.sync();  // rings zeroed and mbarriers initialised before a neighbour CTA may touch them

    // ring of the lane d block rows back and its slot stride; a lane without such a neighbour reads its own ring
    // (times a zero coefficient)
    const T* rem[ND];
    int rstride[ND];
#pragma unroll
    for (int d = 1; d <= ND; ++d) {
        const int gj = g - d;
        rstride[d - 1] = kTriThreads;
        if (gj < 0) {
            rem[d - 1] = ring + tid;
        } else if (lid >= d) {
            rem[d - 1] = ring + tid - d;
        } else if (push) {
            rem[d - 1] = halo + (grp * ND + (d - lid - 1)) * p.D;
            rstride[d - 1] = 1;
        } else {
            const int Gj = gj / kTriGroup;
            rem[d - 1] = cluster.map_shared_rank(ring, Gj % kTriCtas) + (Gj / kTriCtas) * kTriGroup + gj % kTriGroup;
        }
    }
    T* const mine = ring + tid;
    T* const cmine = coef + tid;
    // step s works on in-block row k = kbase + ksgn*kr, kr = s - delta*g - kk0, if 0 <= kr < rK
    const int kk0 = p.upper ? p.maxr - rK : 0;
    const int kbase = p.upper ? rK - 1 : 0, ksgn = p.upper ? -1 : 1;
    int kr = -delta * g - kk0;
    T* src = p.packed + (g % Wd);   // + s*Wd

    // producer side of the halo: my values go to halo ring q of the next lane group
    const int q_out = kTriGroup - 1 - lid;
    const bool pushes = push && q_out < ND && G + 1 < kTriGroups * kTriCtas;
    uint32_t out_ring = 0, out_bar = 0;
    if (pushes) {
        const int Gn = G + 1;
        out_ring = tri_mapa(smem_u32(halo + ((Gn / kTriCtas) * ND + q_out) * p.D), Gn % kTriCtas);
        out_bar = tri_mapa(smem_u32(&hbar[((Gn / kTriCtas) * ND + q_out) * 2]), Gn % kTriCtas);
    }
    // consumer side (first warp of the group): halo ring q receives a value in step t iff its producer is active
    // then.  That value completes a transaction on mbarrier [q][t&1], which the consumer armed in step t-1 (right
    // after it was done with the value of step t-2 on the same mbarrier, and before its own arrive, hence before
    // the producer can get to step t) and waits for in step t+1.
    const bool halo_warp = push && lid < 32 && G > 0;
    int krP[ND], rKP[ND];
    uint32_t par = 0, failed = 0;   // bit 2q+b of par = phase parity of mbarrier [q][b]
#pragma unroll
    for (int q = 0; q < ND; ++q) {
        const int gP = G * kTriGroup - 1 - q;
        rKP[q] = halo_warp ? block_size(gP) : 0;
        krP[q] = -1 - delta * gP - (p.upper ? p.maxr - rKP[q] : 0);   // the producer's kr at step s-1
        if (halo_warp && lid == 0 && (unsigned)(krP[q] + 1) < (unsigned)rKP[q]) mbar_arrive_expect_tx(&hbar[(grp * ND + q) * 2], (uint32_t)sizeof(T));
    }

    auto issue = [&](int st, int krs, const T* sp) {
        if ((unsigned)krs < (unsigned)rK) {
            T* cb = cmine + (size_t)((st & (PF - 1)) * NC) * kTriThreads;
            tri_cp_async(cb, sp + xq);
            if (!unit) tri_cp_async(cb + kTriThreads, sp + tot);
#pragma unroll
            for (int c = 2; c < NC; ++c) tri_cp_async(cb + c * kTriThreads, sp + c * tot);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int t = 0; t < PF - 1; ++t) issue(t, kr + t, src + (i64)t * Wd);

    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    for (int s = 0; s < iters; ++s) {
        issue(s + PF - 1, kr + PF - 1, src + (PF - 1) * Wd);
        asm volatile("cp.async.wait_group %0;" ::"n"(PF - 1) : "memory");
        const bool active = (unsigned)kr < (unsigned)rK;
        const int k = kbase + ksgn * kr;
        T acc = T(0), rinv = T(1), dg = T(1), co[ND][W], cl[NDIAG > 0 ? NDIAG : 1];
        const T* pr[ND][W];
        const T* pl[NDIAG > 0 ? NDIAG : 1];
        T* const pme = mine + (k & Dm) * kTriThreads;
        if (active) {
            const T* cb = cmine + (size_t)((s & (PF - 1)) * NC) * kTriThreads;
            acc = cb[0];
            if (!unit) { dg = cb[kTriThreads]; rinv = T(1) / dg; }
            const int km = k + mu;
#pragma unroll
            for (int d = 0; d < ND; ++d)
#pragma unroll
                for (int ii = 0; ii < W; ++ii) {
                    co[d][ii] = cb[(2 + d * W + ii) * kTriThreads];
                    pr[d][ii] = rem[d] + ((km - ii) & Dm) * rstride[d];
                }
#pragma unroll
            for (int t = 0; t < NDIAG; ++t) {
                cl[t] = cb[(2 + ND * W + t) * kTriThreads];
                pl[t] = mine + ((km - ii0 - t) & Dm) * kTriThreads;
            }
        }
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // step s-1 is complete cluster-wide
        if (halo_warp) {
#pragma unroll
            for (int q = 0; q < ND; ++q) {
                uint64_t* bar = &hbar[(grp * ND + q) * 2 + ((s + 1) & 1)];   // the mbarrier of steps s-1 and s+1
                const int bit = 2 * q + ((s + 1) & 1);
                if ((unsigned)krP[q] < (unsigned)rKP[q] && !failed) {          // a value was pushed in step s-1
                    int spin = 0;
                    while (!mbar_try_wait(bar, (par >> bit) & 1u))
                        if (++spin > (1 << 14)) { failed = 1; break; }   // never hang: a lost value poisons the result
                    par ^= 1u << bit;
                }
                if (lid == 0 && (unsigned)(krP[q] + 2) < (unsigned)rKP[q]) mbar_arrive_expect_tx(bar, (uint32_t)sizeof(T));
            }
        }
        T xk = T(0);
        if (active) {
#pragma unroll
            for (int d = 0; d < ND; ++d)
#pragma unroll
                for (int ii = 0; ii < W; ++ii) acc = fma(-co[d][ii], *pr[d][ii], acc);
#pragma unroll
            for (int t = 0; t < NDIAG; ++t) acc = fma(-cl[t], *pl[t], acc);
            xk = acc;
            if (!unit) {   // acc / dg from the reciprocal prepared above (quotient, exact residual, correction)
                const T q = acc * rinv;
                xk = fma(fma(-dg, q, acc), rinv, q);
            }
            *pme = xk;
            if (pushes) tri_st_async(out_ring + (uint32_t)((k & Dm) * sizeof(T)), xk, out_bar + 8u * (s & 1));
        }
        // only the shared-memory ring has to be ordered across the step boundary (see the generic kernel)
        asm volatile("fence.acq_rel.cta;" ::: "memory");
        asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
        if (active) src[xq] = xk;
        ++kr;
#pragma unroll
        for (int q = 0; q < ND; ++q) ++krP[q];
        src += Wd;
    }
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (failed) p.status[rhs] = 1;
    cluster.sync();  // nothing may leave while a neighbour could still read its ring
}

template <typename T, int PF, int ND, int W, int NDIAG>
cudaError_t launch_lean(const TriWaveParams<T>& p, size_t smem, cudaStream_t stream, int nq) {
    cudaError_t e = cudaFuncSetAttribute(bbb_trisolve_lean_kernel<T, PF, ND, W, NDIAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    bbb_trisolve_lean_kernel<T, PF, ND, W, NDIAG><<<kTriCtas * nq, kTriThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

template <typename T, int PF>
cudaError_t launch_wave(const TriWaveParams<T>& p, size_t smem, cudaStream_t stream, int nq) {
    cudaError_t e = cudaFuncSetAttribute(bbb_trisolve_wave_kernel<T, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    bbb_trisolve_wave_kernel<T, PF><<<kTriCtas * nq, kTriThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

// mode 0: pack the matrix planes into the handle's workspace, then solve (the one-call path)
// mode 1: only pack the matrix planes into a buffer of the caller's (*ext, allocated here): bbm_bbb_trsv_prepare_*
// mode 2: solve with the matrix planes already in *ext: bbm_bbb_trsv_solve_*  (a Gauss-Seidel sweep re-uses them: the matrix
//         does not change between sweeps, and packing it is 1.2 GB written + read per solve at cfg2's size)
// Right-hand sides are packed, solved and unpacked in batches of up to "tri.rhs_batch" (default 16): one 8-CTA cluster
// each, concurrently on different SMs.  Returns BBM_OK when done, 1 when this variant does not apply (caller falls back).
template <typename T>
int32_t bbb_trsv_wave(bbm_handle_t h, const BbbDescView& v, const TriSolveParams<T>& sp, T* d_B, i64 ldb, i64 nrhs, int mode = 0,
                      void** ext = nullptr, size_t* ext_bytes = nullptr) {
    const int upper = sp.upper;
    const i64 delta = upper ? v.lam + 1 : v.mu + 1;
    const i64 nd = upper ? v.u : v.l;
    if (sp.iters + 8 >= (i64)1 << 31 || delta * (sp.N + (i64)kTriLanes) >= (i64)1 << 31) return 1;   // 32-bit step counters
    TriWaveParams<T> p{};
    p.data = sp.data; p.R = sp.R; p.N = (int)sp.N; p.maxr = (int)sp.maxr; p.iters = (int)sp.iters;
    p.l = sp.l; p.u = sp.u; p.lam = sp.lam; p.mu = sp.mu; p.w = sp.w; p.P = sp.P;
    p.upper = upper; p.unit = sp.unit; p.passes = sp.passes; p.D = sp.D; p.nc = sp.nc; p.relaxed = sp.relaxed;
    const size_t budget = std::min<size_t>(h->smem_optin, 200 * 1024);
    const size_t fixed = (size_t)p.passes * (kTriThreads * p.D * sizeof(T) + (kTriGroups * (kTriGroup + nd) + kTriThreads) * sizeof(int)) + 16;
    int pf = 0;
    const int want_pf = (int)bbm_opt(h, "tri.pf", 4);
    for (int cand : {8, 4, 2})
        if (cand <= want_pf && pf == 0 && fixed + (size_t)p.passes * cand * kTriThreads * p.nc * sizeof(T) <= budget) pf = cand;
    if (pf == 0) return 1;
    // the block rows active in one step are fewer than Wd consecutive ones
    i64 Wd = std::min<i64>(sp.N, sp.maxr / delta + 1);
    Wd = (Wd + 15) / 16 * 16;
    p.Wd = Wd;
    p.total = (i64)p.iters * Wd;
    const size_t plane = ((size_t)p.total * sizeof(T) + 255) / 256 * 256;
    const size_t need = plane * (size_t)(p.nc - 1);                  // matrix planes 1 .. nc-1 (plane 0 is the right-hand side's)
    const int batch = (int)std::max<i64>(1, std::min<i64>(std::min<i64>(nrhs, bbm_opt(h, "tri.rhs_batch", 16)), 64));
    const size_t xneed = 256 + plane * (size_t)batch;                // status words + right-hand-side planes
    if (need + xneed > (size_t)bbm_opt(h, "tri.ws_max_mb", 16384) * 1048576) return 1;
    T* mat = nullptr;
    char* xws = nullptr;
    if (mode == 1) {
        cudaError_t e = cudaMalloc(ext, need + 256);
        if (e != cudaSuccess) { cudaGetLastError(); return bbm_fail(h, BBM_ERR_ALLOC, "cudaMalloc(%zu) for the wave-ordered matrix planes: %s", need, cudaGetErrorString(e)); }
        *ext_bytes = need;
        mat = static_cast<T*>(*ext);
    } else if (mode == 2) {
        if (!ext || !*ext || *ext_bytes != need) return bbm_fail(h, BBM_ERR_ARG, "prepared plan does not match this solve");
        if (bbm_ws_reserve(h, &h->ws_w, &h->ws_w_bytes, xneed + 256) != BBM_OK) return 1;
        mat = static_cast<T*>(*ext);
        xws = static_cast<char*>(h->ws_w);
    } else {
        if (bbm_ws_reserve(h, &h->ws_w, &h->ws_w_bytes, need + xneed + 256) != BBM_OK) return 1;
        mat = static_cast<T*>(h->ws_w);
        xws = static_cast<char*>(h->ws_w) + need;
    }
    p.packed = mat - plane / sizeof(T);          // plane c of the matrix is packed + c*total', c >= 1 (plane 0 is never addressed)
    p.total = (i64)(plane / sizeof(T));          // plane stride in elements (rounded to 256 bytes)
    p.push = bbm_opt(h, "tri.push", 1) != 0 ? 1 : 0;
    p.ldb = ldb;
    // transposing tiles of 32 block rows x 32 steps; a block-row tile is active for maxr + 31*delta steps
    const dim3 grid((unsigned)((p.maxr + 31 * delta + 31) / 32), (unsigned)((p.N + 31) / 32));
    const bool fused_pack = mode == 0 && nrhs == 1;   // one launch packs the matrix planes and the right-hand side together
    if (mode != 2 && !fused_pack) {
        p.x = nullptr; p.what = 2; p.xoff = 0; p.status = nullptr;
        bbb_tri_pack_kernel<T><<<grid, 256, 0, h->stream>>>(p);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "bbb_tri_pack_kernel launch: %s", cudaGetErrorString(e));
        h->launches += 1;
        if (mode == 1) return BBM_OK;
    }
    p.status = reinterpret_cast<int*>(xws);
    p.xoff = (reinterpret_cast<T*>(xws + 256) - p.packed);
    const size_t smem = fixed + (size_t)p.passes * pf * kTriThreads * p.nc * sizeof(T);
    // lean instantiations: one pass, the two common shapes (l or u = 1 with 3 sub-block bands, 2 with 5)
    int lean = 0;
    if (p.passes == 1 && bbm_opt(h, "tri.lean", 1) != 0 &&
        (size_t)kTriThreads * (p.D + 4 * p.nc) * sizeof(T) + (size_t)kTriGroups * nd * (p.D * sizeof(T) + 2 * sizeof(uint64_t)) <= budget) {
        const int ndiag = upper ? p.mu : p.lam;
        if (nd == 1 && p.w == 3 && ndiag == 1) lean = 13;
        if (nd == 2 && p.w == 5 && ndiag == 2) lean = 25;
    }
    for (i64 q0 = 0; q0 < nrhs; q0 += batch) {
        const int nq = (int)std::min<i64>(batch, nrhs - q0);
        p.x = d_B + q0 * ldb;
        p.what = fused_pack ? 3 : 1;
        BBM_CUDA(h, cudaMemsetAsync(p.status, 0, sizeof(int) * 64, h->stream));
        bbb_tri_pack_kernel<T><<<dim3(grid.x, grid.y, (unsigned)nq), 256, 0, h->stream>>>(p);
        cudaError_t e;
        const size_t lsmem = (size_t)kTriThreads * (p.D + 4 * p.nc) * sizeof(T) + (size_t)kTriGroups * nd * (p.D * sizeof(T) + 2 * sizeof(uint64_t));
        if (lean == 13) e = launch_lean<T, 4, 1, 3, 1>(p, lsmem, h->stream, nq);
        else if (lean == 25) e = launch_lean<T, 4, 2, 5, 2>(p, lsmem, h->stream, nq);
        else e = pf == 8 ? launch_wave<T, 8>(p, smem, h->stream, nq) : pf == 4 ? launch_wave<T, 4>(p, smem, h->stream, nq) : launch_wave<T, 2>(p, smem, h->stream, nq);
        if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "bbb_trisolve_wave_kernel launch: %s", cudaGetErrorString(e));
        bbb_tri_unpack_kernel<T><<<dim3(grid.x, grid.y, (unsigned)nq), 256, 0, h->stream>>>(p);
        e = cudaGetLastError();
        if (e != cudaSuccess) return bbm_fail(h, BBM_ERR_CUDA, "bbb_tri_unpack_kernel launch: %s", cudaGetErrorString(e));
        h->launches += 3;
    }
    return BBM_OK;
}

template <typename T>
int32_t bbb_trsv_impl(bbm_handle_t h, bbm_bbb_desc_t d, int32_t uplo, int32_t unit, const T* d_data, T* d_B, i64 ldb, i64 nrhs, int mode = 0,
                      void** ext = nullptr, size_t* ext_bytes = nullptr) {
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
    if (v.m == 0 || (nrhs == 0 && mode != 1)) return BBM_OK;
    if (v.P == 0 || v.l < 0 || v.u < 0 || v.lam < 0 || v.mu < 0) return bbm_fail(h, BBM_ERR_BAND, "BandError: the diagonal is outside the bands");
    if ((mode != 2 && !d_data) || (mode != 1 && !d_B)) return bbm_fail(h, BBM_ERR_ARG, "null device pointer");
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
    p.relaxed = bbm_opt(h, "tri.sync", 1) != 0 ? 1 : 0;  // 0 = cluster.sync(), 1 = CTA fence + relaxed arrive (staged / generic wave kernels)
    const int which = (int)bbm_opt(h, "tri.kernel", 0);  // 0 = wave-ordered, 1 = staged, 2 = direct
    if (which == 0 || mode != 0) {
        const int32_t rc = bbb_trsv_wave<T>(h, v, p, d_B, ldb, nrhs, mode, ext, ext_bytes);
        if (rc != 1) return rc;
        if (mode != 0) return bbm_fail(h, BBM_ERR_UNSUPPORTED, "the wave-ordered solve does not apply to this structure: use bbm_bbb_trsv_* directly");
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

struct bbm_bbb_tri_plan {
    uint32_t magic = 0x7B1A50F1u;
    bbm_handle_t h = nullptr;
    bbm_bbb_desc_t d = nullptr;
    int uplo = 'L', unit = 0, elem = 8;
    void* packed = nullptr;      // wave-ordered matrix planes
    size_t bytes = 0;
};

namespace {
template <typename T>
int32_t tri_prepare(bbm_handle_t h, bbm_bbb_desc_t d, int32_t uplo, int32_t unit, const T* d_data, bbm_bbb_tri_plan_t* out) {
    BBM_REQUIRE_HANDLE(h);
    if (!out) return bbm_fail(h, BBM_ERR_ARG, "null out pointer");
    *out = nullptr;
    bbm_bbb_tri_plan* pl = new (std::nothrow) bbm_bbb_tri_plan();
    if (!pl) return bbm_fail(h, BBM_ERR_ALLOC, "out of host memory");
    pl->h = h; pl->d = d; pl->uplo = uplo; pl->unit = unit ? 1 : 0; pl->elem = (int)sizeof(T);
    int32_t rc = bbb_trsv_impl<T>(h, d, uplo, unit, d_data, nullptr, 1, 0, 1, &pl->packed, &pl->bytes);
    if (rc != BBM_OK) { if (pl->packed) cudaFree(pl->packed); delete pl; return rc; }
    *out = pl;
    return BBM_OK;
}
template <typename T>
int32_t tri_solve(bbm_bbb_tri_plan_t pl, T* d_B, i64 ldb, i64 nrhs) {
    if (!pl || pl->magic != 0x7B1A50F1u || !bbm_valid(pl->h)) return BBM_ERR_HANDLE;
    if (pl->elem != (int)sizeof(T)) return bbm_fail(pl->h, BBM_ERR_ARG, "the plan was prepared for another element type");
    if (!pl->packed) return BBM_OK;   // empty matrix
    return bbb_trsv_impl<T>(pl->h, pl->d, pl->uplo, pl->unit, nullptr, d_B, ldb, nrhs, 2, &pl->packed, &pl->bytes);
}
}  // namespace

extern "C" {
int32_t bbm_bbb_trsv_prepare_f64(bbm_handle_t h, bbm_bbb_desc_t d, int32_t uplo, int32_t unit, const double* d_data, bbm_bbb_tri_plan_t* out) {
    return tri_prepare<double>(h, d, uplo, unit, d_data, out);
}
int32_t bbm_bbb_trsv_prepare_f32(bbm_handle_t h, bbm_bbb_desc_t d, int32_t uplo, int32_t unit, const float* d_data, bbm_bbb_tri_plan_t* out) {
    return tri_prepare<float>(h, d, uplo, unit, d_data, out);
}
int32_t bbm_bbb_trsv_solve_f64(bbm_bbb_tri_plan_t p, double* d_B, int64_t ldb, int64_t nrhs) { return tri_solve<double>(p, d_B, ldb, nrhs); }
int32_t bbm_bbb_trsv_solve_f32(bbm_bbb_tri_plan_t p, float* d_B, int64_t ldb, int64_t nrhs) { return tri_solve<float>(p, d_B, ldb, nrhs); }
int32_t bbm_bbb_tri_plan_destroy(bbm_bbb_tri_plan_t p) {
    if (!p || p->magic != 0x7B1A50F1u) return BBM_ERR_HANDLE;
    if (bbm_valid(p->h)) cudaSetDevice(p->h->device);
    if (p->packed) cudaFree(p->packed);
    p->magic = 0;
    delete p;
    return BBM_OK;
}
int32_t bbm_bbb_trsv_f64(bbm_handle_t h, bbm_bbb_desc_t d, int32_t uplo, int32_t unit, const double* d_data, double* d_B, int64_t ldb,
                         int64_t nrhs) {
    return bbb_trsv_impl<double>(h, d, uplo, unit, d_data, d_B, ldb, nrhs);
}
int32_t bbm_bbb_trsv_f32(bbm_handle_t h, bbm_bbb_desc_t d, int32_t uplo, int32_t unit, const float* d_data, float* d_B, int64_t ldb,
                         int64_t nrhs) {
    return bbb_trsv_impl<float>(h, d, uplo, unit, d_data, d_B, ldb, nrhs);
}
}
